#!/usr/bin/env python
"""Summarise an `ncu --page source --csv --print-source sass` dump: opcode mix (executed / sampled) and the hottest
instructions with their top stall reasons.   python tools/ncu_source_summary.py dump.csv [n_top]"""
import collections
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[h]
    data = [r for r in rows[h + 1:] if len(r) == len(hdr) and r[0] != "Address"]
    ix = {k: i for i, k in enumerate(hdr)}

    def num(r, k):
        try:
            return int(float(r[ix[k]] or 0))
        except ValueError:
            return 0
    tot = sum(num(r, "# Samples") for r in data)
    texec = sum(num(r, "Instructions Executed") for r in data)
    print("samples", tot, "warp-instructions", texec, "sass lines", len(data))
    ops, samp = collections.Counter(), collections.Counter()
    for r in data:
        toks = r[ix["Source"]].strip().split()
        if not toks:
            continue
        op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
        op = ".".join(op.split(".")[:2]) if op.startswith(("FFMA", "HFMA", "FHFMA", "LDS", "STS", "LDG", "STG")) else op.split(".")[0]
        ops[op] += num(r, "Instructions Executed")
        samp[op] += num(r, "# Samples")
    for op, c in ops.most_common(28):
        print("%-16s exec %5.1f%%  samples %5.1f%%" % (op, 100 * c / max(texec, 1), 100 * samp[op] / max(tot, 1)))
    stalls = [k for k in hdr if k.startswith("stall_") and "(Not" not in k]
    agg = collections.Counter()
    for r in data:
        for k in stalls:
            agg[k] += num(r, k)
    print("stall totals:", [(k, v) for k, v in agg.most_common(8)])
    for r in sorted(data, key=lambda r: -num(r, "# Samples"))[:ntop]:
        st = sorted(((k, num(r, k)) for k in stalls), key=lambda kv: -kv[1])[:3]
        print(r[ix["Address"]][-5:], "%6d %8d" % (num(r, "# Samples"), num(r, "Instructions Executed")), r[ix["Source"]][:72], st)


if __name__ == "__main__":
    main()

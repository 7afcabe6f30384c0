#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv --log-file X` launch list of
`bench.py --steps 1 --warmup 1 --kernels-only`: the LAST training step (from its preprocess_u8 launch on), kernels grouped by
name with their share of the summed durations.   python tools/launch_list_summary.py launches.csv [title]"""
import collections
import csv
import re
import sys


def main():
    rows = []
    with open(sys.argv[1], newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ix = {k: i for i, k in enumerate(hdr)}
    for r in rd:
        if len(r) != len(hdr) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        v = float(r[ix["Metric Value"]].replace(",", ""))
        unit = r[ix["Metric Unit"]]
        us = v / 1e3 if unit in ("ns", "nsecond") else (v * 1e3 if unit in ("ms", "msecond") else v)
        rows.append((r[ix["Kernel Name"]], us))
    starts = [i for i, (n, _) in enumerate(rows) if "preprocess_u8" in n]
    step = rows[starts[-1]:] if starts else rows
    total = sum(us for _, us in step)
    groups = collections.defaultdict(lambda: [0.0, 0])
    for n, us in step:
        n = re.sub(r"^void ", "", n)
        n = re.sub(r"\(.*$", "", n)
        n = re.sub(r"mvfb::<unnamed>::|<unnamed>::|unnamed>::", "", n)
        groups[n[:110]][0] += us
        groups[n[:110]][1] += 1
    if len(sys.argv) > 2:
        print("# " + sys.argv[2])
    print("# the timed step only (from preprocess_u8 of the last step to the end); per-launch times are cold-cache and serialised: compare SHARES, not absolutes")
    print("step total %.1f us, %d launches" % (total, len(step)))
    own = sum(v[0] for k, v in groups.items() if not k.startswith(("at::", "cutlass", "cudnn", "nccl", "void at::")) and "cutlass" not in k and "cudnn" not in k)
    print("own kernels: %.1f %% of the summed time" % (100 * own / total))
    for k, (us, n) in sorted(groups.items(), key=lambda kv: -kv[1][0]):
        print("%5.1f%% %10.1f us %4d  %s" % (100 * us / total, us, n, k))


if __name__ == "__main__":
    main()

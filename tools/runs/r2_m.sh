#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== mvf tests (default build)"; timeout 900 python -m pytest tests/test_mvf_gpu.py tests/test_mvf_production_gpu.py -m gpu -q -x > gpurun_out/m_mvf.log 2>&1; echo "rc=$?" >> gpurun_out/m_mvf.log; tail -15 gpurun_out/m_mvf.log | cut -c1-300
for v in default mid2 fp32stats centrelo; do
  echo "== variant $v"
  if [ $v = default ]; then L=""; else L="--lib variants/$v/libmvf_b200.so"; fi
  timeout 600 python tools/mvf_microbench.py --iters 20 --clips 160 --fwd-only $L --out gpurun_out/m_micro_$v.jsonl 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.rstrip()[:200]); continue
    print(d['C'], d['H'], 'train' if d['training'] else 'eval ', '%.1f us  frac %.3f' % (d['us_median'], d['frac']))
"
done

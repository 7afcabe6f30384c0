#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for v in fp32stats centrelo mid2 mid2exact; do
  echo "== variant $v"
  MVFB_TEST_LIB=variants/$v/libmvf_b200.so timeout 900 python -m pytest tests/test_mvf_gpu.py tests/test_mvf_production_gpu.py -m gpu -q 2>&1 | grep -E "^FAILED|passed|failed|AssertionError: assert" | cut -c1-200 | head -12
done
echo "== mid2exact timing"
timeout 600 python tools/mvf_microbench.py --iters 20 --clips 160 --fwd-only --lib variants/mid2exact/libmvf_b200.so --out gpurun_out/n_micro.jsonl 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.rstrip()[:200]); continue
    print(d['C'], d['H'], 'train' if d['training'] else 'eval ', '%.1f us  frac %.3f' % (d['us_median'], d['frac']))
"

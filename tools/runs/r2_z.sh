#!/bin/bash
# final artefacts of the round: smoke, full GPU suite, default bench, reference arm, per-shape list, ncu launch list
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
echo "== pytest gpu (all)"; timeout 1500 python -m pytest tests -m gpu -q -x --durations=3 > gpurun_out/z_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/z_pytest.log; tail -8 gpurun_out/z_pytest.log | cut -c1-300
echo "== bench (default flags)"; timeout 1500 python bench.py > gpurun_out/z_bench.json 2> gpurun_out/z_bench.err; echo "rc=$?"; tail -c 600 gpurun_out/z_bench.json; tail -3 gpurun_out/z_bench.err
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/z_ref.json 2> gpurun_out/z_ref.err; tail -c 400 gpurun_out/z_ref.json
echo "== by shape"; timeout 500 python tools/step_by_shape.py --families gemm1x1,conv3x3,wgrad1x1,wgrad3x3,bn_fwd,bn_bwd,maxpool,stem_im2col,mvf_fwd,mvf_bwd --out gpurun_out/z_by_shape.json > gpurun_out/z_by_shape.txt 2>&1; tail -3 gpurun_out/z_by_shape.txt
echo "== launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/z_launches.csv python bench.py --steps 1 --warmup 1 --batch 160 --kernels-only > gpurun_out/z_ncu_list.log 2>&1; tail -1 gpurun_out/z_ncu_list.log
python tools/launch_list_summary.py gpurun_out/z_launches.csv "round 2 final build, B=160 clips, uint8 input: ncu --metrics gpu__time_duration.sum --clock-control none python bench.py --steps 1 --warmup 1 --batch 160 --kernels-only" > gpurun_out/z_launches.txt; head -12 gpurun_out/z_launches.txt

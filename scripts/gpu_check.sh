#!/usr/bin/env bash
# One gpurun call: GPU tests, smoke, bench (both arms), ncu launch list + full capture.
# usage: scripts/gpu_check.sh <tag>   (outputs under gpurun_out/<tag>_*)
set -u
tag=${1:-r01}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/${tag}_gpu.txt 2>&1
python -m pytest tests -x -q -m gpu > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${tag}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/${tag}_smoke.log
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err; echo "bench ref rc=$?"
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/${tag}_bench.json
SMALL="python bench.py --steps 6 --warmup 3 --no-cpu --dataset-rows 131072 --big-batch 65536 --decode-rows 262144"
$SMALL > gpurun_out/${tag}_small.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv $SMALL > gpurun_out/${tag}_ncu1.log 2>&1
echo "ncu list rc=$?"
$SMALL > gpurun_out/${tag}_small2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:train_tc_fused_kernel -s 5 -c 1 -f -o gpurun_out/${tag}_prof_fused $SMALL > gpurun_out/${tag}_ncu2.log 2>&1
echo "ncu full fused rc=$?"
ncu --set full --clock-control none --import-source on -k regex:wgrad_kernel -s 5 -c 1 -f -o gpurun_out/${tag}_prof_wgrad $SMALL > gpurun_out/${tag}_ncu3.log 2>&1
echo "ncu full wgrad rc=$?"
ncu --set full --clock-control none --import-source on -k regex:decode_tc_kernel -s 4 -c 1 -f -o gpurun_out/${tag}_prof_decode_tc $SMALL > gpurun_out/${tag}_ncu4.log 2>&1
echo "ncu full decode_tc rc=$?"
ls -la gpurun_out

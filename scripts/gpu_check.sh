#!/usr/bin/env bash
# One gpurun call: GPU tests, smoke, bench (both arms), ncu launch list + full captures.
# usage: scripts/gpu_check.sh <tag>   (outputs under gpurun_out/<tag>_*)
set -u
tag=${1:-r02}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/${tag}_gpu.txt 2>&1
python -m pytest tests -x -q -m gpu > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${tag}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/${tag}_smoke.log
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err; echo "bench ref rc=$?"
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench_driver_flags.json 2> gpurun_out/${tag}_bench_driver_flags.err; echo "bench (driver flags) rc=$?"
tail -c 1500 gpurun_out/${tag}_bench.json
python scripts/trace_chain.py 4096 > gpurun_out/${tag}_fused_trace.txt 2>&1
SMALL="python bench.py --steps 6 --warmup 3 --no-cpu --dataset-rows 131072 --big-batch 65536 --decode-rows 1048576 --mpc-rows 65536"
$SMALL > gpurun_out/${tag}_small.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches.csv $SMALL > gpurun_out/${tag}_ncu1.log 2>&1
echo "ncu list rc=$?"
for k in train_tc_fused_kernel:5 chain_kernel:2 wgrad_kernel:2 decode_tc_kernel:4 reduce_tc_kernel:5 train_kernel:2 mpc_track_kernel:1; do
  name=${k%%:*}; skip=${k##*:}
  ncu --set full --clock-control none --import-source on -k regex:$name -s $skip -c 1 -f -o gpurun_out/${tag}_prof_${name} $SMALL > gpurun_out/${tag}_ncu_${name}.log 2>&1
  echo "ncu full $name rc=$?"
done
timeout 900 python scripts/sweep.py --out gpurun_out/${tag}_sweep.jsonl > gpurun_out/${tag}_sweep.txt 2>&1; echo "sweep rc=$?"
python scripts/quick_mpc_bench.py 1048576 20 > gpurun_out/${tag}_mpc_1m.txt 2>&1
ls -la gpurun_out | tail -20

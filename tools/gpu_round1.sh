#!/bin/bash
# First GPU bring-up: probe, parity tests per file (each under its own timeout), then the bench.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo skip-probe-sweep >> gpurun_out/summary.txt
for f in test_gpu_tc_probe test_gpu_preprocess test_gpu_forward test_gpu_train; do
  timeout 900 python -m pytest tests/$f.py -q -m gpu --tb=line > gpurun_out/$f.log 2>&1; echo "$f exit $?" >> gpurun_out/summary.txt
  tail -3 gpurun_out/$f.log >> gpurun_out/summary.txt
done
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/summary.txt
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt

#!/bin/bash
# round 2, run A (1 GPU): all parity tests, phase stamps, bench
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests -q -m gpu --tb=short -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest -m gpu exit $?" >> gpurun_out/summary.txt
tail -15 gpurun_out/pytest_gpu.log >> gpurun_out/summary.txt
timeout 300 python tools/phase_timing.py > gpurun_out/phase_timing.log 2>&1; echo "phase exit $?" >> gpurun_out/summary.txt
cat gpurun_out/phase_timing.log >> gpurun_out/summary.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/summary.txt
tail -5 gpurun_out/bench.err >> gpurun_out/summary.txt
python - <<'PY' >> gpurun_out/summary.txt 2>&1
import json
d = json.load(open('gpurun_out/bench.json'))
print({k: d[k] for k in ('value', 'ms_per_step', 'repeats', 'timed_region_ms', 'clocks', 'tc_status', 'status_ok', 'gpu_launches')})
print('e2e', d['e2e'])
print('roofline', {k: d['roofline'][k] for k in ('achieved', 'frac', 'kernel_ms', 'kernel_ms_fwd_bwd_only')})
for k in ('fwd', 'train_fp32', 'fwd_fp32', 'train_ref_default_shape', 'train_ref_default_shape_fp32', 'fwd_config1_latency', 'fwd_config1_latency_fp32', 'preprocess', 'stream', 'launch_floor_us'):
    print(k, json.dumps(d.get(k))[:400])
PY
cat gpurun_out/summary.txt

#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_train.py -q -m gpu --tb=short -x > gpurun_out/test_gpu_train.log 2>&1; echo "train tests exit $?" >> gpurun_out/summary.txt
tail -4 gpurun_out/test_gpu_train.log >> gpurun_out/summary.txt
timeout 300 python tools/phase_timing.py > gpurun_out/phase_timing.log 2>&1; cat gpurun_out/phase_timing.log >> gpurun_out/summary.txt
timeout 600 python bench.py --skip-extras > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(d['value']/1e6, 'Mf/s', d['ms_per_step']*1e3, 'us', 'launches', d['gpu_launches'], 'kernel', d['roofline']['kernel_ms']*1e3, 'e2e', d['e2e']['value']/1e6)"

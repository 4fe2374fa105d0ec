#!/bin/bash
# N=1: programmatic dependent launch on/off: phase stamps, train/forward tests, quick bench
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
S=gpurun_out/summary.txt
for M in 1 0; do
  B2H_PDL=$M timeout 200 python tools/phase_timing.py > gpurun_out/phase_pdl$M.log 2>&1; echo "phase pdl=$M exit $?" >> $S
  grep -E "us/step" gpurun_out/phase_pdl$M.log >> $S
  grep -E "rep 2" -A1 gpurun_out/phase_pdl$M.log | grep -v "^\[r[1-7]" | tail -3 >> $S
done
B2H_PDL=1 B2H_NONCOOP=1 timeout 200 python tools/phase_timing.py > gpurun_out/phase_pdl_noncoop.log 2>&1; echo "phase pdl noncoop exit $?" >> $S
grep -E "us/step" gpurun_out/phase_pdl_noncoop.log >> $S
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_forward.py -q -m gpu --tb=short -x > gpurun_out/pytest_train.log 2>&1; echo "train+fwd tests exit $?" >> $S
tail -5 gpurun_out/pytest_train.log >> $S
for M in 1 0; do
  B2H_PDL=$M timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_pdl$M.json 2> gpurun_out/bench_pdl$M.err; echo "bench pdl=$M exit $?" >> $S
done
python - <<PY >> $S 2>&1
import json
for f in ('bench_pdl1', 'bench_pdl0'):
    try:
        d = json.load(open('gpurun_out/%s.json' % f))
        print(f, 'us/step', d['ms_per_step'] * 1e3, 'e2e us', d['e2e']['us_per_step'], 'fwd', d['fwd'].get('ms_per_batch'), 'fp32', d['train_fp32'].get('ms_per_step'), 'fwd_fp32', d['fwd_fp32'].get('ms_per_batch'), 'T200', d['train_ref_default_shape'].get('ms_per_step'), 'lat', d['fwd_config1_latency'].get('ms_per_batch'), 'stream16', d['stream']['stride16'].get('unique_frames_per_sec'), d.get('tc_status'), d.get('status_ok'))
    except Exception as e:
        print(f, 'ERR', e)
PY
cat $S

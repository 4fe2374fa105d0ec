#!/bin/bash
# 8-GPU box: the driver's scaling sequence N = 1, 2, 4, 8 (bench.py --steps 20 --warmup 5) + DP stamps at 8
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
S=gpurun_out/summary.txt
timeout 200 python bench.py --steps 20 --warmup 5 --skip-extras > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench n1 exit $?" >> $S
for N in 2 4 8; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29530 + N)) bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench n$N exit $?" >> $S
done
B2H_MULTICAST=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 tools/phase_timing.py > gpurun_out/phase_n8_mc1.log 2>&1; echo "phase n8 exit $?" >> $S
grep -E "\[r0\]" gpurun_out/phase_n8_mc1.log | grep -E "us/step" >> $S
python - <<PY >> $S 2>&1
import json
base = None
for n in (1, 2, 4, 8):
    try:
        d = json.load(open('gpurun_out/bench_n%d.json' % n))
        base = base or d['value']
        print('N', d['n_gpus'], 'value', round(d['value'] / 1e6, 1), 'M  us/step', round(d['ms_per_step'] * 1e3, 3), 'x', round(d['value'] / base, 3), 'e2e', round(d['e2e']['value'] / 1e6, 1), 'M e2e us', d['e2e']['segment_us_per_step'], (d.get('dp_parity') or {}).get('replicas_bit_identical'), (d.get('dp_parity') or {}).get('weights_frac_within_1e-4'), d.get('tc_status'), d.get('status_ok'))
    except Exception as e:
        print(n, 'ERR', e)
PY
cat $S

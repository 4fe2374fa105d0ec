#!/bin/bash
# N-GPU box: strong scaling (global batch 256 windows split over the ranks), N = 1, 2, 4[, 8]
NMAX=${1:-4}
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
S=gpurun_out/summary.txt
timeout 200 python bench.py --steps 20 --warmup 5 --skip-extras --strong > gpurun_out/strong_n1.json 2> gpurun_out/strong_n1.err; echo "strong n1 exit $?" >> $S
for N in 2 4 8; do
  [ $N -le $NMAX ] || continue
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29530 + N)) bench.py --gpus $N --steps 20 --warmup 5 --strong > gpurun_out/strong_n$N.json 2> gpurun_out/strong_n$N.err; echo "strong n$N exit $?" >> $S
done
python - <<PY >> $S 2>&1
import json
base = None
for n in (1, 2, 4, 8):
    try:
        d = json.load(open('gpurun_out/strong_n%d.json' % n))
    except Exception as e:
        continue
    base = base or d['value']
    print('strong N', d['n_gpus'], 'global batch', d['config']['global_batch'], 'value', round(d['value'] / 1e6, 1), 'M frames/s  us/step', round(d['ms_per_step'] * 1e3, 3), 'x', round(d['value'] / base, 3), 'e2e', round(d['e2e']['value'] / 1e6, 1), 'M', (d.get('dp_parity') or {}).get('replicas_bit_identical'), d.get('tc_status'), d.get('status_ok'))
PY
cat $S

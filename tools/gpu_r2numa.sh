#!/bin/bash
# N-GPU: host topology + bench with / without NUMA placement of the ranks' host side
N=${1:-8}
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
S=gpurun_out/summary.txt
{ nvidia-smi topo -m; lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)"; cat /sys/fs/cgroup/cpuset.cpus.effective 2>/dev/null; nproc;
  for n in /sys/devices/system/node/node*; do echo "$n $(cat $n/cpulist) $(grep MemTotal $n/meminfo)"; done; } > gpurun_out/topo.txt 2>&1
for MODE in bind nobind; do
  FLAG=""; [ $MODE = nobind ] && FLAG="--no-numa-bind"
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 $FLAG > gpurun_out/bench_n${N}_$MODE.json 2> gpurun_out/bench_n${N}_$MODE.err; echo "bench n$N $MODE exit $?" >> $S
  grep "host placement" gpurun_out/bench_n${N}_$MODE.err | head -8 >> $S
done
python - <<PY >> $S 2>&1
import json
for f in ('bench_n${N}_bind', 'bench_n${N}_nobind'):
    try:
        d = json.load(open('gpurun_out/%s.json' % f))
        print(f, d['n_gpus'], 'value', d['value'], 'us/step', d['ms_per_step'] * 1e3, 'e2e', d['e2e']['value'], 'e2e us', d['e2e']['us_per_step'], 'h2d GB/s', d['e2e']['h2d_gbs'], d.get('status_ok'))
    except Exception as e:
        print(f, 'ERR', e)
PY
tail -30 gpurun_out/topo.txt >> $S
cat $S

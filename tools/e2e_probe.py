"""Where does the end-to-end step time go?  Variants of the pipelined loop (B200, batch 256 x 64, bf16 inputs)."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hand_pose_sl_b200 as b2h
from hand_pose_sl_b200 import synthetic
from hand_pose_sl_b200.runner import TrainStepRunner, pipelined_steps
dev = torch.device("cuda:0")
torch.manual_seed(0)
m = b2h.ConvModel(30, "ReLU", False, precision="bf16").to(dev)
opt = b2h.FusedAdam(m.parameters(), lr=2e-4)
r = TrainStepRunner(m, opt, 256, 64, "L1", n_slots=2, x_dtype=torch.bfloat16)
staged = [r.host_stage(synthetic.model_batch(256, 64, seed=s)) for s in range(40)]
N = 300
def timeit(name, fn):
    fn(20); torch.cuda.synchronize()
    t0 = time.perf_counter(); fn(N); torch.cuda.synchronize()
    print(f"{name}: {(time.perf_counter() - t0) / N * 1e6:.1f} us/step", flush=True)
timeit("pipelined_steps (as benched)", lambda n: [0 for _ in pipelined_steps(r, (staged[i % 40] for i in range(n)))])
cs = torch.cuda.Stream(dev)
def copy_only(n):
    for i in range(n):
        with torch.cuda.stream(cs):
            r.load_staged(staged[i % 40], slot=i & 1)
    cs.synchronize()
timeit("copies only, back to back on a copy stream", copy_only)
def copy_sync(n):
    for i in range(n):
        with torch.cuda.stream(cs):
            r.load_staged(staged[i % 40], slot=i & 1)
        cs.synchronize()
timeit("copy + stream sync each", copy_sync)
def step_sync(n):
    ev = torch.cuda.Event()
    for i in range(n):
        l = r.step(i & 1, to_host=True); ev.record(); ev.synchronize(); float(l)
timeit("step + event sync each (no copy)", step_sync)
def step_nosync(n):
    for i in range(n):
        r.step(i & 1)
timeit("steps back to back, direct launches (no copy, no sync)", step_nosync)
def overlap_nodep(n):      # copies and steps on two streams with no dependency between them, sync per step
    ev = torch.cuda.Event()
    for i in range(n):
        with torch.cuda.stream(cs):
            r.load_staged(staged[i % 40], slot=(i + 1) & 1)
        l = r.step(i & 1, to_host=True); ev.record(); ev.synchronize(); float(l)
    cs.synchronize()
timeit("copy stream + step stream without dependencies, sync per step", overlap_nodep)
r.finish()

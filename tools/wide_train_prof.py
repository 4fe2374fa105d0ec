"""Wide training step (C = 256, B = 256 x 64): a few graph-free steps for an ncu launch list (per-kernel device times)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hand_pose_sl_b200 as b2h
from hand_pose_sl_b200 import synthetic
from hand_pose_sl_b200.runner import TrainStepRunner
C = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda:0")
torch.manual_seed(0)
m = b2h.ConvModel(C, "ReLU", False, precision="bf16").to(dev)
opt = b2h.FusedAdam(m.parameters(), lr=2e-4)
r = TrainStepRunner(m, opt, 256, 64, "L1", n_slots=2)
for s in range(2):
    r.load(synthetic.model_batch(256, 64, seed=5 + s), slot=s, non_blocking=False)
for i in range(6):
    r.step(i & 1)
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for i in range(20):
    r.step(i & 1)
ev1.record(); torch.cuda.synchronize()
r.finish()
print(f"C={C}: {ev0.elapsed_time(ev1) / 20 * 1e3:.1f} us/step (direct launches), loss {float(r.loss[0]):.5f}")

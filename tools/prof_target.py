"""Short profiling target: a few launches of every hot kernel (K0 preprocess, tcgen05 forward, fused train
step, Adam) at the BASELINE sizes.  Run plain first, then under ncu (see B200_PROFILING.md)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hand_pose_sl_b200 as b2h
from hand_pose_sl_b200 import synthetic
from hand_pose_sl_b200.runner import ForwardRunner, TrainStepRunner

dev = torch.device("cuda:0")
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
F = 108000
pose, lh, rh = synthetic.synthetic_clip(F, seed=1234)
tp, tl, tr = (torch.from_numpy(a).to(dev) for a in (pose, lh, rh))
pre = b2h.PreprocessRightHand()
starts = torch.zeros(1, dtype=torch.int64, device=dev)
for _ in range(reps):
    pre(tp, tl, tr, starts, F)
torch.manual_seed(0)
for prec in ("bf16",):
    fm = b2h.ConvModel(30, "ReLU", False, precision=prec).to(dev)
    fr = ForwardRunner(fm, 512, 64, x_dtype=torch.bfloat16 if prec == "bf16" else torch.float32)
    fr.x[0].copy_(synthetic.model_batch(512, 64, seed=99)["input_kp"])
    for _ in range(reps):
        fr.run(0)
m = b2h.ConvModel(30, "ReLU", False, precision=(sys.argv[2] if len(sys.argv) > 2 else "bf16")).to(dev)
opt = b2h.FusedAdam(m.parameters(), lr=2e-4)
r = TrainStepRunner(m, opt, 256, 64, x_dtype=torch.bfloat16 if (len(sys.argv) <= 2 or sys.argv[2] == "bf16") else None)   # as benched: bf16 keypoint input
r.load(synthetic.model_batch(256, 64, seed=1234), non_blocking=False)
for _ in range(reps):
    r.step(0)
torch.cuda.synchronize()
print("prof_target done, loss", float(r.loss[0]))

# wide variant (conv_channels = 256): streamed-weight forward kernel, 4 tiles per SM
wm = b2h.ConvModel(256, "ReLU", False, precision="bf16").to(dev)
wr = ForwardRunner(wm, 1184, 126, x_dtype=torch.bfloat16)
wr.x[0].copy_(synthetic.model_batch(1184, 126, seed=300)["input_kp"])
for _ in range(reps):
    wr.run(0)
torch.cuda.synchronize()

# fp32 mode on the tensor pipe (bf16 high/low operand pairs): forward B=512 and the fused train step B=256
f32 = b2h.ConvModel(30, "ReLU", False, precision="fp32").to(dev)
fr32 = ForwardRunner(f32, 512, 64)
fr32.x[0].copy_(synthetic.model_batch(512, 64, seed=99)["input_kp"])
o32 = b2h.FusedAdam(f32.parameters(), lr=2e-4)
r32 = TrainStepRunner(f32, o32, 256, 64)
r32.load(synthetic.model_batch(256, 64, seed=1234), non_blocking=False)
for _ in range(reps):
    fr32.run(0)
    r32.step(0)
torch.cuda.synchronize()

# wide training (conv_channels = 256, batch 256 x 64): forward+criterion / dgrad chain / split-K wgrad / Adam
wt = b2h.ConvModel(256, "ReLU", False, precision="bf16").to(dev)
wo = b2h.FusedAdam(wt.parameters(), lr=2e-4)
wr2 = TrainStepRunner(wt, wo, 256, 64)
wr2.load(synthetic.model_batch(256, 64, seed=5), non_blocking=False)
for _ in range(reps):
    wr2.step(0)
torch.cuda.synchronize()
print("prof_target: fp32 loss", float(r32.loss[0]), "wide loss", float(wr2.loss[0]))

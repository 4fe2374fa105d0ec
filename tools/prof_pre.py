import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hand_pose_sl_b200 as b2h
from hand_pose_sl_b200 import synthetic
dev = torch.device("cuda:0")
F = 216000
pose, lh, rh = synthetic.synthetic_clip(F, seed=1)
tp, tl, tr = (torch.from_numpy(a).to(dev) for a in (pose, lh, rh))
pre = b2h.PreprocessRightHand()
starts = torch.zeros(1, dtype=torch.int64, device=dev)
out = pre(tp, tl, tr, starts, F)
for _ in range(3): pre(tp, tl, tr, starts, F, out=out)
torch.cuda.synchronize(); print("ok")

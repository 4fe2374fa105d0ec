"""Device time of the forward and the fused train step over (B, T, precision) shapes outside the headline configs,
e.g. the reference default --max-frames 200.  CUDA events around 10 replays of a 20-launch CUDA graph after warm-up; prints one line per shape."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hand_pose_sl_b200 as b2h
from hand_pose_sl_b200 import synthetic
from hand_pose_sl_b200.runner import ForwardRunner, TrainStepRunner

dev = torch.device("cuda:0")
SHAPES = [(256, 64), (128, 128), (82, 200), (164, 200), (64, 256)]


def timed(fn, n=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n   # us


for prec in ("bf16", "fp32"):
    for B, T in SHAPES:
        torch.manual_seed(0)
        m = b2h.ConvModel(30, "ReLU", False, precision=prec).to(dev)
        fr = ForwardRunner(m, B, T)
        fr.x[0].copy_(synthetic.model_batch(B, T, seed=99)["input_kp"])
        opt = b2h.FusedAdam(m.parameters(), lr=2e-4)
        tr = TrainStepRunner(m, opt, B, T)
        tr.load(synthetic.model_batch(B, T, seed=1234), non_blocking=False)
        fr.capture(20); tr.capture(20)              # 20 launches per graph replay: device time, not host enqueue time
        f_us = timed(lambda: fr.graph.replay(), 10) / 20
        t_us = timed(lambda: tr.replay(), 10) / 20
        fr_s = B * T / f_us
        print(f"{prec} B={B:4d} T={T:3d}: fwd {f_us:7.1f} us ({fr_s:7.1f} Mframes/s)   train step {t_us:7.1f} us ({B * T / t_us:7.1f} Mframes/s)",
              flush=True)

"""Device time of the streamed-weight wide forward kernel (conv_channels = 256, bf16 mode): CUDA-graph replay of 10
launches, CUDA events; prints frames/s and algorithmic TFLOP/s (2*(5T-6)*S(C) per window)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hand_pose_sl_b200 as b2h
from hand_pose_sl_b200 import synthetic
from hand_pose_sl_b200.runner import ForwardRunner

dev = torch.device("cuda:0")
C = int(sys.argv[1]) if len(sys.argv) > 1 else 256
S = 24 * C + 2 * C * C + 42 * C
for B, T in [(444, 64), (512, 64), (1776, 64), (296, 126), (1184, 126), (148, 200), (592, 200)]:
    torch.manual_seed(0)
    m = b2h.ConvModel(C, "ReLU", False, precision="bf16").to(dev)
    fr = ForwardRunner(m, B, T, x_dtype=torch.bfloat16)
    fr.x[0].copy_(synthetic.model_batch(B, T, seed=99)["input_kp"])
    fr.capture(10)
    for _ in range(3):
        fr.graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        fr.graph.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / 50
    flops = 2.0 * (5 * T - 6) * S * B
    print(f"C={C} B={B:5d} T={T:3d}: {us:8.1f} us  {B * T / us:8.1f} Mframes/s  {flops / us * 1e-6:7.1f} TFLOP/s algorithmic", flush=True)

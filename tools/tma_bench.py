"""1-D bulk-copy (cp.async.bulk) L2 -> shared throughput per SM vs copy size, ring depth and number of busy SMs."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hand_pose_sl_b200 import _lib
lib = _lib.load()
out = torch.zeros(2, dtype=torch.int64, device="cuda")
print("copy_bytes depth grid | bytes/cycle/SM")
for grid in (1, 148):
    for depth in (4, 11):
        for cb in (8192, 2048, 512, 256):
            for _ in range(2):
                _lib.check(lib.b2h_tc_bench(_lib.ptr(out), 1, cb, 50, depth, grid, _lib.stream_ptr()))
            torch.cuda.synchronize()
            cyc, nbytes = out.cpu().tolist()
            print(f"{cb:6d} {depth:3d} {grid:4d} | {nbytes / cyc:7.1f}", flush=True)

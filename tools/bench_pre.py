"""K0 micro-benchmark: graph-replayed preprocessing over clips of several lengths; prints GB/s (1452 B/frame)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hand_pose_sl_b200 as b2h
from hand_pose_sl_b200 import synthetic
dev = torch.device("cuda:0")
for F in (108000, 216000, 864000):
    pose, lh, rh = synthetic.synthetic_clip(F, seed=1)
    tp, tl, tr = (torch.from_numpy(a).to(dev) for a in (pose, lh, rh))
    pre = b2h.PreprocessRightHand()
    starts = torch.zeros(1, dtype=torch.int64, device=dev)
    out = pre(tp, tl, tr, starts, F)
    for _ in range(3): pre(tp, tl, tr, starts, F, out=out)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(4): pre(tp, tl, tr, starts, F, out=out)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): g.replay()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"F={F}: {ms*1e3:.1f} us/launch, {F*1452/ms/1e6:.0f} GB/s, {F/ms*1e3/1e9:.2f} Gframes/s")
    del tp, tl, tr, out

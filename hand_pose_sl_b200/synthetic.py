"""Seeded synthetic How2Sign-shaped inputs.  No real How2Sign data exists in the
reference tree, so every test and benchmark uses these generators.  Pure numpy/torch CPU code; the
arrays are produced directly in fp32 so JSON double->float rounding is not part of parity."""
from __future__ import annotations

import numpy as np
import torch

FPS = 30


def synthetic_clip(n_frames: int, seed: int = 1234, undetected: float = 0.05):
    """One OpenPose clip: pose25 (F,25,3), hand_left (F,21,3), hand_right (F,21,3) fp32 rows of
    [x, y, c] -- the arrays `people[0].{pose,hand_left,hand_right}_keypoints_2d` hold per frame
    (schema: reference How2Sign/util_scripts/build_dataset.py:56-74).  x in [0,1280), y in [0,720)
    follow a sigma=3 px/frame random walk; c ~ U(0,1); a fraction `undetected` of keypoints is
    x=y=c=0 (OpenPose's convention for a missed detection)."""
    rng = np.random.default_rng(seed)

    def part(k):
        start = np.stack([rng.uniform(0, 1280, size=k), rng.uniform(0, 720, size=k)], axis=-1)
        walk = rng.normal(0.0, 3.0, size=(n_frames, k, 2)).cumsum(axis=0)
        xy = start[None] + walk
        xy[..., 0] = np.clip(xy[..., 0], 0, 1279.0)
        xy[..., 1] = np.clip(xy[..., 1], 0, 719.0)
        c = rng.uniform(0, 1, size=(n_frames, k, 1))
        arr = np.concatenate([xy, c], axis=-1).astype(np.float32)
        miss = rng.uniform(size=(n_frames, k)) < undetected
        arr[miss] = 0.0
        return np.ascontiguousarray(arr)

    return part(25), part(21), part(21)


def window_starts(n_frames: int, T: int = 64, stride: int = 64):
    """Start frame of every sliding window over a clip (config 5: stride 64 disjoint, 16 overlap x4);
    the last window may be partial and is padded by the reference rule."""
    return np.arange(0, max(n_frames, 1), stride, dtype=np.int64)


def model_batch(B: int, T: int = 64, seed: int = 1234, ragged: bool = False, len_seed: int = 4321,
                n_kp: int = 12):
    """Model-only inputs (configs 1-4): input_kp (B,T,12,2), target_kp (B,T,21,2), target_conf (B,T,21)
    fp32 in the range the reference transforms produce (about +-1), lengths (B,) int64: all T for
    throughput runs, randint(8, T+1) for parity runs."""
    g = torch.Generator().manual_seed(seed)
    base = torch.randn(B, 1, n_kp, 2, generator=g) * 0.15
    walk = (torch.randn(B, T, n_kp, 2, generator=g) * (3.0 / 1280.0)).cumsum(dim=1)
    input_kp = (base + walk).float().contiguous()
    input_kp[:, :, 1, :] = 0.0                      # ChestDifference makes row 1 exactly 0
    tb = torch.randn(B, 1, 21, 2, generator=g) * 0.05
    tw = (torch.randn(B, T, 21, 2, generator=g) * (3.0 / 1280.0)).cumsum(dim=1)
    target_kp = (tb + tw).float().contiguous()
    target_conf = torch.rand(B, T, 21, generator=g).float().contiguous()
    if ragged:
        gl = torch.Generator().manual_seed(len_seed)
        lengths = torch.randint(min(8, T), T + 1, (B,), generator=gl, dtype=torch.int64)
    else:
        lengths = torch.full((B,), T, dtype=torch.int64)
    return {"input_kp": input_kp, "target_kp": target_kp, "target_conf": target_conf, "n_frames": lengths}

"""CPU ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A CPU restatement of the body2hand temporal-conv hot path of benoriol/hand_pose_sl, used only as
the checker in tests/, in __graft_entry__.smoke() and as the `cpu_baseline` / `--impl reference`
arm of bench.py.  Nothing under hand_pose_sl_b200/ imports this file; the product path fails
loudly when the CUDA library is missing and never falls back to this code.

Parity status: the reference ships NO tests, golden vectors or known-answer fixtures for this path
(SURVEY.md §4/§8c).  The restatement is therefore pinned against the *reference's own classes
executed in the build container* (oracle/ref_loader.py loads the unmodified files from
/root/reference; oracle/make_golden.py writes the vectors committed under tests/golden/;
tests/test_oracle_golden.py re-checks this file against them on every run, and
tests/test_oracle_vs_reference.py against the live reference when it is mounted).

Arithmetic engine: the reference's floating-point math is PyTorch's (torch is a third-party
dependency of the reference, un-pinned: no requirements/setup/lock file exists; .pyc files are
cpython-37 ~ torch 1.5-1.7).  Oracle of record = torch 2.11.0 CPU fp32 in this image, i.e. the same
aten ops the reference calls (aten::conv1d -> mkldnn_convolution, relu, l1_loss, Adam).
Index / windowing / gather logic is restated in numpy and is bit-exact.

Every function cites the reference file:line it follows (paths relative to /root/reference).
"""
from __future__ import annotations

import math
import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------
# constants
# --------------------------------------------------------------------------------------
# body2hand/src/dataloaders/text_pose_dataset.py:14
BODY_HEAD_KEYPOINTS = [0, 1, 2, 3, 4, 5, 6, 7, 15, 16, 17, 18]
RIGHT_WRIST_INDEX = 4   # body2hand/src/steps/utils.py:198  (index inside the 12-subset)
CHEST_INDEX = 1         # body2hand/src/steps/utils.py:207
NORM_FACTOR = 1280      # body2hand/src/run.py:90
PAD_REPEAT_FIRST = 0    # JSON datasets: text_pose_dataset.py:512-518
PAD_ZEROS = 1           # H5 dataset:    text_pose_dataset.py:616-622


# --------------------------------------------------------------------------------------
# a9: keypoint parsing / 12-of-25 gather   (text_pose_dataset.py:16-50)
# --------------------------------------------------------------------------------------
def format_keypoints(flat, n_dim=2):
    """[x1,y1,c1,x2,...] -> [[x1,y1,c1],...]   text_pose_dataset.py:16-26"""
    n = n_dim + 1
    flat = np.asarray(flat)
    return flat[..., : (flat.shape[-1] // n) * n].reshape(flat.shape[:-1] + (-1, n))


def load_keypoints_arrays(pose75, lhand63, rhand63):
    """Array form of load_keypoints (text_pose_dataset.py:29-50) for a stack of frames.
    pose75 (F,75), lhand63 (F,63), rhand63 (F,63) -> 6-tuple in the reference's return order
    (r_kp, r_conf, l_kp, l_conf, body_kp, body_conf)."""
    body = format_keypoints(pose75)[:, BODY_HEAD_KEYPOINTS, :]      # :38-40
    lh = format_keypoints(lhand63)                                   # :42
    rh = format_keypoints(rhand63)                                   # :43
    return (rh[..., :2], rh[..., 2], lh[..., :2], lh[..., 2], body[..., :2], body[..., 2])  # :46-50


# --------------------------------------------------------------------------------------
# a10/a11: windowing   (text_pose_dataset.py:52-68, 447, 511-529, 614-635)
# --------------------------------------------------------------------------------------
def select_window(n_total, n, selection_type, rand_start=None):
    """select_jsons (text_pose_dataset.py:52-68) as index math: returns (start, stop).
    `rand_start` is the value random.randint(0, n_total-n) returned (inclusive both ends)."""
    if n_total <= n:
        return 0, n_total                     # :60-61
    if selection_type == "first":
        return 0, n                           # :63-64
    if selection_type == "randomcrop":
        assert 0 <= rand_start <= n_total - n
        return rand_start, rand_start + n     # :66-68
    raise ValueError("selection_type must be given for long clips (reference returns None: :52-68)")


def window_frame_index(start, stop, n, pad_mode):
    """Source-frame index of every slot of an n-slot window cropped to [start,stop):
    pad (repeat crop frame 0, :512-518; zeros -> -1, :616-622) then clip (:520-529)."""
    idx = np.arange(start, min(stop, start + n), dtype=np.int64)
    if idx.shape[0] < n:
        fill = start if pad_mode == PAD_REPEAT_FIRST else -1
        idx = np.concatenate([idx, np.full(n - idx.shape[0], fill, dtype=np.int64)])
    return idx


def n_frames_of(n_frames_meta, n):
    """item['n_frames'] = min(metadata n_frames, max_frames)   text_pose_dataset.py:447 / :661-662"""
    return min(int(n_frames_meta), int(n))


def gather_frames(arr, idx):
    """Apply window_frame_index to an (F, ...) array; -1 -> zeros (numpy.pad default, :616-622)."""
    out = np.zeros((idx.shape[0],) + arr.shape[1:], dtype=arr.dtype)
    ok = idx >= 0
    out[ok] = arr[idx[ok]]
    return out


# --------------------------------------------------------------------------------------
# a12-a15: transforms   (steps/utils.py:180-210, 261-277)
# --------------------------------------------------------------------------------------
def wrist_difference(item):
    """steps/utils.py:199-201"""
    item = dict(item)
    item["right_hand_kp"] = item["right_hand_kp"] - item["body_kp"][:, RIGHT_WRIST_INDEX][:, None]
    return item


def chest_difference(item):
    """steps/utils.py:208-210"""
    item = dict(item)
    item["body_kp"] = item["body_kp"] - item["body_kp"][:, CHEST_INDEX][:, None]
    return item


def normalize_fixed_factor(item, factor=NORM_FACTOR):
    """steps/utils.py:184-190 -- true fp32 division, confidences untouched"""
    item = dict(item)
    f = np.float32(factor)
    for k in ("body_kp", "right_hand_kp", "left_hand_kp"):
        item[k] = (item[k] / f).astype(np.float32)
    return item


def build_right_hand_item(item):
    """steps/utils.py:263-277"""
    item = dict(item)
    item["input_kp"] = item["body_kp"]
    item["target_kp"] = item["right_hand_kp"]
    item["input_conf"] = item["body_conf"]
    item["target_conf"] = item["right_hand_conf"]
    return item


def apply_transforms(item, dif_encoding=True, normalize=True, factor=NORM_FACTOR):
    """The Compose built in run.py:85-107: [WristDifference, ChestDifference] (if dif_encoding),
    NormalizeFixedFactor(1280) (unless --no-normalize), BuildRightHandItem."""
    if dif_encoding:
        item = wrist_difference(item)      # run.py:86  (uses the RAW wrist: runs before ChestDifference)
        item = chest_difference(item)      # run.py:87
    if normalize:
        item = normalize_fixed_factor(item, factor)   # run.py:90
    return build_right_hand_item(item)     # run.py:102


def array2item(array):
    """TextPoseH5Dataset.array2item  text_pose_dataset.py:587-612.
    (N,150) rows [x0..x49 | y0..y49 | c0..c49] -> body 8 kpts, left hand 8:29, right hand 29:50."""
    n = array.shape[0]
    a = array.reshape((n, 3, -1)).transpose(0, 2, 1)
    kp, conf = a[:, :, :2], a[:, :, 2]
    return {
        "body_kp": kp[:, :8, :], "body_conf": conf[:, :8],
        "left_hand_kp": kp[:, 8:29, :], "left_hand_conf": conf[:, 8:29],
        "right_hand_kp": kp[:, 29:, :], "right_hand_conf": conf[:, 29:],
    }


def preprocess_windows(pose25, lhand, rhand, win_start, T, pad_mode=PAD_REPEAT_FIRST,
                       dif_encoding=True, normalize=True, factor=NORM_FACTOR):
    """The whole data-item construction of SURVEY.md §3.4 for a list of windows over one clip:
    load_keypoints (a9) -> crop [start,start+T) (a10) -> pad/clip/n_frames (a11) -> float32 (:537-542)
    -> transforms (a12-a15).  pose25 (F,25,3), lhand/rhand (F,21,3) fp32 OpenPose [x,y,c].
    Returns dict of stacked (W,T,...) float32 arrays + n_frames (W,) int64."""
    Fr = pose25.shape[0]
    r_kp, r_cf, l_kp, l_cf, b_kp, b_cf = load_keypoints_arrays(
        pose25.reshape(Fr, 75), lhand.reshape(Fr, 63), rhand.reshape(Fr, 63))
    keys = ["body_kp", "body_conf", "right_hand_kp", "right_hand_conf", "left_hand_kp", "left_hand_conf",
            "input_kp", "input_conf", "target_kp", "target_conf"]
    out = {k: [] for k in keys}
    nfr = []
    for s in win_start:
        s = int(s)
        idx = window_frame_index(s, Fr, T, pad_mode)
        item = {
            "body_kp": gather_frames(b_kp, idx).astype(np.float32),
            "body_conf": gather_frames(b_cf, idx).astype(np.float32),
            "right_hand_kp": gather_frames(r_kp, idx).astype(np.float32),
            "right_hand_conf": gather_frames(r_cf, idx).astype(np.float32),
            "left_hand_kp": gather_frames(l_kp, idx).astype(np.float32),
            "left_hand_conf": gather_frames(l_cf, idx).astype(np.float32),
        }
        item = apply_transforms(item, dif_encoding, normalize, factor)
        for k in keys:
            out[k].append(item[k])
        nfr.append(n_frames_of(Fr - s, T))
    res = {k: np.stack(v).astype(np.float32) for k, v in out.items()}
    res["n_frames"] = np.asarray(nfr, dtype=np.int64)
    return res


# --------------------------------------------------------------------------------------
# a1-a3: ConvModel   (models/HandPoseModels.py:17-84)
# --------------------------------------------------------------------------------------
PARAM_NAMES = ["conv1.weight", "conv1.bias", "conv2.weight", "conv2.bias",
               "conv3.weight", "conv3.bias", "conv4.weight", "conv4.bias"]


def init_params(conv_channels=30, pos_emb=False, seed=0, n_in=24):
    """Parameters with the reference's construction order and torch default Conv1d init
    (HandPoseModels.py:24-32) so that `torch.manual_seed(seed)` yields the same weights as
    `ConvModel(conv_channels,'ReLU',pos_emb)` built right after the same seed."""
    torch.manual_seed(seed)
    cin = n_in + (1 if pos_emb else 0)
    convs = [torch.nn.Conv1d(cin, conv_channels, 5, padding=2),
             torch.nn.Conv1d(conv_channels, conv_channels, 5, padding=2),
             torch.nn.Conv1d(conv_channels, conv_channels, 5, padding=2),
             torch.nn.Conv1d(conv_channels, 42, 5, padding=2)]
    sd = {}
    for i, c in enumerate(convs, 1):
        sd[f"conv{i}.weight"] = c.weight.detach().clone()
        sd[f"conv{i}.bias"] = c.bias.detach().clone()
    return sd


def linear_positional_embedding(x, max_len=100):
    """LinearPositionalEmbedding.forward  HandPoseModels.py:68-84: prepend channel t/max_len."""
    pe = (torch.arange(max_len).float() / max_len)[None, None, :]      # :70-75
    pe = pe.expand(x.shape[0], -1, -1)                                 # :80
    return torch.cat([pe, x], dim=1)                                   # :82  (T must equal max_len)


def conv_model_forward(sd, inp, pos_emb=False):
    """ConvModel.forward  HandPoseModels.py:40-64.  inp (B,T,K,2) fp32 -> (B,T,21,2)."""
    x = inp.permute(0, 2, 3, 1)                                  # :43
    bs, nk, dim, ln = x.shape
    x = x.reshape(bs, nk * dim, ln)                              # :46 (view on the permuted tensor)
    if pos_emb:
        x = linear_positional_embedding(x)                       # :53
    out = F.relu(F.conv1d(x, sd["conv1.weight"], sd["conv1.bias"], padding=2))     # :55
    out = F.relu(F.conv1d(out, sd["conv2.weight"], sd["conv2.bias"], padding=2))   # :56
    out = F.relu(F.conv1d(out, sd["conv3.weight"], sd["conv3.bias"], padding=2))   # :57
    out = F.conv1d(out, sd["conv4.weight"], sd["conv4.bias"], padding=2)           # :58
    out = out.view(bs, -1, dim, ln)                              # :60
    return out.permute(0, 3, 1, 2)                               # :62


def conv_model_forward_f64(sd, inp, pos_emb=False):
    """Independent float64 numpy evaluation of the same formula
    out[b,o,t] = bias[o] + sum_{i,k} W[o,i,k] * in[b,i,t+k-2] (zeros outside [0,T)) -- the arbiter
    when the fp32 oracle and a kernel disagree close to tolerance."""
    x = np.asarray(inp, dtype=np.float64)
    B, T = x.shape[:2]
    a = x.reshape(B, T, -1)                       # NWC: channel c = 2*kp + d  (HandPoseModels.py:43-46)
    if pos_emb:
        pe = (np.arange(100, dtype=np.float32) / np.float32(100)).astype(np.float64)
        a = np.concatenate([np.broadcast_to(pe[None, :, None], (B, T, 1)), a], axis=2)
    for li in range(1, 5):
        W = sd[f"conv{li}.weight"].double().numpy()      # (Cout,Cin,5)
        b = sd[f"conv{li}.bias"].double().numpy()
        ap = np.pad(a, ((0, 0), (2, 2), (0, 0)))
        out = np.zeros((B, T, W.shape[0]))
        for k in range(5):
            out += ap[:, k:k + T, :] @ W[:, :, k].T
        out += b
        a = np.maximum(out, 0.0) if li < 4 else out
    return a.reshape(B, T, 21, 2)


# --------------------------------------------------------------------------------------
# a4-a6: mask_output and the two criteria   (steps/utils.py:309-312, 413-452)
# --------------------------------------------------------------------------------------
def mask_output(output, lengths):
    """steps/utils.py:309-312 (in place, returns the same tensor)"""
    for i, ln in enumerate(lengths):
        output[i, int(ln):, :] = 0
    return output


def masked_pose_l1(prediction, target, lengths):
    """maskedPoseL1.forward  steps/utils.py:420-428 -- literal per-sample loop."""
    loss = 0
    i = 0
    for i, seq_len in enumerate(lengths):
        loss = loss + F.l1_loss(prediction[i, :int(seq_len)], target[i, :int(seq_len)], reduction="mean")
    return loss / (i + 1)


def poderated_pose_l1(prediction, target, lengths, scores):
    """poderatedPoseL1.forward  steps/utils.py:437-452 -- batch SUM of per-sample means."""
    loss = 0
    for i, seq_len in enumerate(lengths):
        n = int(seq_len)
        s = scores[i, :n].unsqueeze(2)
        loss = loss + F.l1_loss(prediction[i, :n] * s, target[i, :n] * s, reduction="mean")
    return loss


def masked_pose_l1_closed_form(prediction, target, lengths):
    """Closed form of a5: (1/B) sum_i sum_{t<len_i}|d| / (len_i*J*D)."""
    B, T = prediction.shape[:2]
    per = prediction.shape[2] * prediction.shape[3]
    ln = torch.as_tensor(lengths, dtype=torch.int64)
    m = (torch.arange(T)[None, :] < ln[:, None]).to(prediction.dtype)
    s = ((prediction - target).abs().sum(dim=(2, 3)) * m).sum(dim=1)
    return (s / (ln.to(prediction.dtype) * per)).sum() / B


# --------------------------------------------------------------------------------------
# a7/a8: train step and validation   (steps/traintest.py:48, 87-123, 168-211)
# --------------------------------------------------------------------------------------
class TrainState:
    """Model parameters + torch.optim.Adam exactly as traintest.py:48 builds them."""

    def __init__(self, sd, lr=2e-4, pos_emb=False):
        self.pos_emb = pos_emb
        self.params = {k: sd[k].detach().clone().requires_grad_(True) for k in PARAM_NAMES}
        self.opt = torch.optim.Adam([self.params[k] for k in PARAM_NAMES], lr=lr)   # traintest.py:48

    def state_dict(self):
        return {k: v.detach().clone() for k, v in self.params.items()}


def train_step(state: TrainState, input_kp, target_kp, lengths, loss_kind="L1", target_conf=None):
    """One iteration of the hot loop, traintest.py:87-123, restated literally:
    forward (:94) -> mask_output (:111) -> criterion (:115/:117) -> zero_grad/backward/step (:119-121)
    -> loss.item() (:123).  Returns (loss_float, grads dict)."""
    prediction = conv_model_forward(state.params, input_kp, state.pos_emb)      # :94
    prediction = mask_output(prediction, lengths)                               # :111
    if loss_kind == "L1":
        loss = masked_pose_l1(prediction, target_kp, lengths)                   # :115
    elif loss_kind == "confL1":
        loss = poderated_pose_l1(prediction, target_kp, lengths, target_conf)   # :117
    else:
        raise ValueError(loss_kind)     # MSE/huber never compute a loss in the reference (:114-117)
    state.opt.zero_grad()               # :119
    loss.backward()                     # :120
    grads = {k: state.params[k].grad.detach().clone() for k in PARAM_NAMES}
    state.opt.step()                    # :121
    return float(loss.item()), grads    # :123


@torch.no_grad()
def validate_batch(sd, input_kp, target_kp, lengths, loss_kind="L1", target_conf=None, pos_emb=False):
    """validate() body, traintest.py:174-207, for one batch."""
    prediction = mask_output(conv_model_forward(sd, input_kp, pos_emb), lengths)
    if loss_kind == "L1":
        return float(masked_pose_l1(prediction, target_kp, lengths))
    return float(poderated_pose_l1(prediction, target_kp, lengths, target_conf))


def train_grads_bf16_emulated(sd, input_kp, target_kp, lengths, loss_kind="L1", target_conf=None):
    """What an IDEAL bf16-operand / fp32-accumulate implementation of traintest.py:94-120 computes: the reference
    formulas with every GEMM operand rounded to bf16 (weights, input, post-ReLU activations, dY, dZ) and everything
    else (accumulation, bias, loss, sign) in fp32.  Used to separate kernel bugs from bf16 rounding: the tensor-core
    kernels must match THIS closely, and this matches the fp32 reference within the bf16 tolerances."""
    bf = lambda t: t.to(torch.bfloat16).float()
    B, T = input_kp.shape[:2]
    ln = torch.as_tensor(lengths, dtype=torch.int64)
    a = [bf(input_kp.reshape(B, T, -1).permute(0, 2, 1))]
    W = [bf(sd[f"conv{i}.weight"]) for i in range(1, 5)]
    b = [sd[f"conv{i}.bias"] for i in range(1, 5)]
    for l in range(3):
        a.append(bf(F.relu(F.conv1d(a[-1], W[l], b[l], padding=2))))
    pred = F.conv1d(a[-1], W[3], b[3], padding=2)                               # (B,42,T) fp32
    mask = (torch.arange(T)[None, :] < ln[:, None]).float()[:, None, :]
    pred = pred * mask                                                           # mask_output  utils.py:309-312
    t42 = target_kp.reshape(B, T, 42).permute(0, 2, 1)
    n_el = (ln.float() * 42)[:, None, None]
    if loss_kind == "L1":
        d = pred - t42
        loss = ((d.abs() * mask).sum(dim=(1, 2)) / n_el[:, 0, 0]).sum() / B
        dy = torch.sign(d) * mask / (B * n_el)
    else:
        s = target_conf.permute(0, 2, 1).repeat_interleave(2, dim=1)
        d = pred * s - t42 * s
        loss = ((d.abs() * mask).sum(dim=(1, 2)) / n_el[:, 0, 0]).sum()
        dy = torch.sign(d) * s * mask / n_el
    grads = {}
    dz = bf(dy)
    for l in (3, 2, 1, 0):
        grads[f"conv{l + 1}.weight"] = torch.nn.grad.conv1d_weight(a[l], W[l].shape, dz, padding=2)
        grads[f"conv{l + 1}.bias"] = dz.sum(dim=(0, 2))
        if l > 0:
            dz = bf(torch.nn.grad.conv1d_input(a[l].shape, W[l], dz, padding=2) * (a[l] > 0).float())
    return float(loss), grads, pred.permute(0, 2, 1).reshape(B, T, 21, 2)


def adam_reference_step(p, g, m, v, step, lr=2e-4, b1=0.9, b2=0.999, eps=1e-8):
    """torch.optim.Adam single-tensor update (defaults of traintest.py:48), written out; float64 in,
    used to cross-check the fused CUDA Adam independent of torch's foreach implementation."""
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    bc1 = 1 - b1 ** step
    bc2 = 1 - b2 ** step
    denom = np.sqrt(v) / math.sqrt(bc2) + eps
    p = p - (lr / bc1) * m / denom
    return p, m, v


def adam_conditioned(grads_per_step, weights, lr, eps=1e-8, margin=3.0):
    """Adam parity is ill-conditioned where a gradient element is close to rounding noise: the update
    is u = lr*g/(|g|+eps), so du/dg = lr*eps/(|g|+eps)^2 explodes for |g| -> eps (1e-8); an element
    whose gradient is a near-cancelling sign-sum (the L1 loss produces them) moves by anything in
    [-lr, lr] depending on summation order, and the reference itself is not reproducible there across
    thread counts.  With a gradient tolerance of tol*max|g| and a weight tolerance of tol*max|w| the
    update is within tolerance iff (|g|+eps)^2 >= lr*eps*max|g|/max|w|.  Returns, per parameter name,
    the mask of elements meeting that bound (times `margin`) at EVERY step (an exact 0 can be an exact
    cancellation in one summation order and 1e-10 in another, so zeros are excluded too); post-step weight parity is asserted on those, and every element is still
    bounded by 2*lr per step."""
    masks = {}
    for grads in grads_per_step:
        for k, g in grads.items():
            g = np.abs(np.asarray(g, dtype=np.float64))
            wmax = max(float(np.abs(np.asarray(weights[k], dtype=np.float64)).max()), 1e-30)
            thr = margin * math.sqrt(lr * eps * max(float(g.max()), 1e-30) / wmax)
            m = g >= thr
            masks[k] = m if k not in masks else (masks[k] & m)
    return masks


def array2open_pose(array):
    """steps/utils.py:355-364: (21,2) -> flat [x,y,1.0]*21 (confidence defaults to ones)."""
    a = np.concatenate((np.asarray(array), np.zeros((21, 1)) + 1.0), axis=1)
    return np.reshape(a, (-1)).astype(np.float32)


def order_and_reshape_toh5(frame_prediction):
    """steps/traintest.py:302-317: (n,21,2) -> (n,63) rows [x*21 | y*21 | 0*21]."""
    fp = np.pad(np.asarray(frame_prediction), ((0, 0), (0, 0), (0, 1)))
    return fp.transpose(0, 2, 1).reshape(fp.shape[0], -1)


def l1_to_pixels(loss, num_joints=21, upsample=1280):
    """L12Pixels  steps/utils.py:291-299"""
    return loss / num_joints * upsample


def rel_err(a, b):
    """Tolerance definition of SURVEY.md §8c: max|a-b| / max(max|b|, tiny)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(float(np.max(np.abs(b))), 1e-30))

"""CPU tier: the oracle restatement (oracle/b2h_oracle.py) against the golden vectors produced by the
UNMODIFIED reference (oracle/make_golden.py).  This is what pins the oracle."""
import numpy as np
import pytest
import torch

import b2h_oracle as oracle
from conftest import golden_sd, load_golden

MODEL_FIXTURES = ["convmodel_c30.npz", "convmodel_c30_b1.npz", "convmodel_c30_posemb.npz", "convmodel_c64.npz",
                  "convmodel_c30_t200.npz"]


@pytest.mark.parametrize("name", MODEL_FIXTURES)
def test_forward_matches_reference(name):
    g = load_golden(name)
    sd = golden_sd(g)
    x = torch.from_numpy(g["input_kp"])
    pe = bool(g["pos_emb"])
    y = oracle.conv_model_forward(sd, x, pe).contiguous().numpy()
    assert oracle.rel_err(y, g["pred"]) < 2e-6
    y64 = oracle.conv_model_forward_f64(sd, x.numpy(), pe)
    assert oracle.rel_err(y64, g["pred"]) < 5e-6
    masked = oracle.mask_output(torch.from_numpy(y.copy()), g["lengths"]).numpy()
    assert np.array_equal(masked, oracle.mask_output(torch.from_numpy(g["pred"].copy()), g["lengths"]).numpy()) or \
        oracle.rel_err(masked, g["pred_masked"]) < 2e-6
    # masking zeroes exactly the rows t >= len
    for i, ln in enumerate(g["lengths"]):
        assert not masked[i, int(ln):].any()


def test_wide_forward_matches_reference():
    """conv_channels = 256: the fixture stores only the reference's outputs; the weights are the reference's default
    init after torch.manual_seed(0), which oracle.init_params reproduces (per-tensor checksums in the fixture)."""
    g = load_golden("convmodel_c256_fwd.npz")
    sd = oracle.init_params(int(g["C"]), False, seed=int(g["seed"]))
    for k, v in sd.items():
        kk = k.replace(".", "_")
        assert abs(v.double().sum().item() - float(g["sum_" + kk])) <= 1e-9 * float(g["abssum_" + kk]), k
        assert abs(v.double().abs().sum().item() - float(g["abssum_" + kk])) <= 1e-9 * float(g["abssum_" + kk]), k
    x = torch.from_numpy(g["input_kp"])
    y = oracle.conv_model_forward(sd, x).contiguous()
    assert oracle.rel_err(y.numpy(), g["pred"]) < 5e-6
    masked = oracle.mask_output(y.clone(), g["lengths"]).numpy()
    assert oracle.rel_err(masked, g["pred_masked"]) < 5e-6
    assert oracle.rel_err(oracle.conv_model_forward_f64(sd, x.numpy()), g["pred"]) < 1e-5


@pytest.mark.parametrize("name", MODEL_FIXTURES)
@pytest.mark.parametrize("kind", ["L1", "confL1"])
def test_train_steps_match_reference(name, kind):
    g = load_golden(name)
    sd = golden_sd(g)
    x, tgt = torch.from_numpy(g["input_kp"]), torch.from_numpy(g["target_kp"])
    conf, lengths = torch.from_numpy(g["target_conf"]), torch.from_numpy(g["lengths"])
    st = oracle.TrainState(sd, lr=float(g["lr"]), pos_emb=bool(g["pos_emb"]))
    steps = len(g[f"loss_{kind}"])
    all_grads = []
    for s in range(steps):
        loss, grads = oracle.train_step(st, x, tgt, lengths, kind, conf)
        all_grads.append({k: v.numpy() for k, v in grads.items()})
        assert abs(loss - g[f"loss_{kind}"][s]) <= 2e-6 * abs(g[f"loss_{kind}"][s])
        if s == 0:
            for k, v in grads.items():
                assert oracle.rel_err(v.numpy(), g[f"grad_{kind}_" + k.replace(".", "_")]) < 1e-5, k
    lr = float(g["lr"])
    masks = oracle.adam_conditioned(all_grads, sd, lr)   # see its docstring: Adam on noise-level gradients
    assert sum(int(m.sum()) for m in masks.values()) > 0.3 * sum(m.size for m in masks.values())
    for k, v in st.state_dict().items():
        want = g[f"w{steps}_{kind}_" + k.replace(".", "_")]
        m = masks[k]
        assert np.abs(v.numpy() - want)[m].max() <= 1e-5 * np.abs(want).max(), k
        assert np.abs(v.numpy() - want).max() <= 2 * lr * steps, k


def test_closed_form_l1_equals_loop():
    g = load_golden("convmodel_c30.npz")
    pred = torch.from_numpy(g["pred_masked"])
    tgt = torch.from_numpy(g["target_kp"])
    a = float(oracle.masked_pose_l1(pred, tgt, g["lengths"]))
    b = float(oracle.masked_pose_l1_closed_form(pred, tgt, g["lengths"]))
    assert abs(a - b) < 1e-6 * abs(a)


@pytest.mark.parametrize("tag,dif", [("dif", True), ("nodif", False)])
def test_preprocess_matches_reference_bit_exact(tag, dif):
    g = load_golden("preprocess.npz")
    out = oracle.preprocess_windows(g["pose25"], g["hand_left"], g["hand_right"], g["win_start"], int(g["T"]),
                                    oracle.PAD_REPEAT_FIRST, dif_encoding=dif)
    for k in ("input_kp", "input_conf", "target_kp", "target_conf", "left_hand_kp", "left_hand_conf", "n_frames"):
        assert np.array_equal(out[k], g[f"{tag}_{k}"]), k
    if dif:
        assert not out["input_kp"][:, :, 1, :].any()      # ChestDifference: row 1 becomes exactly 0


@pytest.mark.parametrize("tag", ["h5short", "h5long"])
def test_h5_path_matches_reference_bit_exact(tag):
    g = load_golden("preprocess.npz")
    T = int(g["T"])
    arr = g[f"{tag}_array"]
    item = oracle.array2item(arr)
    n = arr.shape[0]
    idx = oracle.window_frame_index(0, n, T, oracle.PAD_ZEROS)
    item = {k: oracle.gather_frames(np.ascontiguousarray(v), idx).astype(np.float32) for k, v in item.items()}
    item = oracle.apply_transforms(item, dif_encoding=True)
    assert oracle.n_frames_of(n, T) == int(g[f"{tag}_n_frames"])
    for k in ("input_kp", "input_conf", "target_kp", "target_conf", "left_hand_kp", "left_hand_conf"):
        assert np.array_equal(item[k], g[f"{tag}_{k}"]), k


def test_windowing_matches_reference():
    cases = load_golden("windowing.npz")["cases"]
    for n_total, n, sel, draw, start, first, last, count in cases:
        s, e = oracle.select_window(int(n_total), int(n), "first" if sel == 0 else "randomcrop", int(draw))
        assert (s, e - 1, e - s) == (int(first), int(last), int(count))
        assert s == int(start)


def test_adam_written_out_matches_torch():
    rng = np.random.default_rng(0)
    p = rng.normal(size=1000); g1 = rng.normal(size=1000)
    tp = torch.tensor(p, dtype=torch.float64, requires_grad=True)
    opt = torch.optim.Adam([tp], lr=2e-4)
    m = np.zeros_like(p); v = np.zeros_like(p); q = p.copy()
    for step in range(1, 4):
        tp.grad = torch.tensor(g1 * step)
        opt.step()
        q, m, v = oracle.adam_reference_step(q, g1 * step, m, v, step)
    assert np.allclose(q, tp.detach().numpy(), rtol=0, atol=1e-12)


def test_writers_match_reference():
    """the writer layouts and the loss-to-pixels scalar against outputs of the reference's own functions."""
    g = load_golden("writers.npz")
    pred = g["pred"]
    for t in range(pred.shape[0]):
        assert np.array_equal(oracle.array2open_pose(pred[t]).astype(np.float64), g["openpose"][t])
    assert np.array_equal(oracle.order_and_reshape_toh5(pred), g["h5"])
    assert oracle.l1_to_pixels(float(g["l1"]), 21, 1280) == float(g["l1_pixels_21"])
    assert oracle.l1_to_pixels(float(g["l1"]), 4, 1280) == float(g["l1_pixels_4"])

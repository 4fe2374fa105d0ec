"""GPU tier: loss, gradients, Adam and the fused train step through the C-ABI vs the oracle / golden vectors."""
import numpy as np
import pytest
import torch

import b2h_oracle as oracle
import hand_pose_sl_b200 as b2h
from conftest import golden_sd, load_golden
from hand_pose_sl_b200 import _lib, synthetic

pytestmark = pytest.mark.gpu
TOL = {"fp32": 1e-4, "fp32-ffma": 1e-4, "bf16": 2e-2}          # BASELINE.json north_star: predicted keypoints and loss
# Gradients in bf16 mode: the L1 criterion's gradient is sign(pred - target); rounding the WEIGHTS to bf16 moves
# the prediction by ~3e-3 and flips the sign of the residuals that are that close to zero, which alone puts an
# ideal bf16-operand implementation at 2.3e-2 (conv4.weight, confL1) against the fp32 reference (CPU emulation:
# only un-rounding the weights brings it to 1e-3; un-rounding activations / dY / dZ does not).  5e-2 bounds it.
GTOL = {"fp32": 1e-4, "fp32-ffma": 1e-4, "bf16": 5e-2}
DEV = "cuda:0"
NAMES = ["conv1.weight", "conv1.bias", "conv2.weight", "conv2.bias", "conv3.weight", "conv3.bias", "conv4.weight", "conv4.bias"]


def _grads_close(v, ref, prec, tol):
    """Gradient check.  bf16 / fp32-ffma: max-norm relative error <= tol.  fp32 = split-operand tensor-core mode: every GEMM
    operand carries bf16 high + low halves (16-17 significant bits, ~6e-6 relative on an activation), so a ReLU or an
    L1-sign decision that the fp32 reference takes within that distance of zero can fall the other way; ONE flipped ReLU
    moves ONE output channel's gradient by one frame's contribution (~1e-3 of its scale at 3 windows).  The check is
    therefore: max-norm error <= tol, or -- a handful of flips -- relative L1 error <= 5e-4 with every element inside 1e-2 (a
    missing operand term or an indexing bug shows as >= 1e-2 across the tensor).  Seen on convmodel_c30_t200 (window 3, conv1
    channel 3) and with 1443 frames at T = 481; fp32-ffma, exact to ~1e-7, agrees with the reference there to 5e-7."""
    err = oracle.rel_err(v, ref)
    if err <= tol:
        return True
    if prec != "fp32":
        return False
    rel_l1 = float(np.abs(v - ref).sum() / np.abs(ref).sum())
    return rel_l1 <= 5e-4 and err <= 1e-2                # a few channels off by ~1e-3 of their scale: tiny in L1, bounded in max


def _model(sd, C, pe, prec):
    m = b2h.ConvModel(C, "ReLU", pe, precision=prec)
    m.load_state_dict(sd)
    return m.to(DEV)


def _batch(g):
    return {"input_kp": torch.from_numpy(g["input_kp"]).to(DEV), "target_kp": torch.from_numpy(g["target_kp"]).to(DEV),
            "target_conf": torch.from_numpy(g["target_conf"]).to(DEV), "n_frames": torch.from_numpy(g["lengths"])}


def _split(model, flat):
    out, off = {}, 0
    for k, p in zip(NAMES, model._ordered_params()):
        out[k] = flat[off:off + p.numel()].view(p.shape).cpu().numpy()
        off += p.numel()
    return out


@pytest.mark.parametrize("prec", ["fp32", "fp32-ffma", "bf16"])
@pytest.mark.parametrize("kind", ["L1", "confL1"])
@pytest.mark.parametrize("name", ["convmodel_c30.npz", "convmodel_c30_b1.npz", "convmodel_c30_posemb.npz", "convmodel_c64.npz",
                                  "convmodel_c30_t200.npz"])
def test_loss_and_gradients_golden(name, kind, prec):
    g = load_golden(name)
    m = _model(golden_sd(g), int(g["C"]), bool(g["pos_emb"]), prec)
    loss, grads, pred = b2h.forward_backward(m, _batch(g), loss=kind, want_pred=True)
    torch.cuda.synchronize()
    assert abs(float(loss) - g[f"loss_{kind}"][0]) <= TOL[prec] * abs(g[f"loss_{kind}"][0])
    assert oracle.rel_err(pred.cpu().numpy(), g["pred_masked"]) <= TOL[prec]
    for k, v in _split(m, grads).items():
        assert _grads_close(v, g[f"grad_{kind}_" + k.replace(".", "_")], prec, GTOL[prec]), k
    if prec == "bf16" and int(g["C"]) <= 32 and not bool(g["pos_emb"]):
        # tensor-core kernel vs the ideal bf16-operand computation (tight: separates kernel bugs from bf16 rounding)
        e_loss, e_g, e_pred = oracle.train_grads_bf16_emulated(golden_sd(g), torch.from_numpy(g["input_kp"]),
                                                               torch.from_numpy(g["target_kp"]), g["lengths"], kind,
                                                               torch.from_numpy(g["target_conf"]))
        assert abs(float(loss) - e_loss) <= 1e-4 * abs(e_loss)
        for k, v in _split(m, grads).items():
            assert oracle.rel_err(v, e_g[k].numpy()) <= 1e-2, k


@pytest.mark.parametrize("prec", ["fp32", "fp32-ffma", "bf16"])
@pytest.mark.parametrize("kind", ["L1", "confL1"])
@pytest.mark.parametrize("name", ["convmodel_c30.npz", "convmodel_c64.npz", "convmodel_c30_t200.npz"])
def test_fused_train_steps_golden(name, kind, prec):
    """k fused steps (fwd+mask+loss+bwd | reduce+Adam+repack) == k reference steps (traintest.py:94-121)."""
    g = load_golden(name)
    sd = golden_sd(g)
    m = _model(sd, int(g["C"]), bool(g["pos_emb"]), prec)
    opt = b2h.FusedAdam(m.parameters(), lr=float(g["lr"]))
    batch = _batch(g)
    steps = len(g[f"loss_{kind}"])
    # oracle trajectory for the conditioning mask (see oracle.adam_conditioned)
    st = oracle.TrainState(sd, lr=float(g["lr"]), pos_emb=bool(g["pos_emb"]))
    all_grads = []
    for s in range(steps):
        loss = b2h.fused_train_step(m, batch, opt, loss=kind)
        assert abs(float(loss) - g[f"loss_{kind}"][s]) <= TOL[prec] * abs(g[f"loss_{kind}"][s]), s
        _, gr = oracle.train_step(st, torch.from_numpy(g["input_kp"]), torch.from_numpy(g["target_kp"]),
                                  torch.from_numpy(g["lengths"]), kind, torch.from_numpy(g["target_conf"]))
        all_grads.append({k: v.numpy() for k, v in gr.items()})
    masks = oracle.adam_conditioned(all_grads, sd, float(g["lr"]))
    for k, v in m.state_dict().items():
        want = g[f"w{steps}_{kind}_" + k.replace(".", "_")]
        d = np.abs(v.cpu().numpy() - want)
        viol = d[masks[k]] > TOL[prec] * np.abs(want).max()
        # fp32 split mode: a flipped ReLU decision (see _grads_close) changes one channel's gradients; Adam's first steps
        # turn a changed SIGN of a small gradient into 2*lr -> allow <= 1 % of the conditioned elements, bounded below
        assert viol.sum() == 0 or (prec == "fp32" and viol.mean() <= 0.01), k
        assert d.max() <= 2 * float(g["lr"]) * steps, k
    # the re-packed operand layouts the Adam kernel wrote == a fresh pack of the new weights
    packed_by_adam = m._packed.clone()
    m.mark_packed_stale()
    assert torch.equal(m.packed_weights(), packed_by_adam)
    # optimiser state in torch layout
    osd = opt.state_dict()
    assert set(osd["state"][0].keys()) >= {"step", "exp_avg", "exp_avg_sq"} and float(osd["state"][0]["step"]) == steps


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_modular_autograd_path_matches_reference_loop(prec):
    """The reference's own call sequence, unchanged: model(x) -> mask_output -> criterion -> zero_grad ->
    backward -> optimizer.step (traintest.py:94-121), with the drop-in classes."""
    g = load_golden("convmodel_c30.npz")
    sd = golden_sd(g)
    m = _model(sd, 30, False, prec)
    opt = b2h.FusedAdam(m.parameters(), lr=float(g["lr"]))
    crit = b2h.maskedPoseL1()
    batch = _batch(g)
    for s in range(3):
        pred = m(batch["input_kp"])
        pred = b2h.mask_output(pred, batch["n_frames"])
        loss = crit(pred, batch["target_kp"], batch["n_frames"])
        opt.zero_grad()
        loss.backward()
        if s == 0:
            for k, p in zip(NAMES, m._ordered_params()):
                assert oracle.rel_err(p.grad.cpu().numpy(), g["grad_L1_" + k.replace(".", "_")]) <= GTOL[prec], k
        opt.step()
        assert abs(loss.item() - g["loss_L1"][s]) <= TOL[prec] * abs(g["loss_L1"][s])


def test_stock_torch_adam_also_works_and_repacks():
    g = load_golden("convmodel_c30.npz")
    m = _model(golden_sd(g), 30, False, "fp32")
    opt = torch.optim.Adam(m.parameters(), lr=float(g["lr"]))          # the reference's optimiser, verbatim
    batch = _batch(g)
    for s in range(3):
        pred = b2h.mask_output(m(batch["input_kp"]), batch["n_frames"])
        loss = b2h.maskedPoseL1()(pred, batch["target_kp"], batch["n_frames"])
        opt.zero_grad(); loss.backward(); opt.step()
        assert abs(loss.item() - g["loss_L1"][s]) <= 1e-4 * abs(g["loss_L1"][s])


@pytest.mark.parametrize("kind", ["L1", "confL1"])
def test_criteria_standalone(kind):
    g = load_golden("convmodel_c30.npz")
    pred = torch.from_numpy(g["pred_masked"]).to(DEV).requires_grad_(True)
    tgt, conf, ln = torch.from_numpy(g["target_kp"]), torch.from_numpy(g["target_conf"]), torch.from_numpy(g["lengths"])
    cp = torch.from_numpy(g["pred_masked"]).requires_grad_(True)
    if kind == "L1":
        loss = b2h.maskedPoseL1()(pred, tgt.to(DEV), ln)
        ref = oracle.masked_pose_l1(cp, tgt, ln)
    else:
        loss = b2h.poderatedPoseL1()(pred, tgt.to(DEV), ln, conf)       # scores arrive on the CPU (utils.py:439)
        ref = oracle.poderated_pose_l1(cp, tgt, ln, conf)
    loss.backward(); ref.backward()
    assert abs(loss.item() - ref.item()) <= 2e-6 * abs(ref.item())
    assert oracle.rel_err(pred.grad.cpu().numpy(), cp.grad.numpy()) <= 1e-5


def test_mask_output_in_place_bit_exact():
    g = load_golden("convmodel_c30.npz")
    y = torch.from_numpy(g["pred"]).to(DEV)
    out = b2h.mask_output(y, torch.from_numpy(g["lengths"]))
    assert out.data_ptr() == y.data_ptr()                                # in place, returns the same tensor
    assert np.array_equal(out.cpu().numpy(), g["pred_masked"])


def test_fused_adam_vs_torch_adam_many_steps():
    torch.manual_seed(1)
    m = b2h.ConvModel(30, "ReLU", False).to(DEV)
    ref_p = [p.detach().clone().cpu().requires_grad_(True) for p in m._ordered_params()]
    ref = torch.optim.Adam(ref_p, lr=1e-3)
    opt = b2h.FusedAdam(m.parameters(), lr=1e-3)
    gen = torch.Generator().manual_seed(2)
    for s in range(10):
        for p, q in zip(m._ordered_params(), ref_p):
            gr = torch.randn(q.shape, generator=gen)
            q.grad = gr.clone(); p.grad = gr.to(DEV)
        opt.step(); ref.step()
    for p, q in zip(m._ordered_params(), ref_p):
        assert oracle.rel_err(p.detach().cpu().numpy(), q.detach().numpy()) <= 1e-6


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_config3_full_size_properties(prec):
    """BASELINE config 3 (B=256 x 64): the batch gradient is the mean of the two half-batch gradients (the same
    identity data-parallel training relies on), determinism, loss == criterion on the masked
    forward, and parity on a 16-window subset against the oracle."""
    sd = oracle.init_params(30, False, seed=0)
    batch = synthetic.model_batch(256, 64, seed=1234, ragged=True)
    m = _model(sd, 30, False, prec)
    db = {k: (v.to(DEV) if k != "n_frames" else v) for k, v in batch.items()}
    loss, grads = b2h.forward_backward(m, db)
    loss2, grads2 = b2h.forward_backward(m, db)
    assert torch.equal(grads, grads2) and torch.equal(loss, loss2)       # deterministic reduction
    halves = []
    for lo in (0, 128):
        hb = {k: v[lo:lo + 128] for k, v in db.items()}
        halves.append(b2h.forward_backward(m, hb))
    gmean = (halves[0][1] + halves[1][1]) / 2
    assert oracle.rel_err(grads.cpu().numpy(), gmean.cpu().numpy()) <= 1e-5     # same kernel, same rounding: tight
    assert abs(float(loss) - float((halves[0][0] + halves[1][0]) / 2)) <= 1e-6 * abs(float(loss))
    assert abs(float(loss) - float(b2h.validate_batch(m, db))) <= TOL[prec] * abs(float(loss))
    sub = {k: v[:16] for k, v in batch.items()}
    st = oracle.TrainState(sd)
    ref_loss, ref_g = oracle.train_step(st, sub["input_kp"], sub["target_kp"], sub["n_frames"])
    l16, g16 = b2h.forward_backward(m, {k: (v[:16].to(DEV) if k != "n_frames" else v[:16]) for k, v in batch.items()})
    assert abs(float(l16) - ref_loss) <= TOL[prec] * abs(ref_loss)
    for k, v in _split(m, g16).items():
        assert oracle.rel_err(v, ref_g[k].numpy()) <= GTOL[prec], k


@pytest.mark.parametrize("graph", [False, True])
@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_train_step_runner_matches_golden(prec, graph):
    """The steady-state runner (static buffers, device-side step counter; in bf16 mode ONE cooperative launch per step
    with the reduction + Adam in the kernel tail; optionally replayed from a CUDA graph) == the reference steps."""
    from hand_pose_sl_b200.runner import TrainStepRunner
    g = load_golden("convmodel_c30.npz")
    sd = golden_sd(g)
    m = _model(sd, 30, False, prec)
    opt = b2h.FusedAdam(m.parameters(), lr=float(g["lr"]))
    B, T = g["input_kp"].shape[:2]
    r = TrainStepRunner(m, opt, B, T, "L1", n_slots=1)
    r.load({"input_kp": torch.from_numpy(g["input_kp"]), "target_kp": torch.from_numpy(g["target_kp"]),
            "n_frames": torch.from_numpy(g["lengths"])}, non_blocking=False)
    steps = len(g["loss_L1"])
    losses = []
    if graph:
        r.capture(1)
        for s in range(steps):
            r.replay()
            losses.append(float(r.loss[0]))
    else:
        for s in range(steps):
            losses.append(float(r.step(0)))
    r.finish()
    torch.cuda.synchronize()
    assert _lib.load().b2h_tc_status() == 0
    for s in range(steps):
        assert abs(losses[s] - g["loss_L1"][s]) <= TOL[prec] * abs(g["loss_L1"][s]), (s, losses)
    st = oracle.TrainState(sd, lr=float(g["lr"]))
    all_grads = []
    for s in range(steps):
        _, gr = oracle.train_step(st, torch.from_numpy(g["input_kp"]), torch.from_numpy(g["target_kp"]), torch.from_numpy(g["lengths"]))
        all_grads.append({k: v.numpy() for k, v in gr.items()})
    masks = oracle.adam_conditioned(all_grads, sd, float(g["lr"]))
    for k, v in m.state_dict().items():
        want = g[f"w{steps}_L1_" + k.replace(".", "_")]
        d = np.abs(v.cpu().numpy() - want)
        assert d[masks[k]].max() <= TOL[prec] * np.abs(want).max(), k
        assert d.max() <= 2 * float(g["lr"]) * steps, k
    assert float(opt.state_dict()["state"][0]["step"]) == steps
    packed_by_kernel = m._packed.clone()
    m.mark_packed_stale()
    assert torch.equal(m.packed_weights(), packed_by_kernel)


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
@pytest.mark.parametrize("kind", ["L1", "confL1"])
@pytest.mark.parametrize("B,T,C", [(5, 37, 30), (3, 9, 16), (6, 101, 24), (3, 128, 30), (4, 129, 30), (3, 200, 30),
                                   (2, 255, 24), (2, 256, 30), (160, 200, 30), (2, 300, 30), (1, 700, 24), (3, 481, 30)])
def test_loss_and_gradients_odd_shapes_vs_oracle(B, T, C, kind, prec):
    """Shapes outside the golden files (odd T -> non-bulk target path, several windows per row segment, T > 64 ->
    one 128-row segment per tile, T > 128 -> two MMA tiles per segment (bf16) / overlapping sub-windows (fp32), T > 256 ->
    sub-windows in both modes, more windows than CTAs, C < 32) against the oracle's literal train step."""
    sd = oracle.init_params(C, False, seed=B + T)
    batch = synthetic.model_batch(B, T, seed=7 * B + T, ragged=True, len_seed=T)
    m = _model(sd, C, False, prec)
    db = {k: (v.to(DEV) if k != "n_frames" else v) for k, v in batch.items()}
    loss, grads, pred = b2h.forward_backward(m, db, loss=kind, want_pred=True)
    torch.cuda.synchronize()
    assert _lib.load().b2h_tc_status() == 0
    st = oracle.TrainState(sd)
    ref_loss, ref_g = oracle.train_step(st, batch["input_kp"], batch["target_kp"], batch["n_frames"], kind, batch["target_conf"])
    ref_pred = oracle.mask_output(oracle.conv_model_forward(sd, batch["input_kp"]).contiguous().clone(), batch["n_frames"])
    assert abs(float(loss) - ref_loss) <= TOL[prec] * abs(ref_loss)
    assert oracle.rel_err(pred.cpu().numpy(), ref_pred.detach().numpy()) <= TOL[prec]
    if prec.startswith("fp32"):
        # 32000 frames: a handful of the 1.3M residuals sit within fp32 accumulation noise of zero, and each sign flip
        # moves a gradient element by 2/n_el (the criterion is L1) -- widen the bound for the large case only
        gtol = GTOL[prec] if B * T < 10000 else 1e-3
        for k, v in _split(m, grads).items():
            assert _grads_close(v, ref_g[k].numpy(), prec, gtol), k
    else:
        # bf16 mode: with few frames a single flipped sign(pred - target) moves a gradient element by percents, so the
        # kernel is checked against the IDEAL bf16-operand computation (oracle.train_grads_bf16_emulated: reference
        # formulas, GEMM operands rounded to bf16, fp32 everywhere else) -- kernel bugs show, bf16 rounding does not.
        e_loss, e_g, e_pred = oracle.train_grads_bf16_emulated(sd, batch["input_kp"], batch["target_kp"], batch["n_frames"],
                                                               kind, batch["target_conf"])
        assert abs(float(loss) - e_loss) <= 1e-4 * abs(e_loss)
        # the tensor core's fp32 accumulation order differs from the emulation's, so an activation that lands on a bf16
        # rounding boundary may round the other way (1 bf16 ulp of one activation ~ 1e-3 of the prediction scale)
        assert oracle.rel_err(pred.cpu().numpy(), e_pred.numpy()) <= 2e-3
        for k, v in _split(m, grads).items():
            assert oracle.rel_err(v, e_g[k].numpy()) <= 1e-2, k


def test_runners_of_different_shapes_share_one_workspace():
    """the fused kernel's grid-barrier words sit in the workspace HEADER (fixed offset), so a small-batch
    runner used after a large-batch one on the same model (remainder batch) never finds stale gradient partials where
    it expects barrier state.  Large -> small -> large, losses must equal fresh single-runner results."""
    from hand_pose_sl_b200.runner import TrainStepRunner
    sd = oracle.init_params(30, False, seed=3)
    big = synthetic.model_batch(256, 64, seed=11, ragged=True)
    small = synthetic.model_batch(6, 64, seed=12, ragged=True)

    def run(order):
        m = _model(sd, 30, False, "bf16")
        opt = b2h.FusedAdam(m.parameters(), lr=2e-4)
        runners = {"big": TrainStepRunner(m, opt, 256, 64, "L1"), "small": TrainStepRunner(m, opt, 6, 64, "L1")}
        assert runners["big"].ws.data_ptr() == runners["small"].ws.data_ptr()      # one shared per-model workspace
        runners["big"].load(big, non_blocking=False); runners["small"].load(small, non_blocking=False)
        out = [float(runners[k].step(0)) for k in order]
        torch.cuda.synchronize()
        assert _lib.load().b2h_tc_status() == 0
        return out, m.flat_parameters().clone()

    l1, w1 = run(["big", "small", "big", "small"])
    l2, w2 = run(["big", "small", "big", "small"])
    assert l1 == l2 and torch.equal(w1, w2)                                         # deterministic, no stale state
    # each step's loss equals the oracle's on the same trajectory
    st = oracle.TrainState(sd, lr=2e-4)
    for i, k in enumerate(["big", "small", "big", "small"]):
        b = big if k == "big" else small
        ref_loss, _ = oracle.train_step(st, b["input_kp"], b["target_kp"], b["n_frames"])
        assert abs(l1[i] - ref_loss) <= TOL["bf16"] * abs(ref_loss), (i, l1, ref_loss)


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_captured_graph_follows_learning_rate_changes(prec):
    """the learning rate lives in device memory, so adjust_learning_rate (traintest.py:83-84) takes effect
    on a captured graph without a re-capture: graph steps with a mid-way lr change == eager steps with the same change."""
    from hand_pose_sl_b200.runner import TrainStepRunner
    g = load_golden("convmodel_c30.npz")
    sd = golden_sd(g)
    batch = {"input_kp": torch.from_numpy(g["input_kp"]), "target_kp": torch.from_numpy(g["target_kp"]),
             "n_frames": torch.from_numpy(g["lengths"])}
    B, T = g["input_kp"].shape[:2]
    results = []
    for graph in (False, True):
        m = _model(sd, 30, False, prec)
        opt = b2h.FusedAdam(m.parameters(), lr=1e-3)
        r = TrainStepRunner(m, opt, B, T, "L1")
        r.load(batch, non_blocking=False)
        if graph:
            r.capture(1)
        for s in range(4):
            if s == 2:
                b2h.adjust_learning_rate(1e-3, 1, opt, 2)          # lr = 1e-5 from step 3 on
            r.replay() if graph else r.step(0)
        r.finish()
        results.append(m.flat_parameters().clone())
    assert torch.equal(results[0], results[1])
    # and the change really took effect: two more steps at lr=1e-3 would have moved the weights further
    m = _model(sd, 30, False, prec)
    opt = b2h.FusedAdam(m.parameters(), lr=1e-3)
    r = TrainStepRunner(m, opt, B, T, "L1")
    r.load(batch, non_blocking=False)
    for s in range(4):
        r.step(0)
    assert not torch.equal(m.flat_parameters(), results[0])


def test_forward_runner_follows_weight_changes():
    """ForwardRunner re-packs when the parameters' version counters moved (load_state_dict after
    construction), also for a captured graph (the packed buffer keeps its address)."""
    from hand_pose_sl_b200.runner import ForwardRunner
    sd_a, sd_b = oracle.init_params(30, False, seed=1), oracle.init_params(30, False, seed=2)
    batch = synthetic.model_batch(4, 64, seed=5)
    m = _model(sd_a, 30, False, "bf16")
    fr = ForwardRunner(m, 4, 64)
    fr.x[0].copy_(batch["input_kp"])
    ya = fr.run(0).clone()
    fr.capture(1)
    m.load_state_dict(sd_b)
    yb_graph = None
    fr.replay(); yb_graph = fr.y[0].clone()
    yb = fr.run(0).clone()
    ref_b = oracle.conv_model_forward(sd_b, batch["input_kp"]).contiguous().numpy()
    assert oracle.rel_err(yb.cpu().numpy(), ref_b) <= TOL["bf16"] and torch.equal(yb, yb_graph)
    assert not torch.equal(ya, yb)


def test_staged_single_copy_inputs_are_bit_identical():
    """runner.host_stage + load_staged (ONE H2D copy, bf16 keypoint input) == load(batch) with fp32 inputs in bf16 mode."""
    from hand_pose_sl_b200.runner import TrainStepRunner
    sd = oracle.init_params(30, False, seed=4)
    batch = synthetic.model_batch(32, 64, seed=21, ragged=True)
    out = []
    for staged in (False, True):
        m = _model(sd, 30, False, "bf16")
        opt = b2h.FusedAdam(m.parameters(), lr=2e-4)
        r = TrainStepRunner(m, opt, 32, 64, "confL1", n_slots=2, x_dtype=torch.bfloat16 if staged else None)
        if staged:
            r.load_staged(r.host_stage(batch), slot=1, non_blocking=False)
        else:
            r.load(batch, slot=1, non_blocking=False)
        losses = [float(r.step(1)) for _ in range(3)]
        r.finish()
        out.append((losses, m.flat_parameters().clone()))
    assert out[0][0] == out[1][0] and torch.equal(out[0][1], out[1][1])
    # the loss written straight into pinned host memory by the kernel (to_host=True) == the device-side loss, and the
    # pipelined loop that uses it yields the same trajectory
    from hand_pose_sl_b200.runner import pipelined_steps
    m = _model(sd, 30, False, "bf16")
    opt = b2h.FusedAdam(m.parameters(), lr=2e-4)
    r = TrainStepRunner(m, opt, 32, 64, "confL1", n_slots=2, x_dtype=torch.bfloat16)
    st = r.host_stage(batch)
    piped = list(pipelined_steps(r, [st, st, st]))
    r.finish()
    assert piped == out[1][0] and torch.equal(m.flat_parameters(), out[1][1])


def test_fp32_mode_split_kernel_matches_ffma_arbiter():
    """fp32 mode on the tensor pipe (bf16 high/low operand pairs, 3 MMAs per product, fp32 accumulation) against the FFMA
    kernel on the same inputs: prediction / loss / gradients agree far inside the 1e-4 budget (the split carries ~16
    mantissa bits per operand: ~6e-6 on the prediction), and the kernel choice says which one ran."""
    lib = _lib.load()
    assert lib.b2h_kernel_choice(64, 24, 30, 0, _lib.FP32, 1) == 2 and lib.b2h_kernel_choice(64, 24, 30, 0, _lib.FP32_FFMA, 1) == 1
    sd = oracle.init_params(30, False, seed=5)
    for B, T in ((64, 64), (5, 37), (3, 128)):
        batch = synthetic.model_batch(B, T, seed=31 + T, ragged=True)
        db = {k: (v.to(DEV) if k != "n_frames" else v) for k, v in batch.items()}
        outs = {}
        for prec in ("fp32", "fp32-ffma"):
            m = _model(sd, 30, False, prec)
            outs[prec] = b2h.forward_backward(m, db, loss="confL1", want_pred=True)
        torch.cuda.synchronize()
        assert lib.b2h_tc_status() == 0
        (l1, g1, p1), (l2, g2, p2) = outs["fp32"], outs["fp32-ffma"]
        assert abs(float(l1) - float(l2)) <= 2e-5 * abs(float(l2))
        assert oracle.rel_err(p1.cpu().numpy(), p2.cpu().numpy()) <= 3e-5
        assert oracle.rel_err(g1.cpu().numpy(), g2.cpu().numpy()) <= 1e-4


@pytest.mark.parametrize("kind", ["L1", "confL1"])
@pytest.mark.parametrize("B,T,C", [(5, 64, 64), (7, 64, 128), (4, 64, 256), (3, 126, 256), (2, 200, 96), (9, 33, 80), (40, 64, 256)])
def test_wide_training_on_tensor_cores_vs_oracle(B, T, C, kind):
    """Wide models (32 < conv_channels <= 256, bf16 mode) train on tcgen05: streamed-weight forward that saves the layer
    inputs + criterion, dgrad chain with the transposed blocks, split-K weight-gradient GEMMs -- against the oracle's
    literal train step (loss, masked prediction at 2e-2) and against the ideal bf16-operand computation (gradients)."""
    lib = _lib.load()
    assert lib.b2h_kernel_choice(T, 24, C, 0, _lib.BF16, 1) == 5
    sd = oracle.init_params(C, False, seed=B + T)
    batch = synthetic.model_batch(B, T, seed=7 * B + T, ragged=True, len_seed=T)
    m = _model(sd, C, False, "bf16")
    db = {k: (v.to(DEV) if k != "n_frames" else v) for k, v in batch.items()}
    loss, grads, pred = b2h.forward_backward(m, db, loss=kind, want_pred=True)
    loss2, grads2 = b2h.forward_backward(m, db, loss=kind)
    torch.cuda.synchronize()
    assert lib.b2h_tc_status() == 0
    assert torch.equal(grads, grads2) and torch.equal(loss, loss2)                           # deterministic split-K reduction
    st = oracle.TrainState(sd)
    ref_loss, ref_g = oracle.train_step(st, batch["input_kp"], batch["target_kp"], batch["n_frames"], kind, batch["target_conf"])
    ref_pred = oracle.mask_output(oracle.conv_model_forward(sd, batch["input_kp"]).contiguous().clone(), batch["n_frames"])
    assert abs(float(loss) - ref_loss) <= TOL["bf16"] * abs(ref_loss)
    assert oracle.rel_err(pred.cpu().numpy(), ref_pred.detach().numpy()) <= TOL["bf16"]
    e_loss, e_g, e_pred = oracle.train_grads_bf16_emulated(sd, batch["input_kp"], batch["target_kp"], batch["n_frames"],
                                                           kind, batch["target_conf"])
    assert abs(float(loss) - e_loss) <= 2e-4 * abs(e_loss)
    assert oracle.rel_err(pred.cpu().numpy(), e_pred.numpy()) <= 3e-3
    for k, v in _split(m, grads).items():
        assert oracle.rel_err(v, e_g[k].numpy()) <= 1e-2, k


def test_wide_training_fused_steps_follow_the_oracle():
    """k fused train steps at conv_channels = 128 (wide kernels + reduce/Adam/re-pack) == k reference steps."""
    C, B, T, steps = 128, 6, 64, 3
    sd = oracle.init_params(C, False, seed=9)
    batch = synthetic.model_batch(B, T, seed=41, ragged=True)
    m = _model(sd, C, False, "bf16")
    opt = b2h.FusedAdam(m.parameters(), lr=2e-4)
    db = {k: (v.to(DEV) if k != "n_frames" else v) for k, v in batch.items()}
    st = oracle.TrainState(sd, lr=2e-4)
    all_grads = []
    for s in range(steps):
        loss = b2h.fused_train_step(m, db, opt)
        ref_loss, gr = oracle.train_step(st, batch["input_kp"], batch["target_kp"], batch["n_frames"])
        all_grads.append({k: v.numpy() for k, v in gr.items()})
        assert abs(float(loss) - ref_loss) <= TOL["bf16"] * abs(ref_loss), s
    masks = oracle.adam_conditioned(all_grads, sd, 2e-4)
    want = st.state_dict()
    for k, v in m.state_dict().items():
        d = np.abs(v.cpu().numpy() - want[k].numpy())
        assert d[masks[k]].max() <= TOL["bf16"] * np.abs(want[k].numpy()).max(), k
        assert d.max() <= 2 * 2e-4 * steps, k
    packed_by_adam = m._packed.clone()
    m.mark_packed_stale()
    assert torch.equal(m.packed_weights(), packed_by_adam)
    # the runner (device-side step counter, CUDA graph) drives the same path
    from hand_pose_sl_b200.runner import TrainStepRunner
    m2 = _model(sd, C, False, "bf16")
    o2 = b2h.FusedAdam(m2.parameters(), lr=2e-4)
    r = TrainStepRunner(m2, o2, B, T, "L1")
    r.load(batch, non_blocking=False)
    r.capture(1)
    for s in range(steps):
        r.replay()
    r.finish()
    assert torch.equal(m2.flat_parameters(), m.flat_parameters())


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
@pytest.mark.parametrize("T,max_len", [(64, 100), (200, 100), (37, 37), (128, 250)])
def test_positional_embedding_beyond_100_frames(T, max_len, prec):
    """the positional row t / max_len for windows other than 100 frames (the reference's torch.cat only
    works for T == max_len).  Oracle = the reference formulas with the row built for the actual T: forward, loss and
    gradients (conv1 has 25 input channels, the row is channel 0)."""
    import torch.nn.functional as F
    B = 4
    sd = oracle.init_params(30, True, seed=T)
    batch = synthetic.model_batch(B, T, seed=3 * T, ragged=True, len_seed=T)
    m = b2h.ConvModel(30, "ReLU", True, precision=prec, pos_emb_max_len=max_len, pos_emb_any_length=True)
    m.load_state_dict(sd)
    m = m.to(DEV)
    db = {k: (v.to(DEV) if k != "n_frames" else v) for k, v in batch.items()}
    loss, grads, pred = b2h.forward_backward(m, db, want_pred=True)
    torch.cuda.synchronize()
    assert _lib.load().b2h_tc_status() == 0
    # reference formulas (HandPoseModels.py:40-84) with pe = arange(T) / max_len
    ps = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    x = batch["input_kp"].reshape(B, T, 24).permute(0, 2, 1)
    pe = (torch.arange(T).float() / max_len)[None, None, :].repeat(B, 1, 1)
    h = torch.cat([pe, x], dim=1)
    for l in range(1, 5):
        h = F.conv1d(h, ps[f"conv{l}.weight"], ps[f"conv{l}.bias"], padding=2)
        if l < 4:
            h = F.relu(h)
    out = h.view(B, 21, 2, T).permute(0, 3, 1, 2).contiguous()
    out = oracle.mask_output(out, batch["n_frames"])
    ref_loss = oracle.masked_pose_l1(out, batch["target_kp"], batch["n_frames"])
    ref_loss.backward()
    assert abs(float(loss) - float(ref_loss)) <= TOL[prec] * abs(float(ref_loss))
    assert oracle.rel_err(pred.cpu().numpy(), out.detach().numpy()) <= TOL[prec]
    if prec == "fp32":
        # 4 windows: one residual within fp32 noise of zero flips sign(pred - target) and moves a gradient element by
        # 2 / (B * len * 42) ~ 1e-4 of its scale -> percent-level on the smallest rows; everything else agrees to ~1e-6
        for k, v in _split(m, grads).items():
            assert oracle.rel_err(v, ps[k].grad.numpy()) <= 1e-2, k
            assert np.median(np.abs(v - ps[k].grad.numpy())) <= 1e-5 * np.abs(ps[k].grad.numpy()).max(), k
    # the reference behaviour is kept unless asked otherwise
    strict = b2h.ConvModel(30, "ReLU", True, precision=prec).to(DEV)
    if T != 100:
        with pytest.raises(RuntimeError):
            strict(db["input_kp"])

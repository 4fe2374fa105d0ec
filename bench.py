#!/usr/bin/env python
"""bench.py -- body2hand frames/sec on B200 (contract: see the task statement / DESIGN.md §Measurement).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Headline workload (N GPUs): BASELINE.json config 3/4 -- the training step (forward + mask_output +
maskedPoseL1 + backward + Adam) on batch 256 x 64 frames per GPU, C=30, synthetic How2Sign-shaped
windows; weak scaling (per-GPU batch fixed) with one gradient all-reduce per step.  The same JSON line
also carries config 2 (forward, batch 512 x 64, bf16 tensor-core path) under "fwd" and the K0
preprocessing throughput under "preprocess" (N=1 only).  One JSON line on stdout (rank 0)."""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

C, T, B_TRAIN, B_FWD, LR = 30, 64, 256, 512, 2e-4
S_MAC = 24 * C + 2 * C * C + 42 * C                          # MAC per tap per frame
FWD_FLOP_PER_WINDOW = 2 * (5 * T - 6) * S_MAC                # 2 373 840 at T=64, C=30
TRAIN_FLOP_PER_WINDOW = 2 * (5 * T - 6) * (3 * S_MAC - 24 * C)   # 6 669 360
PRE_BYTES_PER_FRAME = 804 + 648                              # K0 figure of record
N_SLOTS = 40                                                 # resident batches rotated so the working set > L2


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "tflops": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "tflops_burst": d["bf16_tflops"], "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "tflops": 1400.0, "tflops_burst": 1590.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower() == "active"})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def cpu_reference_arm(steps, warmup, budget_s=25.0):
    """The reference's CPU path for the headline config: its ConvModel + mask_output + maskedPoseL1 + Adam
    (restated in oracle/b2h_oracle.py from traintest.py:87-123; the reference is pure Python/PyTorch, nothing
    to compile -> kind 'port'), all host threads, same seeded batch."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import b2h_oracle as oracle
    from hand_pose_sl_b200 import synthetic
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    batch = synthetic.model_batch(B_TRAIN, T, seed=1234)
    st = oracle.TrainState(oracle.init_params(C, False, seed=0), lr=LR)
    for _ in range(max(1, warmup)):
        oracle.train_step(st, batch["input_kp"], batch["target_kp"], batch["n_frames"])
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):
        oracle.train_step(st, batch["input_kp"], batch["target_kp"], batch["n_frames"])
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    # also the "loops vectorised" variant (closed-form masked L1, one multiply for mask_output), so the
    # reader can see how much of the CPU number is the reference's per-sample Python loops
    vec_ms = None
    try:
        lens = torch.as_tensor(batch["n_frames"], dtype=torch.int64)
        keep = (torch.arange(T)[None, :] < lens[:, None]).float()[:, :, None, None]
        n_vec = max(3, min(done, 40))
        t1 = time.perf_counter()
        for _ in range(n_vec):
            pred = oracle.conv_model_forward(st.params, batch["input_kp"], st.pos_emb) * keep
            loss = oracle.masked_pose_l1_closed_form(pred, batch["target_kp"], batch["n_frames"])
            st.opt.zero_grad()
            loss.backward()
            st.opt.step()
        vec_ms = (time.perf_counter() - t1) / n_vec * 1e3
    except Exception:
        vec_ms = None
    model = ""
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                model = line.split(":", 1)[1].strip()
                break
    except Exception:
        pass
    return {"value": done * B_TRAIN * T / dt, "unit": "frames/s", "cores": cores, "kind": "port",
            "sample": f"{done} train steps of batch {B_TRAIN}x{T} (C={C}, fp32, reference per-sample loss loops) in {dt:.2f} s",
            "threads": torch.get_num_threads(), "cpu": model, "ms_per_step": dt / done * 1e3,
            "ms_per_step_loops_vectorised": vec_ms,
            "value_loops_vectorised": (B_TRAIN * T / (vec_ms * 1e-3)) if vec_ms else None}, done, dt


def run_reference(args, rank):
    if rank != 0:
        return
    base, done, dt = cpu_reference_arm(args.steps, args.warmup, budget_s=150.0)
    line = {"impl": "reference", "metric": "body2hand_train_frames_per_sec", "value": base["value"], "unit": "frames/s",
            "n_gpus": args.gpus, "steps": done, "warmup": args.warmup, "ms_per_step": dt / done * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"train step (fwd+mask+L1+bwd+Adam), batch {B_TRAIN}x{T} frames, C={C} (BASELINE config 3)",
                       "global_batch": B_TRAIN, "frames_per_window": T, "conv_channels": C},
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=40)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default="auto", choices=["auto", "fp32", "bf16"])
    ap.add_argument("--skip-extras", action="store_true", help="only the headline train-step number")
    ap.add_argument("--exchange", default="auto", choices=["auto", "p2p", "nccl"], help="multi-GPU gradient exchange")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    import hand_pose_sl_b200 as b2h
    from hand_pose_sl_b200 import _lib, synthetic
    from hand_pose_sl_b200.parallel import DataParallelTrainer
    from hand_pose_sl_b200.runner import ForwardRunner, TrainStepRunner

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: hand_pose_sl_b200 has no CPU path (use --impl reference for the CPU arm)")
    # libraries (NCCL's version banner, ...) write to stdout; keep fd 1 clean for the ONE JSON line
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ["NCCL_DEBUG"] = os.environ.get("B2H_NCCL_DEBUG", "WARN")    # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)
    pk = peaks()
    train_prec = "bf16" if args.precision == "auto" else args.precision
    fwd_prec = "bf16" if args.precision == "auto" else args.precision

    # ---------------- headline: training step, batch 256 x 64 per GPU ----------------
    torch.manual_seed(0)
    model = b2h.ConvModel(C, "ReLU", False, precision=train_prec).to(dev)
    opt = b2h.FusedAdam(model.parameters(), lr=LR)
    if world > 1:
        runner = DataParallelTrainer(model, opt, B_TRAIN, T, "L1", n_slots=N_SLOTS, exchange=args.exchange)
    else:
        runner = TrainStepRunner(model, opt, B_TRAIN, T, "L1", n_slots=N_SLOTS)
    host_batches = []
    for s in range(N_SLOTS):
        b = synthetic.model_batch(B_TRAIN, T, seed=1234 + 1000 * rank + s)
        host_batches.append({k: v.pin_memory() for k, v in b.items()})
        runner.load(host_batches[-1], slot=s, non_blocking=False)
    chunk = min(N_SLOTS, args.steps)
    graphed = True
    try:
        runner.capture(chunk)
    except Exception as e:                                    # e.g. NCCL capture refused: direct launches instead
        graphed = False
        sys.stderr.write(f"[bench] CUDA graph capture unavailable ({type(e).__name__}: {e}); direct launches\n")
        torch.cuda.synchronize()

    def run_steps(n):
        done = 0
        if graphed:
            while done + chunk <= n:
                runner.replay()
                done += chunk
        while done < n:
            runner.step(done % N_SLOTS)
            done += 1

    run_steps(args.warmup)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = _lib.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ev0.record()
    run_steps(args.steps)
    ev1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    clocks = sampler.stop() if rank == 0 else None
    torch.cuda.synchronize()
    c0 = _lib.launch_count()
    runner.step(0)                                            # count the kernels of ONE step (outside the timed region)
    torch.cuda.synchronize()
    launches_per_step = _lib.launch_count() - c0
    gpu_launches = args.steps * launches_per_step
    value = args.steps * B_TRAIN * T * world / (ms * 1e-3)
    final_loss = float(runner.loss[0].item())

    # ---------------- e2e: host buffers -> H2D -> step -> D2H loss, every step ----------------
    e2e_steps = min(args.steps, 100)
    h2d = sum(host_batches[0][k].numel() * host_batches[0][k].element_size() for k in ("input_kp", "target_kp")) + B_TRAIN * 4
    lengths32 = [hb["n_frames"].to(torch.int32).pin_memory() for hb in host_batches]
    for hb, l32 in zip(host_batches, lengths32):
        hb["n_frames"] = l32
    for i in range(5):
        runner.load(host_batches[i % N_SLOTS], slot=0)
        float(runner.step(0).item())
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        runner.load(host_batches[i % N_SLOTS], slot=0)        # pinned host -> device, inside the timed region
        loss_val = float(runner.step(0).item())               # device -> host read of the step's result
    torch.cuda.synchronize()
    e2e_dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_dt], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_dt = float(t.item())
    e2e = {"value": e2e_steps * B_TRAIN * T * world / e2e_dt, "unit": "frames/s", "h2d_bytes_per_step": h2d,
           "d2h_bytes_per_step": 4, "steps": e2e_steps, "api": "TrainStepRunner.load(pinned batch) + .step() + loss.item()",
           "mode": "sequential"}
    if world == 1:
        # Same bytes, same per-step loss read, but batch i+1 travels on a copy stream while step i computes
        # (runner.pipelined_steps: what a prefetching DataLoader gives the reference loop).  Single-process only: an
        # exception on one rank of a multi-rank run would desynchronise the collectives.  Guarded: on any failure the
        # sequential number above stands.
        try:
            from hand_pose_sl_b200.runner import pipelined_steps
            for _ in pipelined_steps(runner, (host_batches[i % N_SLOTS] for i in range(6))):
                pass
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            n_p = 0
            for loss_val in pipelined_steps(runner, (host_batches[i % N_SLOTS] for i in range(e2e_steps))):
                n_p += 1
            torch.cuda.synchronize()
            p_dt = time.perf_counter() - t0
            p_val = n_p * B_TRAIN * T / p_dt
            e2e["sequential_value"] = e2e["value"]
            e2e["pipelined_value"] = p_val
            if n_p == e2e_steps and loss_val == loss_val and p_val > e2e["value"]:      # finite loss, all steps ran
                e2e["value"] = p_val
                e2e["mode"] = "pipelined (double-buffered H2D on a copy stream)"
                e2e["api"] = "runner.pipelined_steps(TrainStepRunner, pinned batches): load + step + loss.item() per step"
        except Exception as ex:   # noqa: BLE001
            e2e["pipelined_error"] = str(ex)[:200]
    runner.finish()

    line = {"metric": "body2hand_train_frames_per_sec", "value": value, "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32" if train_prec == "fp32" else "bf16", "data": "synthetic",
            "config": {"workload": f"train step (fwd+mask+L1+bwd+Adam), batch {B_TRAIN}x{T} frames per GPU, C={C} (BASELINE config 3/4)",
                       "global_batch": B_TRAIN * world, "frames_per_window": T, "conv_channels": C,
                       "parallelism": f"dp{world}", "cuda_graph": graphed,
                       "grad_exchange": (getattr(runner, "exchange", None) if world > 1 else None),
                       "l2": f"{N_SLOTS} resident batches rotated ({N_SLOTS * (h2d) / 1e6:.0f} MB inputs + "
                             f"{_lib.workspace_bytes(B_TRAIN, T, 24, C, 0, _lib.PRECISIONS[train_prec]) / 1e6:.0f} MB gradient partials) > 126 MB L2"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(gpu_launches), "final_loss": final_loss}

    if rank == 0 and world == 1:
        # ---------------- roofline of the dominant kernel (fused fwd+loss+bwd), timed alone ----------------
        lib = _lib.load()
        n_in, Cc, pe = model._geometry()

        def train_kernel_only(slot):
            _lib.check(lib.b2h_train_forward_backward(
                _lib.ptr(runner.x[slot]), _lib.DT_F32, _lib.ptr(runner.target[slot]), None, _lib.ptr(runner.lengths[slot]),
                _lib.ptr(model._flat), _lib.ptr(runner.packed), None, None, None, B_TRAIN, T, n_in, Cc, pe, _lib.LOSS_L1,
                _lib.PRECISIONS[model.precision], None, _lib.ptr(runner.ws), runner.ws.numel(), _lib.stream_ptr(dev)))

        for i in range(10):
            train_kernel_only(i % N_SLOTS)
        torch.cuda.synchronize()
        kgraph = torch.cuda.CUDAGraph()                       # graph replay: device time, not Python launch rate
        with torch.cuda.graph(kgraph):
            for i in range(N_SLOTS):
                train_kernel_only(i)
        kgraph.replay()
        torch.cuda.synchronize()
        greps = 5
        reps = greps * N_SLOTS
        ev0.record()
        for _ in range(greps):
            kgraph.replay()
        ev1.record()
        torch.cuda.synchronize()
        k_only_ms = ev0.elapsed_time(ev1) / reps            # forward+loss+backward kernel without the fused tail
        fused = launches_per_step == 1                      # bf16 mode: the whole step IS one kernel launch
        k_ms = (ms / args.steps) if fused else k_only_ms
        tfl = TRAIN_FLOP_PER_WINDOW * B_TRAIN / (k_ms * 1e-3) / 1e12
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.isfile(tpath):
            traffic = json.load(open(tpath)).get("train_kernel_dram_bytes_per_launch")
        line["roofline"] = {"bound": "tensor", "achieved": tfl, "peak": pk["tflops"], "unit": "TFLOP/s", "frac": tfl / pk["tflops"],
                            "traffic": traffic, "kernel": "conv_fp32_kernel<train>" if train_prec == "fp32" else "conv_tc_tile_kernel<train>",
                            "kernel_ms": k_ms, "kernel_ms_fwd_bwd_only": k_only_ms, "algorithmic_flop_per_launch": TRAIN_FLOP_PER_WINDOW * B_TRAIN,
                            "peak_source": pk["source"] + ", bf16 dense sustained",
                            "note": "bf16 mode: one cooperative launch per step (fwd+loss+bwd, grid barrier, reduction, Adam), timed over the graph-replayed steps (inter-kernel gaps included); kernel_ms_fwd_bwd_only = the same kernel without its reduction/Adam tail; per-launch device times are in profiles/"}

        if not args.skip_extras:
            # ---------------- config 2: forward, batch 512 x 64, bf16 tensor-core path ----------------
            torch.manual_seed(0)
            fmodel = b2h.ConvModel(C, "ReLU", False, precision=fwd_prec).to(dev)
            fr = ForwardRunner(fmodel, B_FWD, T, n_slots=N_SLOTS, x_dtype=torch.bfloat16 if fwd_prec == "bf16" else torch.float32)
            for s in range(N_SLOTS):
                fr.x[s].copy_(synthetic.model_batch(B_FWD, T, seed=99 + s)["input_kp"])
            fr.capture(N_SLOTS)
            for _ in range(3):
                fr.graph.replay()
            torch.cuda.synchronize()
            freps = 10
            ev0.record()
            for _ in range(freps):
                fr.graph.replay()
            ev1.record()
            torch.cuda.synchronize()
            f_ms = ev0.elapsed_time(ev1) / (freps * N_SLOTS)
            f_tfl = FWD_FLOP_PER_WINDOW * B_FWD / (f_ms * 1e-3) / 1e12
            in_b = 2 if fwd_prec == "bf16" else 4
            f_bytes = B_FWD * T * (24 * in_b + 42 * 4)
            line["fwd"] = {"metric": "body2hand_forward_frames_per_sec", "value": B_FWD * T / (f_ms * 1e-3), "unit": "frames/s",
                           "ms_per_batch": f_ms, "dtype": fwd_prec, "workload": f"forward, batch {B_FWD}x{T} (BASELINE config 2), CUDA graph",
                           "roofline": {"bound": "tensor", "achieved": f_tfl, "peak": pk["tflops"], "unit": "TFLOP/s",
                                        "frac": f_tfl / pk["tflops"], "hbm_gbs": f_bytes / (f_ms * 1e-3) / 1e9,
                                        "hbm_frac": f_bytes / (f_ms * 1e-3) / 1e9 / pk["hbm_gbs"]},
                           "tc_status": int(lib.b2h_tc_status())}
            # ---------------- wide variant (C = 256 puts the convs on the tensor-pipe roofline) ----------------
            if fwd_prec == "bf16":
                CW = 256
                sw = 24 * CW + 2 * CW * CW + 42 * CW
                wide = {}
                for (bw, tw) in ((1776, 64), (1184, 126)):       # 4 tiles of 256 rows per SM: 3 x 64-frame / 2 x 126-frame windows per tile
                    torch.manual_seed(0)
                    wm = b2h.ConvModel(CW, "ReLU", False, precision="bf16").to(dev)
                    wr = ForwardRunner(wm, bw, tw, n_slots=4, x_dtype=torch.bfloat16)
                    for s in range(4):
                        wr.x[s].copy_(synthetic.model_batch(bw, tw, seed=300 + s)["input_kp"])
                    wr.capture(8)
                    for _ in range(3):
                        wr.graph.replay()
                    torch.cuda.synchronize()
                    ev0.record()
                    for _ in range(5):
                        wr.graph.replay()
                    ev1.record()
                    torch.cuda.synchronize()
                    w_ms = ev0.elapsed_time(ev1) / 40
                    w_tfl = 2 * (5 * tw - 6) * sw * bw / (w_ms * 1e-3) / 1e12
                    w_traffic = None
                    if (bw, tw) == (1184, 126) and os.path.isfile(tpath):
                        w_traffic = json.load(open(tpath)).get("wide_fwd_dram_bytes_per_launch")
                    wide[f"{bw}x{tw}"] = {"value": bw * tw / (w_ms * 1e-3), "unit": "frames/s", "ms_per_batch": w_ms,
                                          "roofline": {"bound": "tensor", "achieved": w_tfl, "peak": pk["tflops"], "unit": "TFLOP/s",
                                                       "frac": w_tfl / pk["tflops"], "traffic": w_traffic}}
                    del wr, wm
                line["fwd_wide"] = {"metric": "body2hand_forward_frames_per_sec", "dtype": "bf16", "conv_channels": CW,
                                    "workload": "forward, conv_channels=256 (streamed-weight tcgen05 kernel), bf16 inputs, CUDA graph of 8 launches over 4 resident batches",
                                    "batches": wide, "tc_status": int(lib.b2h_tc_status())}
            # ---------------- K0 preprocessing: 2 h of 30 fps frames ----------------
            F = 216000
            pose, lh, rh = synthetic.synthetic_clip(F, seed=1234)
            tp, tl, tr = (torch.from_numpy(a).to(dev) for a in (pose, lh, rh))
            pre = b2h.PreprocessRightHand()
            starts = torch.zeros(1, dtype=torch.int64, device=dev)
            pout = pre(tp, tl, tr, starts, F)
            for _ in range(3):
                pre(tp, tl, tr, starts, F, out=pout)
            torch.cuda.synchronize()
            pgraph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(pgraph):
                for _ in range(4):
                    pre(tp, tl, tr, starts, F, out=pout)
            pgraph.replay()
            torch.cuda.synchronize()
            preps = 20
            ev0.record()
            for _ in range(preps // 4):
                pgraph.replay()
            ev1.record()
            torch.cuda.synchronize()
            p_ms = ev0.elapsed_time(ev1) / preps
            gbs = F * PRE_BYTES_PER_FRAME / (p_ms * 1e-3) / 1e9
            ptraffic = None
            if os.path.isfile(tpath):
                pf = json.load(open(tpath)).get("preprocess_dram_bytes_per_frame")
                ptraffic = pf * F if pf else None
            line["preprocess"] = {"metric": "body2hand_preprocess_frames_per_sec", "value": F / (p_ms * 1e-3), "unit": "frames/s",
                                  "ms_per_launch": p_ms, "workload": f"{F} frames (2 h at 30 fps), full reference item (1452 B/frame)",
                                  "roofline": {"bound": "hbm", "achieved": gbs, "peak": pk["hbm_gbs"], "unit": "GB/s",
                                               "frac": gbs / pk["hbm_gbs"], "traffic": ptraffic,
                                               "note": "CUDA-graph replay, outputs reused; 314 MB moved per launch > 126 MB L2"}}
        if not args.skip_extras:
            # ---------------- config 5: streaming K0 -> K1 over 1 h of frames (stride 64 and 16) ----------------
            Fs = 108000
            sp_, sl_, sr_ = (t_[:Fs].contiguous() for t_ in (tp, tl, tr))
            pre_s = b2h.PreprocessRightHand(with_left_hand=False, emit_bf16=(fwd_prec == "bf16"))
            line["stream"] = {"workload": f"{Fs} frames (1 h at 30 fps): K0 preprocessing (+bf16 copy) -> forward over 64-frame windows, CUDA graph",
                              "dtype": fwd_prec}
            for stride in (64, 16):
                st_ = torch.from_numpy(b2h.sliding_window_starts(Fs - T + 1, T, stride)).to(dev)
                Wn = st_.numel()
                so = pre_s(sp_, sl_, sr_, st_, T)
                xin = so["input_kp_bf16"] if fwd_prec == "bf16" else so["input_kp"]
                frs = ForwardRunner(fmodel, Wn, T, n_slots=1, x_dtype=xin.dtype, out_scale=1280.0)
                frs.x = xin.view(1, Wn, T, 12, 2)              # the net reads K0's output in place (no copy)
                pre_s(sp_, sl_, sr_, st_, T, out=so); frs.run(0)
                torch.cuda.synchronize()
                sg = torch.cuda.CUDAGraph()
                with torch.cuda.graph(sg):
                    for _ in range(4):
                        pre_s(sp_, sl_, sr_, st_, T, out=so)
                        frs.run(0)
                sg.replay(); torch.cuda.synchronize()
                ev0.record()
                for _ in range(5):
                    sg.replay()
                ev1.record()
                torch.cuda.synchronize()
                s_ms = ev0.elapsed_time(ev1) / 20
                line["stream"][f"stride{stride}"] = {"windows": Wn, "ms": s_ms, "unique_frames_per_sec": Fs / (s_ms * 1e-3),
                                                     "window_frames_per_sec": Wn * T / (s_ms * 1e-3)}
        # ---------------- CPU baseline (reference path on this box's host cores) ----------------
        line["cpu_baseline"], _, _ = cpu_reference_arm(200, 2, budget_s=15.0)
        if not args.skip_extras:
            # ---------------- launch-latency floor: 40 dependent one-block kernels per graph replay ----------------
            # Last on purpose and fully guarded: nothing measured above can be affected by it.
            try:
                tiny = torch.zeros((1, 1, 21, 2), dtype=torch.float32, device=dev)
                tlen = torch.ones(1, dtype=torch.int32, device=dev)
                from hand_pose_sl_b200 import _lib as _l

                def _tiny():
                    _l.check(lib.b2h_mask_output(_l.ptr(tiny), _l.ptr(tlen), 1, 1, 42, _l.stream_ptr(dev)))
                _tiny()
                torch.cuda.synchronize()
                fg = torch.cuda.CUDAGraph()
                with torch.cuda.graph(fg):
                    for _ in range(40):
                        _tiny()
                fg.replay()
                torch.cuda.synchronize()
                ev0.record()
                for _ in range(10):
                    fg.replay()
                ev1.record()
                torch.cuda.synchronize()
                line["launch_floor_us"] = ev0.elapsed_time(ev1) * 1e3 / 400
            except Exception as e:   # noqa: BLE001
                line["launch_floor_us"] = None
                line["launch_floor_error"] = str(e)[:200]

    sys.stdout.flush()
    try:                                   # NCCL's banner sits in the C stdio buffer: push it to the redirected fd first
        import ctypes
        ctypes.CDLL(None).fflush(None)
    except Exception:
        pass
    os.dup2(real_stdout, 1)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        # Tearing the NCCL communicator down while a captured CUDA graph still holds its kernels hangs in
        # destroy_process_group (seen on B200, torch 2.11 / NCCL 2.28): synchronise, then leave without the teardown.
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()

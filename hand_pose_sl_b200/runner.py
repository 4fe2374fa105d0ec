"""Low-overhead steady-state runners: static device buffers, pre-bound C-ABI arguments, optional CUDA
graph capture.  The fused train step is two kernel launches (~µs of GPU time at the reference's batch
sizes), so the per-call Python/torch bookkeeping of the modular API would dominate; these runners are
what a training / serving loop uses once shapes are fixed (launch latency, not the
tensor pipe or HBM, bounds every BASELINE config)."""
from __future__ import annotations

import torch

from . import _lib
from .models import ConvModel
from .steps import FusedAdam


class TrainStepRunner:
    """steps/traintest.py:94-121 for a fixed (B, T): `load(batch)` copies a batch into the static
    buffers, `step()` enqueues forward+mask+loss+backward and reduce+Adam+repack (2 launches).
    The Adam step counter lives on the device so a captured graph can be replayed."""

    def __init__(self, model: ConvModel, optimizer: FusedAdam, B: int, T: int, loss: str = "L1", n_slots: int = 1):
        self.model, self.opt, self.B, self.T = model, optimizer, B, T
        self.kind = _lib.LOSSES[loss]
        flat = model.flat_parameters()
        dev = flat.device
        _lib.require_device(flat, "model")
        _lib.require_sm100(dev)
        self.dev = dev
        group = optimizer.param_groups[0]
        if optimizer._owner(group) is not model:
            raise RuntimeError("TrainStepRunner needs FusedAdam(model.parameters()) over exactly this model")
        self.state = optimizer._group_state(0, group, model)
        K = model.n_in // 2
        self.n_slots = n_slots
        self.x = torch.zeros((n_slots, B, T, K, 2), dtype=torch.float32, device=dev)
        self.target = torch.zeros((n_slots, B, T, 21, 2), dtype=torch.float32, device=dev)
        self.conf = torch.zeros((n_slots, B, T, 21), dtype=torch.float32, device=dev) if self.kind == _lib.LOSS_CONFL1 else None
        self.lengths = torch.full((n_slots, B), T, dtype=torch.int32, device=dev)
        self.loss = torch.zeros((n_slots,), dtype=torch.float32, device=dev)
        self.step_dev = torch.full((1,), int(self.state["step"]), dtype=torch.int64, device=dev)
        self.ws = model.workspace(B, T)
        self.packed = model.packed_weights()
        self.lib = _lib.load()
        self.graph = None
        self._graph_steps = 0
        self.host_steps = int(self.state["step"])

    def load(self, batch, slot=0, non_blocking=True):
        """Copy one reference-style batch dict (CPU pinned or device tensors) into slot `slot`."""
        self.x[slot].copy_(batch["input_kp"], non_blocking=non_blocking)
        self.target[slot].copy_(batch["target_kp"], non_blocking=non_blocking)
        if self.conf is not None:
            self.conf[slot].copy_(batch["target_conf"], non_blocking=non_blocking)
        self.lengths[slot].copy_(batch["n_frames"], non_blocking=non_blocking)

    def step(self, slot=0):
        """Enqueue one training step on the current stream; returns the 0-dim device loss tensor."""
        m, g = self.model, self.opt.param_groups[0]
        n_in, C, pe = m._geometry()
        b1, b2 = g["betas"]
        conf = None if self.conf is None else self.conf[slot]
        _lib.check(self.lib.b2h_train_step(
            _lib.ptr(self.x[slot]), _lib.DT_F32, _lib.ptr(self.target[slot]), _lib.ptr(conf), _lib.ptr(self.lengths[slot]),
            _lib.ptr(m._flat), _lib.ptr(self.packed), _lib.ptr(self.state["m"]), _lib.ptr(self.state["v"]),
            _lib.ptr(self.loss[slot:slot + 1]), self.B, self.T, n_in, C, pe, self.kind, _lib.PRECISIONS[m.precision],
            float(g["lr"]), b1, b2, g["eps"], 0, _lib.ptr(self.step_dev), _lib.ptr(self.ws), self.ws.numel(),
            _lib.stream_ptr(self.dev)))
        self.host_steps += 1
        self.state["step"] = self.host_steps
        return self.loss[slot]

    def capture(self, n_steps=None):
        """Capture `n_steps` consecutive steps (slot i % n_slots) into one CUDA graph."""
        n_steps = n_steps or self.n_slots
        torch.cuda.synchronize(self.dev)
        saved = (self.host_steps, self.step_dev.clone(), self.model._flat.clone(), self.state["m"].clone(), self.state["v"].clone())
        s = torch.cuda.Stream(self.dev)
        s.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(s):
            for i in range(2):                       # warm-up outside capture (lazy module load, attributes)
                self.step(i % self.n_slots)
        torch.cuda.current_stream(self.dev).wait_stream(s)
        torch.cuda.synchronize(self.dev)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for i in range(n_steps):
                self.step(i % self.n_slots)
        # undo the warm-up / capture bookkeeping: graphs only record, they do not execute
        self.host_steps = saved[0]
        self.step_dev.copy_(saved[1]); self.model._flat.copy_(saved[2])
        self.state["m"].copy_(saved[3]); self.state["v"].copy_(saved[4])
        self.model.mark_packed_stale(); self.packed = self.model.packed_weights()
        self.state["step"] = self.host_steps
        torch.cuda.synchronize(self.dev)
        self.graph, self._graph_steps = g, n_steps
        return g

    def replay(self):
        self.graph.replay()
        self.host_steps += self._graph_steps
        self.state["step"] = self.host_steps

    def finish(self):
        """Publish the step count to the optimiser's torch-layout state (for state_dict())."""
        self.state["step_t"].fill_(float(self.host_steps))
        self.model.packed_weights(fresh_from_kernel=True)


def pipelined_steps(runner, batches):
    """Training loop with a double-buffered input pipeline (what a DataLoader with prefetch gives the reference loop,
    steps/traintest.py:87-123): while step i runs, batch i+1 travels host -> device on a copy stream into the other
    slot.  `runner` is a TrainStepRunner or parallel.DataParallelTrainer with n_slots >= 2; `batches` yields
    reference-style batch dicts in pinned host memory.  Yields the loss of every step as a float (device -> host
    read, traintest.py:123), so each step's result is observed before the next one is enqueued."""
    if runner.n_slots < 2:
        raise RuntimeError("pipelined_steps needs a runner with n_slots >= 2")
    dev = runner.x.device
    main = torch.cuda.current_stream(dev)
    copy_stream = torch.cuda.Stream(dev)
    ready = [torch.cuda.Event(), torch.cuda.Event()]     # batch landed in slot s
    freed = [torch.cuda.Event(), torch.cuda.Event()]     # the step that read slot s has finished
    used = [False, False]
    it = iter(batches)

    def issue(i, batch):
        s = i & 1
        if used[s]:
            copy_stream.wait_event(freed[s])             # do not overwrite a slot a step is still reading
        with torch.cuda.stream(copy_stream):
            runner.load(batch, slot=s)
            ready[s].record(copy_stream)

    nxt = next(it, None)
    if nxt is None:
        return
    issue(0, nxt)
    i = 0
    while True:
        s = i & 1
        nxt = next(it, None)
        if nxt is not None:
            issue(i + 1, nxt)                            # travels while step i computes
        main.wait_event(ready[s])
        loss = runner.step(s)
        freed[s].record(main)
        used[s] = True
        yield float(loss.item())
        if nxt is None:
            return
        i += 1


class ForwardRunner:
    """ConvModel.forward for a fixed (B, T) with static buffers (inference serving loop)."""

    def __init__(self, model: ConvModel, B: int, T: int, n_slots: int = 1, x_dtype=torch.float32, out_scale: float = 1.0):
        self.model, self.B, self.T = model, B, T
        flat = model.flat_parameters()
        dev = flat.device
        _lib.require_device(flat, "model")
        _lib.require_sm100(dev)
        self.dev = dev
        K = model.n_in // 2
        self.x = torch.zeros((n_slots, B, T, K, 2), dtype=x_dtype, device=dev)
        self.y = torch.zeros((n_slots, B, T, 21, 2), dtype=torch.float32, device=dev)
        self.n_slots = n_slots
        self.out_scale = float(out_scale)
        self.packed = model.packed_weights()
        self.lib = _lib.load()
        self.graph = None

    def run(self, slot=0):
        m = self.model
        n_in, C, pe = m._geometry()
        _lib.check(self.lib.b2h_conv_forward(
            _lib.ptr(self.x[slot]), _lib.DT_BF16 if self.x.dtype == torch.bfloat16 else _lib.DT_F32, _lib.ptr(m._flat),
            _lib.ptr(self.packed), None, _lib.ptr(self.y[slot]), self.B, self.T, n_in, C, pe, _lib.PRECISIONS[m.precision], 0,
            self.out_scale, _lib.stream_ptr(self.dev)))
        return self.y[slot]

    def capture(self, n_runs=None):
        n_runs = n_runs or self.n_slots
        torch.cuda.synchronize(self.dev)
        s = torch.cuda.Stream(self.dev)
        s.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(s):
            self.run(0)
        torch.cuda.current_stream(self.dev).wait_stream(s)
        torch.cuda.synchronize(self.dev)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for i in range(n_runs):
                self.run(i % self.n_slots)
        self.graph = g
        return g

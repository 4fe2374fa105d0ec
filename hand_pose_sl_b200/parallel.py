"""Data-parallel training.  The reference is single-process; windows are independent
units (per-sample zero padding, no cross-sample op: HandPoseModels.py:55-58), so training shards by
batch with ONE gradient all-reduce per step over the flat fp32 gradient buffer (76 KB at C=30) and
inference / preprocessing shard with no collective at all.

Combine rule: maskedPoseL1 divides by the local batch size (steps/utils.py:428), so with equal
per-rank batches  global_grad = (1/W) * sum_r grad_r;  poderatedPoseL1 sums over the batch
(steps/utils.py:452), so  global_grad = sum_r grad_r."""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import _lib
from .models import ConvModel
from .steps import FusedAdam


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous block partition in rank order (rank r gets items [lo, hi))."""
    per, rem = divmod(n_items, world)
    lo = rank * per + min(rank, rem)
    return lo, lo + per + (1 if rank < rem else 0)


def shard_batch(batch: dict, rank: int, world: int):
    """Rank r's slice of a reference-style batch dict (tensors indexed on dim 0, lists sliced)."""
    B = batch["input_kp"].shape[0]
    if B % world != 0:
        raise RuntimeError(f"global batch {B} is not divisible by world size {world}: the loss combine rule "
                           "(mean of per-rank means) needs equal per-rank batches")
    lo, hi = shard_range(B, rank, world)
    return {k: (v[lo:hi] if hasattr(v, "__getitem__") else v) for k, v in batch.items()}


def grad_scale_for(loss: str, world: int) -> float:
    return 1.0 / world if loss == "L1" else 1.0


def allreduce_flat(flat_grads: torch.Tensor, group=None):
    """The one collective of a training step: SUM over ranks of the flat gradient buffer, in place."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM, group=group)
    return flat_grads


def combine_losses(loss: torch.Tensor, kind: str, group=None):
    """Logging only: the global loss value from per-rank losses."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        out = loss.detach().clone()
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
        return out / dist.get_world_size(group) if kind == "L1" else out
    return loss


def broadcast_parameters(model: ConvModel, src: int = 0, group=None):
    """Identical replicas: rank `src`'s flat parameter buffer to every rank."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        flat = model.flat_parameters()
        dist.broadcast(flat, src=src, group=group)
        model.mark_packed_stale()


def _symmetric_exchange_buffer(n_floats: int, device, group):
    """Peer-mapped buffer [2][P] fp32 gradients + [world] int64 flags on every rank (torch symmetric memory:
    cuMem allocations mapped into every peer over NVLink).  Returns (local tensor, device array of peer base
    pointers indexed by rank)."""
    import torch.distributed._symmetric_memory as symm_mem
    world = dist.get_world_size(group)
    buf = symm_mem.empty(n_floats, dtype=torch.float32, device=device)
    buf.zero_()
    hdl = symm_mem.rendezvous(buf, group if group is not None else dist.group.WORLD)
    ptrs = [int(p) for p in hdl.buffer_ptrs]
    if len(ptrs) != world or hdl.rank != dist.get_rank(group):
        raise RuntimeError("symmetric memory rendezvous returned an unexpected peer table")
    torch.cuda.synchronize(device)
    dist.barrier(group)                                   # every rank's buffer is zeroed before anyone signals
    return buf, hdl, torch.tensor(ptrs, dtype=torch.int64, device=device)


class DataParallelTrainer:
    """One process per GPU.  exchange="p2p" (default when symmetric memory is available): step() = [forward+mask+
    loss+backward kernel, partial-reduce kernel -> peer-mapped buffer] -> [ONE kernel: flag exchange with all peers,
    sum of the peers' gradients over NVLink, Adam + re-pack] -- the collective is fused with the optimiser, no NCCL
    call on the step path.  exchange="nccl": ... -> dist.all_reduce(flat grads) -> Adam kernel (baseline).
    Static buffers; the whole step can be captured in a CUDA graph either way."""

    def __init__(self, model: ConvModel, optimizer: FusedAdam, B: int, T: int, loss: str = "L1", group=None,
                 n_slots: int = 1, exchange: str = "auto"):
        self.model, self.opt, self.B, self.T, self.loss_name, self.group = model, optimizer, B, T, loss, group
        self.kind = _lib.LOSSES[loss]
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        flat = model.flat_parameters()
        self.dev = flat.device
        _lib.require_device(flat, "model")
        _lib.require_sm100(self.dev)
        broadcast_parameters(model, 0, group)
        g = optimizer.param_groups[0]
        if optimizer._owner(g) is not model:
            raise RuntimeError("DataParallelTrainer needs FusedAdam(model.parameters()) over exactly this model")
        self.state = optimizer._group_state(0, g, model)
        K = model.n_in // 2
        dev = self.dev
        self.n_slots = n_slots
        self.x = torch.zeros((n_slots, B, T, K, 2), dtype=torch.float32, device=dev)
        self.target = torch.zeros((n_slots, B, T, 21, 2), dtype=torch.float32, device=dev)
        self.conf = torch.zeros((n_slots, B, T, 21), dtype=torch.float32, device=dev) if self.kind == _lib.LOSS_CONFL1 else None
        self.lengths = torch.full((n_slots, B), T, dtype=torch.int32, device=dev)
        self.loss = torch.zeros((n_slots,), dtype=torch.float32, device=dev)
        self.grads = torch.zeros_like(flat)
        self.step_dev = torch.full((1,), int(self.state["step"]), dtype=torch.int64, device=dev)
        self.ws = model.workspace(B, T)
        self.packed = model.packed_weights()
        self.lib = _lib.load()
        self.host_steps = int(self.state["step"])
        self.graph, self._graph_steps = None, 0
        if exchange not in ("auto", "p2p", "nccl"):
            raise ValueError("exchange must be 'auto', 'p2p' or 'nccl'")
        self.exchange = "nccl"
        self.sym = self.sym_hdl = self.peer_ptrs = None
        if self.world > 1 and exchange in ("auto", "p2p"):
            try:
                n_in_, C_, pe_ = model._geometry()
                nfl = int(self.lib.b2h_dp_exchange_floats(n_in_, C_, pe_, self.world))
                self.sym, self.sym_hdl, self.peer_ptrs = _symmetric_exchange_buffer(nfl, dev, group)
                self.epoch_dev = torch.zeros((1,), dtype=torch.int64, device=dev)
                self.rank = dist.get_rank(group)
                self.exchange = "p2p"
            except Exception as e:  # noqa: BLE001  (no peer access / symmetric memory unavailable)
                if exchange == "p2p":
                    raise
                import warnings
                warnings.warn(f"symmetric-memory gradient exchange unavailable ({type(e).__name__}: {e}); using NCCL all-reduce")

    def load(self, batch, slot=0, non_blocking=True):
        self.x[slot].copy_(batch["input_kp"], non_blocking=non_blocking)
        self.target[slot].copy_(batch["target_kp"], non_blocking=non_blocking)
        if self.conf is not None:
            self.conf[slot].copy_(batch["target_conf"], non_blocking=non_blocking)
        self.lengths[slot].copy_(batch["n_frames"], non_blocking=non_blocking)

    def step(self, slot=0):
        m, g = self.model, self.opt.param_groups[0]
        n_in, C, pe = m._geometry()
        b1, b2 = g["betas"]
        conf = None if self.conf is None else self.conf[slot]
        sp = _lib.stream_ptr(self.dev)
        if self.exchange == "p2p":
            _lib.check(self.lib.b2h_train_step_dp(
                _lib.ptr(self.x[slot]), _lib.DT_F32, _lib.ptr(self.target[slot]), _lib.ptr(conf), _lib.ptr(self.lengths[slot]),
                _lib.ptr(m._flat), _lib.ptr(self.packed), _lib.ptr(self.state["m"]), _lib.ptr(self.state["v"]),
                _lib.ptr(self.loss[slot:slot + 1]), self.B, self.T, n_in, C, pe, self.kind, _lib.PRECISIONS[m.precision],
                float(g["lr"]), b1, b2, g["eps"], _lib.ptr(self.step_dev), _lib.ptr(self.epoch_dev), _lib.ptr(self.sym),
                _lib.ptr(self.peer_ptrs), self.rank, self.world, grad_scale_for(self.loss_name, self.world),
                _lib.ptr(self.ws), self.ws.numel(), sp))
            self.host_steps += 1
            self.state["step"] = self.host_steps
            return self.loss[slot]
        _lib.check(self.lib.b2h_train_forward_backward(
            _lib.ptr(self.x[slot]), _lib.DT_F32, _lib.ptr(self.target[slot]), _lib.ptr(conf), _lib.ptr(self.lengths[slot]),
            _lib.ptr(m._flat), _lib.ptr(self.packed), _lib.ptr(self.grads), _lib.ptr(self.loss[slot:slot + 1]), None,
            self.B, self.T, n_in, C, pe, self.kind, _lib.PRECISIONS[m.precision], _lib.ptr(self.step_dev),
            _lib.ptr(self.ws), self.ws.numel(), sp))
        allreduce_flat(self.grads, self.group)
        _lib.check(self.lib.b2h_adam_step(
            _lib.ptr(m._flat), _lib.ptr(self.grads), _lib.ptr(self.state["m"]), _lib.ptr(self.state["v"]), m._flat.numel(),
            float(g["lr"]), b1, b2, g["eps"], 0, _lib.ptr(self.step_dev), grad_scale_for(self.loss_name, self.world),
            _lib.ptr(self.packed), n_in, C, pe, sp))
        self.host_steps += 1
        self.state["step"] = self.host_steps
        return self.loss[slot]

    def capture(self, n_steps=None):
        n_steps = n_steps or self.n_slots
        torch.cuda.synchronize(self.dev)
        saved = (self.host_steps, self.step_dev.clone(), self.model._flat.clone(), self.state["m"].clone(), self.state["v"].clone())
        s = torch.cuda.Stream(self.dev)
        s.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(s):
            for i in range(3):
                self.step(i % self.n_slots)
        torch.cuda.current_stream(self.dev).wait_stream(s)
        torch.cuda.synchronize(self.dev)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for i in range(n_steps):
                self.step(i % self.n_slots)
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
        self.state["step_t"].fill_(float(self.host_steps))
        self.model.packed_weights(fresh_from_kernel=True)

"""CPU tier: the data-parallel host logic (sharding, the one gradient all-reduce, loss combine, parameter
broadcast) on a world_size-2 gloo group.  Gradients come from the oracle (test infrastructure); the identity
checked is the one DP training relies on:  grad(full batch) == (1/W) sum_r grad(shard r)  for
maskedPoseL1 and  == sum_r grad(shard r)  for poderatedPoseL1."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, kind, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import b2h_oracle as oracle
    import hand_pose_sl_b200 as b2h
    from hand_pose_sl_b200 import parallel, synthetic
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    try:
        batch = synthetic.model_batch(8, 32, seed=5, ragged=True)
        sd = oracle.init_params(30, False, seed=0)
        # replicas start from different weights; broadcast makes them identical
        torch.manual_seed(100 + rank)
        model = b2h.ConvModel(30, "ReLU", False)
        parallel.broadcast_parameters(model, 0)
        got = torch.cat([p.detach().reshape(-1) for p in model._ordered_params()])
        torch.manual_seed(100)
        want = torch.cat([p.detach().reshape(-1) for p in b2h.ConvModel(30, "ReLU", False)._ordered_params()])
        assert torch.equal(got, want) and model._is_flat()
        # shard, local gradients (oracle), the collective, the combine rule
        shard = parallel.shard_batch(batch, rank, world)
        assert shard["input_kp"].shape[0] == 4 and torch.equal(shard["input_kp"], batch["input_kp"][4 * rank:4 * rank + 4])
        st = oracle.TrainState(sd)
        loss, grads = oracle.train_step(st, shard["input_kp"], shard["target_kp"], shard["n_frames"], kind, shard["target_conf"])
        flat = torch.cat([grads[k].reshape(-1) for k in oracle.PARAM_NAMES])
        parallel.allreduce_flat(flat)
        flat *= parallel.grad_scale_for(kind, world)
        gl = parallel.combine_losses(torch.tensor(loss), kind)
        full = oracle.TrainState(sd)
        floss, fgrads = oracle.train_step(full, batch["input_kp"], batch["target_kp"], batch["n_frames"], kind, batch["target_conf"])
        fflat = torch.cat([fgrads[k].reshape(-1) for k in oracle.PARAM_NAMES])
        err = float((flat - fflat).abs().max() / fflat.abs().max())
        assert err < 1e-5, err
        assert abs(float(gl) - floss) < 1e-5 * abs(floss)
        ret[rank] = "ok"
    except Exception as e:  # noqa: BLE001
        ret[rank] = f"{type(e).__name__}: {e}"
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("kind", ["L1", "confL1"])
def test_dp_host_logic_world2(kind):
    world = 2
    port = 29500 + (os.getpid() % 2000) + (0 if kind == "L1" else 1)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, kind, ret), nprocs=world, join=True)
    assert dict(ret) == {0: "ok", 1: "ok"}


def test_shard_range_and_errors():
    from hand_pose_sl_b200 import parallel
    assert [parallel.shard_range(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    with pytest.raises(RuntimeError):
        parallel.shard_batch({"input_kp": torch.zeros(5, 4, 12, 2)}, 0, 2)
    assert parallel.grad_scale_for("L1", 8) == 0.125 and parallel.grad_scale_for("confL1", 8) == 1.0

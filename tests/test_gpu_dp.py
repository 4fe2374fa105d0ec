"""GPU tier (needs >= 2 GPUs, skipped otherwise): W-rank data-parallel training with per-rank batch b ==
1-rank training with batch W*b concatenated in rank order: loss, post-step weights within fp32
tolerance (summation order differs), replicas bit-identical to each other."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, prec, kind, exchange, C, ret):
    sys.path.insert(0, ROOT)
    import hand_pose_sl_b200 as b2h
    from hand_pose_sl_b200 import parallel, synthetic
    from hand_pose_sl_b200.runner import TrainStepRunner
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        B, T, steps = 16 * world, 64, 4
        batch = synthetic.model_batch(B, T, seed=77, ragged=True)
        torch.manual_seed(0)
        model = b2h.ConvModel(C, "ReLU", False, precision=prec).to(dev)
        opt = b2h.FusedAdam(model.parameters(), lr=2e-4)
        tr = parallel.DataParallelTrainer(model, opt, B // world, T, kind, exchange=exchange)
        assert tr.exchange == exchange
        tr.load(parallel.shard_batch(batch, rank, world), non_blocking=False)
        losses = []
        for s in range(steps):
            l = tr.step(0)
            losses.append(float(parallel.combine_losses(l, kind)))
        tr.finish()
        flat = model.flat_parameters().clone()
        # replicas identical
        gathered = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(gathered, flat)
        assert all(torch.equal(gathered[0], g) for g in gathered), "replicas diverged"
        # a captured graph of the same step keeps working (NCCL inside the graph)
        tr.capture(2)
        tr.replay()
        torch.cuda.synchronize()
        from hand_pose_sl_b200 import _lib
        assert _lib.load().b2h_dp_status() == 0, "peer-flag wait timed out"
        assert _lib.load().b2h_tc_status() == 0, "fused exchange: wait for a peer's gradient words timed out"
        tr.check_status()
        if rank == 0:
            torch.manual_seed(0)
            ref = b2h.ConvModel(C, "ReLU", False, precision=prec).to(dev)
            ropt = b2h.FusedAdam(ref.parameters(), lr=2e-4)
            rr = TrainStepRunner(ref, ropt, B, T, kind)
            rr.load(batch, non_blocking=False)
            rl = [float(rr.step(0)) for _ in range(steps)]
            for a, b in zip(losses, rl):
                assert abs(a - b) <= 1e-4 * abs(b), (losses, rl)
            d = (flat - ref.flat_parameters()).abs().max().item()
            assert d <= 2 * 2e-4 * steps
            frac_close = ((flat - ref.flat_parameters()).abs() <= 1e-4 * ref.flat_parameters().abs().max()).float().mean().item()
            assert frac_close > 0.97, frac_close      # Adam on noise-level gradients: see oracle.adam_conditioned
        ret[rank] = "ok"
    except Exception as e:  # noqa: BLE001
        import traceback
        ret[rank] = f"{type(e).__name__}: {e}\n{traceback.format_exc()}"
    finally:
        # destroy_process_group hangs while a captured graph still holds NCCL kernels: leave without the teardown
        try:
            torch.cuda.synchronize()
        except Exception:  # noqa: BLE001
            pass
        sys.stdout.flush()
        os._exit(0 if ret.get(rank) == "ok" else 1)


def _spawn(prec, kind, exchange, C):
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    port = 29600 + (os.getpid() % 1000) + (7 if exchange == "p2p" else 0) + (len(prec) + len(kind)) * 11 + C
    ret = mp.Manager().dict()
    try:
        mp.spawn(_worker, args=(world, port, prec, kind, exchange, C, ret), nprocs=world, join=True)
    except Exception as e:  # noqa: BLE001  (a worker exited non-zero: its message is in `ret`)
        assert False, (str(e)[:300], dict(ret))
    assert len(ret) == world and all(v == "ok" for v in ret.values()), dict(ret)


@pytest.mark.parametrize("exchange", ["p2p", "nccl"])
@pytest.mark.parametrize("prec,kind", [("fp32", "L1"), ("bf16", "L1"), ("bf16", "confL1")])
def test_data_parallel_matches_single_gpu(prec, kind, exchange):
    _spawn(prec, kind, exchange, 30)


@pytest.mark.parametrize("prec,C", [("fp32-ffma", 30), ("bf16", 64)])
def test_data_parallel_three_launch_path(prec, C):
    """Shapes without the one-launch step (CUDA-core fp32 mode, wide models): forward/backward -> slice reduction into
    the symmetric buffer -> `adam_dp_kernel` reading every peer's gradients over NVLink (b2h_train_step_dp's other branch)."""
    _spawn(prec, "L1", "p2p", C)

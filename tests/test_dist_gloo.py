"""World-size-2 check of the data-parallel scheme on CPU with the gloo backend.

The CUDA kernels cannot run here, so each rank builds its shard's partial joints with the oracle, the
PRODUCT's exchange hook (``iic_b200.ops._maybe_allreduce`` under ``set_data_parallel(True)``) combines
them, and the oracle's epilogue/backward finish locally.  What is asserted is the scheme itself
(SURVEY.md section 8e): after ONE sum-all-reduce of the joints every rank holds the loss of the GLOBAL
batch and the gradient of that loss with respect to its own shard.
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _views(seed, B, K, H, W):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import iic_oracle as O
    rng = np.random.default_rng(seed)
    base = rng.standard_normal((B, K, H // 4, W // 4)).repeat(4, axis=2).repeat(4, axis=3) * 3
    x = O.softmax(base + 0.5 * rng.standard_normal((B, K, H, W))).astype(np.float32)
    y = O.softmax(base + 0.5 * rng.standard_normal((B, K, H, W))).astype(np.float32)
    gx = O.softmax(rng.standard_normal((B, K)) * 2).astype(np.float32)
    gy = O.softmax(rng.standard_normal((B, K)) * 2).astype(np.float32)
    return x, y, gx, gy


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import iic_oracle as O
    import iic_b200
    from iic_b200 import ops as iops

    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        B, K, H, W, pad = 4, 5, 16, 20, 1
        x, y, gx, gy = _views(7, B, K, H, W)
        lo, hi = rank * B // world, (rank + 1) * B // world
        xs, ys = x[lo:hi], y[lo:hi]

        iic_b200.set_data_parallel(True)
        # local term: partial joint of the shard -> the product's exchange hook -> epilogue + backward locally
        J = torch.from_numpy(O.local_joint(xs, ys, pad)[None])
        J = iops._maybe_allreduce(J)
        loss, GA = O.local_loss_from_joint(J[0].numpy(), 1.0)
        dxs, dys = O.local_backward_from_GA(xs, ys, GA, pad)
        # global term: J = x^T y
        Jg = torch.from_numpy(gx[lo:hi].astype(np.float64).T @ gy[lo:hi].astype(np.float64))
        Jg = iops._maybe_allreduce(Jg)
        iic_b200.set_data_parallel(False)
        # disabled hook must be the identity even with a process group up
        same = torch.ones(3, dtype=torch.float64)
        assert torch.equal(iops._maybe_allreduce(same.clone()), same)

        full_loss, full_dx, full_dy = O.iid_segmentation_loss(x, y, pad, with_grads=True)
        np.testing.assert_allclose(loss, full_loss, rtol=1e-12)
        np.testing.assert_allclose(dxs, full_dx[lo:hi], rtol=1e-10, atol=1e-14)
        np.testing.assert_allclose(dys, full_dy[lo:hi], rtol=1e-10, atol=1e-14)
        np.testing.assert_allclose(Jg.numpy(), gx.astype(np.float64).T @ gy.astype(np.float64), rtol=1e-13)
        out.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        out.put((rank, f"{type(e).__name__}: {e}"))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_joint_allreduce_world2():
    if not dist.is_available() or not dist.is_gloo_available():
        pytest.skip("gloo backend not available")
    pytest.importorskip("iic_b200")
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    results = [out.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, "ok"), (1, "ok")], results


# ---- gradient semantics under DistributedDataParallel (SURVEY.md section 7 "DDP semantics", section 8f row 3) -----
def _ddp_model():
    torch.manual_seed(3)
    return torch.nn.Linear(6, 5).double()


def _ddp_losses(model, a, b, allreduce):
    """mean-type term (MSE of the two views) + the global IIC term on the exchanged joint (iic_loss.py:43-94)."""
    pa, pb = model(a).softmax(1), model(b).softmax(1)
    mse = ((pa - pb) ** 2).mean()
    J = pa.t() @ pb                                        # this rank's partial joint
    J = J + (allreduce(J.detach().clone()) - J.detach())   # value: the global joint; gradient: identity to the shard
    P = (J + J.t()) / 2.0
    P = P / P.sum()
    pi, pj = P.sum(1, keepdim=True).expand_as(P), P.sum(0, keepdim=True).expand_as(P)
    iic = (-P * (torch.log(P + 1e-10) - torch.log(pj + 1e-10) - torch.log(pi + 1e-10))).sum()
    return mse, iic


def _ddp_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT)
    import iic_b200
    from iic_b200 import ops as iops
    from torch.nn.parallel import DistributedDataParallel as DDP

    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(5)
        a = torch.randn(8, 6, generator=g, dtype=torch.float64)
        b = a + 0.3 * torch.randn(8, 6, generator=g, dtype=torch.float64)
        lo, hi = rank * 8 // world, (rank + 1) * 8 // world

        # reference: one process, the full batch
        ref = _ddp_model()
        mse, iic = _ddp_losses(ref, a, b, lambda J: J)
        (mse + 0.3 * iic).backward()

        assert iic_b200.ddp_loss_scale() == 1.0                       # exchange off
        iic_b200.set_data_parallel(True)
        assert iic_b200.ddp_loss_scale() == float(world)
        model = DDP(_ddp_model())
        mse_r, iic_r = _ddp_losses(model, a[lo:hi], b[lo:hi], iops._maybe_allreduce)
        (mse_r + 0.3 * iic_b200.ddp_loss_scale() * iic_r).backward()  # DDP averages the gradients over the ranks
        iic_b200.set_data_parallel(False)

        np.testing.assert_allclose(iic_r.item(), iic.item(), rtol=1e-12)         # every rank: the global-batch loss
        for p_ref, p in zip(ref.parameters(), model.module.parameters()):
            np.testing.assert_allclose(p.grad.numpy(), p_ref.grad.numpy(), rtol=1e-9, atol=1e-13)
        out.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        out.put((rank, f"{type(e).__name__}: {e}"))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_ddp_gradient_semantics_world2():
    """Stock DDP (gradient mean) + ddp_loss_scale() on the IIC term == the single-process full-batch gradient."""
    if not dist.is_available() or not dist.is_gloo_available():
        pytest.skip("gloo backend not available")
    pytest.importorskip("iic_b200")
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ddp_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    results = [out.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, "ok"), (1, "ok")], results

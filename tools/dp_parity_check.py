"""Parity of the PRODUCT loss path under data parallelism against the full-batch fp64 oracle.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29541 tools/dp_parity_check.py

Every rank holds a contiguous shard of ONE global batch (same numpy seed everywhere).  With
``iic_b200.set_data_parallel(True)`` each rank's loss must equal the oracle's loss on the CONCATENATED batch (and be
bit-identical on all ranks), and each rank's gradient must equal the slice of the oracle's full-batch gradient that
belongs to its shard.  Checked for both transports -- the NVLink peer-memory exchange folded into
the finish launch (csrc/finish.cu) and NCCL -- for the public modules one call at a time, for ``iic_losses`` (several
terms, ONE exchange), eagerly and from a replayed CUDA graph.  Prints ``DP PARITY PASS`` on rank 0.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import iic_oracle as O  # noqa: E402  (the checker)
import iic_b200  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
LOSS_RTOL, GRAD_RTOL = 1e-5, 1e-4
ok = True


def relmax(a, ref):
    return float(np.abs(np.asarray(a, np.float64) - ref).max() / max(np.abs(ref).max(), 1e-300))


def views(rng, B, K, H, W):
    base = rng.standard_normal((B, K, max(H // 4, 1), max(W // 4, 1))).repeat(4, axis=2).repeat(4, axis=3)[:, :, :H, :W] * 3
    x = O.softmax(base + 0.5 * rng.standard_normal((B, K, H, W))).astype(np.float32)
    y = O.softmax(base + 0.5 * rng.standard_normal((B, K, H, W))).astype(np.float32)
    return x, y


def fail(msg):
    global ok
    ok = False
    print(f"[rank {rank}] FAIL {msg}", flush=True)


def same_on_all_ranks(t, what):
    ref = t.detach().clone()
    dist.broadcast(ref, 0)
    if not torch.equal(ref, t.detach()):
        fail(f"{what}: not bit-identical across ranks")


per = 3
Bt = per * world
lo, hi = rank * per, (rank + 1) * per
rng = np.random.default_rng(20261)
cases = []
for (K, H, W, pad, patch) in [(10, 32, 48, 1, 512), (20, 24, 32, 3, 1024), (6, 40, 44, 2, 16), (20, 32, 64, 1, 512)]:
    x, y = views(rng, Bt, K, H, W)
    cases.append(("local", K, pad, patch, x, y, O.iid_segmentation_small_path_loss(x, y, pad, patch, with_grads=True)))
gb = rng.standard_normal((Bt, 10)) * 2
gx = O.softmax(gb + 0.7 * rng.standard_normal((Bt, 10))).astype(np.float32)
gy = O.softmax(gb + 0.7 * rng.standard_normal((Bt, 10))).astype(np.float32)
g_ref = (O.iid_loss(gx, gy, 1.0), O.iid_loss_grads(gx, gy, 1.0))

for transport in ("peer_memory", "nccl"):
    iic_b200.set_data_parallel(True, peer_memory=(transport == "peer_memory"))
    if iic_b200.data_parallel_transport() != transport:
        fail(f"transport {iic_b200.data_parallel_transport()} != {transport}")
    # ---- the public modules, one call at a time ----
    for kind, K, pad, patch, x, y, (ol, ogx, ogy) in cases:
        xd = torch.from_numpy(x[lo:hi]).to(dev).requires_grad_(True)
        yd = torch.from_numpy(y[lo:hi]).to(dev).requires_grad_(True)
        loss = iic_b200.IIDSegmentationSmallPathLoss(padding=pad, patch_size=patch)(xd, yd)
        loss.backward()
        if abs(loss.item() - ol) > LOSS_RTOL * max(abs(ol), 0.05):
            fail(f"{transport} local K={K} pad={pad}: loss {loss.item()} vs full-batch oracle {ol}")
        ex, ey = relmax(xd.grad.cpu().numpy(), ogx[lo:hi]), relmax(yd.grad.cpu().numpy(), ogy[lo:hi])
        # the shard's gradient is measured against the max-norm of the FULL gradient it is a slice of
        scale = np.abs(ogx[lo:hi]).max() / np.abs(ogx).max()
        if ex * scale > GRAD_RTOL or ey * scale > GRAD_RTOL:
            fail(f"{transport} local K={K} pad={pad}: shard gradient err {ex:.2e} / {ey:.2e}")
        same_on_all_ranks(loss, f"{transport} local K={K} pad={pad} loss")
    a = torch.from_numpy(gx[lo:hi]).to(dev).requires_grad_(True)
    b = torch.from_numpy(gy[lo:hi]).to(dev).requires_grad_(True)
    gl, gnl, P = iic_b200.IIDLoss()(a, b)
    gl.backward()
    if abs(gl.item() - g_ref[0][0]) > LOSS_RTOL * abs(g_ref[0][0]) or relmax(P.detach().cpu().numpy(), g_ref[0][2]) > 1e-6:
        fail(f"{transport} global: loss {gl.item()} vs {g_ref[0][0]}")
    if relmax(a.grad.cpu().numpy(), g_ref[1][0][lo:hi]) > GRAD_RTOL or relmax(b.grad.cpu().numpy(), g_ref[1][1][lo:hi]) > GRAD_RTOL:
        fail(f"{transport} global: shard gradient")
    same_on_all_ranks(gl, f"{transport} global loss")
    # ---- several terms, ONE exchange (iic_losses), eagerly and from a replayed CUDA graph ----
    ins = []
    for kind, K, pad, patch, x, y, _ in cases:
        ins.append((iic_b200.IIDSegmentationSmallPathLoss(padding=pad, patch_size=patch),
                    torch.from_numpy(x[lo:hi]).to(dev).requires_grad_(True), torch.from_numpy(y[lo:hi]).to(dev).requires_grad_(True)))
    ins.append((iic_b200.IIDLoss(), torch.from_numpy(gx[lo:hi]).to(dev).requires_grad_(True),
                torch.from_numpy(gy[lo:hi]).to(dev).requires_grad_(True)))

    def step(ins=ins):
        out = iic_b200.iic_losses(ins)
        flat = [o[0] if isinstance(o, tuple) else o for o in out]
        grads = torch.autograd.grad(sum(flat), [t for _, p, q in ins for t in (p, q)])
        return torch.stack(flat), grads

    refs = [c[6][0] for c in cases] + [g_ref[0][0]]
    with iic_b200.check_mode("deferred"):
        flat, grads = step()
        for i, (v, r) in enumerate(zip(flat.tolist(), refs)):
            if abs(v - r) > LOSS_RTOL * max(abs(r), 0.05):
                fail(f"{transport} batched term {i}: {v} vs {r}")
        if relmax(grads[0].cpu().numpy(), cases[0][6][1][lo:hi]) > GRAD_RTOL * 10:
            fail(f"{transport} batched: gradient of term 0")
        same_on_all_ranks(flat, f"{transport} batched losses")
        if transport == "peer_memory":
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                # fresh leaves made on the capture stream (autograd ties a leaf's gradient to the stream it first saw)
                ins_g = [(c, p.detach().clone().requires_grad_(True), q.detach().clone().requires_grad_(True)) for c, p, q in ins]
                step(ins_g)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=side):
                gflat, ggrads = step(ins_g)
            for _ in range(4):
                graph.replay()
            torch.cuda.synchronize()
            if not torch.equal(gflat, flat):
                fail("graph replay of the batched step differs from the eager run")
            if not torch.equal(ggrads[0], grads[0]):
                fail("graph replay: gradient differs")
        iic_b200.raise_if_flagged(dev)

dist.barrier(device_ids=[local])
t = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("DP PARITY PASS" if int(t.item()) else "DP PARITY FAIL", flush=True)
torch.cuda.synchronize()
sys.stdout.flush()
os._exit(0 if int(t.item()) else 1)

"""Check the NVLink peer-memory joint exchange (csrc/xchg.cu) against torch.distributed.all_reduce.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/xchg_check.py

Every rank: random fp64 tensors of the joint sizes of the BASELINE configs, exchanged eagerly and from a replayed CUDA
graph (the sequence counter lives on the device), compared with NCCL's sum; then the latency of both transports.
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import iic_b200  # noqa: E402
from iic_b200 import ops as O  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
iic_b200.set_data_parallel(True, peer_memory=True)
assert iic_b200.data_parallel_transport() == "peer_memory"
ok = True
g = torch.Generator(device=dev).manual_seed(100 + rank)
for E in (100, 900, 19600, 147456, 900, 100, 7):
    for rep in range(3):
        J = torch.randn(E, dtype=torch.float64, device=dev, generator=g)
        ref = J.clone()
        dist.all_reduce(ref)
        out = O._maybe_allreduce(J.clone())
        torch.cuda.synchronize()
        err = (out - ref).abs().max().item()
        # bit-identical on all ranks: compare with rank 0's copy
        chk = out.clone()
        dist.broadcast(chk, 0)
        same = bool((chk == out).all().item())
        if err > 1e-12 or not same:
            ok = False
            print(f"[rank {rank}] E={E} rep={rep}: max err {err:.3e}, identical across ranks {same}", flush=True)
# graph replay
J = torch.randn(900, dtype=torch.float64, device=dev, generator=g)
src = J.clone()
side = torch.cuda.Stream(device=dev)
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    O._maybe_allreduce(J)
    J.copy_(src)
torch.cuda.current_stream().wait_stream(side)
torch.cuda.synchronize()
graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph, stream=side):
    J.copy_(src)
    O._maybe_allreduce(J)
ref = src.clone()
dist.all_reduce(ref)
for _ in range(5):
    graph.replay()
torch.cuda.synchronize()
gerr = (J - ref).abs().max().item()
if gerr > 1e-12:
    ok = False
    print(f"[rank {rank}] graph replay: max err {gerr:.3e}", flush=True)


def timed(fn, reps=200):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for E in (900, 147456):
    J = torch.randn(E, dtype=torch.float64, device=dev, generator=g)
    t_p2p = timed(lambda: O._maybe_allreduce(J))
    t_nccl = timed(lambda: dist.all_reduce(J))
    if rank == 0:
        print(f"E={E}: peer-memory exchange {t_p2p:.1f} us, NCCL all_reduce {t_nccl:.1f} us (eager launches, world {world})", flush=True)
okt = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(okt, op=dist.ReduceOp.MIN)
if rank == 0:
    print("XCHG " + ("PASS" if int(okt.item()) else "FAIL"), flush=True)
torch.cuda.synchronize()
dist.barrier()
os._exit(0)

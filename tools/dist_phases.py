"""Diagnostic (torchrun): time the phases of the sharded apply separately. usage: torchrun ... tools/dist_phases.py [spins]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import bench
import qob200 as Q
from qob200.dist import ShardedLazySum, axis_swap

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 31
p = world.bit_length() - 1
nloc = n - p
B, H = bench.build_chain(Q, n)


def timeit(fn, reps=3):
    fn()
    dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / reps], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


x = torch.empty(1 << nloc, dtype=torch.complex128, device="cuda")
Q.fill_state(x, 7, 2.0 ** (-n / 2), offset=rank << nloc)
y = torch.empty_like(x)
res = {}
for overlap in (False, True):
    sh = ShardedLazySum(H, rank, world, overlap=overlap)
    res[f"total_overlap={overlap}"] = timeit(lambda: sh.mul_(y, x, 0.5 - 1j, 0.0))
    if not overlap:
        b1, b2 = sh._buf
        res["swap"] = timeit(lambda: axis_swap(x, sh.swap_lo, sh.p, b1))
        res["swap_GBps_per_dir"] = 16 * (1 << nloc) * (1 - 1 / world) / 1e9 / (res["swap"] * 1e-3)
        res["local"] = timeit(lambda: sh._apply(sh.plan_local, 0.5, x, 0.0, y))
        res["swapped"] = timeit(lambda: sh._apply(sh.plan_swapped, 0.5, b1, 0.0, b2))
        res["add"] = timeit(lambda: y.add_(b1))
    else:
        res["localA"] = timeit(lambda: sh._apply(sh.plan_local, 0.5, x, 0.0, y))
        if sh.plan_local_b is not None:
            res["localB"] = timeit(lambda: sh._apply(sh.plan_local_b, 0.5, x, 1.0, y))
    if rank == 0:
        print(sh.describe())
try:
    sh = ShardedLazySum(H, rank, world)
    xs = sh.empty_state()
    xs.copy_(x)
    for k in (8, 16, 32, 48):
        sh.swap_sms = k
        res[f"fused_overlap_k{k}"] = timeit(lambda: sh.mul_fused_(y, xs, 0.5 - 1j, 0.0))
    sh.overlap = False
    res["fused_serial"] = timeit(lambda: sh.mul_fused_(y, xs, 0.5 - 1j, 0.0))
    zh, zptrs, _ = sh._symm[sh._zbuf.data_ptr()]
    _, xptrs, _ = sh._symm[xs.data_ptr()]
    for k in (8, 16, 32, 148):
        res[f"peer_pass_alone_k{k}"] = timeit(lambda: sh._apply_ex(sh.plan_swapped, 0.5, None, 0.0, None, peers=(xptrs, zptrs), sm_budget=k))
    res["peer_GBps_each_way_k148"] = 16 * (1 << nloc) * (1 - 1 / world) / 1e9 / (res["peer_pass_alone_k148"] * 1e-3)
except Exception as e:
    import traceback
    traceback.print_exc()
    res["fused_error"] = repr(e)
if rank == 0:
    print({k: (round(v, 2) if isinstance(v, float) else v) for k, v in res.items()}, "n", n, "world", world)
dist.destroy_process_group()

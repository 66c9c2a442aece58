"""Debug aid: the 'all communication-free terms in one plan' program of a 2-rank N=33 chain (2^32 amplitudes per rank) against the
two-plan schedule (known good), on ONE GPU, by checksums.  Usage: python tools/debug_nbits32.py [nloc]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qob200 as Q
from bench import build_chain
from qob200.dist import ShardedLazySum

nloc = int(sys.argv[1]) if len(sys.argv) > 1 else 32
n = nloc + 1
B, H = build_chain(Q, n)
x = torch.empty(1 << nloc, dtype=torch.complex128, device="cuda")
y = torch.empty(1 << nloc, dtype=torch.complex128, device="cuda")
Q.fill_state(x, 7, 2.0 ** (-n / 2))
idx = torch.tensor([0, 1, 4095, 4096, (1 << nloc) - 1, (1 << (nloc - 1)) + 5, 0x5555555, 0xAAAAAAA, (1 << 27) + 3, (1 << 31) % (1 << nloc) + 77], device="cuda")


def checks(tag):
    torch.cuda.synchronize()
    print(tag, "norm2", Q.norm2(y), "dot", Q.dot(x, y), "samples", y[idx].cpu().numpy()[:4], flush=True)
    return Q.norm2(y), Q.dot(x, y), y[idx].cpu().numpy()


for rank in (0, 1):
    sh = ShardedLazySum(H, rank, 2)
    y.zero_()
    sh._apply_ex(sh.plan_local, 1.0, x, 1.0, y)
    sh._apply_ex(sh.plan_local_b, 1.0, x, 1.0, y)
    ref = checks(f"rank {rank} two plans ")
    one = ShardedLazySum(H, rank, 2, overlap=False)
    print(one.describe()[:1500])
    y.zero_()
    one._apply_ex(one.plan_local, 1.0, x, 1.0, y)
    got = checks(f"rank {rank} one plan  ")
    print("  MATCH" if abs(got[0] - ref[0]) <= 1e-9 * abs(ref[0]) and abs(got[2] - ref[2]).max() <= 1e-12 * abs(ref[2]).max() else "  MISMATCH", abs(got[2] - ref[2]))
    y.zero_()
    one._apply_ex(one.plan_local, 1.0, x, 1.0, y, sm_budget=-32)
    got = checks(f"rank {rank} one plan, 116 SMs")
    print("  MATCH" if abs(got[0] - ref[0]) <= 1e-9 * abs(ref[0]) and abs(got[2] - ref[2]).max() <= 1e-12 * abs(ref[2]).max() else "  MISMATCH", abs(got[2] - ref[2]))

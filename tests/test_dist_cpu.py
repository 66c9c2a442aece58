"""world_size-2 gloo tests (CPU) of the sharded LazySum apply's host logic: the all-to-all axis swap, the choice
of layouts, and the complete orchestration of ShardedLazySum.mul_ with the per-rank tile programs replaced by a
numpy emulation of what qob_layout_plan_apply computes (the CUDA programs themselves are covered by the -m gpu
tests).  Rendezvous on 127.0.0.1."""
import os
import socket

import numpy as np
import pytest
import scipy.sparse as sp

import helpers as H
from helpers import O

PAULI = [np.array([[0, 1], [1, 0]], dtype=complex), np.array([[0, -1j], [1j, 0]], dtype=complex),
         np.array([[1, 0], [0, -1]], dtype=complex)]


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def chain_spec(n, seed):
    rng = np.random.default_rng(seed)
    spec = []
    for i in range(1, n + 1):
        j = i % n + 1
        for a in range(3):
            spec.append((float(rng.uniform(0.5, 1.5)), sorted([i, j]), a))
    return spec


def emulate_layout_apply(spec, info, nloc, alpha, x, beta, y):
    """numpy restatement of qob_layout_plan_apply: virtual index v = a | hi_value << nloc, subsystem k's bit at
    position bitpos[k]; y[a] = beta*y[a] + alpha * sum_t c_t sum_j prod_k A_k[i_k, j_k] x[a with bits j]."""
    a = np.arange(1 << nloc, dtype=np.int64)
    v = a | (info["hi_value"] << nloc)
    acc = np.zeros(1 << nloc, dtype=complex)
    for sel, (c, idx, pa) in zip(info["select"], spec):
        if not sel:
            continue
        A = PAULI[pa]
        p1, p2 = info["bitpos"][idx[0] - 1], info["bitpos"][idx[1] - 1]
        i1, i2 = (v >> p1) & 1, (v >> p2) & 1
        for j1 in (0, 1):
            for j2 in (0, 1):
                w = A[i1, j1] * A[i2, j2]
                if not np.any(w):
                    continue
                src = v & ~((1 << p1) | (1 << p2)) | (j1 << p1) | (j2 << p2)
                # an off-diagonal factor must never leave the local slab
                assert np.all((src >> nloc)[w != 0] == info["hi_value"])
                acc += c * w * x[src & ((1 << nloc) - 1)]
    return alpha * acc + (beta * y if beta != 0 else 0)


def _worker(rank, world, port, n, beta, out_dir):
    import torch
    import torch.distributed as dist

    import qob200 as Q
    from qob200.dist import ShardedLazySum, axis_swap

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        p = world.bit_length() - 1
        nloc = n - p
        # ---- axis swap: bit permutation of the global index, and an involution
        g = torch.arange(1 << nloc, dtype=torch.float64) + (rank << nloc)
        t = torch.complex(g, -g)
        s = 5
        out = torch.empty_like(t)
        axis_swap(t, s, p, out)
        a = np.arange(1 << nloc, dtype=np.int64)
        win = (a >> s) & (world - 1)
        expect = (a & ~((world - 1) << s)) | (rank << s) | (win << nloc)   # global index that must now sit at a
        assert np.array_equal(out.real.numpy().astype(np.int64), expect)
        back = torch.empty_like(t)
        axis_swap(out, s, p, back)
        assert torch.equal(back, t)

        # ---- full orchestration with emulated per-rank compute
        spec = chain_spec(n, 11)
        b = Q.SpinBasis(0.5)
        B = Q.tensor(*[b] * n)
        sig = (Q.sigmax(b), Q.sigmay(b), Q.sigmaz(b))
        Hq = Q.LazySum([c for c, _, _ in spec], [Q.LazyTensor(B, idx, (sig[a], sig[a])) for _, idx, a in spec])

        class Emu(ShardedLazySum):
            def _apply(self, plan, alpha, x, beta, y):
                r = emulate_layout_apply(spec, self.plan_info[plan], self.nloc, alpha, x.numpy(), beta, y.numpy())
                y.copy_(torch.from_numpy(np.ascontiguousarray(r)))

        sh = Emu(Hq, rank, world, ctx=Q.context(-1))
        assert sh.n_remote > 0 and sh.plan_swapped is not None
        xfull = O.fill_state(1 << n, 3, 2.0 ** (-n / 2))
        yfull0 = O.fill_state(1 << n, 4, 1.0)
        x = torch.from_numpy(xfull[rank << nloc:(rank + 1) << nloc].copy())
        y = torch.from_numpy(yfull0[rank << nloc:(rank + 1) << nloc].copy())
        alpha = 0.7 - 0.2j
        sh.mul_(y, x, alpha, beta)
        np.save(os.path.join(out_dir, f"y{rank}.npy"), y.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("beta", [0.0, 0.5 + 0.25j])
@pytest.mark.parametrize("world", [2, 4])
def test_sharded_apply_orchestration_gloo(tmp_path, world, beta):
    import torch.multiprocessing as mp

    n = 14 if world == 2 else 15
    port = free_port()
    mp.spawn(_worker, args=(world, port, n, beta, str(tmp_path)), nprocs=world, join=True)
    got = np.concatenate([np.load(tmp_path / f"y{r}.npy") for r in range(world)])
    # oracle: the reference's per-term sparse recursion on the full state
    dims = (2,) * n
    spec = chain_spec(n, 11)
    terms = [O.LazyTensor(dims, dims, idx, [O.Op((2,), (2,), sp.csc_matrix(PAULI[a]))] * 2) for _, idx, a in spec]
    Ho = O.LazySum(dims, dims, [c for c, _, _ in spec], terms)
    y = O.Ket(dims, O.fill_state(1 << n, 4, 1.0))
    O.mul(y, Ho, O.Ket(dims, O.fill_state(1 << n, 3, 2.0 ** (-n / 2))), 0.7 - 0.2j, beta)
    assert H.rel_err(got, y.data) <= 1e-12


@pytest.mark.parametrize("world,n", [(2, 20), (4, 20), (8, 21), (8, 33), (2, 32), (8, 17)])
def test_chunk_bits_are_fixed_in_both_single_pass_plans(world, n):
    """Planning only (no GPU): the exchange pass and the last local group may be launched in chunks only on index bits that
    are fixed (not free) in both plans, outside the swap window, and every rank must choose the same bits."""
    import qob200 as Q
    from qob200.dist import ShardedLazySum

    spec = chain_spec(n, 11)
    b = Q.SpinBasis(0.5)
    B = Q.tensor(*[b] * n)
    sig = (Q.sigmax(b), Q.sigmay(b), Q.sigmaz(b))
    Hq = Q.LazySum([c for c, _, _ in spec], [Q.LazyTensor(B, idx, (sig[a], sig[a])) for _, idx, a in spec])
    chosen = set()
    for rank in (0, world - 1):
        sh = ShardedLazySum(Hq, rank, world, ctx=Q.context(-1))
        (ns, fs), (nb, fb) = sh._plan_info(sh.plan_swapped), sh._plan_info(sh.plan_local_b)
        if sh.nchunks == 1:
            assert n == 17 and (ns != 1 or nb != 1)
            return
        assert ns == 1 and nb == 1 and sh.nchunks == 4
        m = sh.chunk_mask
        assert bin(m).count("1") == 2 and m & fs == m and m & fb == m
        assert m & (((1 << sh.p) - 1) << sh.swap_lo) == 0 and m < (1 << sh.nloc)
        chosen.add(m)
    assert len(chosen) == 1


@pytest.mark.parametrize("world,n", [(2, 20), (4, 22), (8, 23), (8, 33), (2, 33), (4, 33), (2, 14)])
def test_c_abi_planner_agrees_with_the_python_orchestration(world, n):
    """Planning only (no GPU): qob_dist_create (csrc/qob_dist.cu) restates the planning of ShardedLazySum (dist.py) — term
    classes, swap window, tile ranges — and must arrive at the same schedule on every rank; the direct mode (the exchange adds
    into the owners' result slabs) is offered exactly when the slab has >= 2^20 amplitudes."""
    import ctypes as C

    import qob200 as Q
    from qob200 import _lib
    from qob200.dist import ShardedLazySum
    from qob200.operators import handle

    spec = chain_spec(n, 11)
    b = Q.SpinBasis(0.5)
    B = Q.tensor(*[b] * n)
    sig = (Q.sigmax(b), Q.sigmay(b), Q.sigmaz(b))
    Hq = Q.LazySum([c for c, _, _ in spec], [Q.LazyTensor(B, idx, (sig[a], sig[a])) for _, idx, a in spec])
    ctx = Q.context(-1)
    for rank in (0, world - 1):
        sh = ShardedLazySum(Hq, rank, world, ctx=ctx)
        d = C.c_void_p()
        _lib.check(_lib.lib.qob_dist_create(handle(Hq, ctx), rank, world, C.byref(d)))
        nloc, nrem, nch = C.c_int32(), C.c_int32(), C.c_int32()
        slab, flagb = C.c_int64(), C.c_int64()
        _lib.check(_lib.lib.qob_dist_info(d, C.byref(nloc), C.byref(nrem), C.byref(nch), C.byref(slab), C.byref(flagb)))
        assert (nloc.value, nrem.value, nch.value) == (sh.nloc, sh.n_remote, sh.nchunks)
        assert slab.value == 16 << sh.nloc
        buf = C.create_string_buffer(1 << 15)
        _lib.check(_lib.lib.qob_dist_describe(d, buf, len(buf)))
        text = buf.value.decode()
        assert f"{sh.n_local} local + {sh.n_remote} exchanged terms" in text
        assert f"exchanged (window bit {sh.swap_lo})" in text
        cap = C.c_int32()
        _lib.check(_lib.lib.qob_dist_direct_capable(d, C.byref(cap)))
        assert bool(cap.value) == (sh.nloc >= 20), text
        _lib.check(_lib.lib.qob_dist_destroy(d))

"""Sharded LazySum apply on the GPU.

(1) every rank's tile programs (qob_layout_plan_*: rank-dependent diagonal weights through hi_value, the swapped
    layout) run one after the other on ONE GPU with the all-to-all done by tensor indexing — this checks the CUDA side of
    the multi-GPU path against the oracle on any box;
(2) the real thing: one process per GPU over NCCL, when >= 2 GPUs are visible (gpurun --gpus 2).
"""
import os
import socket

import numpy as np
import pytest
import scipy.sparse as sp

import helpers as H
from helpers import O

pytestmark = pytest.mark.gpu
PAULI = [np.array([[0, 1], [1, 0]], dtype=complex), np.array([[0, -1j], [1j, 0]], dtype=complex),
         np.array([[1, 0], [0, -1]], dtype=complex)]


def chain_spec(n, seed):
    rng = np.random.default_rng(seed)
    return [(float(rng.uniform(0.5, 1.5)), sorted([i, i % n + 1]), a) for i in range(1, n + 1) for a in range(3)]


def oracle_result(n, spec, alpha, beta, xfull, yfull0):
    dims = (2,) * n
    terms = [O.LazyTensor(dims, dims, idx, [O.Op((2,), (2,), sp.csc_matrix(PAULI[a]))] * 2) for _, idx, a in spec]
    Ho = O.LazySum(dims, dims, [c for c, _, _ in spec], terms)
    y = O.Ket(dims, yfull0.copy())
    O.mul(y, Ho, O.Ket(dims, xfull), alpha, beta)
    return y.data


def build_q(Q, n, spec):
    b = Q.SpinBasis(0.5)
    B = Q.tensor(*[b] * n)
    sig = (Q.sigmax(b), Q.sigmay(b), Q.sigmaz(b))
    return Q.LazySum([c for c, _, _ in spec], [Q.LazyTensor(B, idx, (sig[a], sig[a])) for _, idx, a in spec])


@pytest.mark.parametrize("world,n", [(2, 14), (4, 15), (8, 16), (8, 20)])
@pytest.mark.parametrize("beta", [0.0, 0.5 + 0.25j])
def test_sharded_apply_all_ranks_on_one_gpu(world, n, beta):
    import torch

    import qob200 as Q
    from qob200.dist import ShardedLazySum

    p = world.bit_length() - 1
    nloc = n - p
    spec = chain_spec(n, 21)
    xfull = O.fill_state(1 << n, 3, 2.0 ** (-n / 2))
    yfull0 = O.fill_state(1 << n, 4, 1.0)
    alpha = 0.7 - 0.2j
    ranks = [ShardedLazySum(build_q(Q, n, spec), r, world) for r in range(world)]
    xs = [torch.from_numpy(xfull[r << nloc:(r + 1) << nloc].copy()).cuda() for r in range(world)]
    ys = [torch.from_numpy(yfull0[r << nloc:(r + 1) << nloc].copy()).cuda() for r in range(world)]
    s = ranks[0].swap_lo
    assert ranks[0].n_remote > 0

    def swap_all(ts):
        # out[r][h, q, :] = in[q][h, r, :]
        lo = 1 << s
        v = [t.view(-1, world, lo) for t in ts]
        return [torch.stack([v[q][:, r, :] for q in range(world)], dim=1).reshape(-1).contiguous() for r in range(world)]

    xsw = swap_all(xs)
    ysw = [torch.full_like(x, float("nan")) for x in xs]
    for r, sh in enumerate(ranks):
        sh._apply(sh.plan_swapped, alpha, xsw[r], 0.0, ysw[r])
    yback = swap_all(ysw)
    for r, sh in enumerate(ranks):
        sh._apply(sh.plan_local, alpha, xs[r], beta, ys[r])
        if sh.plan_local_b is not None:
            sh._apply(sh.plan_local_b, alpha, xs[r], 1.0, ys[r])
        ys[r] += yback[r]
    got = np.concatenate([y.cpu().numpy() for y in ys])
    assert H.rel_err(got, oracle_result(n, spec, alpha, beta, xfull, yfull0)) <= 1e-12


# (2, 20) ... (8, 21) run chunked; from 2^20 amplitudes per rank on, the exchange and the budgeted local launches run the round-2
# kernel (peer-addressed: contiguous pieces by bulk copy; tile ranges by re-numbering the tiles)
@pytest.mark.parametrize("world,n", [(2, 14), (4, 16), (8, 17), (2, 20), (4, 20), (8, 21), (2, 21), (4, 22), (8, 23)])
@pytest.mark.parametrize("beta", [0.0, 0.5 + 0.25j])
def test_fused_peer_exchange_all_ranks_on_one_gpu(world, n, beta):
    """The PEER variant of the tile kernel (loads x tiles from the owners' slabs, stores contributions into the owners'
    buffers) with every rank's slab living on ONE GPU: same addresses and kernels as over NVLink."""
    import ctypes as C

    import torch

    import qob200 as Q
    from qob200.dist import ShardedLazySum

    p = world.bit_length() - 1
    nloc = n - p
    spec = chain_spec(n, 23)
    xfull = O.fill_state(1 << n, 5, 2.0 ** (-n / 2))
    yfull0 = O.fill_state(1 << n, 6, 1.0)
    alpha = -0.3 + 0.9j
    ranks = [ShardedLazySum(build_q(Q, n, spec), r, world) for r in range(world)]
    assert (ranks[0].nchunks > 1) == (n >= 20)
    xs = [torch.from_numpy(xfull[r << nloc:(r + 1) << nloc].copy()).cuda() for r in range(world)]
    ys = [torch.from_numpy(yfull0[r << nloc:(r + 1) << nloc].copy()).cuda() for r in range(world)]
    zs = [torch.full((1 << nloc,), float("nan"), dtype=torch.complex128, device="cuda") for _ in range(world)]
    xptrs, zptrs = [t.data_ptr() for t in xs], [t.data_ptr() for t in zs]
    l3, l4 = Q.launch_count(3), Q.launch_count(4)
    for r, sh in enumerate(ranks):
        for c in reversed(range(sh.nchunks)):   # chunked launches of the exchange pass, in any order
            sh._apply_ex(sh.plan_swapped, alpha, None, 0.0, None, peers=(xptrs, zptrs), sm_budget=8 + r, chunk=(c, sh.nchunks))
    torch.cuda.synchronize()
    assert (Q.launch_count(3) - l3) + (Q.launch_count(4) - l4) >= world * ranks[0].nchunks
    if nloc >= 20:   # the round-2 kernel ran the exchange
        assert Q.launch_count(3) - l3 == world * ranks[0].nchunks and Q.launch_count(4) == l4
    for r, sh in enumerate(ranks):
        assert torch.isfinite(torch.view_as_real(zs[r])).all()   # every element written exactly once
        if sh.plan_local_b is not None:
            sh._apply_ex(sh.plan_local, alpha, xs[r], beta, ys[r], sm_budget=100)
            for c in range(sh.nchunks):
                sh._apply_ex(sh.plan_local_b, alpha, xs[r], 1.0, ys[r], zadd=zs[r], chunk=(c, sh.nchunks))
        else:
            sh._apply_ex(sh.plan_local, alpha, xs[r], beta, ys[r], zadd=zs[r])
    got = np.concatenate([y.cpu().numpy() for y in ys])
    assert H.rel_err(got, oracle_result(n, spec, alpha, beta, xfull, yfull0)) <= 1e-12


@pytest.mark.parametrize("world,n", [(2, 21), (8, 23)])
@pytest.mark.parametrize("beta", [0.0, 0.5 + 0.25j])
def test_peer_exchange_adds_into_the_owners_result(world, n, beta):
    """The exchange without a contribution buffer: the peer-addressed pass ADDS its results into the owners' y slabs (f64 add
    performed by the owner's L2: cp.reduce.async.bulk), in any order relative to the local passes, once y holds beta*y."""
    import torch

    import qob200 as Q
    from qob200.dist import ShardedLazySum

    p = world.bit_length() - 1
    nloc = n - p
    spec = chain_spec(n, 31)
    xfull = O.fill_state(1 << n, 5, 2.0 ** (-n / 2))
    yfull0 = O.fill_state(1 << n, 6, 1.0)
    alpha = 0.8 + 0.1j
    ranks = [ShardedLazySum(build_q(Q, n, spec), r, world) for r in range(world)]
    xs = [torch.from_numpy(xfull[r << nloc:(r + 1) << nloc].copy()).cuda() for r in range(world)]
    ys = [torch.from_numpy(yfull0[r << nloc:(r + 1) << nloc].copy()).cuda() for r in range(world)]
    xptrs, yptrs = [t.data_ptr() for t in xs], [t.data_ptr() for t in ys]
    for y in ys:
        y.mul_(beta)          # y prepared on every rank before any contribution arrives
    order = list(range(world))
    for r in order[::2]:      # some ranks exchange before their local passes, some after
        ranks[r]._apply_ex(ranks[r].plan_swapped, alpha, None, 1.0, None, peers=(xptrs, yptrs), sm_budget=16)
    for r, sh in enumerate(ranks):
        sh._apply_ex(sh.plan_local, alpha, xs[r], 1.0, ys[r], sm_budget=-32)
        if sh.plan_local_b is not None:
            sh._apply_ex(sh.plan_local_b, alpha, xs[r], 1.0, ys[r], sm_budget=-32)
    for r in order[1::2]:
        ranks[r]._apply_ex(ranks[r].plan_swapped, alpha, None, 1.0, None, peers=(xptrs, yptrs), sm_budget=16)
    got = np.concatenate([y.cpu().numpy() for y in ys])
    assert H.rel_err(got, oracle_result(n, spec, alpha, beta, xfull, yfull0)) <= 1e-12


@pytest.mark.parametrize("world,n", [(4, 20), (4, 22)])
def test_chunks_cover_the_same_amplitudes_in_both_plans(world, n):
    """chunk c of the exchange pass writes exactly the contribution amplitudes that chunk c of the last local group reads"""
    import torch

    import qob200 as Q
    from qob200.dist import ShardedLazySum

    p, nloc = 2, n - 2
    spec = chain_spec(n, 29)
    ranks = [ShardedLazySum(build_q(Q, n, spec), r, world) for r in range(world)]
    if ranks[0].nchunks < 2:
        pytest.skip("plans are not single-pass at this size")
    nc, mask = ranks[0].nchunks, ranks[0].chunk_mask
    xs = [torch.randn(1 << nloc, dtype=torch.complex128, device="cuda") for _ in range(world)]
    zs = [torch.full((1 << nloc,), float("nan"), dtype=torch.complex128, device="cuda") for _ in range(world)]
    xptrs, zptrs = [t.data_ptr() for t in xs], [t.data_ptr() for t in zs]
    c = 1
    for r, sh in enumerate(ranks):
        sh._apply_ex(sh.plan_swapped, 1.0, None, 0.0, None, peers=(xptrs, zptrs), chunk=(c, nc))
    torch.cuda.synchronize()
    idx = torch.arange(1 << nloc, device="cuda")
    bits = [b for b in range(nloc) if mask >> b & 1]
    val = sum(((idx >> b) & 1) << i for i, b in enumerate(bits))
    for r in range(world):
        written = torch.isfinite(zs[r].real)
        assert torch.equal(written, val == c)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _nccl_worker(rank, world, port, n, beta, out_dir, fused):
    import torch
    import torch.distributed as dist

    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    try:
        import qob200 as Q
        from qob200.dist import ShardedLazySum

        p = world.bit_length() - 1
        nloc = n - p
        spec = chain_spec(n, 21)
        capi = fused in ("capi", "capi_z")
        if capi:   # everything behind the C ABI: qob_dist_* (IPC mapping, device barriers, choreography)
            from qob200.dist import DistLazySum

            # "capi": direct mode when the plan allows it (>= 2^20 amplitudes per rank: the exchange adds into the owners'
            # result slabs); "capi_z": contribution slabs
            sh = DistLazySum(build_q(Q, n, spec), rank, world, direct=None if fused == "capi" else False)
            assert sh.direct == (fused == "capi" and nloc >= 20)
            x = sh.x
        else:
            sh = ShardedLazySum(build_q(Q, n, spec), rank, world)
            x = sh.empty_state() if fused else torch.empty(1 << nloc, dtype=torch.complex128, device="cuda")
        Q.fill_state(x, 3, 2.0 ** (-n / 2), offset=rank << nloc)
        y = sh.y if capi and sh.direct else torch.empty(1 << nloc, dtype=torch.complex128, device="cuda")
        Q.fill_state(y, 4, 1.0, offset=rank << nloc)
        for _ in range(2):   # twice: the second call reuses the contribution buffer
            Q.fill_state(y, 4, 1.0, offset=rank << nloc)
            if capi:
                sh.mul_(y, 0.7 - 0.2j, beta)
            else:
                (sh.mul_fused_ if fused else sh.mul_)(y, x, 0.7 - 0.2j, beta)
        torch.cuda.synchronize()
        out = y.cpu().numpy()
        if capi:
            sh.close()
        np.save(os.path.join(out_dir, f"y{rank}.npy"), out)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [18, 21])   # 21 spins: the fused schedule runs its exchange / fold-in passes in 4 chunks
@pytest.mark.parametrize("fused", [False, True, "capi", "capi_z"])
@pytest.mark.parametrize("beta", [0.0, 0.5 + 0.25j])
def test_sharded_apply_nccl(tmp_path, beta, fused, n):
    import torch
    import torch.multiprocessing as mp

    world = 1
    while world * 2 <= min(torch.cuda.device_count(), 8):
        world *= 2
    if world < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    mp.spawn(_nccl_worker, args=(world, _free_port(), n, beta, str(tmp_path), fused), nprocs=world, join=True)
    got = np.concatenate([np.load(tmp_path / f"y{r}.npy") for r in range(world)])
    spec = chain_spec(n, 21)
    ref = oracle_result(n, spec, 0.7 - 0.2j, beta, O.fill_state(1 << n, 3, 2.0 ** (-n / 2)), O.fill_state(1 << n, 4, 1.0))
    assert H.rel_err(got, ref) <= 1e-12


def _spot(spec, index, xval):
    """(H x)[index] from the definition (each output amplitude of a 2-site-term LazySum depends on <= 1 + n_terms inputs)"""
    acc = 0.0 + 0.0j
    for c, idx, a in spec:
        A = PAULI[a]
        k1, k2 = idx[0] - 1, idx[1] - 1
        i1, i2 = (index >> k1) & 1, (index >> k2) & 1
        for j1 in (0, 1):
            for j2 in (0, 1):
                w = A[i1, j1] * A[i2, j2]
                if w != 0:
                    acc += c * w * xval((index & ~((1 << k1) | (1 << k2))) | (j1 << k1) | (j2 << k2))
    return acc


def _full_size_worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist

    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    try:
        import qob200 as Q
        from qob200.dist import ShardedLazySum

        p = world.bit_length() - 1
        free, _ = torch.cuda.mem_get_info()
        n = 33
        while 3 * 16 * (1 << (n - p)) > 0.85 * free and n > 20:   # x, y, contributions: the same rule as bench.py
            n -= 1
        nloc = n - p
        spec = chain_spec(n, 37)
        sh = ShardedLazySum(build_q(Q, n, spec), rank, world)
        seed, scale, al = 4321, 2.0 ** (-n / 2), 0.7 - 0.4j
        x = sh.empty_state()
        Q.fill_state(x, seed, scale, offset=rank << nloc)
        y = torch.full((1 << nloc,), float("nan"), dtype=torch.complex128, device="cuda")   # beta = 0 must not read y
        sh.mul_fused_(y, x, al, 0.0)
        sh.mul_fused_(y, x, al, 0.0)   # again: the contribution buffer is reused
        # ---- exact spot oracle: slab boundaries, tile / chunk boundaries, random places
        L = 1 << nloc
        rng = np.random.default_rng(500 + rank)
        loc = [0, 1, 4095, 4096, L - 1, L - 2, L // 2, L // 2 - 1, L // 4, 3 * (L // 4) - 1, 0x55555555 % L, 0x2AAAAAAA % L]
        loc += [int(v) for v in rng.integers(0, L, 52)]
        got = y[torch.tensor(loc, device="cuda")].cpu().numpy()
        ref = np.array([al * _spot(spec, (rank << nloc) | i, lambda j: O.state_at(seed, j, scale)) for i in loc])
        err = float(np.abs(got - ref).max() / np.abs(ref).max())
        # ---- Hermiticity: <x|Hx> summed over the ranks is real for real coefficients; all of y is finite
        d = torch.tensor([complex(Q.dot(x, y) / al)], dtype=torch.complex128, device="cuda")
        dr = torch.view_as_real(d).clone()
        dist.all_reduce(dr)
        nrm = torch.tensor([Q.norm2(y)], dtype=torch.float64, device="cuda")
        dist.all_reduce(nrm)
        with open(os.path.join(out_dir, f"r{rank}.txt"), "w") as f:
            f.write(f"{n} {err:.3e} {float(dr[0, 0]):.17g} {float(dr[0, 1]):.3e} {float(nrm[0]):.17g} {sh.nchunks}\n")
    finally:
        dist.destroy_process_group()


def test_sharded_apply_full_size_spot_oracle(tmp_path):
    """BASELINE config 5 at the size bench.py runs (N=33 on 4/8 GPUs, N=32 on 2): the fused sharded apply against the
    exact per-amplitude oracle at sampled indices of every slab, plus <x|Hx> real and a finite norm."""
    import torch
    import torch.multiprocessing as mp

    world = 1
    while world * 2 <= min(torch.cuda.device_count(), 8):
        world *= 2
    if world < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    mp.spawn(_full_size_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    lines = []
    for r in range(world):
        n, err, re, im, nrm, nch = open(tmp_path / f"r{r}.txt").read().split()
        assert float(err) <= 1e-12, f"rank {r}: spot oracle {err}"
        assert abs(float(im)) <= 1e-12 * max(1.0, abs(float(re)))
        assert np.isfinite(float(nrm)) and float(nrm) > 0
        lines.append(f"rank {r}: N={n} chunks={nch} spot_oracle_max_rel_err={err} <x|Hx>={re}{float(im):+.1e}i |y|^2={nrm}")
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, f"full_size_spot_{world}gpu.txt"), "w") as f:
            f.write("\n".join(lines) + "\n")


def test_two_devices_in_one_process():
    """One process, one qob_ctx per device (what the Julia glue's per-device CONTEXTS does): the shared-memory opt-in of the tile
    kernels and the SM count are per-device state, so the FIRST launch on the second device must configure it again."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    import qob200 as Q

    n = 21
    spec = chain_spec(n, 31)
    results = []
    for dev in (1, 0, 1):      # the second device first: nothing is configured for it yet
        torch.cuda.set_device(dev)
        Hs = build_q(Q, n, spec)
        b = Q.SpinBasis(0.5)
        B = Q.tensor(*[b] * n)
        x = Q.Ket(B)
        Q.fill_state(x.data, 9, 2.0 ** (-n / 2))
        y = Q.Ket(B)
        assert x.data.device.index == dev
        Q.mul_(y, Hs, x, 0.7 - 0.2j, 0.0)
        # the same operator through the round-1 tile kernel and the mixed-radix / gather kernels on this device
        e = Q.expect(Hs, x)
        torch.cuda.synchronize()
        results.append((y.data.cpu().numpy(), e))
    torch.cuda.set_device(0)
    for yv, e in results[1:]:
        assert H.rel_err(yv, results[0][0]) <= 1e-14
        assert abs(e - results[0][1]) <= 1e-12 * max(1.0, abs(results[0][1]))
    xfull = O.fill_state(1 << n, 9, 2.0 ** (-n / 2))
    ref = oracle_result(n, spec, 0.7 - 0.2j, 0.0, xfull, np.zeros(1 << n, dtype=complex))
    assert H.rel_err(results[0][0], ref) <= 1e-12


@pytest.mark.parametrize("n", [16, 20])
def test_dist_c_abi_single_rank(n):
    """qob_dist_* with world = 1 on any box: library-owned slab (qob_dist_alloc), bind, apply == the ordinary mul!."""
    import torch

    import qob200 as Q
    from qob200.dist import DistLazySum

    spec = chain_spec(n, 41)
    Hs = build_q(Q, n, spec)
    sh = DistLazySum(Hs, 0, 1)
    assert sh.nloc == n and sh.n_remote == 0
    Q.fill_state(sh.x, 3, 2.0 ** (-n / 2))
    y = torch.empty(1 << n, dtype=torch.complex128, device="cuda")
    Q.fill_state(y, 4, 1.0)
    sh.mul_(y, 0.7 - 0.2j, 0.5 + 0.25j)
    torch.cuda.synchronize()
    ref = oracle_result(n, spec, 0.7 - 0.2j, 0.5 + 0.25j, O.fill_state(1 << n, 3, 2.0 ** (-n / 2)), O.fill_state(1 << n, 4, 1.0))
    assert H.rel_err(y.cpu().numpy(), ref) <= 1e-12
    sh.close()

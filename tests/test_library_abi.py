"""CPU-side checks: the C-ABI library loads and exports every symbol include/qob200.h declares, fails loudly
without a GPU (no CPU fallback), validates constructor arguments like the reference, and its planner produces
the expected pass structure.  No compute calls are made here."""
import ctypes
import os
import re

import numpy as np
import pytest
import scipy.sparse as sp

import helpers as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def Q():
    import qob200

    return qob200


def header_functions():
    src = open(os.path.join(ROOT, "include", "qob200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(qob_[a-z0-9_]+)\s*\(", src)))


def test_every_header_symbol_is_exported(Q):
    lib = ctypes.CDLL(Q.LIB_PATH)
    names = header_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/qob200.h but not exported by libqob200.so"
    assert set(names) == set(Q.EXPORTED), set(names) ^ set(Q.EXPORTED)


def test_library_is_sm100a_only():
    import subprocess

    import qob200

    out = subprocess.run(["cuobjdump", "--list-elf", qob200.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_cpu_fallback(Q):
    import torch

    if torch.cuda.is_available():
        pytest.skip("this check is for the GPU-less container")
    h = ctypes.c_void_p()
    st = Q.lib.qob_ctx_create(0, ctypes.byref(h))
    assert st == 5 and b"no CPU fallback" in Q.lib.qob_last_error()
    with pytest.raises(Q.CudaError):
        Q.context()
    with pytest.raises(Q.CudaError):
        Q.Ket(Q.GenericBasis(4))
    # a planning-only context can build and describe plans but never computes
    ctx = Q.context(-1)
    b = Q.SpinBasis(0.5)
    B = Q.tensor(*[b] * 20)
    op = Q.LazySum([1.0], [Q.LazyTensor(B, [3, 4], (Q.sigmax(b), Q.sigmax(b)))])
    hnd = Q.handle(op, ctx)
    st = Q.lib.qob_op_apply(hnd, 0, Q._lib.c64.of(1), ctypes.c_void_p(4096), Q._lib.c64.of(0), ctypes.c_void_p(1 << 40), 1, None)
    assert st == 5 and b"no CPU fallback" in Q.lib.qob_last_error()


def test_status_strings(Q):
    assert Q.lib.qob_status_string(0) == b"ok"
    assert Q.lib.qob_status_string(1) == b"DimensionMismatch"
    assert Q.lib.qob_version() >= 100


def chain(Q, n, kind="heis"):
    b = Q.SpinBasis(0.5)
    B = Q.tensor(*[b] * n)
    sig = (Q.sigmax(b), Q.sigmay(b), Q.sigmaz(b))
    terms = []
    for i in range(1, n + 1):
        j = i % n + 1
        idx = sorted([i, j])
        if kind == "heis":
            terms += [Q.LazyTensor(B, idx, (s, s)) for s in sig]
        else:
            terms += [Q.LazyTensor(B, [i], (sig[0],)), Q.LazyTensor(B, idx, (sig[2], sig[2]))]
    return B, Q.LazySum([1.0] * len(terms), terms)


def test_planner_pass_structure(Q, monkeypatch):
    """N=28 Heisenberg chain: 28 bonds need 3 tile passes of 12 free bits (11 + 8 + 8 + wrap), 28 distinct flip masks.
    Round-2 kernel (qreg): passes 1 and 2 are chained through L2 in one launch (their free bits span 20 index bits =
    256 chunks of 256 tiles), 3 bonds per pass are register-resident, 19 gathers in total."""
    ctx = Q.context(-1)
    _, Hs = chain(Q, 28)
    d = Q.describe(Hs, ctx=ctx)
    assert "qreg[bits=28,T=12,passes=3,launches=2,components=56]" in d, d
    assert "{chained chunks=256 x 256 tiles, lag=1: [free:0-11 R:8,9,10,11" in d and "[free:0-2,11-19 R:16,17,18,19" in d, d
    assert "{single [free:0-2,19-27 R:24,25,26,27" in d, d
    # the lone pass (DRAM bound) runs first and only writes y; the chained pair (shared-memory bound) accumulates behind it
    assert d.index("{single") < d.index("{chained"), d
    assert sum(int(v) for v in re.findall(r"in-register:(\d+)", d)) == 9
    assert sum(int(v) for v in re.findall(r"gathers:(\d+)", d)) == 19
    # the per-GPU slab of config 5 (2^30 amplitudes): 4 tile passes in 2 chained launches = 2 trips through DRAM
    _, H30 = chain(Q, 30)
    d30 = Q.describe(H30, ctx=ctx)
    assert "qreg[bits=30,T=12,passes=4,launches=2" in d30 and d30.count("{chained") == 2, d30
    # density-matrix apply: 2N index bits, left terms on the low N, right terms (transposed) on the high N
    _, Hd = chain(Q, 10)
    assert "qreg[bits=20" in Q.describe(Hd, "left", 1 << 10, ctx=ctx)
    assert "qreg[bits=20" in Q.describe(Hd, "right", 1 << 10, ctx=ctx)
    assert "gather" in Q.describe(Hd, "left", 3, ctx=ctx)  # non power-of-two batch -> generic kernel
    _, Ht = chain(Q, 12, "tfim")
    assert "gather[terms=24,maxfac=2]" in Q.describe(Ht, ctx=ctx)
    # the round-1 tile kernel stays behind it (sharded layouts, states below 2^20 amplitudes)
    monkeypatch.setenv("QOB_DISABLE_QREG", "1")
    _, Hs = chain(Q, 28)
    d = Q.describe(Hs, ctx=ctx)
    assert "qtile[bits=28,T=12,L=3,passes=3,components=56]" in d, d
    # pass 0: 28 diagonal bonds collapse into 1 per-amplitude table + 5 per-thread tables; the 28 flip masks are spread
    # over the passes (bonds inside the low block may run in any pass and are moved off the shared-memory-bound pass 0)
    assert "lookups:1 diag+0 multi+" in d and ", 5 per thread" in d, d
    assert sum(int(v) for v in re.findall(r"multi\+(\d+) single", d)) == 28
    assert "{free:0-11" in d and "free:0-2,11-19" in d and "free:0-2,19-27" in d
    _, Hd = chain(Q, 10)
    assert "qtile[bits=20" in Q.describe(Hd, "left", 1 << 10, ctx=ctx)
    assert "qtile[bits=20" in Q.describe(Hd, "right", 1 << 10, ctx=ctx)


def test_planner_routes_dense_factors_to_dmma(Q):
    ctx = Q.context(-1)
    f1, f2, f3 = Q.FockBasis(47), Q.FockBasis(47), Q.NLevelBasis(3)
    B = Q.tensor(f1, f2, f3)
    rng = np.random.default_rng(0)
    A1, A2 = Q.Operator(f1, f1, H.rnd(rng, 48, 48)), Q.Operator(f2, f2, H.rnd(rng, 48, 48))
    op = Q.LazyTensor(B, [1, 2], (A1, A2))
    d = Q.describe(op, "left", 4096, ctx=ctx)
    assert d.count("dmma") == 2, d
    # sparse site operators of the same size stay on the fused gather kernel
    op2 = Q.LazyTensor(B, [1, 2], (Q.destroy(f1), Q.create(f2)))
    assert "gather" in Q.describe(op2, ctx=ctx)


def test_constructor_validation_mirrors_reference(Q):
    ctx = Q.context(-1)
    b2, b3 = Q.GenericBasis(2), Q.GenericBasis(3)
    B = Q.CompositeBasis([b2, b3])
    a2, a3 = Q.Operator(b2, b2, np.eye(2)), Q.Operator(b3, b3, np.eye(3))
    with pytest.raises(AssertionError):
        Q.LazyTensor(B, [2, 1], (a3, a2))            # issorted(indices), operators_lazytensor.jl:26
    with pytest.raises(AssertionError):
        Q.LazyTensor(B, [1], (a3,))                  # basis of the site operator, :30-31
    with pytest.raises(Q.ArgumentError):
        Q.LazyTensor(B, [3], (a3,))                  # check_indices
    with pytest.raises(Q.ArgumentError):
        Q.LazySum([1.0, 2.0], [Q.LazyTensor(B, [1], (a2,))])   # operators_lazysum.jl:46
    with pytest.raises(Q.IncompatibleBases):
        Q.LazySum(B, B, [1.0], [Q.LazyTensor(Q.CompositeBasis([b3, b2]), [1], (a3,))])
    with pytest.raises(Q.IncompatibleBases):
        Q.LazyProduct(Q.LazyTensor(B, [1], (a2,)), Q.LazyTensor(Q.CompositeBasis([b3, b2]), [1], (a3,)))
    with pytest.raises(Q.ArgumentError):
        Q.LazyProduct()
    with pytest.raises(Q.DimensionMismatch):
        Q.Operator(b2, b3, np.zeros((3, 2)))        # operators_dense.jl:16-17
    # the C ABI validates on its own as well (a Julia caller bypasses the python mirror)
    dl = (ctypes.c_int64 * 2)(2, 3)
    sites = (ctypes.c_int32 * 2)(2, 1)
    keep = []
    facs = (Q._lib.Factor * 2)(Q.operators._factor_struct(a3.data, keep), Q.operators._factor_struct(a2.data, keep))
    out = ctypes.c_void_p()
    st = Q.lib.qob_lazytensor_create(ctx, 2, dl, dl, 2, sites, facs, Q._lib.c64.of(1), ctypes.byref(out))
    assert st == 3 and b"sorted" in Q.lib.qob_last_error()
    sites = (ctypes.c_int32 * 1)(1)
    facs = (Q._lib.Factor * 1)(Q.operators._factor_struct(a3.data, keep))
    st = Q.lib.qob_lazytensor_create(ctx, 2, dl, dl, 1, sites, facs, Q._lib.c64.of(1), ctypes.byref(out))
    assert st == 3
    # unsupported factor type -> MethodError (test_operators_lazytensor.jl:409-415)
    class Weird(Q.operators.AbstractOperator):
        basis_l = basis_r = b2
    with pytest.raises(Q.MethodError):
        Q.handle(Q.LazyTensor(B, [1], (Weird(),)), ctx)
    # sparse x sparse is "not implemented" in the reference (sparsematrix.jl:177-188): no method here either
    with pytest.raises(Q.MethodError):
        Q.mul_(Q.Operator(b2, b2, sp.eye(2)), Q.Operator(b2, b2, sp.eye(2)), Q.Operator(b2, b2, sp.eye(2)))


def test_lazysum_dimension_check_in_c_abi(Q):
    ctx = Q.context(-1)
    b2, b3 = Q.GenericBasis(2), Q.GenericBasis(3)
    t = Q.LazyTensor(Q.CompositeBasis([b2, b3]), [1], (Q.Operator(b2, b2, np.eye(2)),))
    h = Q.handle(t, ctx)
    arr = (ctypes.c_void_p * 1)(h.value)
    cf = (Q._lib.c64 * 1)(Q._lib.c64.of(1))
    out = ctypes.c_void_p()
    assert Q.lib.qob_lazysum_create(ctx, 5, 5, 1, cf, arr, ctypes.byref(out)) == 1
    assert Q.lib.qob_lazysum_create(ctx, 6, 6, 1, cf, arr, ctypes.byref(out)) == 0
    assert Q.lib.qob_lazysum_set_coefs(out, 2, cf) == 3
    dl, dr = ctypes.c_int64(), ctypes.c_int64()
    assert Q.lib.qob_op_dims(out, ctypes.byref(dl), ctypes.byref(dr)) == 0 and (dl.value, dr.value) == (6, 6)
    Q.lib.qob_op_destroy(out)


def test_layout_plans_and_term_masks(Q):
    """host logic of the sharded apply: which terms are local, where the swap window goes"""
    from qob200.dist import ShardedLazySum, swap_window, swapped_bitpos

    ctx = Q.context(-1)
    n = 24
    _, Hs = chain(Q, n)
    h = Q.handle(Hs, ctx)
    od, al = ctypes.c_uint64(), ctypes.c_uint64()
    Q._lib.check(Q.lib.qob_lazysum_term_masks(h, 0, ctypes.byref(od), ctypes.byref(al)))   # sx sx on sites 1,2
    assert (od.value, al.value) == (0b11, 0b11)
    Q._lib.check(Q.lib.qob_lazysum_term_masks(h, 2, ctypes.byref(od), ctypes.byref(al)))   # sz sz on sites 1,2
    assert (od.value, al.value) == (0, 0b11)
    sh = ShardedLazySum(Hs, rank=5, world=8, ctx=ctx)
    # bonds (21,22),(22,23),(23,24),(24,1) have sigma_x/sigma_y on sharded bits 21..23 (0-based) -> 4 bonds x 2 terms
    assert sh.n_remote == 8 and sh.n_local == 3 * n - 8
    assert sh.swap_lo == 17 and swap_window(21, 3, (1 << 20) | 1) == 17
    assert swapped_bitpos(24, 21, 3, 17) == list(range(17)) + [21, 22, 23, 20, 17, 18, 19]
    d = sh.describe()
    assert "local[64 terms]" in d and "swapped[8 terms, window bit 17]" in d


def _bose_hubbard(Q, sites, cutoff):
    f = Q.FockBasis(cutoff)
    B = Q.tensor(*[f] * sites)
    a, ad, n = Q.destroy(f), Q.create(f), Q.number(f)
    terms = []
    for i in range(1, sites):
        terms += [Q.LazyTensor(B, [i, i + 1], (ad, a)), Q.LazyTensor(B, [i, i + 1], (a, ad))]
    terms += [Q.LazyTensor(B, [i], (n,)) for i in range(1, sites + 1)]
    return Q.LazySum([1.0] * len(terms), terms)


def test_mixed_radix_tile_planner(Q, monkeypatch):
    """Planning only: a Bose-Hubbard chain of 8 x Fock(7) (2^24 amplitudes) runs as 3 mixed-radix tile passes: the first
    four sites contiguously, then {site 1} + a window of three higher sites twice; every off-diagonal bond is in exactly one
    pass, the diagonal terms are spread; small states and a disabled planner keep the gather kernel."""
    ctx = Q.context(-1)
    d = Q.describe(_bose_hubbard(Q, 8, 7), ctx=ctx)
    assert "dtile[axes=8,passes=3]" in d and "gather[" not in d, d
    assert "{free axes:1,2,3,4 tile:4096 run:4096" in d and "{free axes:1,4,5,6 tile:4096 run:8" in d and "{free axes:1,6,7,8 tile:4096 run:8" in d, d
    assert sum(int(s.split(" ")[0]) for s in d.split("components:")[1:]) == 22 and d.count(" real}") == 3, d
    # X*op with 96 rows: the batch becomes two leading axes (32 x 3), 8 tensor axes in all
    assert "dtile[axes=10," in Q.describe(_bose_hubbard(Q, 8, 7), "right", 96, ctx=ctx)
    # op*rho with more columns than a tile holds: the batch axis is never free, one contiguous pass does it all
    d = Q.describe(_bose_hubbard(Q, 4, 7), "left", 5000, ctx=ctx)
    assert "dtile[axes=5,passes=1]" in d and "run:4096" in d, d
    # rho*op with 5000 rows: 50 x 100; a prime number of rows cannot be split and stays with the gather kernel
    assert "run:50 " in Q.describe(_bose_hubbard(Q, 4, 7), "right", 5000, ctx=ctx)
    assert "gather[" in Q.describe(_bose_hubbard(Q, 4, 7), "right", 4099, ctx=ctx)
    # below the size threshold, or switched off: the generic gather kernel
    assert "gather[" in Q.describe(_bose_hubbard(Q, 4, 3), ctx=ctx)
    monkeypatch.setenv("QOB_DISABLE_DTILE", "1")
    assert "gather[" in Q.describe(_bose_hubbard(Q, 8, 7), ctx=ctx)
    monkeypatch.delenv("QOB_DISABLE_DTILE")
    # a term that expands into more than 64 single-gather components makes the planner decline the whole sum
    import numpy as np

    g = Q.GenericBasis(9)
    Bg = Q.tensor(g, g, g, g, g, g)
    rng = np.random.default_rng(0)
    dense9 = [Q.Operator(g, g, rng.standard_normal((9, 9)) + 0j) for _ in range(2)]
    heavy = Q.LazyTensor(Bg, [1, 2], tuple(dense9))
    d = Q.describe(Q.LazySum([1.0], [heavy]), ctx=ctx)
    assert "gather[" in d and "dtile" not in d, d
    # ... and only that term when there are others
    light = Q.LazyTensor(Bg, [3, 4], (Q.Operator(g, g, np.diag(np.arange(9.0)) + 0j), Q.Operator(g, g, np.eye(9, k=1) + 0j)))
    d = Q.describe(Q.LazySum([1.0, 2.0], [heavy, light]), ctx=ctx)
    assert "dtile[" in d and "terms:1 components:1" in d and "gather[terms=1," in d, d


def test_mixed_radix_planner_fuzz_accounts_for_every_term(Q, monkeypatch):
    """Planning only, 150 random systems (dims 2..9 and a few large axes, 1-3 factor terms, shifted / diagonal / general
    sparse / small dense factors, both sides, random batch): the planner terminates, every term lands in exactly one tile
    pass or in the gather kernel, tiles stay within 4096 amplitudes."""
    import re

    import numpy as np
    import scipy.sparse as sp

    monkeypatch.setenv("QOB_DTILE_MIN_ELEMS", "1")
    ctx = Q.context(-1)
    rng = np.random.default_rng(2024)

    def factor(b, d):
        kind = rng.integers(0, 5)
        if kind == 0:
            m = sp.diags(rng.standard_normal(d)).tocsc()
        elif kind == 1:
            k = int(rng.integers(1, d)) * (1 if rng.random() < 0.5 else -1)
            m = sp.diags(rng.standard_normal(d - abs(k)) + 1.5, k, shape=(d, d)).tocsc()
        elif kind == 2:
            m = sp.random(d, d, density=min(1.0, 2.0 / d), random_state=np.random.RandomState(int(rng.integers(1 << 30))), format="csc")
        elif kind == 3 and d <= 4:
            m = rng.standard_normal((d, d))
        else:
            m = (sp.diags(np.ones(d - 1), 1) + sp.diags(np.ones(d - 1), -1)).tocsc()
        return Q.Operator(b, b, m.astype(complex) if sp.issparse(m) else m + 0j)

    seen_dtile = seen_gather = 0
    for it in range(150):
        n = int(rng.integers(2, 9))
        dims = [int(rng.choice([2, 3, 4, 5, 7, 8, 9])) for _ in range(n)]
        if rng.random() < 0.15:
            dims[int(rng.integers(0, n))] = int(rng.choice([64, 300, 2100, 5000]))
        bases = [Q.GenericBasis(d) for d in dims]
        B = Q.tensor(*bases) if n > 1 else bases[0]
        nterms = int(rng.integers(1, 10))
        terms = []
        for _ in range(nterms):
            k = int(rng.integers(1, min(3, n) + 1))
            idx = sorted(int(v) + 1 for v in rng.choice(n, size=k, replace=False))
            terms.append(Q.LazyTensor(B, idx, tuple(factor(bases[i - 1], dims[i - 1]) for i in idx)))
        S = Q.LazySum([1.0] * nterms, terms)
        side = "left" if rng.random() < 0.6 else "right"
        batch = int(rng.choice([1, 1, 3, 8, 96, 97, 640, 5000]))
        d = Q.describe(S, side, batch, ctx=ctx)
        in_tiles = sum(int(v) for v in re.findall(r" terms:(\d+) components", d))
        in_gather = sum(int(v) for v in re.findall(r"gather\[terms=(\d+),", d))
        in_seq = sum(int(v) for v in re.findall(r"seq\[terms=(\d+):", d))
        assert in_tiles + in_gather + in_seq == nterms, (dims, side, batch, d)
        assert all(int(v) <= 4096 for v in re.findall(r" tile:(\d+) ", d)), d
        seen_dtile += "dtile[" in d
        seen_gather += "gather[" in d
    assert seen_dtile >= 100 and seen_gather >= 5, (seen_dtile, seen_gather)


def test_lindblad_host_assembly(Q):
    """Planning only: the matrices the fused master-equation kernel works from — Heff = H - i/2 sum r_k J_k^+ J_k (by rows),
    G = H + i/2 sum r_k J_k^+ J_k (by columns), sqrt(r_k) J_k — against numpy, for a NON-Hermitian H, sparse and host-dense
    jump operators; the closed form -i(Heff rho - rho G) + sum J rho J^+ equals the reference's call pattern."""
    import numpy as np
    import scipy.sparse as sp

    from oracle import qob_oracle as O

    ctx = Q.context(-1)
    rng = np.random.default_rng(12)
    D = 23
    bas = Q.GenericBasis(D)

    def sprand(density):
        m = sp.random(D, D, density=density, random_state=np.random.RandomState(int(rng.integers(1 << 30))), format="csc").astype(complex)
        m.data = rng.standard_normal(m.nnz) + 1j * rng.standard_normal(m.nnz)
        return m

    Hm, J1, J2 = sprand(0.15), sprand(0.1), sprand(0.08)
    J3 = (rng.standard_normal((D, D)) + 1j * rng.standard_normal((D, D))) * (rng.uniform(0, 1, (D, D)) < 0.1)
    rates = [0.7, 0.0, 2.5]
    L = Q.LindbladRHS(Q.Operator(bas, bas, Hm), [Q.Operator(bas, bas, j) for j in (J1, J2, J3)], rates, ctx=ctx)
    Jd = [J1.toarray(), J2.toarray(), J3]
    JdJ = sum(r * (j.conj().T @ j) for r, j in zip(rates, Jd))
    Hd = Hm.toarray()
    assert np.abs(L.assembled(0) - (Hd - 0.5j * JdJ)).max() <= 1e-13
    assert np.abs(L.assembled(1) - (Hd + 0.5j * JdJ)).max() <= 1e-13
    for k in range(3):
        assert np.abs(L.assembled(2 + k) - np.sqrt(rates[k]) * Jd[k]).max() <= 1e-14
    rho = rng.standard_normal((D, D)) + 1j * rng.standard_normal((D, D))
    closed = -1j * (L.assembled(0) @ rho - rho @ L.assembled(1)) + sum(L.assembled(2 + k) @ rho @ L.assembled(2 + k).conj().T for k in range(3))
    assert np.abs(closed - O.lindblad_rhs(Hd, Jd, rho, rates)).max() <= 1e-12
    assert "lindblad 23x23" in L.describe() and "jumps=3" in L.describe()
    # errors: non-square H, mismatched jump operator, negative rate; no GPU -> apply refuses (no CPU fallback)
    with pytest.raises(Q.IncompatibleBases):
        Q.LindbladRHS(Q.Operator(bas, Q.GenericBasis(D + 1), sp.csc_matrix((D, D + 1), dtype=complex)), ctx=ctx)
    with pytest.raises(Q.ArgumentError):
        Q.LindbladRHS(Q.Operator(bas, bas, Hm), [Q.Operator(bas, bas, J1)], [-0.1], ctx=ctx)
    st = Q.lib.qob_lindblad_apply(L._handle, Q._lib.c64.of(1), ctypes.c_void_p(4096), Q._lib.c64.of(0), ctypes.c_void_p(1 << 40), None)
    assert st == 5 and b"no CPU fallback" in Q.lib.qob_last_error()


def test_ptrace_argument_checks_and_directsum_planning(Q):
    """check_ptrace_arguments (src/operators.jl:153-176) is enforced by the library before any device work, so it can be
    exercised on a planning-only context; LazyDirectSum handles add up their blocks' dimensions."""
    import ctypes as C

    from qob200 import _lib

    ctx = Q.context(-1)
    dl, dr = (C.c_int64 * 3)(2, 3, 4), (C.c_int64 * 3)(5, 3, 2)

    def call(tr):
        t = (C.c_int32 * max(len(tr), 1))(*tr)
        st = _lib.lib.qob_ptrace_op(ctx, 3, dl, dr, len(tr), t, None, None, None)
        return st, _lib.lib.qob_last_error().decode()

    assert call([1]) == (3, "Partial trace can only be applied onto subsystems that have the same left and right dimension.")
    assert call([1, 2, 3])[0] == 3 and "use tr() instead" in call([1, 2, 3])[1]
    assert call([4])[0] == 3 and call([2, 2])[0] == 3 and call([0])[0] == 3
    assert call([2])[0] == 5          # valid arguments: only the missing GPU stops it (no CPU fallback)
    st = _lib.lib.qob_ptrace_state(ctx, 3, dl, 3, (C.c_int32 * 3)(1, 2, 3), 0, None, None, None)
    assert st == 3
    # LazyDirectSum of a 3x3 sparse block and a 2-spin LazyTensor (4x4)
    b3, bs = Q.GenericBasis(3), Q.SpinBasis(0.5)
    import scipy.sparse as sp
    A = Q.SparseOperator(b3, b3, sp.identity(3, dtype=complex, format="csc"))
    B = Q.LazyTensor(Q.tensor(bs, bs), [1], (Q.sigmax(bs),))
    S = Q.LazyDirectSum(A, B)
    assert len(S.basis_l) == 7 and len(S.basis_r) == 7
    d = Q.describe(S, ctx=ctx)
    assert d.startswith("lazydirectsum[") and "sparse 3x3" in d, d
    S2 = Q.LazyDirectSum(S, A)      # nested sums are flattened (src/spinors.jl:166-168)
    assert len(S2.operators) == 3 and len(S2.basis_l) == 10


def test_dist_planning_behind_the_abi(Q):
    """qob_dist_create: term classification, swap window and chunking of the sharded apply are planned inside the library
    (planning-only context: no GPU needed); 8 ranks x 2^30 amplitudes = BASELINE config 5."""
    import ctypes as C

    from qob200 import _lib
    from qob200.operators import handle

    ctx = Q.context(-1)
    _, Hs = chain(Q, 33)
    h = handle(Hs, ctx)
    d = C.c_void_p()
    _lib.check(_lib.lib.qob_dist_create(h, 5, 8, C.byref(d)))
    nloc, nrem, nch = C.c_int32(), C.c_int32(), C.c_int32()
    slab, flagb = C.c_int64(), C.c_int64()
    _lib.check(_lib.lib.qob_dist_info(d, C.byref(nloc), C.byref(nrem), C.byref(nch), C.byref(slab), C.byref(flagb)))
    assert (nloc.value, slab.value) == (30, 16 << 30) and flagb.value >= 64
    assert nrem.value == 8          # XX and YY of the four bonds that touch a sharded axis; their ZZ parts are rank-dependent weights
    assert nch.value == 4
    buf = C.create_string_buffer(1 << 15)
    _lib.check(_lib.lib.qob_dist_describe(d, buf, len(buf)))
    text = buf.value.decode()
    assert text.startswith("dist[rank 5/8, 2^30 amplitudes per rank, 91 local + 8 exchanged terms, chunks=4]"), text[:200]
    assert "exchanged (window bit" in text and "qreg[bits=30" in text
    # apply without bound buffers / on a planning-only context fails loudly
    assert _lib.lib.qob_dist_apply(d, _lib.c64.of(1.0), _lib.c64.of(0.0), C.c_void_p(16), None) == 3
    assert _lib.lib.qob_dist_create(h, 3, 6, C.byref(C.c_void_p())) == 3      # world must be a power of two
    # direct mode: the exchange pass of this plan can add into the owners' result slabs (round-2 kernel, 4 KiB pieces)
    cap = C.c_int32(-1)
    _lib.check(_lib.lib.qob_dist_direct_capable(d, C.byref(cap)))
    assert cap.value == 1 and "direct mode available" in text
    assert "exchanged (window bit 26)" in text and "free:0-7,26-29" in text
    assert _lib.lib.qob_dist_bind_result(d, None) == 3
    ms, cnt, nb = C.c_double(-1), C.c_int32(-1), C.c_int64()
    _lib.check(_lib.lib.qob_dist_exchange_timing(d, 1))
    _lib.check(_lib.lib.qob_dist_exchange_ms(d, C.byref(ms), C.byref(cnt), C.byref(nb)))
    assert (ms.value, cnt.value) == (0.0, 0) and nb.value == int(2 * (7 / 8) * 16 * 2 ** 30)
    _lib.check(_lib.lib.qob_dist_destroy(d))
    # small slabs: the round-2 kernel is not planned below 2^20 amplitudes per rank, so no direct mode
    _, Hs = chain(Q, 18)
    d2 = C.c_void_p()
    _lib.check(_lib.lib.qob_dist_create(handle(Hs, ctx), 1, 2, C.byref(d2)))
    _lib.check(_lib.lib.qob_dist_direct_capable(d2, C.byref(cap)))
    assert cap.value == 0
    assert _lib.lib.qob_dist_bind_result(d2, (C.c_void_p * 2)(16, 32)) == 4   # QOB_STATUS_UNSUPPORTED
    _lib.check(_lib.lib.qob_dist_destroy(d2))

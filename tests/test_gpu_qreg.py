"""Parity of the round-2 tile kernel (qob_kernels_qreg.cu: register-blocked, TMA-staged, pass pairs chained through L2)
against the oracle's restatement of the reference's per-term sparse recursion (operators_lazysum.jl:189-200 ->
operators_lazytensor.jl:652-685), and against the independent generic gather kernel on full vectors.  The kernel is
forced on for small chains (QOB_QREG_MIN_BITS) so that the oracle finishes in seconds; chunk size and queue lag of the
chained launches are swept through their corner cases (one chunk, many chunks, lag > 1, chaining off)."""
import numpy as np
import pytest
import scipy.sparse as sp

import helpers as H
from helpers import O

pytestmark = pytest.mark.gpu
TOL = 1e-12


@pytest.fixture(scope="module")
def Q():
    import qob200

    return qob200


def _pauli():
    sx = np.array([[0, 1], [1, 0]], dtype=complex)
    sy = np.array([[0, -1j], [1j, 0]], dtype=complex)
    sz = np.array([[1, 0], [0, -1]], dtype=complex)
    return sx, sy, sz


def _chain(n, kind, rng, complex_coefs=False):
    sx, sy, sz = _pauli()
    dims = (2,) * n
    terms, coefs = [], []

    def cf():
        c = rng.uniform(0.5, 1.5)
        return c * np.exp(1j * rng.uniform(0, 6.28)) if complex_coefs else c

    for i in range(1, n + 1):
        j = i % n + 1
        idx = sorted([i, j])
        if kind == "tfim":
            terms.append(H.lazytensor(dims, dims, [i], [sp.csc_matrix(sx)]))
            coefs.append(-cf())
            terms.append(H.lazytensor(dims, dims, idx, [sp.csc_matrix(sz), sp.csc_matrix(sz)]))
            coefs.append(-cf())
        else:
            for s in (sx, sy, sz):
                terms.append(H.lazytensor(dims, dims, idx, [sp.csc_matrix(s), sp.csc_matrix(s)]))
                coefs.append(cf())
    return dims, coefs, terms


@pytest.mark.parametrize("n", [12, 13, 14, 16, 17, 18])
@pytest.mark.parametrize("kind", ["tfim", "heis"])
def test_qreg_chain_vs_oracle(Q, monkeypatch, n, kind):
    monkeypatch.setenv("QOB_QREG_MIN_BITS", "12")
    rng = np.random.default_rng(500 + n)
    dims, coefs, terms = _chain(n, kind, rng)
    s = H.lazysum(dims, dims, coefs, terms)
    assert "qreg" in Q.describe(s.q), Q.describe(s.q)
    H.check_mul(s, dims, dims, rng, tol=TOL, kinds=("ket", "bra"))
    if n <= 14:
        H.check_mul(s, dims, dims, rng, tol=TOL, kinds=("opl", "opr"), nbatch=4, scalars=((1, 0), (1.5, 2.1)))


@pytest.mark.parametrize("chain_bits,lag_tiles", [(0, 192), (13, 192), (14, 1), (15, 8), (16, 192), (20, 192), (20, 1)])
def test_qreg_chunked_chaining_corner_cases(Q, monkeypatch, chain_bits, lag_tiles):
    """n = 18: three tile passes; the chunk size (union of the free bits of a chained pair) and the lag of the second pass
    decide how the tile queue is ordered — every ordering must give the same result."""
    monkeypatch.setenv("QOB_QREG_MIN_BITS", "12")
    monkeypatch.setenv("QOB_QREG_CHAIN_BITS", str(chain_bits))
    monkeypatch.setenv("QOB_QREG_LAG_TILES", str(lag_tiles))
    rng = np.random.default_rng(640 + chain_bits)
    n = 18
    dims, coefs, terms = _chain(n, "heis", rng, complex_coefs=(chain_bits % 2 == 1))
    s = H.lazysum(dims, dims, coefs, terms)
    d = Q.describe(s.q)
    assert "qreg" in d, d
    H.check_mul(s, dims, dims, rng, tol=TOL, kinds=("ket",), scalars=((1, 0), (0.3 - 0.2j, 1.7), (-1j, 0)))


def test_qreg_general_2x2_factors_and_three_site_terms(Q, monkeypatch):
    monkeypatch.setenv("QOB_QREG_MIN_BITS", "12")
    rng = np.random.default_rng(660)
    n = 15
    dims = (2,) * n
    terms, coefs = [], []
    for idx in ([1], [7], [15], [1, 15], [3, 9], [2, 3, 4], [1, 8, 15], [12, 13], [9, 10, 11], [10, 12], [11, 12], [13, 14, 15]):
        datas = [H.rnd(rng, 2, 2) if rng.uniform() < 0.5 else H.sprnd(rng, 2, 2, 0.7) for _ in idx]
        terms.append(H.lazytensor(dims, dims, idx, datas, rng.uniform(0.5, 1.0)))
        coefs.append(H.rnd(rng, 1)[0])
    s = H.lazysum(dims, dims, coefs, terms)
    assert "qreg" in Q.describe(s.q)
    H.check_mul(s, dims, dims, rng, tol=TOL, kinds=("ket", "bra"))


@pytest.mark.parametrize("seed", range(10))
def test_qreg_random_term_sets_vs_oracle(Q, monkeypatch, seed):
    """random chain lengths, 1-3-site terms on random sites with random dense / sparse / diagonal / adjoint 2x2 factors,
    repeated site sets (several selector sets per mask), real or complex coefficients"""
    rng = np.random.default_rng(7000 + seed)
    n = int(rng.integers(12, 18))
    monkeypatch.setenv("QOB_QREG_MIN_BITS", "12")
    monkeypatch.setenv("QOB_QREG_CHAIN_BITS", str(int(rng.integers(12, 21))))
    dims = (2,) * n
    terms, coefs = [], []
    site_sets = []
    real = seed % 3 == 0
    for _ in range(int(rng.integers(5, 26))):
        k = int(rng.integers(1, 4))
        if site_sets and rng.uniform() < 0.3:
            idx = site_sets[int(rng.integers(len(site_sets)))]
        else:
            idx = sorted(int(v) for v in rng.choice(np.arange(1, n + 1), size=k, replace=False))
            site_sets.append(idx)
        datas = []
        for _s in idx:
            kind = rng.integers(4)
            m = H.rnd(rng, 2, 2)
            if real:
                m = m.real.astype(complex)
            if kind == 0:
                datas.append(m)
            elif kind == 1:
                datas.append(sp.csc_matrix(m * (rng.uniform(0, 1, (2, 2)) < 0.6)))
            elif kind == 2:
                datas.append(sp.csc_matrix(np.diag(np.diag(m))))
            else:
                datas.append(("adj", m))
        terms.append(H.lazytensor(dims, dims, idx, datas, rng.uniform(0.5, 1.0) if real else H.rnd(rng, 1)[0]))
        coefs.append(rng.uniform(-1, 1) if real else H.rnd(rng, 1)[0])
    s = H.lazysum(dims, dims, coefs, terms)
    H.check_mul(s, dims, dims, rng, tol=TOL, kinds=("ket", "bra"), scalars=((1, 0), (0.3 - 0.2j, 1.7)))


@pytest.mark.parametrize("n", [9, 10])
def test_qreg_density_matrix_commutator(Q, monkeypatch, n):
    """-i[H, rho] on a full 2^n x 2^n density matrix: left and right application run as tile passes over 2n index bits"""
    monkeypatch.setenv("QOB_QREG_MIN_BITS", "12")
    rng = np.random.default_rng(777 + n)
    dims, coefs, terms = _chain(n, "heis", rng)
    s = H.lazysum(dims, dims, coefs, terms)
    D = 1 << n
    assert "qreg[bits=%d" % (2 * n) in Q.describe(s.q, "left", D) and "qreg[bits=%d" % (2 * n) in Q.describe(s.q, "right", D)
    rho = H.rnd(rng, D, D)
    st, r = H.denseop(dims, dims, rho), H.denseop(dims, dims, np.zeros((D, D), dtype=complex))
    O.mul(r.o, s.o, st.o, -1j, 0)
    O.mul(r.o, st.o, s.o, 1j, 1)
    Q.mul_(r.q, s.q, st.q, -1j, 0)
    Q.mul_(r.q, st.q, s.q, 1j, 1)
    assert H.rel_err(r.q.to_host(), r.o.data) <= TOL


def test_qreg_nan_kill_and_coefficient_updates(Q, monkeypatch):
    """beta == 0 must not read y (operators_lazytensor.jl:719-720); new coefficients only refill the weight tables"""
    monkeypatch.setenv("QOB_QREG_MIN_BITS", "12")
    rng = np.random.default_rng(690)
    n = 16
    dims, coefs, terms = _chain(n, "heis", rng)
    s = H.lazysum(dims, dims, coefs, terms)
    x = H.rnd(rng, 1 << n)
    st = H.ket(dims, x)
    r = H.ket(dims, np.full(1 << n, np.nan + 1j * np.nan))
    O.mul(r.o, s.o, st.o, 0.5 + 0.1j, 0)
    Q.mul_(r.q, s.q, st.q, 0.5 + 0.1j, 0)
    assert H.rel_err(r.q.to_host(), r.o.data) <= TOL
    used = None
    for trial in range(3):
        newc = [complex(c * rng.uniform(0.5, 2.0) * (1j if trial == 1 else 1.0)) for c in coefs]   # trial 1: complex weight tables
        so = O.LazySum(dims, dims, newc, [t_.o for t_ in terms])
        s.q.factors[:] = newc
        O.mul(r.o, so, st.o, 1.0, 0.25)
        before = Q.launch_count()
        Q.mul_(r.q, s.q, st.q, 1.0, 0.25)
        n_launch = Q.launch_count() - before
        used = n_launch if used is None else used
        assert n_launch == used             # same plan every time: only the weight tables are refilled
        assert H.rel_err(r.q.to_host(), r.o.data) <= TOL


@pytest.mark.parametrize("n", [22, 24])
def test_qreg_vs_gather_kernel_full_vector(Q, monkeypatch, n):
    """default settings at sizes with many chunks: the chained tile passes against the independent gather kernel"""
    import torch

    b = Q.SpinBasis(0.5)
    B = Q.tensor(*[b] * n)
    sig = (Q.sigmax(b), Q.sigmay(b), Q.sigmaz(b))
    rng = np.random.default_rng(800 + n)
    coefs = list(rng.uniform(0.5, 1.5, 3 * n))

    def build():
        terms = []
        for i in range(1, n + 1):
            j = i % n + 1
            for s in sig:
                terms.append(Q.LazyTensor(B, sorted([i, j]), (s, s)))
        return Q.LazySum(coefs, terms)

    Ht = build()
    assert "qreg" in Q.describe(Ht), Q.describe(Ht)
    x = Q.randstate(B, seed=5)
    yt = Q.Ket(B)
    yt.data.fill_(float("nan"))
    Q.mul_(yt, Ht, x, 0.7 - 0.2j, 0.0)
    Q.mul_(yt, Ht, x, 0.1j, 1.0)   # accumulate: first launch read-modify-writes y
    monkeypatch.setenv("QOB_DISABLE_QREG", "1")
    monkeypatch.setenv("QOB_DISABLE_QTILE", "1")
    Hg = build()
    assert "qreg" not in Q.describe(Hg) and "qtile" not in Q.describe(Hg)
    yg = Q.Ket(B)
    Q.mul_(yg, Hg, x, 0.7 - 0.1j, 0.0)
    diff = yt.data - yg.data
    assert np.sqrt(Q.norm2(diff) / Q.norm2(yg.data)) <= TOL
    torch.cuda.synchronize()


@pytest.mark.parametrize("world,n", [(2, 14), (4, 16), (8, 17)])
@pytest.mark.parametrize("beta", [0.0, 0.5 + 0.25j])
def test_qreg_in_sharded_layout_plans(monkeypatch, world, n, beta):
    """The communication-free groups and the swapped-layout group of a sharded apply, with every rank's tile programs run on ONE
    GPU: plain launches of a layout plan go through the round-2 kernel (rank-dependent diagonal weights enter through the index
    bits above the local address)."""
    monkeypatch.setenv("QOB_QREG_MIN_BITS", "12")
    import test_gpu_dist as TD

    import qob200 as Q
    from qob200.dist import ShardedLazySum

    sh = ShardedLazySum(TD.build_q(Q, n, TD.chain_spec(n, 21)), 1, world)
    assert "qreg[" in sh.describe(), sh.describe()
    TD.test_sharded_apply_all_ranks_on_one_gpu(world, n, beta)
    # the same with the fused peer exchange: the local group then runs the two-buffer variant that leaves room for the
    # exchange kernel's CTAs on every SM (sm_budget > 0)
    TD.test_fused_peer_exchange_all_ranks_on_one_gpu(world, n, beta)


@pytest.mark.parametrize("n,force_old", [(22, False), (18, True)])
def test_coefficient_updates_are_ordered_across_streams(Q, monkeypatch, n, force_old):
    """One handle, two streams, coefficients changed between the launches (TimeDependentSum set_time! with per-task streams):
    the weight tables are one device buffer per program, so the upload for the second launch must wait for the first launch's
    kernels, and the second stream must see the upload — whatever stream it was issued on."""
    import torch

    if force_old:
        monkeypatch.setenv("QOB_DISABLE_QREG", "1")
        monkeypatch.setenv("QOB_QTILE_MIN_BITS", "10")
    b = Q.SpinBasis(0.5)
    B = Q.tensor(*[b] * n)
    sig = (Q.sigmax(b), Q.sigmay(b), Q.sigmaz(b))
    rng = np.random.default_rng(1234 + n)
    terms = []
    for i in range(1, n + 1):
        j = i % n + 1
        for s in sig:
            terms.append(Q.LazyTensor(B, sorted([i, j]), (s, s)))
    c1 = list(rng.uniform(0.5, 1.5, len(terms)))
    c2 = list(rng.uniform(-1.5, -0.5, len(terms)))
    Hs = Q.LazySum(list(c1), terms)
    x = Q.randstate(B, seed=11)
    ref1, ref2 = Q.Ket(B), Q.Ket(B)
    Q.mul_(ref1, Hs, x, 1.0, 0.0)
    Hs.factors[:] = c2
    Q.mul_(ref2, Hs, x, 1.0, 0.0)
    torch.cuda.synchronize()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for trial in range(6):
        y1, y2 = Q.Ket(B), Q.Ket(B)
        torch.cuda.synchronize()
        Hs.factors[:] = c1
        with torch.cuda.stream(s1):
            for _ in range(3):
                Q.mul_(y1, Hs, x, 1.0, 0.0)
        Hs.factors[:] = c2
        with torch.cuda.stream(s2):
            Q.mul_(y2, Hs, x, 1.0, 0.0)
        with torch.cuda.stream(s1):      # equal coefficients again on the first stream: it must see stream 2's upload
            y3 = Q.Ket(B)
            Q.mul_(y3, Hs, x, 1.0, 0.0)
        torch.cuda.synchronize()
        assert np.sqrt(Q.norm2(y1.data - ref1.data) / Q.norm2(ref1.data)) <= TOL
        assert np.sqrt(Q.norm2(y2.data - ref2.data) / Q.norm2(ref2.data)) <= TOL
        assert np.sqrt(Q.norm2(y3.data - ref2.data) / Q.norm2(ref2.data)) <= TOL


@pytest.mark.parametrize("n,kind", [(13, "tfim"), (18, "heis"), (16, "heis")])
def test_qreg_fused_expect(Q, monkeypatch, n, kind):
    """expect(op, psi) on the round-2 kernel: every tile pass reduces conj(x) * (its share of op x) — no result vector is written,
    no separate dot product; the launch count is the plan's launch count and the value is reproducible bit for bit."""
    import re

    monkeypatch.setenv("QOB_QREG_MIN_BITS", "12")
    rng = np.random.default_rng(1500 + n)
    dims, coefs, terms = _chain(n, kind, rng, complex_coefs=(n == 16))
    s = H.lazysum(dims, dims, coefs, terms)
    d = Q.describe(s.q)
    assert "qreg[" in d
    nlaunch = int(re.search(r"launches=(\d+)", d).group(1))
    x = H.rnd(rng, 1 << n)
    x /= np.linalg.norm(x)
    hx = H.ket(dims, np.zeros(1 << n, dtype=complex))
    O.mul(hx.o, s.o, H.ket(dims, x).o, 1.0, 0.0)
    ref = np.vdot(x, hx.o.data)
    st = H.ket(dims, x).q
    Q.expect(s.q, st)                      # builds the handle, loads the tables
    l0 = Q.launch_count()
    e1 = Q.expect(s.q, st)
    assert Q.launch_count() - l0 == nlaunch, (Q.launch_count() - l0, nlaunch)
    e2 = Q.expect(s.q, st)
    assert e1 == e2                        # static tile assignment + fixed-order reduction
    assert abs(e1 - ref) <= 1e-12 * max(1.0, abs(ref))

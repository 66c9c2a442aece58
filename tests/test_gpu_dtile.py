"""Mixed-radix tile passes (qob_kernels_dtile.cu): the fused LazySum / LazyTensor kernel for large states on subsystems of
any dimension, against the oracle's restatement of the reference (per-term sparse recursion / dense-factor path).
The kernel is forced on for small states (QOB_DTILE_MIN_ELEMS=1) so that the oracle finishes in seconds."""
import numpy as np
import pytest
import scipy.sparse as sp

import helpers as H
from helpers import O

pytestmark = pytest.mark.gpu
TOL = 1e-12


@pytest.fixture(params=["one-cta-per-tile", "persistent"])
def Q(monkeypatch, request):
    """every test runs with both launch schemes: one CTA per tile, and the persistent double-buffered CTAs (forced on
    here; by default they are used once there are >= 8 tiles per SM)"""
    import qob200

    monkeypatch.setenv("QOB_DTILE_MIN_ELEMS", "1")
    monkeypatch.setenv("QOB_DTILE_PERSIST", "2" if request.param == "persistent" else "0")
    return qob200


_USED = []


def _shift_op(rng, d, s):
    """one shifted diagonal (like destroy / create / transition operators)"""
    m = sp.lil_matrix((d, d), dtype=complex)
    for i in range(d):
        if 0 <= i + s < d:
            m[i, i + s] = complex(rng.uniform(0.5, 1.5), rng.uniform(-1, 1))
    return sp.csc_matrix(m)


def _random_factor(rng, d):
    kind = rng.integers(0, 6)
    if kind == 0:
        return sp.csc_matrix(np.diag(H.rnd(rng, d)))                       # diagonal (number-like)
    if kind == 1:
        return _shift_op(rng, d, int(rng.integers(1, d)) * (1 if rng.random() < 0.5 else -1))
    if kind == 2:
        return H.sprnd(rng, d, d, 0.35)                                     # general sparse
    if kind == 3 and d <= 4:
        return H.rnd(rng, d, d)                                             # small dense factor
    if kind == 4:
        return ("adj", H.sprnd(rng, d, d, 0.4))                             # lazy adjoint
    return _shift_op(rng, d, 1) + _shift_op(rng, d, -1)                     # x-like: two diagonals


@pytest.mark.parametrize("seed", range(12))
def test_random_mixed_radix_lazysum(Q, seed):
    rng = np.random.default_rng(700 + seed)
    n = int(rng.integers(3, 7))
    dims = tuple(int(rng.choice([2, 3, 4, 5, 7])) for _ in range(n))
    nterms = int(rng.integers(1, 9))
    pairs, coefs = [], []
    for _ in range(nterms):
        k = int(rng.integers(1, min(3, n) + 1))
        idx = sorted(int(v) + 1 for v in rng.choice(n, size=k, replace=False))
        pairs.append(H.lazytensor(dims, dims, idx, [_random_factor(rng, dims[i - 1]) for i in idx], factor=complex(rng.uniform(0.5, 1.5), 0.3)))
        coefs.append(complex(rng.uniform(-1, 1), rng.uniform(-1, 1)))
    s = H.lazysum(dims, dims, coefs, pairs)
    d = Q.describe(s.q)
    assert "dtile" in d or "gather[" in d, d     # terms the tile planner declines run through the gather kernel
    _USED.append("dtile" in d)
    H.check_mul(s, dims, dims, rng, tol=TOL, nbatch=int(rng.choice([1, 3, 8, 24])))


def test_random_sweep_mostly_used_the_tile_passes():
    """(runs after the sweep) single terms that expand into > 64 components are declined, the rest of the sum is taken"""
    assert len(_USED) == 24 and sum(_USED) >= 20, _USED


@pytest.mark.parametrize("sites,cutoff", [(6, 4), (4, 9), (7, 7)])
def test_bose_hubbard_chain(Q, sites, cutoff):
    """hopping a_i^dag a_{i+1} + h.c. and on-site n(n-1): the site operators of src/fock.jl on a Fock lattice"""
    rng = np.random.default_rng(31)
    d = cutoff + 1
    dims = (d,) * sites
    a, ad, num = O.destroy(cutoff).data, O.create(cutoff).data, O.number(cutoff).data
    nn = sp.csc_matrix(num @ num - num)
    pairs, coefs = [], []
    for i in range(1, sites):
        pairs += [H.lazytensor(dims, dims, [i, i + 1], [ad, a]), H.lazytensor(dims, dims, [i, i + 1], [a, ad])]
        coefs += [-1.0, -1.0]
    for i in range(1, sites + 1):
        pairs.append(H.lazytensor(dims, dims, [i], [nn]))
        coefs.append(0.5 + 0.1 * i)
    s = H.lazysum(dims, dims, coefs, pairs)
    desc = Q.describe(s.q)
    assert "dtile" in desc
    kinds = ("ket",) if d ** sites > 100000 else ("ket", "bra", "opl", "opr")
    H.check_mul(s, dims, dims, rng, tol=TOL, kinds=kinds, nbatch=2, scalars=((1, 0), (-1j, 0.5)))
    # TimeDependentSum-style coefficient update reaches the tile programs without replanning
    coefs2 = [c * (1.5 - 0.5j) for c in coefs]
    s2 = H.lazysum(dims, dims, coefs2, pairs)
    s.q.factors = list(coefs2)          # what TimeDependentSum.set_time_ does (src/time_dependent_operator.jl:279-290)
    x = H.rnd(rng, d ** sites)
    r = H.ket(dims, np.zeros(d ** sites, dtype=complex))
    O.mul(r.o, s2.o, H.ket(dims, x).o, 1.0, 0.0)
    Q.mul_(r.q, s.q, H.ket(dims, x).q, 1.0, 0.0)
    assert H.rel_err(r.q.to_host(), r.o.data) <= TOL


def test_single_lazytensor_and_nan_kill(Q):
    """a single LazyTensor (not inside a LazySum) takes the same path; beta = 0 must not read y; alpha = 0 only scales"""
    rng = np.random.default_rng(41)
    dims = (3, 4, 5, 3)
    op = H.lazytensor(dims, dims, [2, 4], [_shift_op(rng, 4, 1), H.sprnd(rng, 3, 3, 0.5)], factor=0.3 - 0.7j)
    assert "dtile" in Q.describe(op.q)
    H.check_mul(op, dims, dims, rng, tol=TOL)
    D = int(np.prod(dims))
    x = H.rnd(rng, D)
    ref = H.ket(dims, np.zeros(D, dtype=complex))
    O.mul(ref.o, op.o, H.ket(dims, x).o, 0.7, 0.0)
    r = H.ket(dims, np.full(D, np.nan + 0j))
    Q.mul_(r.q, op.q, H.ket(dims, x).q, 0.7, 0.0)
    out = r.q.to_host()
    assert np.all(np.isfinite(out)) and H.rel_err(out, ref.o.data) <= TOL


def test_right_side_with_long_batch(Q):
    """X*op with 96 rows: the batch axis is the fastest tensor axis and is split (32 x 3) so that a short run of it is
    the coalesced low block; 97 rows (prime) cannot be split and stays whole"""
    rng = np.random.default_rng(43)
    dims = (4, 3, 5)
    pairs = [H.lazytensor(dims, dims, [1, 2], [_shift_op(rng, 4, -1), _shift_op(rng, 3, 1)]),
             H.lazytensor(dims, dims, [3], [H.sprnd(rng, 5, 5, 0.5)]),
             H.lazytensor(dims, dims, [2, 3], [sp.csc_matrix(np.diag(H.rnd(rng, 3))), _shift_op(rng, 5, 2)])]
    s = H.lazysum(dims, dims, [0.5, -1.2j, 0.8], pairs)
    for nb in (96, 97):
        assert "dtile" in Q.describe(s.q, "right", nb)
        H.check_mul(s, dims, dims, rng, tol=TOL, kinds=("opr", "opl"), nbatch=nb, scalars=((1.5, 2.1), (1, 0)))


def test_declines_to_gather_when_it_does_not_fit(Q):
    """two dense 8x8 factors expand into 225 components: the planner declines, the generic gather kernel takes over"""
    rng = np.random.default_rng(47)
    dims = (8, 8, 3)
    op = H.lazytensor(dims, dims, [1, 2], [H.rnd(rng, 8, 8), H.rnd(rng, 8, 8)])
    d = Q.describe(op.q)
    assert "gather[" in d and "dtile" not in d
    H.check_mul(op, dims, dims, rng, tol=TOL, kinds=("ket", "opl"), nbatch=3, scalars=((1, 0),))


def test_heavy_term_beside_tile_terms(Q):
    """one term with two dense 8x8 factors (225 components) inside a Bose-Hubbard-like sum: that term goes through the
    gather kernel, the others through the tile passes, beta is applied once"""
    rng = np.random.default_rng(53)
    dims = (8, 8, 4, 3)
    pairs = [H.lazytensor(dims, dims, [1, 2], [H.rnd(rng, 8, 8), H.rnd(rng, 8, 8)]),
             H.lazytensor(dims, dims, [2, 3], [_shift_op(rng, 8, 1), _shift_op(rng, 4, -1)]),
             H.lazytensor(dims, dims, [4], [sp.csc_matrix(np.diag(H.rnd(rng, 3)))]),
             H.lazytensor(dims, dims, [1, 4], [_shift_op(rng, 8, -2), H.sprnd(rng, 3, 3, 0.6)])]
    s = H.lazysum(dims, dims, [0.5, -1.2j, 0.8, 1.1], pairs)
    d = Q.describe(s.q)
    assert "dtile" in d and "gather[terms=1," in d, d
    H.check_mul(s, dims, dims, rng, tol=TOL, nbatch=3)


def test_density_matrix_commutator_default_threshold(monkeypatch):
    """-i[H, rho] for a 3-site Bose-Hubbard H (D = 512) on a dense 512 x 512 rho as the two mul! calls of a master
    equation: 2^18 amplitudes, so both sides take the tile passes with the default settings (rho*H: the 512 rows are the
    fastest axis, split 64 x 8)"""
    import qob200 as Q

    monkeypatch.delenv("QOB_DTILE_MIN_ELEMS", raising=False)
    monkeypatch.delenv("QOB_DTILE_PERSIST", raising=False)
    rng = np.random.default_rng(61)
    sites, cutoff = 3, 7
    d = cutoff + 1
    dims = (d,) * sites
    a, ad, num = O.destroy(cutoff).data, O.create(cutoff).data, O.number(cutoff).data
    pairs, coefs = [], []
    for i in range(1, sites):
        pairs += [H.lazytensor(dims, dims, [i, i + 1], [ad, a]), H.lazytensor(dims, dims, [i, i + 1], [a, ad])]
        coefs += [-1.0, -1.0]
    for i in range(1, sites + 1):
        pairs.append(H.lazytensor(dims, dims, [i], [sp.csc_matrix(num @ num - num)]))
        coefs.append(0.5)
    s = H.lazysum(dims, dims, coefs, pairs)
    D = d ** sites
    assert "dtile" in Q.describe(s.q, "left", D) and "dtile" in Q.describe(s.q, "right", D)
    rho = H.rnd(rng, D, D)
    st = H.denseop(dims, dims, rho)
    r = H.denseop(dims, dims, np.full((D, D), np.nan + 0j))
    ro = H.denseop(dims, dims, np.zeros((D, D), dtype=complex))
    O.mul(ro.o, s.o, st.o, -1j, 0.0)
    O.mul(ro.o, st.o, s.o, 1j, 1.0)
    Q.mul_(r.q, s.q, st.q, -1j, 0.0)
    Q.mul_(r.q, st.q, s.q, 1j, 1.0)
    assert H.rel_err(r.q.to_host(), ro.o.data) <= TOL


@pytest.mark.parametrize("sites", [7, 8])
def test_large_state_default_threshold(monkeypatch, sites):
    """no override: a 2^21 / 2^24-amplitude Bose-Hubbard state takes the tile passes by default (8 sites: 4096 tiles per
    pass, the persistent CTAs); Hermiticity + agreement with the gather kernel (QOB_DISABLE_DTILE=1) on the full vector"""
    import qob200 as Q

    monkeypatch.delenv("QOB_DTILE_MIN_ELEMS", raising=False)
    monkeypatch.delenv("QOB_DTILE_PERSIST", raising=False)
    cutoff = 7
    f = Q.FockBasis(cutoff)
    B = Q.tensor(*[f] * sites)
    a, ad, n = Q.destroy(f), Q.create(f), Q.number(f)

    def build():
        terms, cf = [], []
        for i in range(1, sites):
            terms += [Q.LazyTensor(B, [i, i + 1], (ad, a)), Q.LazyTensor(B, [i, i + 1], (a, ad))]
            cf += [-1.0, -1.0]
        for i in range(1, sites + 1):
            terms.append(Q.LazyTensor(B, [i], (n,)))
            cf.append(0.25 * i)
        return Q.LazySum(cf, terms)

    Ht = build()
    assert "dtile" in Q.describe(Ht)
    D = 8 ** sites
    x, y = Q.Ket(B), Q.Ket(B)
    Q.fill_state(x.data, 9, D ** -0.5)
    y.data.fill_(float("nan"))
    Q.mul_(y, Ht, x, 0.7 - 0.4j, 0.0)
    d = Q.dot(x.data, y.data) / (0.7 - 0.4j)
    assert abs(d.imag) <= 1e-12 * max(1.0, abs(d.real))
    monkeypatch.setenv("QOB_DISABLE_DTILE", "1")
    Hg = build()
    assert "gather[" in Q.describe(Hg)
    y2 = Q.Ket(B)
    Q.mul_(y2, Hg, x, 0.7 - 0.4j, 0.0)
    assert np.sqrt(Q.norm2(y.data - y2.data) / Q.norm2(y2.data)) <= TOL

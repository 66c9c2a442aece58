"""GPU parity tests: libqob200 (through the C ABI, via the qob200 mirror) against the CPU oracle on the
same seeded inputs.  Tolerance: relative 2-norm error <= 1e-12 in ComplexF64 (BASELINE.json north_star);
the reference's own tests use 1e-11 … 1e-13 absolute on O(1) data (test/test_operators_lazytensor.jl:237-416).

The structure follows the reference's tests: test/test_operators_lazytensor.jl (mul! for Ket/Bra/Op-left/
Op-right, dense vs sparse factors, (alpha,beta) in {(1,0),(1.5,2.1)}, isometries, NaN kill),
test/test_operators_lazysum.jl:255-347, test/test_operators_lazyproduct.jl:171-244,
test/test_operators_sparse.jl:318-437.
"""
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


def test_native_library_is_loaded(Q):
    """The product path is the CUDA library: it must be loaded from the tree and launch kernels."""
    import torch

    assert torch.cuda.is_available()
    with open("/proc/self/maps") as f:
        assert "libqob200.so" in f.read()
    before = Q.launch_count()
    b = Q.SpinBasis(0.5)
    B = Q.tensor(b, b, b)
    op = Q.LazyTensor(B, [2], (Q.sigmax(b),))
    x = Q.randstate(B, seed=1)
    y = Q.Ket(B)
    Q.mul_(y, op, x)
    assert Q.launch_count() > before
    xh = x.to_host().reshape(2, 2, 2, order="F")
    assert np.allclose(y.to_host().reshape(2, 2, 2, order="F"), xh[:, ::-1, :], atol=1e-15)


# ------------------------------------------------------------------ LazyTensor (test_operators_lazytensor.jl:237-416)
@pytest.mark.parametrize("sparse", [False, True])
@pytest.mark.parametrize("indices", [[1, 2, 3], [1, 2], [2, 3], [1, 3], [1], [2], [3]])
def test_lazytensor_rectangular(Q, sparse, indices):
    rng = np.random.default_rng(10 + len(indices))
    # b_l = b1a⊗b2a⊗b3a, b_r = b1b⊗b2b⊗b3b with equal sizes where no factor acts
    dl_full, dr_full = (2, 4, 3), (3, 2, 5)
    dims_l = tuple(dl_full[k] if (k + 1) in indices else 3 for k in range(3))
    dims_r = tuple(dr_full[k] if (k + 1) in indices else 3 for k in range(3))
    datas = []
    for i in indices:
        a = H.rnd(rng, dims_l[i - 1], dims_r[i - 1])
        datas.append(sp.csc_matrix(a) if sparse else a)
    op = H.lazytensor(dims_l, dims_r, indices, datas, 0.1)
    H.check_mul(op, dims_l, dims_r, rng, tol=TOL)


def test_lazytensor_mixed_and_adjoint_factors(Q):
    rng = np.random.default_rng(3)
    dims_l, dims_r = (2, 3, 4), (3, 2, 4)
    a1 = H.rnd(rng, 2, 3)
    a2 = ("adj", H.rnd(rng, 2, 3))            # explicit adjoint dense (test_operators_lazytensor.jl:239-240)
    a3 = ("adj", H.sprnd(rng, 4, 4, 0.6))     # adjoint sparse
    op = H.lazytensor(dims_l, dims_r, [1, 2, 3], [a1, a2, a3], 0.1)
    H.check_mul(op, dims_l, dims_r, rng, tol=TOL)
    op = H.lazytensor(dims_l, dims_r, [1, 2], [sp.csc_matrix(a1), a2], -0.7j)
    H.check_mul(op, dims_l, dims_r, rng, tol=TOL)


def test_lazytensor_no_factors_scaled_identity_and_isometry(Q):
    """:373-407 — LazyTensor with no operators: scaled identity / scaled isometry."""
    rng = np.random.default_rng(4)
    op = H.lazytensor((2, 3), (2, 3), [], [], 0.5 - 0.25j)
    H.check_mul(op, (2, 3), (2, 3), rng, tol=TOL)
    op = H.lazytensor((2, 3, 2), (2, 2, 3), [], [], 1.5)
    H.check_mul(op, (2, 3, 2), (2, 2, 3), rng, tol=TOL)


def test_lazytensor_explicit_identities_and_isometries(Q):
    """:446-494 — explicit identityoperator factors (Eye), non-square isometry, NaN kill, alpha=0."""
    rng = np.random.default_rng(5)
    dims_l, dims_r = (3, 2, 2, 1, 2), (3, 2, 2, 2, 1)
    num = sp.csc_matrix(np.diag([0.0, 1.0, 2.0]).astype(complex))
    sx = sp.csc_matrix(np.array([[0, 1], [1, 0]], dtype=complex))
    iso = ("eye", 2, 1)
    n1 = H.lazytensor(dims_l, dims_r, [1, 3], [num, sx])
    n1_sp = H.lazytensor(dims_l, dims_r, [1, 2, 3, 5], [num, ("eye", 2, 2), sx, iso])
    n1_de = H.lazytensor(dims_l, dims_r, [1, 2, 3, 5], [num.toarray(), ("eye", 2, 2), sx, iso])
    Dl, Dr = int(np.prod(dims_l)), int(np.prod(dims_r))
    for op in (n1, n1_sp, n1_de):
        H.check_mul(op, dims_l, dims_r, rng, tol=TOL)
        # beta = 0 must kill NaNs, alpha = 0 must only scale (:476-488)
        x = H.rnd(rng, Dr, 4)
        s = H.denseop(dims_r, (4,), x)
        for (al, be) in [(0.7, 0), (0, 0.3), (0, 0)]:
            y0 = H.rnd(rng, Dl, 4)
            ref = H.denseop(dims_l, (4,), y0)
            O.mul(ref.o, op.o, s.o, al, be)
            bad = y0 * (np.nan if be == 0 else 1.0)
            r = H.denseop(dims_l, (4,), bad)
            Q.mul_(r.q, op.q, s.q, al, be)
            out = r.q.to_host()
            assert np.all(np.isfinite(out))
            assert H.rel_err(out, ref.o.data) <= TOL


def test_lazytensor_many_factors_and_higher_dims(Q):
    """terms with more factors than the fused kernel handles go through the sequential path"""
    rng = np.random.default_rng(6)
    dims = (2, 3, 2, 2, 3, 2)
    datas = [H.rnd(rng, d, d) for d in dims]
    op = H.lazytensor(dims, dims, [1, 2, 3, 4, 5, 6], datas, 0.3)
    H.check_mul(op, dims, dims, rng, tol=TOL, nbatch=3)
    datas = [H.sprnd(rng, d, d, 0.7) for d in dims[:5]]
    op = H.lazytensor(dims, dims, [1, 2, 3, 4, 5], datas, 0.3)
    H.check_mul(op, dims, dims, rng, tol=TOL, nbatch=3)


def test_lazytensor_errors(Q):
    b2, b3 = Q.GenericBasis(2), Q.GenericBasis(3)
    B = Q.CompositeBasis([b2, b3])
    a = Q.Operator(b2, b2, np.eye(2))
    with pytest.raises(AssertionError):  # unsorted indices (operators_lazytensor.jl:26)
        Q.LazyTensor(B, [2, 1], (Q.Operator(b3, b3, np.eye(3)), a))
    with pytest.raises(AssertionError):  # wrong site basis (:30-31)
        Q.LazyTensor(B, [2], (a,))
    lt = Q.LazyTensor(B, [1], (a,))
    x = Q.Ket(Q.CompositeBasis([b2, b2]), np.ones(4))
    with pytest.raises(Q.DimensionMismatch):
        Q.mul_(Q.Ket(B), lt, x)
    k = Q.Ket(B, np.ones(6))
    with pytest.raises(Q.ArgumentError):  # aliasing (:704-708)
        Q.mul_(k, lt, k)


# ------------------------------------------------------------------ dense d>=16 factors: DMMA axis kernel
@pytest.mark.parametrize("dims,indices", [((20, 3, 17), [1, 3]), ((3, 24, 2), [2]), ((16, 16), [1, 2]), ((5, 33), [2]),
                                           ((48, 48, 3), [1, 2])])
def test_lazytensor_dense_axis_kernel(Q, dims, indices):
    rng = np.random.default_rng(7)
    datas = [H.rnd(rng, dims[i - 1], dims[i - 1]) / np.sqrt(dims[i - 1]) for i in indices]
    op = H.lazytensor(dims, dims, indices, datas, 0.9)
    assert "dmma" in Q.describe(op.q)
    H.check_mul(op, dims, dims, rng, tol=TOL, nbatch=6)


def test_lazytensor_dense_axis_rectangular(Q):
    rng = np.random.default_rng(8)
    dims_l, dims_r = (18, 3, 40), (25, 3, 17)
    datas = [H.rnd(rng, 18, 25) / 5, H.rnd(rng, 40, 17) / 5]
    op = H.lazytensor(dims_l, dims_r, [1, 3], datas, 1.1)
    H.check_mul(op, dims_l, dims_r, rng, tol=TOL, nbatch=4)


# ------------------------------------------------------------------ LazySum (test_operators_lazysum.jl:255-347)
def test_lazysum_of_lazytensors_and_mixed_terms(Q):
    rng = np.random.default_rng(11)
    dims = (2, 3, 4)
    D = int(np.prod(dims))
    t1 = H.lazytensor(dims, dims, [1, 3], [H.rnd(rng, 2, 2), H.sprnd(rng, 4, 4, 0.5)], 0.3)
    t2 = H.lazytensor(dims, dims, [2], [H.rnd(rng, 3, 3)], 1.0)
    t3 = H.operator(dims, dims, H.sprnd(rng, D, D, 0.1))           # SparseOperator term
    t4 = H.operator(dims, dims, H.rnd(rng, D, D) / D)               # dense Operator term
    t5 = H.lazyproduct([t2, t1], 0.5)                               # LazyProduct term
    inner = H.lazysum(dims, dims, [0.5, -0.25j], [t1, t3])          # nested LazySum
    s = H.lazysum(dims, dims, [0.1, 0.3 + 0.1j, -0.7, 0.2, 1.0, 2.0], [t1, t2, t3, t4, t5, inner])
    H.check_mul(s, dims, dims, rng, tol=TOL)
    s2 = H.lazysum(dims, dims, [0.1, 0.3], [t1, t2])
    H.check_mul(s2, dims, dims, rng, tol=TOL)
    # coefficient update (TimeDependentSum set_time!, time_dependent_operator.jl:279-290)
    s2.q.factors[0] = 2.5 - 1j
    s2.o.factors[0] = 2.5 - 1j
    H.check_mul(s2, dims, dims, rng, tol=TOL)


def test_lazysum_empty_and_nan_kill(Q):
    """:311,335 — empty sum: result = beta*result, beta=0 kills NaN; dimension errors :118-121"""
    rng = np.random.default_rng(12)
    dims = (2, 3)
    s = H.lazysum(dims, dims, [], [])
    x = H.ket(dims, H.rnd(rng, 6))
    y0 = H.rnd(rng, 6)
    r = H.ket(dims, y0 * np.nan)
    Q.mul_(r.q, s.q, x.q, 1.0, 0.0)
    assert np.all(r.q.to_host() == 0)
    r = H.ket(dims, y0)
    Q.mul_(r.q, s.q, x.q, 1.0, 2.0)
    assert H.rel_err(r.q.to_host(), 2.0 * y0) <= TOL
    with pytest.raises(Q.DimensionMismatch):
        Q.mul_(Q.Ket(Q.CompositeBasis([Q.GenericBasis(2), Q.GenericBasis(2)])), s.q, x.q)


# ------------------------------------------------------------------ LazyProduct (test_operators_lazyproduct.jl:171-244)
@pytest.mark.parametrize("nops", [1, 2, 3])
def test_lazyproduct_chain(Q, nops):
    rng = np.random.default_rng(20 + nops)
    chain_dims = [(2, 3), (3, 2), (2, 2), (4, 1)][: nops + 1]
    ops = []
    for k in range(nops):
        dl, dr = chain_dims[k], chain_dims[k + 1]
        if k % 2 == 0:
            ops.append(H.operator(dl, dr, H.sprnd(rng, int(np.prod(dl)), int(np.prod(dr)), 0.6)))
        else:
            ops.append(H.lazytensor(dl, dr, [1, 2], [H.rnd(rng, dl[0], dr[0]), H.rnd(rng, dl[1], dr[1])], 0.7))
    p = H.lazyproduct(ops, 0.5 + 0.5j)
    H.check_mul(p, chain_dims[0], chain_dims[nops], rng, tol=TOL,
                scalars=((1, 0), (1.5, 2.1), (0, 1.3)))


# ------------------------------------------------------------------ SparseOperator (test_operators_sparse.jl:318-437)
@pytest.mark.parametrize("shape", [(3, 5, 7), (50, 60, 55)])  # small gemm and the nnz > 550 branch
def test_sparse_gemm_gemv(Q, shape):
    rng = np.random.default_rng(30)
    m, k, n = shape
    M = H.sprnd(rng, m, k, 0.5)
    op = H.operator((m,), (k,), M)
    H.check_mul(op, (m,), (k,), rng, tol=TOL, nbatch=n)
    # lazy adjoint sparse (gemm! only; gemv! has no adjoint method, operators_sparse.jl:201-202)
    opa = H.operator((k,), (m,), ("adj", M))
    H.check_mul(opa, (k,), (m,), rng, tol=TOL, nbatch=n, kinds=("opl", "opr"))
    with pytest.raises(Q.MethodError):
        Q.mul_(Q.Ket(opa.q.basis_l), opa.q, Q.Ket(opa.q.basis_r))
    with pytest.raises(Q.DimensionMismatch):
        Q.mul_(Q.Ket(Q.GenericBasis(m + 1)), op.q, Q.Ket(op.q.basis_r))


@pytest.mark.parametrize("m,k,n", [(1201, 1201, 1103), (1300, 1100, 517), (700, 650, 901)])
def test_sparse_gemm_bandwidth_regime(Q, m, k, n):
    """Operands large enough for the several-outputs-per-thread SpMM kernels (4 and 2 columns / rows per thread, ragged
    edges); the square case is >= 1024 wide, so the right-side product walks its columns in Cuthill-McKee order."""
    rng = np.random.default_rng(33)
    M = sp.random(m, k, density=4.0 / k, random_state=np.random.RandomState(5), format="csc", dtype=float).astype(complex)
    M.data = H.rnd(rng, M.nnz)
    op = H.operator((m,), (k,), M)
    if m == k:
        assert "Cuthill-McKee" in Q.describe(op.q)
    H.check_mul(op, (m,), (k,), rng, tol=TOL, nbatch=n, kinds=("opl", "opr"), scalars=((1.5, 2.1), (-1j, 0)))
    opa = H.operator((k,), (m,), ("adj", M))
    H.check_mul(opa, (k,), (m,), rng, tol=TOL, nbatch=n, kinds=("opl", "opr"), scalars=((0.3 - 0.2j, 1),))


def test_jaynes_cummings_large_cutoff_commutator(Q):
    """BASELINE config 2 in its bandwidth regime (cutoff 1024, dim 2050): the partner column of the coupling term is
    ~D/2 columns away; same two mul! calls as the small case, against the oracle's CSC loops."""
    nc = 1024
    nf = nc + 1
    a, ad, num = O.destroy(nc).data, O.create(nc).data, O.number(nc).data
    sz, spl, smi = O.sigmaz().data, O.sigmap().data, O.sigmam().data
    i2, inf = sp.identity(2, format="csc"), sp.identity(nf, format="csc")
    Hm = (1.0 * sp.kron(i2, num) + 0.45 * sp.kron(sz, inf) + 0.1 * (sp.kron(spl, a) + sp.kron(smi, ad))).tocsc()
    D = 2 * nf
    rng = np.random.default_rng(34)
    op = H.operator((D,), (D,), Hm)
    assert "Cuthill-McKee" in Q.describe(op.q)
    rho = H.rnd(rng, D, D)
    s = H.denseop((D,), (D,), rho)
    r = H.denseop((D,), (D,), np.zeros((D, D), dtype=complex))
    O.mul(r.o, op.o, s.o, -1j, 0.0)
    O.mul(r.o, s.o, op.o, 1j, 1.0)
    Q.mul_(r.q, op.q, s.q, -1j, 0.0)
    Q.mul_(r.q, s.q, op.q, 1j, 1.0)
    assert H.rel_err(r.q.to_host(), r.o.data) <= TOL


def test_dense_operator_as_operator(Q):
    rng = np.random.default_rng(31)
    for (m, k) in [(7, 5), (40, 33)]:
        op = H.operator((m,), (k,), H.rnd(rng, m, k) / k)
        H.check_mul(op, (m,), (k,), rng, tol=TOL, nbatch=9)
        opa = H.operator((k,), (m,), ("adj", H.rnd(rng, m, k) / k))
        H.check_mul(opa, (k,), (m,), rng, tol=TOL, nbatch=9)


# ------------------------------------------------------------------ BASELINE configs
def _pauli():
    sx = np.array([[0, 1], [1, 0]], dtype=complex)
    sy = np.array([[0, -1j], [1j, 0]], dtype=complex)
    sz = np.array([[1, 0], [0, -1]], dtype=complex)
    return sx, sy, sz


def _chain_terms(n, kind, dense, rng):
    """TFIM (2N terms) or Heisenberg (3N terms) on a periodic chain, sparse factors as spin.jl builds them."""
    sx, sy, sz = _pauli()
    wrap = (lambda a: a) if dense else sp.csc_matrix
    dims = (2,) * n
    terms, coefs = [], []
    for i in range(1, n + 1):
        j = i % n + 1
        idx = sorted([i, j])
        if kind == "tfim":
            terms.append(H.lazytensor(dims, dims, [i], [wrap(sx)]))
            coefs.append(-rng.uniform(0.5, 1.5))
            terms.append(H.lazytensor(dims, dims, idx, [wrap(sz), wrap(sz)]))
            coefs.append(-rng.uniform(0.5, 1.5))
        else:
            for s in (sx, sy, sz):
                terms.append(H.lazytensor(dims, dims, idx, [wrap(s), wrap(s)]))
                coefs.append(rng.uniform(0.5, 1.5))
    return dims, coefs, terms


@pytest.mark.parametrize("dense", [False, True])
def test_config1_tfim_n12(Q, dense):
    """BASELINE config 1: TFIM N=12, LazySum of 2N embedded terms on a random Ket (generic fused kernel)."""
    rng = np.random.default_rng(40)
    dims, coefs, terms = _chain_terms(12, "tfim", dense, rng)
    s = H.lazysum(dims, dims, coefs, terms)
    assert "gather" in Q.describe(s.q)
    H.check_mul(s, dims, dims, rng, tol=TOL, kinds=("ket", "bra", "opl"), nbatch=3)


@pytest.mark.parametrize("n,T,L", [(12, 10, 3), (14, 12, 3), (15, 11, 2), (16, 12, 4), (17, 13, 3), (18, 12, 3)])
@pytest.mark.parametrize("kind", ["tfim", "heis"])
def test_qtile_chain_vs_oracle(Q, monkeypatch, n, T, L, kind):
    """The tile kernel (forced on for small chains) against the reference's per-term sparse recursion."""
    monkeypatch.setenv("QOB_QTILE_MIN_BITS", "10")
    monkeypatch.setenv("QOB_DISABLE_QREG", "1")   # this test pins the round-1 tile kernel (tests/test_gpu_qreg.py covers the round-2 one)
    monkeypatch.setenv("QOB_QTILE_T", str(T))
    monkeypatch.setenv("QOB_QTILE_L", str(L))
    rng = np.random.default_rng(50 + n)
    dims, coefs, terms = _chain_terms(n, kind, False, rng)
    s = H.lazysum(dims, dims, coefs, terms)
    assert "qtile" in Q.describe(s.q)
    H.check_mul(s, dims, dims, rng, tol=TOL, kinds=("ket", "bra"))
    if n <= 14:
        H.check_mul(s, dims, dims, rng, tol=TOL, kinds=("opl", "opr"), nbatch=4, scalars=((1, 0), (1.5, 2.1)))


def test_qtile_general_2x2_factors_and_three_site_terms(Q, monkeypatch):
    monkeypatch.setenv("QOB_QTILE_MIN_BITS", "10")
    monkeypatch.setenv("QOB_DISABLE_QREG", "1")   # this test pins the round-1 tile kernel (tests/test_gpu_qreg.py covers the round-2 one)
    rng = np.random.default_rng(60)
    n = 14
    dims = (2,) * n
    terms, coefs = [], []
    for idx in ([1], [7], [14], [1, 14], [3, 9], [2, 3, 4], [1, 8, 14], [12, 13]):
        datas = [H.rnd(rng, 2, 2) if rng.uniform() < 0.5 else H.sprnd(rng, 2, 2, 0.7) for _ in idx]
        terms.append(H.lazytensor(dims, dims, idx, datas, rng.uniform(0.5, 1.0)))
        coefs.append(H.rnd(rng, 1)[0])
    s = H.lazysum(dims, dims, coefs, terms)
    assert "qtile" in Q.describe(s.q)
    H.check_mul(s, dims, dims, rng, tol=TOL, kinds=("ket", "bra"))


def test_config2_jaynes_cummings_liouvillian(Q):
    """BASELINE config 2: H (sparse, 258 nnz) x rho (dense 130x130): mul!(drho,H,rho,-i,0); mul!(drho,rho,H,i,1)."""
    rng = np.random.default_rng(70)
    nf = 65
    a = O.destroy(64).data
    ad = O.create(64).data
    num = O.number(64).data
    sz, sp_, sm = O.sigmaz().data, O.sigmap().data, O.sigmam().data
    i2, inf = sp.identity(2, format="csc"), sp.identity(nf, format="csc")
    # tensor(a, b) = kron(b, a): subsystem 1 (cavity) is the fast index
    Hm = (1.0 * sp.kron(i2, num) + 0.5 * 0.9 * sp.kron(sz, inf) + 0.1 * (sp.kron(sp_, a) + sp.kron(sm, ad))).tocsc()
    assert Hm.nnz == 258
    op = H.operator((nf, 2), (nf, 2), Hm)
    rho = H.rnd(rng, 130, 130)
    s = H.denseop((nf, 2), (nf, 2), rho)
    r = H.denseop((nf, 2), (nf, 2), np.zeros((130, 130), dtype=complex))
    O.mul(r.o, op.o, s.o, -1j, 0)
    O.mul(r.o, s.o, op.o, 1j, 1)
    Q.mul_(r.q, op.q, s.q, -1j, 0)
    Q.mul_(r.q, s.q, op.q, 1j, 1)
    assert H.rel_err(r.q.to_host(), r.o.data) <= TOL
    # the same H as a LazySum of LazyTensors
    dims = (nf, 2)
    terms = [H.lazytensor(dims, dims, [1], [num]), H.lazytensor(dims, dims, [2], [sz]),
             H.lazytensor(dims, dims, [1, 2], [a, sp_]), H.lazytensor(dims, dims, [1, 2], [ad, sm])]
    ls = H.lazysum(dims, dims, [1.0, 0.45, 0.1, 0.1], terms)
    r2 = H.denseop(dims, dims, np.zeros((130, 130), dtype=complex))
    Q.mul_(r2.q, ls.q, s.q, -1j, 0)
    Q.mul_(r2.q, s.q, ls.q, 1j, 1)
    assert H.rel_err(r2.q.to_host(), r.o.data) <= TOL


def test_config3_two_mode_fock_dense_factors(Q):
    """BASELINE config 3 (reduced batch): dims (48,48,3), LazyTensor with two dense d=48 factors on a Ket batch."""
    rng = np.random.default_rng(80)
    dims = (48, 48, 3)
    A1, A2 = H.rnd(rng, 48, 48) / 7, H.rnd(rng, 48, 48) / 7
    op = H.lazytensor(dims, dims, [1, 2], [A1, A2])
    assert Q.describe(op.q).count("dmma") == 2
    H.check_mul(op, dims, dims, rng, tol=TOL, kinds=("ket", "opl"), nbatch=8, scalars=((1, 0), (1.5, 2.1)))
    H.check_mul(op, dims, dims, rng, tol=TOL, kinds=("bra", "opr"), nbatch=3, scalars=((1, 0),))


def test_apply_host_end_to_end(Q):
    rng = np.random.default_rng(90)
    dims, coefs, terms = _chain_terms(10, "heis", False, rng)
    s = H.lazysum(dims, dims, coefs, terms)
    x = H.rnd(rng, 1 << 10)
    ref = H.ket(dims, np.zeros(1 << 10, dtype=complex))
    O.mul(ref.o, s.o, H.ket(dims, x).o, 1.0, 0.0)
    y = Q.apply_host(s.q, x)
    assert H.rel_err(y, ref.o.data) <= TOL


@pytest.mark.parametrize("pipe_min,batch", [(16384, 7), (40000, 7), (16384, 2), (0, 5)])
@pytest.mark.parametrize("beta", [0.0, 0.4 - 1.1j])
def test_apply_host_batch_of_kets_is_pipelined(Q, monkeypatch, pipe_min, batch, beta):
    """qob_op_apply_host on a batch of host kets streams column groups through two device lanes (H2D / apply / D2H on
    separate streams): groups of 1 and of 2 columns with a ragged last group, the two-group minimum, and the plain
    path (pipe_min = 0 disables the pipeline).  beta != 0 also stages y."""
    monkeypatch.setenv("QOB_HOST_PIPE_MIN_BYTES", str(pipe_min))
    rng = np.random.default_rng(91)
    dims, coefs, terms = _chain_terms(10, "heis", False, rng)
    s = H.lazysum(dims, dims, coefs, terms)
    D = 1 << 10
    x, y0 = H.rnd(rng, D, batch), H.rnd(rng, D, batch)
    ref = H.denseop(dims, (batch,), y0)
    O.mul(ref.o, s.o, H.denseop(dims, (batch,), x).o, 0.3 + 0.2j, beta)
    y = np.ascontiguousarray(y0.reshape(-1, order="F"))
    l0 = Q.launch_count()
    Q.apply_host(s.q, x, alpha=0.3 + 0.2j, beta=beta, y=y, batch=batch)
    assert H.rel_err(y, np.asarray(ref.o.data).reshape(-1, order="F")) <= TOL
    col = 16 * D
    groups = 1 if pipe_min == 0 or col * batch < 2 * pipe_min else -(-batch // max(1, min(batch // 2, pipe_min // col)))
    assert Q.launch_count() - l0 >= groups   # one apply per column group


def test_apply_host_pipelined_isometry(Q, monkeypatch):
    """different input and output dimensions per ket (Eye isometry factor) through the pipelined host path"""
    monkeypatch.setenv("QOB_HOST_PIPE_MIN_BYTES", "4096")
    rng = np.random.default_rng(92)
    dl, dr = (3, 4, 5), (3, 6, 5)
    op = H.lazytensor(dl, dr, [1, 3], [H.sprnd(rng, 3, 3), H.rnd(rng, 5, 5)], factor=0.7)
    batch = 9
    x = H.rnd(rng, 90, batch)
    ref = H.denseop(dl, (batch,), np.zeros((60, batch), dtype=complex))
    O.mul(ref.o, op.o, H.denseop(dr, (batch,), x).o, 1.0, 0.0)
    y = Q.apply_host(op.q, x, batch=batch)
    assert H.rel_err(y, np.asarray(ref.o.data).reshape(-1, order="F")) <= TOL


def test_expect_and_variance(Q):
    """expect / variance built on mul! + a device reduction (src/operators.jl:119-150)."""
    rng = np.random.default_rng(95)
    dims, coefs, terms = _chain_terms(11, "heis", False, rng)
    s = H.lazysum(dims, dims, coefs, terms)
    x = H.rnd(rng, 1 << 11)
    x /= np.linalg.norm(x)
    Dm = O.dense(s.o)
    k = H.ket(dims, x).q
    e = Q.expect(s.q, k)
    ref = np.vdot(x, Dm @ x)
    assert abs(e - ref) <= 1e-12 * max(1.0, abs(ref))
    v = Q.variance(s.q, k)
    refv = np.vdot(x, Dm @ (Dm @ x)) - ref * ref
    assert abs(v - refv) <= 1e-11 * max(1.0, abs(refv))
    rho = np.outer(x, x.conj())
    er = Q.expect(s.q, H.denseop(dims, dims, rho).q)
    assert abs(er - ref) <= 1e-12 * max(1.0, abs(ref))


def test_tile_planner_declines_and_generic_kernel_takes_over(Q, monkeypatch):
    """78 terms X_1 Z_j Z_k share one flip mask with 78 different selector sets: more shared-mask lookups than one tile pass
    carries -> the planner declines and the generic fused kernel computes the same map."""
    monkeypatch.setenv("QOB_QTILE_MIN_BITS", "10")
    monkeypatch.setenv("QOB_DISABLE_QREG", "1")   # this test pins the round-1 tile kernel (tests/test_gpu_qreg.py covers the round-2 one)
    rng = np.random.default_rng(96)
    n = 14
    dims = (2,) * n
    sx, _, sz = _pauli()
    terms, coefs = [], []
    for j in range(2, n + 1):
        for k in range(j + 1, n + 1):
            terms.append(H.lazytensor(dims, dims, [1, j, k], [sp.csc_matrix(sx), sp.csc_matrix(sz), sp.csc_matrix(sz)]))
            coefs.append(rng.uniform(-1, 1))
    s = H.lazysum(dims, dims, coefs, terms)
    d = Q.describe(s.q)
    assert "qtile" not in d and "gather" in d, d
    H.check_mul(s, dims, dims, rng, tol=TOL, kinds=("ket",), scalars=((1, 0), (1.5, 2.1)))


@pytest.mark.parametrize("seed", range(8))
def test_qtile_random_term_sets_vs_oracle(Q, monkeypatch, seed):
    """Property-style sweep over the tile planner's corner cases: random chain lengths, tile sizes, 1-3-site terms on
    random sites with random dense / sparse / diagonal / adjoint 2x2 factors, repeated site sets (shared masks)."""
    rng = np.random.default_rng(1000 + seed)
    n = int(rng.integers(12, 17))
    T = int(rng.integers(10, min(n, 13) + 1))
    monkeypatch.setenv("QOB_QTILE_MIN_BITS", "10")
    monkeypatch.setenv("QOB_DISABLE_QREG", "1")   # this test pins the round-1 tile kernel (tests/test_gpu_qreg.py covers the round-2 one)
    monkeypatch.setenv("QOB_QTILE_T", str(T))
    dims = (2,) * n
    terms, coefs = [], []
    site_sets = []
    for _ in range(int(rng.integers(5, 30))):
        k = int(rng.integers(1, 4))
        if site_sets and rng.uniform() < 0.3:
            idx = site_sets[int(rng.integers(len(site_sets)))]   # same sites again -> shared masks / selector sets
        else:
            idx = sorted(int(v) for v in rng.choice(np.arange(1, n + 1), size=k, replace=False))
            site_sets.append(idx)
        datas = []
        for _s in idx:
            kind = rng.integers(4)
            m = H.rnd(rng, 2, 2)
            if kind == 0:
                datas.append(m)
            elif kind == 1:
                datas.append(sp.csc_matrix(m * (rng.uniform(0, 1, (2, 2)) < 0.6)))
            elif kind == 2:
                datas.append(sp.csc_matrix(np.diag(np.diag(m))))
            else:
                datas.append(("adj", m))
        terms.append(H.lazytensor(dims, dims, idx, datas, H.rnd(rng, 1)[0]))
        coefs.append(H.rnd(rng, 1)[0])
    s = H.lazysum(dims, dims, coefs, terms)
    H.check_mul(s, dims, dims, rng, tol=TOL, kinds=("ket", "bra"), scalars=((1, 0), (0.3 - 0.2j, 1.7)))


@pytest.mark.parametrize("n", [9, 10])
def test_qtile_density_matrix_commutator(Q, monkeypatch, n):
    """-i[H, rho] on a full 2^n x 2^n density matrix: left and right application run as tile passes over 2n index bits."""
    monkeypatch.setenv("QOB_QTILE_MIN_BITS", "10")
    monkeypatch.setenv("QOB_DISABLE_QREG", "1")   # this test pins the round-1 tile kernel (tests/test_gpu_qreg.py covers the round-2 one)
    rng = np.random.default_rng(77 + n)
    dims, coefs, terms = _chain_terms(n, "heis", False, rng)
    s = H.lazysum(dims, dims, coefs, terms)
    D = 1 << n
    assert "qtile[bits=%d" % (2 * n) in Q.describe(s.q, "left", D) and "qtile[bits=%d" % (2 * n) in Q.describe(s.q, "right", D)
    rho = H.rnd(rng, D, D)
    st, r = H.denseop(dims, dims, rho), H.denseop(dims, dims, np.zeros((D, D), dtype=complex))
    O.mul(r.o, s.o, st.o, -1j, 0)
    O.mul(r.o, st.o, s.o, 1j, 1)
    Q.mul_(r.q, s.q, st.q, -1j, 0)
    Q.mul_(r.q, st.q, s.q, 1j, 1)
    assert H.rel_err(r.q.to_host(), r.o.data) <= TOL


def test_dense_device_times_dense_device_is_library_gemm(Q):
    """operators_dense.jl:394-396: dense x dense goes to BLAS in the reference, to cuBLAS (through torch) here."""
    rng = np.random.default_rng(97)
    a, b, r0 = H.rnd(rng, 6, 5), H.rnd(rng, 5, 7), H.rnd(rng, 6, 7)
    A, B_, R = H.denseop((6,), (5,), a), H.denseop((5,), (7,), b), H.denseop((6,), (7,), r0)
    Q.mul_(R.q, A.q, B_.q, 1.5, 2.1)
    assert H.rel_err(R.q.to_host(), 1.5 * a @ b + 2.1 * r0) <= TOL
    R = H.denseop((6,), (7,), r0 * np.nan)
    Q.mul_(R.q, A.q, B_.q, 1.0, 0.0)
    assert H.rel_err(R.q.to_host(), a @ b) <= TOL
    x, y0 = H.rnd(rng, 5), H.rnd(rng, 6)
    k, out = H.ket((5,), x), H.ket((6,), y0)
    Q.mul_(out.q, A.q, k.q, -1j, 0.5)
    assert H.rel_err(out.q.to_host(), -1j * a @ x + 0.5 * y0) <= TOL
    xb, yb0 = H.rnd(rng, 6), H.rnd(rng, 5)
    bb, outb = H.bra((6,), xb), H.bra((5,), yb0)
    Q.mul_(outb.q, bb.q, A.q, 2.0, 1.0)
    assert H.rel_err(outb.q.to_host(), 2.0 * xb @ a + yb0) <= TOL


@pytest.mark.parametrize("n", [12, 14])
def test_qtile_more_masks_than_records_per_pass(Q, monkeypatch, n):
    """All-to-all sigma_x sigma_x couplings: 91 distinct flip masks, more than the 56 lookup records a pass carries in its
    kernel parameters -> the planner chunks them into extra passes over the same tiles."""
    monkeypatch.setenv("QOB_QTILE_MIN_BITS", "10")
    monkeypatch.setenv("QOB_DISABLE_QREG", "1")   # this test pins the round-1 tile kernel (tests/test_gpu_qreg.py covers the round-2 one)
    rng = np.random.default_rng(98)
    dims = (2,) * n
    sx, sy, sz = _pauli()
    terms, coefs = [], []
    for i in range(1, n + 1):
        for j in range(i + 1, n + 1):
            terms.append(H.lazytensor(dims, dims, [i, j], [sp.csc_matrix(sx), sp.csc_matrix(sx)]))
            coefs.append(rng.uniform(-1, 1) / abs(i - j) ** 3)
    terms.append(H.lazytensor(dims, dims, [1, 7, n], [sy, sz, sy]))   # three separate selector runs
    coefs.append(0.37)
    s = H.lazysum(dims, dims, coefs, terms)
    d = Q.describe(s.q)
    assert "qtile" in d and d.count("{free:") >= (2 if n == 12 else 3), d
    H.check_mul(s, dims, dims, rng, tol=TOL, kinds=("ket", "bra"), scalars=((1, 0), (1.5, 2.1)))


def test_time_dependent_sum_coefficient_updates(Q, monkeypatch):
    """TimeDependentSum: set_time! rewrites the LazySum factors (time_dependent_operator.jl:279-290); the device plan keeps
    its passes and only refills its weight tables.  Checked on the tile-kernel path at several times."""
    monkeypatch.setenv("QOB_QTILE_MIN_BITS", "10")
    monkeypatch.setenv("QOB_DISABLE_QREG", "1")   # this test pins the round-1 tile kernel (tests/test_gpu_qreg.py covers the round-2 one)
    rng = np.random.default_rng(99)
    n = 13
    dims, coefs, terms = _chain_terms(n, "heis", False, rng)
    fns = [(lambda t, c=c, k=k: c * np.cos(0.3 * k * t) + 0.1j * np.sin(t)) if k % 3 else c for k, c in enumerate(coefs)]
    td = Q.TimeDependentSum(fns, [t_.q for t_ in terms])
    assert "qtile" in Q.describe(td.static_op)
    x = H.rnd(rng, 1 << n)
    launches = None
    for t in (0.0, 0.7, 1.9):
        td.set_time_(t)
        now = [complex(f(t)) if callable(f) else complex(f) for f in fns]
        so = O.LazySum(dims, dims, now, [t_.o for t_ in terms])
        ref = H.ket(dims, np.zeros(1 << n, dtype=complex))
        O.mul(ref.o, so, H.ket(dims, x).o, 1.0, 0.0)
        out = H.ket(dims, np.zeros(1 << n, dtype=complex))
        before = Q.launch_count()
        Q.mul_(out.q, td, H.ket(dims, x).q, 1.0, 0.0)
        used = Q.launch_count() - before
        launches = used if launches is None else launches
        assert used == launches            # same plan every time
        assert H.rel_err(out.q.to_host(), ref.o.data) <= TOL


def test_lindblad_rhs_call_pattern_with_bool_scalars(Q):
    """The master-equation right-hand side as ODE solvers call it (test/test_sciml_broadcast_interfaces.jl:36-43):
    3-argument mul!, Bool alpha/beta, lazy-adjoint jump operators, left and right application on a dense rho."""
    rng = np.random.default_rng(101)
    n = 6
    dims, coefs, terms = _chain_terms(n, "heis", False, rng)
    Hs = H.lazysum(dims, dims, coefs, terms)
    sm = np.array([[0, 0], [1, 0]], dtype=complex)                       # sigma_minus as spin.jl builds it (lower diagonal)
    J = H.lazytensor(dims, dims, [3], [sp.csc_matrix(sm)], 0.8)
    Jd = H.lazytensor(dims, dims, [3], [("adj", sp.csc_matrix(sm))], 0.8)  # lazy dagger of the site operator
    mh = H.lazytensor(dims, dims, [3], [-0.5 * 0.64 * (sm.conj().T @ sm)])
    D = 1 << n
    rho0 = H.rnd(rng, D, D)
    rho = H.denseop(dims, dims, rho0)
    tmp = H.denseop(dims, dims, np.full((D, D), np.nan + 0j))
    drho = H.denseop(dims, dims, np.full((D, D), np.nan + 0j))
    Q.mul_(tmp.q, rho.q, Jd.q)                  # tmp  = rho J^dagger          (alpha = true, beta = false)
    Q.mul_(drho.q, J.q, tmp.q)                  # drho = J rho J^dagger
    Q.mul_(drho.q, rho.q, mh.q, True, True)     # drho += rho (-J^dagger J / 2)
    Q.mul_(drho.q, mh.q, rho.q, True, True)
    Q.mul_(drho.q, Hs.q, rho.q, -1j, 1)         # drho += -i H rho
    Q.mul_(drho.q, rho.q, Hs.q, 1j, True)       # drho += +i rho H
    Hd, Jm = O.dense(Hs.o), O.dense(J.o)
    ref = Jm @ rho0 @ Jm.conj().T - 0.5 * (Jm.conj().T @ Jm @ rho0 + rho0 @ Jm.conj().T @ Jm) - 1j * (Hd @ rho0 - rho0 @ Hd)
    assert H.rel_err(drho.q.to_host(), ref) <= TOL

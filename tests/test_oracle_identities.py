"""Pins the CPU oracle (oracle/qob_oracle.{c,py}) the way the reference pins its own hot path: by identities.

The reference's tests store no golden vectors (Julia RNG); every numeric test builds a lazy / sparse operator
and its explicit dense kron twin and compares (test/test_operators_lazytensor.jl:16-19,237-416,446-494,
test/test_operators_lazysum.jl:255-347, test/test_operators_lazyproduct.jl:171-244,
test/test_operators_sparse.jl:318-437), plus known-answer algebra for the site operators
(test/test_spin.jl:16-118) and embed ordering (test/test_embed.jl:47-90).  The same identities are run here on
the restatement; tolerances as in those tests (1e-12 / 1e-13 absolute on O(1) data).
"""
import numpy as np
import pytest
import scipy.sparse as sp

import helpers as H
from helpers import O


def D(a, b):
    return np.linalg.norm(np.asarray(a).reshape(-1) - np.asarray(b).reshape(-1))


def ops_for(rng, dims_l, dims_r, indices, kind):
    out = []
    for i in indices:
        a = H.rnd(rng, dims_l[i - 1], dims_r[i - 1])
        if kind == "dense":
            out.append(O.Op((a.shape[0],), (a.shape[1],), a))
        elif kind == "sparse":
            out.append(O.Op((a.shape[0],), (a.shape[1],), sp.csc_matrix(a)))
        else:  # adjoint-wrapped dense, as `randoperator(b2b, b2a)'` in the reference test (:239-240)
            out.append(O.Op((a.shape[0],), (a.shape[1],), O.Adj(np.asfortranarray(a.conj().T))))
    return out


@pytest.mark.parametrize("kind", ["dense", "sparse", "adjoint"])
@pytest.mark.parametrize("indices", [[1, 2, 3], [1, 2], [1, 3], [2, 3], [1], [2], [3], []])
@pytest.mark.parametrize("scal", [(1.0, 0.0), (1.5, 2.1)])
def test_lazytensor_mul_matches_dense_kron(kind, indices, scal):
    rng = np.random.default_rng(1)
    full_l, full_r = (2, 4, 3), (3, 2, 5)
    dims_l = tuple(full_l[k] if (k + 1) in indices else 3 for k in range(3))
    dims_r = tuple(full_r[k] if (k + 1) in indices else 3 for k in range(3))
    lt = O.LazyTensor(dims_l, dims_r, indices, ops_for(rng, dims_l, dims_r, indices, kind), 0.1)
    Dm = O.dense(lt)
    Dl, Dr = Dm.shape
    al, be = scal
    x, y0 = H.rnd(rng, Dr), H.rnd(rng, Dl)
    y = O.Ket(dims_l, y0.copy())
    O.mul(y, lt, O.Ket(dims_r, x), al, be)
    assert D(y.data, al * Dm @ x + be * y0) < 1e-13
    xb, yb0 = H.rnd(rng, Dl), H.rnd(rng, Dr)
    yb = O.Bra(dims_r, yb0.copy())
    O.mul(yb, O.Bra(dims_l, xb), lt, al, be)
    assert D(yb.data, al * xb @ Dm + be * yb0) < 1e-13
    X, Y0 = H.rnd(rng, Dr, 4), H.rnd(rng, Dl, 4)
    Y = O.Op(dims_l, (4,), Y0.copy())
    O.mul(Y, lt, O.Op(dims_r, (4,), X), al, be)
    assert D(Y.data, al * Dm @ X + be * Y0) < 1e-12
    X, Y0 = H.rnd(rng, 4, Dl), H.rnd(rng, 4, Dr)
    Y = O.Op((4,), dims_r, Y0.copy())
    O.mul(Y, O.Op((4,), dims_l, X), lt, al, be)
    assert D(Y.data, al * X @ Dm + be * Y0) < 1e-12


def test_pure_sparse_path_equals_dense_path():
    """the two reference code paths (_mul_puresparse! vs _tp_sum_matmul!) compute the same map"""
    rng = np.random.default_rng(2)
    dims = (3, 2, 4, 2)
    mats = [H.sprnd(rng, d, d, 0.6) for d in dims]
    for idx in ([1, 3], [2, 4], [1, 2, 3, 4], [4]):
        sp_ops = [O.Op((dims[i - 1],), (dims[i - 1],), mats[i - 1]) for i in idx]
        de_ops = [O.Op((dims[i - 1],), (dims[i - 1],), mats[i - 1].toarray()) for i in idx]
        a, b = O.LazyTensor(dims, dims, idx, sp_ops, 0.7j), O.LazyTensor(dims, dims, idx, de_ops, 0.7j)
        assert O._is_pure_sparse(a.operators) and not O._is_pure_sparse(b.operators)
        x = H.rnd(rng, int(np.prod(dims)))
        ya, yb = O.Ket(dims, H.rnd(rng, x.size)), None
        yb = O.Ket(dims, ya.data.copy())
        O.mul(ya, a, O.Ket(dims, x), 1.5, 2.1)
        O.mul(yb, b, O.Ket(dims, x), 1.5, 2.1)
        assert D(ya.data, yb.data) < 1e-13


def test_explicit_identities_isometries_nan_and_alpha_zero():
    """test/test_operators_lazytensor.jl:446-494"""
    rng = np.random.default_rng(3)
    dims_l, dims_r = (3, 2, 2, 1, 2), (3, 2, 2, 2, 1)
    num = O.number(2)
    sx = O.sigmax()
    iso = O.Op((2,), (1,), O.Eye(2, 1))
    ident = O.Op((2,), (2,), O.Eye(2, 2))
    n1 = O.LazyTensor(dims_l, dims_r, [1, 3], [num, sx])
    n1_sp = O.LazyTensor(dims_l, dims_r, [1, 2, 3, 5], [num, ident, sx, iso])
    n1_de = O.LazyTensor(dims_l, dims_r, [1, 2, 3, 5], [O.Op((3,), (3,), num.data.toarray()), ident, sx, iso])
    assert D(O.dense(n1), O.dense(n1_sp)) == 0 and D(O.dense(n1), O.dense(n1_de)) == 0
    Dl, Dr = int(np.prod(dims_l)), int(np.prod(dims_r))
    state = O.Op(dims_r, dims_r, H.rnd(rng, Dr, Dr))
    out = H.rnd(rng, Dl, Dr)
    for (al, be) in [(0.7, -1.3), (0.7, 0), (0, -1.3), (0, 0)]:
        init = out * np.nan if be == 0 else out
        refs = []
        for op in (n1, n1_sp, n1_de):
            r = O.Op(dims_l, dims_r, init.copy())
            O.mul(r, op, state, al, be)
            assert np.all(np.isfinite(r.data))
            refs.append(r.data)
        assert D(refs[0], al * O.dense(n1) @ state.data + (0 if be == 0 else be * out)) < 1e-12
        assert D(refs[0], refs[1]) < 1e-12 and D(refs[0], refs[2]) < 1e-12
    state = O.Op(dims_l, dims_l, H.rnd(rng, Dl, Dl))
    for op in (n1, n1_sp, n1_de):
        r = O.Op(dims_l, dims_r, out.copy())
        O.mul(r, state, op, 0.7, -1.3)
        assert D(r.data, 0.7 * state.data @ O.dense(n1) - 1.3 * out) < 1e-12


def test_lazysum_and_empty_sum():
    """test/test_operators_lazysum.jl:255-347 incl. :311,335 (NaN kill) and :118-121"""
    rng = np.random.default_rng(4)
    dims = (2, 3, 4)
    Dn = 24
    t1 = O.LazyTensor(dims, dims, [1, 3], ops_for(rng, dims, dims, [1, 3], "sparse"), 0.3)
    t2 = O.LazyTensor(dims, dims, [2], ops_for(rng, dims, dims, [2], "dense"), 1.0)
    t3 = O.Op(dims, dims, H.sprnd(rng, Dn, Dn, 0.2))
    s = O.LazySum(dims, dims, [0.1, 0.3 + 0.1j, -0.7], [t1, t2, t3])
    Dm = O.dense(s)
    x, y0 = H.rnd(rng, Dn), H.rnd(rng, Dn)
    y = O.Ket(dims, y0.copy())
    O.mul(y, s, O.Ket(dims, x), 1.5, 2.1)
    assert D(y.data, 1.5 * Dm @ x + 2.1 * y0) < 1e-12
    yb = O.Bra(dims, y0.copy())
    O.mul(yb, O.Bra(dims, x), s, 1.5, 2.1)
    assert D(yb.data, 1.5 * x @ Dm + 2.1 * y0) < 1e-12
    empty = O.LazySum(dims, dims, [], [])
    y = O.Ket(dims, y0 * np.nan)
    O.mul(y, empty, O.Ket(dims, x), 1.0, 0.0)
    assert np.all(y.data == 0)
    y = O.Ket(dims, y0.copy())
    O.mul(y, empty, O.Ket(dims, x), 1.0, 2.0)
    assert D(y.data, 2 * y0) < 1e-15
    with pytest.raises(O.IncompatibleBases):
        O.LazySum(dims, dims, [1.0], [O.LazyTensor((2, 3), (2, 3), [1], ops_for(rng, (2, 3), (2, 3), [1], "dense"))])


@pytest.mark.parametrize("nops", [1, 2, 3])
def test_lazyproduct(nops):
    """test/test_operators_lazyproduct.jl:171-244"""
    rng = np.random.default_rng(5)
    chain = [(2, 3), (3, 2), (2, 2), (4, 1)][: nops + 1]
    ops = []
    for k in range(nops):
        dl, dr = chain[k], chain[k + 1]
        if k % 2 == 0:
            ops.append(O.Op(dl, dr, H.sprnd(rng, int(np.prod(dl)), int(np.prod(dr)), 0.6)))
        else:
            ops.append(O.LazyTensor(dl, dr, [1, 2], ops_for(rng, dl, dr, [1, 2], "dense"), 0.7))
    p = O.LazyProduct(ops, 0.5 + 0.5j)
    Dm = O.dense(p)
    dl, dr = chain[0], chain[nops]
    x, y0 = H.rnd(rng, Dm.shape[1]), H.rnd(rng, Dm.shape[0])
    for (al, be) in [(1, 0), (1.5, 2.1), (0, 1.3)]:
        y = O.Ket(dl, y0.copy())
        O.mul(y, p, O.Ket(dr, x), al, be)
        assert D(y.data, al * Dm @ x + be * y0) < 1e-12
        X, Y0 = H.rnd(rng, 3, Dm.shape[0]), H.rnd(rng, 3, Dm.shape[1])
        Y = O.Op((3,), dr, Y0.copy())
        O.mul(Y, O.Op((3,), dl, X), p, al, be)
        assert D(Y.data, al * X @ Dm + be * Y0) < 1e-12


@pytest.mark.parametrize("shape", [(3, 5, 7), (50, 60, 55)])
def test_sparse_gemm_gemv_all_variants(shape):
    """test/test_operators_sparse.jl:318-437: gemv, small gemm, big gemm (nnz > 550), lazy adjoint."""
    rng = np.random.default_rng(6)
    m, k, n = shape
    M = H.sprnd(rng, m, k, 0.5)
    if shape[0] >= 50:
        assert M.nnz > 550
    Md = M.toarray()
    for (al, be) in [(1, 0), (1.5, 2.1), (1, 1)]:
        B, R0 = H.rnd(rng, k, n), H.rnd(rng, m, n)
        R = np.asfortranarray(R0.copy())
        O.gemm(al, M, np.asfortranarray(B), be, R)
        assert D(R, al * Md @ B + be * R0) < 1e-11
        B, R0 = H.rnd(rng, n, m), H.rnd(rng, n, k)
        R = np.asfortranarray(R0.copy())
        O.gemm(al, np.asfortranarray(B), M, be, R)
        assert D(R, al * B @ Md + be * R0) < 1e-11
        B, R0 = H.rnd(rng, m, n), H.rnd(rng, k, n)
        R = np.asfortranarray(R0.copy())
        O.gemm(al, O.Adj(M), np.asfortranarray(B), be, R)
        assert D(R, al * Md.conj().T @ B + be * R0) < 1e-11
        B, R0 = H.rnd(rng, n, k), H.rnd(rng, n, m)
        R = np.asfortranarray(R0.copy())
        O.gemm(al, np.asfortranarray(B), O.Adj(M), be, R)
        assert D(R, al * B @ Md.conj().T + be * R0) < 1e-11
        v, r0 = H.rnd(rng, k), H.rnd(rng, m)
        r = r0.copy()
        O.gemv(al, M, v, be, r)
        assert D(r, al * Md @ v + be * r0) < 1e-12
        v, r0 = H.rnd(rng, m), H.rnd(rng, k)
        r = r0.copy()
        O.gemv(al, v, M, be, r)
        assert D(r, al * v @ Md + be * r0) < 1e-12
    with pytest.raises(O.DimensionMismatch):
        O.gemm(1, M, np.asfortranarray(H.rnd(rng, k + 1, n)), 0, np.asfortranarray(H.rnd(rng, m, n)))
    with pytest.raises(O.MethodError):
        O.gemm(1, M, M.T.tocsc(), 0, np.zeros((m, m), dtype=complex, order="F"))


def test_aliasing_and_dimension_errors():
    rng = np.random.default_rng(7)
    dims = (2, 3)
    lt = O.LazyTensor(dims, dims, [1], ops_for(rng, dims, dims, [1], "sparse"))
    k = O.Ket(dims, H.rnd(rng, 6))
    with pytest.raises(O.ArgumentError):  # operators_lazytensor.jl:704-708
        O.mul(k, lt, k)
    with pytest.raises(O.IncompatibleBases):
        O.mul(O.Ket((2, 2), np.zeros(4)), lt, k)


# ---------------------------------------------------------------- known answers (test/test_spin.jl:16-118)
@pytest.mark.parametrize("spin", [0.5, 1.0, 1.5, 2.5])
def test_spin_operator_algebra(spin):
    sx, sy, sz = (O.dense(f(spin)) for f in (O.sigmax, O.sigmay, O.sigmaz))
    sp_, sm = O.dense(O.sigmap(spin)), O.dense(O.sigmam(spin))
    comm = lambda a, b: a @ b - b @ a
    assert D(comm(sx, sy), 2j * sz) < 1e-12 and D(comm(sy, sz), 2j * sx) < 1e-12 and D(comm(sz, sx), 2j * sy) < 1e-12
    assert D(sp_, sm.conj().T) < 1e-14
    assert D(sx, sp_ + sm) < 1e-12 and D(sy, -1j * (sp_ - sm)) < 1e-12
    n = int(2 * spin + 1)
    casimir = sx @ sx + sy @ sy + sz @ sz
    assert D(casimir, 4 * spin * (spin + 1) * np.eye(n)) < 1e-11
    if spin == 0.5:
        for s in (sx, sy, sz):
            assert D(s @ s, np.eye(2)) < 1e-14
        assert np.array_equal(sx, np.array([[0, 1], [1, 0]])) and np.array_equal(sz, np.diag([1, -1]))
        assert np.array_equal(sy, np.array([[0, -1j], [1j, 0]]))


def test_fock_and_nlevel_known_answers():
    a, ad, num = O.dense(O.destroy(6)), O.dense(O.create(6)), O.dense(O.number(6))
    assert D(ad @ a, num) < 1e-13
    c = a @ ad - ad @ a  # [a, a†] = 1 except at the cutoff
    assert D(c[:-1, :-1], np.eye(6)) < 1e-13
    assert D(a, np.diag(np.sqrt(np.arange(1, 7)), 1)) < 1e-15
    t = O.dense(O.transition(3, 1, 2))
    assert t[0, 1] == 1 and np.count_nonzero(t) == 1
    with pytest.raises(IndexError):
        O.transition(3, 4, 1)


def test_embed_ordering_known_answer():
    """subsystem 1 is the fastest index: op on site 1 of (2,3) acts as kron(I3, op)  (test/test_embed.jl:47-90)"""
    op = np.array([[1, 2], [3, 4]], dtype=complex)
    lt = O.LazyTensor((2, 3), (2, 3), [1], [O.Op((2,), (2,), op)])
    assert np.array_equal(O.dense(lt), np.kron(np.eye(3), op))
    x = np.zeros(6, dtype=complex)
    x[1 + 2 * 2] = 1.0  # |i1=1, i2=2>
    y = O.Ket((2, 3), np.zeros(6))
    O.mul(y, lt, O.Ket((2, 3), x))
    expect = np.zeros(6, dtype=complex)
    expect[0 + 2 * 2], expect[1 + 2 * 2] = 2, 4
    assert np.array_equal(y.data, expect)


def test_generator_is_deterministic():
    a = O.fill_state(64, 9, 0.5)
    b = O.fill_state(32, 9, 0.5, offset=32)
    assert np.array_equal(a[32:], b)
    assert O.state_at(9, 40, 0.5) == a[40]
    assert np.all(np.abs(a.real) <= 0.5) and np.all(np.abs(a.imag) <= 0.5)


# ------------------------------------------------------------------ SURVEY §8(f) rows restated ahead of the device code
def _kron(*ms):
    """a (x) b (x) c with subsystem 1 fastest = kron(c, kron(b, a)) in numpy's convention"""
    out = np.array([[1.0 + 0j]])
    for m in ms:
        out = np.kron(m, out)
    return out


def test_ptrace_identities():
    """test/test_abstractdata.jl:198-219: ptrace(op1 (x) op2 (x) op3, k) == (the others) * tr(op_k); argument errors."""
    rng = np.random.default_rng(7)
    dl, dr = (2, 3, 4), (2, 3, 4)
    o1, o2, o3 = H.rnd(rng, 2, 2), H.rnd(rng, 3, 3), H.rnd(rng, 4, 4)
    o123 = _kron(o1, o2, o3)
    cases = {(3,): _kron(o1, o2) * np.trace(o3), (2,): _kron(o1, o3) * np.trace(o2), (1,): _kron(o2, o3) * np.trace(o1),
             (2, 3): o1 * np.trace(o2) * np.trace(o3), (1, 3): o2 * np.trace(o1) * np.trace(o3),
             (1, 2): o3 * np.trace(o1) * np.trace(o2)}
    for idx, ref in cases.items():
        kl, kr, r = O.ptrace_op(dl, dr, o123, list(idx))
        assert r.shape == ref.shape and np.allclose(r, ref, rtol=1e-13, atol=1e-13)
        assert kl == tuple(d for k, d in enumerate(dl) if k + 1 not in idx)
    with pytest.raises(O.ArgumentError):
        O.ptrace_op(dl, dr, o123, [1, 2, 3])
    with pytest.raises(O.ArgumentError):
        O.ptrace_op((2, 3), (2, 4), H.rnd(rng, 6, 8), [2])      # traced subsystem with unequal dimensions
    # rectangular kept subsystems are fine: trace subsystem 1 of a (2*3) x (2*5) operator
    a = H.rnd(rng, 6, 10)
    _, _, r = O.ptrace_op((2, 3), (2, 5), a, [1])
    ref = sum(a[t::2, t::2] for t in range(2))
    assert np.allclose(r, ref)
    # definition by explicit loops on a random (non-product) operator
    a = H.rnd(rng, 24, 24)
    _, _, r = O.ptrace_op(dl, dr, a, [2])
    ref = np.zeros((8, 8), dtype=complex)
    for i1 in range(2):
        for i3 in range(4):
            for j1 in range(2):
                for j3 in range(4):
                    for t in range(3):
                        ref[i1 + 2 * i3, j1 + 2 * j3] += a[i1 + 2 * t + 6 * i3, j1 + 2 * t + 6 * j3]
    assert np.allclose(r, ref, rtol=1e-13, atol=1e-13)


def test_ptrace_states_and_expect():
    """test/test_abstractdata.jl:233-234: expect(k, op, psi) == expect(op, ptrace(psi, others)); bra = conj of ket"""
    rng = np.random.default_rng(8)
    dims = (2, 3, 4)
    psi = H.rnd(rng, 24)
    psi /= np.linalg.norm(psi)
    o2 = H.rnd(rng, 3, 3)
    full = _kron(np.eye(2), o2, np.eye(4))
    _, _, rho2 = O.ptrace_ket(dims, psi, [1, 3])
    assert np.isclose(np.trace(o2 @ rho2), np.vdot(psi, full @ psi))
    assert np.isclose(np.trace(rho2), 1.0) and np.allclose(rho2, rho2.conj().T)
    _, _, rb = O.ptrace_bra(dims, psi.conj(), [1, 3])
    assert np.allclose(rb, rho2)


def test_lindblad_rhs_equals_the_mul_call_pattern():
    """test/test_sciml_broadcast_interfaces.jl:36-43: the six mul! calls of a master-equation step, through the oracle's own
    mul (sparse H and J on a dense rho), equal the closed form; trace preserving and Hermiticity preserving."""
    rng = np.random.default_rng(9)
    nc = 6
    nf = nc + 1
    a, ad = O.destroy(nc).data, O.create(nc).data
    sm, sp_ = O.sigmam().data, O.sigmap().data
    i2, inf = sp.identity(2, format="csc"), sp.identity(nf, format="csc")
    Hm = (sp.kron(i2, ad @ a) + 0.3 * (sp.kron(sp_, a) + sp.kron(sm, ad))).tocsc()
    Jm = [np.sqrt(0.7) * sp.kron(i2, a).tocsc(), np.sqrt(0.2) * sp.kron(sm, inf).tocsc()]
    D = 2 * nf
    dims = (nf, 2)
    x = H.rnd(rng, D, D)
    rho = x @ x.conj().T
    rho /= np.trace(rho)
    ref = O.lindblad_rhs(Hm.toarray(), [j.toarray() for j in Jm], rho)
    assert abs(np.trace(ref)) < 1e-13 and np.allclose(ref, ref.conj().T)
    Ho = O.Op(dims, dims, Hm)
    st = O.Op(dims, dims, np.asfortranarray(rho))
    out = O.Op(dims, dims, np.zeros((D, D), dtype=complex, order="F"))
    tmp = O.Op(dims, dims, np.zeros((D, D), dtype=complex, order="F"))
    O.mul(out, Ho, st, -1j, 0.0)
    O.mul(out, st, Ho, 1j, 1.0)
    for j in Jm:
        Jo, Jd = O.Op(dims, dims, j), O.Op(dims, dims, sp.csc_matrix(j.conj().T))
        JdJ = O.Op(dims, dims, sp.csc_matrix(j.conj().T @ j))
        O.mul(tmp, Jo, st, 1.0, 0.0)
        O.mul(out, tmp, Jd, 1.0, 1.0)
        O.mul(out, JdJ, st, -0.5, 1.0)
        O.mul(out, st, JdJ, -0.5, 1.0)
    assert H.rel_err(out.data, ref) <= 1e-13

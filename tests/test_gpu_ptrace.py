"""Device partial traces, LazyDirectSum mul! and expect / variance against the oracle's
restatement of src/operators_dense.jl:191-215,311-383, src/spinors.jl:221-247 and src/operators.jl:119-142, plus the
identities the reference's own tests use (test/test_abstractdata.jl:198-234: ptrace of a product state/operator gives the
factors back, tr is preserved)."""
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


@pytest.mark.parametrize("dims,traced", [((2, 3, 4), [2]), ((2, 3, 4), [1, 3]), ((3, 2, 2, 5), [4]), ((3, 2, 2, 5), [1, 2, 3]),
                                         ((2,) * 10, [1, 2, 3, 4, 5, 6]), ((2,) * 10, [2, 5, 9]), ((7, 5), [1]), ((6, 4, 3), [])])
def test_ptrace_dense_operator_vs_oracle(Q, dims, traced):
    rng = np.random.default_rng(900 + len(dims) + len(traced))
    D = int(np.prod(dims))
    a = H.rnd(rng, D, D)
    cb = Q.CompositeBasis([Q.GenericBasis(d) for d in dims])
    got = Q.ptrace(Q.DenseOperator(cb, cb, a), traced)
    if traced:
        _, _, ref = O.ptrace_op(dims, dims, a, traced)
    else:
        ref = a
    assert got.data.shape == ref.shape
    assert H.rel_err(got.to_host(), ref) <= TOL
    # the trace survives (test/test_abstractdata.jl:226-234)
    assert abs(np.trace(got.to_host()) - np.trace(a)) <= 1e-11 * max(1.0, abs(np.trace(a)))


def test_ptrace_rectangular_operator(Q):
    """kept subsystems may have different left / right dimensions; traced ones must be square"""
    rng = np.random.default_rng(911)
    dl, dr = (2, 3, 4), (5, 3, 2)
    a = H.rnd(rng, int(np.prod(dl)), int(np.prod(dr)))
    bl = Q.CompositeBasis([Q.GenericBasis(d) for d in dl])
    br = Q.CompositeBasis([Q.GenericBasis(d) for d in dr])
    got = Q.ptrace(Q.DenseOperator(bl, br, a), [2])
    _, _, ref = O.ptrace_op(dl, dr, a, [2])
    assert got.data.shape == (8, 10) and H.rel_err(got.to_host(), ref) <= TOL
    with pytest.raises(Q.ArgumentError):
        Q.ptrace(Q.DenseOperator(bl, br, a), [1])        # 2 != 5
    with pytest.raises(Q.ArgumentError):
        Q.ptrace(Q.DenseOperator(bl, br, a), [1, 2, 3])  # use tr() instead
    with pytest.raises(Q.ArgumentError):
        Q.ptrace(Q.DenseOperator(bl, br, a), [4])


@pytest.mark.parametrize("dims,traced", [((2, 3, 4), [2]), ((2, 3, 4), [1, 3]), ((3, 2, 2, 5), [1, 2, 3]), ((2,) * 14, list(range(5, 15))),
                                         ((2,) * 14, [1, 3, 5, 7, 9, 11, 13]), ((2,) * 18, list(range(1, 15))), ((33, 17), [1])])
@pytest.mark.parametrize("bra", [False, True])
def test_ptrace_ket_and_bra_vs_oracle(Q, dims, traced, bra):
    rng = np.random.default_rng(930 + len(dims))
    D = int(np.prod(dims))
    psi = H.rnd(rng, D)
    psi /= np.linalg.norm(psi)
    cb = Q.CompositeBasis([Q.GenericBasis(d) for d in dims])
    st = Q.Bra(cb, psi) if bra else Q.Ket(cb, psi)
    got = Q.ptrace(st, traced)
    # exact reference without the D x D outer product: reshape to (kept, traced) and contract
    n = len(dims)
    keep = [k for k in range(n) if (k + 1) not in traced]
    t = psi.reshape(dims, order="F").transpose(keep + [k - 1 for k in traced]).reshape(
        int(np.prod([dims[k] for k in keep])), -1, order="F")
    ref = (t.conj() @ t.T) if bra else (t @ t.conj().T)
    assert H.rel_err(got.to_host(), ref) <= TOL
    if D <= 4096:
        _, _, ref2 = (O.ptrace_bra if bra else O.ptrace_ket)(dims, psi, traced)
        assert H.rel_err(got.to_host(), ref2) <= TOL
    assert abs(np.trace(got.to_host()) - 1.0) <= 1e-12


def test_ptrace_product_state_gives_the_factors_back(Q):
    """test/test_abstractdata.jl:198-224: ptrace(a (x) b (x) c, [1, 3]) == b * tr(a) * tr(c)"""
    rng = np.random.default_rng(950)
    a, b, c = H.rnd(rng, 3, 3), H.rnd(rng, 4, 4), H.rnd(rng, 2, 2)
    full = np.kron(c, np.kron(b, a))   # subsystem 1 fastest
    cb = Q.CompositeBasis([Q.GenericBasis(d) for d in (3, 4, 2)])
    got = Q.ptrace(Q.DenseOperator(cb, cb, full), [1, 3])
    assert H.rel_err(got.to_host(), b * np.trace(a) * np.trace(c)) <= TOL
    red = Q.reduced(Q.DenseOperator(cb, cb, full), [1])
    assert H.rel_err(red.to_host(), a * np.trace(b) * np.trace(c)) <= TOL


def test_lazydirectsum_mul_ket_and_bra(Q):
    rng = np.random.default_rng(960)
    sizes = [3, 5, 8]
    o_ops, q_ops = [], []
    for k, n in enumerate(sizes):
        m = H.sprnd(rng, n, n, 0.5) if k != 1 else H.rnd(rng, n, n)
        pair = H.operator((n,), (n,), m)
        o_ops.append(pair.o)
        q_ops.append(pair.q)
    S = Q.LazyDirectSum(*q_ops)
    D = sum(sizes)
    assert len(S.basis_l) == D and "lazydirectsum" in Q.describe(S)
    for (al, be) in ((1, 0), (0.3 - 0.2j, 1.7), (-1j, 0)):
        x, y0 = H.rnd(rng, D), H.rnd(rng, D)
        ref = O.directsum_mul(y0.copy(), o_ops, x, al, be)
        r = Q.Ket(S.basis_l, y0.copy())
        Q.mul_(r, S, Q.Ket(S.basis_r, x), al, be)
        assert H.rel_err(r.to_host(), ref) <= TOL
        refb = O.directsum_mul(y0.copy(), o_ops, x, al, be, bra=True)
        rb = Q.Bra(S.basis_r, y0.copy())
        Q.mul_(rb, Q.Bra(S.basis_l, x), S, al, be)
        assert H.rel_err(rb.to_host(), refb) <= TOL
    # the reference defines no LazyDirectSum method for operator states
    rho = Q.DenseOperator(S.basis_r, S.basis_r, H.rnd(rng, D, D))
    with pytest.raises(Q.MethodError):
        Q.mul_(Q.DenseOperator(S.basis_l, S.basis_r), S, rho)


def test_expect_and_variance_through_the_c_abi(Q, monkeypatch):
    monkeypatch.setenv("QOB_QREG_MIN_BITS", "12")
    rng = np.random.default_rng(970)
    n = 13
    sx = np.array([[0, 1], [1, 0]], dtype=complex)
    sz = np.array([[1, 0], [0, -1]], dtype=complex)
    dims = (2,) * n
    terms, coefs = [], []
    for i in range(1, n + 1):
        j = i % n + 1
        terms.append(H.lazytensor(dims, dims, [i], [sp.csc_matrix(sx)]))
        coefs.append(rng.uniform(0.5, 1.5))
        terms.append(H.lazytensor(dims, dims, sorted([i, j]), [sp.csc_matrix(sz), sp.csc_matrix(sz)]))
        coefs.append(rng.uniform(0.5, 1.5))
    s = H.lazysum(dims, dims, coefs, terms)
    x = H.rnd(rng, 1 << n)
    x /= np.linalg.norm(x)
    hx = H.ket(dims, np.zeros(1 << n, dtype=complex))
    O.mul(hx.o, s.o, H.ket(dims, x).o, 1.0, 0.0)
    hhx = H.ket(dims, np.zeros(1 << n, dtype=complex))
    O.mul(hhx.o, s.o, hx.o, 1.0, 0.0)
    e_ref = np.vdot(x, hx.o.data)
    v_ref = np.vdot(x, hhx.o.data) - e_ref ** 2
    st = H.ket(dims, x).q
    l0 = Q.launch_count()
    e = Q.expect(s.q, st)
    assert abs(e - e_ref) <= 1e-12 * max(1.0, abs(e_ref))
    assert Q.launch_count() > l0
    v = Q.variance(s.q, st)
    assert abs(v - v_ref) <= 1e-11 * max(1.0, abs(v_ref))

"""Fused master-equation right-hand side (qob_lindblad_*) against the oracle's closed form and against
the reference's own call pattern (six mul! per jump operator, test/test_sciml_broadcast_interfaces.jl:36-43) run through
the device mul!."""
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


def jaynes_cummings(nc, non_hermitian=False):
    nf = nc + 1
    a, ad = O.destroy(nc).data, O.create(nc).data
    sm, spl, sz = O.sigmam().data, O.sigmap().data, O.sigmaz().data
    i2, inf = sp.identity(2, format="csc"), sp.identity(nf, format="csc")
    Hm = (sp.kron(i2, ad @ a) + 0.45 * sp.kron(sz, inf) + 0.1 * (sp.kron(spl, a) + sp.kron(sm, ad))).tocsc()
    if non_hermitian:
        Hm = (Hm + 0.05j * sp.kron(spl, inf)).tocsc()
    J = [sp.kron(i2, a).tocsc(), sp.kron(sm, inf).tocsc()]
    return (nf, 2), Hm.astype(complex), [j.astype(complex) for j in J]


def device_pair(Q, dims, arr):
    return H.denseop(dims, dims, arr)


@pytest.mark.parametrize("nc,non_hermitian", [(6, False), (6, True), (64, False)])
def test_jaynes_cummings_against_closed_form(Q, nc, non_hermitian):
    rng = np.random.default_rng(3)
    dims, Hm, J = jaynes_cummings(nc, non_hermitian)
    D = int(np.prod(dims))
    bas = H._cb(dims)
    rates = [0.7, 0.2]
    L = Q.LindbladRHS(Q.Operator(bas, bas, Hm), [Q.Operator(bas, bas, j) for j in J], rates)
    rho = H.rnd(rng, D, D)
    ref = O.lindblad_rhs(Hm.toarray(), [j.toarray() for j in J], rho, rates)
    r = device_pair(Q, dims, rho)
    for (al, be) in [(1.0, 0.0), (0.5 - 0.2j, 1.5), (-1j, 0.25 + 1j)]:
        y0 = H.rnd(rng, D, D)
        d = device_pair(Q, dims, y0 * (np.nan if be == 0 else 1.0))   # beta = 0 must not read drho
        L.apply_(d.q, r.q, al, be)
        out = d.q.to_host()
        want = al * ref + (be * y0 if be != 0 else 0)
        assert np.all(np.isfinite(out)) and H.rel_err(out, want) <= TOL, (al, be)
    # alpha = 0: only the beta update
    y0 = H.rnd(rng, D, D)
    d = device_pair(Q, dims, y0)
    L.apply_(d.q, r.q, 0.0, 0.3)
    assert H.rel_err(d.q.to_host(), 0.3 * y0) <= TOL
    with pytest.raises(Q.ArgumentError):
        L.apply_(r.q, r.q)


@pytest.mark.parametrize("D", [150, 700])
def test_random_sparse_and_dense_jump_operators(Q, D):
    """general (non-Hermitian) sparse H, sparse jump operators with a few entries per row and one dense jump operator; D not a
    multiple of the block size"""
    rng = np.random.default_rng(D)
    dims = (D,)
    bas = H._cb(dims)

    def sprand():
        m = sp.random(D, D, density=3.0 / D, random_state=np.random.RandomState(int(rng.integers(1 << 30))), format="csc").astype(complex)
        m.data = H.rnd(rng, m.nnz)
        return m

    Hm = sprand()
    J = [sprand(), sprand()]
    Jd = H.rnd(rng, D, D) * (rng.uniform(0, 1, (D, D)) < 2.0 / D)      # host-dense data with mostly zeros
    rates = [0.4, 1.1, 0.05]
    L = Q.LindbladRHS(Q.Operator(bas, bas, Hm), [Q.Operator(bas, bas, j) for j in J] + [Q.Operator(bas, bas, Jd)], rates)
    rho = H.rnd(rng, D, D)
    ref = O.lindblad_rhs(Hm.toarray(), [j.toarray() for j in J] + [Jd], rho, rates)
    r, d = device_pair(Q, dims, rho), device_pair(Q, dims, np.zeros((D, D), dtype=complex))
    L.apply_(d.q, r.q)
    assert H.rel_err(d.q.to_host(), ref) <= TOL
    # no jump operators: the plain commutator
    L0 = Q.LindbladRHS(Q.Operator(bas, bas, Hm))
    L0.apply_(d.q, r.q)
    assert H.rel_err(d.q.to_host(), -1j * (Hm @ rho - rho @ Hm.toarray())) <= TOL


def test_equals_the_reference_call_pattern_on_device(Q):
    """the mul! sequence a master-equation step makes in the reference, run through the device mul!, vs the fused kernel"""
    rng = np.random.default_rng(5)
    dims, Hm, J = jaynes_cummings(64)
    D = int(np.prod(dims))
    bas = H._cb(dims)
    Hq = Q.Operator(bas, bas, Hm)
    x = H.rnd(rng, D, D)
    rho = x @ x.conj().T
    rho /= np.trace(rho)
    r = device_pair(Q, dims, rho)
    out, tmp, fused = (device_pair(Q, dims, np.zeros((D, D), dtype=complex)) for _ in range(3))
    Q.mul_(out.q, Hq, r.q, -1j, 0.0)
    Q.mul_(out.q, r.q, Hq, 1j, 1.0)
    for j in J:
        Jq, Jdq = Q.Operator(bas, bas, j), Q.Operator(bas, bas, sp.csc_matrix(j.conj().T))
        JdJ = Q.Operator(bas, bas, sp.csc_matrix(j.conj().T @ j))
        Q.mul_(tmp.q, Jq, r.q, 1.0, 0.0)
        Q.mul_(out.q, tmp.q, Jdq, 1.0, 1.0)
        Q.mul_(out.q, JdJ, r.q, -0.5, 1.0)
        Q.mul_(out.q, r.q, JdJ, -0.5, 1.0)
    Q.LindbladRHS(Hq, [Q.Operator(bas, bas, j) for j in J]).apply_(fused.q, r.q)
    a, b = fused.q.to_host(), out.q.to_host()
    assert H.rel_err(a, b) <= TOL
    assert abs(np.trace(a)) <= 1e-13 and np.allclose(a, a.conj().T, atol=1e-14)   # trace and Hermiticity preserving


def test_errors(Q):
    bas, bas2 = H._cb((4,)), H._cb((5,))
    Hq = Q.Operator(bas, bas, sp.identity(4, format="csc", dtype=complex))
    with pytest.raises(Q.IncompatibleBases):
        Q.LindbladRHS(Hq, [Q.Operator(bas2, bas2, sp.identity(5, format="csc", dtype=complex))])
    with pytest.raises(Q.ArgumentError):
        Q.LindbladRHS(Hq, [Hq], rates=[-1.0])
    with pytest.raises(Q.ArgumentError):
        Q.LindbladRHS(Hq, [Hq], rates=[1.0, 2.0])
    L = Q.LindbladRHS(Hq, [Hq])
    with pytest.raises(Q.IncompatibleBases):
        L.apply_(H.denseop((5,), (5,), np.zeros((5, 5), dtype=complex)).q, H.denseop((5,), (5,), np.zeros((5, 5), dtype=complex)).q)

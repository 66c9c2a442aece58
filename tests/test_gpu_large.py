"""Full-size (BASELINE config 4: Heisenberg N=28, 4 GiB state) and mid-size checks through properties the
domain offers, because the CPU oracle cannot finish a 2^28 x 84-term apply in seconds:
  * exact per-amplitude spot oracle on a counter-based input (each output amplitude of a 1-/2-site-term
    LazySum depends on <= 1 + n_terms inputs) — SURVEY.md §8c;
  * linearity, Hermiticity (<x|Hx> real), and agreement between the two independent device code paths
    (tile kernel vs generic gather kernel) on the full vector.
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


def heisenberg(Q, n, rng, coefs=None):
    b = Q.SpinBasis(0.5)
    B = Q.tensor(*[b] * n)
    sig = (Q.sigmax(b), Q.sigmay(b), Q.sigmaz(b))
    terms, cf, spec = [], [], []
    for i in range(1, n + 1):
        j = i % n + 1
        idx = sorted([i, j])
        for a, s in enumerate(sig):
            terms.append(Q.LazyTensor(B, idx, (s, s)))
            c = rng.uniform(0.5, 1.5) if coefs is None else coefs[len(cf)]
            cf.append(c)
            spec.append((c, idx, a))
    return B, Q.LazySum(cf, terms), spec


PAULI = [np.array([[0, 1], [1, 0]], dtype=complex), np.array([[0, -1j], [1j, 0]], dtype=complex),
         np.array([[1, 0], [0, -1]], dtype=complex)]


def spot_oracle(spec, index, xval):
    """(H x)[index] from the definition: sum_t c_t sum_j prod_k A_k[i_k, j_k] x[j]  (SURVEY.md §3.1 item 1)."""
    acc = 0.0 + 0.0j
    for c, idx, a in spec:
        A = PAULI[a]
        k1, k2 = idx[0] - 1, idx[1] - 1
        i1, i2 = (index >> k1) & 1, (index >> k2) & 1
        for j1 in (0, 1):
            for j2 in (0, 1):
                w = A[i1, j1] * A[i2, j2]
                if w != 0:
                    jidx = (index & ~((1 << k1) | (1 << k2))) | (j1 << k1) | (j2 << k2)
                    acc += c * w * xval(jidx)
    return acc


@pytest.mark.parametrize("n", [20, 28])
def test_heisenberg_spot_oracle_and_properties(Q, n):
    import torch

    free, _ = torch.cuda.mem_get_info()
    if free < 8 * (16 << n):
        pytest.skip("not enough device memory")
    rng = np.random.default_rng(100 + n)
    B, Hs, spec = heisenberg(Q, n, rng)
    desc = Q.describe(Hs)
    assert "qtile" in desc or "qreg" in desc
    seed, scale = 1234, 2.0 ** (-n / 2)
    x = Q.Ket(B)
    Q.fill_state(x.data, seed, scale)
    y = Q.Ket(B)
    y.data.fill_(float("nan"))  # beta = 0 must not read y
    al = 0.7 - 0.4j
    Q.mul_(y, Hs, x, al, 0.0)
    # --- spot oracle at tile / pass boundaries and random places
    D = 1 << n
    idxs = [0, 1, 7, 8, 4095, 4096, D - 1, D - 2, D // 2, D // 2 - 1, (1 << (n - 1)) + 5, 0x5555555 % D, 0xAAAAAAA % D]
    idxs += [int(v) for v in rng.integers(0, D, 120)]
    it = torch.tensor(idxs, device="cuda")
    got = y.data[it].cpu().numpy()
    ref = np.array([al * spot_oracle(spec, i, lambda j: O.state_at(seed, j, scale)) for i in idxs])
    err = np.abs(got - ref).max() / np.abs(ref).max()
    assert err <= TOL, f"spot oracle: {err:.3e}"
    # --- Hermiticity: <x|Hx> is real for real coefficients
    d = Q.dot(x.data, y.data) / al
    assert abs(d.imag) <= 1e-12 * max(1.0, abs(d.real))
    # --- beta path: y2 = al*H x + be*y  ==  (1+be)*y
    y2 = y.copy()
    Q.mul_(y2, Hs, x, al, 0.5)
    diff = (y2.data - 1.5 * y.data)
    assert np.sqrt(Q.norm2(diff) / Q.norm2(y.data)) <= TOL
    del diff, y2
    # --- linearity: H(a x + b x2) = a Hx + b Hx2
    x2 = Q.Ket(B)
    Q.fill_state(x2.data, seed + 1, scale)
    comb = Q.Ket(B, 0.3 * x.data + (0.2 + 0.9j) * x2.data)
    yc = Q.Ket(B)
    Q.mul_(yc, Hs, comb, al, 0.0)
    del comb
    Q.mul_(y, Hs, x2, (0.2 + 0.9j) * al, 0.3)  # y <- b*al*H x2 + a*(al*H x)
    diff = yc.data - y.data
    assert np.sqrt(Q.norm2(diff) / Q.norm2(yc.data)) <= TOL


@pytest.mark.parametrize("n,T,L", [(22, 12, 3), (22, 13, 3), (21, 11, 4), (22, 12, 5)])
def test_tile_kernel_vs_gather_kernel_full_vector(Q, monkeypatch, n, T, L):
    rng = np.random.default_rng(200 + n)
    coefs = list(rng.uniform(0.5, 1.5, 3 * n))
    monkeypatch.setenv("QOB_QTILE_T", str(T))
    monkeypatch.setenv("QOB_QTILE_L", str(L))
    monkeypatch.setenv("QOB_DISABLE_QREG", "1")   # this test pins the round-1 tile kernel; tests/test_gpu_qreg.py covers the new one
    B, Ht, _ = heisenberg(Q, n, rng, coefs)
    x = Q.randstate(B, seed=5)
    yt = Q.Ket(B)
    assert "qtile" in Q.describe(Ht)
    Q.mul_(yt, Ht, x, 1.0, 0.0)
    monkeypatch.setenv("QOB_DISABLE_QTILE", "1")
    _, Hg, _ = heisenberg(Q, n, rng, coefs)
    assert "qtile" not in Q.describe(Hg)
    yg = Q.Ket(B)
    Q.mul_(yg, Hg, x, 1.0, 0.0)
    diff = yt.data - yg.data
    assert np.sqrt(Q.norm2(diff) / Q.norm2(yg.data)) <= TOL


def test_heisenberg_n20_full_vector_vs_oracle(Q):
    """largest size the reference's scalar recursion (restated in C) finishes in seconds"""
    n = 20
    rng = np.random.default_rng(300)
    dims = (2,) * n
    sx, sy, sz = [sp.csc_matrix(p) for p in PAULI]
    terms, coefs = [], []
    for i in range(1, n + 1):
        j = i % n + 1
        for s in (sx, sy, sz):
            terms.append(H.lazytensor(dims, dims, sorted([i, j]), [s, s]))
            coefs.append(rng.uniform(0.5, 1.5))
    s = H.lazysum(dims, dims, coefs, terms)
    xh = O.fill_state(1 << n, 77, 2.0 ** (-n / 2))
    x, r = H.ket(dims, xh), H.ket(dims, np.zeros(1 << n, dtype=complex))
    O.mul(r.o, s.o, x.o, -1j, 0)
    Q.mul_(r.q, s.q, x.q, -1j, 0)
    assert H.rel_err(r.q.to_host(), r.o.data) <= TOL
    # the device generator is bit-identical to the oracle's
    xd = Q.Ket(s.q.basis_r)
    Q.fill_state(xd.data, 77, 2.0 ** (-n / 2))
    assert np.array_equal(xd.to_host(), xh)

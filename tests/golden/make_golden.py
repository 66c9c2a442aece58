"""Generates tests/golden/*.npz — small known-answer fixtures for the mul! hot path.

The reference (Julia) cannot run in this image and holds no stored vectors, so the EXPECTED outputs here are
computed from the definition with explicit dense Kronecker products in numpy (tensor(a,b) = kron(b,a),
src/operators_dense.jl:134) — independent of both the oracle's lazy/sparse restatements and the CUDA kernels.
Both are then checked against these files (tests/test_golden.py).  Re-run: python tests/golden/make_golden.py
"""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PAULI = [np.array([[0, 1], [1, 0]], dtype=complex), np.array([[0, -1j], [1j, 0]], dtype=complex),
         np.array([[1, 0], [0, -1]], dtype=complex)]


def rnd(rng, *shape):
    return rng.uniform(-1, 1, shape) + 1j * rng.uniform(-1, 1, shape)


def embed(dims_l, dims_r, indices, mats):
    out = np.ones((1, 1), dtype=complex)
    for k in range(len(dims_l)):
        m = mats[indices.index(k + 1)] if (k + 1) in indices else np.eye(dims_l[k], dims_r[k], dtype=complex)
        out = np.kron(m, out)
    return out


def save(name, **kw):
    np.savez(os.path.join(HERE, name + ".npz"), **kw)


def main():
    rng = np.random.default_rng(20261017)
    # 1. LazyTensor with rectangular factors and an untouched rectangular axis (isometry)
    dims_l, dims_r, idx = (2, 3, 4), (3, 2, 4), [1, 3]
    dims_l, dims_r = (2, 3, 4), (3, 2, 5)
    mats = [rnd(rng, 2, 3), rnd(rng, 4, 5)]
    M = 0.1 * embed(dims_l, dims_r, idx, mats)
    x, y0 = rnd(rng, M.shape[1]), rnd(rng, M.shape[0])
    xb, yb0 = rnd(rng, M.shape[0]), rnd(rng, M.shape[1])
    save("lazytensor_rect", dims_l=dims_l, dims_r=dims_r, indices=idx, m0=mats[0], m1=mats[1], factor=0.1, alpha=1.5, beta=2.1,
         x=x, y0=y0, y=1.5 * M @ x + 2.1 * y0, xb=xb, yb0=yb0, yb=1.5 * xb @ M + 2.1 * yb0)
    # 2. Heisenberg chain N=6 (periodic) as a LazySum of 18 two-site terms
    n = 6
    dims = (2,) * n
    coefs, sites, kinds = [], [], []
    Hm = np.zeros((1 << n, 1 << n), dtype=complex)
    for i in range(1, n + 1):
        j = i % n + 1
        for a in range(3):
            c = rng.uniform(0.5, 1.5)
            s = sorted([i, j])
            Hm += c * embed(dims, dims, s, [PAULI[a], PAULI[a]])
            coefs.append(c), sites.append(s), kinds.append(a)
    x, y0 = rnd(rng, 1 << n), rnd(rng, 1 << n)
    rho = rnd(rng, 1 << n, 1 << n)
    save("heisenberg_n6", n=n, coefs=coefs, sites=sites, kinds=kinds, alpha=-1j, beta=0.5, x=x, y0=y0,
         y=-1j * Hm @ x + 0.5 * y0, rho=rho, comm=-1j * (Hm @ rho) + 1j * (rho @ Hm))
    # 3. sparse x dense, both sides and adjoint
    m, k, c = 5, 7, 4
    S = rnd(rng, m, k) * (rng.uniform(0, 1, (m, k)) < 0.5)
    B, R0 = rnd(rng, k, c), rnd(rng, m, c)
    B2, R20 = rnd(rng, c, m), rnd(rng, c, k)
    B3, R30 = rnd(rng, m, c), rnd(rng, k, c)
    save("sparse_gemm", S=S, alpha=0.3 - 0.7j, beta=1.25, B=B, R0=R0, R=(0.3 - 0.7j) * S @ B + 1.25 * R0,
         B2=B2, R20=R20, R2=(0.3 - 0.7j) * B2 @ S + 1.25 * R20,
         B3=B3, R30=R30, R3=(0.3 - 0.7j) * S.conj().T @ B3 + 1.25 * R30)
    # 4. Jaynes-Cummings, Fock cutoff 5 (x) spin-1/2: H rho and rho H
    nf = 6
    a = np.diag(np.sqrt(np.arange(1, nf)), 1).astype(complex)
    num = np.diag(np.arange(nf)).astype(complex)
    sp_ = np.array([[0, 1], [0, 0]], dtype=complex)
    Hjc = 1.0 * np.kron(np.eye(2), num) + 0.45 * np.kron(PAULI[2], np.eye(nf)) + 0.1 * (np.kron(sp_, a) + np.kron(sp_.T, a.T))
    rho = rnd(rng, 2 * nf, 2 * nf)
    save("jaynes_cummings_nf6", nf=nf, H=Hjc, rho=rho, drho=-1j * Hjc @ rho + 1j * rho @ Hjc)


if __name__ == "__main__":
    main()

"""Known-answer fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py from explicit dense Kronecker
products): the oracle must reproduce them (CPU), and so must the CUDA path (-m gpu)."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

import helpers as H
from helpers import O

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PAULI = [np.array([[0, 1], [1, 0]], dtype=complex), np.array([[0, -1j], [1j, 0]], dtype=complex),
         np.array([[1, 0], [0, -1]], dtype=complex)]
TOL = 1e-12


def load(name):
    return np.load(os.path.join(G, name + ".npz"))


class OracleBackend:
    name = "oracle"

    @staticmethod
    def run(op, kind, dims_out, dims_in, x, y0, al, be, batch_dims=None):
        o = op.o
        if kind == "ket":
            r = O.Ket(dims_out, y0.copy())
            O.mul(r, o, O.Ket(dims_in, x), al, be)
        elif kind == "bra":
            r = O.Bra(dims_out, y0.copy())
            O.mul(r, O.Bra(dims_in, x), o, al, be)
        elif kind == "opl":
            r = O.Op(dims_out, batch_dims, y0.copy())
            O.mul(r, o, O.Op(dims_in, batch_dims, x), al, be)
        else:
            r = O.Op(batch_dims, dims_out, y0.copy())
            O.mul(r, O.Op(batch_dims, dims_in, x), o, al, be)
        return r.data


class CudaBackend:
    name = "cuda"

    @staticmethod
    def run(op, kind, dims_out, dims_in, x, y0, al, be, batch_dims=None):
        Q = H.qo()
        q = op.q
        if kind == "ket":
            r = H.ket(dims_out, y0).q
            Q.mul_(r, q, H.ket(dims_in, x).q, al, be)
        elif kind == "bra":
            r = H.bra(dims_out, y0).q
            Q.mul_(r, H.bra(dims_in, x).q, q, al, be)
        elif kind == "opl":
            r = H.denseop(dims_out, batch_dims, y0).q
            Q.mul_(r, q, H.denseop(dims_in, batch_dims, x).q, al, be)
        else:
            r = H.denseop(batch_dims, dims_out, y0).q
            Q.mul_(r, H.denseop(batch_dims, dims_in, x).q, q, al, be)
        return r.to_host()


BACKENDS = [pytest.param(OracleBackend, id="oracle"), pytest.param(CudaBackend, id="cuda", marks=pytest.mark.gpu)]


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("sparse", [False, True])
def test_golden_lazytensor_rect(backend, sparse):
    g = load("lazytensor_rect")
    dl, dr = tuple(int(v) for v in g["dims_l"]), tuple(int(v) for v in g["dims_r"])
    mats = [g["m0"], g["m1"]]
    if sparse:
        mats = [sp.csc_matrix(m) for m in mats]
    op = H.lazytensor(dl, dr, [int(i) for i in g["indices"]], mats, complex(g["factor"]))
    al, be = complex(g["alpha"]), complex(g["beta"])
    assert H.rel_err(backend.run(op, "ket", dl, dr, g["x"], g["y0"], al, be), g["y"]) <= TOL
    assert H.rel_err(backend.run(op, "bra", dr, dl, g["xb"], g["yb0"], al, be), g["yb"]) <= TOL


@pytest.mark.parametrize("backend", BACKENDS)
def test_golden_heisenberg_n6(backend):
    g = load("heisenberg_n6")
    n = int(g["n"])
    dims = (2,) * n
    terms = [H.lazytensor(dims, dims, [int(v) for v in s], [sp.csc_matrix(PAULI[int(a)])] * 2) for s, a in zip(g["sites"], g["kinds"])]
    s = H.lazysum(dims, dims, [float(c) for c in g["coefs"]], terms)
    y = backend.run(s, "ket", dims, dims, g["x"], g["y0"], complex(g["alpha"]), complex(g["beta"]))
    assert H.rel_err(y, g["y"]) <= TOL
    # -i [H, rho] as two mul! calls (the Liouvillian call pattern of BASELINE config 2)
    z = np.zeros_like(g["rho"])
    r1 = backend.run(s, "opl", dims, dims, g["rho"], z, -1j, 0.0, batch_dims=dims)
    r2 = backend.run(s, "opr", dims, dims, g["rho"], np.asarray(r1).reshape(g["rho"].shape, order="F"), 1j, 1.0, batch_dims=dims)
    assert H.rel_err(np.asarray(r2).reshape(g["rho"].shape, order="F"), g["comm"]) <= TOL


@pytest.mark.parametrize("backend", BACKENDS)
def test_golden_sparse_gemm(backend):
    g = load("sparse_gemm")
    S = sp.csc_matrix(g["S"])
    m, k = S.shape
    c = g["B"].shape[1]
    al, be = complex(g["alpha"]), complex(g["beta"])
    op = H.operator((m,), (k,), S)
    assert H.rel_err(backend.run(op, "opl", (m,), (k,), g["B"], g["R0"], al, be, batch_dims=(c,)), g["R"]) <= TOL
    assert H.rel_err(backend.run(op, "opr", (k,), (m,), g["B2"], g["R20"], al, be, batch_dims=(c,)), g["R2"]) <= TOL
    opa = H.operator((k,), (m,), ("adj", S))
    assert H.rel_err(backend.run(opa, "opl", (k,), (m,), g["B3"], g["R30"], al, be, batch_dims=(c,)), g["R3"]) <= TOL


@pytest.mark.parametrize("backend", BACKENDS)
def test_golden_jaynes_cummings(backend):
    g = load("jaynes_cummings_nf6")
    nf = int(g["nf"])
    dims = (nf, 2)
    op = H.operator(dims, dims, sp.csc_matrix(g["H"]))
    z = np.zeros_like(g["rho"])
    r1 = backend.run(op, "opl", dims, dims, g["rho"], z, -1j, 0.0, batch_dims=dims)
    r1 = np.asarray(r1).reshape(g["rho"].shape, order="F")
    r2 = backend.run(op, "opr", dims, dims, g["rho"], r1, 1j, 1.0, batch_dims=dims)
    assert H.rel_err(np.asarray(r2).reshape(g["rho"].shape, order="F"), g["drho"]) <= TOL

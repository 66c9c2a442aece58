"""Build the SAME operator in the oracle's object model and in the qob200 mirror from raw arrays."""
import os
import sys

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import qob_oracle as O  # noqa: E402  (tests are allowed to import the oracle)

C128 = np.complex128


def rnd(rng, *shape):
    return (rng.uniform(-1, 1, shape) + 1j * rng.uniform(-1, 1, shape)).astype(C128)


def sprnd(rng, m, n, density=0.4):
    a = rnd(rng, m, n) * (rng.uniform(0, 1, (m, n)) < density)
    return sp.csc_matrix(a)


class Pair:
    """oracle object + qob200 object for one operator"""

    def __init__(self, o, q):
        self.o, self.q = o, q


def qo():
    import qob200

    return qob200


def _bases(dims):
    Q = qo()
    return [Q.GenericBasis(d) for d in dims]


def _cb(dims):
    Q = qo()
    return Q.CompositeBasis(_bases(dims))


def _basis_for(dims):
    """composite basis for >1 subsystems, plain basis for one"""
    return _cb(dims)


def data_pair(d):
    """raw data spec -> (oracle data, qob200 data).  spec: ndarray | scipy sparse | ('eye', m, n) | ('adj', spec)"""
    Q = qo()
    if isinstance(d, tuple) and d[0] == "eye":
        return O.Eye(d[1], d[2]), Q.Eye(d[1], d[2])
    if isinstance(d, tuple) and d[0] == "adj":
        a, b = data_pair(d[1])
        return O.Adj(a), Q.Adjoint(b)
    if sp.issparse(d):
        return sp.csc_matrix(d), sp.csc_matrix(d)
    return np.asfortranarray(d), np.asfortranarray(d)


def spec_shape(d):
    if isinstance(d, tuple) and d[0] == "eye":
        return (d[1], d[2])
    if isinstance(d, tuple) and d[0] == "adj":
        s = spec_shape(d[1])
        return (s[1], s[0])
    return d.shape


def lazytensor(dims_l, dims_r, indices, datas, factor=1.0):
    Q = qo()
    bl, br = _cb(dims_l), _cb(dims_r)
    o_ops, q_ops = [], []
    for i, d in zip(indices, datas):
        od, qd = data_pair(d)
        o_ops.append(O.Op((dims_l[i - 1],), (dims_r[i - 1],), od))
        q_ops.append(Q.Operator(bl.bases[i - 1], br.bases[i - 1], qd))
    return Pair(O.LazyTensor(dims_l, dims_r, list(indices), o_ops, factor),
                Q.LazyTensor(bl, br, list(indices), tuple(q_ops), factor))


def operator(dims_l, dims_r, data):
    """a plain Operator definition (dense / sparse / adjoint) on composite dims"""
    Q = qo()
    od, qd = data_pair(data)
    return Pair(O.Op(dims_l, dims_r, od), Q.Operator(_cb(dims_l), _cb(dims_r), qd))


def lazysum(dims_l, dims_r, factors, pairs):
    Q = qo()
    return Pair(O.LazySum(dims_l, dims_r, list(factors), [p.o for p in pairs]),
                Q.LazySum(_cb(dims_l), _cb(dims_r), list(factors), [p.q for p in pairs]))


def lazyproduct(pairs, factor=1.0):
    Q = qo()
    return Pair(O.LazyProduct([p.o for p in pairs], factor), Q.LazyProduct([p.q for p in pairs], factor))


def ket(dims, data):
    Q = qo()
    return Pair(O.Ket(dims, data.copy()), Q.Ket(_cb(dims), data.copy()))


def bra(dims, data):
    Q = qo()
    return Pair(O.Bra(dims, data.copy()), Q.Bra(_cb(dims), data.copy()))


def denseop(dims_l, dims_r, data):
    Q = qo()
    return Pair(O.Op(dims_l, dims_r, np.asfortranarray(data.copy())), Q.DenseOperator(_cb(dims_l), _cb(dims_r), data.copy()))


def rel_err(a, b):
    a = np.asarray(a).reshape(-1)
    b = np.asarray(b).reshape(-1)
    nb = np.linalg.norm(b)
    return np.linalg.norm(a - b) / (nb if nb > 0 else 1.0)


def check_mul(op, dims_l, dims_r, rng, kinds=("ket", "bra", "opl", "opr"), scalars=((1, 0), (1.5, 2.1), (-1j, 0), (0.3 - 0.2j, 1)),
              tol=1e-12, nbatch=5):
    """Run mul! on both sides for every state kind / scalar pair and compare (relative 2-norm <= tol)."""
    Q = qo()
    Dl, Dr = int(np.prod(dims_l)), int(np.prod(dims_r))
    worst = 0.0
    for kind in kinds:
        for (al, be) in scalars:
            if kind == "ket":
                x, y0 = rnd(rng, Dr), rnd(rng, Dl)
                s, r = ket(dims_r, x), ket(dims_l, y0)
                O.mul(r.o, op.o, s.o, al, be)
                Q.mul_(r.q, op.q, s.q, al, be)
            elif kind == "bra":
                x, y0 = rnd(rng, Dl), rnd(rng, Dr)
                s, r = bra(dims_l, x), bra(dims_r, y0)
                O.mul(r.o, s.o, op.o, al, be)
                Q.mul_(r.q, s.q, op.q, al, be)
            elif kind == "opl":
                x, y0 = rnd(rng, Dr, nbatch), rnd(rng, Dl, nbatch)
                s, r = denseop(dims_r, (nbatch,), x), denseop(dims_l, (nbatch,), y0)
                O.mul(r.o, op.o, s.o, al, be)
                Q.mul_(r.q, op.q, s.q, al, be)
            else:
                x, y0 = rnd(rng, nbatch, Dl), rnd(rng, nbatch, Dr)
                s, r = denseop((nbatch,), dims_l, x), denseop((nbatch,), dims_r, y0)
                O.mul(r.o, s.o, op.o, al, be)
                Q.mul_(r.q, s.q, op.q, al, be)
            e = rel_err(r.q.to_host(), r.o.data)
            worst = max(worst, e)
            assert e <= tol, f"{kind} alpha={al} beta={be}: rel err {e:.3e}"
    return worst

"""A minimal cvxpy-compatible expression layer, just large enough to run the UNMODIFIED reference class
``direct_data_driven_mpc.direct_data_driven_mpc_controller.DirectDataDrivenMPCController`` in a container that has
no cvxpy (TEST INFRASTRUCTURE: used by ``make_golden.py`` only, never by the product).

It covers exactly the API surface the reference touches (controller.py:434-445, 537-545, 580-581, 626-627, 674-675,
709-716, 736-737, 753, 767, 778, 804-805): ``Variable``, slicing, ``+ - @`` with NumPy constants, ``vstack``,
``==``, ``quad_form``, ``norm(x, 2) ** 2``, ``norm(x, "inf") <= c``, ``Minimize``, ``Problem(...).solve()``,
``.status``, ``.value``.  Every expression is tracked as an affine map of the variables, so ``Problem.solve`` sees
the QP exactly as the reference's own code states it

    min  sum_k w_k (A_k z + b_k)^T M_k (A_k z + b_k)   s.t.  E z = f,   |G z + h|_inf <= c

and solves it with a generic dense KKT solve (LU, least squares when the KKT matrix is singular as in the NOMINAL
scheme) plus a generic active-set loop on the inf-norm rows.  Nothing in here knows about MPC, Hankel matrices or
the variable layout: that knowledge stays in the reference's code, which is what the fixtures are meant to pin.
"""
from __future__ import annotations

import itertools

import numpy as np

_ids = itertools.count()


def _col(a):
    a = np.asarray(a, dtype=float)
    if a.ndim == 0:
        return a.reshape(1, 1)
    if a.ndim == 1:
        return a.reshape(-1, 1)
    return a


class Expression:
    """rows x 1 affine expression  sum_v A_v v + b."""
    __array_ufunc__ = None            # NumPy defers to our reflected operators (ndarray @ Expression, ndarray - Expression)
    __hash__ = object.__hash__

    def __init__(self, terms, const):
        self.terms, self.const = terms, _col(const)

    @property
    def shape(self):
        return self.const.shape

    def __getitem__(self, key):
        if isinstance(key, tuple):
            key = key[0]
        return Expression({v: A[key] for v, A in self.terms.items()}, self.const[key])

    @staticmethod
    def _wrap(other, rows):
        if isinstance(other, Expression):
            return other
        if isinstance(other, Constant):
            other = other.array
        c = _col(other)
        if c.shape[0] == 1 and rows != 1:
            c = np.full((rows, 1), float(c[0, 0]))
        return Expression({}, c)

    def __add__(self, other):
        o = self._wrap(other, self.shape[0])
        terms = dict(self.terms)
        for v, A in o.terms.items():
            terms[v] = terms[v] + A if v in terms else A
        return Expression(terms, self.const + o.const)

    __radd__ = __add__

    def __neg__(self):
        return Expression({v: -A for v, A in self.terms.items()}, -self.const)

    def __sub__(self, other):
        return self + (-self._wrap(other, self.shape[0]))

    def __rsub__(self, other):
        return self._wrap(other, self.shape[0]) + (-self)

    def __mul__(self, s):
        s = float(s)
        return Expression({v: s * A for v, A in self.terms.items()}, s * self.const)

    __rmul__ = __mul__

    def __rmatmul__(self, M):
        M = M.array if isinstance(M, Constant) else np.asarray(M, dtype=float)
        return Expression({v: M @ A for v, A in self.terms.items()}, M @ self.const)

    def __eq__(self, other):            # noqa: D105 - cvxpy semantics: builds a constraint
        return Equality(self - self._wrap(other, self.shape[0]))

    @property
    def value(self):
        out = self.const.copy()
        for v, A in self.terms.items():
            if v._value is None:
                return None
            out = out + A @ v._value
        return out


class Variable(Expression):
    def __init__(self, shape):
        rows = shape[0] if isinstance(shape, tuple) else int(shape)
        self.id, self.size, self._value = next(_ids), rows, None
        super().__init__({self: np.eye(rows)}, np.zeros((rows, 1)))

    @property
    def value(self):
        return self._value


class Constant:
    """Result of vstack over NumPy arrays only (cp.vstack([HLn_ud, HLn_yd]) @ alpha, vstack([u_past, y_past]))."""
    __array_ufunc__ = None

    def __init__(self, array):
        self.array = np.asarray(array, dtype=float)

    def __matmul__(self, expr):
        return expr.__rmatmul__(self.array)


def vstack(items):
    if all(not isinstance(i, Expression) for i in items):
        return Constant(np.vstack([i.array if isinstance(i, Constant) else _col(i) for i in items]))
    exprs = [i if isinstance(i, Expression) else Expression({}, _col(i.array if isinstance(i, Constant) else i))
             for i in items]
    variables = {v for e in exprs for v in e.terms}
    terms = {v: np.vstack([e.terms.get(v, np.zeros((e.shape[0], v.size))) for e in exprs]) for v in variables}
    return Expression(terms, np.vstack([e.const for e in exprs]))


class Equality:
    def __init__(self, expr):
        self.expr = expr


class InfNormBound:
    def __init__(self, expr, bound):
        self.expr, self.bound = expr, float(bound)


class Quadratic:
    """sum_k w_k x_k^T M_k x_k."""

    def __init__(self, parts):
        self.parts = parts

    def __add__(self, other):
        return Quadratic(self.parts + other.parts)

    __radd__ = __add__

    def __mul__(self, s):
        return Quadratic([(float(s) * w, x, M) for w, x, M in self.parts])

    __rmul__ = __mul__


class _Norm2:
    def __init__(self, x):
        self.x = x

    def __pow__(self, k):
        assert k == 2, "only norm(x, 2) ** 2 is supported"
        return Quadratic([(1.0, self.x, np.eye(self.x.shape[0]))])


class _NormInf:
    def __init__(self, x):
        self.x = x

    def __le__(self, bound):
        return InfNormBound(self.x, bound)


def norm(x, p=2):
    if p == 2:
        return _Norm2(x)
    if p == "inf":
        return _NormInf(x)
    raise NotImplementedError(p)


def quad_form(x, M):
    return Quadratic([(1.0, x, np.asarray(M, dtype=float))])


class Minimize:
    def __init__(self, cost):
        self.cost = cost


class Constraint:           # annotation target only (List[cp.Constraint])
    pass


class Problem:
    def __init__(self, objective, constraints):
        self.objective, self.constraints = objective, list(constraints)
        self.status, self.value = None, None

    def _dense(self, expr, offs, nz):
        A = np.zeros((expr.shape[0], nz))
        for v, Av in expr.terms.items():
            A[:, offs[v]:offs[v] + v.size] += Av
        return A, expr.const[:, 0]

    def solve(self, **_):
        exprs = [x for _, x, _ in self.objective.cost.parts] + [c.expr for c in self.constraints]
        variables = sorted({v for e in exprs for v in e.terms}, key=lambda v: v.id)
        offs, nz = {}, 0
        for v in variables:
            offs[v] = nz
            nz += v.size
        P, q, c0 = np.zeros((nz, nz)), np.zeros(nz), 0.0
        for w, x, M in self.objective.cost.parts:
            A, b = self._dense(x, offs, nz)
            Ms = 0.5 * (M + M.T)
            P += w * (A.T @ Ms @ A)
            q += 2.0 * w * (A.T @ (Ms @ b))
            c0 += w * float(b @ Ms @ b)
        eq = [self._dense(c.expr, offs, nz) for c in self.constraints if isinstance(c, Equality)]
        E = np.vstack([a for a, _ in eq]) if eq else np.zeros((0, nz))
        f = -np.concatenate([b for _, b in eq]) if eq else np.zeros(0)
        G, h, bound = np.zeros((0, nz)), np.zeros(0), np.zeros(0)
        for c in self.constraints:
            if isinstance(c, InfNormBound):
                A, b = self._dense(c.expr, offs, nz)
                G, h, bound = np.vstack([G, A]), np.concatenate([h, b]), np.concatenate([bound, np.full(len(b), c.bound)])

        def kkt(Ea, fa):
            me = Ea.shape[0]
            K = np.zeros((nz + me, nz + me))
            K[:nz, :nz] = 2.0 * P
            K[:nz, nz:] = Ea.T
            K[nz:, :nz] = Ea
            rhs = np.concatenate([-q, fa])
            try:
                sol = np.linalg.solve(K, rhs)
                if not np.all(np.isfinite(sol)) or np.abs(K @ sol - rhs).max() > 1e-7 * (1.0 + np.abs(rhs).max()):
                    raise np.linalg.LinAlgError
            except np.linalg.LinAlgError:          # singular KKT (NOMINAL: alpha is not unique): min-norm solution
                sol = np.linalg.lstsq(K, rhs, rcond=1e-13)[0]
            return sol[:nz], sol[nz:]

        z, nu = kkt(E, f)
        if np.abs(E @ z - f).max(initial=0.0) > 1e-6 * (1.0 + np.abs(f).max(initial=0.0)):
            self.status = "infeasible"
            return None
        active = {}                                 # inf-norm row -> +1 / -1
        for _ in range(20 * len(h) + 20):
            if active:
                idx = sorted(active)
                Ea = np.vstack([E, G[idx]])
                fa = np.concatenate([f, [active[j] * bound[j] - h[j] for j in idx]])
                z, nu = kkt(Ea, fa)
                mu = nu[E.shape[0]:]
            else:
                idx, mu = [], np.zeros(0)
                z, nu = kkt(E, f)
            if not len(h):
                break
            r = G @ z + h
            viol = np.abs(r) - bound - 1e-13
            for j in idx:
                viol[j] = -np.inf
            j = int(np.argmax(viol))
            if viol[j] > 0.0:
                active[j] = 1 if r[j] > 0 else -1
                continue
            bad = [(active[jj] * mu[k], jj) for k, jj in enumerate(idx) if active[jj] * mu[k] < -1e-12]
            if bad:
                del active[min(bad)[1]]
                continue
            break
        else:
            self.status = "solver_error"
            return None
        for v in variables:
            v._value = z[offs[v]:offs[v] + v.size].reshape(-1, 1)
        self.status = "optimal"
        self.value = float(z @ P @ z + q @ z + c0)
        self.n_active = len(active)
        return self.value

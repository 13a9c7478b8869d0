"""L-BFGS-B with forward-difference gradients as a coroutine.

``scipy.optimize.minimize(fun, x0, method='L-BFGS-B', bounds=...)`` -- what the reference calls for every stage of a
fit (``pastml/ml.py:231``) -- is a thin Python loop around the reverse-communication routine ``_lbfgsb.setulb``: the
routine asks for the objective and its gradient at a point, scipy evaluates ``fun`` there and at ``n`` forward-difference
points (absolute step 1e-8, flipped at an upper bound) and hands both back.  :func:`minimize_lbfgsb` is that same loop,
with the same routine, the same step rule (scipy's own ``_adjust_scheme_to_bounds``) and the same arithmetic, written as
a generator: instead of calling ``fun`` it YIELDS the list of points it needs (the point and its ``n`` perturbations,
one request) and is sent their values.  The fits of all the columns of a batch can then advance in ONE thread, every
round of requests being one device pass -- no optimiser thread per column taking turns under the interpreter lock
(which was 80 % of the fit's wall time on 100k-tip trees), and the base point travels with its perturbations instead
of in a pass of its own.

The iterates are those of ``scipy.optimize.minimize`` bit for bit (tests/test_host_cpu.py compares them on bounded test
functions); the defaults below are scipy's for L-BFGS-B."""
import numpy as np
from scipy.optimize import _lbfgsb
from scipy.optimize._numdiff import _adjust_scheme_to_bounds, _eps_for_method

try:   # 64-bit integer builds of the routine
    from scipy.optimize._lbfgsb_py import HAS_ILP64
except ImportError:   # pragma: no cover
    HAS_ILP64 = False


class Result(object):
    """The fields of scipy's OptimizeResult that the fit reads."""

    def __init__(self, x, fun, success, nit, nfev, status):
        self.x, self.fun, self.success, self.nit, self.nfev, self.status = x, fun, success, nit, nfev, status


def forward_difference_points(x, lb, ub, eps=1e-8):
    """(points [n][n], dx [n]) of scipy's '2-point' scheme with absolute step ``eps`` at ``x`` (``approx_derivative`` with
    ``abs_step=eps`` and bounds, scipy/optimize/_numdiff.py)."""
    sign_x0 = (x >= 0).astype(x.dtype) * 2 - 1
    h = eps
    dx = ((x + h) - x)
    h = np.where(dx == 0, _eps_for_method(x.dtype, np.dtype(np.float64), '2-point') * sign_x0 * np.maximum(1.0, np.abs(x)), h)
    h = h.astype(x.dtype)
    h, _ = _adjust_scheme_to_bounds(x, h, 1, '1-sided', lb, ub)
    points = []
    for i in range(x.size):
        x1 = np.copy(x)
        x1[i] = x[i] + h[i]
        points.append(x1)
    return points, [(x[i] + h[i]) - x[i] for i in range(x.size)]


def minimize_lbfgsb(x0, bounds, maxcor=10, ftol=2.2204460492503131e-09, gtol=1e-5, eps=1e-8, maxfun=15000,
                    maxiter=15000, maxls=20):
    """Generator.  Yields lists of points (1-D arrays) and expects to be sent the list of their objective values;
    returns (``StopIteration.value``) a :class:`Result`.  ``bounds``: array [n][2]."""
    m = maxcor
    pgtol = gtol
    factr = ftol / np.finfo(float).eps
    bounds = np.asarray(bounds, dtype=np.float64)
    lb, ub = bounds[:, 0].copy(), bounds[:, 1].copy()
    x0 = np.clip(np.asarray(x0, dtype=np.float64).ravel(), lb, ub)
    n = x0.size
    int_dtype = np.int64 if HAS_ILP64 else np.int32
    nbd = np.zeros(n, dtype=int_dtype)
    low_bnd = np.zeros(n, dtype=np.float64)
    upper_bnd = np.zeros(n, dtype=np.float64)
    bounds_map = {(-np.inf, np.inf): 0, (1, np.inf): 1, (1, 1): 2, (-np.inf, 1): 3}
    for i in range(n):
        lo, hi = lb[i], ub[i]
        if not np.isinf(lo):
            low_bnd[i] = lo
            lo = 1
        if not np.isinf(hi):
            upper_bnd[i] = hi
            hi = 1
        nbd[i] = bounds_map[lo, hi]

    nfev = 0
    at = {'x': None, 'f': None, 'g': None}

    def fun_and_grad(x):
        # ScalarFunction.fun_and_grad: nothing is recomputed at the point of the previous request
        nonlocal nfev
        if at['x'] is None or not np.array_equal(x, at['x']):
            xc = np.array(x, dtype=np.float64)
            points, dx = forward_difference_points(xc, lb, ub, eps)
            values = yield [np.copy(xc)] + points
            f = values[0]
            nfev += 1
            g = np.empty(n, dtype=np.float64)
            for i in range(n):
                g[i] = (values[1 + i] - f) / dx[i]
            at['x'], at['f'], at['g'] = xc, f, g
        return at['f'], at['g']

    yield from fun_and_grad(x0)   # ScalarFunction evaluates the starting point when it is built

    x = np.array(x0, dtype=np.float64)
    f = np.array(0.0, dtype=np.float64)
    g = np.zeros((n,), dtype=np.float64)
    wa = np.zeros(2 * m * n + 5 * n + 11 * m * m + 8 * m, np.float64)
    iwa = np.zeros(3 * n, dtype=int_dtype)
    task = np.zeros(2, dtype=int_dtype)
    ln_task = np.zeros(2, dtype=int_dtype)
    lsave = np.zeros(4, dtype=int_dtype)
    isave = np.zeros(44, dtype=int_dtype)
    dsave = np.zeros(29, dtype=np.float64)
    n_iterations = 0
    while True:
        g = g.astype(np.float64)
        _lbfgsb.setulb(m, x, low_bnd, upper_bnd, nbd, f, g, factr, pgtol, wa, iwa, task, lsave, isave, dsave, maxls, ln_task)
        if task[0] == 3:
            f, g = yield from fun_and_grad(x)
        elif task[0] == 1:
            n_iterations += 1
            if n_iterations >= maxiter:
                task[0] = 5
                task[1] = 504
            elif nfev > maxfun:
                task[0] = 5
                task[1] = 502
        else:
            break
    if task[0] == 4:
        warnflag = 0
    elif nfev > maxfun or n_iterations >= maxiter:
        warnflag = 1
    else:
        warnflag = 2
    return Result(x, f, warnflag == 0, n_iterations, nfev, warnflag)

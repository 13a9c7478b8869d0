"""Python handle on one engine problem: a forest x C columns that share K and the model family.

This is a thin, array-level wrapper over the C ABI (``include/pastml_b200.h``); the pastml-compatible API
(``acr``, ``ml_acr``, ``parsimonious_acr``, model classes) sits on top of it in ``pastml_b200.ml`` etc.
"""
import ctypes as C

import numpy as np

from pastml_b200 import _lib
from pastml_b200._lib import (MODEL_F81, MODEL_HKY, MODEL_EIG, MASK_ANNOTATION, MASK_JOINT, MASK_MAP, MASK_MPPA,
                              MASK_CUSTOM, PARS_ACCTRAN, PARS_DOWNPASS, PARS_DELTRAN)


class EngineError(RuntimeError):
    pass


class ZeroLikelihoodError(EngineError):
    """Device-side twin of the reference's PastMLLikelihoodError (``/root/reference/pastml/ml.py:139-145``)."""

    def __init__(self, column, param_set, parent, child):
        EngineError.__init__(self, 'zero likelihood in column {} between nodes {} and {}'.format(column, parent, child))
        self.column, self.param_set, self.parent, self.child = column, param_set, parent, child


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def pack_params(kind, K, sf, tau, tau_factor, pi, mu=None, kappa=None, d=None, A=None, Ainv=None):
    """One lane's parameter block (layout documented in include/pastml_b200.h)."""
    stride = param_stride(kind, K)
    out = np.zeros(stride, dtype=np.float64)
    out[0], out[1], out[2] = sf, tau, tau_factor
    pi = np.asarray(pi, dtype=np.float64)
    out[4:4 + K] = pi
    if kind == MODEL_F81:
        if mu is None:
            with np.errstate(divide='ignore'):
                mu = np.float64(1.) / (1. - pi.dot(pi))   # F81Model.get_mu, same numpy expression
        out[3] = mu
    elif kind == MODEL_HKY:
        out[3] = kappa
    else:
        out[4 + K:4 + 2 * K] = np.asarray(d, dtype=np.float64)
        out[4 + 2 * K:4 + 2 * K + K * K] = np.asarray(A, dtype=np.float64).reshape(-1)
        out[4 + 2 * K + K * K:] = np.asarray(Ainv, dtype=np.float64).reshape(-1)
    return out


_STRIDES = {}


def param_stride(kind, K):
    """Doubles per lane of a parameter block (pml_param_stride; cached: it is asked for with every packed block)."""
    key = (int(kind), int(K))
    if key not in _STRIDES:
        _STRIDES[key] = int(_lib.load().pml_param_stride(*key))
    return _STRIDES[key]


class Problem(object):
    """Device-resident ACR problem.

    :param flat: :class:`pastml_b200.tree.FlatForest`
    :param annot_mask: uint64 [C][N] annotated state sets (0 = none)
    :param has_state: uint8 [C][N] node carries the annotation feature
    """

    def __init__(self, flat, n_states, model_kind, annot_mask, has_state, max_param_sets=1, device=0):
        self.lib = _lib.load()
        annot_mask = np.ascontiguousarray(np.atleast_2d(annot_mask), dtype=np.uint64)
        has_state = np.ascontiguousarray(np.atleast_2d(has_state), dtype=np.uint8)
        self.C, self.N = annot_mask.shape
        if self.N != flat.n_nodes or has_state.shape != annot_mask.shape:
            raise ValueError('mask arrays must be [columns][nodes]')
        self.K, self.kind, self.B = int(n_states), int(model_kind), int(max_param_sets)
        self.stride = param_stride(self.kind, self.K)
        self.flat = flat
        handle = C.c_void_p()
        rc = self.lib.pml_problem_create(self.N, _p(flat.parent), _p(flat.dist), self.C, self.K, self.kind,
                                         _p(annot_mask), _p(has_state), self.B, device, C.byref(handle))
        self._h = handle if rc == 0 else None
        self._check(rc)

    # ------------------------------------------------------------------ plumbing
    def _check(self, rc):
        if rc == 0:
            return
        if rc == _lib.PML_ERR_ZERO_LIKELIHOOD:
            col, ps, par, ch = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
            self.lib.pml_zero_likelihood_info(self._h, C.byref(col), C.byref(ps), C.byref(par), C.byref(ch))
            raise ZeroLikelihoodError(col.value, ps.value, par.value, ch.value)
        raise EngineError('pastml_b200 engine error {}: {}'.format(rc, _lib.last_error()))

    def close(self):
        if getattr(self, '_h', None) is not None:
            self.lib.pml_problem_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _params(self, params, n_sets):
        params = np.ascontiguousarray(params, dtype=np.float64)
        if params.size != n_sets * self.C * self.stride:
            raise ValueError('params must be [n_sets={}][C={}][stride={}]'.format(n_sets, self.C, self.stride))
        return params

    # ------------------------------------------------------------------ passes
    def loglik(self, params, n_sets=1, mask_kind=MASK_ANNOTATION, is_marginal=True, alter=True, tolerate_zero=False):
        """[n_sets][C] log-likelihoods (get_bottom_up_loglikelihood summed over the trees).

        With ``tolerate_zero`` a zero-likelihood lane does not raise: the call returns (values, {lane: (parent,
        child)}) with -inf in the failing lanes."""
        params = self._params(params, n_sets)
        out = np.empty((n_sets, self.C), dtype=np.float64)
        rc = self.lib.pml_loglik(self._h, _p(params), n_sets, mask_kind, 1 if is_marginal else 0,
                                 1 if alter else 0, _p(out))
        if not tolerate_zero:
            self._check(rc)
            return out
        bad = {}
        if rc == _lib.PML_ERR_ZERO_LIKELIHOOD:
            pairs = np.empty((n_sets * self.C, 2), dtype=np.int32)
            self._check(self.lib.pml_zero_likelihood_all(self._h, _p(pairs), n_sets * self.C))
            bad = {int(j): (int(pairs[j, 0]), int(pairs[j, 1])) for j in np.flatnonzero(pairs[:, 0] >= 0)}
        else:
            self._check(rc)
        return out, bad

    def supports_lane_lists(self):
        return bool(self.lib.pml_supports_lane_lists(self._h))

    def loglik_lanes(self, params, lane_columns):
        """The objective for an explicit list of lanes: lane i = column ``lane_columns[i]`` with the parameter block
        ``params[i]``; returns (values [n_lanes], {lane: (parent, child)} for zero-likelihood lanes).  The number of
        lanes must be a multiple of 4 (pad with copies of a lane)."""
        params = np.ascontiguousarray(params, dtype=np.float64)
        cols = np.ascontiguousarray(lane_columns, dtype=np.int32)
        n = len(cols)
        if params.shape != (n, self.stride):
            raise ValueError('params must be [n_lanes][{}]'.format(self.stride))
        out = np.empty(n, dtype=np.float64)
        rc = self.lib.pml_loglik_lanes(self._h, _p(params), _p(cols), n, _p(out))
        bad = {}
        if rc == _lib.PML_ERR_ZERO_LIKELIHOOD:
            pairs = np.empty((n, 2), dtype=np.int32)
            self._check(self.lib.pml_zero_likelihood_all(self._h, _p(pairs), n))
            bad = {int(j): (int(pairs[j, 0]), int(pairs[j, 1])) for j in np.flatnonzero(pairs[:, 0] >= 0)}
        else:
            self._check(rc)
        return out, bad

    def joint(self, params, want_states=True):
        params = self._params(params, 1)
        lh = np.empty(self.C, dtype=np.float64)
        states = np.empty((self.N, self.C), dtype=np.int32) if want_states else None
        self._check(self.lib.pml_joint(self._h, _p(params), _p(lh), _p(states)))
        return lh, states

    def marginal(self, params, force_joint=True, want_probs=True, want_node_lh=False, want_masks=True,
                 reuse_bottom_up=False, results_on_device=False):
        """``reuse_bottom_up``: the bottom-up state of the immediately preceding ``loglik(params)`` call (one set,
        annotation masks, marginal, alter) is reused instead of recomputed."""
        params = self._params(params, 1)
        lh = np.empty(self.C, dtype=np.float64)
        probs = np.empty((self.N, self.C, self.K), dtype=np.float64) if want_probs else None
        node_lh = np.empty((self.N, self.C), dtype=np.float64) if want_node_lh else None
        map_mask = np.empty((self.N, self.C), dtype=np.uint64) if want_masks else None
        mppa_mask = np.empty((self.N, self.C), dtype=np.uint64) if want_masks else None
        k_hist = np.empty((self.C, self.K + 1), dtype=np.int64)
        self._check(self.lib.pml_marginal(self._h, _p(params), (1 if force_joint else 0) | (2 if reuse_bottom_up else 0) | (4 if results_on_device else 0),
                                          _p(lh), _p(probs),
                                          _p(node_lh), _p(map_mask), _p(mppa_mask), _p(k_hist)))
        return dict(loglh=lh, probs=probs, node_log10lh=node_lh, map_mask=map_mask, mppa_mask=mppa_mask,
                    k_hist=k_hist)

    def mppa_finalize(self, force_joint=True, want_masks=True):
        mppa_mask = np.empty((self.N, self.C), dtype=np.uint64) if want_masks else None
        k_hist = np.empty((self.C, self.K + 1), dtype=np.int64)
        self._check(self.lib.pml_mppa_finalize(self._h, 1 if force_joint else 0, _p(mppa_mask), _p(k_hist)))
        return mppa_mask, k_hist

    def parsimony(self, methods=PARS_ACCTRAN | PARS_DOWNPASS | PARS_DELTRAN, want_masks=True):
        """Returns {method name: (mask[N][C] or None, steps[C])}."""
        names = [n for n, f in (('ACCTRAN', PARS_ACCTRAN), ('DOWNPASS', PARS_DOWNPASS), ('DELTRAN', PARS_DELTRAN))
                 if methods & f]
        masks = np.empty((len(names), self.N, self.C), dtype=np.uint64) if want_masks else None
        steps = np.empty((len(names), self.C), dtype=np.int64)
        self._check(self.lib.pml_parsimony(self._h, methods, _p(masks), _p(steps)))
        return {n: (masks[i] if want_masks else None, steps[i]) for i, n in enumerate(names)}

    def set_custom_masks(self, mask_cn):
        mask_cn = np.ascontiguousarray(mask_cn, dtype=np.uint64)
        if mask_cn.shape != (self.C, self.N):
            raise ValueError('custom masks must be [C][N]')
        self._check(self.lib.pml_set_custom_masks(self._h, _p(mask_cn)))

    def get_masks(self, mask_kind):
        out = np.empty((self.N, self.C), dtype=np.uint64)
        self._check(self.lib.pml_get_masks(self._h, mask_kind, _p(out)))
        return out

    def get_initial(self):
        has = np.empty((self.N, self.C), dtype=np.uint8)
        init = np.empty((self.N, self.C), dtype=np.uint64)
        self._check(self.lib.pml_get_initial(self._h, _p(has), _p(init)))
        return has, init

    def bottom_up_state(self):
        """(vec [NI][K][C], ex [NI][C], allowed [N][C], altered [N][C]) left by the last single-set sum-product pass;
        the slot of internal node v is ``flat.slot[v]``."""
        vec = np.empty((self.flat.n_internal, self.K, self.C), dtype=np.float64)
        ex = np.empty((self.flat.n_internal, self.C), dtype=np.int32)
        allowed = np.empty((self.N, self.C), dtype=np.uint64)
        altered = np.empty((self.N, self.C), dtype=np.uint8)
        self._check(self.lib.pml_get_bottom_up(self._h, _p(vec), _p(ex), _p(allowed), _p(altered)))
        return vec, ex, allowed, altered

    # ------------------------------------------------------------------ subtree sharding hooks
    def root_partials(self, n_sets=1):
        """Bottom-up vectors of the roots after a loglik pass: (vec [T][K][J], ex [T][J])."""
        T, J = len(self.flat.roots), n_sets * self.C
        vec = np.empty((T, self.K, J), dtype=np.float64)
        ex = np.empty((T, J), dtype=np.int32)
        self._check(self.lib.pml_get_root_partials(self._h, n_sets, _p(vec), _p(ex)))
        return vec, ex

    def set_tip_partials(self, nodes, vec, ex, n_sets=1):
        nodes = np.ascontiguousarray(nodes, dtype=np.int32)
        vec = np.ascontiguousarray(vec, dtype=np.float64)
        ex = np.ascontiguousarray(ex, dtype=np.int32)
        if vec.shape != (len(nodes), self.K, n_sets * self.C) or ex.shape != (len(nodes), n_sets * self.C):
            raise ValueError('external partials must be [n][K][J] and [n][J]')
        self._check(self.lib.pml_set_tip_partials(self._h, len(nodes), _p(nodes), _p(vec), _p(ex), n_sets))
        self._n_ext = len(nodes)

    def tip_td(self):
        n = self._n_ext
        vec = np.empty((n, self.K, self.C), dtype=np.float64)
        ex = np.empty((n, self.C), dtype=np.int32)
        self._check(self.lib.pml_get_tip_td(self._h, _p(vec), _p(ex)))
        return vec, ex

    def set_root_td(self, vec, ex):
        T = len(self.flat.roots)
        vec = np.ascontiguousarray(vec, dtype=np.float64)
        ex = np.ascontiguousarray(ex, dtype=np.int32)
        if vec.shape != (T, self.K, self.C) or ex.shape != (T, self.C):
            raise ValueError('root top-down vectors must be [T][K][C] and [T][C]')
        self._check(self.lib.pml_set_root_td(self._h, _p(vec), _p(ex)))

    def last_call_stats(self):
        ms, n = C.c_double(), C.c_int64()
        self._check(self.lib.pml_last_call_stats(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def marginal_counts(self, params, n_repetitions, seed):
        """[C][K][K] expected transition counts per sampled scenario (pml_marginal_counts)."""
        params = self._params(params, 1)
        out = np.empty((self.C, self.K, self.K), dtype=np.float64)
        self._check(self.lib.pml_marginal_counts(self._h, _p(params), int(n_repetitions), int(seed) & (2 ** 64 - 1), _p(out)))
        return out

    def last_kernels(self):
        """{'bottom_up': {kernel: (launches, items)}, 'top_down': {...}} of the last launch sequences, as the engine
        recorded them when it enqueued the sequence (items: the internal nodes / nodes those launches cover)."""
        buf = C.create_string_buffer(1024)
        self._check(self.lib.pml_last_kernels(self._h, buf, len(buf)))
        out = {}
        for part in buf.value.decode().split(';'):
            label, _, items = part.partition(':')
            out[label] = {k: tuple(int(x) for x in v.split('/')) for k, v in (item.split('=') for item in items.split(',') if item)}
        return out

    def bench_pass(self, params, which, n_iter, n_sets=1, flush_l2=True):
        params = self._params(params, n_sets)
        avg, best, n = C.c_double(), C.c_double(), C.c_int64()
        self._check(self.lib.pml_bench_pass(self._h, _p(params), n_sets, which, n_iter, 1 if flush_l2 else 0,
                                            C.byref(avg), C.byref(best), C.byref(n)))
        return avg.value, best.value, n.value

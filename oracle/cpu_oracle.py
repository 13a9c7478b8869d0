"""TEST INFRASTRUCTURE ONLY -- Python front of the CPU oracle (``oracle/pastml_oracle.c``).

Builds the C restatement with gcc into ``oracle/_build/`` and chains its functions in the order
``pastml.ml.ml_acr`` / ``pastml.parsimony.parsimonious_acr`` do (``/root/reference/pastml/ml.py:640-750``,
``/root/reference/pastml/parsimony.py:245-331``) at FIXED model parameters.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may import this module;
the product (``pastml_b200``) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD_DIR = os.path.join(HERE, '_build')
LIB_PATH = os.path.join(BUILD_DIR, 'libpastml_oracle.so')
SRC_PATH = os.path.join(HERE, 'pastml_oracle.c')

F81, HKY, EIG = 0, 1, 2

ORC_OK, ORC_ERR_ZERO_LIKELIHOOD, ORC_ERR_EMPTY_RESCALE, ORC_ERR_ALLOC = 0, 1, 2, 3


class OracleZeroLikelihood(Exception):
    """The reference would raise PastMLLikelihoodError (ml.py:139-145)."""

    def __init__(self, parent, child):
        Exception.__init__(self, 'zero likelihood between parent {} and child {}'.format(parent, child))
        self.parent, self.child = parent, child


def build(force=False):
    """Compiles the oracle (gcc -O2, no fast-math, no FMA contraction so that sums keep their written order)."""
    if not force and os.path.exists(LIB_PATH) and os.path.getmtime(LIB_PATH) >= os.path.getmtime(SRC_PATH):
        return LIB_PATH
    os.makedirs(BUILD_DIR, exist_ok=True)
    cmd = ['gcc', '-O2', '-fPIC', '-shared', '-std=c99', '-ffp-contract=off', '-fno-fast-math',
           '-o', LIB_PATH, SRC_PATH, '-lm']
    subprocess.check_call(cmd)
    return LIB_PATH


class _Model(C.Structure):
    _fields_ = [('kind', C.c_int32), ('K', C.c_int32), ('sf', C.c_double), ('tau', C.c_double),
                ('tau_factor', C.c_double), ('mu', C.c_double), ('kappa', C.c_double),
                ('pi', C.c_void_p), ('d', C.c_void_p), ('A', C.c_void_p), ('Ainv', C.c_void_p)]


class _Tree(C.Structure):
    _fields_ = [('n_nodes', C.c_int32), ('parent', C.c_void_p), ('first_child', C.c_void_p),
                ('n_children', C.c_void_p), ('dist', C.c_void_p)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB_PATH)
        _lib.orc_pars_steps.restype = C.c_int64
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class OracleModel(object):
    """Fixed-parameter description of a substitution model (what ``Model.get_Pij_t`` needs)."""

    def __init__(self, kind, pi, sf, tau=0., tau_factor=1., kappa=4., d=None, A=None, Ainv=None):
        self.kind = kind
        self.pi = np.ascontiguousarray(pi, dtype=np.float64)
        self.K = len(self.pi)
        self.sf, self.tau, self.tau_factor, self.kappa = float(sf), float(tau), float(tau_factor), float(kappa)
        with np.errstate(divide='ignore'):
            # F81Model.get_mu, /root/reference/pastml/models/F81Model.py:26 (same numpy expression, same bits)
            self.mu = float(np.float64(1.) / (1. - self.pi.dot(self.pi)))
        self.d = None if d is None else np.ascontiguousarray(d, dtype=np.float64)
        self.A = None if A is None else np.ascontiguousarray(A, dtype=np.float64)
        self.Ainv = None if Ainv is None else np.ascontiguousarray(Ainv, dtype=np.float64)

    def c_struct(self):
        return _Model(self.kind, self.K, self.sf, self.tau, self.tau_factor, self.mu, self.kappa,
                      _ptr(self.pi), _ptr(self.d), _ptr(self.A), _ptr(self.Ainv))

    def pij(self, t):
        P = np.empty((self.K, self.K), dtype=np.float64)
        m = self.c_struct()
        lib().orc_pij(C.byref(m), C.c_double(t), _ptr(P))
        return P


class OracleTree(object):
    def __init__(self, flat):
        self.flat = flat
        self.n = flat.n_nodes
        self._c = _Tree(flat.n_nodes, _ptr(flat.parent), _ptr(flat.first_child), _ptr(flat.n_children),
                        _ptr(flat.dist))

    def ref(self):
        return C.byref(self._c)


def _u64(a):
    return np.ascontiguousarray(a, dtype=np.uint64)


def alter_zero_clusters(tree, has_state, K, allowed, initial, has_initial):
    altered = np.zeros(tree.n, dtype=np.uint8)
    n_alt = lib().orc_alter_zero_clusters(tree.ref(), _ptr(has_state), C.c_int(K), _ptr(allowed), _ptr(initial),
                                          _ptr(has_initial), _ptr(altered))
    return altered if n_alt else None


def bottom_up(tree, model, allowed, is_marginal=True, want_ptr=None):
    """Returns (loglh, bu, bu_sf, joint_ptr) -- get_bottom_up_loglikelihood without alter/unalter."""
    n, K = tree.n, model.K
    bu = np.zeros((n, K), dtype=np.float64)
    bu_sf = np.zeros(n, dtype=np.float64)
    ptr = np.zeros((n, K), dtype=np.int32) if not is_marginal else None
    lh = C.c_double(0.)
    err = np.zeros(2, dtype=np.int32)
    m = model.c_struct()
    rc = lib().orc_bottom_up(tree.ref(), C.byref(m), _ptr(allowed), C.c_int(1 if is_marginal else 0),
                             _ptr(bu), _ptr(bu_sf), _ptr(ptr), C.byref(lh), _ptr(err))
    if rc == ORC_ERR_ZERO_LIKELIHOOD:
        raise OracleZeroLikelihood(int(err[0]), int(err[1]))
    if rc != ORC_OK:
        raise RuntimeError('oracle bottom-up failed with code {}'.format(rc))
    return lh.value, bu, bu_sf, ptr


def loglikelihood(tree, model, allowed, has_state, is_marginal=True, alter=True, initial=None, has_initial=None):
    """get_bottom_up_loglikelihood(tree, character, model, is_marginal, alter) for marginal mode, including the
    alter / unalter bracket (``/root/reference/pastml/ml.py:100-119``).  ``allowed`` is restored on return;
    ``initial`` / ``has_initial`` (optional) keep the '.initial' side effect."""
    allowed = _u64(allowed).copy()
    if initial is None:
        initial = np.zeros(tree.n, dtype=np.uint64)
        has_initial = np.zeros(tree.n, dtype=np.uint8)
    if alter and model.tau == 0:
        alter_zero_clusters(tree, has_state, model.K, allowed, initial, has_initial)
    return bottom_up(tree, model, allowed, is_marginal)[0]


def ml_acr_fixed(flat, model, annot_mask, has_state, method='MPPA', force_joint=True):
    """The state machine of ``ml_acr`` (``/root/reference/pastml/ml.py:640-750``) for one character at fixed
    parameters (the likelihood optimisation, ml.py:656-658, only evaluates the objective, which leaves no trace
    except the '.initial' features of altered nodes, reproduced here by one evaluation).

    annot_mask[N] uint64: annotated state set of every node (0 = no / empty annotation);
    has_state[N] uint8: the node carries the annotation feature (ml.py:329-331).
    Returns a dict of numpy arrays / scalars (masks are uint64 bit sets in node-id order).
    """
    tree = OracleTree(flat)
    n, K = flat.n_nodes, model.K
    full = np.uint64((1 << K) - 1) if K < 64 else np.uint64(0xFFFFFFFFFFFFFFFF)
    annot_mask = _u64(annot_mask)
    has_state = np.ascontiguousarray(has_state, dtype=np.uint8)
    L = lib()
    out = {}

    def init_allowed():
        # initialize_allowed_states, ml.py:293-318
        return np.where(annot_mask == 0, full, annot_mask).astype(np.uint64)

    initial = np.zeros(n, dtype=np.uint64)
    has_initial = np.zeros(n, dtype=np.uint8)

    def bu_pass(allowed, marginal=True, alter=True):
        # get_bottom_up_loglikelihood, ml.py:82-121
        altered = None
        if alter and model.tau == 0:
            altered = alter_zero_clusters(tree, has_state, K, allowed, initial, has_initial)
        lh, bu, bu_sf, ptr = bottom_up(tree, model, allowed, marginal)
        if altered is not None and marginal:
            L.orc_unalter_allowed(C.c_int(n), _ptr(altered), _ptr(allowed), _ptr(initial))
        return lh, bu, bu_sf, ptr, altered

    # optimise_likelihood: initialize_allowed_states + objective evaluations, ml.py:865-871
    allowed = init_allowed()
    out['log_likelihood'] = bu_pass(allowed)[0]

    m = model.c_struct()
    joint_state = None
    if method != 'MAP':
        # ml.py:687-695
        lh, bu, bu_sf, ptr, altered = bu_pass(allowed, marginal=False)
        out['log_likelihood_restricted_JOINT'] = lh
        joint_state = np.zeros(n, dtype=np.int32)
        L.orc_choose_joint(tree.ref(), C.byref(m), _ptr(bu), _ptr(ptr), _ptr(altered), _ptr(initial),
                           _ptr(joint_state), _ptr(allowed))
        out['joint_state'] = joint_state.copy()
        out['joint_mask'] = allowed.copy()

    if method in ('MPPA', 'MAP', 'ML', 'ALL'):
        # ml.py:697-714
        allowed = init_allowed()
        altered = None
        if model.tau == 0:
            altered = alter_zero_clusters(tree, has_state, K, allowed, initial, has_initial)
        lh0, bu, bu_sf, _, _ = bu_pass(allowed, alter=False)
        td = np.zeros((n, K), dtype=np.float64)
        td_sf = np.zeros(n, dtype=np.float64)
        rc = L.orc_top_down(tree.ref(), C.byref(m), _ptr(bu), _ptr(bu_sf), _ptr(td), _ptr(td_sf))
        if rc != ORC_OK:
            raise RuntimeError('oracle top-down failed with code {}'.format(rc))
        lh = np.zeros((n, K), dtype=np.float64)
        lh_sf = np.zeros(n, dtype=np.float64)
        probs = np.zeros((n, K), dtype=np.float64)
        rc = L.orc_marginal(tree.ref(), C.byref(m), _ptr(allowed), _ptr(bu), _ptr(bu_sf), _ptr(td), _ptr(td_sf),
                            _ptr(lh), _ptr(lh_sf), _ptr(probs))
        if rc != ORC_OK:
            raise RuntimeError('oracle marginal failed with code {}'.format(rc))
        out['marginal_probabilities'] = probs
        out['marginal_bu'], out['marginal_bu_sf'], out['td'], out['td_sf'] = bu, bu_sf, td, td_sf
        if altered is not None:
            L.orc_unalter_allowed(C.c_int(n), _ptr(altered), _ptr(allowed), _ptr(initial))
        L.orc_choose_map(C.c_int(n), C.c_int(K), _ptr(lh), _ptr(initial), _ptr(has_initial), _ptr(allowed))
        out['lh_after_map'] = lh.copy()
        out['lh_sf'] = lh_sf
        out['map_mask'] = allowed.copy()
        out['log_likelihood_restricted_MAP'] = bu_pass(allowed)[0]

        if method in ('MPPA', 'ML', 'ALL'):
            # (the ALL branch additionally runs parsimony + restricted likelihoods in between, ml.py:718-733;
            # those passes can alter '.initial' and are replayed by the caller when needed)
            k_hist = np.zeros(K + 1, dtype=np.int64)
            best_k = np.zeros(n, dtype=np.int32)
            L.orc_choose_mppa(C.c_int(n), C.c_int(K), _ptr(lh), _ptr(initial), _ptr(has_initial),
                              _ptr(joint_state), C.c_int(1 if force_joint else 0), _ptr(allowed), _ptr(k_hist),
                              _ptr(best_k))
            out['mppa_mask'] = allowed.copy()
            out['mppa_k'] = best_k
            out['k_hist'] = k_hist
            num_scenarios = 1
            for k in range(1, K + 1):
                num_scenarios *= k ** int(k_hist[k])
            out['num_scenarios'] = num_scenarios
            out['num_unresolved_nodes'] = int(k_hist[2:].sum())
            out['num_states'] = int((np.arange(K + 1) * k_hist).sum())
            out['log_likelihood_restricted_MPPA'] = bu_pass(allowed)[0]
    return out


def marginal_pipeline(flat, model, annot_mask, has_state, force_joint=False, joint_state=None):
    """The fixed-parameter MPPA pipeline the benchmark times (metric i): zero-branch
    alteration, bottom-up pass, top-down pass, marginal likelihoods + probabilities, MAP and MPPA selection
    (``/root/reference/pastml/ml.py:697-712, 735-742``).  Returns (loglh, probs, map_mask, mppa_mask)."""
    tree = OracleTree(flat)
    n, K = flat.n_nodes, model.K
    full = np.uint64((1 << K) - 1) if K < 64 else np.uint64(0xFFFFFFFFFFFFFFFF)
    annot_mask = _u64(annot_mask)
    has_state = np.ascontiguousarray(has_state, dtype=np.uint8)
    L = lib()
    m = model.c_struct()
    allowed = np.where(annot_mask == 0, full, annot_mask).astype(np.uint64)
    initial = np.zeros(n, dtype=np.uint64)
    has_initial = np.zeros(n, dtype=np.uint8)
    altered = alter_zero_clusters(tree, has_state, K, allowed, initial, has_initial) if model.tau == 0 else None
    loglh, bu, bu_sf, _ = bottom_up(tree, model, allowed, True)
    td = np.zeros((n, K), dtype=np.float64)
    td_sf = np.zeros(n, dtype=np.float64)
    L.orc_top_down(tree.ref(), C.byref(m), _ptr(bu), _ptr(bu_sf), _ptr(td), _ptr(td_sf))
    lh = np.zeros((n, K), dtype=np.float64)
    lh_sf = np.zeros(n, dtype=np.float64)
    probs = np.zeros((n, K), dtype=np.float64)
    L.orc_marginal(tree.ref(), C.byref(m), _ptr(allowed), _ptr(bu), _ptr(bu_sf), _ptr(td), _ptr(td_sf),
                   _ptr(lh), _ptr(lh_sf), _ptr(probs))
    if altered is not None:
        L.orc_unalter_allowed(C.c_int(n), _ptr(altered), _ptr(allowed), _ptr(initial))
    L.orc_choose_map(C.c_int(n), C.c_int(K), _ptr(lh), _ptr(initial), _ptr(has_initial), _ptr(allowed))
    map_mask = allowed.copy()
    k_hist = np.zeros(K + 1, dtype=np.int64)
    L.orc_choose_mppa(C.c_int(n), C.c_int(K), _ptr(lh), _ptr(initial), _ptr(has_initial), _ptr(joint_state),
                      C.c_int(1 if force_joint else 0), _ptr(allowed), _ptr(k_hist), None)
    return loglh, probs, map_mask, allowed


def parsimony_fixed(flat, K, annot_mask):
    """parsimonious_acr(..., prediction_method=MP) (``/root/reference/pastml/parsimony.py:245-331``): returns the
    ACCTRAN, DOWNPASS and DELTRAN state masks and step counts."""
    tree = OracleTree(flat)
    n = flat.n_nodes
    annot_mask = _u64(annot_mask)
    L = lib()
    down = np.zeros(n, dtype=np.uint64)
    preset = np.zeros(n, dtype=np.uint64)
    L.orc_pars_uppass(tree.ref(), C.c_int(K), _ptr(annot_mask), _ptr(down), _ptr(preset))
    out = {'uppass_mask': down.copy()}
    acc = np.zeros(n, dtype=np.uint64)
    L.orc_pars_acctran(tree.ref(), _ptr(down), _ptr(acc))
    out['ACCTRAN_mask'] = acc
    out['ACCTRAN_steps'] = int(L.orc_pars_steps(tree.ref(), C.c_int(K), _ptr(acc)))
    L.orc_pars_downpass(tree.ref(), C.c_int(K), _ptr(down), _ptr(preset))
    out['DOWNPASS_mask'] = preset.copy()
    out['DOWNPASS_steps'] = int(L.orc_pars_steps(tree.ref(), C.c_int(K), _ptr(preset)))
    L.orc_pars_deltran(tree.ref(), _ptr(preset))
    out['DELTRAN_mask'] = preset.copy()
    out['DELTRAN_steps'] = int(L.orc_pars_steps(tree.ref(), C.c_int(K), _ptr(preset)))
    return out

"""Maximum-likelihood ancestral character reconstruction on the CUDA engine.

Host-side mirror of ``pastml.ml`` (``/root/reference/pastml/ml.py``): the same entry points (``ml_acr``,
``optimise_likelihood``, ``get_bottom_up_loglikelihood``), result-dictionary keys and node features, but every
likelihood pass is one device-resident launch sequence over *all* the columns that share a model family and a
number of states (a :class:`ColumnBatch`), instead of per-node numpy code per column.

What stays on the host, exactly as in the reference: the two-stage parameter fit with
``scipy.optimize.minimize(method='L-BFGS-B', bounds=...)``, its starting points, restarts and acceptance rule
(``ml.py:174-237, 865-920``).  Columns of a batch are fitted concurrently -- one optimiser thread per column whose
objective evaluations meet in a :class:`LockstepEvaluator`, so that one launch sequence serves one evaluation of
every column still being fitted.
"""
import logging
import threading

import numpy as np
import pandas as pd
from scipy.optimize import minimize

from pastml_b200 import get_personalized_feature_name, CHARACTER, METHOD, NUM_SCENARIOS, NUM_UNRESOLVED_NODES, \
    NUM_STATES_PER_NODE, PERC_UNRESOLVED, STATES
from pastml_b200.engine import Problem, ZeroLikelihoodError, MASK_ANNOTATION, MASK_JOINT, MASK_MAP, MASK_MPPA, \
    MASK_CUSTOM
from pastml_b200.models import ModelWithFrequencies
from pastml_b200.tree import flatten_forest

LOG_LIKELIHOOD = 'log_likelihood'
RESTRICTED_LOG_LIKELIHOOD_FORMAT_STR = '{}_restricted_{{}}'.format(LOG_LIKELIHOOD)

JOINT = 'JOINT'
MPPA = 'MPPA'
MAP = 'MAP'
ALL = 'ALL'
ML = 'ML'

MARGINAL_PROBABILITIES = 'marginal_probabilities'
MODEL = 'model'

MARGINAL_ML_METHODS = {MPPA, MAP}
ML_METHODS = MARGINAL_ML_METHODS | {JOINT}
META_ML_METHODS = {ML, ALL}

BU_LH = 'BOTTOM_UP_LIKELIHOOD'
TD_LH = 'TOP_DOWN_LIKELIHOOD'
LH = 'LIKELIHOOD'
LH_SF = 'LIKELIHOOD_SF'
BU_LH_SF = 'BOTTOM_UP_LIKELIHOOD_SF'
BU_LH_JOINT_STATES = 'BOTTOM_UP_LIKELIHOOD_JOINT_STATES'
TD_LH_SF = 'TOP_DOWM_LIKELIHOOD_SF'
ALLOWED_STATES = 'ALLOWED_STATES'
JOINT_STATE = 'JOINT_STATE'


def is_marginal(method):
    return method in MARGINAL_ML_METHODS or method in META_ML_METHODS


def is_ml(method):
    return method in ML_METHODS or method in META_ML_METHODS


def is_meta_ml(method):
    return method in META_ML_METHODS


def get_default_ml_method():
    return MPPA


class PastMLLikelihoodError(Exception):

    def __init__(self, *args):
        self.message = args[0] if args else None

    def __str__(self):
        if self.message:
            return 'PastMLLikelihoodError, {}'.format(self.message)
        return 'PastMLLikelihoodError has been raised.'


# ---------------------------------------------------------------------------------------- annotation masks

def state_masks(flat, character, states):
    """(annot_mask uint64 [N], has_state uint8 [N]) of one character from the node features, the inputs of
    ``initialize_allowed_states`` / ``get_zero_clusters_with_states`` (``ml.py:293-349``)."""
    s2i = {s: i for i, s in enumerate(states)}
    mask = np.zeros(flat.n_nodes, dtype=np.uint64)
    has = np.zeros(flat.n_nodes, dtype=np.uint8)
    for i, node in enumerate(flat.nodes):
        value = getattr(node, character, None)
        if value is None or (isinstance(value, str) and value == ''):
            continue
        has[i] = 1
        m = 0
        for s in value:
            m |= 1 << s2i[s]
        mask[i] = m
    return mask, has


def mask_to_sets(mask_column, states):
    """uint64 masks -> list of fresh python sets of state names (what pastml stores in ``node.<character>``)."""
    cache = {}
    out = []
    for m in mask_column.tolist():
        proto = cache.get(m)
        if proto is None:
            proto = cache[m] = frozenset(states[k] for k in range(len(states)) if (m >> k) & 1)
        out.append(set(proto))
    return out


# ---------------------------------------------------------------------------------------- lock-step objective

def batched_gradients():
    """L-BFGS-B differentiates the objective by forward differences (scipy's default, as in the reference).  scipy
    hands the n_params perturbed points of one gradient to a map-like ``workers`` option; we evaluate them as
    parameter sets of ONE device pass.  The points, their order and the values are those of the default path
    (checked in tests), only the number of launch sequences changes.  PASTML_B200_BATCHED_FD=0 disables it."""
    import os
    return os.environ.get('PASTML_B200_BATCHED_FD', '1') != '0'


def padded_sets(n_sets, n_columns):
    """Number of parameter sets a batched pass is run with: the engine's TMA-staged bottom-up sweep takes lanes
    (sets x columns) in tiles of 8 (or 4), so a multi-set pass is padded with copies of its last set when that costs
    at most a quarter more lanes.  Single-set passes are left alone."""
    if n_sets <= 1:
        return n_sets
    for tile in (8, 4):
        n = n_sets
        while (n * n_columns) % tile:
            n += 1
        if 4 * n <= 5 * n_sets:
            return n
    return n_sets


def lane_list_size(n):
    """Lanes a lane-list pass is padded to (copies of its last lane): 4, 8, then multiples of 8; from 64 lanes on
    multiples of ``PASTML_B200_LANE_PAD`` (default 32: measured 1.20 s against 1.27 s for 8) -- the engine tiles a pass by the widest of 32 / 16 / 8 / 4 lanes
    that divides it."""
    import os
    if n <= 4:
        return 4
    pad = int(os.environ.get('PASTML_B200_LANE_PAD', '32')) if n >= 64 else 8
    return pad * ((n + pad - 1) // pad)


class LockstepEvaluator(object):
    """One batched device pass for the requests of all the columns that are still being fitted.  A request is a list
    of parameter blocks (one for a plain evaluation, the point and its n_params forward differences for one L-BFGS-B
    step); a pass runs over exactly the requested lanes (lane lists) or over max(len) parameter sets x all columns.
    The coroutine driver (``_fit_batch_coroutines``) fills ``pending`` and calls ``_flush`` itself; for the threads
    driver this is also the rendezvous (``evaluate`` blocks until every active column has asked)."""

    def __init__(self, batch):
        self.batch = batch
        self.cond = threading.Condition()
        self.active = set(range(batch.C))
        self.pending = {}
        self.results = {}
        self.errors = {}
        self.last_block = [m.engine_block() for m in batch.models]
        self.n_passes = 0
        self.n_lane_evaluations = 0
        self.device_ms = 0.   # device time of the objective passes (the rest of the fit's wall time is the host's)

    def _flush(self):
        cols = sorted(self.pending)
        try:
            n_lanes = lane_list_size(sum(len(self.pending[c]) for c in cols))
            if self.batch.lane_lists and n_lanes <= self.batch.max_sets * self.batch.C:
                self._flush_lanes(cols)
            else:
                self._flush_grid(cols)
        except BaseException as e:
            # the device pass itself failed (engine error, out of memory, a zero likelihood the batch did not
            # tolerate): every column that waits for this pass gets the error instead of waiting forever
            for c in cols:
                self.errors[c] = e
                self.results[c] = None
        finally:
            stats = getattr(getattr(self.batch, 'problem', None), 'last_call_stats', None)
            if stats is not None:
                try:
                    self.device_ms += stats()[0]
                except Exception:   # measurement only
                    pass
            self.pending.clear()
            self.cond.notify_all()

    def _flush_grid(self, cols):
        """One pass over (parameter sets) x (all columns): columns without a request ride along on their last block."""
        n_sets = min(padded_sets(max(len(self.pending[c]) for c in cols), self.batch.C), self.batch.max_sets)
        params = np.empty((n_sets, self.batch.C, self.batch.problem.stride), dtype=np.float64)
        for c in range(self.batch.C):
            blocks = self.pending.get(c)
            if blocks is None:
                params[:, c, :] = self.last_block[c]
            else:
                for b in range(n_sets):
                    params[b, c, :] = blocks[min(b, len(blocks) - 1)]
                self.last_block[c] = blocks[-1]
        values, errs = self.batch.loglik_sets(params, n_sets)
        self.n_passes += 1
        for c in cols:
            n = len(self.pending[c])
            self.n_lane_evaluations += n
            self.results[c] = [float(values[b, c]) for b in range(n)]
            bad = [errs[(b, c)] for b in range(n) if (b, c) in errs]
            if bad:
                self.errors[c] = bad[0]

    def _flush_lanes(self, cols):
        """One pass over exactly the requested (column, parameter block) lanes."""
        lane_cols, blocks = [], []
        for c in cols:
            for block in self.pending[c]:
                lane_cols.append(c)
                blocks.append(block)
            self.last_block[c] = self.pending[c][-1]
        n = len(lane_cols)
        padded = lane_list_size(n)
        lane_cols += [lane_cols[-1]] * (padded - n)
        blocks += [blocks[-1]] * (padded - n)
        values, errs = self.batch.loglik_lanes(np.asarray(blocks, dtype=np.float64), lane_cols)
        self.n_passes += 1
        at = 0
        for c in cols:
            k = len(self.pending[c])
            self.n_lane_evaluations += k
            self.results[c] = [float(values[at + b]) for b in range(k)]
            bad = [errs[at + b] for b in range(k) if at + b in errs]
            if bad:
                self.errors[c] = bad[0]
            at += k

    def evaluate(self, c, blocks):
        """Log-likelihoods of column c for the given parameter blocks."""
        with self.cond:
            self.pending[c] = list(blocks)
            self.results.pop(c, None)
            if set(self.pending) >= self.active:
                self._flush()
            else:
                while c not in self.results:
                    self.cond.wait()
            if c in self.errors:
                raise self.errors.pop(c)
            return self.results[c]

    def retire(self, c):
        """Column c takes no further part (fitted, or failed); never raises (it runs in a ``finally``)."""
        with self.cond:
            self.active.discard(c)
            self.pending.pop(c, None)
            if self.pending and set(self.pending) >= self.active:
                self._flush()   # delivers failures through self.errors


class ColumnBatch(object):
    """Columns of one forest that share the model family and the number of states: one engine problem."""

    def __init__(self, flat, characters, models, device=0, annot_mask=None, has_state=None):
        """``annot_mask`` / ``has_state`` ([C][N]) may be given directly (array-native callers); otherwise they are
        read from the node features named by ``characters``."""
        self.flat = flat
        self.characters = list(characters)
        self.models = list(models)
        self.C = len(self.characters)
        m0 = self.models[0]
        self.K = len(m0.states)
        self.kind = m0.engine_kind
        if annot_mask is None:
            masks, has = zip(*(state_masks(flat, ch, list(m.states)) for ch, m in zip(self.characters, self.models)))
            annot_mask, has_state = np.stack(masks), np.stack(has)
        self.annot_mask = np.ascontiguousarray(annot_mask, dtype=np.uint64)
        self.has_state = np.ascontiguousarray(has_state, dtype=np.uint8)
        # capacity for the finite-difference batch of one gradient: every free parameter + the base point
        self.max_sets = padded_sets(1 + max(m.get_num_params() for m in self.models), self.C) if batched_gradients() else 1
        self.batched_fd = self.max_sets > 1
        # room for the largest lane-list pass: every column asking for a full gradient, padded to whole lane tiles
        self.max_sets = max(self.max_sets, -(-lane_list_size(self.max_sets * self.C) // self.C))
        self.problem = Problem(flat, self.K, self.kind, self.annot_mask, self.has_state,
                               max_param_sets=self.max_sets, device=device)
        # every objective pass of the fit runs over exactly the requested (column, parameter block) lanes (engine:
        # pml_loglik_lanes): columns that have finished no longer cost anything, and the base point and the perturbed
        # points of a finite difference go through the same kernel, lane by lane.  PASTML_B200_LANE_LISTS=0 keeps the
        # (sets x all columns) grid
        import os
        self.lane_lists = os.environ.get('PASTML_B200_LANE_LISTS', '1') != '0' and self.problem.supports_lane_lists()

    def close(self):
        self.problem.close()

    def params(self):
        return np.stack([m.engine_block() for m in self.models])

    def _error_for(self, c, parent, child):
        names = self.flat.names
        return PastMLLikelihoodError(
            'The parent node {} and its child node {} have non-intersecting states, and are connected by a '
            'zero-length ({:g}) branch. This creates a zero likelihood value. To avoid this issue check the '
            'restrictions on these node states and/or use a smoothing factor (tau).'
            .format(names[parent], names[child], self.flat.dist[child]))

    def loglik_sets(self, params, n_sets):
        """(values [n_sets][C], {(set, column): PastMLLikelihoodError}) for n_sets parameter sets x all columns."""
        values, bad = self.problem.loglik(params, n_sets, MASK_ANNOTATION, True, True, tolerate_zero=True)
        errs = {(j // self.C, j % self.C): self._error_for(j % self.C, *pc) for j, pc in bad.items()}
        return values, errs

    def loglik_lanes(self, params, lane_columns):
        """(values [n_lanes], {lane: PastMLLikelihoodError}) for an explicit list of (column, parameter block) lanes."""
        values, bad = self.problem.loglik_lanes(params, lane_columns)
        errs = {j: self._error_for(lane_columns[j], *pc) for j, pc in bad.items()}
        return values, errs

    def loglik_columns(self, cols=None, mask_kind=MASK_ANNOTATION, is_marginal=True, alter=True):
        """({column: log-likelihood}, {column: PastMLLikelihoodError}) for one batched pass."""
        values, bad = self.problem.loglik(self.params(), 1, mask_kind, is_marginal, alter, tolerate_zero=True)
        cols = range(self.C) if cols is None else cols
        errs = {c: self._error_for(c, *bad[c]) for c in cols if c in bad}
        return {c: float(values[0, c]) for c in cols}, errs


# ---------------------------------------------------------------------------------------- parameter fit

def optimize_likelihood_params(evaluate, observed_frequencies, model, evaluate_many=None, rng=None):
    """One stage of the fit for one column (``ml.py:174-237``): L-BFGS-B from the current parameters, then from the
    observed frequencies, then from uniform random points, returning at the first successful run that is not worse
    than the better starting point.  ``evaluate()`` returns the log-likelihood at the model's current parameters;
    ``evaluate_many(blocks)`` (optional) the log-likelihoods for several parameter blocks in one device pass.
    ``rng``: source of the random restart points (default: numpy's global state, like the reference).  With
    ``evaluate_many`` scipy's finite differences go through ``workers`` and are not counted in ``nfev``: the
    ``maxfun`` limit (15000) then counts base-point evaluations only -- it is never reached on either path for the
    models here (a fit takes a few hundred evaluations)."""
    bounds = model.get_bounds()

    def get_v(ps):
        if np.any(pd.isnull(ps)):
            return np.nan
        model.set_params_from_optimised(ps)
        res = evaluate()
        return np.inf if pd.isnull(res) else -res

    def get_v_many(_fun, points):
        # scipy's finite-difference driver maps the objective over the perturbed points of one gradient: same points,
        # same order as its sequential loop (the model ends up at the last one, as it would there)
        points = list(points)
        ok = (~np.isnan(np.asarray(points, dtype=np.float64)).any(axis=1)).tolist() if points else []
        blocks = model.engine_blocks([ps for ps, good in zip(points, ok) if good])
        values = iter(evaluate_many(blocks) if blocks else [])
        out = []
        for good in ok:
            if not good:
                out.append(np.nan)
            else:
                res = next(values)
                out.append(np.inf if pd.isnull(res) else -res)
        return out

    options = {'workers': get_v_many} if evaluate_many is not None else {}

    if np.any(observed_frequencies <= 0):
        observed_frequencies = np.maximum(observed_frequencies, 1e-10)

    x0_JC = model.get_optimised_parameters()
    optimise_frequencies = isinstance(model, ModelWithFrequencies) and model._optimise_frequencies
    x0_EFT = x0_JC
    if optimise_frequencies:
        model.frequencies = observed_frequencies
        x0_EFT = model.get_optimised_parameters()
    log_lh_JC = -get_v(x0_JC)
    log_lh_EFT = log_lh_JC if not optimise_frequencies else -get_v(x0_EFT)
    best_log_lh = max(log_lh_JC, log_lh_EFT)

    for i in range(100):
        if i == 0:
            vs = x0_JC
        elif optimise_frequencies and i == 1:
            vs = x0_EFT
        else:
            vs = (rng or np.random).uniform(bounds[:, 0], bounds[:, 1])
        fres = minimize(get_v, x0=vs, method='L-BFGS-B', bounds=bounds, options=options)
        if fres.success and not np.any(np.isnan(fres.x)):
            if -fres.fun >= best_log_lh:
                model.set_params_from_optimised(fres.x)
                return -fres.fun
    model.set_params_from_optimised(x0_JC if log_lh_JC >= log_lh_EFT else x0_EFT)
    return best_log_lh


def _fit_one(evaluate, character, model, observed_frequencies, evaluate_many=None, rng=None):
    """optimise_likelihood for one column (``ml.py:865-920``) given its objective."""
    logger = logging.getLogger('pastml')
    likelihood = evaluate()
    if np.isnan(likelihood):
        raise PastMLLikelihoodError('Failed to calculate the likelihood for your tree, please check that you do not '
                                    'have contradicting {} states specified for internal tree nodes, and if not - '
                                    'submit a bug at https://github.com/evolbioinfo/pastml/issues'.format(character))
    if not model.get_num_params():
        logger.debug('All the parameters are fixed for {}:\n{}\tlog likelihood:\t{:.6f}.'
                     .format(character, model._print_parameters(), likelihood))
        return likelihood
    logger.debug('Initial values for {} parameter optimisation:\n{}\tlog likelihood:\t{:.6f}.'
                 .format(character, model._print_parameters(), likelihood))
    failure = 'Failed to optimise the likelihood for your tree, please check that you do not have contradicting {} ' \
              'states specified for internal tree nodes, and if not - submit a bug at ' \
              'https://github.com/evolbioinfo/pastml/issues'.format(character)
    if not model.basic_params_fixed():
        model.fix_extra_params()
        likelihood = optimize_likelihood_params(evaluate, observed_frequencies, model, evaluate_many, rng)
        if np.isnan(likelihood) or likelihood == -np.inf:
            raise PastMLLikelihoodError(failure)
        model.unfix_extra_params()
        if not model.extra_params_fixed():
            logger.debug('Pre-optimised basic parameters for {}:\n{}\tlog likelihood:\t{:.6f}.'
                         .format(character, model._print_basic_parameters(), likelihood))
    if not model.extra_params_fixed():
        likelihood = optimize_likelihood_params(evaluate, observed_frequencies, model, evaluate_many, rng)
        if np.isnan(likelihood) or likelihood == -np.inf:
            raise PastMLLikelihoodError(failure)
    logger.debug('Optimised parameters for {}:\n{}\tlog likelihood:\t{:.6f}'
                 .format(character, model._print_parameters(), likelihood))
    return likelihood


# ---- the same fit as coroutines: every column's optimiser in ONE thread, one device pass per round of requests

def _values_gen(model, points):
    """Objective values (-log-likelihood; inf where the likelihood is undefined, nan for a nan point) at the given
    optimiser vectors: yields the parameter blocks, is sent their log-likelihoods (what ``get_v`` does point by point)."""
    points = list(points)
    ok = (~np.isnan(np.asarray(points, dtype=np.float64)).any(axis=1)).tolist() if points else []
    blocks = model.engine_blocks([ps for ps, good in zip(points, ok) if good])
    results = iter((yield blocks) if blocks else [])
    out = []
    for good in ok:
        if not good:
            out.append(np.nan)
        else:
            res = next(results)
            out.append(np.inf if pd.isnull(res) else -res)
    return out


def _minimize_gen(model, x0, bounds):
    """One bounded L-BFGS-B run from x0 (:mod:`pastml_b200.lbfgsb`: scipy's iterates bit for bit)."""
    from pastml_b200.lbfgsb import minimize_lbfgsb
    run = minimize_lbfgsb(x0, bounds)
    try:
        request = next(run)
        while True:
            values = yield from _values_gen(model, request)
            request = run.send(values)
    except StopIteration as stop:
        return stop.value


def _optimize_stage_gen(model, observed_frequencies, rng):
    """:func:`optimize_likelihood_params` as a coroutine (same starting points, same acceptance rule)."""
    bounds = model.get_bounds()
    if np.any(observed_frequencies <= 0):
        observed_frequencies = np.maximum(observed_frequencies, 1e-10)
    x0_JC = model.get_optimised_parameters()
    optimise_frequencies = isinstance(model, ModelWithFrequencies) and model._optimise_frequencies
    x0_EFT = x0_JC
    if optimise_frequencies:
        model.frequencies = observed_frequencies
        x0_EFT = model.get_optimised_parameters()
    starts = yield from _values_gen(model, [x0_JC, x0_EFT] if optimise_frequencies else [x0_JC])
    log_lh_JC = -starts[0]
    log_lh_EFT = -starts[1] if optimise_frequencies else log_lh_JC
    best_log_lh = max(log_lh_JC, log_lh_EFT)
    for i in range(100):
        if i == 0:
            vs = x0_JC
        elif optimise_frequencies and i == 1:
            vs = x0_EFT
        else:
            vs = (rng or np.random).uniform(bounds[:, 0], bounds[:, 1])
        fres = yield from _minimize_gen(model, vs, bounds)
        if fres.success and not np.any(np.isnan(fres.x)):
            if -fres.fun >= best_log_lh:
                model.set_params_from_optimised(fres.x)
                return -fres.fun
    model.set_params_from_optimised(x0_JC if log_lh_JC >= log_lh_EFT else x0_EFT)
    return best_log_lh


def _fit_one_gen(character, model, observed_frequencies, rng):
    """:func:`_fit_one` as a coroutine: yields lists of parameter blocks, is sent their log-likelihoods, returns the
    fitted log-likelihood."""
    logger = logging.getLogger('pastml')
    likelihood = (yield [model.engine_block()])[0]
    if np.isnan(likelihood):
        raise PastMLLikelihoodError('Failed to calculate the likelihood for your tree, please check that you do not '
                                    'have contradicting {} states specified for internal tree nodes, and if not - '
                                    'submit a bug at https://github.com/evolbioinfo/pastml/issues'.format(character))
    if not model.get_num_params():
        logger.debug('All the parameters are fixed for {}:\n{}\tlog likelihood:\t{:.6f}.'
                     .format(character, model._print_parameters(), likelihood))
        return likelihood
    logger.debug('Initial values for {} parameter optimisation:\n{}\tlog likelihood:\t{:.6f}.'
                 .format(character, model._print_parameters(), likelihood))
    failure = 'Failed to optimise the likelihood for your tree, please check that you do not have contradicting {} ' \
              'states specified for internal tree nodes, and if not - submit a bug at ' \
              'https://github.com/evolbioinfo/pastml/issues'.format(character)
    if not model.basic_params_fixed():
        model.fix_extra_params()
        likelihood = yield from _optimize_stage_gen(model, observed_frequencies, rng)
        if np.isnan(likelihood) or likelihood == -np.inf:
            raise PastMLLikelihoodError(failure)
        model.unfix_extra_params()
        if not model.extra_params_fixed():
            logger.debug('Pre-optimised basic parameters for {}:\n{}\tlog likelihood:\t{:.6f}.'
                         .format(character, model._print_basic_parameters(), likelihood))
    if not model.extra_params_fixed():
        likelihood = yield from _optimize_stage_gen(model, observed_frequencies, rng)
        if np.isnan(likelihood) or likelihood == -np.inf:
            raise PastMLLikelihoodError(failure)
    logger.debug('Optimised parameters for {}:\n{}\tlog likelihood:\t{:.6f}'
                 .format(character, model._print_parameters(), likelihood))
    return likelihood


def _fit_batch_coroutines(batch, observed_frequencies):
    """All the columns' fits advanced in lock step by ONE thread: a round collects every running column's request,
    evaluates them in one device pass (:class:`LockstepEvaluator` does the packing) and sends the values back."""
    ev = LockstepEvaluator(batch)
    out = [None] * batch.C
    failures = [None] * batch.C
    rngs = [None] if batch.C == 1 else [np.random.default_rng(int(seed)) for seed in np.random.randint(0, 2 ** 31 - 1, size=batch.C)]
    running, requests = {}, {}
    for c in range(batch.C):
        gen = _fit_one_gen(batch.characters[c], batch.models[c], observed_frequencies[c], rngs[c])
        try:
            requests[c] = next(gen)
            running[c] = gen
        except StopIteration as stop:
            out[c] = stop.value
        except BaseException as e:
            failures[c] = e
    while running:
        with ev.cond:
            ev.pending = {c: list(requests[c]) for c in running}
            ev.results.clear()
            ev._flush()
        for c in list(running):
            gen = running[c]
            try:
                if c in ev.errors:
                    requests[c] = gen.throw(ev.errors.pop(c))
                else:
                    requests[c] = gen.send(ev.results[c])
            except StopIteration as stop:
                out[c] = stop.value
                del running[c]
            except BaseException as e:
                failures[c] = e
                del running[c]
    batch.last_fit_device_ms = ev.device_ms
    for e in failures:
        if e is not None:
            raise e
    return out, ev.n_passes


def fit_batch(batch, observed_frequencies):
    """Fits all the columns of a batch concurrently; returns ([log-likelihood per column], n device passes).

    Default driver: the coroutine one above.  PASTML_B200_FIT_DRIVER=threads takes the original one (one scipy
    ``minimize`` call per column in its own thread, meeting in :class:`LockstepEvaluator`), which is also what columns
    with frequency smoothing use: their model is stateful across evaluations (``models/__init__.py:331-335``), so the
    order of the evaluations is part of the result."""
    import os
    stateful = any(getattr(m, '_frequency_smoothing', False) for m in batch.models)
    if os.environ.get('PASTML_B200_FIT_DRIVER', 'coroutines') != 'threads' and batch.batched_fd and not stateful:
        try:
            import pastml_b200.lbfgsb  # noqa: F401  (binds scipy's reverse-communication routine: private names)
        except ImportError:   # another scipy layout: the threads driver goes through scipy's public minimize()
            logging.getLogger('pastml').debug('coroutine L-BFGS-B driver unavailable with this scipy, using threads')
        else:
            return _fit_batch_coroutines(batch, observed_frequencies)
    ev = LockstepEvaluator(batch)
    out = [None] * batch.C
    failures = [None] * batch.C
    # restart points: one generator per column, seeded from numpy's global state before the threads start, so that a
    # run is reproducible under np.random.seed whatever the thread interleaving (a single column draws from the
    # global state itself, like the reference)
    rngs = [None] if batch.C == 1 else [np.random.default_rng(int(seed)) for seed in np.random.randint(0, 2 ** 31 - 1, size=batch.C)]

    def work(c):
        try:
            model = batch.models[c]
            many = (lambda blocks: ev.evaluate(c, blocks)) if batch.batched_fd else None
            out[c] = _fit_one(lambda: ev.evaluate(c, [model.engine_block()])[0], batch.characters[c], model,
                              observed_frequencies[c], many, rngs[c])
        except BaseException as e:  # re-raised in the caller's thread
            failures[c] = e
        finally:
            ev.retire(c)

    if batch.C == 1:
        work(0)
    else:
        threads = [threading.Thread(target=work, args=(c,), daemon=True) for c in range(batch.C)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
    batch.last_fit_device_ms = ev.device_ms
    for e in failures:
        if e is not None:
            raise e
    return out, ev.n_passes


def optimise_likelihood(forest, character, model, observed_frequencies):
    """Drop-in for ``pastml.ml.optimise_likelihood``: fits one character's model on the forest."""
    flat = forest if hasattr(forest, 'level_ptr') else flatten_forest(forest)
    batch = ColumnBatch(flat, [character], [model])
    try:
        return fit_batch(batch, [observed_frequencies])[0][0]
    finally:
        batch.close()


def get_bottom_up_loglikelihood(tree, character, model, is_marginal=True, alter=True):
    """Drop-in for ``pastml.ml.get_bottom_up_loglikelihood`` (``ml.py:82-121``): log-likelihood of one tree for one
    character.  The allowed states are the ``<character>_ALLOWED_STATES`` node features when present
    (0/1 arrays, as ``initialize_allowed_states`` writes them), otherwise the annotation sets."""
    flat = flatten_forest(tree)
    batch = ColumnBatch(flat, [character], [model])
    try:
        feature = get_personalized_feature_name(character, ALLOWED_STATES)
        mask_kind = MASK_ANNOTATION
        if all(hasattr(n, feature) for n in flat.nodes):
            custom = np.zeros((1, flat.n_nodes), dtype=np.uint64)
            for i, n in enumerate(flat.nodes):
                bits = 0
                for k, v in enumerate(getattr(n, feature)):
                    if v:
                        bits |= 1 << k
                custom[0, i] = bits
            batch.problem.set_custom_masks(custom)
            mask_kind = MASK_CUSTOM
        values, errs = batch.loglik_columns([0], mask_kind, is_marginal, alter)
        if errs:
            raise errs[0]
        return values[0]
    finally:
        batch.close()


# ---------------------------------------------------------------------------------------- allowed-state bookkeeping
#
# Host-side helpers with the reference's names and side effects on the nodes, for code that prepares or inspects the
# ``<character>_ALLOWED_STATES`` features itself and then calls :func:`get_bottom_up_loglikelihood` (which reads them).
# No likelihood arithmetic happens here; inside :func:`ml_acr` the same steps run on the device (``alter_kernel``).

def _allowed_feature(feature):
    return get_personalized_feature_name(feature, ALLOWED_STATES)


def get_state2allowed_states(states, by_name=True):
    """(all-ones array, {state (or its index): one-hot array}); missing values map to all ones (``ml.py:625-637``)."""
    n = len(states)
    everything = np.ones(n, int)
    table = {}
    for index, state in enumerate(states):
        table[state if by_name else index] = (np.arange(n) == index).astype(int)
    if by_name:
        table[None] = table[''] = everything
    return everything, table


def initialize_allowed_states(tree, feature, states):
    """0/1 array per node from its annotation set: every state where the set is empty or missing (``ml.py:293-318``)."""
    index = {state: i for i, state in enumerate(states)}
    n = len(states)
    for node in tree.traverse():
        annotated = getattr(node, feature, set())
        allowed = np.zeros(n, dtype=int)
        if annotated:
            allowed[[index[state] for state in annotated]] = 1
        else:
            allowed[:] = 1
        node.add_feature(_allowed_feature(feature), allowed)


def get_zero_clusters_with_states(tree, feature):
    """Groups of at least two annotated nodes connected by zero-length branches (``ml.py:321-349``)."""
    def annotated(node):
        value = getattr(node, feature, None)
        return value is not None and value != ''

    tops = [tree]
    while tops:
        members, frontier = [], [tops.pop()]
        while frontier:
            node = frontier.pop()
            if annotated(node):
                members.append(node)
            for child in node.children:
                (frontier if child.dist == 0 else tops).append(child)
        if len(members) > 1:
            yield members


def alter_zero_node_allowed_states(tree, feature):
    """Where the members of a zero-branch cluster share no allowed state, every member gets the running union of the
    members' arrays (one shared array object, as in the reference) and keeps its own in ``.initial``; returns the
    altered nodes (``ml.py:352-387``)."""
    allowed_feature = _allowed_feature(feature)
    altered = []
    for members in get_zero_clusters_with_states(tree, feature):
        votes = np.sum([getattr(node, allowed_feature) for node in members], axis=0)
        if votes.max() == len(members):
            continue
        union = None
        for node in members:
            own = getattr(node, allowed_feature).copy()
            if union is None:
                union = own.copy()
            else:
                union[np.nonzero(own)] = 1
            node.add_feature(allowed_feature, union)
            node.add_feature(allowed_feature + '.initial', own)
            altered.append(node)
    return altered


def unalter_zero_node_allowed_states(altered_nodes, feature):
    """Back to the node's own states: what is left of them in the current array, or all of them (``ml.py:390-405``)."""
    allowed_feature = _allowed_feature(feature)
    for node in altered_nodes:
        own = getattr(node, allowed_feature + '.initial')
        kept = getattr(node, allowed_feature) & own
        node.add_feature(allowed_feature, kept if np.any(kept > 0) else own)


def unalter_zero_node_joint_states(altered_nodes, feature):
    """Joint pointers of altered nodes that point outside the node's own states are redirected to its first own
    state (``ml.py:408-428``)."""
    pointer_feature = get_personalized_feature_name(feature, BU_LH_JOINT_STATES)
    for node in altered_nodes:
        own = getattr(node, _allowed_feature(feature) + '.initial')
        first = np.argmax(own)
        if np.count_nonzero(own > 0) == 1:
            node.add_feature(pointer_feature, np.ones(len(own), int) * first)
            continue
        pointers = getattr(node, pointer_feature)
        for i in range(len(own)):
            if not own[pointers[i]]:
                pointers[i] = first


def convert_allowed_states2feature(tree, feature, states, out_feature=None):
    """``node.<out_feature>`` = the set of state names the node's array allows (``ml.py:923-928``)."""
    states = np.asarray(states)
    for node in tree.traverse():
        node.add_feature(out_feature or feature, set(states[getattr(node, _allowed_feature(feature)).astype(bool)]))


def _parsimonious_states2allowed_states(tree, ps_feature, feature, states):
    """Allowed-state arrays from the state sets a parsimonious reconstruction left in ``ps_feature`` (``ml.py:931-940``)."""
    index = {state: i for i, state in enumerate(states)}
    for node in tree.traverse():
        allowed = np.zeros(len(states), dtype=int)
        allowed[[index[state] for state in getattr(node, ps_feature)]] = 1
        node.add_feature(_allowed_feature(feature), allowed)


# ---------------------------------------------------------------------------------------- reconstruction

STATE_MASKS = 'state_masks'   # array-native result: uint64 [N], bit k = state k, in the forest's level order


def _write_states(flat, mask_column, states, feature):
    for node, value in zip(flat.nodes, mask_to_sets(mask_column, states)):
        node.add_feature(feature, value)


def reconstruct_batch(batch, prediction_method, likelihoods, force_joint=True, write_likelihood_features=True,
                      parsimony_hook=None):
    """Everything ``ml_acr`` does after the fit (``ml.py:659-750``) for all the columns of a batch; returns one list
    of result dictionaries per column."""
    logger = logging.getLogger('pastml')
    flat, pb = batch.flat, batch.problem
    params = batch.params()
    C, N = batch.C, flat.n_nodes
    order = flat.ref_order
    names = flat.names
    index = [names[i] for i in order]
    results = [[] for _ in range(C)]
    current = [{LOG_LIKELIHOOD: likelihoods[c], CHARACTER: batch.characters[c], METHOD: prediction_method,
                MODEL: batch.models[c], STATES: batch.models[c].states} for c in range(C)]

    def emit(method, mask_nc):
        if method != prediction_method and not is_meta_ml(prediction_method):
            return
        for c in range(C):
            character = batch.characters[c]
            feature = get_personalized_feature_name(character, method) if prediction_method != method else character
            res = current[c].copy()
            if flat.nodes is not None:
                _write_states(flat, mask_nc[:, c], batch.models[c].states, feature)
            else:   # array-native forest: the predicted states travel with the result instead of the node objects
                res[STATE_MASKS] = np.ascontiguousarray(mask_nc[:, c]) & np.uint64((1 << batch.K) - 1)
            res[CHARACTER] = feature
            res[METHOD] = method
            results[c].append(res)

    def note_restricted(method, mask_kind):
        values, errs = batch.loglik_columns(None, mask_kind, True, True)
        for c in range(C):
            if c in errs:
                raise errs[c]
            logger.debug('Log likelihood for {} after {} state selection:\t{:.6f}'
                         .format(batch.characters[c], method, values[c]))
            current[c][RESTRICTED_LOG_LIKELIHOOD_FORMAT_STR.format(method)] = values[c]

    have_joint = False
    if prediction_method != MAP:
        lh_joint, joint_state = pb.joint(params, want_states=True)
        have_joint = True
        for c in range(C):
            current[c][RESTRICTED_LOG_LIKELIHOOD_FORMAT_STR.format(JOINT)] = float(lh_joint[c])
        if prediction_method == JOINT or is_meta_ml(prediction_method):
            emit(JOINT, np.uint64(1) << joint_state.astype(np.uint64))

    if is_marginal(prediction_method):
        use_fj = bool(force_joint) and have_joint
        res = pb.marginal(params, force_joint=use_fj, want_probs=True, want_node_lh=write_likelihood_features,
                          want_masks=True)
        probs = res['probs']
        for c in range(C):
            states = batch.models[c].states
            current[c][MARGINAL_PROBABILITIES] = pd.DataFrame(probs[order, c, :], index=index, columns=states)
        if write_likelihood_features and flat.nodes is not None:
            has_init, init = pb.get_initial()
            for c in range(C):
                lh_f = get_personalized_feature_name(batch.characters[c], LH)
                sf_f = get_personalized_feature_name(batch.characters[c], LH_SF)
                K = batch.K
                bits = ((init[:, c][:, None] >> np.arange(K, dtype=np.uint64)) & np.uint64(1)).astype(np.float64)
                for i, node in enumerate(flat.nodes):
                    vec = probs[i, c, :].copy()
                    if has_init[i, c]:
                        vec *= bits[i]
                    node.add_feature(lh_f, vec)
                    node.add_feature(sf_f, -float(res['node_log10lh'][i, c]))
        note_restricted(MAP, MASK_MAP)
        emit(MAP, res['map_mask'])

        if prediction_method == MPPA or is_meta_ml(prediction_method):
            if prediction_method == ALL and parsimony_hook is not None:
                parsimony_hook(batch, current, results, note_restricted)
            mppa_mask, k_hist = pb.mppa_finalize(force_joint=use_fj, want_masks=True)
            for c in range(C):
                num_nodes = batch.models[c].forest_stats.num_nodes
                scenarios, unresolved, n_states = 1, 0, 0
                for k in range(1, batch.K + 1):
                    cnt = int(k_hist[c, k])
                    scenarios *= k ** cnt
                    n_states += k * cnt
                    if k > 1:
                        unresolved += cnt
                current[c][NUM_SCENARIOS] = scenarios
                current[c][NUM_UNRESOLVED_NODES] = unresolved
                current[c][NUM_STATES_PER_NODE] = n_states / num_nodes
                current[c][PERC_UNRESOLVED] = unresolved * 100 / num_nodes
                logger.debug('{} node{} unresolved ({:.2f}%) for {} by {}, i.e. {:.4f} state{} per node in average.'
                             .format(unresolved, 's are' if unresolved != 1 else ' is', current[c][PERC_UNRESOLVED],
                                     batch.characters[c], MPPA, current[c][NUM_STATES_PER_NODE],
                                     's' if current[c][NUM_STATES_PER_NODE] > 1 else ''))
            note_restricted(MPPA, MASK_MPPA)
            emit(MPPA, mppa_mask)
    return results


def ml_acr(forest, character, prediction_method, model, observed_frequencies, force_joint=True):
    """Drop-in for ``pastml.ml.ml_acr`` (``ml.py:640-750``): fits the model of one character and reconstructs its
    ancestral states; returns the list of result dictionaries (one per produced method)."""
    return ml_acr_batch(forest, [character], prediction_method, [model], [observed_frequencies], force_joint)[0]


def ml_acr_batch(forest, characters, prediction_method, models, observed_frequencies, force_joint=True,
                 flat=None, device=0, write_likelihood_features=True):
    """``ml_acr`` for several characters that share the model family and the number of states, in one engine
    problem.  Returns one list of result dictionaries per character."""
    if flat is None:
        flat = flatten_forest(forest)
    batch = ColumnBatch(flat, characters, models, device=device)
    try:
        likelihoods, _ = fit_batch(batch, observed_frequencies)
        hook = None
        if prediction_method == ALL:
            from pastml_b200.parsimony import all_methods_hook
            hook = all_methods_hook
        return reconstruct_batch(batch, prediction_method, likelihoods, force_joint,
                                 write_likelihood_features=write_likelihood_features, parsimony_hook=hook)
    finally:
        batch.close()


# ---------------------------------------------------------------------------------------- transition counts

def marginal_counts(forest, character, model, n_repetitions=1_000):
    """Expected numbers of state transitions along the forest, by sampling ancestral scenarios from the marginal
    posterior (drop-in for ``pastml.ml.marginal_counts``, ``/root/reference/pastml/ml.py:753-862``): the root state is
    drawn ``n_repetitions`` times from its marginal probabilities, then every branch is walked in level order and the
    child's state is drawn given the parent's state and the subtree below the child
    (``P(i | j) ~ L_child[i] P_ij(t) pi_i allowed_child[i]``).  Returns the K x K matrix of counts per repetition, the
    diagonal holding the number of lineages that stay in a state.

    Everything runs on the device (``pml_marginal_counts``): the alteration of zero-branch clusters, the bottom-up
    sweep, and one sampling kernel per level in which every repetition carries its own state down the tree --  the same
    distribution of counts as the reference's per-state multinomial draws, on a different random stream (Philox,
    counter-based).  The stream's seed is taken from numpy's global random state, so ``np.random.seed`` controls the
    result like it does for the reference."""
    if not isinstance(forest, (list, tuple)):
        forest = [forest]
    states = list(model.states)
    flat = flatten_forest(forest)
    mask, has = state_masks(flat, character, states)
    seed = int(np.random.randint(0, 2 ** 62))
    pb = Problem(flat, len(states), model.engine_kind, mask[None, :], has[None, :])
    try:
        return pb.marginal_counts(model.engine_block(), n_repetitions, seed)[0]
    finally:
        pb.close()

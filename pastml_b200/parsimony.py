"""Maximum-parsimony ancestral character reconstruction on the CUDA engine.

Host-side mirror of ``pastml.parsimony`` (``/root/reference/pastml/parsimony.py``): ``parsimonious_acr`` with the
ACCTRAN, DOWNPASS, DELTRAN and MP (all three) methods.  State sets are bit masks on the device; the uppass, the
three top-down passes and the step count are level-synchronous kernels over all columns of a batch.
"""
import logging

import numpy as np

from pastml_b200 import get_personalized_feature_name, METHOD, STATES, CHARACTER, NUM_SCENARIOS, \
    NUM_UNRESOLVED_NODES, NUM_NODES, NUM_TIPS, NUM_STATES_PER_NODE, PERC_UNRESOLVED
from pastml_b200.engine import Problem, MASK_CUSTOM, PARS_ACCTRAN, PARS_DOWNPASS, PARS_DELTRAN
from pastml_b200.tree import flatten_forest

STEPS = 'steps'

DOWNPASS = 'DOWNPASS'
ACCTRAN = 'ACCTRAN'
DELTRAN = 'DELTRAN'
MP = 'MP'

MP_METHODS = {DOWNPASS, ACCTRAN, DELTRAN}
META_MP_METHODS = {MP}

_FLAGS = {ACCTRAN: PARS_ACCTRAN, DOWNPASS: PARS_DOWNPASS, DELTRAN: PARS_DELTRAN}


def is_meta_mp(method):
    return method in META_MP_METHODS


def get_default_mp_method():
    return DOWNPASS


def is_parsimonious(method):
    return method in MP_METHODS | {MP}


def _popcount(masks):
    """Set bits per uint64 (np.bitwise_count needs numpy >= 2: older numpy takes the byte-table route)."""
    masks = np.ascontiguousarray(masks, dtype=np.uint64)
    if hasattr(np, 'bitwise_count'):
        return np.bitwise_count(masks).astype(np.int64)
    table = np.array([bin(b).count('1') for b in range(256)], dtype=np.int64)
    return table[masks.view(np.uint8).reshape(masks.shape + (8,))].sum(axis=-1)


def _set_stats(mask_column, num_nodes):
    """(num_scenarios, num_unresolved_nodes, states per node, % unresolved) of one column of state masks
    (``choose_parsimonious_states``, ``parsimony.py:334-357``)."""
    sizes = _popcount(mask_column)
    hist = np.bincount(sizes)
    scenarios = 1
    for k, cnt in enumerate(hist.tolist()):
        if cnt:
            scenarios *= k ** cnt
    unresolved = int((sizes > 1).sum())
    return scenarios, unresolved, int(sizes.sum()) / num_nodes, unresolved * 100 / num_nodes


def parsimony_on_problem(problem, flat, characters, states_per_column, prediction_method, num_nodes, num_tips,
                         write_features=True):
    """Runs the requested parsimony method(s) for all the columns of an engine problem; returns one list of result
    dictionaries per column, ordered like the reference's (ACCTRAN, DOWNPASS, DELTRAN)."""
    from pastml_b200.ml import mask_to_sets
    logger = logging.getLogger('pastml')
    wanted = [m for m in (ACCTRAN, DOWNPASS, DELTRAN) if prediction_method in (m, MP)]
    flags = 0
    for m in wanted:
        flags |= _FLAGS[m]
    out = problem.parsimony(flags, want_masks=True)
    results = [[] for _ in characters]
    for m in wanted:
        masks, steps = out[m]
        for c, character in enumerate(characters):
            feature = get_personalized_feature_name(character, m) if prediction_method != m else character
            if write_features and flat.nodes is not None:
                for node, value in zip(flat.nodes, mask_to_sets(masks[:, c], states_per_column[c])):
                    node.add_feature(feature, value)
            scenarios, unresolved, per_node, perc = _set_stats(masks[:, c], num_nodes)
            logger.debug('{} node{} unresolved ({:.2f}%) for {} by {}, i.e. {:.4f} state{} per node in average.'
                         .format(unresolved, 's are' if unresolved != 1 else ' is', perc, character, m, per_node,
                                 's' if per_node > 1 else ''))
            results[c].append({STATES: states_per_column[c], NUM_NODES: num_nodes, NUM_TIPS: num_tips,
                               STEPS: int(steps[c]), NUM_SCENARIOS: scenarios, NUM_UNRESOLVED_NODES: unresolved,
                               NUM_STATES_PER_NODE: per_node, PERC_UNRESOLVED: perc, CHARACTER: feature, METHOD: m})
            if flat.nodes is None:   # array-native forest: the predicted states travel with the result
                from pastml_b200.ml import STATE_MASKS
                results[c][-1][STATE_MASKS] = np.ascontiguousarray(masks[:, c]) & \
                    np.uint64((1 << len(states_per_column[c])) - 1)
    for c, character in enumerate(characters):
        if results[c]:
            logger.debug('Parsimonious reconstruction for {} requires {} state changes.'
                         .format(character, results[c][-1][STEPS]))
    return results, out


def parsimonious_acr_batch(forest, characters, prediction_method, states_per_column, num_nodes, num_tips, flat=None,
                           device=0):
    """``parsimonious_acr`` for several characters with the same number of states in one engine problem."""
    from pastml_b200.ml import state_masks
    if flat is None:
        flat = flatten_forest(forest)
    masks, has = zip(*(state_masks(flat, ch, list(st)) for ch, st in zip(characters, states_per_column)))
    problem = Problem(flat, len(states_per_column[0]), 0, np.stack(masks), np.stack(has), device=device)
    try:
        return parsimony_on_problem(problem, flat, characters, states_per_column, prediction_method, num_nodes,
                                    num_tips)[0]
    finally:
        problem.close()


def parsimonious_acr(forest, character, prediction_method, states, num_nodes, num_tips):
    """Drop-in for ``pastml.parsimony.parsimonious_acr`` (``parsimony.py:245-331``)."""
    return parsimonious_acr_batch(forest, [character], prediction_method, [states], num_nodes, num_tips)[0]


def all_methods_hook(batch, current, results, note_restricted):
    """The parsimony part of the ALL meta-method (``ml.py:718-733``): the three MP reconstructions are added to the
    results, and the likelihood restricted to each of them is noted (or the inconsistency logged)."""
    from pastml_b200.ml import RESTRICTED_LOG_LIKELIHOOD_FORMAT_STR
    logger = logging.getLogger('pastml')
    stats = batch.models[0].forest_stats
    pars_results, raw = parsimony_on_problem(batch.problem, batch.flat, batch.characters,
                                             [m.states for m in batch.models], MP, stats.num_nodes, stats.num_tips)
    for c in range(batch.C):
        results[c].extend(pars_results[c])
    for m in (ACCTRAN, DOWNPASS, DELTRAN):
        masks, _ = raw[m]
        batch.problem.set_custom_masks(np.ascontiguousarray(masks.T))
        values, errs = batch.loglik_columns(None, MASK_CUSTOM, True, True)
        for c in range(batch.C):
            if c in errs:
                logger.error('{}\n{} parsimonious state selection is inconsistent in terms of ML.'
                             .format(errs[c].message, m))
            else:
                current[c][RESTRICTED_LOG_LIKELIHOOD_FORMAT_STR.format(m)] = values[c]

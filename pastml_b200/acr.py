"""``acr()``: ancestral character reconstruction for the columns of an annotation table.

Drop-in for ``pastml.acr.acr`` (``/root/reference/pastml/acr.py:76-279``): same signature, same result
dictionaries, same side effects on the tree nodes.  Where the reference maps a thread pool over the characters,
the columns are grouped here by (method, model family, number of states) and every group runs as one batched
problem on the GPU.
"""
import logging
from collections import OrderedDict

import numpy as np
import pandas as pd

from pastml_b200 import value2list, STATES, METHOD, CHARACTER
from pastml_b200.annotation import preannotate_forest, ForestStats
from pastml_b200.ml import is_ml, MPPA, ml_acr_batch
from pastml_b200.models import model2class, F81, HKY, JTT, HKY_STATES, JTT_STATES
from pastml_b200.parsimony import is_parsimonious, parsimonious_acr_batch, MP_METHODS
from pastml_b200.ml import ML_METHODS
from pastml_b200.tree import flatten_forest, resolve_trees, unresolve_trees, IS_POLYTOMY

COPY = 'COPY'

from pastml_b200 import _set_up_pastml_logger  # noqa: E402,F401  (pastml.acr exposes it)


def _serialize_acr(args):
    """``pastml.acr._serialize_acr((acr_result, work_dir))``: the parameter file and the marginal-probability table of one
    reconstruction (``acr.py:45-73``)."""
    from pastml_b200.pipeline import serialize_acr
    serialize_acr(*args)


def calculate_observed_freqs(character, forest, states):
    """Tip-state frequencies (a tip with n states counts 1/n for each) and the fraction of unannotated tips
    (``acr.py:282-299``)."""
    n = len(states)
    missing_data = 0.
    state2index = dict(zip(states, range(n)))
    observed_frequencies = np.zeros(n, np.float64)
    for tree in forest:
        for tip in tree:
            state = getattr(tip, character, set())
            if state:
                share = 1. / len(state)
                for s in state:
                    observed_frequencies[state2index[s]] += share
            else:
                missing_data += 1
    total_count = observed_frequencies.sum() + missing_data
    observed_frequencies /= observed_frequencies.sum()
    missing_data /= total_count
    return missing_data, observed_frequencies, state2index


def acr(forest, df=None, columns=None, column2states=None, prediction_method=MPPA, model=F81,
        column2parameters=None, column2rates=None, force_joint=True, threads=0, reoptimise=False, tau=0,
        resolve_polytomies=False, frequency_smoothing=False, device=0):
    """Reconstructs ancestral states for the given tree(s) and characters; returns the list of ACR result
    dictionaries.  ``threads`` is accepted for compatibility (the columns are batched on the device instead)."""
    logger = logging.getLogger('pastml')
    if not isinstance(forest, (list, tuple)):
        forest = [forest]
    if columns is None:
        if df is None:
            raise ValueError('Either the tree should be preannotated with character values and columns and '
                             'column2states specified, or an annotation dataframe provided!')
        columns = df.columns
        column2states = {column: np.array(sorted([_ for _ in df[column].unique() if not pd.isna(_) and '' != _]))
                         for column in columns}
        preannotate_forest(forest, df=df)
    columns = list(columns)

    flat = flatten_forest(forest)
    forest_stats = ForestStats(flat=flat)
    # the reconstruction overwrites the features with its predictions: polytomy resolution starts again from these
    n2c2states = {}
    for node in flat.nodes:
        c2states = {c: getattr(node, c) for c in columns if getattr(node, c, None)}
        if c2states:
            n2c2states[node] = c2states
    logger.debug('\n=============ACR===============================')
    column2parameters = column2parameters if column2parameters else {}
    column2rates = column2rates if column2rates else {}
    prediction_methods = value2list(len(columns), prediction_method, MPPA)
    models = value2list(len(columns), model, F81)

    def get_states(method, model_name, column):
        initial_states = column2states[column]
        if not is_ml(method) or model_name not in {HKY, JTT}:
            return initial_states
        states = HKY_STATES if HKY == model_name else JTT_STATES
        if not set(initial_states) & set(states):
            raise ValueError('The allowed states for model {} are {}, but your annotation file specifies {} as states '
                             'in column {}.'.format(model_name, ', '.join(states), ', '.join(initial_states), column))
        state_set = set(states)
        for node in flat.nodes:
            if hasattr(node, column):
                node.add_feature(column, state_set & getattr(node, column))
        return states

    # per character: method, model instance (ML) or states, observed frequencies
    settings = OrderedDict()
    for character, method, model_name in zip(columns, prediction_methods, models):
        logger.debug('ACR settings for {}:\n\tMethod:\t{}{}.'.format(
            character, method, '\n\tModel:\t{}'.format(model_name) if model_name and is_ml(method) else ''))
        if COPY == method or is_parsimonious(method):
            settings[character] = [method, get_states(method, model_name, character), None]
        elif is_ml(method):
            optimise_tau = tau is None or reoptimise
            if tau is None:
                tau = 0
            states = get_states(method, model_name, character)
            missing_data, observed_frequencies, state2index = calculate_observed_freqs(character, forest, states)
            logger.debug('Observed frequencies for {}:{}{}.'.format(
                character, ''.join('\n\tfrequency of {}:\t{:.6f}'.format(s, observed_frequencies[state2index[s]])
                                   for s in states),
                '\n\tfraction of missing data:\t{:.6f}'.format(missing_data) if missing_data else ''))
            instance = model2class[model_name](parameter_file=column2parameters.get(character),
                                               rate_matrix_file=column2rates.get(character), reoptimise=reoptimise,
                                               frequency_smoothing=frequency_smoothing, tau=tau,
                                               optimise_tau=optimise_tau, states=states, forest_stats=forest_stats,
                                               observed_frequencies=observed_frequencies, character=character)
            settings[character] = [method, instance, observed_frequencies]
        else:
            raise ValueError('Method {} is unknown, should be one of ML ({}), one of MP ({}) or {}'
                             .format(method, ', '.join(ML_METHODS), ', '.join(MP_METHODS), COPY))

    # group the characters into engine batches
    groups = OrderedDict()
    for character, (method, model_or_states, _) in settings.items():
        if COPY == method:
            key = (COPY, character)
        elif is_ml(method):
            key = ('ML', method, model_or_states.engine_kind, len(model_or_states.states))
        else:
            key = ('MP', method, len(model_or_states))
        groups.setdefault(key, []).append(character)

    def reconstruct(flat, forest_stats):
        """One reconstruction of every character on the forest as it is now."""
        per_character = {}
        for key, members in groups.items():
            if key[0] == COPY:
                ch = members[0]
                per_character[ch] = {CHARACTER: ch, STATES: settings[ch][1], METHOD: COPY}
            elif key[0] == 'ML':
                res = ml_acr_batch(forest, members, key[1], [settings[ch][1] for ch in members],
                                   [settings[ch][2] for ch in members], force_joint=force_joint, flat=flat, device=device)
                per_character.update(zip(members, res))
            else:
                res = parsimonious_acr_batch(forest, members, key[1], [settings[ch][1] for ch in members],
                                             forest_stats.num_nodes, forest_stats.num_tips, flat=flat, device=device)
                per_character.update(zip(members, res))
        results = []
        for character in settings:
            r = per_character[character]
            if isinstance(r, list):
                results.extend(r)
            else:
                results.append(r)
        return results

    acr_results = reconstruct(flat, forest_stats)
    if not resolve_polytomies:
        return acr_results

    # polytomy resolution (acr.py:234-278): children of a multifurcation predicted in one state that differs from their
    # parent's get a parent of their own; the likelihoods are recalculated with the fitted parameters frozen, and
    # resolutions that the recalculation does not support are undone until none is left to undo
    column2states = {res[CHARACTER]: res[STATES] for res in acr_results}
    column2copy = {res[CHARACTER]: res[METHOD] == COPY for res in acr_results}
    if not resolve_trees(column2states, forest):
        return acr_results
    level = logger.level
    logger.setLevel(logging.ERROR)
    for root in forest:
        for n in root.traverse():
            c2states = n2c2states.get(n, {})
            for c in columns:
                if c in c2states:
                    n.add_feature(c, c2states[c])
                elif not getattr(n, IS_POLYTOMY, False) or not column2copy.get(c, False):
                    n.del_feature(c)   # a COPY character keeps the state its new node was created with
    flat = flatten_forest(forest)
    forest_stats = ForestStats(flat=flat)
    for method, model_or_states, _ in settings.values():
        if is_ml(method):
            model_or_states.freeze()
            model_or_states.forest_stats = forest_stats
    acr_results = reconstruct(flat, forest_stats)
    while unresolve_trees(column2states, forest):
        flat = flatten_forest(forest)
        acr_results = reconstruct(flat, forest_stats)
    logger.setLevel(level)
    return acr_results

def pastml_pipeline(*args, **kwargs):
    """``pastml.acr.pastml_pipeline`` lives here in the reference; the implementation is in
    :mod:`pastml_b200.pipeline` (imported lazily: it builds on :func:`acr`)."""
    from pastml_b200.pipeline import pastml_pipeline as run
    return run(*args, **kwargs)

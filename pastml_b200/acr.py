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
from pastml_b200.tree import Tree, flatten_forest

COPY = 'COPY'


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
    if resolve_polytomies:
        raise NotImplementedError('resolve_polytomies is outside the accelerated path (SURVEY.md section 8f-4)')

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

    acr_results = []
    for character in settings:
        r = per_character[character]
        if isinstance(r, list):
            acr_results.extend(r)
        else:
            acr_results.append(r)
    return acr_results


def pastml_pipeline(*args, **kwargs):
    """``pastml.acr.pastml_pipeline`` lives here in the reference; the implementation is in
    :mod:`pastml_b200.pipeline` (imported lazily: it builds on :func:`acr`)."""
    from pastml_b200.pipeline import pastml_pipeline as run
    return run(*args, **kwargs)

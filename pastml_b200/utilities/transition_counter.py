"""``count_transitions()``: transition counts between the states of a character, written as a table.

Drop-in for the table-writing part of ``pastml.utilities.transition_counter.count_transitions``
(``/root/reference/pastml/utilities/transition_counter.py:24-173``): input validation as in ``pastml_pipeline``, the
model set up from a parameter file / dict (no optimisation), :func:`pastml_b200.ml.marginal_counts`, and the K x K table
``from`` x ``to``.  The html rendering of the result is visualisation and stays with the reference.
"""
import logging

import pandas as pd

from pastml_b200.acr import calculate_observed_freqs
from pastml_b200.annotation import ForestStats
from pastml_b200.ml import marginal_counts
from pastml_b200.models import model2class, F81, HKY, JTT, HKY_STATES, JTT_STATES
from pastml_b200.pipeline import validate_input


def count_transitions(tree, data, column, parameters, out_transitions, data_sep='\t', id_index=0, model=F81, threshold=1,
                      n_repetitions=1000, rate_matrix=None, work_dir=None, html=None, verbose=False, offline=False,
                      colours=None):
    """Counts the state transitions of ``column`` along the tree(s) over ``n_repetitions`` ancestral scenarios drawn
    from the marginal posterior under ``model`` with the given ``parameters`` (a parameter file as written by
    ``pastml_pipeline``, or a dict), and saves them to ``out_transitions``.  Returns the table."""
    if html:
        raise NotImplementedError('html: the rendering of the transition graph is outside the accelerated ACR path')
    logging.getLogger('pastml').setLevel(logging.DEBUG if verbose else logging.ERROR)
    forest, columns, column2states, parameters, rates = \
        validate_input(tree, [column], data, data_sep, id_index, parameters=[parameters] if parameters else None,
                       rates=[rate_matrix] if rate_matrix else None)
    column = columns[0]
    states = column2states[column]
    if model in {HKY, JTT}:
        initial_states = states
        states = HKY_STATES if HKY == model else JTT_STATES
        if not set(initial_states) & set(states):
            raise ValueError('The allowed states for model {} are {}, but your annotation file specifies {} as states '
                             'in column {}.'.format(model, ', '.join(states), ', '.join(initial_states), column))
    _, observed_frequencies, _ = calculate_observed_freqs(column, forest, states)
    instance = model2class[model](parameter_file=parameters.get(column), rate_matrix_file=rates.get(column),
                                  reoptimise=False, states=states, forest_stats=ForestStats(forest),
                                  observed_frequencies=observed_frequencies)
    state_set = set(states)
    for root in forest:
        for n in root.traverse():
            if hasattr(n, column):
                n.add_feature(column, state_set & getattr(n, column))
    result = marginal_counts(forest, column, instance, n_repetitions=n_repetitions)
    df = pd.DataFrame(data=result, columns=states, index=states)
    df.to_csv(out_transitions, sep='\t', index_label='from')
    logging.getLogger('pastml').info('Transition counts are saved as {}.'.format(out_transitions))
    return df

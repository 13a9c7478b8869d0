"""Forward simulation of character states down a tree (the recipe of ``pastml/utilities/state_simulator.py:6-31``:
root ~ the model's frequencies, child ~ the row of P(t) selected by its parent's state), vectorised over repetitions.

Every node receives the feature ``character`` = int array [n_repetitions] of state indices, as the reference's
function leaves it (``tests/MRANDJCTest.py`` and ``tests/CUSTOM_RATESTest.py`` read it that way).
"""
import numpy as np


def simulate_states(tree, model, character, n_repetitions=1_000, rng=None):
    rng = np.random.default_rng() if rng is None else rng
    k = len(model.states)
    for node in tree.traverse('levelorder'):
        u = rng.random(n_repetitions)
        if node.is_root():
            cdf = np.cumsum(np.asarray(model.frequencies, dtype=np.float64))
            states = np.searchsorted(cdf / cdf[-1], u, side='right')
        else:
            # inverse-CDF draw from the parent's row: one comparison table, no loop over the parent states
            p = np.maximum(model.get_Pij_t(node.dist), 0.)
            cdf = np.cumsum(p, axis=1)
            cdf /= cdf[:, -1:]
            rows = cdf[getattr(node.up, character)]
            states = (u[:, None] >= rows).sum(axis=1)
        node.add_feature(character, np.minimum(states, k - 1).astype(int))
    return tree

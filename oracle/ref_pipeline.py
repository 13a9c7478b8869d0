"""TEST INFRASTRUCTURE ONLY -- runs the UNMODIFIED reference's fixed-parameter MPPA pipeline on flat arrays.

The call sequence is the marginal block of ``pastml.ml.ml_acr`` (``/root/reference/pastml/ml.py:697-712, 735-742``):
initialize_allowed_states, alter_zero_node_allowed_states, get_bottom_up_loglikelihood(alter=False),
calculate_top_down_likelihood, calculate_marginal_likelihoods, convert_likelihoods_to_probabilities,
unalter_zero_node_allowed_states, choose_ancestral_states_map, choose_ancestral_states_mppa -- every function is the
reference's own, imported from `/root/reference` or from its staged copy `oracle/_ref` (oracle/stage_reference.py);
`pastml_b200.tree.Tree` plays ete3.Tree (oracle/ref_env.py).  Used by bench.py's reference arm and by tests.
"""
import os
import time

import numpy as np


def reference_root():
    """Where an importable copy of the reference lives, or None."""
    here = os.path.dirname(os.path.abspath(__file__))
    for cand in (os.environ.get('PASTML_REFERENCE_ROOT'), '/root/reference', os.path.join(here, '_ref')):
        if cand and os.path.isdir(os.path.join(cand, 'pastml')):
            return cand
    return None


def _install():
    root = reference_root()
    if root is None:
        raise RuntimeError('no copy of the reference is available (run oracle/stage_reference.py in the build container)')
    os.environ['PASTML_REFERENCE_ROOT'] = root
    from oracle import ref_env
    ref_env.REFERENCE_ROOT = root
    ref_env.install()


def run_column(parent, dist, annot_mask, K, pi, sf, repeats=1):
    """One column through the reference's marginal pipeline, `repeats` times on the same tree objects.  Returns
    (seconds per repeat, log-likelihood, probabilities [N][K] in level order, MPPA masks uint64 [N])."""
    _install()
    import pandas as pd  # noqa: F401
    from pastml.annotation import ForestStats
    from pastml.ml import (initialize_allowed_states, alter_zero_node_allowed_states, get_bottom_up_loglikelihood,
                           calculate_top_down_likelihood, calculate_marginal_likelihoods, convert_likelihoods_to_probabilities,
                           unalter_zero_node_allowed_states, choose_ancestral_states_map, choose_ancestral_states_mppa,
                           ALLOWED_STATES)
    from pastml.models.F81Model import F81Model
    from pastml import get_personalized_feature_name
    from pastml_b200.tree import forest_from_arrays
    roots = forest_from_arrays(np.asarray(parent), np.asarray(dist), ['n{}'.format(i) for i in range(len(parent))])
    nodes, frontier = [], list(roots)
    while frontier:
        nodes.extend(frontier)
        frontier = [ch for n in frontier for ch in n.children]
    states = np.array(['s{:02d}'.format(k) for k in range(K)])
    character = 'col'
    for node, m in zip(nodes, np.asarray(annot_mask).tolist()):
        if m:
            node.add_feature(character, {states[k] for k in range(K) if (m >> k) & 1})
    model = F81Model(states=states, forest_stats=ForestStats(roots), sf=float(sf), frequencies=np.asarray(pi, dtype=float))
    t0 = time.perf_counter()
    for _ in range(repeats):
        for tree in roots:
            initialize_allowed_states(tree, character, model.states)
            altered = alter_zero_node_allowed_states(tree, character) if 0 == model.tau else []
            lh = get_bottom_up_loglikelihood(tree=tree, character=character, is_marginal=True, model=model, alter=False)
            calculate_top_down_likelihood(tree, character, model=model)
            calculate_marginal_likelihoods(tree, character, model.frequencies)
            probs = convert_likelihoods_to_probabilities(tree, character, model.states)
            if altered:
                unalter_zero_node_allowed_states(altered, character)
            choose_ancestral_states_map(tree, character, model.states)
            choose_ancestral_states_mppa(tree, character, model.states, force_joint=False)
    dt = (time.perf_counter() - t0) / repeats
    feature = get_personalized_feature_name(character, ALLOWED_STATES)
    masks = np.array([int(sum(1 << k for k, v in enumerate(getattr(n, feature)) if v)) for n in nodes], dtype=np.uint64)
    return dt, float(lh), probs.to_numpy(), masks

"""``acr(..., resolve_polytomies=True)`` against the unmodified reference (``acr.py:234-278``, ``tree.py:344-489``) on
seeded trees with many multifurcations: resolved topology, node names, predicted states and log-likelihood
(``oracle/make_golden_polytomies.py`` -> ``tests/golden/polytomies/*.npz``; fixed model parameters)."""
import glob
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

BASE = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'polytomies')
STATES = ['s0', 's1', 's2']
COLUMN = 'trait'
PARAMS = {'scaling_factor': 0.6, 's0': 0.5, 's1': 0.3, 's2': 0.2}


@pytest.mark.parametrize('path', sorted(glob.glob(os.path.join(BASE, '*.npz'))), ids=os.path.basename)
def test_polytomy_resolution_matches_reference(path):
    from pastml_b200.acr import acr
    from pastml_b200.pipeline import annotate_dist_to_root
    from pastml_b200.tree import forest_from_arrays, flatten_forest
    z = np.load(path)
    roots = forest_from_arrays(z['in_parent'], z['in_dist'], [str(n) for n in z['in_names']])
    nodes = flatten_forest(roots).nodes
    for node, m in zip(nodes, z['in_annot'].tolist()):
        if m:
            node.add_feature(COLUMN, {STATES[k] for k in range(3) if (m >> k) & 1})
    annotate_dist_to_root(roots)
    res = acr(roots, columns=[COLUMN], column2states={COLUMN: np.array(STATES)}, prediction_method='MPPA', model='F81',
              column2parameters={COLUMN: dict(PARAMS)}, force_joint=True, resolve_polytomies=True)[0]
    flat = flatten_forest(roots)
    assert np.array_equal(flat.parent, z['parent'])
    np.testing.assert_allclose(flat.dist, z['dist'], rtol=0, atol=1e-12)
    assert [n.name for n in flat.nodes] == [str(n) for n in z['names']]
    masks = np.array([sum(1 << STATES.index(s) for s in getattr(n, COLUMN)) for n in flat.nodes], dtype=np.uint64)
    assert np.array_equal(masks, z['mask'])
    assert res['log_likelihood'] == pytest.approx(float(z['log_likelihood']), rel=1e-10)

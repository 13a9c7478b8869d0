"""Options of the fit the main fixtures do not exercise -- smoothing (tau optimised), frequency smoothing,
re-optimisation from a parameter file, parameter files of JC / HKY, DENV2 under MAP / JOINT / MPPA x JC / EFT, and
the batch of HIV1C drug-resistance columns in one call -- against outcomes of the unmodified reference
(``oracle/make_golden_options.py`` -> ``tests/golden/options/``)."""
import glob
import json
import os

import numpy as np
import pytest

from helpers import GOLDEN_DIR, masks_from_feature
from pastml_b200.tree import FlatForest, forest_from_arrays, flatten_forest

pytestmark = pytest.mark.gpu

OPT_DIR = os.path.join(GOLDEN_DIR, 'options')
SINGLE = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(OPT_DIR, '*.npz'))
                if not os.path.basename(p).startswith('hiv1c_drms'))


def _forest(z, column, states):
    names = [str(_) for _ in z['names']]
    roots = forest_from_arrays(z['parent'], z['dist'], names)
    nodes, frontier = [], list(roots)
    while frontier:
        nodes.extend(frontier)
        frontier = [ch for n in frontier for ch in n.children]
    for node, m, h in zip(nodes, z['annot_mask'].tolist(), z['has_state'].tolist()):
        if h:
            node.add_feature(column, {states[k] for k in range(len(states)) if (int(m) >> k) & 1})
    return roots


@pytest.mark.parametrize('name', SINGLE)
def test_option_fixture(name, tmp_path):
    from pastml_b200.acr import acr
    z = dict(np.load(os.path.join(OPT_DIR, name + '.npz'), allow_pickle=False))
    meta = json.loads(str(z['meta']))
    column, model = meta['column'], meta['model']
    states = [str(_) for _ in z['states']]
    options = dict(meta['options'])
    if 'parameter_file_text' in z:
        pfile = str(tmp_path / 'params.tab')
        with open(pfile, 'w') as f:
            f.write(str(z['parameter_file_text']))
        options['column2parameters'] = {column: pfile}
    forest = _forest(z, column, states)
    res = acr(forest, columns=[column], column2states={column: np.array(states)}, model=model, **options)
    r = res[-1]
    mdl = r['model']
    ref_lh = float(z['log_likelihood'])
    ens = np.concatenate((z['ens_log_likelihood'], [ref_lh])) if 'ens_log_likelihood' in z else np.array([ref_lh])
    spread = float(np.abs(ens - ref_lh).max() / abs(ref_lh))
    flat = flatten_forest(forest)
    if spread < 1e-8:
        # fixed parameters, or an optimum the reference itself reproduces under one ulp of objective noise
        fixed = 'ens_log_likelihood' not in z and 'column2parameters' in options
        assert r['log_likelihood'] == pytest.approx(ref_lh, rel=1e-12 if fixed else max(1e-11, 2 * spread))
        ens_sf = np.concatenate((z['ens_sf'], [float(z['sf'])])) if 'ens_sf' in z else np.array([float(z['sf'])])
        self_sf = float(np.abs(ens_sf - float(z['sf'])).max() / float(z['sf']))
        assert float(mdl.sf) == pytest.approx(float(z['sf']), rel=1e-14 if fixed else max(2e-6, 2 * self_sf))
        assert float(mdl.tau) == pytest.approx(float(z['tau']), abs=1e-9)
        if hasattr(mdl, 'frequencies'):
            np.testing.assert_allclose(mdl.frequencies, z['pi'], rtol=1e-13 if fixed else 2e-4, atol=1e-9)
        if model == 'HKY':
            # kappa is the least identified parameter: the reference moves it by ~1.2e-4 under one ulp of noise
            ens_k = np.concatenate((z['ens_kappa'], [float(z['kappa'])])) if 'ens_kappa' in z else np.array([float(z['kappa'])])
            self_k = float(np.abs(ens_k - float(z['kappa'])).max() / float(z['kappa']))
            assert float(mdl.kappa) == pytest.approx(float(z['kappa']), rel=1e-14 if fixed else max(1e-4, 2 * self_k))
        for m in ('JOINT', 'MAP', 'MPPA'):
            key = 'log_likelihood_restricted_' + m
            if key in z and key in r:
                # (a restricted likelihood is not at an optimum: it moves in first order with the fitted parameters,
                # which carry the optimiser's own 1e-6 .. 1e-4 spread)
                assert r[key] == pytest.approx(float(z[key]), rel=1e-12 if fixed else 1e-5), key
        assert np.array_equal(masks_from_feature(flat.nodes, states, column), z['state_mask'].astype(np.uint64))
        for key in ('num_unresolved_nodes', 'num_states_per_node_avg'):
            if key in z:
                assert r[key] == pytest.approx(float(z[key]), rel=1e-12)
    else:
        # a surface on which the reference's own optimum moves (smoothing: tau sits on its bound, sf / pi trade off)
        margin = 0.5 * (ens.max() - ens.min())
        assert ens.min() - margin <= r['log_likelihood'] <= ens.max() + margin, (r['log_likelihood'], ens)
        params = {'scaling_factor': float(z['sf']), 'smoothing_factor': float(z['tau'])}
        params.update({s: float(f) for s, f in zip(states, z['pi'])})
        forest = _forest(z, column, states)
        kw = {k: v for k, v in options.items() if k not in ('tau', 'frequency_smoothing')}
        at_ref = acr(forest, columns=[column], column2states={column: np.array(states)}, model=model,
                     column2parameters={column: params}, **kw)[-1]
        assert at_ref['log_likelihood'] == pytest.approx(ref_lh, rel=1e-11)
        assert np.array_equal(masks_from_feature(flatten_forest(forest).nodes, states, column),
                              z['state_mask'].astype(np.uint64))


def test_hiv1c_drm_batch():
    """Every binary drug-resistance column of the HIV1C example in ONE acr() call (JC): the columns are batched on
    the device as lanes of one problem and fitted in lockstep; each must land where the reference's separate fits
    did."""
    from pastml_b200.acr import acr
    import pandas as pd
    path = os.path.join(OPT_DIR, 'hiv1c_drms_jc.npz')
    if not os.path.exists(path):
        pytest.skip('fixture not generated')
    z = dict(np.load(path, allow_pickle=False))   # (an NpzFile decompresses an array again at every access)
    names = [str(_) for _ in z['names']]
    columns = [str(_) for _ in z['columns']]
    roots = forest_from_arrays(z['parent'], z['dist'], names)
    flat0 = FlatForest(z['parent'], z['dist'], names=names)
    data = {}
    for c, col in enumerate(columns):
        states = str(z['states'][c]).split('|')
        annot, has = z['annot_mask'][c].tolist(), z['has_state'][c].tolist()
        data[col] = {names[i]: (states[annot[i].bit_length() - 1] if annot[i] else None)
                     for i in range(flat0.n_nodes) if has[i]}
    df = pd.DataFrame(data)
    res = acr(roots[0], df=df, prediction_method='MPPA', model='JC')
    flat = flatten_forest(roots)
    by_col = {r['character']: r for r in res}
    worst = 0.
    for c, col in enumerate(columns):
        r = by_col[col]
        states = str(z['states'][c]).split('|')
        assert list(r['states']) == states
        ref_lh = float(z['log_likelihood'][c])
        assert r['log_likelihood'] == pytest.approx(ref_lh, rel=1e-9, abs=1e-12), col
        if ref_lh != 0:   # (a column whose tips in the tree all share one state has likelihood 1 at any rate)
            assert float(r['model'].sf) == pytest.approx(float(z['sf'][c]), rel=1e-4), col
            worst = max(worst, abs(r['log_likelihood'] - ref_lh) / abs(ref_lh))
        ours = masks_from_feature(flat.nodes, states, col)
        assert int((ours != z['mppa_mask'][c].astype(np.uint64)).sum()) == 0, col
    print('hiv1c drm batch: {} columns, worst |dlh|/|lh| = {:.2e}'.format(len(columns), worst))

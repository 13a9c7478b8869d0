"""``pastml_b200.ml.marginal_counts`` against the reference's ``marginal_counts`` (``ml.py:753-862``).  The function
samples ancestral scenarios, so the comparison is statistical: the mean over several runs must agree with the mean of
the reference's runs (``oracle/make_golden_counts.py`` -> ``tests/golden/counts/counts_*.npz``) within the Monte-Carlo error."""
import os

import numpy as np
import pytest

from helpers import Golden, annotated_forest

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
CASES = ['albania_f81', 'synth_s4_n30_k3_f81_forest', 'sim152_hky_skew', 'synth_s2_n60_k7_f81']


def _model(g, roots):
    from pastml_b200.annotation import ForestStats
    from pastml_b200.models.F81Model import F81Model
    from pastml_b200.models.HKYModel import HKYModel
    fs = ForestStats(roots)
    kw = dict(forest_stats=fs, sf=float(g['sf']), tau=float(g['tau']))
    if g.kind == 0:
        return F81Model(states=np.array(g.states), frequencies=np.array(g['pi'], dtype=float), **kw)
    return HKYModel(states=np.array(g.states), frequencies=np.array(g['pi'], dtype=float), kappa=float(g['kappa']), **kw)


@pytest.mark.parametrize('name', CASES)
def test_marginal_counts_match_reference_statistics(name):
    from pastml_b200.ml import marginal_counts
    g = Golden(name)
    want = np.load(os.path.join(GOLDEN, 'counts', 'counts_{}.npz'.format(name)))
    runs_ref, n_rep = int(want['runs']), int(want['n_repetitions'])
    runs = []
    for seed in range(runs_ref):
        roots = annotated_forest(g)
        np.random.seed(500 + seed)
        runs.append(marginal_counts(roots, g.meta['column'], _model(g, roots), n_repetitions=n_rep))
    runs = np.array(runs)
    assert runs.shape[1:] == want['mean'].shape
    mean, std = runs.mean(axis=0), runs.std(axis=0, ddof=1)
    sigma = np.sqrt((std ** 2 + want['std'] ** 2) / runs_ref)
    dev = np.abs(mean - want['mean'])
    assert np.all(dev <= 5. * sigma + 0.02), (dev.max(), sigma.max())
    # the number of changes per scenario (off-diagonal mass), a tighter aggregate
    off = lambda m: m.sum() - np.trace(m)
    assert off(mean) == pytest.approx(off(want['mean']), rel=0.01, abs=0.05)


def test_count_transitions_table(tmp_path):
    """The table-writing wrapper on the Albania pipeline inputs with the reference-fitted parameter file: labels, and
    a number of changes per scenario in line with the reference's marginal_counts runs on the same tree."""
    import pandas as pd
    from pastml_b200.utilities.transition_counter import count_transitions
    src = os.path.join(GOLDEN, 'pipeline', 'albania_mppa_f81')
    out = str(tmp_path / 'transitions.tab')
    np.random.seed(3)
    count_transitions(tree=os.path.join(src, 'input.nwk'), data=os.path.join(src, 'input.tab'), column='Country',
                      parameters=os.path.join(src, 'expected', 'params.character_Country.method_MPPA.model_F81.tab'),
                      out_transitions=out, model='F81', n_repetitions=4000)
    df = pd.read_csv(out, sep='\t', index_col=0)
    want = np.load(os.path.join(GOLDEN, 'counts', 'counts_albania_f81.npz'))
    assert list(df.index) == list(df.columns) == sorted(df.columns) and df.index.name == 'from'
    assert df.values.shape == want['mean'].shape
    off = lambda m: m.sum() - np.trace(m)
    assert off(df.values) == pytest.approx(off(want['mean']), rel=0.02)
    with pytest.raises(NotImplementedError):
        count_transitions(tree=os.path.join(src, 'input.nwk'), data=os.path.join(src, 'input.tab'), column='Country',
                          parameters=None, out_transitions=out, html=str(tmp_path / 'x.html'))


@pytest.mark.parametrize('name', CASES)
def test_marginal_counts_many_repetitions(name):
    """The sampler runs on the device, so 200k scenarios cost nothing: its own Monte-Carlo error becomes negligible
    and the comparison is against the reference's mean within the REFERENCE's standard error alone.  Same seed, same
    counts (the random stream is counter-based)."""
    from pastml_b200.ml import marginal_counts
    g = Golden(name)
    want = np.load(os.path.join(GOLDEN, 'counts', 'counts_{}.npz'.format(name)))
    runs_ref = int(want['runs'])
    out = []
    for _ in range(2):
        roots = annotated_forest(g)
        np.random.seed(77)
        out.append(marginal_counts(roots, g.meta['column'], _model(g, roots), n_repetitions=200_000))
    assert np.array_equal(out[0], out[1])
    sigma = want['std'] / np.sqrt(runs_ref)
    dev = np.abs(out[0] - want['mean'])
    assert np.all(dev <= 5. * sigma + 0.02), (dev.max(), sigma.max())
    off = lambda m: m.sum() - np.trace(m)
    assert off(out[0]) == pytest.approx(off(want['mean']), rel=0.01, abs=0.05)

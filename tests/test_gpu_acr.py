"""The pastml-compatible Python API (acr / ml_acr / parsimonious_acr) on the GPU engine against the golden
fixtures of the unmodified reference: fixed parameters (strict) and through the optimiser (SURVEY.md section 0:
log-likelihood strict, parameters ~1e-6)."""
import numpy as np
import pytest

from helpers import Golden, golden_names, annotated_forest, masks_from_feature

pytestmark = pytest.mark.gpu


def _fixed_params(g):
    params = {'scaling_factor': float(g['sf']), 'smoothing_factor': float(g['tau'])}
    if g.model_name in ('F81', 'HKY', 'CUSTOM_RATES'):
        params.update({s: float(f) for s, f in zip(g.states, g['pi'])})
    if g.model_name == 'HKY':
        params['kappa'] = float(g['kappa'])
    return params


def _rates_file(g, tmp_path):
    if g.model_name != 'CUSTOM_RATES':
        return None
    from pastml_b200.models.generator import save_matrix
    path = str(tmp_path / 'rates.txt')
    save_matrix(g.states, g['rate_matrix'], path)
    return path


def _acr(g, method, tmp_path, params=None, **kw):
    from pastml_b200.acr import acr
    from pastml_b200.tree import flatten_forest
    col = g.meta['column']
    forest = annotated_forest(g)
    kwargs = dict(columns=[col], column2states={col: np.array(g.states)}, prediction_method=method,
                  model=g.model_name, force_joint=g.meta['force_joint'], tau=float(g['tau']) if params is None else 0)
    if params is not None:
        kwargs['column2parameters'] = {col: params}
    rates = _rates_file(g, tmp_path)
    if rates:
        kwargs['column2rates'] = {col: rates}
    kwargs.update(kw)
    res = acr(forest, **kwargs)
    return flatten_forest(forest), res


@pytest.mark.parametrize('name', golden_names())
def test_fixed_parameters_all_methods(name, tmp_path):
    g = Golden(name)
    col = g.meta['column']
    flat, res = _acr(g, 'MPPA', tmp_path, _fixed_params(g))
    r = res[0]
    assert r['log_likelihood'] == pytest.approx(float(g['log_likelihood']), rel=1e-10)
    for m in ('JOINT', 'MAP', 'MPPA'):
        assert r['log_likelihood_restricted_' + m] == pytest.approx(float(g['lh_restricted_' + m]), rel=1e-10)
    mp = r['marginal_probabilities']
    assert list(mp.index) == [g.flat.names[i] for i in g.flat.ref_order]
    assert list(mp.columns) == g.states
    np.testing.assert_allclose(mp.values, g['marginal_probabilities'][g.flat.ref_order], rtol=0, atol=1e-9)
    assert np.array_equal(masks_from_feature(flat.nodes, g.states, col), g['mppa_mask'])
    assert str(r['num_scenarios']) == g.meta['num_scenarios']
    assert r['num_unresolved_nodes'] == int(g['num_unresolved_nodes'])
    assert r['num_states_per_node_avg'] == pytest.approx(float(g['num_states_per_node_avg']), rel=1e-12)
    # the LIKELIHOOD / LIKELIHOOD_SF node features encode the same total likelihood at every node
    tot = [np.log10(getattr(n, col + '_LIKELIHOOD').sum()) - getattr(n, col + '_LIKELIHOOD_SF') for n in flat.nodes]
    ref = np.log10(g['lh'].sum(axis=1)) - g['lh_sf']
    np.testing.assert_allclose(tot, ref, rtol=0, atol=1e-8)

    flat, res = _acr(g, 'JOINT', tmp_path, _fixed_params(g))
    assert np.array_equal(masks_from_feature(flat.nodes, g.states, col), g['joint_mask'])
    flat, res = _acr(g, 'MAP', tmp_path, _fixed_params(g))
    assert np.array_equal(masks_from_feature(flat.nodes, g.states, col), g['map_mask'])
    if 'DOWNPASS_mask' in g:
        for m in ('ACCTRAN', 'DOWNPASS', 'DELTRAN'):
            flat, res = _acr(g, m, tmp_path, None, model=None)
            assert np.array_equal(masks_from_feature(flat.nodes, g.states, col), g[m + '_mask']), m
            assert res[0]['steps'] == int(g[m + '_steps'])


OPTIMISED = [n for n in golden_names() if Golden(n).meta['optimised']]


@pytest.mark.parametrize('name', OPTIMISED)
def test_through_the_optimiser(name, tmp_path):
    """Same starting points, same scipy L-BFGS-B on the host.  L-BFGS-B differentiates the objective by forward
    differences with h = 1e-8, so the optimum is only reproducible as far as the reference reproduces ITSELF when
    its own objective is perturbed by one ulp: the fixtures carry that ensemble (``ens_lh``, made by
    ``oracle/make_golden.py --ensemble``).  Where the ensemble is tight the comparison is strict; where the
    likelihood surface is flat (the DENV2 columns: the reference's own optimum moves by ~1 log unit) the result
    must fall inside the reference's spread."""
    g = Golden(name)
    col = g.meta['column']
    flat, res = _acr(g, 'MPPA', tmp_path, None)
    r = res[0]
    mdl = r['model']
    ref_lh = float(g['log_likelihood'])
    ens = np.concatenate((g['ens_lh'], [ref_lh]))
    ens_sf = np.concatenate((g['ens_sf'], [float(g['sf'])]))
    spread = float(np.abs(ens - ref_lh).max() / abs(ref_lh))
    if spread < 1e-8:
        # tolerances from the recorded outcome (profiles/r2_optimiser_parity.json, tools/optimiser_parity.py): the
        # engine's fit lies INSIDE the spread the reference shows against itself under one ulp of objective noise
        # (achieved: |dlh|/|lh| <= 3.5e-12, |dsf|/sf <= 2.6e-6, |dpi|/pi <= 3.1e-5, |d marginal| <= 3.5e-6)
        self_lh = spread
        self_sf = float(np.abs(ens_sf - float(g['sf'])).max() / float(g['sf']))
        assert abs(r['log_likelihood'] - ref_lh) <= max(1e-11, 2 * self_lh) * abs(ref_lh)
        assert abs(float(mdl.sf) - float(g['sf'])) <= max(1e-6, 2 * self_sf) * float(g['sf'])
        if hasattr(mdl, 'frequencies'):
            np.testing.assert_allclose(mdl.frequencies, g['pi'], rtol=1e-4, atol=1e-9)
        if g.model_name == 'HKY':
            assert float(mdl.kappa) == pytest.approx(float(g['kappa']), rel=5e-5)
        np.testing.assert_allclose(r['marginal_probabilities'].values, g['marginal_probabilities'][g.flat.ref_order],
                                   rtol=0, atol=1e-5)
        ours = masks_from_feature(flat.nodes, g.states, col)
        mismatch = int((ours != g['mppa_mask']).sum())
        assert mismatch == 0, '{} nodes differ after optimisation'.format(mismatch)
    else:
        margin = 0.5 * (ens.max() - ens.min())
        assert ens.min() - margin <= r['log_likelihood'] <= ens.max() + margin, (r['log_likelihood'], ens)
        # (the likelihood AT the reference's reported parameters is compared strictly in
        # test_fixed_parameters_all_methods; re-optimising from them is no yardstick -- the reference itself moves away
        # from its own optimum on these columns: -35.986 -> -36.391 for denv2_genotype_f81)
    # whatever the optimum, the reported likelihood is the likelihood at the reported parameters
    from pastml_b200.ml import ColumnBatch
    from pastml_b200.tree import flatten_forest
    forest = annotated_forest(g)
    batch = ColumnBatch(flatten_forest(forest), [col], [mdl])
    values, errs = batch.loglik_columns()
    batch.close()
    assert not errs and values[0] == pytest.approx(r['log_likelihood'], rel=1e-12)


def test_all_meta_method(tmp_path):
    """ALL returns JOINT, MAP, the three parsimony reconstructions and MPPA, with restricted likelihoods."""
    g = Golden('albania_f81')
    col = g.meta['column']
    flat, res = _acr(g, 'ALL', tmp_path, _fixed_params(g))
    methods = [r['method'] for r in res]
    assert methods == ['JOINT', 'MAP', 'ACCTRAN', 'DOWNPASS', 'DELTRAN', 'MPPA']
    assert np.array_equal(masks_from_feature(flat.nodes, g.states, col + '_JOINT'), g['joint_mask'])
    assert np.array_equal(masks_from_feature(flat.nodes, g.states, col + '_MAP'), g['map_mask'])
    assert np.array_equal(masks_from_feature(flat.nodes, g.states, col + '_MPPA'), g['mppa_mask'])
    for m in ('ACCTRAN', 'DOWNPASS', 'DELTRAN'):
        assert np.array_equal(masks_from_feature(flat.nodes, g.states, col + '_' + m), g[m + '_mask'])
    last = res[-1]
    for m in ('JOINT', 'MAP', 'MPPA'):
        assert last['log_likelihood_restricted_' + m] == pytest.approx(float(g['lh_restricted_' + m]), rel=1e-10)
    # on this tree the parsimonious states contradict each other across zero-length internal branches: like the
    # reference (ml.py:725-733) the engine reports it and leaves those restricted likelihoods out
    for m in ('ACCTRAN', 'DOWNPASS', 'DELTRAN'):
        assert 'log_likelihood_restricted_' + m not in last
    assert sorted(res[0].keys()) == ['character', 'log_likelihood', 'log_likelihood_restricted_JOINT', 'method',
                                     'model', 'states']


def test_columns_batched_like_separate_runs(tmp_path):
    """Two columns of the same tree in one acr() call give what two separate calls give."""
    from pastml_b200.acr import acr
    from pastml_b200.tree import flatten_forest
    ga, gb = Golden('denv2_location_f81'), Golden('denv2_genotype_f81')
    forest = annotated_forest(ga)
    other = annotated_forest(gb)
    for n, m in zip(flatten_forest(forest).nodes, flatten_forest(other).nodes):
        if hasattr(m, 'genotype'):
            n.add_feature('genotype', getattr(m, 'genotype'))
    res = acr(forest, columns=['Location', 'genotype'],
              column2states={'Location': np.array(ga.states), 'genotype': np.array(gb.states)},
              column2parameters={'Location': _fixed_params(ga), 'genotype': _fixed_params(gb)})
    assert [r['character'] for r in res] == ['Location', 'genotype']
    assert res[0]['log_likelihood'] == pytest.approx(float(ga['log_likelihood']), rel=1e-10)
    assert res[1]['log_likelihood'] == pytest.approx(float(gb['log_likelihood']), rel=1e-10)
    flat = flatten_forest(forest)
    assert np.array_equal(masks_from_feature(flat.nodes, ga.states, 'Location'), ga['mppa_mask'])
    assert np.array_equal(masks_from_feature(flat.nodes, gb.states, 'genotype'), gb['mppa_mask'])


def test_allowed_state_helpers_feed_the_likelihood():
    """The reference's own calling pattern around ``get_bottom_up_loglikelihood`` (``ml.py:786-790``): allowed-state
    arrays initialised and altered by the caller, then the likelihood with ``alter=False`` -- equal to letting the engine
    alter, and to the reference's value at the same parameters."""
    from pastml_b200.annotation import ForestStats
    from pastml_b200.ml import (get_bottom_up_loglikelihood, initialize_allowed_states, alter_zero_node_allowed_states,
                                unalter_zero_node_allowed_states)
    from pastml_b200.models.F81Model import F81Model
    g = Golden('albania_f81')
    col = g.meta['column']
    roots = annotated_forest(g)
    tree = roots[0]
    model = F81Model(states=np.array(g.states), forest_stats=ForestStats(roots), sf=float(g['sf']),
                     frequencies=np.array(g['pi'], dtype=float), tau=float(g['tau']))
    model.freeze()
    by_engine = get_bottom_up_loglikelihood(tree, col, model, is_marginal=True, alter=True)
    initialize_allowed_states(tree, col, np.array(g.states))
    altered = alter_zero_node_allowed_states(tree, col)
    assert altered   # this tree has contradicting annotations across zero-length branches
    by_caller = get_bottom_up_loglikelihood(tree, col, model, is_marginal=True, alter=False)
    unalter_zero_node_allowed_states(altered, col)
    assert by_caller == pytest.approx(by_engine, rel=1e-13)
    assert by_engine == pytest.approx(float(g['log_likelihood']), rel=1e-10)

"""Pins the CPU oracle (oracle/pastml_oracle.c) against outputs of the unmodified reference (tests/golden)."""
import numpy as np
import pytest

from helpers import Golden, golden_names
from oracle import cpu_oracle


@pytest.fixture(scope='module', params=golden_names())
def case(request):
    g = Golden(request.param)
    res = cpu_oracle.ml_acr_fixed(g.flat, g.oracle_model(), g['annot_mask'], g['has_state'], method='MPPA',
                                  force_joint=g.meta['force_joint'])
    return g, res


def test_loglikelihood(case):
    g, res = case
    assert res['log_likelihood'] == pytest.approx(float(g['log_likelihood']), rel=1e-12)
    for m in ('JOINT', 'MAP', 'MPPA'):
        assert res['log_likelihood_restricted_' + m] == pytest.approx(float(g['lh_restricted_' + m]), rel=1e-12)


def test_marginal_probabilities(case):
    g, res = case
    np.testing.assert_allclose(res['marginal_probabilities'], g['marginal_probabilities'], rtol=0, atol=1e-12)


def test_lh_features(case):
    g, res = case
    # LIKELIHOOD after MAP/MPPA masking and its scaling factor (ml.py:459-460, 541-542, 593-594).  The shift
    # chosen by rescale_log depends on the smallest (possibly rounding-noise) entry, so compare what the pair
    # encodes: the normalised vector and log10(sum) - SF, the node's total log10 likelihood.
    ours, ref = res['lh_after_map'], g['lh']
    np.testing.assert_allclose(ours / ours.sum(axis=1, keepdims=True), ref / ref.sum(axis=1, keepdims=True),
                               rtol=0, atol=1e-12)
    np.testing.assert_allclose(np.log10(ours.sum(axis=1)) - res['lh_sf'], np.log10(ref.sum(axis=1)) - g['lh_sf'],
                               rtol=0, atol=1e-9)


def test_states(case):
    g, res = case
    assert np.array_equal(res['joint_mask'], g['joint_mask'])
    assert np.array_equal(res['map_mask'], g['map_mask'])
    assert np.array_equal(res['mppa_mask'], g['mppa_mask'])
    assert str(res['num_scenarios']) == g.meta['num_scenarios']
    assert res['num_unresolved_nodes'] == int(g['num_unresolved_nodes'])


def test_parsimony(case):
    g, _ = case
    if 'DOWNPASS_mask' not in g:
        pytest.skip('no parsimony outputs in this fixture')
    res = cpu_oracle.parsimony_fixed(g.flat, g.K, g['annot_mask'])
    for m in ('ACCTRAN', 'DOWNPASS', 'DELTRAN'):
        assert np.array_equal(res[m + '_mask'], g[m + '_mask']), m
        assert res[m + '_steps'] == int(g[m + '_steps']), m

"""GPU parity tests: the CUDA engine, called through its C ABI, against (i) the golden fixtures produced by the
unmodified reference and (ii) the CPU oracle on seeded synthetic inputs.

Tolerances (BASELINE.json north_star): log-likelihood 1e-10 relative, marginal probabilities 1e-9 absolute,
state assignments exact.
"""
import numpy as np
import pytest

from helpers import Golden, golden_names, engine_params

pytestmark = pytest.mark.gpu

LH_RTOL = 1e-10
PROB_ATOL = 1e-9


def _problem(g, **kw):
    from pastml_b200.engine import Problem
    return Problem(g.flat, g.K, g.kind, g['annot_mask'][None, :], g['has_state'][None, :], **kw)


def _run_ml(pb, params, force_joint):
    """The call sequence of ml_acr (ml.py:640-750) at fixed parameters on the engine."""
    from pastml_b200.engine import MASK_ANNOTATION, MASK_MAP, MASK_MPPA
    out = {}
    out['lh'] = pb.loglik(params, mask_kind=MASK_ANNOTATION)[0]
    out['lh_joint'], out['joint_state'] = pb.joint(params)
    out.update(pb.marginal(params, force_joint=force_joint, want_node_lh=True))
    out['lh_map'] = pb.loglik(params, mask_kind=MASK_MAP)[0]
    out['mppa_mask'], out['k_hist'] = pb.mppa_finalize(force_joint=force_joint)
    out['lh_mppa'] = pb.loglik(params, mask_kind=MASK_MPPA)[0]
    return out


@pytest.fixture(scope='module', params=golden_names())
def golden_run(request):
    g = Golden(request.param)
    pb = _problem(g)
    res = _run_ml(pb, engine_params(g), g.meta['force_joint'])
    res['pars'] = pb.parsimony()
    pb.close()
    return g, res


def test_golden_loglikelihoods(golden_run):
    g, r = golden_run
    assert r['lh'][0] == pytest.approx(float(g['log_likelihood']), rel=LH_RTOL)
    assert r['loglh'][0] == pytest.approx(float(g['log_likelihood']), rel=LH_RTOL)
    assert r['lh_joint'][0] == pytest.approx(float(g['lh_restricted_JOINT']), rel=LH_RTOL)
    assert r['lh_map'][0] == pytest.approx(float(g['lh_restricted_MAP']), rel=LH_RTOL)
    assert r['lh_mppa'][0] == pytest.approx(float(g['lh_restricted_MPPA']), rel=LH_RTOL)


def test_golden_marginal_probabilities(golden_run):
    g, r = golden_run
    np.testing.assert_allclose(r['probs'][:, 0, :], g['marginal_probabilities'], rtol=0, atol=PROB_ATOL)


def test_golden_node_likelihoods(golden_run):
    g, r = golden_run
    # log10(sum LIKELIHOOD) - LIKELIHOOD_SF of the reference, for nodes whose vector MAP/MPPA did not mask in place
    ref = np.log10(g['lh'].sum(axis=1)) - g['lh_sf']
    keep = g['has_initial'] == 0
    np.testing.assert_allclose(r['node_log10lh'][keep, 0], ref[keep], rtol=0, atol=1e-8)


def test_golden_states(golden_run):
    g, r = golden_run
    assert np.array_equal(np.uint64(1) << r['joint_state'][:, 0].astype(np.uint64), g['joint_mask'])
    assert np.array_equal(r['map_mask'][:, 0], g['map_mask'])
    assert np.array_equal(r['mppa_mask'][:, 0], g['mppa_mask'])
    num_scenarios = 1
    for k, cnt in enumerate(r['k_hist'][0]):
        num_scenarios *= k ** int(cnt)
    assert str(num_scenarios) == g.meta['num_scenarios']
    assert int(r['k_hist'][0][2:].sum()) == int(g['num_unresolved_nodes'])


def test_golden_parsimony(golden_run):
    g, r = golden_run
    if 'DOWNPASS_mask' not in g:
        pytest.skip('no parsimony outputs in this fixture')
    for m in ('ACCTRAN', 'DOWNPASS', 'DELTRAN'):
        mask, steps = r['pars'][m]
        assert np.array_equal(mask[:, 0], g[m + '_mask']), m
        assert int(steps[0]) == int(g[m + '_steps']), m


def test_columns_and_parameter_sets_batched():
    """Three Albania columns (F81, JC, EFT parameters) x 3 parameter sets in one problem."""
    from pastml_b200.engine import Problem
    gs = [Golden(n) for n in ('albania_f81', 'albania_jc', 'albania_eft')]
    g0 = gs[0]
    masks = np.stack([g['annot_mask'] for g in gs])
    has = np.stack([g['has_state'] for g in gs])
    pb = Problem(g0.flat, g0.K, 0, masks, has, max_param_sets=3)
    base = np.stack([engine_params(g) for g in gs])            # [C][stride]
    sets = np.stack([base, base, base])                          # [B][C][stride]
    sets[1, :, 0] *= 1.5
    sets[2, :, 0] *= 0.25
    lh = pb.loglik(sets, n_sets=3)
    from oracle import cpu_oracle
    for c, g in enumerate(gs):
        assert lh[0, c] == pytest.approx(float(g['log_likelihood']), rel=LH_RTOL)
        for b in (1, 2):
            m = g.oracle_model()
            m.sf = float(sets[b, c, 0])
            tree = cpu_oracle.OracleTree(g.flat)
            want = cpu_oracle.loglikelihood(tree, m, np.where(g['annot_mask'] == 0, np.uint64((1 << g.K) - 1),
                                                              g['annot_mask']), g['has_state'])
            assert lh[b, c] == pytest.approx(want, rel=LH_RTOL)
    res = pb.marginal(base, force_joint=False)
    for c, g in enumerate(gs):
        np.testing.assert_allclose(res['probs'][:, c, :], g['marginal_probabilities'], rtol=0, atol=PROB_ATOL)
    pb.close()


@pytest.mark.parametrize('n_tips,K,C,kind', [(3000, 10, 4, 'yule'), (2000, 20, 3, 'coalescent'), (1500, 3, 8, 'yule'),
                                            (1000, 30, 2, 'yule'), (2000, 20, 8, 'yule'), (600, 45, 2, 'yule'),
                                            (400, 62, 1, 'coalescent'), (800, 12, 20, 'yule'), (700, 5, 1, 'yule'),
                                            (900, 2, 16, 'coalescent'), (700, 10, 8, 'yule'), (600, 16, 8, 'coalescent'),
                                            (500, 5, 8, 'yule'), (400, 40, 8, 'yule'), (500, 6, 32, 'yule')])
def test_synthetic_against_oracle(n_tips, K, C, kind):
    from pastml_b200.engine import Problem, MASK_MAP, MASK_MPPA, pack_params
    from pastml_b200 import synth
    from oracle import cpu_oracle
    rng = np.random.default_rng(1000 + n_tips + K)
    flat = synth.yule_tree(n_tips, rng) if kind == 'yule' else synth.coalescent_tree(n_tips, rng)
    masks, has, pis, sfs = synth.simulate_tip_masks(flat, K, C, rng, missing=0.05)
    params = np.stack([pack_params(0, K, sfs[c], 0., 1., pis[c]) for c in range(C)])
    pb = Problem(flat, K, 0, masks, has)
    r = _run_ml(pb, params, True)
    pars = pb.parsimony()
    pb.close()
    for c in range(C):
        model = cpu_oracle.OracleModel(0, pis[c], sfs[c])
        want = cpu_oracle.ml_acr_fixed(flat, model, masks[c], has[c], method='MPPA', force_joint=True)
        assert r['lh'][c] == pytest.approx(want['log_likelihood'], rel=LH_RTOL)
        assert r['lh_joint'][c] == pytest.approx(want['log_likelihood_restricted_JOINT'], rel=LH_RTOL)
        assert r['lh_map'][c] == pytest.approx(want['log_likelihood_restricted_MAP'], rel=LH_RTOL)
        assert r['lh_mppa'][c] == pytest.approx(want['log_likelihood_restricted_MPPA'], rel=LH_RTOL)
        np.testing.assert_allclose(r['probs'][:, c, :], want['marginal_probabilities'], rtol=0, atol=PROB_ATOL)
        assert np.array_equal(r['joint_state'][:, c], want['joint_state'])
        assert np.array_equal(r['map_mask'][:, c], want['map_mask'])
        assert np.array_equal(r['mppa_mask'][:, c], want['mppa_mask'])
        wp = cpu_oracle.parsimony_fixed(flat, K, masks[c])
        for m in ('ACCTRAN', 'DOWNPASS', 'DELTRAN'):
            assert np.array_equal(pars[m][0][:, c], wp[m + '_mask']), m
            assert int(pars[m][1][c]) == wp[m + '_steps'], m


def test_zero_likelihood_is_reported():
    """Two zero-distance tips with different states under tau = 0 are merged by the alteration; with alter off the
    engine must report the reference's PastMLLikelihoodError condition (ml.py:139-145)."""
    from pastml_b200.engine import Problem, ZeroLikelihoodError, pack_params
    from pastml_b200.tree import FlatForest
    flat = FlatForest(np.array([-1, 0, 0], dtype=np.int32), np.array([0., 0., 0.]))
    masks = np.array([[0, 1, 2]], dtype=np.uint64)
    has = np.array([[0, 1, 1]], dtype=np.uint8)
    pb = Problem(flat, 2, 0, masks, has)
    params = pack_params(0, 2, 1., 0., 1., [0.5, 0.5])
    assert np.isfinite(pb.loglik(params, alter=True)[0, 0])
    with pytest.raises(ZeroLikelihoodError) as e:
        pb.loglik(params, alter=False)
    assert (e.value.parent, e.value.child) == (0, 2)
    pb.close()


@pytest.mark.parametrize('K,C,B', [(20, 8, 3), (10, 16, 5), (12, 8, 4), (6, 4, 3), (20, 8, 21), (7, 3, 4), (10, 16, 4), (5, 8, 20)])
def test_many_parameter_sets_tiled(K, C, B):
    """More than 16 lanes (parameter sets x columns): the cooperative bottom-up kernel runs one grid row per tile of 8
    lanes; every (set, column) must equal the single-set evaluation and the oracle."""
    from pastml_b200.engine import Problem, pack_params
    from pastml_b200 import synth
    from oracle import cpu_oracle
    rng = np.random.default_rng(50 + K + C)
    flat = synth.yule_tree(1500, rng)
    masks, has, pis, sfs = synth.simulate_tip_masks(flat, K, C, rng, missing=0.05)
    base = np.stack([pack_params(0, K, sfs[c], 0., 1., pis[c]) for c in range(C)])
    sets = np.stack([base] * B)
    for b in range(B):
        sets[b, :, 0] *= 1. + 0.17 * b
    pb = Problem(flat, K, 0, masks, has, max_param_sets=B)
    got = pb.loglik(sets, n_sets=B)
    for b in range(B):
        one = pb.loglik(sets[b], n_sets=1)[0]
        np.testing.assert_allclose(got[b], one, rtol=1e-13)
    pb.close()
    full = np.uint64((1 << K) - 1)
    tree = cpu_oracle.OracleTree(flat)
    for b, c in ((B - 1, 0), (1, C - 1)):
        model = cpu_oracle.OracleModel(0, pis[c], float(sets[b, c, 0]))
        want = cpu_oracle.loglikelihood(tree, model, np.where(masks[c] == 0, full, masks[c]), has[c])
        assert got[b, c] == pytest.approx(want, rel=LH_RTOL)


@pytest.mark.parametrize('n_tips,C,kind', [(1500, 8, 'yule'), (800, 16, 'coalescent')])
def test_jtt_columns_on_tensor_cores(n_tips, C, kind):
    """JTT on several columns (one shared eigen-decomposition, lanes a multiple of 8): the bottom-up sweep runs as
    FP64 tensor-core GEMMs (pml_mma.cu); log-likelihoods and marginal probabilities against the oracle, and two
    parameter sets in one pass."""
    from pastml_b200.annotation import ForestStats
    from pastml_b200.engine import Problem
    from pastml_b200.models.JTTModel import JTTModel
    from pastml_b200 import synth
    from oracle import cpu_oracle
    rng = np.random.default_rng(77 + n_tips)
    flat = synth.yule_tree(n_tips, rng) if kind == 'yule' else synth.coalescent_tree(n_tips, rng)
    K = 20
    masks, has, pis, sfs = synth.simulate_tip_masks(flat, K, C, rng, missing=0.05)
    fs = ForestStats(flat=flat)
    models = [JTTModel(forest_stats=fs, sf=float(sfs[c])) for c in range(C)]
    params = np.stack([m.engine_block() for m in models])
    pb = Problem(flat, K, models[0].engine_kind, masks, has, max_param_sets=2)
    lh = pb.loglik(params)[0]
    res = pb.marginal(params, force_joint=False)
    sets = np.stack([params, params])
    sets[1, :, 0] *= 1.3
    lh2 = pb.loglik(sets, n_sets=2)
    lh_joint, joint_state = pb.joint(params)   # max-product sweep: P formed on the tensor cores (bu_maxp_mma_kernel)
    assert 'bu_maxp_mma_kernel' in pb.last_kernels()['bottom_up']
    pb.close()
    full = np.uint64((1 << K) - 1)
    tree = cpu_oracle.OracleTree(flat)
    for c in range(C):
        m = models[c]
        kw = dict(d=np.real(m.D_DIAGONAL), A=np.real(m.A), Ainv=np.real(m.A_INV))
        om = cpu_oracle.OracleModel(2, m.frequencies, float(m.sf), **kw)
        allowed = np.where(masks[c] == 0, full, masks[c])
        want = cpu_oracle.loglikelihood(tree, om, allowed, has[c])
        assert lh[c] == pytest.approx(want, rel=LH_RTOL)
        assert lh2[0, c] == pytest.approx(want, rel=LH_RTOL)
        om2 = cpu_oracle.OracleModel(2, m.frequencies, float(m.sf) * 1.3, **kw)
        assert lh2[1, c] == pytest.approx(cpu_oracle.loglikelihood(tree, om2, allowed, has[c]), rel=LH_RTOL)
        if c in (0, C - 1):
            ref = cpu_oracle.ml_acr_fixed(flat, om, masks[c], has[c], method='MAP', force_joint=False)
            np.testing.assert_allclose(res['probs'][:, c, :], ref['marginal_probabilities'], rtol=0, atol=PROB_ATOL)
            assert np.array_equal(res['map_mask'][:, c], ref['map_mask'])
            refj = cpu_oracle.ml_acr_fixed(flat, om, masks[c], has[c], method='JOINT')
            assert np.array_equal(joint_state[:, c], refj['joint_state'])
            assert lh_joint[c] == pytest.approx(refj['log_likelihood_restricted_JOINT'], rel=LH_RTOL)


def test_lane_list_objective_equals_grid():
    """pml_loglik_lanes: an explicit list of (column, parameter block) lanes gives the values of the same lanes in the
    (sets x columns) grid; a zero-likelihood lane is reported for its column."""
    from pastml_b200.engine import Problem, pack_params
    from pastml_b200 import synth
    rng = np.random.default_rng(321)
    K, C, B = 10, 8, 3
    flat = synth.yule_tree(1200, rng)
    masks, has, pis, sfs = synth.simulate_tip_masks(flat, K, C, rng, missing=0.05)
    base = np.stack([pack_params(0, K, sfs[c], 0., 1., pis[c]) for c in range(C)])
    sets = np.stack([base] * B)
    for b in range(B):
        sets[b, :, 0] *= 1. + 0.21 * b
    pb = Problem(flat, K, 0, masks, has, max_param_sets=B)
    assert pb.supports_lane_lists()
    grid = pb.loglik(sets, n_sets=B)
    lanes = [(5, 0), (5, 1), (5, 2), (1, 2), (7, 0), (7, 1), (2, 1), (2, 1)]   # (column, set); 8 lanes
    got, bad = pb.loglik_lanes(np.stack([sets[b, c] for c, b in lanes]), [c for c, _ in lanes])
    assert not bad
    np.testing.assert_allclose(got, [grid[b, c] for c, b in lanes], rtol=1e-13)
    got4, _ = pb.loglik_lanes(np.stack([sets[0, 3]] * 4), [3] * 4)
    np.testing.assert_allclose(got4, grid[0, 3], rtol=1e-13)
    with pytest.raises(Exception):
        pb.loglik_lanes(np.stack([sets[0, 3]] * 3), [3] * 3)   # not a multiple of 4
    pb.close()


@pytest.mark.parametrize('n_tips,K,C', [(300, 40, 2), (200, 61, 1), (250, 28, 3)])
def test_custom_rates_with_many_states(n_tips, K, C):
    """Rate-matrix models beyond 32 states (a 61-state codon-sized matrix): per-column decompositions, every pass of the
    ML pipeline against the oracle.  (The 48- and 62-state instantiations of the eigen kernels run rolled loops.)"""
    from pastml_b200.engine import Problem, pack_params
    from pastml_b200.models.generator import get_diagonalisation
    from pastml_b200 import synth
    from oracle import cpu_oracle
    rng = np.random.default_rng(31 * K + n_tips)
    flat = synth.yule_tree(n_tips, rng)
    masks, has, pis, sfs = synth.simulate_tip_masks(flat, K, C, rng, missing=0.05)
    blocks, models = [], []
    for c in range(C):
        rates = rng.random((K, K)) + 0.05
        rates = rates + rates.T
        np.fill_diagonal(rates, 0.)
        d, A, Ainv = get_diagonalisation(pis[c], rates)
        d, A, Ainv = np.real(d), np.real(A), np.real(Ainv)
        blocks.append(pack_params(2, K, sfs[c], 0., 1., pis[c], d=d, A=A, Ainv=Ainv))
        models.append(cpu_oracle.OracleModel(2, pis[c], sfs[c], d=d, A=A, Ainv=Ainv))
    params = np.stack(blocks)
    pb = Problem(flat, K, 2, masks, has)
    r = _run_ml(pb, params, True)
    pb.close()
    for c in range(C):
        want = cpu_oracle.ml_acr_fixed(flat, models[c], masks[c], has[c], method='MPPA', force_joint=True)
        assert r['lh'][c] == pytest.approx(want['log_likelihood'], rel=LH_RTOL)
        assert r['lh_joint'][c] == pytest.approx(want['log_likelihood_restricted_JOINT'], rel=LH_RTOL)
        assert r['lh_mppa'][c] == pytest.approx(want['log_likelihood_restricted_MPPA'], rel=LH_RTOL)
        np.testing.assert_allclose(r['probs'][:, c, :], want['marginal_probabilities'], rtol=0, atol=PROB_ATOL)
        assert np.array_equal(r['joint_state'][:, c], want['joint_state'])
        assert np.array_equal(r['map_mask'][:, c], want['map_mask'])
        assert np.array_equal(r['mppa_mask'][:, c], want['mppa_mask'])

"""GPU parity AT SCALE: trees whose depth levels are far wider than what the persistent sweeps hold in one wave
(148 SMs x resident warps x nodes per warp, a few thousand nodes), so that every warp runs its software-pipelined
steady-state loop (staging-slot reuse, mbarrier phase flips, descriptor rotation, ready flags of the single-launch
sweeps) many times -- the regime the benchmark runs in and the small fixtures never reach.

Everything is compared with the CPU oracle (oracle/pastml_oracle.c, pinned on the reference's own outputs by
tests/test_oracle_golden.py) on the same seeded inputs: log-likelihoods 1e-10 relative, marginal probabilities 1e-9
absolute, JOINT / MAP / MPPA / parsimony state sets and k histograms exact.  The oracle runs one column per host thread.
Reference: /root/reference/pastml/ml.py:82-148 (bottom-up), 240-290 (top-down), 431-465 (marginals), 505-622
(MPPA / MAP / JOINT), /root/reference/pastml/parsimony.py:92-242, 360-380.
"""
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

LH_RTOL = 1e-10
PROB_ATOL = 1e-9


def _tree(kind, n_tips, rng):
    from pastml_b200 import synth
    return synth.yule_tree(n_tips, rng) if kind == 'yule' else synth.coalescent_tree(n_tips, rng)


def _widest_internal_level(flat):
    lp = np.asarray(flat.level_ptr)
    internal = np.asarray(flat.n_children) > 0
    ci = np.concatenate([[0], np.cumsum(internal)])
    return int((ci[lp[1:]] - ci[lp[:-1]]).max())


def _oracle_columns(fn, cols):
    with ThreadPoolExecutor(max_workers=min(8, len(cols))) as pool:
        return list(pool.map(fn, cols))


def _check_ml(r, want, c):
    assert r['lh'][c] == pytest.approx(want['log_likelihood'], rel=LH_RTOL)
    assert r['loglh'][c] == pytest.approx(want['log_likelihood'], rel=LH_RTOL)
    assert r['lh_joint'][c] == pytest.approx(want['log_likelihood_restricted_JOINT'], rel=LH_RTOL)
    assert r['lh_map'][c] == pytest.approx(want['log_likelihood_restricted_MAP'], rel=LH_RTOL)
    assert r['lh_mppa'][c] == pytest.approx(want['log_likelihood_restricted_MPPA'], rel=LH_RTOL)
    assert np.abs(r['probs'][:, c, :] - want['marginal_probabilities']).max() <= PROB_ATOL
    assert np.array_equal(r['joint_state'][:, c], want['joint_state'])
    assert np.array_equal(r['map_mask'][:, c], want['map_mask'])
    assert np.array_equal(r['mppa_mask'][:, c], want['mppa_mask'])
    assert np.array_equal(r['k_hist'][c], want['k_hist'])


def _run_ml(pb, params, force_joint=True):
    from pastml_b200.engine import MASK_ANNOTATION, MASK_MAP, MASK_MPPA
    out = {}
    out['lh'] = pb.loglik(params, mask_kind=MASK_ANNOTATION)[0]
    out['lh_joint'], out['joint_state'] = pb.joint(params)
    out.update(pb.marginal(params, force_joint=force_joint))
    out['lh_map'] = pb.loglik(params, mask_kind=MASK_MAP)[0]
    out['mppa_mask'], out['k_hist'] = pb.mppa_finalize(force_joint=force_joint)
    out['lh_mppa'] = pb.loglik(params, mask_kind=MASK_MPPA)[0]
    return out


@pytest.mark.parametrize('n_tips,K,C,kind,missing', [
    (250_000, 20, 8, 'yule', 0.),          # the benchmark's shape (TMA-staged / single-launch sweeps, 8 lanes)
    (200_000, 20, 8, 'coalescent', 0.05),  # ultrametric, many near-zero branches, unannotated tips
    (200_000, 10, 16, 'yule', 0.05),       # config 3's shape: 16 lanes
    (220_000, 4, 8, 'yule', 0.02),
    (200_000, 7, 3, 'coalescent', 0.05),   # no compile-time variant: the generic cooperative kernels
])
def test_f81_pipeline_at_scale(n_tips, K, C, kind, missing):
    """The whole fixed-parameter call sequence of ml_acr (objective, JOINT, marginal + MAP + MPPA, the restricted
    likelihoods) and the three parsimony methods, F81 family."""
    from pastml_b200.engine import Problem, pack_params
    from pastml_b200 import synth
    from oracle import cpu_oracle
    rng = np.random.default_rng(20_000 + n_tips + K)
    flat = _tree(kind, n_tips, rng)
    assert _widest_internal_level(flat) > 12_000   # more than one persistent wave of either sweep
    masks, has, pis, sfs = synth.simulate_tip_masks(flat, K, C, rng, missing=missing)
    params = np.stack([pack_params(0, K, sfs[c], 0., 1., pis[c]) for c in range(C)])
    pb = Problem(flat, K, 0, masks, has)
    r = _run_ml(pb, params)
    pars = pb.parsimony()
    pb.close()
    cols = sorted({0, C // 2, C - 1})

    def oracle(c):
        model = cpu_oracle.OracleModel(0, pis[c], sfs[c])
        return (cpu_oracle.ml_acr_fixed(flat, model, masks[c], has[c], method='MPPA', force_joint=True),
                cpu_oracle.parsimony_fixed(flat, K, masks[c]))

    for c, (want, wp) in zip(cols, _oracle_columns(oracle, cols)):
        _check_ml(r, want, c)
        for m in ('ACCTRAN', 'DOWNPASS', 'DELTRAN'):
            assert np.array_equal(pars[m][0][:, c], wp[m + '_mask']), m
            assert int(pars[m][1][c]) == wp[m + '_steps'], m


def test_f81_with_smoothing_at_scale():
    """tau > 0 (no zero-branch alteration, every branch lengthened) on a coalescent tree."""
    from pastml_b200.engine import Problem, pack_params
    from pastml_b200 import synth
    from oracle import cpu_oracle
    rng = np.random.default_rng(31)
    K, C = 12, 8
    flat = _tree('coalescent', 200_000, rng)
    masks, has, pis, sfs = synth.simulate_tip_masks(flat, K, C, rng, missing=0.05)
    tau = 0.3 * float(flat.dist[flat.dist > 0].mean())
    n, length = flat.n_nodes, float(flat.dist.sum())
    tauf = length / (length + tau * (n - 1))
    params = np.stack([pack_params(0, K, sfs[c], tau, tauf, pis[c]) for c in range(C)])
    pb = Problem(flat, K, 0, masks, has)
    r = _run_ml(pb, params)
    pb.close()
    cols = [0, C - 1]

    def oracle(c):
        return cpu_oracle.ml_acr_fixed(flat, cpu_oracle.OracleModel(0, pis[c], sfs[c], tau, tauf), masks[c], has[c])

    for c, want in zip(cols, _oracle_columns(oracle, cols)):
        _check_ml(r, want, c)


def test_objective_parameter_sets_at_scale():
    """The batched finite-difference objective (21 parameter sets x 8 columns on the tile-major layout, and a lane
    list) on a 200k-tip tree: every lane equals its single-set evaluation, two lanes are checked against the oracle."""
    from pastml_b200.engine import Problem, pack_params
    from pastml_b200 import synth
    from oracle import cpu_oracle
    rng = np.random.default_rng(77)
    K, C, B = 20, 8, 21
    flat = _tree('yule', 200_000, rng)
    masks, has, pis, sfs = synth.simulate_tip_masks(flat, K, C, rng, missing=0.05)
    base = np.stack([pack_params(0, K, sfs[c], 0., 1., pis[c]) for c in range(C)])
    sets = np.stack([base] * B)
    for b in range(B):
        sets[b, :, 0] *= 1. + 0.07 * b
    pb = Problem(flat, K, 0, masks, has, max_param_sets=B)
    got = pb.loglik(sets, n_sets=B)
    for b in (0, 7, B - 1):
        np.testing.assert_allclose(got[b], pb.loglik(sets[b], n_sets=1)[0], rtol=1e-13)
    lanes = [(c, b) for c in range(C) for b in (0, 3, 20, 11)]   # 32 lanes
    vals, bad = pb.loglik_lanes(np.stack([sets[b, c] for c, b in lanes]), [c for c, _ in lanes])
    assert not bad
    np.testing.assert_allclose(vals, [got[b, c] for c, b in lanes], rtol=1e-13)
    pb.close()
    full = np.uint64((1 << K) - 1)
    tree = cpu_oracle.OracleTree(flat)
    for b, c in ((B - 1, 0), (5, C - 1)):
        model = cpu_oracle.OracleModel(0, pis[c], float(sets[b, c, 0]))
        want = cpu_oracle.loglikelihood(tree, model, np.where(masks[c] == 0, full, masks[c]), has[c])
        assert got[b, c] == pytest.approx(want, rel=LH_RTOL)


def test_jtt_at_scale():
    """Config 4's model: JTT (K = 20, one shared eigen-decomposition, FP64 tensor-core sweeps) on a 200k-tip
    coalescent tree x 8 columns: objective, JOINT, marginal probabilities, MAP and MPPA states."""
    from pastml_b200.annotation import ForestStats
    from pastml_b200.engine import Problem
    from pastml_b200.models.JTTModel import JTTModel
    from pastml_b200 import synth
    from oracle import cpu_oracle
    rng = np.random.default_rng(404)
    K, C = 20, 8
    flat = _tree('coalescent', 200_000, rng)
    masks, has, pis, sfs = synth.simulate_tip_masks(flat, K, C, rng, missing=0.05)
    fs = ForestStats(flat=flat)
    models = [JTTModel(forest_stats=fs, sf=float(sfs[c])) for c in range(C)]
    params = np.stack([m.engine_block() for m in models])
    pb = Problem(flat, K, models[0].engine_kind, masks, has)
    r = _run_ml(pb, params)
    pb.close()
    cols = [0, C - 1]

    def oracle(c):
        m = models[c]
        om = cpu_oracle.OracleModel(2, m.frequencies, float(m.sf), d=np.real(m.D_DIAGONAL), A=np.real(m.A),
                                    Ainv=np.real(m.A_INV))
        return cpu_oracle.ml_acr_fixed(flat, om, masks[c], has[c], method='MPPA', force_joint=True)

    for c, want in zip(cols, _oracle_columns(oracle, cols)):
        _check_ml(r, want, c)


def test_hky_at_scale():
    """HKY (K = 4, kappa = 4) on a 250k-tip Yule tree x 8 columns."""
    from pastml_b200.engine import Problem, pack_params
    from pastml_b200 import synth
    from oracle import cpu_oracle
    rng = np.random.default_rng(505)
    K, C = 4, 8
    flat = _tree('yule', 250_000, rng)
    masks, has, pis, sfs = synth.simulate_tip_masks(flat, K, C, rng, missing=0.05)
    params = np.stack([pack_params(1, K, sfs[c], 0., 1., pis[c], kappa=4.) for c in range(C)])
    pb = Problem(flat, K, 1, masks, has)
    r = _run_ml(pb, params)
    pb.close()
    cols = [0, C - 1]

    def oracle(c):
        return cpu_oracle.ml_acr_fixed(flat, cpu_oracle.OracleModel(1, pis[c], sfs[c], kappa=4.), masks[c], has[c])

    for c, want in zip(cols, _oracle_columns(oracle, cols)):
        _check_ml(r, want, c)


def test_forest_of_trees_at_scale():
    """Config 5's shape scaled to the oracle's reach: a forest of 60 independent 5k-tip trees analysed jointly (one
    log-likelihood = the sum over the trees, ml.py:206-207), F81, MPPA + JOINT."""
    from pastml_b200.engine import Problem, pack_params
    from pastml_b200 import synth
    from oracle import cpu_oracle
    rng = np.random.default_rng(606)
    K, C = 10, 8
    flat = synth.forest([synth.yule_tree(5_000, rng) for _ in range(60)])
    masks, has, pis, sfs = synth.simulate_tip_masks(flat, K, C, rng, missing=0.05)
    params = np.stack([pack_params(0, K, sfs[c], 0., 1., pis[c]) for c in range(C)])
    pb = Problem(flat, K, 0, masks, has)
    r = _run_ml(pb, params)
    pb.close()
    cols = [0, C - 1]

    def oracle(c):
        return cpu_oracle.ml_acr_fixed(flat, cpu_oracle.OracleModel(0, pis[c], sfs[c]), masks[c], has[c])

    for c, want in zip(cols, _oracle_columns(oracle, cols)):
        _check_ml(r, want, c)

"""Subtree sharding of one tree: the plan (CPU) and, on a GPU, equality with the unsharded engine when the ranks are
simulated one after the other on one device."""
import numpy as np
import pytest


def _tree(n_tips, seed, kind='yule'):
    from pastml_b200 import synth
    rng = np.random.default_rng(seed)
    return (synth.yule_tree if kind == 'yule' else synth.coalescent_tree)(n_tips, rng), rng


@pytest.mark.parametrize('world', [2, 3, 8])
def test_plan_partitions_the_tree(world):
    from pastml_b200.sharded import SubtreePlan
    flat, _ = _tree(3000, 11)
    plan = SubtreePlan(flat, world)
    covered = np.concatenate([plan.nodes[r] for r in range(world)] + [plan.top_ids])
    # every node is in the top tree or in exactly one clade forest; the clade roots are in both
    counts = np.bincount(covered, minlength=flat.n_nodes)
    assert np.all(counts[plan.boundary] == 2)
    rest = np.ones(flat.n_nodes, dtype=bool)
    rest[plan.boundary] = False
    assert np.all(counts[rest] == 1)
    # balanced although the tree is not: no rank carries more than 1.3x the mean
    assert plan.load.max() <= 1.3 * plan.load.mean()
    top = plan.top_forest()
    assert top.n_trees == 1 and top.n_nodes == plan.n_top
    inner = top.parent >= 0
    assert np.array_equal(plan.top_ids[top.parent[inner]], flat.parent[plan.top_ids[inner]])
    assert np.all(top.n_children[plan.boundary_top] == 0)
    assert np.array_equal(plan.top_ids[plan.boundary_top], plan.boundary)
    for r in range(world):
        f = plan.clade_forest(r)
        ids = plan.nodes[r]
        assert f.n_trees == len(plan.roots_pos[r])
        assert np.array_equal(ids[:f.n_trees], plan.boundary[plan.roots_pos[r]])
        inner = f.parent >= 0
        assert np.array_equal(ids[f.parent[inner]], flat.parent[ids[inner]])
        assert np.array_equal(f.dist, flat.dist[ids])


@pytest.mark.gpu
@pytest.mark.parametrize('n_tips,K,C,world,kind', [(4000, 20, 8, 2, 'yule'), (3000, 7, 3, 3, 'coalescent'),
                                                  (2500, 12, 20, 4, 'yule')])
def test_sharded_equals_unsharded(n_tips, K, C, world, kind):
    from pastml_b200 import synth
    from pastml_b200.engine import Problem, pack_params
    from pastml_b200.sharded import SubtreePlan, ShardedProblem, merge_outputs
    flat, rng = _tree(n_tips, 100 + n_tips + K, kind)
    masks, has, pis, sfs = synth.simulate_tip_masks(flat, K, C, rng, missing=0.05)
    params = np.stack([pack_params(0, K, sfs[c], 0., 1., pis[c]) for c in range(C)])
    sets = np.stack([params, params])
    sets[1, :, 0] *= 1.3
    whole = Problem(flat, K, 0, masks, has, max_param_sets=2)
    want_lh = whole.loglik(sets, n_sets=2)
    want = whole.marginal(params, force_joint=False)
    whole.close()

    plan = SubtreePlan(flat, world, clades_per_rank=4)
    ranks = [ShardedProblem(plan, r, K, 0, masks, has, max_param_sets=2) for r in range(world)]
    # objective: every rank prunes its clades, the vectors are exchanged, every rank finishes the top tree
    gathered = [sp.local_root_partials(sets, n_sets=2) for sp in ranks]
    for sp in ranks:
        vec, ex = sp.assemble(gathered)
        got = sp.finish_loglik(sets, vec, ex, n_sets=2)
        np.testing.assert_allclose(got, want_lh, rtol=1e-12, atol=0)
    # fused sweep
    gathered = [sp.local_root_partials(params) for sp in ranks]
    pieces = []
    for sp in ranks:
        vec, ex = sp.assemble(gathered)
        pieces.append(sp.marginal(params, vec, ex))
    for key in ('probs', 'map_mask', 'mppa_mask'):
        got = merge_outputs(plan, pieces, key, flat.n_nodes)
        if key == 'probs':
            np.testing.assert_allclose(got, want[key], rtol=0, atol=1e-12)
        else:
            assert np.array_equal(got, want[key]), key
    np.testing.assert_allclose(pieces[0][0]['loglh'], want['loglh'], rtol=1e-12)
    for sp in ranks:
        sp.close()


@pytest.mark.gpu
@pytest.mark.parametrize('n_tips,K,C,world,kind', [(4000, 20, 8, 2, 'yule'), (3000, 7, 3, 3, 'coalescent'),
                                                  (60_000, 20, 8, 8, 'yule'), (2500, 12, 20, 4, 'yule')])
def test_device_shards_equal_unsharded(n_tips, K, C, world, kind):
    """The device-resident split (pml_shard_*: pack kernel -> gather buffer -> external tips of the top tree; top-down
    vectors handed to the clade roots on the device; TMA-staged clade sweeps) with the ranks simulated in one process:
    objective with two parameter sets, restricted likelihood on the MAP masks, and the whole marginal pipeline."""
    from pastml_b200 import synth
    from pastml_b200.engine import Problem, pack_params, MASK_MAP
    from pastml_b200.sharded import SubtreePlan, DeviceShard, exchange_local, merge_outputs
    flat, rng = _tree(n_tips, 300 + n_tips + K, kind)
    masks, has, pis, sfs = synth.simulate_tip_masks(flat, K, C, rng, missing=0.05)
    params = np.stack([pack_params(0, K, sfs[c], 0., 1., pis[c]) for c in range(C)])
    sets = np.stack([params, params])
    sets[1, :, 0] *= 1.3
    whole = Problem(flat, K, 0, masks, has, max_param_sets=2)
    want_lh = whole.loglik(sets, n_sets=2)
    want = whole.marginal(params, force_joint=False)
    want_map = whole.loglik(params, mask_kind=MASK_MAP)
    from pastml_b200.engine import MASK_JOINT
    want_lh_joint, want_joint_state = whole.joint(params)
    want_joint_restricted = whole.loglik(params, mask_kind=MASK_JOINT)
    want_forced = whole.marginal(params, force_joint=True)
    whole.close()

    plan = SubtreePlan(flat, world, clades_per_rank=4)
    ranks = [DeviceShard(plan, r, K, 0, masks, has, max_param_sets=2) for r in range(world)]
    for sh in ranks:
        sh.bottom_up(sets, n_sets=2)
    exchange_local(ranks)
    for sh in ranks:
        np.testing.assert_allclose(sh.finish(sets, n_sets=2), want_lh, rtol=1e-12, atol=0)
    for sh in ranks:
        sh.bottom_up(params)
    exchange_local(ranks)
    pieces = []
    for sh in ranks:
        lh = sh.finish(params)
        np.testing.assert_allclose(lh[0], want['loglh'], rtol=1e-12)
        sh.top_down()
        pieces.append((sh.results(0), sh.results(1)))
    for key in ('probs', 'map_mask', 'mppa_mask'):
        got = merge_outputs(plan, pieces, key, flat.n_nodes)
        if key == 'probs':
            np.testing.assert_allclose(got, want[key], rtol=0, atol=1e-12)
        else:
            assert np.array_equal(got, want[key]), key
    hist = sum(piece[1]['k_hist'] for piece in pieces) + pieces[0][0]['k_hist']
    # (the clade roots are counted twice: once as tips of the top tree, once as roots of their clade)
    assert int(hist.sum()) == (flat.n_nodes + len(plan.boundary)) * 1 * C
    # restricted likelihood: every problem applies its own MAP masks
    for sh in ranks:
        sh.bottom_up(params, mask_kind=MASK_MAP)
    exchange_local(ranks)
    for sh in ranks:
        np.testing.assert_allclose(sh.finish(params), want_map, rtol=1e-12)
    # JOINT: max-product sweeps on both sides of the cut, the clades' traceback starts from the top tree's choice
    for sh in ranks:
        sh.joint_bottom_up(params)
    exchange_local(ranks)
    joint_pieces = []
    for sh in ranks:
        np.testing.assert_allclose(sh.joint_finish(params), want_lh_joint, rtol=1e-12)
        joint_pieces.append(({'joint': sh.joint_states(0)}, {'joint': sh.joint_states(1)}))
    assert np.array_equal(merge_outputs(plan, joint_pieces, 'joint', flat.n_nodes), want_joint_state)
    for sh in ranks:
        sh.bottom_up(params, mask_kind=MASK_JOINT)
    exchange_local(ranks)
    for sh in ranks:
        np.testing.assert_allclose(sh.finish(params), want_joint_restricted, rtol=1e-12)
    # ... and MPPA with the joint state forced into every set (ml.py:545)
    for sh in ranks:
        sh.bottom_up(params)
    exchange_local(ranks)
    forced = []
    for sh in ranks:
        sh.finish(params)
        sh.top_down(force_joint=True)
        forced.append((sh.results(0), sh.results(1)))
    assert np.array_equal(merge_outputs(plan, forced, 'mppa_mask', flat.n_nodes), want_forced['mppa_mask'])
    for sh in ranks:
        sh.close()


@pytest.mark.gpu
def test_fit_on_a_sharded_tree_matches_the_plain_fit():
    """The lock-step fit driven by the sharded objective (one rank owning all clades: the exchange is the identity)
    gives what the plain ColumnBatch gives."""
    from pastml_b200 import synth
    from pastml_b200.annotation import ForestStats
    from pastml_b200.ml import ColumnBatch, fit_batch
    from pastml_b200.models.F81Model import F81Model
    from pastml_b200.sharded import SubtreePlan, ShardedColumnBatch
    flat, rng = _tree(3000, 77)
    K, C = 6, 3
    masks, has, pis, sfs = synth.simulate_tip_masks(flat, K, C, rng)
    fs = ForestStats(flat=flat)
    states = np.array(['s%d' % i for i in range(K)])
    obs = [np.full(K, 1. / K)] * C

    def models():
        return [F81Model(states=states, forest_stats=fs, character='c%d' % c) for c in range(C)]

    plain = ColumnBatch(flat, ['c0', 'c1', 'c2'], models(), annot_mask=masks, has_state=has)
    want, _ = fit_batch(plain, obs)
    plain.close()
    plan = SubtreePlan(flat, 1, clades_per_rank=6)
    ms = models()
    sharded = ShardedColumnBatch(plan, 0, ['c0', 'c1', 'c2'], ms, masks, has, gather=lambda vec, ex: [(vec, ex)])
    got, _ = fit_batch(sharded, obs)
    sharded.close()
    # the two objectives run on different kernels (rounding differs in the last bits), so the optimiser stops at
    # slightly different points: agreement to its own tolerance, not to the bit
    np.testing.assert_allclose(got, want, rtol=1e-6)

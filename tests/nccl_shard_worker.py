"""Worker of tests/test_gpu_nccl.py (one process per GPU under torchrun): a tree cut by subtree over the ranks, the
clade-root vectors all-gathered by NCCL on the engine's stream; every rank checks the log-likelihoods against the
unsharded engine, rank 0 also the merged marginal probabilities and state sets."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import torch.distributed as dist
    from pastml_b200 import synth
    from pastml_b200.engine import Problem, pack_params, MASK_MAP
    from pastml_b200.parallel import gather_objects
    from pastml_b200.sharded import SubtreePlan, DeviceShard, merge_outputs
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    rng = np.random.default_rng(2024)
    K, C = 20, 8
    flat = synth.yule_tree(30_000, rng)
    masks, has, pis, sfs = synth.simulate_tip_masks(flat, K, C, rng, missing=0.05)
    params = np.stack([pack_params(0, K, sfs[c], 0., 1., pis[c]) for c in range(C)])
    sets = np.stack([params, params])
    sets[1, :, 0] *= 1.3
    whole = Problem(flat, K, 0, masks, has, max_param_sets=2, device=local)
    want_lh = whole.loglik(sets, n_sets=2)
    want = whole.marginal(params, force_joint=False)
    want_map = whole.loglik(params, mask_kind=MASK_MAP)
    want_lh_joint, want_joint_state = whole.joint(params)
    whole.close()
    plan = SubtreePlan(flat, world, clades_per_rank=4)
    sh = DeviceShard(plan, rank, K, 0, masks, has, max_param_sets=2, device=local)
    sh.connect()
    np.testing.assert_allclose(sh.loglik(sets, n_sets=2), want_lh, rtol=1e-12, atol=0)
    loglh, top, mine = sh.marginal(params)
    np.testing.assert_allclose(loglh, want['loglh'], rtol=1e-12)
    np.testing.assert_allclose(sh.loglik(params, mask_kind=MASK_MAP), want_map, rtol=1e-12)
    lh_joint, top_states, my_states = sh.joint(params)   # max-product sweeps around the same all-gather
    np.testing.assert_allclose(lh_joint, want_lh_joint, rtol=1e-12)
    ms = sh.bench_step(params, 3)
    pieces = gather_objects((top, mine))
    joint_pieces = gather_objects(({'joint': top_states}, {'joint': my_states}))
    if rank == 0:
        assert np.array_equal(merge_outputs(plan, joint_pieces, 'joint', flat.n_nodes), want_joint_state)
        for key in ('probs', 'map_mask', 'mppa_mask'):
            got = merge_outputs(plan, pieces, key, flat.n_nodes)
            if key == 'probs':
                np.testing.assert_allclose(got, want[key], rtol=0, atol=1e-12)
            else:
                assert np.array_equal(got, want[key]), key
        print('NCCL_SHARD_OK world={} step_ms={:.3f}'.format(world, ms))
    sh.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()

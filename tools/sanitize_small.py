"""Development tool: a small tour of the engine for compute-sanitizer (memcheck / racecheck):

    compute-sanitizer --tool memcheck python tools/sanitize_small.py

F81 pipeline (TMA-staged sweeps, lane lists, JOINT, parsimony), scenario sampling, JTT on the tensor cores (bottom-up,
top-down, max-product), a 61-state rate matrix, HKY, and a tree cut by subtree with a local exchange."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry  # noqa: E402
from pastml_b200 import synth  # noqa: E402
from pastml_b200.annotation import ForestStats  # noqa: E402
from pastml_b200.engine import Problem, pack_params  # noqa: E402

entry.smoke()
rng = np.random.default_rng(3)

# scenario sampling, with zero-length branches (altered clusters)
flat = synth.coalescent_tree(300, rng)
flat.dist[rng.integers(1, flat.n_nodes, 30)] = 0.
flat = type(flat)(flat.parent, flat.dist)
masks, has, pis, sfs = synth.simulate_tip_masks(flat, 5, 3, rng, missing=0.1)
params = np.stack([pack_params(0, 5, sfs[c], 0., 1., pis[c]) for c in range(3)])
pb = Problem(flat, 5, 0, masks, has)
print('counts', pb.marginal_counts(params, 500, 7)[0].sum())
pb.close()

# JTT: tensor-core sweeps
from pastml_b200.models.JTTModel import JTTModel  # noqa: E402
flat = synth.yule_tree(500, rng)
masks, has, pis, sfs = synth.simulate_tip_masks(flat, 20, 8, rng, missing=0.05)
fs = ForestStats(flat=flat)
params = np.stack([JTTModel(forest_stats=fs, sf=float(sfs[c])).engine_block() for c in range(8)])
pb = Problem(flat, 20, 2, masks, has)
print('jtt', pb.loglik(params)[0][:2], pb.joint(params)[0][:2], pb.marginal(params, force_joint=True)['loglh'][:2])
pb.close()

# 61 states, per-column decomposition
from pastml_b200.models.generator import get_diagonalisation  # noqa: E402
K = 61
flat = synth.yule_tree(120, rng)
masks, has, pis, sfs = synth.simulate_tip_masks(flat, K, 1, rng)
rates = rng.random((K, K)) + .05
rates = rates + rates.T
np.fill_diagonal(rates, 0.)
d, A, Ainv = get_diagonalisation(pis[0], rates)
params = pack_params(2, K, sfs[0], 0., 1., pis[0], d=np.real(d), A=np.real(A), Ainv=np.real(Ainv))[None, :]
pb = Problem(flat, K, 2, masks, has)
print('k61', pb.loglik(params)[0], pb.joint(params)[0], pb.marginal(params)['loglh'])
pb.close()

# HKY
flat = synth.yule_tree(400, rng)
masks, has, pis, sfs = synth.simulate_tip_masks(flat, 4, 8, rng)
params = np.stack([pack_params(1, 4, sfs[c], 0., 1., pis[c], kappa=3.) for c in range(8)])
pb = Problem(flat, 4, 1, masks, has)
print('hky', pb.loglik(params)[0][:2], pb.marginal(params, force_joint=False)['loglh'][:2])
pb.close()

# one tree cut by subtree, ranks simulated in one process
from pastml_b200.sharded import SubtreePlan, DeviceShard, exchange_local  # noqa: E402
flat = synth.yule_tree(3000, rng)
masks, has, pis, sfs = synth.simulate_tip_masks(flat, 6, 4, rng)
params = np.stack([pack_params(0, 6, sfs[c], 0., 1., pis[c]) for c in range(4)])
plan = SubtreePlan(flat, 2, clades_per_rank=3)
ranks = [DeviceShard(plan, r, 6, 0, masks, has) for r in range(2)]
for sh in ranks:
    sh.joint_bottom_up(params)
exchange_local(ranks)
print('shard joint', [sh.joint_finish(params)[0] for sh in ranks])
for sh in ranks:
    sh.bottom_up(params)
exchange_local(ranks)
for sh in ranks:
    sh.finish(params)
    sh.top_down(force_joint=True)
for sh in ranks:
    sh.close()
print('SANITIZE_TOUR_DONE')

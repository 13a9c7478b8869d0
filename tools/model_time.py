"""Development tool: device time of the MPPA pipeline step for a model on a synthetic tree.
python tools/model_time.py MODEL [tips] [columns] [tree]   (MODEL: F81 | JTT | HKY)"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pastml_b200 import synth  # noqa: E402
from pastml_b200.annotation import ForestStats  # noqa: E402
from pastml_b200.engine import Problem, pack_params  # noqa: E402

model = sys.argv[1] if len(sys.argv) > 1 else 'JTT'
tips = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
C = int(sys.argv[3]) if len(sys.argv) > 3 else 8
kind_of_tree = sys.argv[4] if len(sys.argv) > 4 else 'coalescent'
rng = np.random.default_rng(4004)
flat = synth.yule_tree(tips, rng) if kind_of_tree == 'yule' else synth.coalescent_tree(tips, rng)
K = {'JTT': 20, 'HKY': 4}.get(model, 20)
masks, has, pis, sfs = synth.simulate_tip_masks(flat, K, C, rng)
if model == 'JTT':
    from pastml_b200.models.JTTModel import JTTModel
    fs = ForestStats(flat=flat)
    ms = [JTTModel(forest_stats=fs, sf=float(sfs[c])) for c in range(C)]
    params, kind = np.stack([m.engine_block() for m in ms]), ms[0].engine_kind
elif model == 'HKY':
    params, kind = np.stack([pack_params(1, K, sfs[c], 0., 1., pis[c], kappa=4.) for c in range(C)]), 1
else:
    params, kind = np.stack([pack_params(0, K, sfs[c], 0., 1., pis[c]) for c in range(C)]), 0
pb = Problem(flat, K, kind, masks, has)
pb.bench_pass(params, which=1, n_iter=3)
step, _, launches = pb.bench_pass(params, which=1, n_iter=10)
bu, _, _ = pb.bench_pass(params, which=0, n_iter=10)
td, _, _ = pb.bench_pass(params, which=3, n_iter=10)
pb.joint(params, want_states=False)
jt, _, _ = pb.bench_pass(params, which=2, n_iter=3)
print('model {} tree {} tips {} C {} levels {}: step_ms={:.3f} bu_ms={:.3f} td_ms={:.3f} joint_ms={:.3f} launches={}'.format(
    model, kind_of_tree, tips, C, flat.n_levels, step, bu, td, jt, launches))
pb.close()

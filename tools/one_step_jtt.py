"""Development tool: one marginal step of the JTT workload (K = 20, 8 columns; tensor-core sweeps) with direct launches,
for ncu captures of bu_mma_kernel / td_mma_kernel / td_select_kernel:  python tools/one_step_jtt.py [tips] [tree]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pastml_b200 import synth  # noqa: E402
from pastml_b200.annotation import ForestStats  # noqa: E402
from pastml_b200.engine import Problem  # noqa: E402
from pastml_b200.models.JTTModel import JTTModel  # noqa: E402

tips = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
kind = sys.argv[2] if len(sys.argv) > 2 else 'yule'
K, C = 20, 8
rng = np.random.default_rng(4004)
flat = synth.yule_tree(tips, rng) if kind == 'yule' else synth.coalescent_tree(tips, rng)
masks, has, pis, sfs = synth.simulate_tip_masks(flat, K, C, rng)
fs = ForestStats(flat=flat)
params = np.stack([JTTModel(forest_stats=fs, sf=float(sfs[c])).engine_block() for c in range(C)])
pb = Problem(flat, K, 2, masks, has)
avg_ms, min_ms, launches = pb.bench_pass(params, which=1, n_iter=1, flush_l2=False)
print('levels', flat.n_levels, 'launches_per_step', launches, 'avg_ms', avg_ms, pb.last_kernels())
pb.close()

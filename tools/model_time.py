"""Development tool: sweep timings for a model family on synthetic data.  python tools/model_time.py JTT|HKY|F81 tips C"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pastml_b200 import synth  # noqa: E402
from pastml_b200.annotation import ForestStats  # noqa: E402
from pastml_b200.engine import Problem  # noqa: E402
from pastml_b200.models import JTTModel, HKYModel, F81Model  # noqa: E402

name = sys.argv[1]
tips = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
C = int(sys.argv[3]) if len(sys.argv) > 3 else 8
rng = np.random.default_rng(5)
flat = synth.yule_tree(tips, rng)
fs = ForestStats(flat=flat)
K = {'JTT': 20, 'HKY': 4}.get(name, 10)
masks, has, pis, sfs = synth.simulate_tip_masks(flat, K, C, rng)
if name == 'JTT':
    models = [JTTModel(forest_stats=fs, sf=float(sfs[c])) for c in range(C)]
elif name == 'HKY':
    models = [HKYModel(forest_stats=fs, sf=float(sfs[c]), frequencies=pis[c], kappa=4.) for c in range(C)]
else:
    models = [F81Model(states=np.arange(K).astype(str), forest_stats=fs, sf=float(sfs[c]), frequencies=pis[c]) for c in range(C)]
params = np.stack([m.engine_block() for m in models])
pb = Problem(flat, K, models[0].engine_kind, masks, has)
for which, label in ((0, 'bottom-up'), (2, 'joint'), (1, 'bu+td')):
    avg, best, n = pb.bench_pass(params, which, 5)
    print('{} tips={} C={} K={}: {} {:.3f} ms ({} launches), {:.3g} node-states/s'.format(
        name, tips, C, K, label, avg, n, flat.n_nodes * K * C / (avg * 1e-3)))
pb.close()

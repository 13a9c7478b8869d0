"""Development tool: wall time of the full ML fit (two-stage L-BFGS-B, all columns in lock step) plus the
reconstruction passes on synthetic array-native inputs.

    python tools/fit_time.py [tips] [K] [C]
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pastml_b200 import synth  # noqa: E402
from pastml_b200.annotation import ForestStats  # noqa: E402
from pastml_b200.ml import ColumnBatch, fit_batch  # noqa: E402
from pastml_b200.models.F81Model import F81Model  # noqa: E402
from pastml_b200.engine import MASK_MAP, MASK_MPPA  # noqa: E402

tips = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
K = int(sys.argv[2]) if len(sys.argv) > 2 else 10
C = int(sys.argv[3]) if len(sys.argv) > 3 else 16
rng = np.random.default_rng(3000)
flat = synth.yule_tree(tips, rng)
masks, has, pis, sfs = synth.simulate_tip_masks(flat, K, C, rng)
fs = ForestStats(flat=flat)
states = np.array(['s{:02d}'.format(i) for i in range(K)])
obs = []
for c in range(C):
    tip_states = np.log2(masks[c][flat.is_tip].astype(np.float64)).astype(int)
    f = np.bincount(tip_states, minlength=K).astype(np.float64)
    obs.append(f / f.sum())
models = [F81Model(states=states, forest_stats=fs, character='c{}'.format(c)) for c in range(C)]
batch = ColumnBatch(flat, ['c{}'.format(c) for c in range(C)], models, annot_mask=masks, has_state=has)
t0 = time.perf_counter()
lhs, passes = fit_batch(batch, obs)
t1 = time.perf_counter()
params = batch.params()
pb = batch.problem
pb.joint(params, want_states=False)
pb.marginal(params, force_joint=True, want_probs=False, want_masks=False)
pb.loglik(params, mask_kind=MASK_MAP)
pb.mppa_finalize(force_joint=True, want_masks=False)
pb.loglik(params, mask_kind=MASK_MPPA)
t2 = time.perf_counter()
print('tips={} K={} C={}: fit {:.3f} s ({} lock-step passes, {:.3f} ms per pass, device {:.3f} s), reconstruction passes {:.3f} s; '
      'log-lh[0]={:.6f} sf[0]={:.6g} (simulated with {:.6g})'.format(tips, K, C, t1 - t0, passes, (t1 - t0) / passes * 1e3, getattr(batch, 'last_fit_device_ms', 0.) / 1e3,
                                                                    t2 - t1, lhs[0], models[0].sf, sfs[0]))
batch.close()

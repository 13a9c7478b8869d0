"""Development probe: device time of the marginal sweep with / without the probability and node-likelihood stores."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pastml_b200 import synth
from pastml_b200.engine import Problem, pack_params
rng = np.random.default_rng(4000)
K, C = 20, 8
flat = synth.yule_tree(1_000_000, rng)
masks, has, pis, sfs = synth.simulate_tip_masks(flat, K, C, rng)
params = np.stack([pack_params(0, K, sfs[c], 0., 1., pis[c]) for c in range(C)])
pb = Problem(flat, K, 0, masks, has)
pb.loglik(params)
for label, kw in (('no probs / no node lh', dict(results_on_device=False)), ('probs + node lh on device', dict(results_on_device=True))):
    ts = []
    for _ in range(6):
        pb.loglik(params)
        bu = pb.last_call_stats()[0]
        r = pb.marginal(params, force_joint=False, want_probs=False, want_masks=False, reuse_bottom_up=True, **kw)
        ts.append(pb.last_call_stats()[0])
    print('{}: TD sweep {:.3f} ms (BU {:.3f} ms)'.format(label, min(ts), bu))
print('k histogram column 0:', r['k_hist'][0][:6], 'of', flat.n_nodes)
pb.close()

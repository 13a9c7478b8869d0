"""Development probe: objective pass time on a clade forest vs the whole tree (one GPU)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pastml_b200 import synth
from pastml_b200.engine import Problem, pack_params
from pastml_b200.sharded import SubtreePlan
rng = np.random.default_rng(4000)
K, C, B = 20, 8, 21
flat = synth.yule_tree(1_000_000, rng)
masks, has, pis, sfs = synth.simulate_tip_masks(flat, K, C, rng)
params = np.stack([pack_params(0, K, sfs[c], 0., 1., pis[c]) for c in range(C)])
sets = np.stack([params] * B)
def timeit(pb, label, n):
    for ns in (1, B):
        pb.loglik(sets[:ns], n_sets=ns)
        ts = []
        for _ in range(4):
            t0 = time.perf_counter(); pb.loglik(sets[:ns], n_sets=ns); ts.append(time.perf_counter() - t0)
        print('{}: nodes {} levels {} n_sets {}: wall {:.3f} ms, device {:.3f} ms'.format(label, n, pb.flat.n_levels, ns, min(ts) * 1e3, pb.last_call_stats()[0]))
whole = Problem(flat, K, 0, masks, has, max_param_sets=B)
timeit(whole, 'whole tree', flat.n_nodes)
whole.close()
for cpr in (8, 2):
    plan = SubtreePlan(flat, 2, clades_per_rank=cpr)
    ids = plan.nodes[0]
    f = plan.clade_forest(0)
    pb = Problem(f, K, 0, masks[:, ids], has[:, ids], max_param_sets=B)
    timeit(pb, 'rank-0 forest (clades_per_rank={}, {} clades)'.format(cpr, len(plan.roots_pos[0])), len(ids))
    pb.close()

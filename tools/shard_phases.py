"""Development tool: the phases of ONE rank's share of the subtree-sharded step, on one GPU (the other ranks' clade
vectors come from a single in-process exchange).  python tools/shard_phases.py [world] [tips]"""
import os
import sys

import numpy as np

os.environ['PASTML_B200_SHARD_BENCH_LOCAL'] = '1'
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pastml_b200 import synth  # noqa: E402
from pastml_b200.engine import pack_params  # noqa: E402
from pastml_b200.sharded import SubtreePlan, DeviceShard, exchange_local  # noqa: E402

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
tips = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
K, C = 20, 8
rng = np.random.default_rng(4000)
flat = synth.yule_tree(tips, rng)
masks, has, pis, sfs = synth.simulate_tip_masks(flat, K, C, rng)
params = np.stack([pack_params(0, K, sfs[c], 0., 1., pis[c]) for c in range(C)])
plan = SubtreePlan(flat, world)
shards = [DeviceShard(plan, r, K, 0, masks, has) for r in range(world)]
for sh in shards:
    sh.bottom_up(params)
exchange_local(shards)
for sh in shards:
    sh.finish(params)
for r in (0, world - 1):
    sh = shards[r]
    sh.bench_step(params, 5)
    ms, ph = sh.bench_step(params, 10, phases=True)
    f = plan.clade_forest(r)
    print('rank', r, 'nodes', f.n_nodes, 'levels', f.n_levels, 'top levels', plan.top_forest().n_levels,
          'step_ms %.3f' % ms, ' '.join('%s=%.3f' % kv for kv in ph.items()))

"""Development tool: wall time of the array-native pipeline (newick file + annotation table in, the reference's result
files out) on a synthetic tree:  python tools/pipeline_flat_time.py [tips] [K] [columns]"""
import cProfile
import os
import pstats
import sys
import tempfile
import time

import numpy as np
import pandas as pd

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pastml_b200 import synth  # noqa: E402
from pastml_b200.arrays import pastml_pipeline_flat  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith('--')]
tips = int(args[0]) if len(args) > 0 else 100000
K = int(args[1]) if len(args) > 1 else 10
C = int(args[2]) if len(args) > 2 else 4
rng = np.random.default_rng(9)
flat = synth.yule_tree(tips, rng)
masks, has, pis, sfs = synth.simulate_tip_masks(flat, K, C, rng)
# newick text from the arrays (children of v are contiguous in level order)
names = np.array(['t%d' % i if flat.is_tip[i] else '' for i in range(flat.n_nodes)], dtype=object)
text = [None] * flat.n_nodes
for v in range(flat.n_nodes - 1, -1, -1):
    f, n = int(flat.first_child[v]), int(flat.n_children[v])
    inner = '(' + ','.join(text[f:f + n]) + ')' if n else ''
    text[v] = '{}{}:{:.6g}'.format(inner, names[v], flat.dist[v])
    for c in range(f, f + n):
        text[c] = None
work = tempfile.mkdtemp()
tree_file, table_file = os.path.join(work, 'tree.nwk'), os.path.join(work, 'data.tab')
with open(tree_file, 'w') as fh:
    fh.write(text[0] + ';')
tip_ids = np.flatnonzero(flat.is_tip)
data = {'col%d' % c: ['s%02d' % int(np.log2(float(masks[c][i]))) for i in tip_ids] for c in range(C)}
pd.DataFrame(data, index=[names[i] for i in tip_ids]).to_csv(table_file, sep='\t', index_label='id')
out = os.path.join(work, 'out')
pastml_pipeline_flat(tree_file, table_file, work_dir=os.path.join(work, 'warm'))   # CUDA context, kernels loaded
t0 = time.perf_counter()
flat2, results = pastml_pipeline_flat(tree_file, table_file, work_dir=out)
print('tips={} K={} columns={}: pastml_pipeline_flat {:.2f} s (parsing + fit + reconstruction + result files)'.format(
    tips, K, C, time.perf_counter() - t0))
if '--profile' in sys.argv:
    pr = cProfile.Profile()
    pr.enable()
    pastml_pipeline_flat(tree_file, table_file, work_dir=out)
    pr.disable()
    pstats.Stats(pr).sort_stats('cumtime').print_stats(16)

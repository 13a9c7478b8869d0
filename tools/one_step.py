"""Development tool: two steps of the MPPA pipeline on the bench workload with CUDA graphs off, for ncu launch lists:

    PASTML_B200_NO_GRAPH=1 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
        --clock-control none --csv --log-file gpurun_out/traffic.csv python tools/one_step.py
    python tools/traffic_summary.py gpurun_out/traffic.csv <launches per step> > profiles/<tag>_dram_traffic.json
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pastml_b200 import synth  # noqa: E402
from pastml_b200.engine import Problem, pack_params  # noqa: E402

tips = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
K = int(sys.argv[2]) if len(sys.argv) > 2 else 20
C = int(sys.argv[3]) if len(sys.argv) > 3 else 8
rng = np.random.default_rng(4000)
flat = synth.yule_tree(tips, rng)
masks, has, pis, sfs = synth.simulate_tip_masks(flat, K, C, rng)
params = np.stack([pack_params(0, K, sfs[c], 0., 1., pis[c]) for c in range(C)])
pb = Problem(flat, K, 0, masks, has)
avg_ms, min_ms, launches = pb.bench_pass(params, which=1, n_iter=2, flush_l2=False)
print('launches_per_step', launches, 'avg_ms', avg_ms)
pb.close()

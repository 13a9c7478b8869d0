"""Development tool: wall time of the drop-in acr() (fit + reconstruction + node write-back) on a golden fixture."""
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, 'tests'))
from helpers import Golden, annotated_forest  # noqa: E402
from pastml_b200.acr import acr  # noqa: E402

for name in sys.argv[1:] or ['albania_f81', 'hiv1c_loc_f81', 'denv2_location_f81']:
    g = Golden(name)
    col = g.meta['column']
    for attempt in range(2):
        forest = annotated_forest(g)
        t0 = time.perf_counter()
        res = acr(forest, columns=[col], column2states={col: np.array(g.states)}, prediction_method='MPPA',
                  model=g.model_name)[0]
        dt = time.perf_counter() - t0
    print('{}: N={} K={} acr() {:.3f} s, log-lh {:.6f} (reference {:.6f}, its run took {} s for MPPA+JOINT+MAP+3xMP), '
          'sf {:.6g}, unresolved {}'.format(name, g.flat.n_nodes, g.K, dt, res['log_likelihood'],
                                           float(g['log_likelihood']), g.meta['reference_seconds'], res['model'].sf,
                                           res['num_unresolved_nodes']))

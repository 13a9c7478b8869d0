"""TEST INFRASTRUCTURE ONLY -- golden transition counts from the reference's ``marginal_counts``
(``/root/reference/pastml/ml.py:753-862``).

The function samples, so the fixtures hold the mean and the spread of several reference runs: for every case the
tree and annotation of a committed array fixture are rebuilt, the reference model is set to the fixture's fitted
parameters and ``marginal_counts`` runs ``RUNS`` times with ``N_REPETITIONS`` repetitions each.  Output:
``tests/golden/counts/counts_<fixture>.npz`` (mean, std, runs, n_repetitions); ``tests/test_gpu_counts.py`` compares
``pastml_b200.ml.marginal_counts`` with them.

    python oracle/make_golden_counts.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, 'tests'))

CASES = ['albania_f81', 'synth_s4_n30_k3_f81_forest', 'sim152_hky_skew', 'synth_s2_n60_k7_f81']
RUNS, N_REPETITIONS = 8, 4000


def reference_model(g):
    from pastml.annotation import ForestStats
    from pastml.models.F81Model import F81Model
    from pastml.models.HKYModel import HKYModel
    from helpers import annotated_forest
    roots = annotated_forest(g)
    fs = ForestStats(roots)
    kw = dict(forest_stats=fs, sf=float(g['sf']), tau=float(g['tau']))
    if g.kind == 0:
        model = F81Model(states=np.array(g.states), frequencies=np.array(g['pi'], dtype=float), **kw)
    else:
        model = HKYModel(states=np.array(g.states), frequencies=np.array(g['pi'], dtype=float), kappa=float(g['kappa']), **kw)
    return roots, model


def main():
    from oracle import ref_env
    ref_env.install()
    from pastml.ml import marginal_counts
    from helpers import Golden
    for name in CASES:
        g = Golden(name)
        column = g.meta['column']
        runs = []
        for seed in range(RUNS):
            roots, model = reference_model(g)
            np.random.seed(100 + seed)
            runs.append(marginal_counts(roots, column, model, n_repetitions=N_REPETITIONS))
        runs = np.array(runs)
        out = os.path.join(REPO, 'tests', 'golden', 'counts', 'counts_{}.npz'.format(name))
        np.savez_compressed(out, mean=runs.mean(axis=0), std=runs.std(axis=0, ddof=1), runs=RUNS, n_repetitions=N_REPETITIONS)
        print(name, 'total transitions per repetition {:.3f} +- {:.3f}'.format(
            (runs.sum(axis=(1, 2)) - np.trace(runs, axis1=1, axis2=2)).mean(),
            (runs.sum(axis=(1, 2)) - np.trace(runs, axis1=1, axis2=2)).std()))


if __name__ == '__main__':
    main()

"""Records how closely the engine's fit lands on the reference's, fixture by fixture (GPU):

    python tools/optimiser_parity.py profiles/r2_optimiser_parity.json

For every optimised golden fixture the fit is run through pastml_b200.acr (same starting points, scipy L-BFGS-B
on the host, objective on the GPU) and compared with what the unmodified reference reached: |dlh|/|lh|,
|dsf|/sf, max |dpi|/pi, max |d marginal|, nodes whose MPPA state set differs -- next to the spread the reference
shows against ITSELF when its objective carries one ulp of noise (``ens_lh``/``ens_sf`` of the fixture).  The
asserts of tests/test_gpu_acr.py::test_through_the_optimiser are set from this record."""
import json
import os
import sys
import tempfile
import pathlib

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, 'tests'))

from helpers import Golden, golden_names, masks_from_feature  # noqa: E402
from test_gpu_acr import _acr  # noqa: E402


def record(name):
    g = Golden(name)
    col = g.meta['column']
    with tempfile.TemporaryDirectory() as tmp:
        flat, res = _acr(g, 'MPPA', pathlib.Path(tmp), None)
    r = res[0]
    mdl = r['model']
    ref_lh = float(g['log_likelihood'])
    ens_lh = np.asarray(g['ens_lh']) if 'ens_lh' in g else np.array([ref_lh])
    ens_sf = np.asarray(g['ens_sf']) if 'ens_sf' in g else np.array([float(g['sf'])])
    out = {
        'model': g.model_name, 'n_nodes': int(g.flat.n_nodes), 'K': g.K,
        'rel_dlh': abs(r['log_likelihood'] - ref_lh) / abs(ref_lh),
        'dlh': r['log_likelihood'] - ref_lh,
        'rel_dsf': abs(float(mdl.sf) - float(g['sf'])) / float(g['sf']),
        'reference_self_rel_dlh': float(np.abs(ens_lh - ref_lh).max() / abs(ref_lh)),
        'reference_self_rel_dsf': float(np.abs(ens_sf - float(g['sf'])).max() / float(g['sf'])),
        'max_d_marginal': float(np.abs(r['marginal_probabilities'].values
                                       - g['marginal_probabilities'][g.flat.ref_order]).max()),
        'mppa_nodes_differ': int((masks_from_feature(flat.nodes, g.states, col) != g['mppa_mask']).sum()),
    }
    if hasattr(mdl, 'frequencies') and g.model_name not in ('JC', 'JTT'):
        out['max_rel_dpi'] = float((np.abs(np.asarray(mdl.frequencies) - g['pi']) / np.maximum(g['pi'], 1e-12)).max())
    if g.model_name == 'HKY':
        out['rel_dkappa'] = abs(float(mdl.kappa) - float(g['kappa'])) / float(g['kappa'])
    return out


def main():
    names = [n for n in golden_names() if Golden(n).meta['optimised']]
    rows = {n: record(n) for n in names}
    for n, row in rows.items():
        print(n, json.dumps(row))
    if len(sys.argv) > 1:
        with open(sys.argv[1], 'w') as f:
            json.dump({'what': __doc__.split('\n\n')[1].replace('\n', ' '), 'fixtures': rows}, f, indent=1, sort_keys=True)


if __name__ == '__main__':
    main()

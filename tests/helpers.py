"""Shared helpers of the test-suite: golden fixture loading and oracle model construction."""
import glob
import json
import os

import numpy as np

from pastml_b200.tree import FlatForest

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')

MODEL_KIND = {'F81': 0, 'JC': 0, 'EFT': 0, 'HKY': 1, 'JTT': 2, 'CUSTOM_RATES': 2}


def golden_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, '*.npz')))


class Golden(object):
    def __init__(self, name):
        self.name = name
        self.z = np.load(os.path.join(GOLDEN_DIR, name + '.npz'), allow_pickle=False)
        self.meta = json.loads(str(self.z['meta']))
        self.flat = FlatForest(self.z['parent'], self.z['dist'], names=[str(_) for _ in self.z['names']])
        self.states = [str(_) for _ in self.z['states']]
        self.K = len(self.states)
        self.model_name = self.meta['model']
        self.kind = MODEL_KIND[self.model_name]

    def __getitem__(self, key):
        return self.z[key]

    def __contains__(self, key):
        return key in self.z.files

    def oracle_model(self):
        from oracle.cpu_oracle import OracleModel
        z = self.z
        kw = {}
        if self.kind == 2:
            kw = dict(d=z['eig_d'], A=z['eig_A'], Ainv=z['eig_Ainv'])
        return OracleModel(self.kind, z['pi'], float(z['sf']), float(z['tau']), float(z['tau_factor']),
                           float(z['kappa']), **kw)


def engine_params(g):
    """One lane's parameter block for the engine from a golden fixture."""
    from pastml_b200.engine import pack_params
    z = g.z
    kw = {}
    if g.kind == 2:
        kw = dict(d=z['eig_d'], A=z['eig_A'], Ainv=z['eig_Ainv'])
    if g.kind == 1:
        kw = dict(kappa=float(z['kappa']))
    return pack_params(g.kind, g.K, float(z['sf']), float(z['tau']), float(z['tau_factor']), z['pi'], **kw)


def annotated_forest(g, column=None):
    """Tree objects of a golden fixture with the annotation sets on the nodes, as acr() expects pre-annotated
    forests (columns + column2states call path)."""
    from pastml_b200.tree import forest_from_arrays
    column = column or g.meta['column']
    roots = forest_from_arrays(g['parent'], g['dist'], g.flat.names)
    nodes = []
    frontier = list(roots)
    while frontier:
        nodes.extend(frontier)
        frontier = [ch for n in frontier for ch in n.children]
    for node, m, h in zip(nodes, g['annot_mask'].tolist(), g['has_state'].tolist()):
        if h:
            node.add_feature(column, {g.states[k] for k in range(g.K) if (m >> k) & 1})
    return roots


def masks_from_feature(flat_nodes, states, feature):
    s2i = {s: i for i, s in enumerate(states)}
    out = []
    for node in flat_nodes:
        m = 0
        for s in getattr(node, feature):
            m |= 1 << s2i[s]
        out.append(m)
    import numpy as np
    return np.array(out, dtype=np.uint64)

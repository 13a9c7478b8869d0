"""TEST INFRASTRUCTURE ONLY -- golden outputs of the reference's ``acr(..., resolve_polytomies=True)``
(``/root/reference/pastml/acr.py:234-278``, ``tree.py:344-489``).

For every case a seeded tree with many multifurcations and clade-wise tip states is generated, the reference runs MPPA +
F81 with FIXED parameters (no optimiser in the comparison) and polytomy resolution on; the input arrays, the resolved
forest (level-ordered arrays, node names), the predicted state masks and the log-likelihood go to
``tests/golden/polytomies/<case>.npz``.

    python oracle/make_golden_polytomies.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, 'tests'))

CASES = {'poly_a': (11, 60), 'poly_b': (12, 90), 'poly_c': (13, 40), 'poly_d': (14, 120)}   # name: (seed, tips)
STATES = ['s0', 's1', 's2']
COLUMN = 'trait'
PARAMS = {'scaling_factor': 0.6, 's0': 0.5, 's1': 0.3, 's2': 0.2}


def random_polytomy_tree(seed, n_tips):
    """A tree with many multifurcations (2-7 children per node) whose tip states evolve down the tree with rare
    changes, so that groups of siblings share a state their parent does not have."""
    from pastml_b200.tree import Tree
    rng = np.random.default_rng(seed)
    root = Tree(name='root')
    root.dist = 0.
    root.add_feature('true_state', 0)
    leaves = [root]
    n_leaves = 1
    while n_leaves < n_tips:
        node = leaves.pop(int(rng.integers(len(leaves))))
        k = int(rng.integers(2, 8))
        for _ in range(k):
            child = node.add_child(dist=float(np.round(rng.exponential(0.3) + 0.01, 5)))
            state = node.true_state if rng.random() > 0.25 else int(rng.integers(3))
            child.add_feature('true_state', state)
            leaves.append(child)
        n_leaves += k - 1
    for i, leaf in enumerate(n for n in root.traverse() if not n.children):
        leaf.name = 'tip{}'.format(i)
        if rng.random() > 0.05:
            leaf.add_feature(COLUMN, {STATES[leaf.true_state]})
    for n in root.traverse():
        n.del_feature('true_state')
    return root


def main():
    from oracle import ref_env
    ref_env.install()
    from pastml.acr import acr
    from pastml.tree import annotate_dates, name_tree
    from pastml_b200.tree import flatten_forest
    out_dir = os.path.join(REPO, 'tests', 'golden', 'polytomies')
    os.makedirs(out_dir, exist_ok=True)
    for name, (seed, n_tips) in CASES.items():
        root = random_polytomy_tree(seed, n_tips)
        name_tree(root)
        annotate_dates([root])
        before = flatten_forest([root])
        annot = np.array([sum(1 << STATES.index(s) for s in getattr(n, COLUMN, set())) for n in before.nodes], dtype=np.uint64)
        in_parent, in_dist, in_names = before.parent.copy(), before.dist.copy(), np.array([n.name for n in before.nodes])
        res = acr([root], columns=[COLUMN], column2states={COLUMN: np.array(STATES)}, prediction_method='MPPA', model='F81',
                  column2parameters={COLUMN: dict(PARAMS)}, force_joint=True, resolve_polytomies=True, threads=1)[0]
        flat = flatten_forest([root])
        masks = np.array([sum(1 << STATES.index(s) for s in getattr(n, COLUMN)) for n in flat.nodes], dtype=np.uint64)
        np.savez_compressed(os.path.join(out_dir, name + '.npz'), in_parent=in_parent, in_dist=in_dist, in_names=in_names,
                            in_annot=annot, parent=flat.parent, dist=flat.dist, names=np.array([n.name for n in flat.nodes]),
                            mask=masks, log_likelihood=res['log_likelihood'])
        print(name, 'nodes', len(in_parent), '->', flat.n_nodes, 'lh', res['log_likelihood'],
              'polytomy nodes', sum('.polytomy_' in n.name for n in flat.nodes))


if __name__ == '__main__':
    main()

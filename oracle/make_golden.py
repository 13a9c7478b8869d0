"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by running the UNMODIFIED reference.

Run in the build container (needs /root/reference):  ``python oracle/make_golden.py [case ...]``

For every case the script builds the tree(s) with :class:`pastml_b200.tree.Tree` (registered as ``ete3.Tree``
through ``oracle/ref_env.py``), annotates them the way ``pastml.acr.acr`` expects pre-annotated forests
(``/root/reference/pastml/acr.py:118-135``), runs the reference ``acr()`` for MPPA (parameters optimised unless the
case fixes them), then for JOINT and MAP at the parameters MPPA found, then for the three parsimony methods, and
stores inputs (flat level-ordered arrays, annotation masks) and outputs (likelihoods, parameters, marginal
probabilities, per-node state sets as bit masks) in one compressed npz.  The fixtures travel to the GPU box; the
reference does not.
"""
import json
import os
import sys
import time

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path = [p for p in sys.path if os.path.abspath(p or '.') != HERE]
sys.path.insert(0, REPO)

from oracle import ref_env  # noqa: E402

ref_env.install()

from pastml.acr import acr  # noqa: E402
from pastml.annotation import ForestStats  # noqa: E402
from pastml.ml import MPPA, MAP, JOINT  # noqa: E402
from pastml.models.JTTModel import JTT_STATES, JTT_RATE_MATRIX, JTT_FREQUENCIES  # noqa: E402
from pastml.models.generator import save_matrix  # noqa: E402
from pastml.models.HKYModel import HKY_STATES  # noqa: E402
from pastml.parsimony import ACCTRAN, DOWNPASS, DELTRAN  # noqa: E402
from pastml_b200.tree import Tree, read_tree, flatten_forest, name_tree  # noqa: E402

GOLDEN_DIR = os.path.join(REPO, 'tests', 'golden')
REF = ref_env.REFERENCE_ROOT


# ----------------------------------------------------------------------------- helpers

def annotate(forest, node2states, column):
    """Puts the annotation sets on the nodes like preannotate_forest does (annotation.py:113-123):
    nodes listed in the table get a (possibly empty) set, the others lose the feature."""
    for root in forest:
        for node in root.traverse():
            if node.name in node2states:
                node.add_feature(column, set(node2states[node.name]))
            else:
                node.del_feature(column)


def masks_of(flat, states, feature):
    s2i = {s: i for i, s in enumerate(states)}
    mask = np.zeros(flat.n_nodes, dtype=np.uint64)
    has = np.zeros(flat.n_nodes, dtype=np.uint8)
    for i, node in enumerate(flat.nodes):
        value = getattr(node, feature, None)
        if value is not None and value != '':
            has[i] = 1
            m = 0
            for s in value:
                m |= 1 << s2i[s]
            mask[i] = m
    return mask, has


def run_case(name, make_forest, column, states, model, fixed_params=None, tau=0, rate_matrix=None,
             force_joint=True, do_parsimony=True, extra_meta=None):
    """make_forest() -> (list of roots, {node name: iterable of states}) -- called afresh for every reference run,
    because acr() leaves its results on the nodes."""
    t0 = time.time()
    states = np.array(sorted(states)) if model not in ('HKY', 'JTT') else \
        (HKY_STATES if model == 'HKY' else JTT_STATES)
    out = {}
    rate_file = None
    if rate_matrix is not None:
        rate_file = '/tmp/pastml_golden_rates_{}.txt'.format(name)
        save_matrix(states, rate_matrix, rate_file)

    def fresh():
        forest, node2states = make_forest()
        annotate(forest, node2states, column)
        return forest

    # inputs
    forest = fresh()
    flat = flatten_forest(forest)
    annot_mask, has_state = masks_of(flat, states, column)
    stats = ForestStats(forest)
    out.update(parent=flat.parent, dist=flat.dist, names=np.array(flat.names), annot_mask=annot_mask,
               has_state=has_state, states=np.array(states),
               forest_stats=np.array([stats.avg_nonzero_brlen, stats.num_nodes, stats.num_tips, stats.forest_length]))

    def call(method, params):
        forest = fresh()
        kwargs = dict(columns=[column], column2states={column: states}, prediction_method=method, model=model,
                      force_joint=force_joint, tau=tau, threads=1)
        if params is not None:
            kwargs['column2parameters'] = {column: params}
        if rate_file is not None:
            kwargs['column2rates'] = {column: rate_file}
        res = acr(forest, **kwargs)
        assert len(res) == 1
        return forest, res[0]

    # MPPA (optimised unless fixed)
    forest, res = call(MPPA, fixed_params)
    flat_r = flatten_forest(forest)
    assert list(flat_r.names) == list(flat.names)
    mdl = res['model']
    pi = np.array(getattr(mdl, 'frequencies', np.ones(len(states)) / len(states)), dtype=np.float64)
    out.update(sf=np.float64(mdl.sf), tau=np.float64(mdl.tau), tau_factor=np.float64(mdl._tau_factor), pi=pi,
               kappa=np.float64(getattr(mdl, 'kappa', 0.)),
               log_likelihood=np.float64(res['log_likelihood']),
               lh_restricted_JOINT=np.float64(res['log_likelihood_restricted_JOINT']),
               lh_restricted_MAP=np.float64(res['log_likelihood_restricted_MAP']),
               lh_restricted_MPPA=np.float64(res['log_likelihood_restricted_MPPA']),
               num_unresolved_nodes=np.int64(res['num_unresolved_nodes']),
               num_states_per_node_avg=np.float64(res['num_states_per_node_avg']))
    if hasattr(mdl, 'D_DIAGONAL'):
        out.update(eig_d=np.array(mdl.D_DIAGONAL), eig_A=np.array(mdl.A), eig_Ainv=np.array(mdl.A_INV),
                   rate_matrix=np.array(mdl.rate_matrix))
    mp = res['marginal_probabilities']
    assert list(mp.index) == [flat.names[i] for i in flat.ref_order]
    probs = np.zeros((flat.n_nodes, len(states)))
    probs[flat.ref_order] = mp.values
    out['marginal_probabilities'] = probs
    out['mppa_mask'] = masks_of(flat_r, states, column)[0]
    out['lh'] = np.array([getattr(n, column + '_LIKELIHOOD') for n in flat_r.nodes])
    out['lh_sf'] = np.array([getattr(n, column + '_LIKELIHOOD_SF') for n in flat_r.nodes], dtype=np.float64)
    out['has_initial'] = np.array([hasattr(n, column + '_ALLOWED_STATES.initial') for n in flat_r.nodes], dtype=np.uint8)
    num_scenarios = int(res['num_scenarios'])

    # JOINT and MAP at the same parameters
    params = {'scaling_factor': float(mdl.sf), 'smoothing_factor': float(mdl.tau)}
    if model in ('F81', 'HKY', 'CUSTOM_RATES'):
        params.update({s: float(f) for s, f in zip(states, pi)})
    if model == 'HKY':
        params['kappa'] = float(mdl.kappa)
    forest, resj = call(JOINT, params)
    assert float(resj['model'].sf) == float(mdl.sf)
    out['joint_mask'] = masks_of(flatten_forest(forest), states, column)[0]
    out['lh_fixed'] = np.float64(resj['log_likelihood'])
    forest, resm = call(MAP, params)
    out['map_mask'] = masks_of(flatten_forest(forest), states, column)[0]

    if do_parsimony:
        for method in (ACCTRAN, DOWNPASS, DELTRAN):
            forest = fresh()
            resp = acr(forest, columns=[column], column2states={column: states}, prediction_method=method,
                       threads=1)[0]
            out[method + '_mask'] = masks_of(flatten_forest(forest), states, column)[0]
            out[method + '_steps'] = np.int64(resp['steps'])

    meta = dict(name=name, column=column, model=model, force_joint=bool(force_joint),
                optimised=fixed_params is None, num_scenarios=str(num_scenarios),
                reference_seconds=round(time.time() - t0, 2))
    meta.update(extra_meta or {})
    out['meta'] = np.array(json.dumps(meta))
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    path = os.path.join(GOLDEN_DIR, name + '.npz')
    np.savez_compressed(path, **out)
    print('{}: N={} K={} lh={:.12f} sf={:.9g} unresolved={} ({:.1f}s) -> {}'
          .format(name, flat.n_nodes, len(states), float(out['log_likelihood']), float(out['sf']),
                  int(out['num_unresolved_nodes']), time.time() - t0, os.path.relpath(path, REPO)))


# ----------------------------------------------------------------------------- cases

def table_forest(tree_path, table_path, column, sep=',', index_col=0, do_name=False):
    df = pd.read_csv(table_path, index_col=index_col, header=0, sep=sep)[[column]]
    df.index = df.index.map(str)
    node2states = {}
    for idx, value in df[column].items():
        node2states.setdefault(idx, set())
        if not pd.isna(value) and value != '':
            node2states[idx].add(str(value))
    states = sorted({s for v in node2states.values() for s in v})

    def make():
        with open(tree_path) as f:
            nwks = f.read().replace('\n', '').split(';')
        roots = [read_tree(nwk + ';') for nwk in nwks[:-1]]
        for i, root in enumerate(roots):
            for n in root.traverse():
                if n.dist < 0:
                    n.dist = 0
            if do_name:
                name_tree(root, suffix='' if len(roots) == 1 else '_{}'.format(i))
        return roots, node2states

    return make, states


def random_tree(rng, n_tips, polytomy_p=0.15, zero_p=0.2, names_prefix='', internal_annot=False):
    """A random rooted tree with some polytomies and zero-length branches (internal and external)."""
    root = Tree()
    root.dist = 0.
    leaves = [root]
    while len(leaves) < n_tips:
        node = leaves.pop(rng.integers(len(leaves)))
        k = 2 if rng.random() > polytomy_p else int(rng.integers(3, 6))
        for _ in range(k):
            leaves.append(node.add_child(dist=0. if rng.random() < zero_p else float(np.round(rng.exponential(0.1), 5))))
    for i, n in enumerate(root.traverse('preorder')):
        n.name = '{}{}{}'.format(names_prefix, 't' if n.is_leaf() else 'n', i)
    return root


def synthetic_case(seed, n_tips, K, n_trees=1, missing_p=0.1, multi_p=0.1, internal_p=0.05):
    rng = np.random.default_rng(seed)
    states = ['s{:02d}'.format(i) for i in range(K)]
    newicks = []
    node2states = {}
    for t in range(n_trees):
        root = random_tree(rng, n_tips, names_prefix='T{}_'.format(t))
        # states evolve down the tree with a sticky random walk so that clades are informative
        cur = {root: int(rng.integers(K))}
        for n in root.traverse('preorder'):
            if n.up is not None:
                cur[n] = cur[n.up] if (n.dist == 0 or rng.random() < 0.8) else int(rng.integers(K))
            if n.is_leaf():
                r = rng.random()
                if r < missing_p:
                    if rng.random() < 0.5:
                        node2states[n.name] = set()   # listed in the table with a missing value
                elif r < missing_p + multi_p:
                    node2states[n.name] = {states[cur[n]], states[int(rng.integers(K))]}
                else:
                    node2states[n.name] = {states[cur[n]]}
            elif rng.random() < internal_p:
                node2states[n.name] = {states[cur[n]]}
        newicks.append(root.write(format=1, format_root_node=True, dist_formatter='%.17g'))

    def make():
        return [Tree(nwk, format=1) for nwk in newicks], node2states

    return make, states


def case_albania(model):
    make, states = table_forest(os.path.join(REF, 'tests/data/Albanian.tree.152tax.tre'),
                                os.path.join(REF, 'tests/data/data.txt'), 'Country')
    run_case('albania_{}'.format(model.lower()), make, 'Country', states, model)


def case_albania_tau():
    make, states = table_forest(os.path.join(REF, 'tests/data/Albanian.tree.152tax.tre'),
                                os.path.join(REF, 'tests/data/data.txt'), 'Country')
    pi = [0.1, 0.05, 0.15, 0.3, 0.4]
    params = {'scaling_factor': 3.5, 'smoothing_factor': 0.004}
    params.update(dict(zip(states, pi)))
    run_case('albania_f81_tau', make, 'Country', states, 'F81', fixed_params=params, do_parsimony=False)


def case_denv2(column):
    make, states = table_forest(os.path.join(REF, 'examples/DENV2/data/real/raxml_tree.dated.nwk'),
                                os.path.join(REF, 'examples/DENV2/data/real/metadata.tab'), column, sep='\t',
                                index_col=0, do_name=True)
    run_case('denv2_{}_f81'.format(column.lower()), make, column, states, 'F81')


def case_hiv1c_loc():
    tree = os.path.join(REF, 'examples/HIV1C/data/best/pastml_phyml_tree.nwk')
    make, states = table_forest(tree, os.path.join(REF, 'examples/HIV1C/data/metadata.tab'), 'Loc', sep='\t')
    pfile = os.path.join(REF, 'examples/HIV1C/data/pastml_params/params.tree_pastml_phyml_tree.loc_Loc.tab')
    pdf = pd.read_csv(pfile, sep='\t', index_col=0)['value']
    params = {'scaling_factor': float(pdf['scaling_factor'])}
    params.update({s: float(pdf[s]) for s in states})
    run_case('hiv1c_loc_f81', make, 'Loc', states, 'F81', fixed_params=params)


def case_hiv1c_drm(column, model, sf):
    tree = os.path.join(REF, 'examples/HIV1C/data/best/pastml_phyml_tree.nwk')
    make, states = table_forest(tree, os.path.join(REF, 'examples/HIV1C/data/metadata.tab'), column, sep='\t')
    run_case('hiv1c_{}_{}'.format(column.replace(':', '_'), model.lower()), make, column, states, model,
             fixed_params={'scaling_factor': sf}, do_parsimony=True)


def case_hky(freq_tag):
    base = os.path.join(REF, 'tests/data/tree.152taxa.sf_0.5.{}'.format(freq_tag))
    df = pd.read_csv(base + '.pastml.tab', sep='\t', index_col=0)
    tree0 = read_tree(base + '.nwk')
    tips = {t.name for t in tree0}
    node2states = {str(k): {v} for k, v in df['Real'].items() if str(k) in tips and not pd.isna(v)}

    def make():
        return [read_tree(base + '.nwk')], node2states

    run_case('sim152_hky_{}'.format('eq' if 'A_0.25' in freq_tag else 'skew'), make, 'nt', list('ACGT'), 'HKY')


def case_aa(model):
    rng = np.random.default_rng(20)
    tree0 = random_tree(rng, 60, polytomy_p=0.05, zero_p=0.05)
    nwk = tree0.write(format=1, format_root_node=True, dist_formatter='%.17g')
    aa = list(JTT_STATES)
    cur = {}
    node2states = {}
    for n in tree0.traverse('preorder'):
        cur[n] = int(rng.integers(20)) if n.up is None else \
            (cur[n.up] if rng.random() < 0.7 else int(rng.integers(20)))
        if n.is_leaf():
            node2states[n.name] = {aa[cur[n]]}

    def make():
        return [Tree(nwk, format=1)], node2states

    if model == 'JTT':
        run_case('aa60_jtt', make, 'aa', aa, 'JTT')
    else:
        # CUSTOM_RATES with the JTT matrix, frequencies optimised from the JTT ones would take long: fix them
        params = {'scaling_factor': 2.0}
        params.update({s: float(f) for s, f in zip(JTT_STATES, JTT_FREQUENCIES)})
        run_case('aa60_custom', make, 'aa', aa, 'CUSTOM_RATES', fixed_params=params, rate_matrix=JTT_RATE_MATRIX)


def case_synth(seed, n_tips, K, model, n_trees=1, force_joint=True, tau=0.):
    make, states = synthetic_case(seed, n_tips, K, n_trees=n_trees)
    rng = np.random.default_rng(seed + 7)
    pi = rng.dirichlet(np.ones(K) * 2.) if model == 'F81' else np.ones(K) / K
    params = {'scaling_factor': float(rng.uniform(2., 8.)), 'smoothing_factor': tau}
    if model == 'F81':
        params.update(dict(zip(states, pi.tolist())))
    run_case('synth_s{}_n{}_k{}_{}{}{}'.format(seed, n_tips, K, model.lower(), '' if n_trees == 1 else '_forest',
                                              '' if force_joint else '_nofj'),
             make, 'c', states, model, fixed_params=params, force_joint=force_joint, tau=tau)


def add_noise_ensemble(name, seeds=tuple(range(1, 13)), noise=1e-15):
    """For a fixture whose parameters were optimised: re-runs the reference with its own objective perturbed by a
    relative noise of 1e-15 (about one ulp) and stores the optima it then reaches (``ens_lh``, ``ens_sf``).
    scipy's L-BFGS-B differentiates the objective by forward differences with h = 1e-8, so on flat likelihood
    surfaces the reference's optimum is not reproducible beyond bit-identical arithmetic; the spread measured
    here is what "parity through the optimiser" can mean for that data set."""
    import pastml.ml as rml
    path = os.path.join(GOLDEN_DIR, name + '.npz')
    z = dict(np.load(path, allow_pickle=False))
    meta = json.loads(str(z['meta']))
    if not meta['optimised']:
        return
    sys.path.insert(0, os.path.join(REPO, 'tests'))
    from helpers import Golden, annotated_forest
    g = Golden(name)
    col = meta['column']
    orig = rml.get_bottom_up_loglikelihood
    lhs, sfs = [], []
    try:
        for seed in seeds:
            rng = np.random.default_rng(seed)
            rml.get_bottom_up_loglikelihood = lambda *a, **k: orig(*a, **k) * (1 + noise * rng.standard_normal())
            forest = annotated_forest(g)
            r = acr(forest, columns=[col], column2states={col: np.array(g.states)}, prediction_method=MPPA,
                    model=meta['model'], force_joint=meta['force_joint'], threads=1)[0]
            lhs.append(float(r['log_likelihood']))
            sfs.append(float(r['model'].sf))
    finally:
        rml.get_bottom_up_loglikelihood = orig
    z['ens_lh'] = np.array(lhs)
    z['ens_sf'] = np.array(sfs)
    np.savez_compressed(path, **z)
    print('{}: golden lh {:.9f}; with 1e-15 objective noise the reference reaches {}'.format(
        name, float(z['log_likelihood']), ', '.join('{:.9f}'.format(v) for v in lhs)))


CASES = {
    'albania_f81': lambda: case_albania('F81'),
    'albania_jc': lambda: case_albania('JC'),
    'albania_eft': lambda: case_albania('EFT'),
    'albania_f81_tau': case_albania_tau,
    'denv2_location': lambda: case_denv2('Location'),
    'denv2_genotype': lambda: case_denv2('genotype'),
    'hiv1c_loc': case_hiv1c_loc,
    'hiv1c_drm_jc': lambda: case_hiv1c_drm('RT:M184V', 'JC', 4.0),
    'hky_eq': lambda: case_hky('A_0.25.C_0.25.G_0.25.T_0.25'),
    'hky_skew': lambda: case_hky('A_0.6.C_0.15.G_0.2.T_0.05'),
    'aa_jtt': lambda: case_aa('JTT'),
    'aa_custom': lambda: case_aa('CUSTOM_RATES'),
    'synth1': lambda: case_synth(1, 40, 4, 'F81'),
    'synth2': lambda: case_synth(2, 60, 7, 'F81'),
    'synth3': lambda: case_synth(3, 50, 9, 'JC', force_joint=False),
    'synth4': lambda: case_synth(4, 30, 3, 'F81', n_trees=3),
    'synth5': lambda: case_synth(5, 80, 12, 'F81'),
    'synth6': lambda: case_synth(6, 45, 5, 'F81', tau=0.01),
    'synth7': lambda: case_synth(7, 70, 2, 'JC'),
}

if __name__ == '__main__':
    import warnings
    warnings.filterwarnings('ignore')
    if sys.argv[1:2] == ['--ensemble']:
        import glob
        names = sys.argv[2:] or sorted(os.path.splitext(os.path.basename(p))[0]
                                       for p in glob.glob(os.path.join(GOLDEN_DIR, '*.npz')))
        for name in names:
            add_noise_ensemble(name)
        sys.exit(0)
    which = sys.argv[1:] or list(CASES)
    for key in which:
        CASES[key]()

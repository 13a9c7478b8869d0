"""CPU-only checks (no CUDA device needed): the C ABI library loads and exports every symbol the header declares,
host-side logic (tree flattening, forest statistics, models, fast division) behaves, and the engine refuses to
run without a GPU instead of falling back."""
import os
import re

import numpy as np
import pandas as pd
import pytest

from helpers import Golden, golden_names, annotated_forest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from pastml_b200 import _lib
    header = open(os.path.join(REPO, 'include', 'pastml_b200.h')).read()
    declared = set(re.findall(r'\b(pml_[a-z0-9_]+)\s*\(', header))
    assert declared == set(_lib.EXPORTS)
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.pml_version() >= 100
    assert lib.pml_param_stride(0, 5) == 9 and lib.pml_param_stride(1, 4) == 8 and lib.pml_param_stride(2, 20) == 844


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    from pastml_b200.engine import Problem, EngineError
    g = Golden('synth_s1_n40_k4_f81')
    with pytest.raises(EngineError):
        Problem(g.flat, g.K, 0, g['annot_mask'][None, :], g['has_state'][None, :])


@pytest.mark.parametrize('name', golden_names())
def test_flatten_round_trip_and_stats(name):
    from pastml_b200.tree import flatten_forest
    from pastml_b200.annotation import ForestStats
    g = Golden(name)
    forest = annotated_forest(g)
    flat = flatten_forest(forest)
    assert np.array_equal(flat.parent, g['parent']) and np.array_equal(flat.dist, g['dist'])
    assert flat.names == g.flat.names
    fs = ForestStats(flat=flat)
    ref = g['forest_stats']
    assert (fs.avg_nonzero_brlen, fs.num_nodes, fs.num_tips, fs.forest_length) == tuple(ref)
    # level order invariants the kernels rely on
    assert np.all(np.diff(flat.parent) >= 0)
    kids = flat.first_child[flat.int_nodes] + flat.n_children[flat.int_nodes]
    assert np.all(kids[:-1] == flat.first_child[flat.int_nodes][1:])
    assert flat.level_ptr[-1] == flat.n_nodes and flat.int_level_ptr[-1] == flat.n_internal


@pytest.mark.parametrize('name', golden_names())
def test_models_reproduce_reference_pij(name):
    """get_Pij_t of the host model classes against the oracle's restatement at the fixture's parameters."""
    from pastml_b200.models import model2class
    from pastml_b200.annotation import ForestStats
    g = Golden(name)
    fs = ForestStats(flat=g.flat)
    params = {'scaling_factor': float(g['sf']), 'smoothing_factor': float(g['tau'])}
    if g.model_name in ('F81', 'HKY', 'CUSTOM_RATES'):
        params.update({s: float(f) for s, f in zip(g.states, g['pi'])})
    if g.model_name == 'HKY':
        params['kappa'] = float(g['kappa'])
    kw = {}
    if g.model_name == 'CUSTOM_RATES':
        kw = dict(rate_matrix=g['rate_matrix'])
    model = model2class[g.model_name](states=np.array(g.states), forest_stats=fs, parameter_file=params,
                                      observed_frequencies=g['pi'], tau=float(g['tau']), **kw)
    assert model.get_num_params() == 0
    om = g.oracle_model()
    for t in (0., 1e-3, 0.05, 1.7):
        np.testing.assert_allclose(model.get_Pij_t(t), om.pij(t), rtol=1e-12, atol=1e-15)
    block = model.engine_block()
    assert block[0] == float(g['sf']) and block[2] == pytest.approx(float(g['tau_factor']), rel=1e-15)


def test_newick_round_trip_and_naming():
    from pastml_b200.tree import Tree, name_tree, collapse_zero_branches
    t = Tree('((A:1,B:0.5):0.25,(C:1,(D:1,E:2):0):0.5,F:3);', format=1)
    name_tree(t)
    assert t.name == 'root' and len({n.name for n in t.traverse()}) == 10
    # the reference's naming rule (tree.py:76-105): a preorder counter that only advances on name clashes
    assert [n.name for n in t.traverse('postorder')] == ['A', 'B', 'n0', 'C', 'D', 'E', 'n1', 'n00', 'F', 'root']
    again = Tree(t.write(format=1, format_root_node=True), format=1)
    assert [n.name for n in again.traverse()] == [n.name for n in t.traverse()]
    collapse_zero_branches([t])
    assert sum(1 for _ in t.traverse()) == 9


def test_padded_sets():
    """Multi-set passes are padded to lane tiles of 8 (or 4) when that costs at most a quarter more lanes."""
    from pastml_b200.ml import padded_sets
    assert padded_sets(1, 1) == 1 and padded_sets(1, 8) == 1
    assert padded_sets(21, 8) == 21          # 168 lanes: already whole tiles
    assert padded_sets(21, 1) == 24
    assert padded_sets(2, 1) == 2            # padding 2 -> 4 would double the work
    assert padded_sets(21, 2) == 24
    assert padded_sets(11, 16) == 11
    assert padded_sets(5, 3) == 5            # 15 lanes -> 24 is too expensive, -> 16 is not a multiple of 3
    for n in range(1, 40):
        for c in range(1, 20):
            p = padded_sets(n, c)
            assert n <= p <= 1.25 * n + 1e-9


# ------------------------------------------------------------------ array-native newick ingest (csrc/pml_newick.cpp)

def _python_flat(text, name=True):
    """The Python path: read_tree per ';'-separated tree + negative lengths to zero + name_tree + flatten_forest."""
    from pastml_b200 import tree as T
    roots = [T.read_tree(nwk + ';') for nwk in text.replace('\n', '').split(';')[:-1] if nwk.strip()]
    for i, root in enumerate(roots):
        for node in root.traverse():
            if node.dist < 0:
                node.dist = 0
        if name:
            T.name_tree(root, suffix='_{}'.format(i) if len(roots) > 1 else '')
    flat = T.flatten_forest(roots)
    return flat, [n.name for n in flat.nodes]


def _random_newick(rng, n_tips, named_internal=0.5, dup=False):
    from pastml_b200 import tree as T
    nodes = [T.Tree(name='x{}'.format(i) if not dup else 'same') for i in range(n_tips)]
    k = 0
    while len(nodes) > 1:
        m = min(len(nodes), int(rng.integers(2, 5)))
        idx = sorted(rng.choice(len(nodes), size=m, replace=False))[::-1]
        parent = T.Tree(name='in{}'.format(k) if rng.random() < named_internal else '')
        k += 1
        for i in idx:
            ch = nodes.pop(i)
            ch.dist = float(np.round(rng.exponential(1.), 6)) if rng.random() < 0.9 else 0.
            ch.up = parent
            parent.children.append(ch)
        nodes.append(parent)
    return nodes[0].write(format=1)


@pytest.mark.parametrize('text', [
    "((A:1,B:2)ab:0.5,(C:-1,'D d':3)[&&NHX:x=1]:4,E)r;",
    "(A,B,(C,D));",
    "A;",
    "A:1.5;",
    "((a:1,b:1):1,(a:1,c:1):2);\n(x:1,(y:2,z:3)w:1)top:7;",
    "( A : 1 , B : 2e-3 ) ;",
    "((A:1,B:1)0.9:1,C:2);",
])
def test_newick_ingest_matches_python_reader(text):
    from pastml_b200 import tree as T
    for name in (False, True):
        want, want_names = _python_flat(text, name)
        got = T.read_forest_flat(text, name_nodes=name)
        assert np.array_equal(got.parent, want.parent)
        assert np.array_equal(got.dist, want.dist)
        assert list(got.names) == want_names


def test_newick_ingest_random_trees_and_naming_rules():
    from pastml_b200 import tree as T
    rng = np.random.default_rng(11)
    for trial in range(12):
        text = _random_newick(rng, int(rng.integers(2, 60)), named_internal=rng.random(), dup=trial % 4 == 3)
        if trial % 3 == 2:
            text += '\n' + _random_newick(rng, 7)
        want, want_names = _python_flat(text)
        got = T.read_forest_flat(text)
        assert np.array_equal(got.parent, want.parent)
        assert np.allclose(got.dist, want.dist, rtol=0, atol=0)
        assert list(got.names) == want_names
        assert len(set(got.names[got.tree_id == 0])) == int((got.tree_id == 0).sum())


def test_newick_ingest_golden_trees_round_trip():
    """Every fixture tree, written to newick from its arrays, comes back with the same level-ordered arrays."""
    from helpers import Golden, golden_names
    from pastml_b200 import tree as T
    for name in golden_names()[:8]:
        g = Golden(name)
        roots = T.forest_from_arrays(g['parent'], g['dist'], g['names'])
        text = '\n'.join(r.write(format=1) for r in roots)
        got = T.read_forest_flat(text, name_nodes=False)
        assert np.array_equal(got.parent, g['parent'])
        np.testing.assert_allclose(got.dist, g['dist'], rtol=1e-6)   # write() prints 6 significant digits
        keep = g['parent'] >= 0   # like ete3, write() leaves the root's name out
        assert list(got.names[keep]) == [str(x) for x in g['names'][keep]]


@pytest.mark.parametrize('text', ["((A,B);", "(A,B));", "(A,B)C)D;", "", "   ", "(A:x,B);", "(A,'B);"])
def test_newick_ingest_errors(text):
    from pastml_b200 import tree as T
    with pytest.raises(ValueError):
        T.read_forest_flat(text)


# ------------------------------------------------------------------ lock-step fit on a stand-in objective (no GPU)

class _FakeProblem(object):
    stride = 4 + 3


class _FakeBatch(object):
    """Stands in for ColumnBatch: the 'log-likelihood' of column c is a concave function of (sf, pi) with its maximum
    at a column-specific point, so that the columns need different numbers of L-BFGS-B iterations."""

    def __init__(self, models, lane_lists):
        self.models = models
        self.C = len(models)
        self.characters = ['c%d' % c for c in range(self.C)]
        self.problem = _FakeProblem()
        self.lane_lists = lane_lists
        self.batched_fd = True
        from pastml_b200.ml import padded_sets, lane_list_size
        self.max_sets = padded_sets(1 + max(m.get_num_params() for m in models), self.C)
        self.max_sets = max(self.max_sets, -(-lane_list_size(self.max_sets * self.C) // self.C))   # as ColumnBatch
        self.lanes_evaluated = 0
        self.calls = []

    @staticmethod
    def _value(c, block):
        sf, pi = block[0], block[4:7]
        target = np.array([.2 + .05 * c, .5 - .03 * c, .3 - .02 * c])
        return -100. * (c + 1) * (np.log(sf) - np.log(0.5 + c)) ** 2 - 50. * ((pi - target) ** 2).sum() - 10.

    def loglik_sets(self, params, n_sets):
        self.calls.append(('grid', n_sets * self.C))
        self.lanes_evaluated += n_sets * self.C
        return np.array([[self._value(c, params[b, c]) for c in range(self.C)] for b in range(n_sets)]), {}

    def loglik_lanes(self, params, lane_columns):
        assert len(lane_columns) % 4 == 0 and len(lane_columns) <= self.max_sets * self.C
        self.calls.append(('lanes', len(lane_columns)))
        self.lanes_evaluated += len(lane_columns)
        return np.array([self._value(c, params[i]) for i, c in enumerate(lane_columns)]), {}


def _fake_fit(lane_lists, C=8):
    from pastml_b200.annotation import ForestStats
    from pastml_b200.ml import fit_batch
    from pastml_b200.models.F81Model import F81Model
    from pastml_b200.tree import FlatForest
    flat = FlatForest(np.array([-1, 0, 0, 1, 1], dtype=np.int32), np.array([0., 1., 2., 1., 1.]))
    fs = ForestStats(flat=flat)
    models = [F81Model(states=np.array(['a', 'b', 'c']), forest_stats=fs, character='c%d' % c) for c in range(C)]
    batch = _FakeBatch(models, lane_lists)
    out, passes = fit_batch(batch, [np.full(3, 1. / 3)] * C)
    return out, passes, batch, models


def test_lockstep_fit_lane_lists_match_the_grid():
    """The lock-step fit with lane-list passes reaches the values of the (sets x columns) grid -- same requests, same
    answers -- while evaluating fewer lanes once columns start to finish; every pass is padded to whole lane tiles."""
    from pastml_b200.ml import lane_list_size
    assert [lane_list_size(n) for n in (1, 4, 5, 8, 9, 63, 64, 65, 147)] == [4, 4, 8, 8, 16, 64, 64, 96, 160]
    grid_out, grid_passes, grid_batch, grid_models = _fake_fit(False)
    lane_out, lane_passes, lane_batch, lane_models = _fake_fit(True)
    assert lane_passes == grid_passes
    np.testing.assert_allclose(lane_out, grid_out, rtol=0, atol=0)
    for a, b in zip(grid_models, lane_models):
        assert a.sf == b.sf and np.array_equal(a.frequencies, b.frequencies)
    assert all(kind == 'lanes' for kind, _ in lane_batch.calls) and all(kind == 'grid' for kind, _ in grid_batch.calls)
    assert lane_batch.lanes_evaluated < grid_batch.lanes_evaluated
    for c, m in enumerate(lane_models):   # the optimum of the stand-in objective
        assert m.sf == pytest.approx(0.5 + c, rel=1e-3)


# ------------------------------------------------------------------ polytomy resolution (tree surgery)

def _random_predicted_tree(rng, n_tips):
    """A tree with polytomies, names, dates and a predicted state set on every node (two columns)."""
    from pastml_b200 import tree as T
    text = _random_newick(rng, n_tips, named_internal=1.)
    root = T.read_tree(text)
    T.name_tree(root)
    for node in root.traverse('preorder'):
        node.add_feature('date', 0. if node.is_root() else node.up.date + node.dist)
        a = {str(rng.choice(['x', 'y', 'z']))}
        if rng.random() < 0.15:
            a.add(str(rng.choice(['x', 'y', 'z'])))
        node.add_feature('colA', a)
        node.add_feature('colB', {str(rng.choice(['p', 'q']))})
    return root


def _describe(root):
    return [(n.name, round(n.dist, 9), n.up.name if n.up else None, sorted(n.colA), sorted(n.colB),
             int(getattr(n, 'polytomy', 0))) for n in root.traverse()]


def test_polytomy_surgery_matches_reference():
    """resolve_trees / unresolve_trees against the reference's functions on the same predicted trees."""
    from oracle import ref_env
    if not ref_env.reference_available():
        pytest.skip('reference sources not available')
    ref_env.install()
    from pastml import tree as R
    from pastml_b200 import tree as T
    column2states = {'colA': np.array(['x', 'y', 'z']), 'colB': np.array(['p', 'q'])}
    rng = np.random.default_rng(5)
    created = removed = 0
    for trial in range(25):
        seed = int(rng.integers(1 << 30))
        mine = _random_predicted_tree(np.random.default_rng(seed), int(rng.integers(6, 60)))
        theirs = _random_predicted_tree(np.random.default_rng(seed), 0) if False else mine.copy()
        n_mine, n_theirs = T.resolve_trees(column2states, [mine]), R.resolve_trees(column2states, [theirs])
        assert n_mine == n_theirs
        assert _describe(mine) == _describe(theirs)
        created += n_mine
        # new predictions (as after a likelihood recalculation), the same on both copies
        flip = np.random.default_rng(seed + 1)
        for a, b in zip(mine.traverse(), theirs.traverse()):
            if flip.random() < 0.3:
                s = {str(flip.choice(['x', 'y', 'z']))}
                a.add_feature('colA', s)
                b.add_feature('colA', set(s))
        r_mine, r_theirs = T.unresolve_trees(column2states, [mine]), R.unresolve_trees(column2states, [theirs])
        assert r_mine == r_theirs
        assert _describe(mine) == _describe(theirs)
        removed += r_mine
    assert created > 10 and removed > 5


# ------------------------------------------------------------------ array-native tables (pastml_b200/arrays.py)

def test_table_masks_and_frequencies_match_the_node_path():
    from helpers import Golden, annotated_forest
    from pastml_b200.acr import calculate_observed_freqs
    from pastml_b200.arrays import table_masks, observed_frequencies
    for name in ('albania_f81', 'denv2_location_f81', 'synth_s2_n60_k7_f81'):
        g = Golden(name)
        column = g.meta['column']
        rows = [(str(g['names'][i]), g.states[k]) for i in range(len(g['names'])) if g['has_state'][i]
                for k in range(g.K) if (int(g['annot_mask'][i]) >> k) & 1]
        rows += [(str(g['names'][i]), None) for i in range(len(g['names'])) if g['has_state'][i] and not g['annot_mask'][i]]
        df = pd.DataFrame({column: [s for _, s in rows]}, index=[n for n, _ in rows])
        states, mask, has = table_masks(g.flat, df, column, np.array(g.states))
        assert list(states) == list(g.states)
        assert np.array_equal(mask, g['annot_mask']) and np.array_equal(has, g['has_state'])
        roots = annotated_forest(g)
        _, want, _ = calculate_observed_freqs(column, roots, list(g.states))
        np.testing.assert_allclose(observed_frequencies(g.flat, mask, g.K), want, rtol=1e-14)


def test_write_results_combined_states_match_the_node_writer(tmp_path):
    """combined_ancestral_states.tab from state masks = the file the node-object writer produces."""
    from pastml_b200.arrays import write_results
    from pastml_b200.ml import STATE_MASKS
    from pastml_b200.pipeline import serialize_predicted_states
    from pastml_b200.tree import read_tree, name_tree, flatten_forest
    rng = np.random.default_rng(9)
    root = read_tree(_random_newick(rng, 40))
    name_tree(root)
    flat = flatten_forest(root)
    states = {'colA': np.array(['b', 'a', 'c']), 'col B': np.array(['x', 'y'])}
    results = []
    for c, st in states.items():
        masks = rng.integers(1, 1 << len(st), size=flat.n_nodes).astype(np.uint64)
        for node, m in zip(flat.nodes, masks.tolist()):
            node.add_feature(c, {st[k] for k in range(len(st)) if (m >> k) & 1})
        results.append({'character': c, 'states': st, 'method': 'DOWNPASS', STATE_MASKS: masks})
    serialize_predicted_states(sorted(states), str(tmp_path / 'nodes.tab'), [root])
    write_results(results, flat, str(tmp_path / 'w'), out_data=str(tmp_path / 'arrays.tab'))
    assert open(tmp_path / 'arrays.tab').read() == open(tmp_path / 'nodes.tab').read()


def test_array_native_parsimony_plumbing(tmp_path, monkeypatch):
    """acr_flat + write_results for a parsimony method with the engine problem replaced by the oracle (host plumbing
    only: masks from the table, result dictionaries, result files) against the reference pipeline's files."""
    from oracle import cpu_oracle
    from pastml_b200 import arrays
    from pastml_b200.tree import read_forest_flat

    class OracleProblem(object):
        def __init__(self, flat, K, kind, masks, has, device=0):
            self.flat, self.K, self.masks = flat, K, masks

        def parsimony(self, flags, want_masks=True):
            out = {}
            per_column = [cpu_oracle.parsimony_fixed(self.flat, self.K, m) for m in self.masks]
            for name, flag in (('ACCTRAN', 1), ('DOWNPASS', 2), ('DELTRAN', 4)):
                if flags & flag:
                    out[name] = (np.stack([r[name + '_mask'] for r in per_column], axis=1),
                                 np.array([r[name + '_steps'] for r in per_column], dtype=np.int64))
            return out

        def close(self):
            pass

    monkeypatch.setattr(arrays, 'Problem', OracleProblem)
    src = os.path.join(REPO, 'tests', 'golden', 'pipeline', 'albania_downpass')
    flat = read_forest_flat(os.path.join(src, 'input.nwk'))
    df = pd.read_csv(os.path.join(src, 'input.tab'), sep='\t', index_col=0, header=0, dtype=str)
    results = arrays.acr_flat(flat, df, columns=['Country'], prediction_method='DOWNPASS')
    work = str(tmp_path / 'out')
    arrays.write_results(results, flat, work)
    want_dir = os.path.join(src, 'expected')
    for name in ('combined_ancestral_states.tab', 'params.character_Country.method_DOWNPASS.tab'):
        assert open(os.path.join(work, name)).read() == open(os.path.join(want_dir, name)).read(), name
    # the whole array-native pipeline: every file of the reference's work directory, with the same text
    work2 = str(tmp_path / 'out2')
    arrays.pastml_pipeline_flat(os.path.join(src, 'input.nwk'), os.path.join(src, 'input.tab'), columns=['Country'],
                                prediction_method='DOWNPASS', work_dir=work2)
    assert sorted(os.listdir(work2)) == sorted(os.listdir(want_dir))
    for name in os.listdir(want_dir):
        assert open(os.path.join(work2, name)).read() == open(os.path.join(want_dir, name)).read(), name
    with pytest.raises(ValueError, match='not found among the annotation columns'):
        arrays.pastml_pipeline_flat(os.path.join(src, 'input.nwk'), os.path.join(src, 'input.tab'), columns=['Nope'],
                                    work_dir=work2)


def test_write_named_tree_matches_the_node_writer(tmp_path):
    from pastml_b200.arrays import write_named_tree
    from pastml_b200.ml import STATE_MASKS
    from pastml_b200.pipeline import annotate_dist_to_root
    from pastml_b200.tree import read_tree, name_tree, flatten_forest
    rng = np.random.default_rng(21)
    roots = []
    for i in range(2):
        root = read_tree(_random_newick(rng, 30 + 10 * i))
        name_tree(root, suffix='_%d' % i)
        roots.append(root)
    annotate_dist_to_root(roots)
    flat = flatten_forest(roots)
    states = {'colA': np.array(['b', 'a', 'c']), 'colB': np.array(['x', 'y'])}
    results = []
    for c, st in states.items():
        masks = rng.integers(1, 1 << len(st), size=flat.n_nodes).astype(np.uint64)
        for node, m in zip(flat.nodes, masks.tolist()):
            node.add_feature(c, {st[k] for k in range(len(st)) if (m >> k) & 1})
        results.append({'character': c, 'states': st, 'method': 'MPPA', STATE_MASKS: masks})
    want = '\n'.join(r.write(format_root_node=True, format=3, features=['date', 'date_CI', 'colA', 'colB']) for r in roots)
    write_named_tree(results, flat, str(tmp_path / 'named.nwk'))
    assert open(tmp_path / 'named.nwk').read() == want


def test_lockstep_fit_reports_a_failing_lane_to_its_column():
    """A zero-likelihood lane of a lane-list pass surfaces as the error of the column that asked for it, the other
    columns finish their fit."""
    from pastml_b200.ml import fit_batch, PastMLLikelihoodError
    from pastml_b200.annotation import ForestStats
    from pastml_b200.models.F81Model import F81Model
    from pastml_b200.tree import FlatForest

    class Failing(_FakeBatch):
        def loglik_lanes(self, params, lane_columns):
            values, _ = _FakeBatch.loglik_lanes(self, params, lane_columns)
            errs = {j: PastMLLikelihoodError('column {} is broken'.format(c)) for j, c in enumerate(lane_columns) if c == 2}
            return values, errs

    flat = FlatForest(np.array([-1, 0, 0, 1, 1], dtype=np.int32), np.array([0., 1., 2., 1., 1.]))
    fs = ForestStats(flat=flat)
    models = [F81Model(states=np.array(['a', 'b', 'c']), forest_stats=fs, character='c%d' % c) for c in range(4)]
    batch = Failing(models, True)
    with pytest.raises(PastMLLikelihoodError, match='column 2 is broken'):
        fit_batch(batch, [np.full(3, 1. / 3)] * 4)
    assert any(kind == 'lanes' for kind, _ in batch.calls)


def test_lockstep_fit_propagates_a_failing_device_pass():
    """When the device pass itself raises (engine error, out of memory), every column thread gets the exception and
    fit_batch returns it to the caller instead of hanging on the rendezvous."""
    import threading
    from pastml_b200.ml import fit_batch
    from pastml_b200.annotation import ForestStats
    from pastml_b200.models.F81Model import F81Model
    from pastml_b200.tree import FlatForest

    class Broken(_FakeBatch):
        def loglik_lanes(self, params, lane_columns):
            if len(self.calls) >= 3:
                raise RuntimeError('device pass failed')
            return _FakeBatch.loglik_lanes(self, params, lane_columns)

    flat = FlatForest(np.array([-1, 0, 0, 1, 1], dtype=np.int32), np.array([0., 1., 2., 1., 1.]))
    fs = ForestStats(flat=flat)
    models = [F81Model(states=np.array(['a', 'b', 'c']), forest_stats=fs, character='c%d' % c) for c in range(4)]
    batch = Broken(models, True)
    outcome = []

    def run():
        try:
            fit_batch(batch, [np.full(3, 1. / 3)] * 4)
            outcome.append(None)
        except BaseException as e:
            outcome.append(e)

    t = threading.Thread(target=run, daemon=True)
    t.start()
    t.join(timeout=60)
    assert not t.is_alive(), 'fit_batch hangs when the device pass raises'
    assert isinstance(outcome[0], RuntimeError) and 'device pass failed' in str(outcome[0])


def test_newick_writers_sanitise_names_and_feature_values(tmp_path):
    """ete3 (the reference's writer) replaces the characters that would break a newick / NHX file by underscores, in
    names and in feature values: states such as 'Congo (DRC)' or 'Korea, South' must not corrupt named.tree_*.nwk."""
    from pastml_b200.arrays import write_named_tree
    from pastml_b200.ml import STATE_MASKS
    from pastml_b200.pipeline import annotate_dist_to_root
    from pastml_b200.tree import read_tree, name_tree, flatten_forest
    root = read_tree('((a:1,b:2):1,(c:1,d:1):2);')
    name_tree(root)
    root.children[0].name = 'odd;name'
    annotate_dist_to_root([root])
    flat = flatten_forest([root])
    states = np.array(['Congo (DRC)', 'Korea, South', 'x=y'])
    masks = np.array([3, 1, 2, 4, 1, 2, 7], dtype=np.uint64)
    for node, m in zip(flat.nodes, masks.tolist()):
        node.add_feature('Country', {states[k] for k in range(3) if (m >> k) & 1})
    text = root.write(format_root_node=True, format=3, features=['date', 'Country'])
    assert 'Country=Congo _DRC_|Korea_ South' in text and 'x_y' in text and 'odd_name' in text
    assert re.search(r'Country=[^:\]]*[(),;=]', text) is None
    write_named_tree([{'character': 'Country', 'states': states, 'method': 'MPPA', STATE_MASKS: masks}], flat,
                     str(tmp_path / 'named.nwk'))
    assert open(tmp_path / 'named.nwk').read() == text
    back = read_tree(text)   # and it parses again
    assert len(list(back.traverse())) == 7


def test_pipeline_input_validation_on_cpu(tmp_path):
    """validate_input (acr.py:677-822): column checks, id matching, states and naming -- no engine involved."""
    from pastml_b200.pipeline import validate_input, get_pastml_parameter_file, get_pastml_marginal_prob_file, \
        get_named_tree_file, get_pastml_work_dir
    src = os.path.join(REPO, 'tests', 'golden', 'pipeline', 'albania_mppa_f81')
    tree, data = os.path.join(src, 'input.nwk'), os.path.join(src, 'input.tab')
    roots, columns, column2states, parameters, rates = validate_input(tree, ['Country'], data)
    assert columns == ['Country'] and len(roots) == 1 and parameters == {} and rates == {}
    assert list(column2states['Country']) == sorted(column2states['Country'])
    annotated = sum(1 for n in roots[0].traverse() if getattr(n, 'Country', None))
    assert annotated > 100
    with pytest.raises(ValueError, match='not found among the annotation columns'):
        validate_input(tree, ['Nope'], data)
    with pytest.raises(ValueError, match='need to provide an annotated tree'):
        validate_input(tree, None, None)
    bad = tmp_path / 'bad.tab'
    bad.write_text('id\tCountry\nnobody\tX\n')
    with pytest.raises(ValueError, match='do not correspond to annotation id'):
        validate_input(tree, ['Country'], str(bad))
    sparse = tmp_path / 'sparse.tab'
    first_tip = next(iter(roots[0])).name
    sparse.write_text('id\tCountry\n{}\tX\n'.format(first_tip))
    with pytest.raises(ValueError, match='are unknown'):
        validate_input(tree, ['Country'], str(sparse))
    _, _, _, params, _ = validate_input(tree, ['Country'], data, parameters={'Country': {'scaling_factor': 1.}})
    assert params == {'Country': {'scaling_factor': 1.}}
    # file names of pastml/file.py
    assert get_pastml_parameter_file('MPPA', 'F81', 'Country') == 'params.character_Country.method_MPPA.model_F81.tab'
    assert get_pastml_parameter_file('DOWNPASS', None, 'Country') == 'params.character_Country.method_DOWNPASS.tab'
    assert get_pastml_marginal_prob_file('MAP', 'JC', 'Country') == 'marginal_probabilities.character_Country.model_JC.tab'
    assert get_pastml_marginal_prob_file('JOINT', 'JC', 'Country') is None
    assert get_named_tree_file('/x/y/my.tree.nwk') == 'named.tree_my.tree.nwk'
    assert get_pastml_work_dir('/x/y/my.tree.nwk') == '/x/y/my.tree_pastml'


# ------------------------------------------------------------------ array-native pipeline: input checks

def test_flat_pipeline_checks_and_popcount():
    """The array-native pipeline applies validate_input's sanity checks (>= 90% unknown tips, a state per tip), and
    the state-set statistics do not depend on numpy >= 2."""
    import pandas as pd
    from pastml_b200.arrays import _check_table
    from pastml_b200.parsimony import _popcount
    from pastml_b200.tree import read_forest_flat
    flat = read_forest_flat('((a:1,b:1):1,(c:1,d:1):1,(e:1,(f:1,(g:1,(h:1,(i:1,(j:1,k:1):1):1):1):1):1):1);')
    names = read_forest_flat('((a:1,b:1):1,(c:1,d:1):1,(e:1,(f:1,(g:1,(h:1,(i:1,(j:1,k:1):1):1):1):1):1):1);', name_nodes=False).names
    ok = pd.DataFrame({'x': ['s', 't', 's', 't', 's']}, index=list('abcde'))
    _check_table(flat, names, ok, ['x'])
    with pytest.raises(ValueError, match='are unknown'):
        _check_table(flat, names, pd.DataFrame({'x': ['s']}, index=['a']), ['x'])
    with pytest.raises(ValueError, match='unique states'):
        _check_table(flat, names, pd.DataFrame({'x': list('pqrstuvwxyz')}, index=list('abcdefghijk')), ['x'])
    masks = np.array([0, 1, 3, 0xffffffffffffffff, 1 << 61, 0x5555], dtype=np.uint64)
    want = np.array([0, 1, 2, 64, 1, 8])
    assert np.array_equal(_popcount(masks), want)
    import pastml_b200.parsimony as P
    saved = getattr(np, 'bitwise_count', None)
    if saved is not None:
        try:
            del np.bitwise_count
            assert np.array_equal(P._popcount(masks), want) and np.array_equal(P._popcount(masks.reshape(2, 3)), want.reshape(2, 3))
        finally:
            np.bitwise_count = saved


# ------------------------------------------------------------------ parameter blocks of a finite-difference batch

def test_engine_blocks_equal_the_point_by_point_route():
    """``Model.engine_blocks(points)`` (all the blocks of one gradient in a few array operations, F81 family) gives bit
    for bit what ``set_params_from_optimised(ps); engine_block()`` gives point by point, for both stages of the fit, and
    leaves the model at the last point; the other models and the smoothing options take the point-by-point route."""
    from pastml_b200.annotation import ForestStats
    from pastml_b200.models.F81Model import F81Model
    from pastml_b200.models.JCModel import JCModel
    from pastml_b200.models.EFTModel import EFTModel
    from pastml_b200.models.HKYModel import HKYModel
    from pastml_b200.tree import FlatForest
    flat = FlatForest(np.array([-1, 0, 0, 1, 1], dtype=np.int32), np.array([0., 1., 2., 1., 1.]))
    fs = ForestStats(flat=flat)
    rng = np.random.default_rng(11)

    def make(kind, K, **kw):
        states = np.array(['s%02d' % i for i in range(K)])
        if kind == 'F81':
            return F81Model(states=states, forest_stats=fs, **kw)
        if kind == 'JC':
            return JCModel(states=states, forest_stats=fs, **kw)
        if kind == 'EFT':
            return EFTModel(states=states, forest_stats=fs, observed_frequencies=rng.dirichlet(np.ones(K)), **kw)
        return HKYModel(states=np.array(list('ACGT')), forest_stats=fs, **kw)

    cases = [('F81', K, {}) for K in (2, 3, 5, 7, 8, 9, 10, 16, 17, 20, 33, 62)] + \
            [('JC', 5, {}), ('EFT', 6, {}), ('HKY', 4, {}), ('F81', 6, dict(frequency_smoothing=True)),
             ('F81', 6, dict(tau=0., optimise_tau=True))]
    for kind, K, kw in cases:
        for fix_extra in (True, False):
            a, b = make(kind, K, **kw), make(kind, K, **kw)
            b._frequencies = a._frequencies if hasattr(a, '_frequencies') else None
            if fix_extra:
                a.fix_extra_params()
                b.fix_extra_params()
            n_par = a.get_num_params() if not fix_extra else len(a.get_optimised_parameters())
            if n_par == 0:
                continue
            lo, hi = a.get_bounds()[:, 0], a.get_bounds()[:, 1]
            points = [np.clip(rng.random(n_par) * 3. + 0.05, lo, hi) for _ in range(n_par + 3)]
            want = []
            for ps in points:
                a.set_params_from_optimised(ps)
                want.append(a.engine_block())
            got = b.engine_blocks(points)
            assert len(got) == len(want)
            for g, w in zip(got, want):
                assert np.array_equal(np.asarray(g), np.asarray(w)), (kind, K, fix_extra)
            assert np.array_equal(a.engine_block(), b.engine_block())   # same state afterwards
            assert np.array_equal(a.get_optimised_parameters(), b.get_optimised_parameters())


# ------------------------------------------------------------------ the coroutine L-BFGS-B driver

def _drive(gen, fun):
    """Runs a minimize_lbfgsb generator against a plain function."""
    try:
        request = next(gen)
        while True:
            request = gen.send([fun(p) for p in request])
    except StopIteration as stop:
        return stop.value


def test_coroutine_lbfgsb_is_scipy_minimize_bit_for_bit():
    """``lbfgsb.minimize_lbfgsb`` (requests batched as one point + its forward differences) walks through the same
    iterates as ``scipy.optimize.minimize(method='L-BFGS-B')`` with bounds: same x, same value, same iteration count and
    status, on objectives shaped like the fit's (log-scale parameter, ratios bounded by [1e-6, 1e7], optimum on a
    bound, a start on a bound, a flat direction, inf on part of the domain)."""
    from scipy.optimize import minimize
    from pastml_b200.lbfgsb import minimize_lbfgsb
    rng = np.random.default_rng(5)

    def rosen(x):
        return float(sum(100. * (x[1:] - x[:-1] ** 2) ** 2 + (1 - x[:-1]) ** 2))

    def likelihood_like(x):
        sf, ratios = x[0], x[1:]
        pi = np.hstack((ratios, [1.]))
        pi = pi / pi.sum()
        target = np.linspace(1, 2, len(pi))
        target = target / target.sum()
        return float(40. * (np.log(sf) - np.log(3.)) ** 2 - 300. * (target * np.log(pi)).sum())

    def with_inf(x):
        return np.inf if x[0] > 7.5 else float((x[0] - 2.) ** 2 + (x[1] + 1.) ** 4 + 0. * x[2])

    cases = [
        (rosen, np.array([-1.2, 1., 0.5]), np.array([[-2., 2.]] * 3)),
        (rosen, np.array([0.5, 0.5, 0.5, 0.5]), np.array([[0., 0.8]] * 4)),                      # optimum on a bound
        (likelihood_like, np.r_[1., np.ones(4)], np.array([[0.001, 10.]] + [[1e-6, 10e6]] * 4)),
        (likelihood_like, np.r_[10., rng.random(9) + .5], np.array([[0.001, 10.]] + [[1e-6, 10e6]] * 9)),   # start on a bound
        (with_inf, np.array([5., 3., 1.]), np.array([[0., 10.], [-5., 5.], [0., 2.]])),
    ]
    for fun, x0, bounds in cases:
        want = minimize(fun, x0=x0, method='L-BFGS-B', bounds=bounds)
        got = _drive(minimize_lbfgsb(x0, bounds), fun)
        assert np.array_equal(got.x, want.x), (fun.__name__, got.x, want.x)
        assert got.fun == want.fun and got.nit == want.nit and got.success == want.success and got.status == want.status


def test_coroutine_fit_equals_the_threaded_fit(monkeypatch):
    """The two drivers of the lock-step fit -- one scipy ``minimize`` per column in its own thread, and all the columns
    as coroutines in one thread (``lbfgsb.minimize_lbfgsb``) -- return the same log-likelihoods and leave the same
    parameters on the models, bit for bit; the coroutine driver needs about half the device passes (a point travels with
    its forward differences)."""
    monkeypatch.setenv('PASTML_B200_FIT_DRIVER', 'threads')
    t_out, t_passes, _, t_models = _fake_fit(True)
    monkeypatch.setenv('PASTML_B200_FIT_DRIVER', 'coroutines')
    c_out, c_passes, c_batch, c_models = _fake_fit(True)
    assert t_out == c_out
    for a, b in zip(t_models, c_models):
        assert a.sf == b.sf and np.array_equal(a.frequencies, b.frequencies)
    assert c_passes < 0.7 * t_passes
    assert all(kind == 'lanes' for kind, _ in c_batch.calls)


# ------------------------------------------------------------------ allowed-state bookkeeping helpers

def test_allowed_state_helpers_match_the_reference():
    """initialize / alter / unalter of the ``_ALLOWED_STATES`` features and their conversions: same features on the
    same nodes as the reference's functions, on random trees with zero-length branches and annotated internal nodes."""
    from oracle import ref_env
    if not ref_env.reference_available():
        pytest.skip('the reference sources are not present')
    ref_env.install()
    import pastml.ml as R
    import pastml_b200.ml as O
    from pastml_b200 import tree as T
    rng = np.random.default_rng(12)
    states = np.array(['a', 'b', 'c', 'd'])
    for trial in range(12):
        text = _random_newick(rng, 25, named_internal=1.)
        trees = [T.read_tree(text), T.read_tree(text)]
        for tree in trees:
            r2 = np.random.default_rng(100 + trial)
            for node in tree.traverse('preorder'):
                if not node.is_root() and r2.random() < .35:
                    node.dist = 0.
                u = r2.random()
                if node.is_leaf() or u < .4:
                    node.add_feature('x', set(r2.choice(states, size=r2.integers(0, 3), replace=False).tolist()))
        ref_tree, our_tree = trees
        R.initialize_allowed_states(ref_tree, 'x', states)
        O.initialize_allowed_states(our_tree, 'x', states)
        ref_alt = R.alter_zero_node_allowed_states(ref_tree, 'x')
        our_alt = O.alter_zero_node_allowed_states(our_tree, 'x')
        assert [n.name for n in ref_alt] == [n.name for n in our_alt]

        def same(feature):
            for a, b in zip(ref_tree.traverse(), our_tree.traverse()):
                va, vb = getattr(a, feature, None), getattr(b, feature, None)
                assert (va is None) == (vb is None), feature
                if va is not None:
                    assert np.array_equal(va, vb) if isinstance(va, np.ndarray) else va == vb, feature
        same('x_ALLOWED_STATES')
        same('x_ALLOWED_STATES.initial')
        for tree, mod, alt in ((ref_tree, R, ref_alt), (our_tree, O, our_alt)):
            for node in tree.traverse():
                node.add_feature('x_BOTTOM_UP_LIKELIHOOD_JOINT_STATES', (np.arange(4) + len(node.name)) % 4)
            mod.unalter_zero_node_joint_states(alt, 'x')
            mod.unalter_zero_node_allowed_states(alt, 'x')
            mod.convert_allowed_states2feature(tree, 'x', states, out_feature='y')
            mod._parsimonious_states2allowed_states(tree, 'y', 'z', states)
        same('x_ALLOWED_STATES')
        same('x_BOTTOM_UP_LIKELIHOOD_JOINT_STATES')
        same('y')
        same('z_ALLOWED_STATES')
    ones, table = O.get_state2allowed_states(states)
    r_ones, r_table = R.get_state2allowed_states(states)
    assert np.array_equal(ones, r_ones) and table.keys() == r_table.keys()
    assert all(np.array_equal(table[k], r_table[k]) for k in table)


def test_table_masks_with_unnamed_and_repeated_node_names():
    """Joining the annotation table on the names as written in the tree file: unnamed internal nodes (all '') match
    nothing, and every node of a repeated name takes the row (preannotate_forest runs before the nodes are renamed)."""
    import pandas as pd
    from pastml_b200.arrays import table_masks
    from pastml_b200.tree import read_forest_flat
    text = '((a:1,b:1):1,(c:1,a:1):1,d:2);'
    flat = read_forest_flat(text)
    raw = read_forest_flat(text, name_nodes=False).names
    assert list(raw).count('') == 3 and list(raw).count('a') == 2
    df = pd.DataFrame({'x': ['s', 't', 's', None]}, index=['a', 'b', 'd', 'zzz'])
    states, mask, has = table_masks(flat, df, 'x', names=raw)
    assert list(states) == ['s', 't']
    raw = np.asarray(raw, dtype=object)
    want_mask = np.zeros(flat.n_nodes, dtype=np.uint64)
    want_mask[raw == 'a'] = 1
    want_mask[raw == 'b'] = 2
    want_mask[raw == 'd'] = 1
    assert np.array_equal(mask, want_mask)
    assert has.sum() == 4 and np.all(has[raw == 'a'] == 1) and not has[raw == ''].any()


def test_probability_table_writer_equals_pandas(tmp_path):
    """The fast writer of the marginal-probability tables produces pandas' ``to_csv(sep='\t', index_label='node')``
    text byte for byte (shortest round-trip floats, exponents, zeros and ones, NaN as an empty cell), and defers to
    pandas for anything that is not a plain float table."""
    import pandas as pd
    from pastml_b200.pipeline import write_probability_table
    rng = np.random.default_rng(2)
    values = rng.random((500, 6))
    values /= values.sum(axis=1)[:, None]
    values[0] = [1., 0., 0., 0., 0., 0.]
    values[1] = [1e-300, 1e-5, 0.1, 1 / 3, 2 / 3, 123456789.125]
    values[2, 3] = np.nan
    values[3] = [5e-324, 1e22, 1e16, 1e15, 0.5, 1e-4]
    values[4] = [-0.0, np.inf, -np.inf, -2.5e-7, 9999999999999998., 1.7976931348623157e308]
    values[100:400] = 10. ** rng.uniform(-320, 308, (300, 6)) * rng.choice([-1., 1.], (300, 6))   # every magnitude
    values[400:450] = rng.integers(0, 10 ** 17, (50, 6)).astype(float)
    df = pd.DataFrame(values, index=['n%d' % i for i in range(500)], columns=list('ABCDEF'))
    fast, slow = str(tmp_path / 'fast.tab'), str(tmp_path / 'slow.tab')
    write_probability_table(df, fast)
    df.to_csv(slow, sep='\t', index_label='node')
    assert open(fast, 'rb').read() == open(slow, 'rb').read()
    odd = pd.DataFrame({'A': [1, 2]}, index=['x\ty', 'z'])   # integers, a tab in a label: pandas' quoting rules
    write_probability_table(odd, fast)
    odd.to_csv(slow, sep='\t', index_label='node')
    assert open(fast, 'rb').read() == open(slow, 'rb').read()


def test_probability_table_writer_edge_cases(tmp_path):
    """Empty tables, one column, non-ASCII labels: still pandas' text."""
    import pandas as pd
    from pastml_b200.pipeline import write_probability_table
    cases = [pd.DataFrame(np.zeros((0, 3)), index=[], columns=['a', 'b', 'c']),
             pd.DataFrame([[0.25], [0.75]], index=['nœud_é', 'узел'], columns=['état']),
             pd.DataFrame(np.full((2, 2), np.nan), index=['x', 'y'], columns=['p', 'q'])]
    for i, df in enumerate(cases):
        fast, slow = str(tmp_path / 'fast{}.tab'.format(i)), str(tmp_path / 'slow{}.tab'.format(i))
        write_probability_table(df, fast)
        df.to_csv(slow, sep='\t', index_label='node')
        assert open(fast, 'rb').read() == open(slow, 'rb').read(), i

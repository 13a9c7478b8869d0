"""Array-native ACR for large trees: newick file + annotation table in, arrays and result files out, without a Python
object per node.

``read_forest_flat`` (C++ newick parser) gives the level-ordered forest, :func:`table_masks` turns the columns of the
annotation table into the engine's bit masks, :func:`acr_flat` runs the same fit and reconstruction as
:func:`pastml_b200.acr.acr` and returns the usual result dictionaries with the predicted states as
``state_masks`` arrays, :func:`write_results` writes the reference's result files from them
(``combined_ancestral_states.tab``, ``params.*.tab``, ``marginal_probabilities.*.tab``; ``acr.py:45-73, 831-860``) and
:func:`write_named_tree` the annotated newick.
"""
import logging
import os

import numpy as np
import pandas as pd

from pastml_b200 import value2list, col_name2cat, CHARACTER, STATES
from pastml_b200.annotation import ForestStats
from pastml_b200.ml import (ColumnBatch, fit_batch, reconstruct_batch, is_ml, MPPA, ALL, STATE_MASKS, ML_METHODS)
from pastml_b200.engine import Problem
from pastml_b200.models import model2class, F81, HKY, JTT, HKY_STATES, JTT_STATES
from pastml_b200.parsimony import is_parsimonious, parsimony_on_problem, MP_METHODS
from pastml_b200.tree import newick_safe


def table_masks(flat, df, column, states=None, names=None):
    """(states, annot_mask [N] uint64, has_state [N] uint8) for one column of an annotation table indexed by node
    names (a node listed several times gets the union of its states, as ``annotation.df2gdf`` does;
    ``annotation.py:104-123``; ``has_state`` marks every node the table lists).  ``states`` defaults to the sorted values of the column (``acr.py:211-213``)."""
    values = df[column]
    text = values.astype(object).where(values.notna(), None)
    filled = np.array([v is not None and str(v) != '' for v in text], dtype=bool)
    if states is None:
        states = np.array(sorted({str(v) for v in text[filled]}))
    codes = np.full(len(values), -1, dtype=np.int64)
    codes[filled] = pd.Index(np.asarray(states, dtype=object)).get_indexer(pd.Index([str(v) for v in text[filled]]))
    # (table row, node) pairs by name: one vectorised hash join.  ``names``: join on these instead of flat.names -- the
    # names as written in the tree file, before name_tree; there a name may repeat, and like preannotate_forest
    # (annotation.py:113-123, which runs before the nodes are renamed) EVERY node of that name takes the row; unnamed
    # nodes ('') match nothing
    node_names = np.asarray(flat.names if names is None else names, dtype=object)
    index = pd.Index(node_names)
    ids = pd.Index(values.index.astype(str))
    if index.is_unique:
        node = index.get_indexer(ids)
        row = np.arange(len(ids))
        keep = node >= 0
        row, node = row[keep], node[keep]
    else:
        named = np.flatnonzero(node_names != '')
        pairs = pd.DataFrame({'name': ids, 'row': np.arange(len(ids))}).merge(
            pd.DataFrame({'name': node_names[named], 'node': named}), on='name', how='inner')
        row, node = pairs['row'].to_numpy(), pairs['node'].to_numpy()
    mask = np.zeros(flat.n_nodes, dtype=np.uint64)
    ok = codes[row] >= 0
    np.bitwise_or.at(mask, node[ok], np.uint64(1) << codes[row][ok].astype(np.uint64))
    # "has a state" in the reference's zero-branch clusters = the node carries the feature at all, and
    # preannotate_forest gives it (as an empty set) to every node the table lists, missing values included (ml.py:329-331)
    has = np.zeros(flat.n_nodes, dtype=np.uint8)
    has[node] = 1
    return np.asarray(states), mask, has


def observed_frequencies(flat, mask, K):
    """Tip-state frequencies, a tip with n states counting 1/n for each (``acr.py:282-299``)."""
    tips = mask[flat.is_tip]
    bits = ((tips[:, None] >> np.arange(K, dtype=np.uint64)) & np.uint64(1)).astype(np.float64)
    n = bits.sum(axis=1)
    freqs = (bits[n > 0] / n[n > 0, None]).sum(axis=0)
    return freqs / freqs.sum()


def acr_flat(flat, df, columns=None, prediction_method=MPPA, model=F81, column2parameters=None, column2rates=None,
             force_joint=True, reoptimise=False, tau=0, frequency_smoothing=False, device=0, join_names=None):
    """:func:`pastml_b200.acr.acr` for an array-native forest (``tree.read_forest_flat``) and an annotation table
    indexed by node names, ML and parsimony methods.  Returns the list of result dictionaries; the predicted states of a
    character are ``result['state_masks']`` (uint64 per node, bit k = ``result['states'][k]``) and the marginal
    probabilities keep their DataFrame (rows in the forest's level order).  ``join_names``: the node names the table
    is joined on when they differ from ``flat.names`` (the pipeline joins on the names as written in the file and names
    the nodes afterwards, like the reference: ``acr.py:487-531``).  Per-column parameters / rate matrices may be keyed by
    the column name as given or normalised (``col_name2cat``)."""
    columns = list(df.columns if columns is None else ([columns] if isinstance(columns, str) else columns))
    methods = value2list(len(columns), prediction_method, MPPA)
    models = value2list(len(columns), model, F81)
    column2parameters = column2parameters or {}
    column2rates = column2rates or {}

    def option(table, column):
        return table.get(column, table.get(col_name2cat(column)))

    forest_stats = ForestStats(flat=flat)
    settings = []
    for column, method, model_name in zip(columns, methods, models):
        character = col_name2cat(column)
        if is_parsimonious(method):
            states, mask, has = table_masks(flat, df, column, names=join_names)
            settings.append((character, method, states, mask, has))
            continue
        if not is_ml(method):
            raise ValueError('Method {} is unknown, should be one of ML ({}) or one of MP ({})'
                             .format(method, ', '.join(ML_METHODS), ', '.join(MP_METHODS)))
        fixed_states = HKY_STATES if model_name == HKY else JTT_STATES if model_name == JTT else None
        states, mask, has = table_masks(flat, df, column, fixed_states, names=join_names)
        if fixed_states is not None and not mask.any():
            raise ValueError('The {} model is only available for characters with states in {}, but none of the values '
                             'of "{}" is one of them.'.format(model_name, ', '.join(fixed_states), column))
        # acr.py:185-187: tau is reassigned inside the loop over the characters, so only the first ML character has
        # its smoothing factor optimised when smoothing is on (and every character when reoptimise is)
        optimise_tau = tau is None or reoptimise
        if tau is None:
            tau = 0
        instance = model2class[model_name](parameter_file=option(column2parameters, column),
                                           rate_matrix_file=option(column2rates, column), reoptimise=reoptimise,
                                           frequency_smoothing=frequency_smoothing, tau=tau,
                                           optimise_tau=optimise_tau, states=states, forest_stats=forest_stats,
                                           observed_frequencies=observed_frequencies(flat, mask, len(states)),
                                           character=character)
        settings.append((character, method, instance, mask, has))
    groups = {}
    for i, (_, method, model_or_states, _, _) in enumerate(settings):
        key = ('ML', method, model_or_states.engine_kind, len(model_or_states.states)) if is_ml(method) \
            else ('MP', method, len(model_or_states))
        groups.setdefault(key, []).append(i)
    per_character = {}
    for key, members in groups.items():
        masks = np.stack([settings[i][3] for i in members])
        has = np.stack([settings[i][4] for i in members])
        characters = [settings[i][0] for i in members]
        if key[0] == 'MP':
            problem = Problem(flat, key[2], 0, masks, has, device=device)
            try:
                res = parsimony_on_problem(problem, flat, characters, [settings[i][2] for i in members], key[1],
                                           forest_stats.num_nodes, forest_stats.num_tips)[0]
            finally:
                problem.close()
        else:
            batch = ColumnBatch(flat, characters, [settings[i][2] for i in members], device=device, annot_mask=masks,
                                has_state=has)
            try:
                obs = [observed_frequencies(flat, settings[i][3], len(settings[i][2].states)) for i in members]
                likelihoods, _ = fit_batch(batch, obs)
                hook = None
                if key[1] == ALL:   # the parsimony part of the ALL meta-method, as in ml_acr_batch
                    from pastml_b200.parsimony import all_methods_hook as hook
                res = reconstruct_batch(batch, key[1], likelihoods, force_joint, write_likelihood_features=False,
                                        parsimony_hook=hook)
            finally:
                batch.close()
        per_character.update(zip(members, res))
    return [r for i in range(len(settings)) for r in per_character[i]]


def write_results(results, flat, work_dir, out_data=None):
    """The reference's result files from array-native results: ``combined_ancestral_states.tab`` (one row per node and
    per rank of its predicted states, ``acr.py:831-860``) and, per character, the parameter and
    marginal-probability files (``acr.py:45-73``)."""
    from pastml_b200.pipeline import serialize_acr, get_combined_ancestral_state_file
    os.makedirs(work_dir, exist_ok=True)
    by_character = {r[CHARACTER]: r for r in results}
    columns = sorted(by_character)
    names = np.asarray(flat.names, dtype=object)
    N = flat.n_nodes
    # per column: the states of every node in increasing order of their names, row r of a node = its r-th state
    ranked, depth = {}, np.zeros(N, dtype=np.int64)
    for c in columns:
        states = np.asarray(by_character[c][STATES], dtype=object)
        order = np.argsort(states.astype(str), kind='stable')
        mask = by_character[c][STATE_MASKS]
        bits = ((mask[:, None] >> order.astype(np.uint64)[None, :]) & np.uint64(1)).astype(bool)   # [N][K] in name order
        rank = np.cumsum(bits, axis=1) - 1
        ranked[c] = (bits, rank, states[order])
        depth = np.maximum(depth, bits.sum(axis=1))
    with open(out_data or os.path.join(work_dir, get_combined_ancestral_state_file()), 'w+') as f:
        f.write('node\t{}\n'.format('\t'.join(columns)))
        rows_of = np.repeat(np.arange(N), depth)
        row_rank = np.arange(len(rows_of)) - np.repeat(np.cumsum(depth) - depth, depth)
        cells = []
        for c in columns:
            bits, rank, state_names = ranked[c]
            col = np.full(len(rows_of), '', dtype=object)
            node_idx, state_idx = np.nonzero(bits)
            # row of (node, its r-th state) = first row of the node + r
            first = np.cumsum(depth) - depth
            col[first[node_idx] + rank[node_idx, state_idx]] = state_names[state_idx]
            cells.append(col)
        table = np.column_stack([names[rows_of]] + cells)
        f.write('\n'.join('\t'.join(map(str, row)) for row in table))
        if len(table):
            f.write('\n')
    for r in results:
        serialize_acr(r, work_dir)
    logging.getLogger('pastml').debug('Serialized the results to {}.'.format(work_dir))


def write_named_tree(results, flat, path):
    """The named, annotated tree file of ``pastml_pipeline`` (``named.tree_<tree>.nwk``: extended newick with the
    distance to the root as ``date`` and the predicted states of every character, ``acr.py:503-507``) written from the
    arrays: same text as ``Tree.write(format=3, format_root_node=True, features=[date, ...])`` on node objects."""
    N = flat.n_nodes
    parent, dist = flat.parent, flat.dist
    dates = [0] * N   # a root's date is the integer 0, a child's its parent's plus the branch (tree.py:51-63)
    for v in range(N):
        if parent[v] >= 0:
            dates[v] = dates[parent[v]] + float(dist[v])
    labels = ['{}:{:g}[&&NHX:date={}'.format(newick_safe(flat.names[v]), float(dist[v]), dates[v]) for v in range(N)]
    for r in results:
        states = np.asarray(r[STATES], dtype=object)
        order = np.argsort(states.astype(str), kind='stable')
        names_sorted = [str(s) for s in states[order]]
        # one piece of text per distinct state set (few of them, whatever the number of nodes)
        uniq, inverse = np.unique(r[STATE_MASKS], return_inverse=True)
        pieces = []
        for m in uniq.tolist():
            chosen = [n for n, k in zip(names_sorted, order.tolist()) if (m >> k) & 1]
            pieces.append(':{}={}'.format(r[CHARACTER], newick_safe('|'.join(chosen))) if chosen else '')
        labels = [label + pieces[i] for label, i in zip(labels, inverse.tolist())]
    # children are contiguous in level order: assemble bottom-up, one string per subtree
    text = [None] * N
    first, count = flat.first_child, flat.n_children
    for v in range(N - 1, -1, -1):
        label = labels[v] + ']'
        if count[v]:
            kids = range(first[v], first[v] + count[v])
            text[v] = '(' + ','.join(text[k] for k in kids) + ')' + label
            for k in kids:
                text[k] = None
        else:
            text[v] = label
    with open(path, 'w+') as f:
        f.write('\n'.join(text[r] + ';' for r in flat.roots))


def _check_table(flat, names, df, columns):
    """validate_input's sanity checks (``acr.py:487-531``) on the table: enough annotated tips, a discrete character."""
    num_tips = int(flat.is_tip.sum())
    listed = df[df.index.isin(set(names))]
    annotated, seen = {}, {}
    for c in columns:
        values = listed[c].dropna()
        values = values[values != '']
        annotated[c] = values.index.nunique()
        seen[c] = set(values)
    sparsest = min(columns, key=lambda c: annotated[c])
    unknown = (num_tips - annotated[sparsest]) / num_tips
    if unknown >= .9:
        raise ValueError('{:.1f}% of tip annotations for character "{}" are unknown, not enough data to infer ancestral '
                         'states. Check your annotation file and if its ids correspond to the tree tip/node names.'
                         .format(unknown * 100, sparsest))
    poorest = min(columns, key=lambda c: len(seen[c]))
    if len(seen[poorest]) > num_tips * .75:
        raise ValueError('Character "{}" has {} unique states annotated in this tree: {}, which is too much to infer on a '
                         '{} with only {} tips. Make sure the character you are analysing is discrete, and if yes use a '
                         'larger tree.'.format(poorest, len(seen[poorest]), seen[poorest],
                                               'tree' if len(flat.roots) == 1 else 'forest', num_tips))


def pastml_pipeline_flat(tree, data, data_sep='\t', id_index=0, columns=None, prediction_method=MPPA, model=F81,
                         parameters=None, rate_matrix=None, out_data=None, work_dir=None, forced_joint=False,
                         reoptimise=False, smoothing=False, frequency_smoothing=False, device=0):
    """``pastml_pipeline`` for trees too large for a Python object per node: C++ newick ingest, annotation table to bit
    masks, the same fit and reconstruction, and the same result files (combined states, named tree, parameters,
    marginal probabilities) written from arrays.  Same argument meaning as :func:`pastml_b200.pipeline.pastml_pipeline`
    (``acr.py:316``); returns ``(flat, results)``."""
    from pastml_b200.pipeline import get_pastml_work_dir, get_named_tree_file, _keyed
    from pastml_b200.tree import read_forest_flat
    flat = read_forest_flat(tree)
    original_names = read_forest_flat(tree, name_nodes=False).names   # what the table ids are matched against
    df = pd.read_csv(data, sep=data_sep, index_col=id_index, header=0, dtype=str)
    df.index = df.index.map(str)
    if columns:
        columns = [columns] if isinstance(columns, str) else list(columns)
        unknown = set(columns) - set(df.columns)
        if unknown:
            raise ValueError('{} of the specified columns ({}) {} not found among the annotation columns: {}.'
                             .format('One' if len(unknown) == 1 else 'Some', ', '.join('"{}"'.format(c) for c in unknown),
                                     'is' if len(unknown) == 1 else 'are', ', '.join('"{}"'.format(c) for c in df.columns)))
        df = df[columns]
    columns = list(df.columns)
    if not len(set(df.index) & set(original_names)):
        raise ValueError('Your tree tip names (e.g. {}) do not correspond to annotation id column values (e.g. {}). '
                         'Check your annotation file.'.format(', '.join(map(str, flat.names[flat.is_tip][:3])),
                                                            ', '.join(map(str, df.index[:3]))))
    _check_table(flat, original_names, df, columns)
    results = acr_flat(flat, df, columns=columns, prediction_method=prediction_method, model=model, join_names=original_names,
                       column2parameters=_keyed(parameters, columns, 'Parameters'),
                       column2rates=_keyed(rate_matrix, columns, 'Rate matrices'), force_joint=forced_joint,
                       reoptimise=reoptimise, tau=None if smoothing else 0, frequency_smoothing=frequency_smoothing,
                       device=device)
    if not work_dir:
        work_dir = get_pastml_work_dir(tree)
    write_results(results, flat, work_dir, out_data=out_data)
    write_named_tree(results, flat, os.path.join(work_dir, get_named_tree_file(tree)))
    return flat, results

"""``pastml_pipeline()``: tree file + annotation table in, reconstruction files out.

Drop-in for the non-visual part of ``pastml.acr.pastml_pipeline`` (``/root/reference/pastml/acr.py:316-560``): input
validation (``acr.py:677-822``), :func:`pastml_b200.acr.acr`, and the output files a pastml user finds in the work
directory -- ``combined_ancestral_states.tab`` (``acr.py:831-860``), ``named.tree_<tree>.nwk``,
``params.character_<c>.method_<m>[.model_<model>].tab`` and ``marginal_probabilities.character_<c>.model_<model>.tab``
(``acr.py:45-73``, file names of ``pastml/file.py``).  Visualisation (html maps, iTOL, pajek), dates and nexus export are
outside the accelerated path: asking for them raises ``NotImplementedError``.
"""
import logging
import os

import numpy as np
import pandas as pd

from pastml_b200 import col_name2cat, METHOD, CHARACTER, STATES
from pastml_b200.acr import acr, COPY
from pastml_b200.annotation import preannotate_forest
from pastml_b200.ml import (MPPA, is_ml, is_marginal, MODEL,
                            MARGINAL_PROBABILITIES, STATE_MASKS)
from pastml_b200.models import F81
from pastml_b200.tree import read_forest, name_tree

from pastml_b200 import PASTML_VERSION   # the reference version whose file formats are reproduced
from pastml_b200.file import (get_pastml_parameter_file, get_pastml_marginal_prob_file, get_combined_ancestral_state_file,  # noqa: F401
                              get_named_tree_file, get_pastml_work_dir)


# ---------------------------------------------------------------------------------------- input checks
#
# What the reference checks before a reconstruction (``acr.py:677-822``, dates left out): the messages are part of its
# observable behaviour and are kept; the checks themselves are independent steps over the annotation table and the
# per-column tallies of the annotated forest.

def _quote(names):
    return ', '.join('"{}"'.format(n) for n in names)


def _keyed(value, columns, what):
    """A per-column option given as one value, a list parallel to the columns, or a dict."""
    if not value:
        return {}
    if isinstance(value, str):
        value = [value]
    if isinstance(value, list):
        return dict(zip(columns, value))
    if isinstance(value, dict):
        return {col_name2cat(col): v for col, v in value.items()}
    raise ValueError('{} should be either a list or a dict, got {}.'.format(what, type(value)))


def _read_table(data, data_sep, id_index, columns):
    """The annotation table restricted to the requested columns, column names normalised."""
    df = pd.read_csv(data, sep=data_sep, index_col=id_index, header=0, dtype=str)
    df.index = df.index.map(str)
    if columns:
        unknown = [c for c in dict.fromkeys(columns) if c not in df.columns]
        if unknown:
            one = len(unknown) == 1
            raise ValueError('{} of the specified columns ({}) {} not found among the annotation columns: {}.'
                             .format('One' if one else 'Some', _quote(unknown), 'is' if one else 'are', _quote(df.columns)))
        df = df[columns]
    df.columns = [col_name2cat(c) for c in df.columns]
    return df


def _match_ids(roots, ids):
    """Node names and table ids must overlap; names that only match without their quotes are unquoted in the tree."""
    nodes = [n for root in roots for n in root.traverse()]
    names = {n.name for n in nodes if n.name}
    if names & ids:
        return
    strip = lambda name: name.strip("'").strip('"')   # noqa: E731
    if {strip(name) for name in names} & ids:
        for n in nodes:
            n.name = strip(n.name)
        return
    some_tips = [tip.name for tip, _ in zip(roots[0].iter_leaves(), range(3))]
    raise ValueError('Your tree tip names (e.g. {}) do not correspond to annotation id column values (e.g. {}). '
                     'Check your annotation file.'.format(', '.join(some_tips), ', '.join(list(ids)[:3])))


def _tally(roots, columns):
    """Per column: the states seen on the nodes and the number of annotated nodes; plus the number of tips."""
    seen = {c: set() for c in columns}
    annotated = dict.fromkeys(columns, 0)
    num_tips = 0
    for root in roots:
        for n in root.traverse():
            num_tips += n.is_leaf()
            for c in columns:
                values = getattr(n, c, None)
                if values:
                    seen[c] |= values
                    annotated[c] += 1
    return seen, annotated, num_tips


def validate_input(tree_nwk, columns=None, data=None, data_sep='\t', id_index=0, copy_only=False, parameters=None,
                   rates=None):
    """Reads the forest and the annotation table, annotates the nodes, collects the states of every character and
    applies the reference's sanity checks.  Returns ``(roots, columns, column2states, parameters, rates)``."""
    if not columns and data is None:
        raise ValueError("If you don't provide the metadata file, "
                         "you need to provide an annotated tree and specify the columns argument, "
                         "which will be used to look for character annotations in your input tree.")
    if isinstance(columns, str):
        columns = [columns]
    roots = read_forest(tree_nwk, columns=columns if data is None else None)
    table_states = {}
    if data:
        df = _read_table(data, data_sep, id_index, columns)
        columns = list(df.columns)
        _match_ids(roots, set(df.index))
        preannotate_forest(roots, df=df)
        table_states = {c: {v for v in df[c].unique() if pd.notnull(v) and v != ''} for c in columns}
    else:
        columns = [col_name2cat(c) for c in columns]

    seen, annotated, num_tips = _tally(roots, columns)
    # the sparsest column decides whether there is enough data at all
    sparsest = min(columns, key=lambda c: annotated[c])
    unknown = (num_tips - annotated[sparsest]) / num_tips
    if unknown >= (1 if copy_only else .9):
        hint = 'Check your annotation file and if its ids correspond to the tree tip/node names.' if data else \
            'You tree file should contain character state annotations, otherwise consider specifying a metadata file.'
        raise ValueError('{:.1f}% of tip annotations for character "{}" are unknown, '
                         'not enough data to infer ancestral states. {}'.format(unknown * 100, sparsest, hint))
    # a character with (almost) as many states as tips is not a discrete character
    poorest = min(columns, key=lambda c: len(seen[c]))
    if len(seen[poorest]) > num_tips * .75 and not copy_only:
        raise ValueError('Character "{}" has {} unique states annotated in this tree: {}, '
                         'which is too much to infer on a {} with only {} tips. '
                         'Make sure the character you are analysing is discrete, and if yes use a larger tree.'
                         .format(poorest, len(seen[poorest]), seen[poorest], 'tree' if len(roots) == 1 else 'forest', num_tips))
    logging.getLogger('pastml').debug('Finished input validation.')
    column2states = {c: np.array(sorted(seen[c] | table_states.get(c, set()))) for c in columns}
    for i, tree in enumerate(roots):
        name_tree(tree, suffix='' if len(roots) == 1 else '_{}'.format(i))
    return roots, columns, column2states, _keyed(parameters, columns, 'Parameters'), _keyed(rates, columns, 'Rate matrices')


# ---------------------------------------------------------------------------------------- output files

def serialize_predicted_states(columns, out_data, roots):
    """``combined_ancestral_states.tab``: one row per node and per combination rank of its predicted states -- a node
    with several states in a column takes several rows, its states in increasing order (``acr.py:831-860``)."""
    with open(out_data, 'w+') as f:
        f.write('node\t{}\n'.format('\t'.join(columns)))
        for root in roots:
            for node in root.traverse():
                pending = {}
                for column in columns:
                    value = getattr(node, column, set())
                    if value:
                        pending[column] = sorted(value, reverse=True)
                while pending:
                    cells = []
                    for column in columns:
                        values = pending.get(column)
                        if values is None:
                            cells.append('')
                            continue
                        cells.append(str(values.pop()))
                        if not values:
                            del pending[column]
                    f.write('{}\t{}\n'.format(node.name, '\t'.join(cells)))
    logging.getLogger('pastml').debug('Serialized reconstructed states to {}.'.format(out_data))


def write_probability_table(df, path):
    """``df.to_csv(path, sep='\\t', index_label='node')`` for a table of floats, the same text (shortest round-trip
    representation of every number, empty cell for NaN; checked against pandas in tests/test_host_cpu.py) without
    pandas' per-chunk float -> str conversion, which took 6 of the 8.5 s of the array-native pipeline on a 100k-tip
    tree.  Anything that is not a plain float table goes through pandas."""
    values = df.to_numpy()
    plain = values.dtype == np.float64 and values.ndim == 2 and all(isinstance(c, str) for c in df.columns) and \
        not any(ch in str(c) for c in df.columns for ch in '\t\n\r"') and df.index.nlevels == 1
    if plain:
        labels = df.index.tolist()
        try:   # labels that would need pandas' quoting rules (or are not text) go through pandas
            text = '\x00'.join(labels)
            plain = not any(ch in text for ch in '\t\n\r"')
        except TypeError:
            plain = False
    if not plain:
        df.to_csv(path, sep='\t', index_label='node')
        return
    # the engine library's host-side writer (csrc/pml_newick.cpp: shortest round-trip digits in Python's repr() layout)
    import ctypes as C
    from pastml_b200 import _lib
    lib = _lib.load()
    encoded = [v.encode() for v in labels]
    offsets = np.zeros(len(encoded) + 1, dtype=np.int64)
    offsets[1:] = np.cumsum([len(v) for v in encoded])
    table = np.ascontiguousarray(values, dtype=np.float64)
    header = ('node\t' + '\t'.join(df.columns)).encode()
    rc = lib.pml_write_float_table(str(path).encode(), header, b''.join(encoded), offsets.ctypes.data_as(C.c_void_p),
                                   table.ctypes.data_as(C.c_void_p), len(encoded), table.shape[1])
    if rc != 0:
        raise IOError(lib.pml_newick_last_error().decode())


def serialize_acr(acr_result, work_dir):
    """The parameter / statistics file of one reconstruction and, for marginal methods, the table of marginal
    probabilities (``acr.py:45-73``)."""
    model_name = acr_result[MODEL].name if MODEL in acr_result else None
    out_param_file = os.path.join(work_dir, get_pastml_parameter_file(method=acr_result[METHOD], model=model_name,
                                                                      column=acr_result[CHARACTER]))
    with open(out_param_file, 'w+') as f:
        f.write('parameter\tvalue\n')
        f.write('pastml_version\t{}\n'.format(PASTML_VERSION))
        for name in sorted(acr_result.keys()):
            if name not in [STATES, MARGINAL_PROBABILITIES, METHOD, MODEL, STATE_MASKS]:
                f.write('{}\t{}\n'.format(name, acr_result[name]))
        f.write('{}\t{}\n'.format(METHOD, acr_result[METHOD]))
        if is_ml(acr_result[METHOD]):
            acr_result[MODEL].save_parameters(f)
    if is_marginal(acr_result[METHOD]):
        out_mp_file = os.path.join(work_dir, get_pastml_marginal_prob_file(method=acr_result[METHOD], model=model_name,
                                                                           column=acr_result[CHARACTER]))
        write_probability_table(acr_result[MARGINAL_PROBABILITIES], out_mp_file)


DATE, DATE_CI = 'date', 'date_CI'


def annotate_dist_to_root(forest):
    """The ``date`` feature of the named tree when no root date is given: the node's distance to its root, unless the
    input tree already carries dates (``tree.py:51-73`` with ``root_dates=None``)."""
    for tree in forest:
        for node in tree.traverse('preorder'):
            if getattr(node, DATE, None) is None:
                if node.is_root():
                    node.add_feature(DATE, 0)
                else:
                    parent_date = getattr(node.up, DATE, None)
                    if parent_date is not None:
                        node.add_feature(DATE, parent_date + node.dist)
            else:
                node.add_feature(DATE, float(getattr(node, DATE)))
            ci = getattr(node, DATE_CI, None)
            if ci and not isinstance(ci, (list, tuple)):
                node.del_feature(DATE_CI)
                if isinstance(ci, str) and '|' in ci:
                    try:
                        node.add_feature(DATE_CI, [float(_) for _ in ci.split('|')])
                    except ValueError:
                        pass


def clear_extra_features(forest, features):
    features = set(features) | {'name', 'dist', 'support'}
    for tree in forest:
        for n in tree.traverse():
            for f in set(n.features) - features:
                if f not in features:
                    n.del_feature(f)


# ---------------------------------------------------------------------------------------- the pipeline

def pastml_pipeline(tree, data=None, data_sep='\t', id_index=0, columns=None, prediction_method=MPPA, model=F81,
                    parameters=None, rate_matrix=None, name_column=None, root_date=None, timeline_type=None,
                    tip_size_threshold=None, colours=None, out_data=None, html_compressed=None, html=None,
                    html_mixed=None, work_dir=None, verbose=False, forced_joint=False, upload_to_itol=False,
                    itol_id=None, itol_project=None, itol_tree_name=None, offline=False, threads=0, reoptimise=False,
                    focus=None, resolve_polytomies=False, smoothing=False, frequency_smoothing=False, pajek=None,
                    pajek_timing=None, recursion_limit=0, device=0):
    """Reconstructs the ancestral states of the characters in ``columns`` for the tree(s) in ``tree`` (newick) annotated
    by the table ``data`` and writes the result files to ``work_dir`` (default ``<tree>_pastml``).  Same arguments as
    the reference; the visualisation options must stay unset.  Returns the list of ACR result dictionaries (the
    reference returns nothing)."""
    for option, value in (('html', html), ('html_compressed', html_compressed), ('html_mixed', html_mixed),
                          ('upload_to_itol', upload_to_itol), ('root_date', root_date), ('pajek', pajek)):
        if value:
            raise NotImplementedError('{}: visualisation and dating are outside the accelerated ACR path '
                                      '(use the reference pastml on the files written to work_dir)'.format(option))
    logging.getLogger('pastml').setLevel(logging.DEBUG if verbose else logging.ERROR)
    copy_only = COPY == prediction_method or (isinstance(prediction_method, list)
                                              and all(COPY == _ for _ in prediction_method))
    roots, columns, column2states, parameters, rates = \
        validate_input(tree, columns, data, data_sep, id_index, copy_only=copy_only, parameters=parameters,
                       rates=rate_matrix)
    annotate_dist_to_root(roots)
    if not work_dir:
        work_dir = get_pastml_work_dir(tree)
    os.makedirs(work_dir, exist_ok=True)

    acr_results = acr(forest=roots, columns=columns, column2states=column2states, prediction_method=prediction_method,
                      model=model, column2parameters=parameters, column2rates=rates, force_joint=forced_joint,
                      threads=threads, reoptimise=reoptimise, tau=None if smoothing else 0,
                      resolve_polytomies=resolve_polytomies, frequency_smoothing=frequency_smoothing, device=device)
    column2states = {res[CHARACTER]: res[STATES] for res in acr_results}

    if not out_data:
        out_data = os.path.join(work_dir, get_combined_ancestral_state_file())
    serialize_predicted_states(sorted(column2states.keys()), out_data, roots)

    features = [DATE, DATE_CI] + list(column2states.keys())
    clear_extra_features(roots, features)
    with open(os.path.join(work_dir, get_named_tree_file(tree)), 'w+') as f:
        f.write('\n'.join(root.write(format_root_node=True, format=3, features=features) for root in roots))
    for res in acr_results:
        serialize_acr(res, work_dir)
    return acr_results

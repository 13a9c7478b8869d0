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
from collections import Counter, defaultdict

import numpy as np
import pandas as pd

from pastml_b200 import col_name2cat, get_personalized_feature_name, METHOD, CHARACTER, STATES
from pastml_b200.acr import acr, COPY
from pastml_b200.annotation import preannotate_forest
from pastml_b200.ml import (MPPA, is_ml, is_marginal, is_meta_ml, get_default_ml_method, MODEL,
                            MARGINAL_PROBABILITIES, STATE_MASKS)
from pastml_b200.models import F81
from pastml_b200.parsimony import is_meta_mp, get_default_mp_method
from pastml_b200.tree import read_forest, name_tree

PASTML_VERSION = '1.9.50'   # the reference version whose file formats are reproduced

PASTML_WORK_DIR = '{tree}_pastml'
COMBINED_ANCESTRAL_STATE_TAB = 'combined_ancestral_states.tab'
NAMED_TREE_NWK = 'named.tree_{tree}.nwk'
PASTML_ML_PARAMS_TAB = 'params.character_{state}.method_{method}.model_{model}.tab'
PASTML_MP_PARAMS_TAB = 'params.character_{state}.method_{method}.tab'
PASTML_MARGINAL_PROBS_TAB = 'marginal_probabilities.character_{state}.model_{model}.tab'


# ---------------------------------------------------------------------------------------- file names (pastml/file.py)

def get_column_method(column, method):
    column = col_name2cat(column)
    if is_meta_ml(method):
        method = get_default_ml_method()
    elif is_meta_mp(method):
        method = get_default_mp_method()
    else:
        return column, method
    return get_personalized_feature_name(column, method), method


def get_pastml_parameter_file(method, model, column):
    template = PASTML_ML_PARAMS_TAB if is_ml(method) else PASTML_MP_PARAMS_TAB
    column, method = get_column_method(column, method)
    return template.format(state=column, method=method, model=model)


def get_pastml_marginal_prob_file(method, model, column):
    if not is_marginal(method):
        return None
    column, method = get_column_method(column, method)
    return PASTML_MARGINAL_PROBS_TAB.format(state=column, model=model)


def get_combined_ancestral_state_file():
    return COMBINED_ANCESTRAL_STATE_TAB


def get_named_tree_file(tree):
    tree_name = os.path.splitext(os.path.basename(tree))[0]
    return NAMED_TREE_NWK.format(tree=tree_name)


def get_pastml_work_dir(tree):
    return PASTML_WORK_DIR.format(tree=os.path.splitext(tree)[0])


# ---------------------------------------------------------------------------------------- input validation

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


def validate_input(tree_nwk, columns=None, data=None, data_sep='\t', id_index=0, copy_only=False, parameters=None,
                   rates=None):
    """Reads the forest and the annotation table, annotates the nodes, collects the states of every character and
    applies the reference's sanity checks (``acr.py:677-822`` without the date handling).  Returns
    ``(roots, columns, column2states, parameters, rates)``."""
    logger = logging.getLogger('pastml')
    if not columns and data is None:
        raise ValueError("If you don't provide the metadata file, "
                         "you need to provide an annotated tree and specify the columns argument, "
                         "which will be used to look for character annotations in your input tree.")
    if columns and isinstance(columns, str):
        columns = [columns]
    roots = read_forest(tree_nwk, columns=columns if data is None else None)
    column2states = defaultdict(set)

    if data:
        df = pd.read_csv(data, sep=data_sep, index_col=id_index, header=0, dtype=str)
        df.index = df.index.map(str)
        if columns:
            unknown = set(columns) - set(df.columns)
            if unknown:
                raise ValueError('{} of the specified columns ({}) {} not found among the annotation columns: {}.'
                                 .format('One' if len(unknown) == 1 else 'Some', _quote(unknown),
                                         'is' if len(unknown) == 1 else 'are', _quote(df.columns)))
            df = df[columns]
        df.columns = [col_name2cat(column) for column in df.columns]
        columns = df.columns
        node_names = set.union(*[{n.name for n in root.traverse() if n.name} for root in roots])
        ids = set(df.index)
        common = node_names & ids
        if not common:   # names quoted in the tree only
            node_names = {_.strip("'").strip('"') for _ in node_names}
            common = node_names & ids
            if common:
                for root in roots:
                    for n in root.traverse():
                        n.name = n.name.strip("'").strip('"')
        if not common:
            tips = [tip.name for tip, _ in zip(roots[0].iter_leaves(), range(3))]
            raise ValueError('Your tree tip names (e.g. {}) do not correspond to annotation id column values (e.g. {}). '
                             'Check your annotation file.'.format(', '.join(tips), ', '.join(list(ids)[:3])))
        preannotate_forest(roots, df=df)
        for c in df.columns:
            column2states[c] |= {_ for _ in df[c].unique() if pd.notnull(_) and _ != ''}
    else:
        columns = [col_name2cat(c) for c in columns]

    num_tips = 0
    column2annotated = Counter()
    column2annotated_states = defaultdict(set)
    for root in roots:
        for n in root.traverse():
            for c in columns:
                vs = getattr(n, c, set())
                column2states[c] |= vs
                column2annotated_states[c] |= vs
                if vs:
                    column2annotated[c] += 1
            if n.is_leaf():
                num_tips += 1
    if column2annotated:
        c, num_annotated = min(column2annotated.items(), key=lambda _: _[1])
    else:
        c, num_annotated = columns[0], 0
    percentage_unknown = (num_tips - num_annotated) / num_tips
    if percentage_unknown >= (.9 if not copy_only else 1):
        raise ValueError('{:.1f}% of tip annotations for character "{}" are unknown, '
                         'not enough data to infer ancestral states. {}'
                         .format(percentage_unknown * 100, c,
                                 'Check your annotation file and if its ids correspond to the tree tip/node names.'
                                 if data else 'You tree file should contain character state annotations, '
                                              'otherwise consider specifying a metadata file.'))
    c, states = min(column2annotated_states.items(), key=lambda _: len(_[1]))
    if len(states) > num_tips * .75 and not copy_only:
        raise ValueError('Character "{}" has {} unique states annotated in this tree: {}, '
                         'which is too much to infer on a {} with only {} tips. '
                         'Make sure the character you are analysing is discrete, and if yes use a larger tree.'
                         .format(c, len(states), states, 'tree' if len(roots) == 1 else 'forest', num_tips))
    logger.debug('Finished input validation.')
    column2states = {c: np.array(sorted(states)) for c, states in column2states.items()}
    parameters = _keyed(parameters, columns, 'Parameters')
    rates = _keyed(rates, columns, 'Rate matrices')
    for i, tree in enumerate(roots):
        name_tree(tree, suffix='' if len(roots) == 1 else '_{}'.format(i))
    return roots, list(columns), column2states, parameters, rates


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
        acr_result[MARGINAL_PROBABILITIES].to_csv(out_mp_file, sep='\t', index_label='node')


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

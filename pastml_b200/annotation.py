"""Forest statistics and table -> node annotation (``/root/reference/pastml/annotation.py``)."""
import logging

import pandas as pd


class ForestStats(object):
    """avg_nonzero_brlen, num_nodes, num_tips, forest_length, num_trees -- the numbers the models start from
    (initial scaling factor 1/avg, optimisation bounds, tau factor).  Every node counts, roots included: a root
    whose branch length is 0 counts as a zero-length node exactly like in the reference
    (``annotation.py:36-101``)."""

    def __init__(self, forest=None, flat=None):
        if flat is None:
            from pastml_b200.tree import flatten_forest
            flat = flatten_forest(forest)
        dist = flat.dist
        nonzero = dist != 0
        n_nonzero = int(nonzero.sum())
        # the reference accumulates tip and internal lengths separately and in traversal order; python floats
        # summed in that order are reproduced with math.fsum-free sequential sums per tree
        total_ext, total_int = 0., 0.
        order = flat.ref_order
        is_tip = flat.is_tip
        for i in order:
            d = dist[i]
            if d:
                if is_tip[i]:
                    total_ext += d
                else:
                    total_int += d
        self.forest_length = total_ext + total_int
        self.num_nodes = int(flat.n_nodes)
        self.num_tips = int(flat.n_tips)
        self.avg_nonzero_brlen = self.forest_length / n_nonzero if n_nonzero else 0
        self.num_trees = int(flat.n_trees)
        logging.getLogger('pastml').debug(
            '\n=============TREE STATISTICS===================\n\tnumber of tips:\t{}\n\tnumber of internal nodes:\t{}\n'
            '\tavg non-zero branch length:\t{:.5f}.'.format(self.num_tips, self.num_nodes - self.num_tips,
                                                           self.avg_nonzero_brlen))


def df2gdf(df):
    """Groups a (node id x columns) table by node id into sets of non-empty values (``annotation.py:104-110``: ids in
    sorted order, one set per id and column, a repeated id collecting the values of all its rows).  One pass over each
    column instead of a pandas ``groupby().apply`` per group -- 300k groups x 82 columns took 6 of the 8.6 s of an
    82-column ``acr()`` call."""
    keys = df.index.unique()
    try:
        keys = keys.sort_values()
    except TypeError:   # ids of mixed types: keep the order of appearance
        pass
    codes = keys.get_indexer(df.index).tolist()
    data = {}
    for c in df.columns:
        column = df[c]
        missing = column.isna().to_numpy()
        values = column.astype(object).to_numpy().tolist()
        sets = [set() for _ in range(len(keys))]
        for k, v, gone in zip(codes, values, missing.tolist()):
            if not gone and v != '':
                sets[k].add(v)
        data[c] = pd.Series(sets, index=keys, dtype=object)
    return pd.DataFrame(data, index=keys, columns=df.columns, dtype=object)


def preannotate_forest(forest, df=None, gdf=None):
    """Sets, on every node named in the table, one feature per column holding the set of its states (an empty set
    for a missing value); the features are removed from the other nodes (``annotation.py:113-123``)."""
    if gdf is None:
        gdf = df2gdf(df)
    rows = {idx: row for idx, row in zip(gdf.index, gdf.to_dict('records'))}
    for tree in forest:
        for node in tree.traverse('postorder'):
            row = rows.get(node.name)
            if row is not None:
                node.add_features(**{c: set(v) for c, v in row.items()})
            else:
                for c in gdf.columns:
                    node.del_feature(c)
    return gdf.columns, gdf

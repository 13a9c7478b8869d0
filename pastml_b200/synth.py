"""Synthetic workloads generated straight into flat level-ordered arrays (no Python tree objects).

Recipes follow SURVEY.md section 8(d): Yule(n) -- split a uniformly chosen leaf until there are n tips, branch
lengths i.i.d. Exp(1) (the recipe of the reference's ``examples/simulations``); Coalescent(n) -- Kingman,
ultrametric, merge a uniform pair after an Exp(k(k-1)/2) waiting time; tip states by forward simulation down the
tree like ``pastml/utilities/state_simulator.py`` (root ~ pi, child ~ row of P(t)) under an F81 process.
"""
import numpy as np

from pastml_b200.tree import FlatForest


def _level_order(left, right, dist, root):
    """Relabels a binary tree given as child arrays (-1 for tips) into level order; returns (parent, dist)."""
    n = len(left)
    order = np.empty(n, dtype=np.int64)
    parent_new = np.empty(n, dtype=np.int32)
    cur = np.array([root], dtype=np.int64)
    cur_parent_new = np.array([-1], dtype=np.int32)
    pos = 0
    while len(cur):
        m = len(cur)
        order[pos:pos + m] = cur
        parent_new[pos:pos + m] = cur_parent_new
        new_ids = np.arange(pos, pos + m, dtype=np.int32)
        internal = left[cur] >= 0
        ci = cur[internal]
        nxt = np.empty(2 * len(ci), dtype=np.int64)
        nxt[0::2] = left[ci]
        nxt[1::2] = right[ci]
        cur_parent_new = np.repeat(new_ids[internal], 2)
        cur = nxt
        pos += m
    assert pos == n
    return parent_new, np.ascontiguousarray(dist[order], dtype=np.float64)


def yule_tree(n_tips, rng, mean_brlen=1.0):
    """Yule tree with n_tips tips as a :class:`FlatForest` (2 n_tips - 1 nodes, root branch length 0)."""
    n = 2 * n_tips - 1
    left = np.full(n, -1, dtype=np.int64)
    right = np.full(n, -1, dtype=np.int64)
    leaves = np.empty(n_tips, dtype=np.int64)
    leaves[0] = 0
    n_leaves, n_nodes = 1, 1
    picks = rng.random(n_tips - 1)
    for i in range(n_tips - 1):
        j = int(picks[i] * n_leaves)
        v = leaves[j]
        left[v], right[v] = n_nodes, n_nodes + 1
        leaves[j] = n_nodes
        leaves[n_leaves] = n_nodes + 1
        n_leaves += 1
        n_nodes += 2
    dist = rng.exponential(mean_brlen, size=n)
    dist[0] = 0.
    parent, dist = _level_order(left, right, dist, 0)
    return FlatForest(parent, dist)


def coalescent_tree(n_tips, rng, scale=1.0):
    """Kingman coalescent with n_tips tips (ultrametric), as a :class:`FlatForest`."""
    n = 2 * n_tips - 1
    left = np.full(n, -1, dtype=np.int64)
    right = np.full(n, -1, dtype=np.int64)
    height = np.zeros(n, dtype=np.float64)
    active = np.arange(n_tips, dtype=np.int64)
    k = n_tips
    ks = np.arange(n_tips, 1, -1, dtype=np.float64)
    waits = rng.exponential(1., size=n_tips - 1) / (ks * (ks - 1.) / 2.) * scale * n_tips
    u1 = rng.random(n_tips - 1)
    u2 = rng.random(n_tips - 1)
    t = 0.
    nxt = n_tips
    parent_of = np.full(n, -1, dtype=np.int64)
    for i in range(n_tips - 1):
        t += waits[i]
        a = int(u1[i] * k)
        b = int(u2[i] * (k - 1))
        if b >= a:
            b += 1
        va, vb = active[a], active[b]
        left[nxt], right[nxt] = va, vb
        parent_of[va] = parent_of[vb] = nxt
        height[nxt] = t
        lo, hi = (a, b) if a < b else (b, a)
        active[lo] = nxt
        active[hi] = active[k - 1]
        k -= 1
        nxt += 1
    root = n - 1
    dist = np.zeros(n, dtype=np.float64)
    has_parent = parent_of >= 0
    dist[has_parent] = height[parent_of[has_parent]] - height[has_parent]
    parent, dist = _level_order(left, right, dist, root)
    return FlatForest(parent, dist)


def simulate_tip_masks(flat, K, n_columns, rng, pi=None, changes_per_avg_branch=0.2, missing=0.0):
    """Forward-simulates n_columns characters with K states under F81 and returns
    (annot_mask uint64 [C][N] with tips only, has_state uint8 [C][N], pi [C][K], sf [C])."""
    N = flat.n_nodes
    nz = flat.dist[flat.dist > 0]
    avg = float(nz.mean()) if len(nz) else 1.
    masks = np.zeros((n_columns, N), dtype=np.uint64)
    has = np.zeros((n_columns, N), dtype=np.uint8)
    pis = np.empty((n_columns, K), dtype=np.float64)
    sfs = np.empty(n_columns, dtype=np.float64)
    tips = flat.is_tip
    for c in range(n_columns):
        p = rng.dirichlet(np.ones(K)) if pi is None else np.asarray(pi, dtype=np.float64)
        p = np.maximum(p, 1e-3)
        p /= p.sum()
        sf = changes_per_avg_branch / avg
        mu = 1. / (1. - p.dot(p))
        state = np.empty(N, dtype=np.int64)
        cdf = np.cumsum(p)
        cdf[-1] = 1.
        roots = flat.roots
        state[roots] = np.searchsorted(cdf, rng.random(len(roots)))
        for lvl in range(1, flat.n_levels):
            lo, hi = flat.level_ptr[lvl], flat.level_ptr[lvl + 1]
            par = flat.parent[lo:hi]
            stay = rng.random(hi - lo) < np.exp(-mu * flat.dist[lo:hi] * sf)
            fresh = np.searchsorted(cdf, rng.random(hi - lo))
            state[lo:hi] = np.where(stay, state[par], fresh)
        m = np.zeros(N, dtype=np.uint64)
        m[tips] = np.uint64(1) << state[tips].astype(np.uint64)
        h = tips.astype(np.uint8)
        if missing > 0:
            drop = tips & (rng.random(N) < missing)
            m[drop] = 0
            h[drop] = 0
        masks[c], has[c], pis[c], sfs[c] = m, h, p, sf
    return masks, has, pis, sfs


def forest(flats):
    """Merges forests (usually single trees) into ONE level-ordered :class:`FlatForest`: depth level d of the result
    is the concatenation of the level-d nodes of every input, input by input, so siblings stay contiguous and parents
    stay sorted.  The stand-in for a set of bootstrap trees analysed together (``Sum_trees``, ml.py:206-207)."""
    n_levels = max(f.n_levels for f in flats)
    total = sum(f.n_nodes for f in flats)
    parent = np.empty(total, dtype=np.int32)
    dist = np.empty(total, dtype=np.float64)
    new_id = [np.empty(f.n_nodes, dtype=np.int64) for f in flats]
    pos = 0
    for lvl in range(n_levels):
        for f, ids in zip(flats, new_id):
            if lvl >= f.n_levels:
                continue
            lo, hi = int(f.level_ptr[lvl]), int(f.level_ptr[lvl + 1])
            ids[lo:hi] = np.arange(pos, pos + hi - lo)
            par = f.parent[lo:hi]
            parent[pos:pos + hi - lo] = np.where(par >= 0, ids[np.maximum(par, 0)], -1)
            dist[pos:pos + hi - lo] = f.dist[lo:hi]
            pos += hi - lo
    return FlatForest(parent, dist)

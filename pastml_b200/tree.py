"""Rooted tree objects and their level-ordered flattening.

pastml keeps its trees as ``ete3.Tree`` objects (``/root/reference/pastml/tree.py:7-9``) and every
numerical routine walks them node by node.  ete3 is not a dependency of this package: :class:`Tree`
below provides the slice of the ete3 node interface that the ACR hot path of the reference touches
(``children/up/dist/name/features``, ``traverse`` in ete3's three orders,
``add_feature(s)/del_feature``, ``is_root/is_leaf``, leaf iteration, ``copy``, ``write``), so that
callers written against pastml keep working, and :func:`flatten_forest` turns such trees *once* into
the level-ordered CSR arrays the CUDA engine consumes.

The newick reader is a single-pass scanner (names, branch lengths, quoted labels, ``[&&NHX:..]``
and ``[&key=value]`` comments).  It is deliberately permissive: pastml's ``read_tree``
(``/root/reference/pastml/tree.py:202-222``) tries ete3 formats 3,2,5,0,1,... until one parses, the net
effect being "labels on internal nodes are names, missing things are allowed".
"""
from __future__ import annotations

import os
from collections import deque

import logging

import re

import numpy as np

DEFAULT_DIST = 1.0
DEFAULT_SUPPORT = 1.0

POSTORDER = 'postorder'
PREORDER = 'preorder'
LEVELORDER = 'levelorder'


class NewickError(ValueError):
    """Raised when a string cannot be parsed as newick (mirrors ete3.parser.newick.NewickError)."""


class Tree(object):
    """A node of a rooted tree (and, through its descendants, the subtree below it)."""

    __slots__ = ('children', 'up', 'dist', 'support', 'name', 'features', '__dict__')

    def __init__(self, newick=None, format=0, dist=None, support=None, name=None, quoted_node_names=False):
        self.children = []
        self.up = None
        self.dist = DEFAULT_DIST if dist is None else float(dist)
        self.support = DEFAULT_SUPPORT if support is None else float(support)
        self.name = '' if name is None else name
        self.features = {'dist', 'support', 'name'}
        if newick is not None:
            if os.path.exists(newick) and not newick.lstrip().startswith('('):
                with open(newick, 'r') as f:
                    newick = f.read()
            _parse_newick_into(self, newick, format)

    # ------------------------------------------------------------------ features
    def add_feature(self, pr_name, pr_value):
        setattr(self, pr_name, pr_value)
        self.features.add(pr_name)

    def add_features(self, **features):
        for key, value in features.items():
            setattr(self, key, value)
            self.features.add(key)

    def del_feature(self, pr_name):
        if pr_name in self.__dict__:
            del self.__dict__[pr_name]
        self.features.discard(pr_name)

    # ------------------------------------------------------------------ topology
    def is_root(self):
        return self.up is None

    def is_leaf(self):
        return not self.children

    def get_children(self):
        return list(self.children)

    def get_tree_root(self):
        node = self
        while node.up is not None:
            node = node.up
        return node

    def add_child(self, child=None, name=None, dist=None, support=None):
        if child is None:
            child = self.__class__()
        if name is not None:
            child.name = name
        if dist is not None:
            child.dist = float(dist)
        if support is not None:
            child.support = float(support)
        # like ete3: the child is NOT taken off its previous parent's list (callers detach it themselves; the
        # reference's polytomy surgery iterates over that list while re-hanging the children)
        child.up = self
        self.children.append(child)
        return child

    def remove_child(self, child):
        try:
            self.children.remove(child)
        except ValueError:
            raise ValueError('child not found')
        child.up = None
        return child

    def detach(self):
        if self.up is not None:
            self.up.children.remove(self)
            self.up = None
        return self

    def delete(self, prevent_nondicotomic=True, preserve_branch_length=False):
        """Removes this node, reconnecting its children to its parent (ete3 semantics)."""
        parent = self.up
        if parent is None:
            return
        if preserve_branch_length:
            for ch in self.children:
                ch.dist += self.dist
        idx = parent.children.index(self)
        parent.children.pop(idx)
        for ch in self.children:
            ch.up = parent
            parent.children.append(ch)
        self.children = []
        self.up = None
        if prevent_nondicotomic and len(parent.children) == 1 and parent.up is not None:
            parent.delete(prevent_nondicotomic=False, preserve_branch_length=preserve_branch_length)

    # ------------------------------------------------------------------ traversal
    def traverse(self, strategy=LEVELORDER, is_leaf_fn=None):
        """Iterates over the subtree in ete3's order for the given strategy."""
        if strategy == PREORDER:
            return self._iter_preorder()
        if strategy == POSTORDER:
            return self._iter_postorder()
        if strategy == LEVELORDER:
            return self._iter_levelorder()
        raise ValueError('unknown traversal strategy {}'.format(strategy))

    def _iter_levelorder(self):
        queue = deque([self])
        while queue:
            node = queue.popleft()
            yield node
            queue.extend(node.children)

    def _iter_preorder(self):
        stack = [self]
        while stack:
            node = stack.pop()
            yield node
            stack.extend(reversed(node.children))

    def _iter_postorder(self):
        stack = [(self, False)]
        while stack:
            node, expanded = stack.pop()
            if expanded or not node.children:
                yield node
            else:
                stack.append((node, True))
                stack.extend((ch, False) for ch in reversed(node.children))

    def iter_descendants(self, strategy=LEVELORDER):
        it = self.traverse(strategy)
        for node in it:
            if node is not self:
                yield node

    def get_descendants(self, strategy=LEVELORDER):
        return list(self.iter_descendants(strategy))

    def iter_leaves(self):
        for node in self._iter_preorder():
            if not node.children:
                yield node

    def get_leaves(self):
        return list(self.iter_leaves())

    def iter_leaf_names(self):
        for leaf in self.iter_leaves():
            yield leaf.name

    def get_leaf_names(self):
        return [leaf.name for leaf in self.iter_leaves()]

    def iter_ancestors(self):
        node = self.up
        while node is not None:
            yield node
            node = node.up

    def get_ancestors(self):
        return list(self.iter_ancestors())

    def search_nodes(self, **conditions):
        return [n for n in self.traverse()
                if all(getattr(n, k, None) == v for k, v in conditions.items())]

    def get_distance(self, target, target2=None, topology_only=False):
        a, b = (self, target) if target2 is None else (target, target2)
        seen = {}
        d, node = 0., a
        while node is not None:
            seen[id(node)] = d
            d += 1 if topology_only else node.dist
            node = node.up
        d, node = 0., b
        while id(node) not in seen:
            d += 1 if topology_only else node.dist
            node = node.up
        return d + seen[id(node)]

    def __iter__(self):
        return self.iter_leaves()

    def __len__(self):
        return sum(1 for _ in self.iter_leaves())

    def __bool__(self):
        return True

    def __contains__(self, item):
        if isinstance(item, Tree):
            return any(n is item for n in self.traverse())
        return any(n.name == item for n in self.traverse())

    def __and__(self, name):
        hits = [n for n in self.traverse() if n.name == name]
        if not hits:
            raise ValueError('Node not found: {}'.format(name))
        return hits[0]

    def __repr__(self):
        return "Tree node '{}' ({})".format(self.name, hex(id(self)))

    # ------------------------------------------------------------------ editing helpers
    def copy(self, method='cpickle'):
        """Deep copy of the subtree (topology, lengths, names and feature values)."""
        mapping = {}
        root_copy = None
        for node in self._iter_preorder():
            twin = self.__class__()
            twin.dist, twin.support, twin.name = node.dist, node.support, node.name
            twin.features = set(node.features)
            for key, value in node.__dict__.items():
                twin.__dict__[key] = _copy_value(value)
            mapping[id(node)] = twin
            if node is self:
                root_copy = twin
            else:
                parent = mapping[id(node.up)]
                twin.up = parent
                parent.children.append(twin)
        return root_copy

    def set_outgroup(self, outgroup):
        """Re-roots this tree (``self`` must be its root and stays the root object) on the branch above
        ``outgroup``: afterwards the root has two children, ``outgroup`` and the rest of the tree, the split
        branch being shared half/half.  Only the reference's re-rooting invariance tests need this."""
        if outgroup is self:
            raise ValueError('Cannot set the root itself as outgroup')
        path = []  # outgroup.up, ..., self
        node = outgroup.up
        while node is not None:
            path.append(node)
            node = node.up
        if not path or path[-1] is not self:
            raise ValueError('outgroup is not a descendant of this node')

        def gather(nodes):
            # a single node is reused as is, several are hung on a fresh zero-length connector
            if len(nodes) == 1:
                return nodes[0]
            connector = self.__class__()
            connector.dist = 0.
            for nd in nodes:
                nd.up = connector
                connector.children.append(nd)
            return connector

        if len(path) == 1:
            rest = gather([ch for ch in self.children if ch is not outgroup])
        else:
            below_root = path[-2]
            remainder = gather([ch for ch in self.children if ch is not below_root])
            old_dists = [nd.dist for nd in path[:-1]]
            for i in range(len(path) - 2):
                lower, upper = path[i], path[i + 1]
                upper.children.remove(lower)
                lower.children.append(upper)
                upper.up = lower
                upper.dist = old_dists[i]
            below_root.children.append(remainder)
            remainder.up = below_root
            remainder.dist += old_dists[-1]
            rest = path[0]
            rest.children.remove(outgroup)
            rest.dist = 0.
        half = (outgroup.dist + rest.dist) / 2.
        outgroup.dist = rest.dist = half
        outgroup.up = rest.up = self
        self.children = [outgroup, rest]

    # ------------------------------------------------------------------ output
    def write(self, features=None, outfile=None, format=0, is_leaf_fn=None, format_root_node=False,
              dist_formatter='%g', support_formatter='%g', name_formatter='%s', quoted_node_names=False):
        """Newick string of the subtree.  ``format`` follows ete3's numbering for the common cases:
        0/2 = leaf names + support on internal nodes, 1/3 = names everywhere, 5 = leaf names and all
        lengths, 9 = leaf names only, 100 = topology only."""
        leaf_name = format != 100
        leaf_dist = format in (0, 1, 2, 3, 4, 5, 7)
        int_label = {0: 'support', 2: 'support', 1: 'name', 3: 'name', 7: 'name', 8: 'name'}.get(format)
        int_dist = format in (0, 1, 2, 3, 5, 6)

        def label(node):
            out = ''
            if not node.children:
                if leaf_name:
                    out += _fmt_name(node.name, name_formatter, quoted_node_names)
                if leaf_dist:
                    out += ':' + dist_formatter % node.dist
            elif node is not self or format_root_node:
                if int_label == 'name':
                    out += _fmt_name(node.name, name_formatter, quoted_node_names)
                elif int_label == 'support':
                    out += support_formatter % node.support
                if int_dist:
                    out += ':' + dist_formatter % node.dist
            if features is not None:
                keys = sorted(k for k in node.features if k not in ('dist', 'support', 'name')) \
                    if not features else [k for k in features if hasattr(node, k)]
                if keys:
                    out += '[&&NHX:' + ':'.join('{}={}'.format(k, _fmt_feature(getattr(node, k))) for k in keys) + ']'
            return out

        pieces = []
        stack = [(self, False)]
        while stack:
            node, closing = stack.pop()
            if node is None:
                pieces.append(',')
            elif closing:
                pieces.append(')' + label(node))
            elif not node.children:
                pieces.append(label(node))
            else:
                pieces.append('(')
                stack.append((node, True))
                for i in range(len(node.children) - 1, -1, -1):
                    stack.append((node.children[i], False))
                    if i:
                        stack.append((None, False))
        text = ''.join(pieces) + ';'
        if outfile is not None:
            with open(outfile, 'w') as f:
                f.write(text)
            return None
        return text


TreeNode = Tree


# characters that cannot appear in an unquoted newick label or an NHX value: ete3 (which the reference writes its
# trees with) replaces each of them by an underscore, so that the files stay parseable
_ILLEGAL_NEWICK = re.compile(r'[:;(),\[\]\t\n\r=]')


def newick_safe(text):
    return _ILLEGAL_NEWICK.sub('_', str(text))


def _fmt_name(name, formatter, quoted):
    text = formatter % (name if name is not None else '')
    if quoted:
        return '"{}"'.format(text)
    return newick_safe(text)


def _fmt_feature(value):
    if isinstance(value, (set, frozenset, list, tuple)):
        return newick_safe('|'.join(str(_) for _ in sorted(value, key=str)))
    return newick_safe(value)


def _copy_value(value):
    if isinstance(value, set):
        return set(value)
    if isinstance(value, list):
        return list(value)
    if isinstance(value, dict):
        return dict(value)
    if isinstance(value, np.ndarray):
        return value.copy()
    return value


# ---------------------------------------------------------------------- newick reading

def _parse_newick_into(root, text, format=0):
    """Fills ``root`` (a fresh node) from a newick string."""
    text = text.strip()
    if not text:
        raise NewickError('Empty newick string')
    if text.endswith(';'):
        text = text[:-1]
    if not text:
        raise NewickError('Empty newick string')
    n = len(text)
    pos = 0
    node = root
    node.dist = 0.
    depth = 0
    expecting_label = False  # right after ')' we read the label of the node we just closed

    def read_label(pos, target):
        # label := [name][:dist][comment]* in any interleaving ete3 accepts
        name_chars = []
        seen_colon = False
        num_chars = []
        while pos < n:
            c = text[pos]
            if c in '(),;':
                break
            if c == '[':
                end = text.find(']', pos)
                if end < 0:
                    raise NewickError('Unclosed comment in newick')
                _apply_comment(target, text[pos + 1:end])
                pos = end + 1
                continue
            if c in '"\'' and not seen_colon and not name_chars:
                end = text.find(c, pos + 1)
                if end < 0:
                    raise NewickError('Unclosed quote in newick')
                name_chars.append(text[pos + 1:end])
                pos = end + 1
                continue
            if c == ':' and not seen_colon:
                seen_colon = True
                pos += 1
                continue
            (num_chars if seen_colon else name_chars).append(c)
            pos += 1
        label = ''.join(name_chars).strip()
        if label:
            if format in (0, 2) and target.children:
                try:
                    target.support = float(label)
                except ValueError:
                    target.name = label
            else:
                target.name = label
        if seen_colon:
            num = ''.join(num_chars).strip()
            if num:
                try:
                    target.dist = float(num)
                except ValueError:
                    raise NewickError('Bad branch length "{}"'.format(num))
        return pos

    if text[0] != '(':
        # a single-node tree: "A;" or "A:1;"
        pos = read_label(0, root)
        if pos != n:
            raise NewickError('Unexpected newick format "{}"'.format(text[:50]))
        return

    while pos < n:
        c = text[pos]
        if c == '(':
            if expecting_label:
                raise NewickError('Unexpected "(" in newick at {}'.format(pos))
            child = root.__class__()
            child.up = node
            node.children.append(child)
            node = child
            depth += 1
            pos += 1
            # a leaf may follow directly
            if pos < n and text[pos] not in '(),':
                pos = read_label(pos, node)
        elif c == ',':
            if depth == 0:
                raise NewickError('Comma outside parentheses in newick')
            parent = node.up
            child = root.__class__()
            child.up = parent
            parent.children.append(child)
            node = child
            expecting_label = False
            pos += 1
            if pos < n and text[pos] not in '(),':
                pos = read_label(pos, node)
        elif c == ')':
            if depth == 0:
                raise NewickError('Unbalanced parentheses in newick')
            node = node.up
            depth -= 1
            pos += 1
            had_dist = node.dist
            if node is root:
                node.dist = 0.
            pos = read_label(pos, node)
            expecting_label = False
        elif c.isspace():
            pos += 1
        else:
            raise NewickError('Unexpected character "{}" in newick at {}'.format(c, pos))
    if depth != 0 or node is not root:
        raise NewickError('Unbalanced parentheses in newick')


def _apply_comment(node, comment):
    """Stores ``[&&NHX:k=v:k=v]`` / ``[&k=v,k=v]`` comment fields as node features."""
    if comment.startswith('&&NHX:'):
        body, sep = comment[6:], ':'
    elif comment.startswith('&'):
        body, sep = comment[1:], ','
    else:
        return
    for field in _split_outside_braces(body, sep):
        if '=' in field:
            key, value = field.split('=', 1)
            key = key.strip()
            if key in ('dist', 'support'):
                try:
                    setattr(node, key, float(value))
                except ValueError:
                    pass
            elif key:
                node.add_feature(key, value.strip().strip('"'))


def _split_outside_braces(body, sep):
    out, cur, level = [], [], 0
    for c in body:
        if c == '{':
            level += 1
        elif c == '}':
            level -= 1
        if c == sep and level == 0:
            out.append(''.join(cur))
            cur = []
        else:
            cur.append(c)
    out.append(''.join(cur))
    return out


def read_tree(tree_path, columns=None):
    """Reads one tree from a newick string or file (drop-in for ``pastml.tree.read_tree``,
    ``/root/reference/pastml/tree.py:202-222``): labels on internal nodes are node names, and
    per-column NHX annotations are split on ``|`` into state sets."""
    try:
        tree = Tree(tree_path, format=1)
    except NewickError:
        raise ValueError('Could not read the tree {}. Is it a valid newick?'.format(tree_path))
    if columns:
        for node in tree.traverse():
            for c in columns:
                vs = set(getattr(node, c).split('|')) if hasattr(node, c) else set()
                if vs:
                    node.add_feature(c, vs)
    return tree


def read_forest(tree_path, columns=None):
    """Reads all the newick trees of a file (``/root/reference/pastml/tree.py:176-199``; nexus input is out of
    scope here).  Negative branch lengths are set to zero like the reference does."""
    with open(tree_path, 'r') as f:
        nwks = f.read().replace('\n', '').split(';')
    roots = [read_tree(nwk + ';', columns) for nwk in nwks[:-1] if nwk.strip()]
    if not roots:
        raise ValueError('Could not find any trees (in newick format) in the file {}.'.format(tree_path))
    for root in roots:
        for node in root.traverse():
            if node.dist < 0:
                node.dist = 0
    return roots


def name_tree(tree, suffix=''):
    """Gives unique names to unnamed / ambiguously named nodes, with the naming rule of
    ``pastml.tree.name_tree`` (``/root/reference/pastml/tree.py:76-105``): preorder counter, ``root``,
    ``n<i>`` for internal nodes and ``t<i>`` for tips."""
    seen = {}
    total = 0
    for node in tree.traverse():
        total += 1
        if node.name:
            seen[node.name] = seen.get(node.name, 0) + 1
            if '.polytomy_' in node.name:
                node.add_feature('polytomy', 1)
    if total == len(seen):
        return
    counter = 0
    taken = set()
    for node in tree.traverse(PREORDER):
        if node.name and seen.get(node.name, 0) < 10:
            prefix = node.name
        else:
            prefix = '{}{}{}'.format('t' if node.is_leaf() else 'n', counter, suffix)
        candidate = 'root{}'.format(suffix) if node.is_root() else prefix
        while candidate is None or candidate in taken:
            candidate = '{}{}{}'.format(prefix, counter, suffix)
            counter += 1
        node.name = candidate
        taken.add(candidate)


def collapse_zero_branches(forest, features_to_be_merged=None):
    """Removes internal nodes hanging on zero-length branches, merging the listed set-valued features
    (same outcome as ``/root/reference/pastml/tree.py:108-145``; used by the state-count tests)."""
    features_to_be_merged = features_to_be_merged or []
    for tree in forest:
        for node in list(tree.traverse(POSTORDER)):
            doomed = [ch for ch in node.children if ch.children and ch.dist <= 0]
            if not doomed:
                continue
            for feature in features_to_be_merged:
                sets = [getattr(ch, feature, set()) for ch in doomed]
                common = set.intersection(*sets) & getattr(node, feature, set())
                value = common if common else (set.union(*sets) | getattr(node, feature, set()))
                if value:
                    node.add_feature(feature, value)
            for ch in doomed:
                node.remove_child(ch)
                for grandchild in list(ch.children):
                    node.add_child(grandchild)


# ---------------------------------------------------------------------- flattening

class FlatForest(object):
    """Level-ordered CSR view of a forest: the engine's only view of the topology.

    Node ids follow a breadth-first order over the *whole forest* (all roots first, then all depth-1
    nodes, ...), so that (i) a depth level is one contiguous id range, (ii) the children of a node are
    contiguous, and (iii) the children of consecutive internal nodes are consecutive.  For a single tree
    this is exactly ete3's default ``traverse()`` order, which is also the row order of the reference's
    marginal-probability table (``/root/reference/pastml/ml.py:498-502``).

    Attributes (numpy, C-contiguous):
      parent[int32 N]        id of the parent, -1 for roots
      first_child[int32 N]   id of the first child (children are first_child .. first_child+n_children-1)
      n_children[int32 N]
      dist[float64 N]        branch length above the node (as stored; roots usually 0)
      level_ptr[int32 L+1]   node-id range of every depth level
      tree_id[int32 N]       index of the tree in the forest
      slot[int32 N]          rank of the node among internal nodes (same order), -1 for tips
      int_nodes[int32 NI]    node id of every internal node ("slot" -> node)
      int_level_ptr[int32 L+1]  slot range of every depth level
      roots[int32 T]
      nodes                  list of the node objects in id order (None for array-built forests)
      ref_order[int64 N]     ids listed tree by tree in levelorder: the row order of reference outputs
    """

    def __init__(self, parent, dist, nodes=None, names=None):
        parent = np.ascontiguousarray(parent, dtype=np.int32)
        n = len(parent)
        self.n_nodes = n
        self.parent = parent
        self.dist = np.ascontiguousarray(dist, dtype=np.float64)
        self.nodes = nodes
        self._names = names
        if n == 0:
            raise ValueError('Empty forest')
        if np.any(parent[1:] < parent[:-1]) or np.any(parent >= np.arange(n)):
            raise ValueError('Nodes are not in level order (parents must be non-decreasing and precede children)')
        self.roots = np.flatnonzero(parent < 0).astype(np.int32)
        n_children = np.bincount(parent[parent >= 0], minlength=n).astype(np.int32)
        first_child = np.zeros(n, dtype=np.int32)
        # children ids are contiguous and ordered by parent: an exclusive scan gives the first child
        first_child[:] = len(self.roots) + np.concatenate(([0], np.cumsum(n_children[:-1], dtype=np.int64)))
        self.n_children = n_children
        self.first_child = first_child
        depth = np.zeros(n, dtype=np.int32)
        level_ptr = [0, len(self.roots)]
        while level_ptr[-1] < n:
            lo, hi = level_ptr[-2], level_ptr[-1]
            nxt = hi + int(n_children[lo:hi].sum())
            if nxt == hi:
                raise ValueError('Disconnected nodes in the forest arrays')
            depth[hi:nxt] = len(level_ptr) - 1
            level_ptr.append(nxt)
        self.level_ptr = np.asarray(level_ptr, dtype=np.int32)
        self.depth = depth
        self.n_levels = len(level_ptr) - 1
        internal = n_children > 0
        self.is_tip = ~internal
        self.n_tips = int(self.is_tip.sum())
        self.n_internal = n - self.n_tips
        slot = np.full(n, -1, dtype=np.int32)
        slot[internal] = np.arange(self.n_internal, dtype=np.int32)
        self.slot = slot
        self.int_nodes = np.flatnonzero(internal).astype(np.int32)
        self.int_level_ptr = np.searchsorted(self.int_nodes, self.level_ptr).astype(np.int32)
        tree_id = np.zeros(n, dtype=np.int32)
        tree_id[self.roots] = np.arange(len(self.roots), dtype=np.int32)
        for lvl in range(1, self.n_levels):
            lo, hi = self.level_ptr[lvl], self.level_ptr[lvl + 1]
            tree_id[lo:hi] = tree_id[parent[lo:hi]]
        self.tree_id = tree_id
        self.ref_order = np.argsort(tree_id, kind='stable')

    @property
    def n_trees(self):
        return len(self.roots)

    @property
    def names(self):
        if self._names is None and self.nodes is not None:
            self._names = [node.name for node in self.nodes]
        return self._names

    def max_children(self):
        return int(self.n_children.max())


def flatten_forest(forest):
    """Flattens a tree (or list of trees) of :class:`Tree`-like nodes into a :class:`FlatForest`."""
    if not isinstance(forest, (list, tuple)):
        forest = [forest]
    nodes = []
    parent = []
    frontier = [(root, -1) for root in forest]
    while frontier:
        nxt = []
        for node, par in frontier:
            idx = len(nodes)
            nodes.append(node)
            parent.append(par)
            for ch in node.children:
                nxt.append((ch, idx))
        frontier = nxt
    dist = np.fromiter((node.dist for node in nodes), dtype=np.float64, count=len(nodes))
    return FlatForest(np.asarray(parent, dtype=np.int32), dist, nodes=nodes)


def read_forest_flat(tree_path_or_newick, name_nodes=True):
    """Array-native ingest for large trees: parses the newick file (or string) with the engine library's C++ parser
    (``csrc/pml_newick.cpp``; same rules as :func:`read_forest` / ``pastml.tree.read_forest``,
    ``/root/reference/pastml/tree.py:176-222``) straight into a level-ordered :class:`FlatForest`, without building a
    Python object per node.  ``name_nodes`` applies :func:`name_tree`'s rule (``tree.py:76-105``).  NHX annotations are
    skipped: annotate through the mask arrays (``ml.ColumnBatch(annot_mask=...)``)."""
    import ctypes as C
    import os
    from pastml_b200 import _lib
    lib = _lib.load()
    text = tree_path_or_newick
    if isinstance(text, str) and len(text) < 4096 and '(' not in text and os.path.exists(text):
        with open(text, 'rb') as f:
            data = f.read()
    else:
        data = text.encode() if isinstance(text, str) else bytes(text)
    handle = C.c_void_p()
    if lib.pml_newick_parse(data, len(data), C.byref(handle)) != 0:
        raise ValueError('Could not read the tree: {}. Is it a valid newick?'.format(lib.pml_newick_last_error().decode()))
    try:
        if name_nodes:
            lib.pml_newick_name_nodes(handle, 1)
        n = lib.pml_newick_num_nodes(handle)
        parent = np.empty(n, dtype=np.int32)
        dist = np.empty(n, dtype=np.float64)
        offsets = np.empty(n + 1, dtype=np.int64)
        buf = np.empty(max(1, lib.pml_newick_names_bytes(handle)), dtype=np.uint8)
        lib.pml_newick_fill(handle, parent.ctypes.data_as(C.c_void_p), dist.ctypes.data_as(C.c_void_p),
                            offsets.ctypes.data_as(C.c_void_p), buf.ctypes.data_as(C.c_void_p))
    finally:
        lib.pml_newick_free(handle)
    raw = buf.tobytes()
    cuts = offsets.tolist()
    if raw.isascii():   # byte offsets are character offsets
        text = raw.decode('ascii')
        names = np.array([text[a:b] for a, b in zip(cuts[:-1], cuts[1:])], dtype=object)
    else:
        names = np.array([raw[a:b].decode() for a, b in zip(cuts[:-1], cuts[1:])], dtype=object)
    return FlatForest(parent, dist, names=names)


def forest_from_arrays(parent, dist, names=None):
    """Builds :class:`Tree` objects from level-ordered arrays (the inverse of :func:`flatten_forest`);
    returns the list of roots."""
    nodes = []
    roots = []
    for i, p in enumerate(parent):
        node = Tree()
        node.dist = float(dist[i])
        node.name = '' if names is None else str(names[i])
        nodes.append(node)
        if p < 0:
            roots.append(node)
        else:
            node.up = nodes[p]
            nodes[p].children.append(node)
    return roots


# ---------------------------------------------------------------------------------------- polytomy resolution

IS_POLYTOMY = 'polytomy'
DATE, DATE_CI = 'date', 'date_CI'


def _prediction_key(node, columns, col2state2i):
    """The node's predicted states as one string (state indices joined per column), used to group siblings."""
    return '.'.join('-'.join(str(i) for i in sorted(col2state2i[c][s] for s in getattr(node, c, set())))
                    for c in columns)


def states_are_different(n1, n2, columns):
    """True when the two nodes share no state in at least one column (``tree.py:383-387`` of the reference)."""
    return any(not (getattr(n1, c, set()) & getattr(n2, c, set())) for c in columns)


def group_children_if_needed(n, children, columns, state):
    """Hangs ``children`` (siblings predicted in the same state, different from their parent's) on a new node between
    them and ``n``, at the distance of the closest of them (``tree.py:390-417``).  Returns True if a node was made."""
    if len(children) <= 1:
        return False
    child = min(children, key=lambda _: _.dist)
    if not states_are_different(n, child, columns):
        return False
    dist = child.dist
    pol = n.add_child(dist=dist, name='{}.polytomy_{}'.format(n.name, state))
    pol.add_feature(IS_POLYTOMY, 1)
    c_date = getattr(child, DATE, None)
    pol.add_feature(DATE, c_date)
    n_ci = getattr(n, DATE_CI, None)
    c_ci = getattr(child, DATE_CI, None)
    pol.add_feature(DATE_CI, None if not n_ci or not isinstance(n_ci, list)
                    else [n_ci[0], c_ci[1] if c_ci and isinstance(c_ci, list) and len(c_ci) > 1 else c_date])
    for c in columns:
        pol.add_feature(c, getattr(child, c))
    for c in children:
        n.remove_child(c)
        pol.add_child(c, dist=c.dist - dist)
    return True


def resolve_trees(column2states, forest):
    """Resolves polytomies from the state predictions (``tree.py:344-380``): when a node with more than two children
    has at least two of them in one predicted state that differs from its own, those children get a parent of their
    own.  Returns the number of nodes created."""
    columns = sorted(column2states.keys())
    col2state2i = {c: dict(zip(states, range(len(states)))) for c, states in column2states.items()}
    num_new_nodes = 0
    for tree in forest:
        todo = [tree]
        while todo:
            n = todo.pop()
            todo.extend(n.children)
            if len(n.children) > 2:
                state2children = {}
                for c in n.children:
                    state2children.setdefault(_prediction_key(c, columns, col2state2i), []).append(c)
                if len(state2children) > 1:
                    for state, children in state2children.items():
                        if group_children_if_needed(n, children, columns, state):
                            num_new_nodes += 1
    logging.getLogger('pastml').debug('Created {} new internal nodes while resolving polytomies'.format(num_new_nodes)
                                      if num_new_nodes else 'Could not resolve any polytomy')
    return num_new_nodes


def unresolve_trees(column2states, forest):
    """Removes the polytomy resolutions whose state no longer matches their children after the likelihoods were
    recalculated (``tree.py:420-489``).  Returns the number of nodes removed.  (The reference also means to regroup
    the freed children under their grandparent, but collects the candidates in a dictionary it never fills, so no
    node is ever created there; the behaviour, not the intent, is reproduced.)"""
    columns = sorted(column2states.keys())
    col2state2i = {c: dict(zip(states, range(len(states)))) for c, states in column2states.items()}
    num_removed_nodes = 0
    for tree in forest:
        for n in list(tree.traverse('postorder')):
            if not getattr(n, IS_POLYTOMY, False):
                continue
            children = list(n.children)
            keys = {_prediction_key(c, columns, col2state2i) for c in children}
            parent = n.up
            if len(keys) == 1 and not states_are_different(n, children[0], columns):
                # still a consistent resolution; useless only if the parent is in the same state
                if states_are_different(n, parent, columns):
                    continue
            num_removed_nodes += 1
            for c in children:
                parent.add_child(c, dist=c.dist + n.dist)
            parent.remove_child(n)
    if num_removed_nodes:
        logging.getLogger('pastml').debug('Removed {} polytomy resolution{} as inconsistent with model parameters.'
                                          .format(num_removed_nodes, 's' if num_removed_nodes > 1 else ''))
    return num_removed_nodes

"""Subtree sharding of ONE big tree across GPUs (second case).

The tree is cut at a depth level with enough clades; each rank owns a balanced set of clades as a forest problem,
and the part of the tree above the cut (the "top tree", a few thousand nodes at most) is replicated on every rank,
seeing the clade roots as tips whose bottom-up vectors come from outside.  One exchange per sweep:

  bottom-up:  every rank prunes its clades  ->  ALL-GATHER of the clade-root vectors ((K + 1) numbers per lane and
              clade)  ->  every rank finishes the top tree redundantly, so every rank holds the log-likelihood;
  top-down:   every rank runs the top tree's top-down sweep itself (it has all the inputs), which leaves the
              top-down vectors of the clade roots; each rank hands its own to its clade forest and continues
              downwards -- no second collective.

Only marginal passes are sharded (objective evaluations and the BU + TD / marginal / MAP / MPPA pipeline without
the forced joint state); the exchange payload is tiny (latency-, not bandwidth-bound), so it travels as one
``all_gather`` of a padded tensor (NCCL on the GPU box, gloo in CPU tests).
"""
import numpy as np

from pastml_b200.engine import Problem
from pastml_b200.tree import FlatForest


class SubtreePlan(object):
    """Where a single tree is cut and who owns which clade.

    The cut follows the sizes, not a depth level: starting from the root, the largest remaining clade is split into
    its children until no clade is larger than ``N / (world * clades_per_rank)`` (Yule / coalescent trees are
    unbalanced: one child of the root may hold half of the tips).  A clade root is never placed below a zero-length
    branch, which would split a zero-branch cluster between two problems.

    Attributes: ``top_ids`` (global ids of the replicated top tree, increasing = level order), ``boundary`` (global ids
    of the clade roots, which are tips of the top tree), ``boundary_top`` (their ids inside the top tree),
    ``owner_of_clade`` [len(boundary)], and per rank ``nodes[r]`` (global ids of its forest in the forest's level
    order), ``roots_pos[r]`` (positions of its clade roots inside ``boundary``).
    """

    def __init__(self, flat, world_size, clades_per_rank=8):
        import heapq
        if flat.n_trees != 1:
            raise ValueError('subtree sharding takes one tree (forests shard by tree)')
        self.flat = flat
        self.world = int(world_size)
        n = flat.n_nodes
        internal = flat.n_children > 0
        # subtree sizes, accumulated level by level from the bottom
        size = np.ones(n, dtype=np.int64)
        for lvl in range(flat.n_levels - 1, 0, -1):
            a, b = flat.level_ptr[lvl], flat.level_ptr[lvl + 1]
            np.add.at(size, flat.parent[a:b], size[a:b])
        target = max(n // (self.world * clades_per_rank), 1)
        in_top = np.zeros(n, dtype=bool)
        heap = [(-int(size[0]), 0)] if internal[0] else []
        in_top[0] = True
        final = []
        while heap:
            neg, v = heapq.heappop(heap)
            must_split = flat.dist[v] == 0 or v == 0
            enough = len(final) + len(heap) + 1 >= self.world
            if -neg <= target and enough and not must_split:
                final.append(v)
                continue
            # v stays in the top tree as an internal node; its children become candidates (or top-tree tips)
            fc, nc = int(flat.first_child[v]), int(flat.n_children[v])
            splittable = any(internal[c] for c in range(fc, fc + nc))
            if not splittable and not must_split and enough:
                final.append(v)
                continue
            in_top[fc:fc + nc] = True
            for c in range(fc, fc + nc):
                if internal[c]:
                    heapq.heappush(heap, (-int(size[c]), c))
        if len(final) < self.world:
            raise ValueError('the tree cannot be cut into {} clades'.format(self.world))
        self.boundary = np.array(sorted(final), dtype=np.int32)
        self.top_ids = np.flatnonzero(in_top).astype(np.int64)
        top_pos = np.full(n, -1, dtype=np.int64)
        top_pos[self.top_ids] = np.arange(len(self.top_ids))
        self.boundary_top = top_pos[self.boundary].astype(np.int32)
        self._top_pos = top_pos
        self.n_top = len(self.top_ids)

        # longest-processing-time assignment of clades to ranks
        order = np.argsort(-size[self.boundary], kind='stable')
        load = np.zeros(self.world, dtype=np.int64)
        self.owner_of_clade = np.zeros(len(self.boundary), dtype=np.int32)
        for i in order:
            r = int(np.argmin(load))
            self.owner_of_clade[i] = r
            load[r] += size[self.boundary[i]]
        self.load = load

        # owner and depth-below-its-clade-root of every node below the cut
        owner = np.full(n, -1, dtype=np.int32)
        rel_depth = np.zeros(n, dtype=np.int32)
        owner[self.boundary] = self.owner_of_clade
        for lvl in range(1, flat.n_levels):
            a, b = flat.level_ptr[lvl], flat.level_ptr[lvl + 1]
            par = flat.parent[a:b]
            inherit = (owner[par] >= 0) & (owner[a:b] < 0)
            idx = np.arange(a, b)[inherit]
            owner[idx] = owner[flat.parent[idx]]
            rel_depth[idx] = rel_depth[flat.parent[idx]] + 1
        self.nodes, self.roots_pos, self._forests = [], [], [None] * self.world
        for r in range(self.world):
            ids = np.flatnonzero(owner == r).astype(np.int64)
            ids = ids[np.lexsort((ids, rel_depth[ids]))]      # the forest's own level order: (depth in clade, id)
            self.nodes.append(ids)
            self.roots_pos.append(np.flatnonzero(self.owner_of_clade == r).astype(np.int64))

    def top_forest(self):
        """The replicated part: everything above the cut plus the clade roots (as tips), in increasing id."""
        par = self.flat.parent[self.top_ids].astype(np.int64)
        new_par = np.where(par < 0, -1, self._top_pos[np.maximum(par, 0)]).astype(np.int32)
        return FlatForest(new_par, self.flat.dist[self.top_ids])

    def clade_forest(self, rank):
        """The rank's clades as a level-ordered forest (roots = its clade roots, in boundary order)."""
        if self._forests[rank] is None:
            ids = self.nodes[rank]
            new_id = np.full(self.flat.n_nodes, -1, dtype=np.int64)
            new_id[ids] = np.arange(len(ids))
            par = self.flat.parent[ids].astype(np.int64)
            is_root = np.isin(ids, self.boundary)
            new_par = np.where(is_root, -1, new_id[np.where(is_root, 0, par)]).astype(np.int32)
            self._forests[rank] = FlatForest(new_par, self.flat.dist[ids])
        return self._forests[rank]


class ShardedProblem(object):
    """One rank's share of a sharded tree: its clade forest and the replicated top tree, both on one device."""

    def __init__(self, plan, rank, n_states, model_kind, annot_mask, has_state, max_param_sets=1, device=0):
        self.plan, self.rank = plan, rank
        annot_mask = np.atleast_2d(annot_mask)
        has_state = np.atleast_2d(has_state)
        ids = plan.nodes[rank]
        self.clades = Problem(plan.clade_forest(rank), n_states, model_kind, annot_mask[:, ids], has_state[:, ids],
                              max_param_sets=max_param_sets, device=device)
        self.top = Problem(plan.top_forest(), n_states, model_kind, annot_mask[:, plan.top_ids],
                           has_state[:, plan.top_ids], max_param_sets=max_param_sets, device=device)
        self.K, self.C = int(n_states), annot_mask.shape[0]

    def close(self):
        self.clades.close()
        self.top.close()

    # ---- bottom-up
    def local_root_partials(self, params, n_sets=1):
        """Prunes this rank's clades; returns (vec [n_mine][K][J], ex [n_mine][J])."""
        self.clades.loglik(params, n_sets=n_sets, tolerate_zero=True)
        return self.clades.root_partials(n_sets)

    def assemble(self, gathered):
        """[(vec, ex) of rank 0, of rank 1, ...] -> arrays in boundary order."""
        n = len(self.plan.boundary)
        J = gathered[0][0].shape[2]
        vec = np.empty((n, self.K, J), dtype=np.float64)
        ex = np.empty((n, J), dtype=np.int32)
        for r, (v, e) in enumerate(gathered):
            pos = self.plan.roots_pos[r]
            vec[pos] = v[:len(pos)]
            ex[pos] = e[:len(pos)]
        return vec, ex

    def finish_loglik(self, params, vec, ex, n_sets=1):
        """Log-likelihoods [n_sets][C] of the whole tree given all clade-root vectors."""
        self.top.set_tip_partials(self.plan.boundary_top, vec, ex, n_sets)
        return self.top.loglik(params, n_sets=n_sets)

    # ---- fused top-down sweep
    def marginal(self, params, vec, ex, want_probs=True, reuse_bottom_up=False, want_masks=True,
                 results_on_device=False):
        """Top tree sweep on every rank, then this rank's clades.  Returns the rank's pieces: dictionaries for the
        top tree (global node ids ``plan.top_ids``) and for its clade forest (global node ids in ``plan.nodes[rank]``)."""
        self.top.set_tip_partials(self.plan.boundary_top, vec, ex, 1)
        top = self.top.marginal(params, force_joint=False, want_probs=want_probs, want_masks=want_masks,
                                results_on_device=results_on_device)
        td_vec, td_ex = self.top.tip_td()
        pos = self.plan.roots_pos[self.rank]
        self.clades.set_root_td(td_vec[pos], td_ex[pos])
        # the clades' bottom-up state is the one local_root_partials() just computed
        mine = self.clades.marginal(params, force_joint=False, want_probs=want_probs, want_masks=want_masks,
                                    reuse_bottom_up=reuse_bottom_up, results_on_device=results_on_device)
        return top, mine


class DeviceShard(object):
    """One rank's share of a sharded tree with the exchange kept on the device (C ABI ``pml_shard_*``): the clade-root
    runs are packed by a kernel, all-gathered by NCCL on the engine's stream and unpacked into the top tree's tables
    without touching the host; the top-down vectors of the rank's clade roots travel the same way in the other
    direction.  ``connect()`` builds the NCCL communicator (the unique id is broadcast through ``torch.distributed``);
    ranks simulated in one process exchange with :func:`exchange_local` instead."""

    def __init__(self, plan, rank, n_states, model_kind, annot_mask, has_state, max_param_sets=1, device=0):
        import ctypes as C
        from pastml_b200 import _lib
        self.plan, self.rank = plan, rank
        annot_mask = np.atleast_2d(annot_mask)
        has_state = np.atleast_2d(has_state)
        ids = plan.nodes[rank]
        self.clades = Problem(plan.clade_forest(rank), n_states, model_kind, annot_mask[:, ids], has_state[:, ids],
                              max_param_sets=max_param_sets, device=device)
        self.top = Problem(plan.top_forest(), n_states, model_kind, annot_mask[:, plan.top_ids],
                           has_state[:, plan.top_ids], max_param_sets=max_param_sets, device=device)
        self.K, self.C = int(n_states), annot_mask.shape[0]
        self.lib = _lib.load()
        boundary_top = np.ascontiguousarray(plan.boundary_top, dtype=np.int32)
        owner = np.ascontiguousarray(plan.owner_of_clade, dtype=np.int32)
        handle = C.c_void_p()
        rc = self.lib.pml_shard_create(self.clades._h, self.top._h, rank, plan.world, len(boundary_top),
                                       boundary_top.ctypes.data_as(C.c_void_p), owner.ctypes.data_as(C.c_void_p),
                                       C.byref(handle))
        self._h = handle if rc == 0 else None
        self.top._check(rc)
        self.connected = False

    def close(self):
        if getattr(self, '_h', None) is not None:
            self.lib.pml_shard_destroy(self._h)
            self._h = None
        self.clades.close()
        self.top.close()

    def connect(self):
        """Builds the NCCL communicator of the sharded step; needs an initialised ``torch.distributed`` group."""
        import torch
        import torch.distributed as dist
        uid = np.zeros(128, dtype=np.uint8)
        if dist.get_rank() == 0:
            self.top._check(self.lib.pml_shard_unique_id(_ptr(uid)))
        t = torch.from_numpy(uid)
        if dist.get_backend() == 'nccl':
            t = t.cuda()
        dist.broadcast(t, src=0)
        uid = np.ascontiguousarray(t.cpu().numpy())
        self.top._check(self.lib.pml_shard_connect(self._h, _ptr(uid)))
        self.connected = True

    # ---- phases
    def bottom_up(self, params, n_sets=1, mask_kind=0):
        params = self.clades._params(params, n_sets)
        self.clades._check(self.lib.pml_shard_bottom_up(self._h, _ptr(params), n_sets, mask_kind))

    def finish(self, params, n_sets=1, fetch=True):
        params = self.top._params(params, n_sets)
        out = np.empty((n_sets, self.C), dtype=np.float64) if fetch else None
        self.top._check(self.lib.pml_shard_finish(self._h, _ptr(params), n_sets, _ptr(out)))
        return out

    def loglik(self, params, n_sets=1, mask_kind=0):
        """Log-likelihoods [n_sets][C] of the whole tree (identical on every rank); needs ``connect()`` (or one rank)."""
        self.bottom_up(params, n_sets, mask_kind)
        if not self.connected:
            exchange_local([self])
        return self.finish(params, n_sets)

    def top_down(self, want_probs=True, want_node_lh=False, force_joint=False):
        self.top._check(self.lib.pml_shard_top_down(self._h, 1 if want_probs else 0, 1 if want_node_lh else 0,
                                                    1 if force_joint else 0))

    def joint_bottom_up(self, params):
        params = self.clades._params(params, 1)
        self.clades._check(self.lib.pml_shard_joint_bottom_up(self._h, _ptr(params)))

    def joint_finish(self, params):
        params = self.top._params(params, 1)
        out = np.empty((1, self.C), dtype=np.float64)
        self.top._check(self.lib.pml_shard_joint_finish(self._h, _ptr(params), _ptr(out)))
        return out[0]

    def joint(self, params):
        """JOINT reconstruction of the sharded tree: (joint log-likelihoods [C] of the whole tree, states of the top
        tree's nodes [n_top][C], states of this rank's clade nodes [n][C]); needs ``connect()`` (or one rank)."""
        self.joint_bottom_up(params)
        if not self.connected:
            exchange_local([self])
        lh = self.joint_finish(params)
        return lh, self.joint_states(0), self.joint_states(1)

    def joint_states(self, which):
        pb = self.clades if which else self.top
        out = np.empty((pb.N, pb.C), dtype=np.int32)
        pb._check(self.lib.pml_shard_joint_states(self._h, which, _ptr(out)))
        return out

    def results(self, which, want_probs=True, want_node_lh=False, want_masks=True):
        pb = self.clades if which else self.top
        probs = np.empty((pb.N, pb.C, pb.K), dtype=np.float64) if want_probs else None
        node_lh = np.empty((pb.N, pb.C), dtype=np.float64) if want_node_lh else None
        map_mask = np.empty((pb.N, pb.C), dtype=np.uint64) if want_masks else None
        mppa_mask = np.empty((pb.N, pb.C), dtype=np.uint64) if want_masks else None
        k_hist = np.empty((pb.C, pb.K + 1), dtype=np.int64)
        pb._check(self.lib.pml_shard_results(self._h, which, _ptr(probs), _ptr(node_lh), _ptr(map_mask), _ptr(mppa_mask),
                                             _ptr(k_hist)))
        return dict(probs=probs, node_log10lh=node_lh, map_mask=map_mask, mppa_mask=mppa_mask, k_hist=k_hist)

    def marginal(self, params, want_probs=True, want_masks=True, force_joint=False):
        """The whole fixed-parameter pipeline on the sharded tree; returns (log-likelihoods [C], top-tree outputs with
        global node ids ``plan.top_ids``, this rank's clade outputs with global node ids ``plan.nodes[rank]``).
        ``force_joint`` (MPPA keeps the joint state, ml.py:545) needs a preceding :meth:`joint`."""
        loglh = self.loglik(params)[0]
        self.top_down(want_probs=want_probs, force_joint=force_joint)
        return loglh, self.results(0, want_probs, False, want_masks), self.results(1, want_probs, False, want_masks)

    def bench_step(self, params, n_iter, flush_l2=True, phases=False):
        """Mean device time (ms) of n_iter sharded MPPA steps; with ``phases`` also the last step's split."""
        import ctypes as C
        params = self.clades._params(params, 1)
        ms = C.c_double()
        ph = np.zeros(5, dtype=np.float64) if phases else None
        self.clades._check(self.lib.pml_shard_bench_step(self._h, _ptr(params), n_iter, 1 if flush_l2 else 0, C.byref(ms),
                                                         _ptr(ph)))
        if phases:
            names = ('clades_bottom_up', 'exchange', 'top_bottom_up', 'top_top_down', 'clades_top_down')
            return ms.value, {k: float(v) for k, v in zip(names, ph)}
        return ms.value


def exchange_local(shards):
    """Stand-in for the all-gather when the ranks of a sharded tree are simulated in ONE process on one device: call
    between ``bottom_up`` and ``finish`` with every rank's :class:`DeviceShard`."""
    import ctypes as C
    arr = (C.c_void_p * len(shards))(*[sh._h for sh in shards])
    shards[0].top._check(shards[0].lib.pml_shard_exchange_local(arr, len(shards)))


def _ptr(a):
    import ctypes as C
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def all_gather_partials(vec, ex, plan, rank):
    """All ranks' (vec, ex), exchanged with one all_gather of a padded tensor (torch.distributed must be
    initialised; the backend decides the transport: NCCL over NVLink on the GPU box, gloo on CPU)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size()
    n_max = max(len(p) for p in plan.roots_pos)
    K, J = vec.shape[1], vec.shape[2]
    use_cuda = dist.get_backend() == 'nccl'
    payload = np.zeros((n_max, K + 1, J), dtype=np.float64)
    payload[:len(vec), :K, :] = vec
    payload[:len(ex), K, :] = ex          # exponents travel as doubles (exact for |ex| < 2^53)
    t = torch.from_numpy(payload)
    if use_cuda:
        t = t.cuda()
    out = torch.empty((world * n_max,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out, t)   # concatenated along dim 0: the form both NCCL and gloo accept
    out = out.cpu().numpy().reshape((world, n_max) + tuple(t.shape[1:]))
    gathered = []
    for r in range(world):
        n = len(plan.roots_pos[r])
        gathered.append((np.ascontiguousarray(out[r, :n, :K, :]), np.ascontiguousarray(out[r, :n, K, :]).astype(np.int32)))
    return gathered


def merge_outputs(plan, per_rank, key, n_nodes):
    """Stitches per-rank [n_local][C](...) outputs into one [N][C](...) array (top tree from rank 0)."""
    top0, _ = per_rank[0]
    sample = top0[key]
    out = np.empty((n_nodes,) + sample.shape[1:], dtype=sample.dtype)
    out[plan.top_ids] = sample
    for r, (_, mine) in enumerate(per_rank):
        out[plan.nodes[r]] = mine[key]
    return out


class ShardedColumnBatch(object):
    """Drop-in for :class:`pastml_b200.ml.ColumnBatch` inside the lock-step fit when ONE tree is sharded by subtree:
    every objective evaluation is clades bottom-up -> all-gather -> top tree, and returns the same values on every
    rank (the top tree is finished redundantly from identical inputs), so the ranks' host optimisers stay in step
    without any further communication."""

    def __init__(self, plan, rank, characters, models, annot_mask, has_state, device=0, gather=None):
        from pastml_b200.ml import batched_gradients
        self.plan, self.rank = plan, rank
        self.flat = plan.flat
        self.characters, self.models = list(characters), list(models)
        self.C = len(self.characters)
        self.K = len(self.models[0].states)
        self.kind = self.models[0].engine_kind
        self.max_sets = (1 + max(m.get_num_params() for m in self.models)) if batched_gradients() else 1
        self.lane_lists = False   # the sharded objective always runs the (sets x columns) grid
        self.batched_fd = self.max_sets > 1
        self.gather = gather
        if gather is None:
            # device-resident exchange (NCCL on the engine's stream); a single rank needs no communicator
            self.sp = DeviceShard(plan, rank, self.K, self.kind, annot_mask, has_state, max_param_sets=self.max_sets,
                                  device=device)
            if plan.world > 1:
                self.sp.connect()
        else:   # host-staged exchange through a caller-supplied gather (tests)
            self.sp = ShardedProblem(plan, rank, self.K, self.kind, annot_mask, has_state, max_param_sets=self.max_sets,
                                     device=device)
        self.problem = self.sp.top          # LockstepEvaluator reads .stride from here
        self.phase_seconds = {}

    def close(self):
        self.sp.close()

    def params(self):
        return np.stack([m.engine_block() for m in self.models])

    def loglik_sets(self, params, n_sets):
        import time
        if self.gather is None:
            t0 = time.perf_counter()
            out = self.sp.loglik(params, n_sets)
            self.phase_seconds['device_step_s'] = self.phase_seconds.get('device_step_s', 0.) + time.perf_counter() - t0
            return out, {}
        t0 = time.perf_counter()
        vec, ex = self.sp.local_root_partials(params, n_sets)
        t1 = time.perf_counter()
        gathered = self.gather(vec, ex)
        t2 = time.perf_counter()
        vec, ex = self.sp.assemble(gathered)
        out = self.sp.finish_loglik(params, vec, ex, n_sets)
        t3 = time.perf_counter()
        for key, dt in (('clades_bottom_up_s', t1 - t0), ('exchange_s', t2 - t1), ('top_tree_s', t3 - t2)):
            self.phase_seconds[key] = self.phase_seconds.get(key, 0.) + dt
        return out, {}

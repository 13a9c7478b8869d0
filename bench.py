#!/usr/bin/env python
"""Benchmark of the ACR hot path: MPPA node-states/s on a synthetic 1M-tip tree, K=20, 8 columns, F81.

A "step" is one pass of the fixed-parameter MPPA pipeline (metric i) over one forest:
zero-branch alteration, bottom-up (pruning) sweep, fused top-down / marginal / probabilities / MAP / MPPA sweep;
node-states = nodes x states x columns.  ``value`` is timed on the device with CUDA events on resident inputs
(L2 flushed between steps; the working set, 2.7 GB of partials, is far larger than L2 anyway); ``e2e`` goes through
the C ABI with pinned host buffers: problem creation (H2D of topology and annotation masks), the pipeline, and D2H
of the marginal probabilities, MAP/MPPA masks and the k histogram, every step (median step; PCIe- and host-bound).
``roofline`` = algorithmic bytes of the dominant sweep / its CUDA-event time / MEASURED_PEAKS.json, with the ncu DRAM
traffic of the same sweep beside it; ``ml_fit`` = metric ii, the whole ML ACR of the columns (two-stage L-BFGS-B fit on
the host with the device objective, then JOINT + marginal + MAP + MPPA + restricted likelihoods); ``cpu_baseline`` and
``--impl reference`` time the C port of the reference's algorithm (oracle/) on the host cores.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--tips T] [--states K] [--columns C]

With --gpus N > 1 the driver launches one rank per GPU through torchrun; every rank owns an independent
(tree x columns) batch of the same shape (weak scaling, no data-path collective), and ``ml_fit.one_tree_columns_sharded``
is the ML ACR of ONE tree with its columns spread over the ranks (strong scaling, no exchange).  ``--sharding subtree`` instead
cuts ONE tree into clades spread over the ranks, with one NCCL all-gather of clade-root vectors per step (strong
scaling; pastml_b200/sharded.py).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

TRAFFIC_FILE = 'r2_w_dram_traffic.json'   # ncu capture of one step of the default workload with the current kernels
METRIC = 'mppa_node_states_per_s'
UNIT = 'node-states/s'


_REAL_STDOUT = None


def emit(line):
    """Prints the one JSON line of the contract on the process's real stdout."""
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + '\n')
    out.flush()


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--tips', type=int, default=1_000_000)
    ap.add_argument('--states', type=int, default=20)
    ap.add_argument('--columns', type=int, default=8)
    ap.add_argument('--tree', default='yule', choices=['yule', 'coalescent'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--e2e-steps', type=int, default=5)
    ap.add_argument('--no-fit', action='store_true')
    ap.add_argument('--no-configs', action='store_true', help='skip the sub-lines of the other named workloads')
    ap.add_argument('--sharding', default='batch', choices=['batch', 'subtree'],
                    help='batch: every GPU owns an independent (tree x columns) batch (weak scaling, default); '
                         'subtree: ONE tree is cut into clades spread over the GPUs, one all-gather per sweep (strong)')
    ap.add_argument('--sweeps-only', action='store_true', help='development: print the sweep timings and exit')
    return ap.parse_args()


def make_workload(args, rank, want_params=True, tips=None):
    from pastml_b200 import synth
    rng = np.random.default_rng(4000 + rank)
    tips = tips or args.tips
    flat = synth.yule_tree(tips, rng) if args.tree == 'yule' else synth.coalescent_tree(tips, rng)
    masks, has, pis, sfs = synth.simulate_tip_masks(flat, args.states, args.columns, rng)
    params = None
    if want_params:   # (the reference arm never touches the engine)
        from pastml_b200.engine import pack_params
        params = np.stack([pack_params(0, args.states, sfs[c], 0., 1., pis[c]) for c in range(args.columns)])
    return flat, masks, has, pis, sfs, params


def workload_config(args, flat):
    return {'workload': 'synthetic {} tree, {} tips ({} nodes, {} levels), K={} states, {} columns, F81, MPPA pipeline '
                        '(bottom-up + top-down + marginal + MAP + MPPA) at fixed parameters'
                        .format(args.tree, args.tips, flat.n_nodes, flat.n_levels, args.states, args.columns),
            'tips': args.tips, 'nodes': int(flat.n_nodes), 'states': args.states, 'columns': args.columns,
            'model': 'F81', 'method': 'MPPA', 'l2': 'flushed between steps (256 MiB write); inputs >> L2',
            'sharding': 'independent (tree x columns) batch per GPU'}


# ------------------------------------------------------------------ algorithmic bytes

def algorithmic_bytes(flat, K, C):
    """Bytes one sweep must move for C columns batched in one launch sequence, by the formulas of SURVEY.md 8(d):
    B_bu = (2 NI - T)(8K + 8) + 4 Ntip per column + 20 N topology bytes per launch sequence;
    B_td = (2 NI - T)(8K + 8) + 4 Ntip + 2 (NI - T)(8K + 8) + 8K N + 8 N per column + 20 N."""
    n, ni, nt, t = flat.n_nodes, flat.n_internal, flat.n_tips, flat.n_trees
    vec = 8 * K + 8
    bu = C * ((2 * ni - t) * vec + 4 * nt) + 20 * n
    td = C * ((2 * ni - t) * vec + 4 * nt + 2 * (ni - t) * vec + 8 * K * n + 8 * n) + 20 * n
    return bu, td


# ------------------------------------------------------------------ clocks

class ClockSampler(object):
    QUERY = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
            'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thread = None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.QUERY,
                                      '--format=csv,noheader,nounits'], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(',')]
                self.samples.append(float(f[0]))
                self.max_mhz = float(f[1])
                for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[2:]):
                    if val.lower().startswith('active'):
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.02)

    def __enter__(self):
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._thread.join(timeout=10)

    def summary(self):
        return {'sm_mhz': float(np.median(self.samples)) if self.samples else None, 'sm_max_mhz': self.max_mhz,
                'reasons': sorted(self.reasons), 'samples': len(self.samples)}


# ------------------------------------------------------------------ CPU arm (oracle port)

def cpu_pipeline_rate(flat, masks, has, pis, sfs, K, columns, threads, repeats=1, check=None):
    """node-states/s of the oracle's MPPA pipeline on `columns` columns of the workload with `threads` host threads.
    ``check(c, loglh, probs, map_mask, mppa_mask)`` (optional) receives every column's oracle outputs: the parity
    check of the GPU results at benchmark scale rides on the work this leg does anyway."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import cpu_oracle
    cpu_oracle.build()
    cpu_oracle.lib()

    def one(c):
        return cpu_oracle.marginal_pipeline(flat, cpu_oracle.OracleModel(0, pis[c], sfs[c]), masks[c], has[c])

    t0 = time.perf_counter()
    for _ in range(repeats):
        with ThreadPoolExecutor(max_workers=threads) as pool:
            results = list(pool.map(one, columns))
    dt = (time.perf_counter() - t0) / repeats
    if check is not None:
        for c, r in zip(columns, results):
            check(c, *r)
    return flat.n_nodes * K * len(columns) / dt, dt


class ParityAtScale(object):
    """GPU outputs of the benchmark workload against the oracle's on the same tree (tolerances of the north star:
    log-likelihood 1e-10 relative, probabilities 1e-9 absolute, MAP / MPPA state sets and k histogram exact)."""

    def __init__(self, lh, probs, map_mask, mppa_mask, k_hist):
        self.lh, self.probs, self.map_mask, self.mppa_mask, self.k_hist = lh, probs, map_mask, mppa_mask, k_hist
        self.cols, self.max_rel_lh, self.max_abs_prob = [], 0., 0.
        self.map_mismatch = self.mppa_mismatch = self.khist_mismatch = 0

    def __call__(self, c, loglh, probs, map_mask, mppa_mask):
        self.cols.append(int(c))
        self.max_rel_lh = max(self.max_rel_lh, abs(float(self.lh[c]) - loglh) / abs(loglh))
        self.max_abs_prob = max(self.max_abs_prob, float(np.abs(self.probs[:, c, :] - probs).max()))
        self.map_mismatch += int((self.map_mask[:, c].view(np.uint64) != map_mask).sum())
        self.mppa_mismatch += int((self.mppa_mask[:, c].view(np.uint64) != mppa_mask).sum())
        sizes = np.bincount(np.bitwise_count(mppa_mask).astype(np.int64), minlength=self.k_hist.shape[1])
        self.khist_mismatch += int((sizes != self.k_hist[c]).sum())

    def summary(self):
        ok = (self.max_rel_lh <= 1e-10 and self.max_abs_prob <= 1e-9 and
              self.map_mismatch == self.mppa_mismatch == self.khist_mismatch == 0 and len(self.cols) > 0)
        return {'ok': bool(ok), 'columns_checked': self.cols, 'max_rel_dev_loglh': self.max_rel_lh,
                'max_abs_dev_probability': self.max_abs_prob, 'map_state_mismatches': self.map_mismatch,
                'mppa_state_mismatches': self.mppa_mismatch, 'k_hist_mismatches': self.khist_mismatch,
                'against': 'oracle/pastml_oracle.c (CPU restatement pinned on the reference) on the same tree, '
                           'outputs of the last end-to-end step (C ABI, host buffers)'}


REFERENCE_SAMPLE_TIPS = 3_000   # down-scale of the workload the unmodified Python reference is timed on


def _reference_column(job):
    """Worker of the reference arm (one process per column): the UNMODIFIED reference's marginal pipeline on one column."""
    parent, dist, mask, K, pi, sf = job
    import warnings
    warnings.filterwarnings('ignore')
    from oracle import ref_pipeline
    return ref_pipeline.run_column(parent, dist, mask, K, pi, sf)[0]


def reference_python_rate(args, procs, steps, warmup):
    """node-states/s of the unmodified Python reference (oracle/_ref or /root/reference) on a down-scaled tree of the
    workload's recipe: every step runs all the columns, one process per column (the reference's own thread pool over
    characters is bound by the interpreter lock; separate processes are the most it can use).  None if no copy of the
    reference is around."""
    from oracle import ref_pipeline
    if ref_pipeline.reference_root() is None:
        return None
    import multiprocessing as mp
    flat, masks, has, pis, sfs, _ = make_workload(args, 0, want_params=False, tips=REFERENCE_SAMPLE_TIPS)
    jobs = [(flat.parent, flat.dist, masks[c], args.states, pis[c], sfs[c]) for c in range(args.columns)]
    ctx = mp.get_context('spawn')
    with ctx.Pool(processes=procs) as pool:
        for _ in range(warmup):
            pool.map(_reference_column, jobs[:procs])
        t0 = time.perf_counter()
        for _ in range(steps):
            pool.map(_reference_column, jobs)
        dt = (time.perf_counter() - t0) / steps
    return {'value': flat.n_nodes * args.states * args.columns / dt, 'unit': UNIT, 'seconds_per_step': dt, 'cores': procs,
            'kind': 'reference',
            'sample': 'unmodified pastml {} functions (ml.py:697-742) on a {}-tip {} tree of the same recipe ({} nodes), all {} '
                      'columns per step, one process per column ({} processes)'.format(
                          'from ' + os.path.relpath(ref_pipeline.reference_root(), REPO) if ref_pipeline.reference_root().startswith(REPO)
                          else 'from /root/reference', REFERENCE_SAMPLE_TIPS, args.tree, flat.n_nodes, args.columns, procs)}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on the host cores.  The value is the
    UNMODIFIED Python reference (staged under oracle/_ref by oracle/stage_reference.py) on a bounded down-scaled sample
    -- its cost per node does not depend on the tree size (interpreter-bound, ~40 us per node and pass); the C port of
    the same algorithm (oracle/pastml_oracle.c) on the full workload is reported beside it.  Without a copy of the
    reference the port is the value."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, args.columns))
    steps, warm = max(1, args.steps), max(0, args.warmup)
    ref = reference_python_rate(args, procs, steps, min(warm, 2))
    flat, masks, has, pis, sfs, _ = make_workload(args, 0, want_params=False)
    threads = procs
    port_rate, port_dt = cpu_pipeline_rate(flat, masks, has, pis, sfs, args.states, list(range(args.columns)), threads, repeats=1)
    port = {'value': port_rate, 'unit': UNIT, 'cores': threads, 'kind': 'port', 'seconds_per_step': port_dt,
            'sample': 'one step of all {} columns of the full workload, C port of the reference algorithm, {} threads'.format(
                args.columns, threads)}
    main = ref or port
    line = {'impl': 'reference', 'metric': METRIC, 'value': main['value'], 'unit': UNIT, 'n_gpus': args.gpus, 'steps': steps,
            'warmup': warm, 'ms_per_step': main['seconds_per_step'] * 1e3, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic', 'config': workload_config(args, flat),
            'cpu_baseline': main, 'cpu_baseline_port': port,
            'e2e': {'value': main['value'], 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0, 'host_cores': cores}
    emit(line)


# ------------------------------------------------------------------ B200 arm

def full_ml_acr(flat, masks, has, cols, K, device, sync, check_column=None):
    """(fit seconds, reconstruction seconds, device passes, log-likelihood of the first column) of the ML ACR of the
    given columns: the call sequence of ml_acr (ml.py:640-750) on array-native inputs, outputs left on the device."""
    from pastml_b200.annotation import ForestStats
    from pastml_b200.engine import MASK_MAP, MASK_MPPA
    from pastml_b200.ml import ColumnBatch, fit_batch
    from pastml_b200.models.F81Model import F81Model
    fs = ForestStats(flat=flat)
    names = np.array(['s{:02d}'.format(i) for i in range(K)])
    obs = []
    for c in cols:
        f = np.bincount(np.log2(masks[c][flat.is_tip].astype(np.float64)).astype(int), minlength=K).astype(np.float64)
        obs.append(f / f.sum())
    models = [F81Model(states=names, forest_stats=fs, character='c{}'.format(c)) for c in cols]
    batch = ColumnBatch(flat, ['c{}'.format(c) for c in cols], models, device=device,
                        annot_mask=np.ascontiguousarray(masks[cols]), has_state=np.ascontiguousarray(has[cols]))
    sync()
    t0 = time.perf_counter()
    lhs, passes = fit_batch(batch, obs)
    t1 = time.perf_counter()
    params = batch.params()
    pb = batch.problem
    lh_joint, _ = pb.joint(params, want_states=False)
    pb.marginal(params, force_joint=True, want_probs=False, want_masks=False, results_on_device=True)
    lh_map = pb.loglik(params, mask_kind=MASK_MAP)[0]
    _, k_hist = pb.mppa_finalize(force_joint=True, want_masks=False)
    lh_mppa = pb.loglik(params, mask_kind=MASK_MPPA)[0]
    sync()
    t2 = time.perf_counter()
    parity = None
    if check_column is not None:
        # outside the timed region: the oracle replays ml_acr at the FITTED parameters of one column
        from pastml_b200.engine import MASK_JOINT
        from oracle import cpu_oracle
        i = cols.index(check_column)
        m = models[i]
        want = cpu_oracle.ml_acr_fixed(flat, cpu_oracle.OracleModel(0, m.frequencies, float(m.sf)), masks[check_column],
                                       has[check_column], method='MPPA', force_joint=True)
        rel = lambda a, b: abs(float(a) - float(b)) / abs(float(b))
        got_masks = {k: pb.get_masks(kind)[:, i] for k, kind in (('joint', MASK_JOINT), ('map', MASK_MAP), ('mppa', MASK_MPPA))}
        parity = {'column': int(check_column),
                  'rel_dev_loglh': rel(lhs[i], want['log_likelihood']),
                  'rel_dev_loglh_joint': rel(lh_joint[i], want['log_likelihood_restricted_JOINT']),
                  'rel_dev_loglh_map': rel(lh_map[i], want['log_likelihood_restricted_MAP']),
                  'rel_dev_loglh_mppa': rel(lh_mppa[i], want['log_likelihood_restricted_MPPA']),
                  'joint_state_mismatches': int((got_masks['joint'] != want['joint_mask']).sum()),
                  'map_state_mismatches': int((got_masks['map'] != want['map_mask']).sum()),
                  'mppa_state_mismatches': int((got_masks['mppa'] != want['mppa_mask']).sum()),
                  'k_hist_mismatches': int((k_hist[i] != want['k_hist']).sum()),
                  'against': 'oracle ml_acr_fixed at the fitted sf and frequencies of this column'}
        parity['ok'] = bool(max(parity[k] for k in parity if k.startswith('rel_dev')) <= 1e-10 and
                            sum(parity[k] for k in parity if k.endswith('mismatches')) == 0)
    batch.close()
    return t1 - t0, t2 - t1, passes, float(lhs[0]), parity


# ------------------------------------------------------------------ the other named workloads (BASELINE.json configs)

def _sweeps(pb, params, steps):
    """(step ms, bottom-up ms, top-down ms) of the MPPA pipeline on a problem, device-timed, L2 flushed between steps."""
    pb.bench_pass(params, which=1, n_iter=3, flush_l2=True)
    step, _, launches = pb.bench_pass(params, which=1, n_iter=steps, flush_l2=True)
    bu, _, _ = pb.bench_pass(params, which=0, n_iter=steps, flush_l2=True)
    td, _, _ = pb.bench_pass(params, which=3, n_iter=steps, flush_l2=True)
    return step, bu, td, launches


def config3_line(device, sync, steps, check):
    """BASELINE config 3: synthetic 100k-tip Yule tree, K = 10, 16 columns, MPPA + F81 with the full scaling-factor /
    frequency fit on one B200."""
    from pastml_b200 import synth
    from pastml_b200.engine import Problem, pack_params
    tips, K, C = 100_000, 10, 16
    rng = np.random.default_rng(3000)
    flat = synth.yule_tree(tips, rng)
    masks, has, pis, sfs = synth.simulate_tip_masks(flat, K, C, rng)
    params = np.stack([pack_params(0, K, sfs[c], 0., 1., pis[c]) for c in range(C)])
    pb = Problem(flat, K, 0, masks, has, device=device)
    step, bu, td, _ = _sweeps(pb, params, steps)
    # scenario sampling (marginal_counts, ml.py:753-862) of all 16 columns: alteration + bottom-up sweep + one sampling
    # kernel per level, 1000 repetitions, result on the host
    pb.marginal_counts(params, 1000, 1)
    t0 = time.perf_counter()
    counts = pb.marginal_counts(params, 1000, 2)
    counts_ms = (time.perf_counter() - t0) * 1e3
    pb.close()
    fit_s, rec_s, passes, lh0, parity = full_ml_acr(flat, masks, has, list(range(C)), K, device, sync,
                                                    check_column=0 if check else None)
    bu_bytes, td_bytes = algorithmic_bytes(flat, K, C)
    return {'workload': 'synthetic Yule tree, 100k tips ({} nodes, {} levels), K=10, 16 columns, F81, MPPA'.format(
                flat.n_nodes, flat.n_levels),
            'ml_acr_seconds': fit_s + rec_s, 'fit_seconds': fit_s, 'reconstruction_seconds': rec_s, 'device_passes': int(passes),
            'step_ms': step, 'node_states_per_s': flat.n_nodes * K * C / (step * 1e-3), 'bu_ms': bu, 'td_ms': td,
            'algorithmic_GBps': {'bu': bu_bytes / bu / 1e6, 'td': td_bytes / td / 1e6},
            'marginal_counts_1000_repetitions_ms': counts_ms,
            'marginal_counts_changes_per_scenario_column0': float(counts[0].sum() - np.trace(counts[0])),
            'parity_at_fitted_parameters': parity}


def config4_line(device, steps, check, tips=1_000_000):
    """BASELINE config 4: synthetic 1M-tip coalescent tree, K = 20 amino-acid JTT, 8 columns, MPPA (tensor-core sweeps)."""
    from pastml_b200 import synth
    from pastml_b200.annotation import ForestStats
    from pastml_b200.engine import Problem
    from pastml_b200.models.JTTModel import JTTModel
    K, C = 20, 8
    rng = np.random.default_rng(4004)
    flat = synth.coalescent_tree(tips, rng)
    masks, has, pis, sfs = synth.simulate_tip_masks(flat, K, C, rng)
    fs = ForestStats(flat=flat)
    models = [JTTModel(forest_stats=fs, sf=float(sfs[c])) for c in range(C)]
    params = np.stack([m.engine_block() for m in models])
    pb = Problem(flat, K, models[0].engine_kind, masks, has, device=device)
    step, bu, td, launches = _sweeps(pb, params, steps)
    lh = pb.loglik(params)[0]
    pb.joint(params, want_states=False)
    joint_ms, _, _ = pb.bench_pass(params, which=2, n_iter=max(2, steps // 3), flush_l2=True)
    pb.close()
    n_edges = flat.n_nodes - 1
    out = {'workload': 'synthetic coalescent tree, {} tips ({} nodes, {} levels), K=20 JTT, 8 columns, MPPA'.format(
               tips, flat.n_nodes, flat.n_levels),
           'step_ms': step, 'node_states_per_s': flat.n_nodes * K * C / (step * 1e-3), 'bu_ms': bu, 'td_ms': td,
           'launches_per_step': int(launches),
           'joint_ms': joint_ms,   # max-product sweep (P formed on the tensor cores, bu_maxp_mma_kernel) + traceback
           'fp64_tensor_TFLOPs': {'bu': 2 * 4 * K * K * n_edges * C / (bu * 1e-3) / 1e12, 'td': 2 * 2 * K * K * n_edges * C / (td * 1e-3) / 1e12,
                                  'note': 'useful flops of the products A^-1 v and A w (2 K^2 multiply-adds each; two applications of P '
                                          'per edge and lane bottom-up incl. the kept P L_child, one top-down)'},
           'kernels': 'bu_mma_kernel / td_mma_kernel / bu_maxp_mma_kernel (mma.sync.m8n8k4.f64) + td_select_kernel'}
    if check:   # log-likelihood of one column against the oracle's bottom-up pass on the same tree
        from oracle import cpu_oracle
        m = models[0]
        om = cpu_oracle.OracleModel(2, m.frequencies, float(m.sf), d=np.real(m.D_DIAGONAL), A=np.real(m.A), Ainv=np.real(m.A_INV))
        full = np.uint64((1 << K) - 1)
        t0 = time.perf_counter()
        want = cpu_oracle.loglikelihood(cpu_oracle.OracleTree(flat), om, np.where(masks[0] == 0, full, masks[0]), has[0])
        out['parity_loglh_column0'] = {'rel_dev': abs(float(lh[0]) - want) / abs(want), 'oracle_seconds': time.perf_counter() - t0,
                                       'ok': bool(abs(float(lh[0]) - want) / abs(want) <= 1e-10)}
    return out


def config5_line(device, rank, world, steps, check, n_trees=1000, tips=10_000):
    """BASELINE config 5: 1,000 bootstrap-like 10k-tip trees x 8 columns, HKY and F81, MPPA + JOINT, sharded by tree:
    every rank takes n_trees / world trees as ONE forest problem (no exchange); times are per rank, the caller takes
    the maximum."""
    from pastml_b200 import synth
    from pastml_b200.engine import Problem, pack_params
    from pastml_b200.parallel import partition
    C = 8
    lo, hi = partition(n_trees, world, rank)
    rng = np.random.default_rng(5000 + rank)
    flat = synth.forest([synth.yule_tree(tips, rng) for _ in range(hi - lo)])
    out = {'trees_this_rank': hi - lo, 'nodes_this_rank': int(flat.n_nodes), 'levels': int(flat.n_levels)}
    for name, kind, K in (('HKY', 1, 4), ('F81', 0, 10)):
        masks, has, pis, sfs = synth.simulate_tip_masks(flat, K, C, rng)
        if kind == 1:
            params = np.stack([pack_params(1, K, sfs[c], 0., 1., pis[c], kappa=4.) for c in range(C)])
        else:
            params = np.stack([pack_params(0, K, sfs[c], 0., 1., pis[c]) for c in range(C)])
        pb = Problem(flat, K, kind, masks, has, device=device)
        pb.joint(params, want_states=False)
        pb.bench_pass(params, which=4, n_iter=3, flush_l2=True)
        mppa_ms, _, _ = pb.bench_pass(params, which=4, n_iter=steps, flush_l2=True)     # BU + TD with the joint state forced
        joint_ms, _, _ = pb.bench_pass(params, which=2, n_iter=steps, flush_l2=True)    # max-product BU + traceback
        entry = {'K': K, 'mppa_step_ms': mppa_ms, 'joint_ms': joint_ms, 'ms': mppa_ms + joint_ms}
        if check and rank == 0:   # the first tree of the forest against the oracle (MPPA + JOINT, one column)
            from oracle import cpu_oracle
            res = pb.marginal(params, force_joint=True, want_probs=False)
            _, joint_state = pb.joint(params)
            first = np.flatnonzero(flat.tree_id == 0)
            sub = _subforest(flat, first)
            om = cpu_oracle.OracleModel(kind, pis[0], sfs[0], kappa=4.)
            want = cpu_oracle.ml_acr_fixed(sub, om, masks[0][first], has[0][first], method='MPPA', force_joint=True)
            entry['parity_first_tree_column0'] = {
                'joint_state_mismatches': int((joint_state[first, 0] != want['joint_state']).sum()),
                'mppa_state_mismatches': int((res['mppa_mask'][first, 0] != want['mppa_mask']).sum()),
                'map_state_mismatches': int((res['map_mask'][first, 0] != want['map_mask']).sum())}
        pb.close()
        out[name] = entry
    return out


def _subforest(flat, ids):
    """The nodes `ids` (one whole tree of a level-ordered forest, increasing) as a FlatForest of their own."""
    from pastml_b200.tree import FlatForest
    new_id = np.full(flat.n_nodes, -1, dtype=np.int64)
    new_id[ids] = np.arange(len(ids))
    par = flat.parent[ids].astype(np.int64)
    return FlatForest(np.where(par < 0, -1, new_id[np.maximum(par, 0)]).astype(np.int32), flat.dist[ids])


def subtree_sharded_step(args, flat, masks, has, params, rank, world, local_rank, steps, warmup, with_fit):
    """ONE tree cut by subtree over the GPUs (pastml_b200/sharded.py): every rank prunes its clades, the clade-root
    vectors are all-gathered by NCCL on the engine's stream, every rank finishes the replicated top tree and sweeps
    back down into its own clades -- nothing goes through the host inside a step.  Device-timed (CUDA events around
    every step on each rank), max over ranks."""
    import torch.distributed as dist
    from pastml_b200.parallel import max_over_ranks
    from pastml_b200.sharded import SubtreePlan, DeviceShard
    K, C = args.states, args.columns
    plan = SubtreePlan(flat, world)
    sh = DeviceShard(plan, rank, K, 0, masks, has, device=local_rank)
    if world > 1:
        sh.connect()
    sh.bench_step(params, max(warmup, 3))
    dist.barrier()
    ms_local, phases = sh.bench_step(params, steps, phases=True)
    ms = max_over_ranks(ms_local, device='cuda')
    loglh = sh.loglik(params)[0]
    out = {'ms_per_step': ms, 'node_states_per_s': flat.n_nodes * K * C / (ms * 1e-3),
           'clades': int(len(plan.boundary)), 'top_nodes': int(plan.n_top), 'top_levels': int(plan.top_forest().n_levels),
           'nodes_per_rank': [int(len(n)) for n in plan.nodes], 'loglh_column0': float(loglh[0]),
           'phases_ms_last_step_this_rank': phases,
           'exchange': 'one ncclAllGather of {} x (K+1) x C doubles per step on the engine stream, no host hop'.format(
               max(len(p) for p in plan.roots_pos)),
           'timing': 'CUDA events around each step on every rank (L2 flushed between steps), mean over {} steps, max over '
                     'ranks'.format(steps)}
    sh.close()
    if with_fit:
        # the two-stage fit of all the columns on the sharded tree: every rank runs the same host optimiser on values
        # that are identical on all ranks, one all-gather per objective batch
        from pastml_b200.annotation import ForestStats
        from pastml_b200.ml import fit_batch
        from pastml_b200.models.F81Model import F81Model
        from pastml_b200.sharded import ShardedColumnBatch
        fs = ForestStats(flat=flat)
        names = np.array(['s{:02d}'.format(i) for i in range(K)])
        obs = []
        for c in range(C):
            f = np.bincount(np.log2(masks[c][flat.is_tip].astype(np.float64)).astype(int), minlength=K).astype(np.float64)
            obs.append(f / f.sum())
        models = [F81Model(states=names, forest_stats=fs, character='c{}'.format(c)) for c in range(C)]
        batch = ShardedColumnBatch(plan, rank, ['c{}'.format(c) for c in range(C)], models, masks, has, device=local_rank)
        dist.barrier()
        t0 = time.perf_counter()
        lhs, passes = fit_batch(batch, obs)
        dist.barrier()
        out['fit'] = {'seconds': max_over_ranks(time.perf_counter() - t0, device='cuda'), 'device_passes': int(passes),
                      'loglh_column0': float(lhs[0]),
                      'what': 'two-stage L-BFGS-B fit of all {} columns of the ONE tree, clades spread over {} GPUs'.format(C, world)}
        batch.close()
    return out


def run_subtree_sharded(args):
    """--sharding subtree: the subtree-sharded step of ONE tree as the headline line (strong scaling)."""
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    flat, masks, has, pis, sfs, params = make_workload(args, 0)     # the SAME tree on every rank
    K, C = args.states, args.columns
    with ClockSampler(local_rank) as clocks:
        sub = subtree_sharded_step(args, flat, masks, has, params, rank, world, local_rank, args.steps, args.warmup,
                                   not args.no_fit)
    line = {'metric': METRIC, 'value': sub['node_states_per_s'], 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': max(args.warmup, 3), 'ms_per_step': sub['ms_per_step'], 'higher_is_better': True, 'scaling': 'strong',
            'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic', 'config': workload_config(args, flat),
            'clocks': clocks.summary(), 'sharding': sub, 'gpu_launches': None}
    line['config']['sharding'] = 'one tree cut into clades over the GPUs (subtree sharding)'
    if rank == 0:   # outside the timed region: the same tree, unsharded, on this GPU
        from pastml_b200.engine import Problem
        whole = Problem(flat, K, 0, masks, has, device=local_rank)
        ref = whole.loglik(params)[0]
        whole.close()
        line['check_rel_dev_loglh_column0_vs_unsharded'] = abs(sub['loglh_column0'] - float(ref[0])) / abs(float(ref[0]))
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        emit(line)


def run_b200(args):
    import torch
    from pastml_b200.engine import Problem

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise RuntimeError('bench.py needs a CUDA device: the engine has no CPU fallback')
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    flat, masks, has, pis, sfs, params = make_workload(args, rank)
    K, C = args.states, args.columns
    node_states = flat.n_nodes * K * C
    pb = Problem(flat, K, 0, masks, has, max_param_sets=1, device=local_rank)

    # ---- device-resident timing: W warm-up steps, then exactly K steps
    pb.bench_pass(params, which=1, n_iter=max(args.warmup, 3), flush_l2=True)
    barrier()
    with ClockSampler(local_rank) as clocks:
        t_wall0 = time.perf_counter()
        avg_ms, min_ms, launches = pb.bench_pass(params, which=1, n_iter=args.steps, flush_l2=True)
        barrier()
        wall = time.perf_counter() - t_wall0
        # per-sweep split for the roofline, same settings
        bu_ms, _, bu_launches = pb.bench_pass(params, which=0, n_iter=args.steps, flush_l2=True)
        td_ms, _, td_launches = pb.bench_pass(params, which=3, n_iter=args.steps, flush_l2=True)
        kernels_used = pb.last_kernels()   # what the engine launched for those sweeps (name -> launches, items)
    if args.sweeps_only:
        print('sweeps step_ms={:.3f} min={:.3f} bu_ms={:.3f} td_ms={:.3f} launches={} lib={}'.format(
            avg_ms, min_ms, bu_ms, td_ms, launches, os.environ.get('PASTML_B200_LIB', 'default')))
        return
    from pastml_b200.parallel import max_over_ranks
    step_ms = max_over_ranks(avg_ms, device='cuda')
    value = node_states * world / (step_ms * 1e-3)

    # ---- end to end through the C ABI with pinned host buffers
    def pinned(shape, dtype):
        t = torch.empty(shape, dtype=dtype, pin_memory=True)
        return t, t.numpy()

    _, h_parent = pinned(flat.n_nodes, torch.int32)
    _, h_dist = pinned(flat.n_nodes, torch.float64)
    _, h_masks = pinned((C, flat.n_nodes), torch.int64)
    _, h_has = pinned((C, flat.n_nodes), torch.uint8)
    h_parent[:] = flat.parent
    h_dist[:] = flat.dist
    h_masks_u = h_masks.view(np.uint64)
    h_masks_u[:] = masks
    h_has[:] = has
    _, h_probs = pinned((flat.n_nodes, C, K), torch.float64)
    _, h_map = pinned((flat.n_nodes, C), torch.int64)
    _, h_mppa = pinned((flat.n_nodes, C), torch.int64)
    pb.close()
    import ctypes as CT
    from pastml_b200 import _lib
    lib = _lib.load()

    def vp(a):
        return a.ctypes.data_as(CT.c_void_p)

    h_lh = np.empty(C)
    h_hist = np.empty((C, K + 1), dtype=np.int64)

    def e2e_step():
        handle = CT.c_void_p()
        rc = lib.pml_problem_create(flat.n_nodes, vp(h_parent), vp(h_dist), C, K, 0, vp(h_masks), vp(h_has), 1,
                                    local_rank, CT.byref(handle))
        assert rc == 0, _lib.last_error()
        rc = lib.pml_marginal(handle, vp(params), 0, vp(h_lh), vp(h_probs), None, vp(h_map), vp(h_mppa), vp(h_hist))
        assert rc == 0, _lib.last_error()
        lib.pml_problem_destroy(handle)

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    e2e_each = []
    for _ in range(args.e2e_steps):
        t1 = time.perf_counter()
        e2e_step()
        e2e_each.append(time.perf_counter() - t1)
    barrier()
    e2e_mean = max_over_ranks((time.perf_counter() - t0) / args.e2e_steps, device='cuda')
    # the steps are PCIe- and host-bound and the box's other tenants show up as outliers of 2-5x: the median step
    # time (max over ranks) is reported, the mean is given beside it
    e2e_dt = max_over_ranks(float(np.median(e2e_each)), device='cuda')
    e2e_value = node_states * world / e2e_dt
    h2d = h_parent.nbytes + h_dist.nbytes + h_masks.nbytes + h_has.nbytes + params.nbytes
    d2h = h_probs.nbytes + h_map.nbytes + h_mppa.nbytes + h_lh.nbytes + h_hist.nbytes

    # ---- ML-fit wall time (metric ii): two-stage L-BFGS-B fit of every column (host optimiser, device objective)
    # + JOINT + marginal + MAP + MPPA + the restricted-likelihood passes, results left on the device
    ml_fit = None
    if not args.no_fit:
        barrier()
        fit_s, rec_s, passes, lh0, fit_parity = full_ml_acr(
            flat, masks, has, list(range(C)), K, local_rank, torch.cuda.synchronize,
            check_column=0 if (rank == 0 and world == 1 and not args.no_cpu_baseline) else None)
        barrier()
        fit_s = max_over_ranks(fit_s, device='cuda')
        rec_s = max_over_ranks(rec_s, device='cuda')
        ml_fit = {'seconds': fit_s + rec_s, 'fit_seconds': fit_s, 'reconstruction_seconds': rec_s, 'device_passes': int(passes),
                  'parity_at_fitted_parameters': fit_parity,
                  'what': 'two-stage L-BFGS-B fit of sf + {} frequencies for all {} columns of this rank in lock step '
                          '(finite-difference gradients batched as parameter sets), then JOINT + marginal + MAP + MPPA + '
                          'restricted-likelihood passes'.format(K - 1, C)}
        if world > 1:
            # strong scaling of the same job: the C columns of ONE tree (rank 0's workload) spread over the GPUs --
            # columns are independent, so there is no exchange; the slowest rank sets the time
            from pastml_b200.parallel import partition
            flat0, masks0, has0 = (flat, masks, has) if rank == 0 else make_workload(args, 0)[:3]
            lo, hi = partition(C, world, rank)
            mine = list(range(lo, hi))
            barrier()
            if mine:
                f1, r1, p1, _, _ = full_ml_acr(flat0, masks0, has0, mine, K, local_rank, torch.cuda.synchronize)
            else:
                f1 = r1 = 0.
                p1 = 0
            barrier()
            tot = max_over_ranks(f1 + r1, device='cuda')
            ml_fit['one_tree_columns_sharded'] = {
                'seconds': tot, 'columns_per_rank': [partition(C, world, r)[1] - partition(C, world, r)[0] for r in range(world)],
                'device_passes_rank0': int(p1),
                'what': 'the {} columns of rank 0\'s tree fitted and reconstructed on {} GPUs, columns partitioned over the '
                        'ranks (independent, no exchange), max over ranks'.format(C, world)}

    # ---- strong scaling of ONE tree cut by subtree over the GPUs (north star config 4's split), in the same run
    one_tree_subtree = None
    if world > 1:
        flat0, masks0, has0, _, _, params0 = (flat, masks, has, pis, sfs, params) if rank == 0 else make_workload(args, 0)
        barrier()
        one_tree_subtree = subtree_sharded_step(args, flat0, masks0, has0, params0, rank, world, local_rank,
                                                args.steps, args.warmup, with_fit=not args.no_fit)
        one_tree_subtree['speedup_vs_this_runs_single_gpu_step'] = None   # filled below from rank 0's own step time
        if rank == 0:
            whole = Problem(flat0, K, 0, masks0, has0, device=local_rank)
            ref = whole.loglik(params0)[0]
            whole.close()
            one_tree_subtree['check_rel_dev_loglh_column0_vs_unsharded'] = \
                abs(one_tree_subtree['loglh_column0'] - float(ref[0])) / abs(float(ref[0]))
        one_tree_subtree['speedup_vs_this_runs_single_gpu_step'] = step_ms / one_tree_subtree['ms_per_step']
        barrier()

    # ---- the other named workloads (BASELINE.json configs 3-5), as sub-lines; the headline stays the 1M-tip F81 step
    configs = None
    if not args.no_configs:
        check = not args.no_cpu_baseline
        configs = {}
        if world == 1:
            configs['config3_100k_yule_K10_C16_F81_full_fit'] = config3_line(local_rank, torch.cuda.synchronize, args.steps, check)
            configs['config4_1M_coalescent_JTT_C8'] = config4_line(local_rank, max(3, args.steps // 2), check)
        barrier()
        c5 = config5_line(local_rank, rank, world, max(3, args.steps // 2), check)
        barrier()
        for name in ('HKY', 'F81'):
            for key in ('mppa_step_ms', 'joint_ms', 'ms'):
                c5[name][key] = max_over_ranks(c5[name][key], device='cuda')
            nodes_total = 1000 * (2 * 10_000 - 1)
            c5[name]['node_states_per_s_mppa_plus_joint'] = nodes_total * c5[name]['K'] * 8 / (c5[name]['ms'] * 1e-3)
        c5['what'] = ('1000 Yule trees of 10k tips x 8 columns as forest problems, trees partitioned over {} GPU(s) (no '
                      'exchange); per model: fused MPPA step with the joint state forced + JOINT (max-product sweep and '
                      'traceback), device-timed, max over ranks'.format(world))
        configs['config5_1000x10k_forest_C8_HKY_F81_MPPA_JOINT'] = c5

    # ---- roofline of the dominant kernel family (per-level launches of one templated kernel)
    peaks_path = os.path.join(REPO, 'MEASURED_PEAKS.json')
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    else:
        peak, peak_src = 6650., 'fallback (B200_PROFILING.md)'
    bu_bytes, td_bytes = algorithmic_bytes(flat, K, C)
    # the kernel of a sweep = the one the engine reports to have covered most of its nodes (pml_last_kernels): thin
    # levels may run on another kernel of the same sweep, listed under 'kernels_of_sweep'
    bu_name = max(kernels_used['bottom_up'], key=lambda k: kernels_used['bottom_up'][k][1])
    td_name = max(kernels_used['top_down'], key=lambda k: kernels_used['top_down'][k][1])
    sweep_of = {bu_name: 'bottom_up', td_name: 'top_down'}
    sweeps = {td_name: (td_bytes, td_ms, td_launches), bu_name: (bu_bytes, bu_ms, bu_launches)}
    dom = max(sweeps, key=lambda k: sweeps[k][1])

    # measured DRAM traffic of one sweep (ncu dram__bytes_read/write.sum over all its launches), when the committed
    # capture matches this workload
    traffic = {}
    tpath = os.path.join(REPO, 'profiles', TRAFFIC_FILE)
    if os.path.exists(tpath) and (args.tips, K, C, args.tree) == (1_000_000, 20, 8, 'yule'):
        traffic = {k: v['dram_bytes_total'] for k, v in json.load(open(tpath))['kernels'].items()}

    def roof(name):
        b, ms, n = sweeps[name]
        ach = b / (ms * 1e-3) / 1e9
        return {'kernel': name, 'bound': 'hbm', 'achieved': ach, 'peak': peak, 'unit': 'GB/s', 'frac': ach / peak,
                'traffic': sum(traffic[k] for k in kernels_used[sweep_of[name]]) if all(k in traffic for k in kernels_used[sweep_of[name]]) else None,
                'traffic_source': 'profiles/{} (all launches of the sweep)'.format(TRAFFIC_FILE) if name in traffic else None,
                'kernels_of_sweep': {k: {'launches': v[0], 'items': v[1]} for k, v in kernels_used[sweep_of[name]].items()},
                'algorithmic_bytes_per_sweep': int(b), 'launches_per_sweep': int(n),
                'avg_launch_us': ms * 1e3 / max(n, 1), 'sweep_ms': ms, 'peak_source': peak_src}

    line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': max(args.warmup, 3), 'ms_per_step': step_ms, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic', 'config': workload_config(args, flat),
            'clocks': clocks.summary(),
            'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h),
                    'ms_per_step': e2e_dt * 1e3, 'mean_ms_per_step': e2e_mean * 1e3, 'steps': args.e2e_steps, 'min_ms_per_step_rank0': min(e2e_each) * 1e3, 'each_ms_rank0': [round(x * 1e3, 2) for x in e2e_each],
                    'what': 'pml_problem_create + pml_marginal (all outputs to pinned host memory) + destroy; median over the steps'},
            'gpu_launches': int(launches * args.steps),
            'roofline': roof(dom), 'roofline_other': roof([k for k in sweeps if k != dom][0]),
            'min_ms_per_step': min_ms, 'wall_s_timed_region': wall,
            'bu_objective_ms': bu_ms, 'bu_node_states_per_s': node_states / (bu_ms * 1e-3), 'ml_fit': ml_fit,
            'one_tree_subtree': one_tree_subtree, 'configs': configs}

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        threads = max(1, min(cores, C))
        cols = list(range(threads))
        parity = ParityAtScale(h_lh, h_probs, h_map, h_mppa, h_hist)
        rate, dt = cpu_pipeline_rate(flat, masks, has, pis, sfs, K, cols, threads, check=parity)
        line['parity_at_scale'] = parity.summary()
        line['reference_python'] = reference_python_rate(args, threads, 1, 0)   # the unmodified reference, down-scaled sample
        line['cpu_baseline'] = {'value': rate, 'unit': UNIT, 'cores': threads, 'kind': 'port',
                                'sample': 'one step of {} of the {} columns of the same tree, one thread per column '
                                          '({:.1f} s)'.format(len(cols), C, dt), 'host_cores': cores}
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        emit(line)


if __name__ == '__main__':
    # stdout carries exactly one JSON line: native libraries (NCCL's version banner) write to file descriptor 1 behind
    # Python's back, so the descriptor is pointed at stderr for the run and the line goes to a duplicate of the real one
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), 'w')
    os.dup2(2, 1)
    a = parse_args()
    if a.impl == 'reference':
        run_reference(a)
    elif a.sharding == 'subtree' and int(os.environ.get('WORLD_SIZE', '1')) > 1:
        run_subtree_sharded(a)
    else:
        run_b200(a)

"""The N > 1 plumbing (one process per GPU, independent column shards, max-over-ranks timing) exercised with two
gloo ranks on CPU.  The per-shard work is a stand-in: on the GPU box it is acr() on the rank's device."""
import os
import subprocess
import sys
import textwrap

import pytest

from pastml_b200.parallel import partition, shard_columns

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_partition_is_balanced_and_complete():
    for n in (0, 1, 7, 8, 9, 100):
        for world in (1, 2, 3, 8):
            slices = [partition(n, world, r) for r in range(world)]
            assert slices[0][0] == 0 and slices[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(slices, slices[1:]))
            sizes = [hi - lo for lo, hi in slices]
            assert max(sizes) - min(sizes) <= 1
    assert shard_columns(list('abcde'), 2, 0) == ['a', 'b', 'c'] and shard_columns(list('abcde'), 2, 1) == ['d', 'e']


WORKER = textwrap.dedent("""
    import os, sys, json
    sys.path.insert(0, {repo!r})
    import torch.distributed as dist
    from pastml_b200.parallel import sharded_acr, max_over_ranks
    dist.init_process_group('gloo', rank=int(os.environ['RANK']), world_size=int(os.environ['WORLD_SIZE']))
    rank = dist.get_rank()
    columns = ['c%d' % i for i in range(5)]
    seen = []
    def run_shard(cols):
        seen.extend(cols)
        return [dict(character=c, rank=rank) for c in cols]
    merged = sharded_acr(run_shard, columns)
    slowest = max_over_ranks(1.0 + rank)
    print(json.dumps(dict(rank=rank, seen=seen, merged=[m['character'] for m in merged],
                          owners=[m['rank'] for m in merged], slowest=slowest)))
    dist.destroy_process_group()
""")


@pytest.mark.timeout(180)
def test_two_gloo_ranks_shard_columns(tmp_path):
    import json
    script = tmp_path / 'worker.py'
    script.write_text(WORKER.format(repo=REPO))
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE='2', MASTER_ADDR='127.0.0.1', MASTER_PORT='29617')
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.PIPE, text=True))
    outs = []
    for p in procs:
        out, err = p.communicate(timeout=170)
        assert p.returncode == 0, err[-2000:]
        outs.append(json.loads(out.strip().splitlines()[-1]))
    outs.sort(key=lambda o: o['rank'])
    assert outs[0]['seen'] == ['c0', 'c1', 'c2'] and outs[1]['seen'] == ['c3', 'c4']
    for o in outs:
        assert o['merged'] == ['c0', 'c1', 'c2', 'c3', 'c4']
        assert o['owners'] == [0, 0, 0, 1, 1]
        assert o['slowest'] == 2.0


SHARD_WORKER = textwrap.dedent("""
    import os, sys, json
    sys.path.insert(0, {repo!r})
    import numpy as np
    import torch.distributed as dist
    from pastml_b200 import synth
    from pastml_b200.sharded import SubtreePlan, all_gather_partials
    dist.init_process_group('gloo', rank=int(os.environ['RANK']), world_size=int(os.environ['WORLD_SIZE']))
    rank = dist.get_rank()
    flat = synth.yule_tree(500, np.random.default_rng(3))
    plan = SubtreePlan(flat, 2, clades_per_rank=3)
    K, J = 4, 3
    n = len(plan.roots_pos[rank])
    # a recognisable payload: clade-root id in the mantissas, rank in the exponents
    vec = np.zeros((n, K, J)) + plan.boundary[plan.roots_pos[rank]][:, None, None]
    ex = np.full((n, J), rank - 7, dtype=np.int32)
    got = all_gather_partials(vec, ex, plan, rank)
    ok = all(np.all(got[r][0] == plan.boundary[plan.roots_pos[r]][:, None, None]) and np.all(got[r][1] == r - 7)
             and got[r][0].shape == (len(plan.roots_pos[r]), K, J) for r in range(2))
    print(json.dumps(dict(rank=rank, ok=bool(ok), clades=[len(p) for p in plan.roots_pos])))
    dist.destroy_process_group()
""")


@pytest.mark.timeout(180)
def test_two_gloo_ranks_exchange_clade_root_partials(tmp_path):
    """The one collective of the subtree-sharded path (all-gather of clade-root vectors) over gloo."""
    import json
    script = tmp_path / 'shard_worker.py'
    script.write_text(SHARD_WORKER.format(repo=REPO))
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE='2', MASTER_ADDR='127.0.0.1', MASTER_PORT='29618')
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.PIPE, text=True))
    for p in procs:
        out, err = p.communicate(timeout=170)
        assert p.returncode == 0, err[-2000:]
        res = json.loads(out.strip().splitlines()[-1])
        assert res['ok'] and sum(res['clades']) >= 6

"""pml_problem_create through the C ABI: the topology is derived on the device (pml_topo.cu), so its input checks and
its degenerate cases are exercised here -- malformed level orders must come back as PML_ERR_INVALID with the engine's
message, and degenerate forests (no internal node at all) must work.  (Trees with polytomies, forests and zero-length
branches go through the same device pass in every parity test.)"""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _create(parent, dist, K=3, cols=1, masks=None, has=None):
    from pastml_b200 import _lib
    lib = _lib.load()
    parent = np.ascontiguousarray(parent, dtype=np.int32)
    dist = np.ascontiguousarray(dist, dtype=np.float64)
    n = len(parent)
    masks = np.zeros((cols, n), dtype=np.uint64) if masks is None else np.ascontiguousarray(masks, dtype=np.uint64)
    has = np.zeros((cols, n), dtype=np.uint8) if has is None else np.ascontiguousarray(has, dtype=np.uint8)
    handle = C.c_void_p()
    vp = lambda a: a.ctypes.data_as(C.c_void_p)   # noqa: E731
    rc = lib.pml_problem_create(n, vp(parent), vp(dist), cols, K, 0, vp(masks), vp(has), 1, 0, C.byref(handle))
    return lib, rc, handle


@pytest.mark.parametrize('parent,message', [
    ([-1, 0, 0, 4, 1], 'a parent must precede its children'),      # node 3 names a later node as its parent
    ([-1, 0, 1, 0, 1], 'parents must be non-decreasing'),          # siblings not contiguous
    ([0, -1, 0], 'level order'),                                   # a root after a child
    ([-1, 0, 0, 3], 'a parent must precede its children'),         # its own parent
])
def test_malformed_level_orders_are_refused(parent, message):
    from pastml_b200 import _lib
    lib, rc, handle = _create(parent, np.ones(len(parent)))
    assert rc == _lib.PML_ERR_INVALID and not handle.value
    assert message in _lib.last_error()


def test_too_many_states_is_refused():
    from pastml_b200 import _lib
    lib, rc, handle = _create([-1, 0, 0], [0., 1., 1.], K=63)
    assert rc == _lib.PML_ERR_INVALID and '[1, 62]' in _lib.last_error()


def test_degenerate_forests():
    """A single node, a forest of single nodes, a cherry: the derived topology handles them (no internal node at all
    in the first two) and the likelihood is the tip term alone."""
    from pastml_b200.engine import Problem, pack_params
    from pastml_b200.tree import FlatForest
    pi = np.array([.2, .3, .5])
    for parent in ([-1], [-1, -1, -1], [-1, 0, 0]):
        n = len(parent)
        flat = FlatForest(np.array(parent, dtype=np.int32), np.where(np.array(parent) < 0, 0., 0.7))
        masks = np.zeros((1, n), dtype=np.uint64)
        has = np.zeros((1, n), dtype=np.uint8)
        tips = [v for v in range(n) if v not in parent]
        for i, v in enumerate(tips):
            masks[0, v] = 1 << (i % 3)
            has[0, v] = 1
        pb = Problem(flat, 3, 0, masks, has)
        lh = pb.loglik(pack_params(0, 3, 1.3, 0., 1., pi)[None, :])[0]
        pb.close()
        if n == 3 and parent[1] == 0:   # the cherry: sum_i pi_i P_i0 P_i1
            mu = 1. / (1. - pi.dot(pi))
            e = np.exp(-mu * 0.7 * 1.3)
            P = (1 - e) * pi[None, :] + e * np.eye(3)
            want = np.log((pi * P[:, 0] * P[:, 1]).sum())
        else:                           # isolated tips: log pi of their state, summed over the trees
            want = sum(np.log(pi[i % 3]) for i in range(len(tips)))
        assert lh[0] == pytest.approx(want, rel=1e-12)

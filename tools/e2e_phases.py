"""Where the end-to-end step (pml_problem_create -> pml_marginal -> destroy, host buffers) spends its time, next to
the PCIe rate of the box for the same bytes (GPU):  python tools/e2e_phases.py [tips] [columns] [states]"""
import ctypes as CT
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pastml_b200 import synth, _lib  # noqa: E402
from pastml_b200.engine import pack_params  # noqa: E402


def main():
    tips = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
    C = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    K = int(sys.argv[3]) if len(sys.argv) > 3 else 10
    rng = np.random.default_rng(1)
    flat = synth.yule_tree(tips, rng)
    masks, has, pis, sfs = synth.simulate_tip_masks(flat, K, C, rng, missing=0.05)
    params = np.stack([pack_params(0, K, sfs[c], 0., 1., pis[c]) for c in range(C)])
    lib = _lib.load()
    N = flat.n_nodes

    def pinned(shape, dtype):
        t = torch.empty(shape, dtype=dtype, pin_memory=True)
        return t, t.numpy()

    keep = []

    def pin_of(a):
        t, v = pinned(a.shape, torch.from_numpy(a[:0].copy().view(np.int64 if a.dtype == np.uint64 else a.dtype)).dtype)
        keep.append(t)
        v = v.view(a.dtype)
        v[...] = a
        return v

    h_parent, h_dist, h_masks, h_has = pin_of(flat.parent.astype(np.int32)), pin_of(flat.dist), pin_of(masks), pin_of(has)
    t_probs, h_probs = pinned((N, C, K), torch.float64)
    t_map, h_map = pinned((N, C), torch.int64)
    t_mppa, h_mppa = pinned((N, C), torch.int64)
    h_lh = np.empty(C)
    h_hist = np.empty((C, K + 1), dtype=np.int64)
    vp = lambda a: a.ctypes.data_as(CT.c_void_p)   # noqa: E731

    # PCIe reference: the same bytes, plain copies
    d = torch.empty(h_probs.size + 2 * h_map.size, dtype=torch.float64, device='cuda')
    for _ in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        t_probs.copy_(d[:h_probs.size].view(N, C, K), non_blocking=True)
        t_map.copy_(d[h_probs.size:h_probs.size + h_map.size].view(torch.int64).view(N, C), non_blocking=True)
        t_mppa.copy_(d[h_probs.size + h_map.size:].view(torch.int64).view(N, C), non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    nbytes = h_probs.nbytes + h_map.nbytes + h_mppa.nbytes
    print('plain D2H of the outputs: {:.1f} ms ({:.1f} GB/s)'.format(dt * 1e3, nbytes / dt / 1e9))
    dm = torch.empty(h_masks.size, dtype=torch.int64, device='cuda')
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    dm.copy_(torch.from_numpy(h_masks.view(np.int64)).view(-1), non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print('plain H2D of the masks: {:.1f} ms ({:.1f} GB/s)'.format(dt * 1e3, h_masks.nbytes / dt / 1e9))
    del d, dm

    for it in range(6):
        handle = CT.c_void_p()
        t0 = time.perf_counter()
        rc = lib.pml_problem_create(N, vp(h_parent), vp(h_dist), C, K, 0, vp(h_masks), vp(h_has), 1, 0, CT.byref(handle))
        assert rc == 0, _lib.last_error()
        t1 = time.perf_counter()
        rc = lib.pml_marginal(handle, vp(params), 0, vp(h_lh), vp(h_probs), None, vp(h_map), vp(h_mppa), vp(h_hist))
        assert rc == 0, _lib.last_error()
        t2 = time.perf_counter()
        lib.pml_problem_destroy(handle)
        t3 = time.perf_counter()
        print('step {}: create {:.1f} ms, marginal {:.1f} ms, destroy {:.1f} ms, total {:.1f} ms'.format(
            it, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t3 - t0) * 1e3))


if __name__ == '__main__':
    main()

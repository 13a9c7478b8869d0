"""Builds the CUDA engine (``libpastml_b200.so``) in-tree with nvcc for sm_100a.

There is exactly one target architecture: B200.  The templated level kernels are instantiated once per
(model family, padded state count) in separate translation units (``csrc/pml_inst.cu`` compiled with
``-DPML_FAMILY -DPML_KP``), which are compiled in parallel and linked with ``csrc/pml_engine.cu``.  Objects go to
``pastml_b200/_build/`` and the shared library lands next to this file so that it travels with the source tree
(both are git-ignored, not gpurun-ignored).
"""
import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OBJ_DIR = os.path.join(HERE, '_build')
LIB_NAME = 'libpastml_b200.so'
LIB_PATH = os.path.join(HERE, LIB_NAME)
ENGINE_SRC = os.path.join(CSRC, 'pml_engine.cu')
INST_SRC = os.path.join(CSRC, 'pml_inst.cu')
COOP_SRC = os.path.join(CSRC, 'pml_coop_inst.cu')
TPI_SRC = os.path.join(CSRC, 'pml_tpi.cu')
MMA_SRC = os.path.join(CSRC, 'pml_mma.cu')
TOPO_SRC = os.path.join(CSRC, 'pml_topo.cu')
NEWICK_SRC = os.path.join(CSRC, 'pml_newick.cpp')
HEADERS = [os.path.join(CSRC, 'pml_kernels.cuh'), os.path.join(CSRC, 'pml_counts.cuh'), os.path.join(CSRC, 'pml_coop.cuh'), os.path.join(CSRC, 'pml_tpi.cuh'),
           os.path.join(CSRC, 'pml_launch.h'),
           os.path.join(os.path.dirname(HERE), 'include', 'pastml_b200.h')]

# (family, KP): 0 = F81/JC/EFT, 1 = HKY, 2 = eigen-decomposition models (CUSTOM_RATES, JTT); keep in sync with
# PML_ALL in pml_engine.cu
INSTANCES = [(0, k) for k in (2, 4, 8, 12, 16, 20, 24, 32, 48, 62)] + [(1, 4)] + \
            [(2, k) for k in (2, 4, 8, 12, 16, 20, 24, 32, 48, 62)]

# (G, KT) of the warp-cooperative F81 sweeps; keep in sync with PML_COOP_ALL in pml_engine.cu
COOP_INSTANCES = [(2, k, 0) for k in (2, 3, 4, 5, 6, 8, 10, 12, 16)] + [(4, k, 0) for k in (2, 3, 4, 5, 6, 8, 10, 12, 16)] + \
                 [(8, k, 0) for k in (2, 3, 4, 5, 6, 8)] + [(16, k, 0) for k in (2, 3, 4)] + [(32, 2, 0)] + \
                 [(4, 5, 8), (2, 5, 16), (2, 5, 8), (4, 5, 4), (2, 2, 16), (2, 2, 8), (2, 10, 8),
                  (4, 3, 8), (4, 4, 8), (4, 2, 8), (2, 3, 8), (2, 4, 8), (4, 6, 8), (4, 8, 8), (2, 6, 8)]   # compile-time lanes / K

NVCC_FLAGS = ['-O3', '-std=c++17', '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo',
              '-Xcompiler', '-fPIC', '-Xcompiler', '-O2']


def find_nvcc():
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found: the CUDA engine cannot be built')


def _digest(paths, extra=''):
    h = hashlib.sha1(extra.encode())
    for path in paths:
        with open(path, 'rb') as f:
            h.update(f.read())
    return h.hexdigest()


def _compile(nvcc, src, obj, defines, stamp, verbose):
    stamp_file = obj + '.stamp'
    if os.path.exists(obj) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
        return False
    cmd = [nvcc] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + defines + ['-c', src, '-o', obj]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError('nvcc failed for {} {}:\n{}'.format(os.path.basename(src), ' '.join(defines), res.stdout))
    if verbose and res.stdout:
        sys.stdout.write(res.stdout)
    with open(stamp_file, 'w') as f:
        f.write(stamp)
    return True


def build(force=False, verbose=False, jobs=None):
    """Compiles whatever is stale and links the shared library; returns its path."""
    nvcc = find_nvcc()
    os.makedirs(OBJ_DIR, exist_ok=True)
    if force:
        for f in os.listdir(OBJ_DIR):
            os.remove(os.path.join(OBJ_DIR, f))
    flags = ' '.join(NVCC_FLAGS)
    tasks = [(ENGINE_SRC, os.path.join(OBJ_DIR, 'pml_engine.o'), [], _digest([ENGINE_SRC] + HEADERS, flags))]
    tasks.append((TPI_SRC, os.path.join(OBJ_DIR, 'pml_tpi.o'), [], _digest([TPI_SRC] + HEADERS, flags)))
    tasks.append((MMA_SRC, os.path.join(OBJ_DIR, 'pml_mma.o'), [], _digest([MMA_SRC] + HEADERS, flags)))
    tasks.append((TOPO_SRC, os.path.join(OBJ_DIR, 'pml_topo.o'), [], _digest([TOPO_SRC, HEADERS[-2]], flags)))
    tasks.append((NEWICK_SRC, os.path.join(OBJ_DIR, 'pml_newick.o'), [], _digest([NEWICK_SRC, HEADERS[-1]], flags)))
    inst_digest = _digest([INST_SRC] + HEADERS, flags)
    for fam, kp in INSTANCES:
        tasks.append((INST_SRC, os.path.join(OBJ_DIR, 'pml_inst_{}_{}.o'.format(fam, kp)),
                      ['-DPML_FAMILY={}'.format(fam), '-DPML_KP={}'.format(kp)], inst_digest))
    coop_digest = _digest([COOP_SRC] + HEADERS, flags)
    for g, kt, jt in COOP_INSTANCES:
        tasks.append((COOP_SRC, os.path.join(OBJ_DIR, 'pml_coop_{}_{}_{}.o'.format(g, kt, jt)),
                      ['-DPML_G={}'.format(g), '-DPML_KT={}'.format(kt), '-DPML_JT={}'.format(jt)], coop_digest))
    # objects of instantiations that are no longer built must not linger (they would be stale, not linked, and travel)
    wanted = {os.path.basename(t[1]) for t in tasks}
    for f in os.listdir(OBJ_DIR):
        if f.endswith('.o') and f not in wanted or f.endswith('.o.stamp') and f[:-len('.stamp')] not in wanted:
            os.remove(os.path.join(OBJ_DIR, f))
    jobs = jobs or max(1, min(len(tasks), os.cpu_count() or 1))
    # biggest instantiations first so that they do not end up alone at the tail
    def weight(t):
        kp = [d for d in t[2] if d.startswith('-DPML_KP=')]
        return -int(kp[0].split('=')[1]) if kp else 0
    tasks.sort(key=weight)
    with ThreadPoolExecutor(max_workers=jobs) as pool:
        rebuilt = list(pool.map(lambda t: _compile(nvcc, t[0], t[1], t[2], t[3], verbose), tasks))
    if any(rebuilt) or not os.path.exists(LIB_PATH):
        objs = [t[1] for t in tasks]
        subprocess.check_call([nvcc, '--shared', '-gencode', 'arch=compute_100a,code=sm_100a', '-o', LIB_PATH] + objs)
    return LIB_PATH


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))

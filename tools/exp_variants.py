"""Development tool: builds kernel-tuning variants of the engine (K=20 F81 instantiation only) into
pastml_b200/_exp/ so that one gpurun call can time several of them:

    python tools/exp_variants.py name1:-DPML_BU_MINB=3,-DPML_TD_MINB=3 name2:-DPML_BLOCK=256 ...
    gpurun -- 'for v in pastml_b200/_exp/*.so; do PASTML_B200_LIB=$v python bench.py --no-cpu-baseline ...; done'
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(REPO, 'pastml_b200', 'csrc')
OUT = os.path.join(REPO, 'pastml_b200', '_exp')
FLAGS = ['-O3', '-std=c++17', '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-Xcompiler', '-fPIC']


def build_variant(spec):
    name, _, defs = spec.partition(':')
    defs = [d for d in defs.split(',') if d]
    os.makedirs(OUT, exist_ok=True)
    o1, o2 = os.path.join(OUT, name + '_engine.o'), os.path.join(OUT, name + '_inst.o')
    subprocess.check_call(['nvcc'] + FLAGS + defs + ['-DPML_ONLY_F81_20', '-c', os.path.join(CSRC, 'pml_engine.cu'), '-o', o1])
    res = subprocess.run(['nvcc'] + FLAGS + defs + ['-Xptxas', '-v', '-DPML_FAMILY=0', '-DPML_KP=20', '-c',
                                                    os.path.join(CSRC, 'pml_inst.cu'), '-o', o2],
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, check=True)
    lines = res.stdout.splitlines()
    info = []
    for i, line in enumerate(lines):
        if 'Compiling entry function' in line and ('bu_level_kernelILi20ENS_7F81LaneELb0' in line or 'td_level' in line):
            info.append(('bu' if 'bu_level' in line else 'td') + ': ' + ' | '.join(x.replace('ptxas info    : ', '').strip() for x in lines[i + 2:i + 4]))
    objs = [o1, o2]
    for g, kt, jt in ((4, 5, 0), (4, 5, 8), (2, 10, 8)):
        o3 = os.path.join(OUT, name + '_coop{}_{}_{}.o'.format(g, kt, jt))
        objs.append(o3)
        res = subprocess.run(['nvcc'] + FLAGS + defs + ['-Xptxas', '-v', '-DPML_G={}'.format(g), '-DPML_KT={}'.format(kt), '-DPML_JT={}'.format(jt),
                                                        '-c', os.path.join(CSRC, 'pml_coop_inst.cu'), '-o', o3],
                             stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, check=True)
        lines = res.stdout.splitlines()
        for i, line in enumerate(lines):
            if 'Compiling entry function' in line and 'coop_kernel' in line:
                info.append(('bu_coop' if 'bu_coop' in line else 'td_coop') + ' G={} KT={} JT={}: '.format(g, kt, jt) + ' | '.join(
                    x.replace('ptxas info    : ', '').strip() for x in lines[i + 2:i + 4]))
    for src, extra, tag in (('pml_tpi.cu', [], 'tpi'),):
        o4 = os.path.join(OUT, name + '_' + tag + '.o')
        objs.append(o4)
        res = subprocess.run(['nvcc'] + FLAGS + defs + extra + ['-Xptxas', '-v', '-c', os.path.join(CSRC, src), '-o', o4],
                             stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, check=True)
        lines = res.stdout.splitlines()
        for i, line in enumerate(lines):
            if 'Compiling entry function' in line and 'tpi_kernel' in line and 'Li8EEE' in line:
                info.append(('bu_tpi' if 'bu_tpi' in line else 'td_tpi') + ' JT=8: ' + ' | '.join(
                    x.replace('ptxas info    : ', '').strip() for x in lines[i + 2:i + 4]))
    lib = os.path.join(OUT, 'lib_' + name + '.so')
    subprocess.check_call(['nvcc', '--shared', '-gencode', 'arch=compute_100a,code=sm_100a', '-o', lib] + objs)
    for o in objs:
        os.remove(o)
    return name, info


if __name__ == '__main__':
    with ThreadPoolExecutor(max_workers=8) as pool:
        for name, info in pool.map(build_variant, sys.argv[1:]):
            print(name)
            for line in info:
                print('   ', line)

"""TEST INFRASTRUCTURE ONLY -- golden outputs of the reference's ``pastml_pipeline`` (``/root/reference/pastml/acr.py:316``).

For every case: writes the input files (a newick tree and an annotation table, DERIVED from the committed array
fixtures ``tests/golden/*.npz`` -- no reference file is copied), runs the unmodified reference pipeline on them in
this container (through ``oracle/ref_env.py``) and keeps its result files under ``tests/golden/pipeline/<case>/``.
``tests/test_gpu_pipeline.py`` runs ``pastml_b200.pipeline.pastml_pipeline`` on the same inputs and compares.

    python oracle/make_golden_pipeline.py
"""
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, 'tests'))

CASES = {
    # case: (fixture, column name, method, model)
    'albania_mppa_f81': ('albania_f81', 'Country', 'MPPA', 'F81'),
    'albania_map_jc': ('albania_jc', 'Country', 'MAP', 'JC'),
    'albania_joint_eft': ('albania_eft', 'Country', 'JOINT', 'EFT'),
    'albania_downpass': ('albania_f81', 'Country', 'DOWNPASS', 'F81'),
    'denv2_mppa_f81': ('denv2_location_f81', 'Location', 'MPPA', 'F81'),
}


def write_inputs(fixture, column, out_dir):
    """input.nwk / input.tab from the fixture's level-ordered arrays and tip annotation."""
    from helpers import Golden
    from pastml_b200.tree import forest_from_arrays
    g = Golden(fixture)
    roots = forest_from_arrays(g['parent'], g['dist'], g['names'])
    with open(os.path.join(out_dir, 'input.nwk'), 'w') as f:
        f.write('\n'.join(r.write(format=3, format_root_node=True, dist_formatter='%.17g') for r in roots) + '\n')
    states = [str(s) for s in g['states']]
    with open(os.path.join(out_dir, 'input.tab'), 'w') as f:
        f.write('id\t{}\n'.format(column))
        for i, name in enumerate(g['names']):
            if not g['has_state'][i]:
                continue
            mask = int(g['annot_mask'][i])
            mine = [states[k] for k in range(len(states)) if (mask >> k) & 1]
            if len(mine) == 1:   # the table format holds one state per row
                f.write('{}\t{}\n'.format(name, mine[0]))


def main():
    from oracle import ref_env
    ref_env.install()
    from pastml.acr import pastml_pipeline
    base = os.path.join(REPO, 'tests', 'golden', 'pipeline')
    for case, (fixture, column, method, model) in CASES.items():
        out_dir = os.path.join(base, case)
        shutil.rmtree(out_dir, ignore_errors=True)
        os.makedirs(out_dir)
        write_inputs(fixture, column, out_dir)
        work = os.path.join(out_dir, 'expected')
        np.random.seed(0)
        pastml_pipeline(tree=os.path.join(out_dir, 'input.nwk'), data=os.path.join(out_dir, 'input.tab'), columns=[column],
                        prediction_method=method, model=model, work_dir=work, threads=1)
        for name in os.listdir(work):   # the nexus copy is produced by a Bio stand-in here: not a reference output
            if name.endswith('.nexus'):
                os.remove(os.path.join(work, name))
        print(case, sorted(os.listdir(work)))


if __name__ == '__main__':
    main()

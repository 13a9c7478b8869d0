"""TEST INFRASTRUCTURE ONLY -- stages the reference for the GPU box.

The reference is pure Python and `/root/reference` does not exist on the GPU box; `gpurun` ships everything under the
repository that is not gpurun-ignored, and `oracle/_ref/` is git-ignored but NOT gpurun-ignored.  This script copies the
reference's package and its unit tests (sources are never committed: `oracle/_ref/` stays out of history) so that, on
the box,
  * `bench.py --impl reference` can time the UNMODIFIED reference (`cpu_baseline.kind: "reference"`), and
  * `tests/test_reference_suite.py` can run the reference's own unit tests against `pastml_b200`.
Run in the build container: ``python oracle/stage_reference.py``; `__graft_entry__.build()` calls it when
`/root/reference` is present.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, '_ref')
SOURCE = os.environ.get('PASTML_REFERENCE_SOURCE', '/root/reference')


def stage(force=False):
    """Copies <reference>/pastml and <reference>/tests into oracle/_ref/; returns the destination or None."""
    if not os.path.isdir(os.path.join(SOURCE, 'pastml')):
        return DEST if os.path.isdir(os.path.join(DEST, 'pastml')) else None
    marker = os.path.join(DEST, '.staged')
    if os.path.exists(marker) and not force:
        return DEST
    os.makedirs(DEST, exist_ok=True)
    for sub in ('pastml', 'tests'):
        dst = os.path.join(DEST, sub)
        if os.path.isdir(dst):
            shutil.rmtree(dst)
        shutil.copytree(os.path.join(SOURCE, sub), dst, ignore=shutil.ignore_patterns('__pycache__', '*.pyc'))
    with open(marker, 'w') as f:
        f.write('staged from {}\n'.format(SOURCE))
    return DEST


if __name__ == '__main__':
    print(stage(force='--force' in sys.argv))

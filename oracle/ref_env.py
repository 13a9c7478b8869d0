"""TEST INFRASTRUCTURE ONLY -- makes the unmodified reference importable in the build container.

The reference (``/root/reference``, pure Python) needs ``ete3``, ``Bio.Phylo`` and ``itolapi`` at import
time (``/root/reference/pastml/tree.py:7-9``, ``/root/reference/pastml/acr.py:10-12``); none of them is installed
here and there is no network.  Only ``ete3.Tree`` needs behaviour, and :class:`pastml_b200.tree.Tree`
implements the part of it the ACR path uses, so this module registers it under the ``ete3`` name plus
empty stand-ins for the other two and puts the reference on ``sys.path``.

Nothing here is shipped or imported by the product path; it is used by ``oracle/make_golden.py`` (fixture
generation) and by CPU tests that are skipped when ``/root/reference`` is absent (e.g. on the GPU box).
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get('PASTML_REFERENCE_ROOT', '/root/reference')
REPO_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'pastml'))


def install():
    """Registers the stand-in modules and returns the imported reference package."""
    if not reference_available():
        raise RuntimeError('reference sources not found under {}'.format(REFERENCE_ROOT))
    sys.dont_write_bytecode = True
    if REPO_ROOT not in sys.path:
        sys.path.insert(0, REPO_ROOT)
    from pastml_b200.tree import Tree, NewickError

    if 'ete3' not in sys.modules:
        ete3 = types.ModuleType('ete3')
        ete3.Tree = Tree
        ete3.TreeNode = Tree
        parser = types.ModuleType('ete3.parser')
        newick = types.ModuleType('ete3.parser.newick')
        newick.NewickError = NewickError
        parser.newick = newick
        ete3.parser = parser
        sys.modules['ete3'] = ete3
        sys.modules['ete3.parser'] = parser
        sys.modules['ete3.parser.newick'] = newick
    if 'Bio' not in sys.modules:
        bio = types.ModuleType('Bio')
        phylo = types.ModuleType('Bio.Phylo')
        newickio = types.ModuleType('Bio.Phylo.NewickIO')
        import io
        newickio.StringIO = io.StringIO
        phylo.NewickIO = newickio
        phylo.write = lambda *a, **k: None
        phylo.parse = lambda *a, **k: iter(())
        bio.Phylo = phylo
        sys.modules['Bio'] = bio
        sys.modules['Bio.Phylo'] = phylo
        sys.modules['Bio.Phylo.NewickIO'] = newickio
    if 'itolapi' not in sys.modules:
        itolapi = types.ModuleType('itolapi')
        itolapi.Itol = type('Itol', (), {})
        sys.modules['itolapi'] = itolapi
    if REFERENCE_ROOT not in sys.path:
        sys.path.append(REFERENCE_ROOT)
    import pastml
    return pastml

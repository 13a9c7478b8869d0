"""The reference's OWN unit tests (`/root/reference/tests/*Test.py`, staged under oracle/_ref by
oracle/stage_reference.py) run against THIS package: `pastml` and its sub-modules are aliased to `pastml_b200`
(`ete3.Tree` to `pastml_b200.tree.Tree`), the test modules are imported unmodified from a scratch copy of the tests
directory (some of them write files next to their data) and run with unittest.  Everything the reference asserts --
log-likelihoods, scaling factors and frequencies to 3 decimals, marginal probabilities, exact state sets at named
nodes, node counts after collapsing zero branches, parsimony step counts, rate-matrix round trips, HKY/F81/JC
nestings, marginal_counts against brute-force simulation -- must hold on the GPU engine.
"""
import importlib
import os
import shutil
import sys
import types
import unittest

import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)


def _reference_tests_dir():
    for cand in ('/root/reference/tests', os.path.join(REPO, 'oracle', '_ref', 'tests')):
        if os.path.isdir(cand):
            return cand
    return None


def _alias_pastml():
    """`import pastml...` resolves to pastml_b200; `from ete3 import Tree` to our tree class."""
    import pastml_b200
    mapping = {}
    for name in ('', '.acr', '.ml', '.parsimony', '.tree', '.annotation', '.file', '.models', '.models.F81Model',
                 '.models.JCModel', '.models.EFTModel', '.models.HKYModel', '.models.JTTModel',
                 '.models.CustomRatesModel', '.models.generator', '.utilities', '.utilities.state_simulator',
                 '.utilities.transition_counter'):
        mapping['pastml' + name] = importlib.import_module('pastml_b200' + name)
    saved = {k: sys.modules.get(k) for k in list(mapping) + ['ete3']}
    sys.modules.update(mapping)
    ete3 = types.ModuleType('ete3')
    ete3.Tree = ete3.TreeNode = pastml_b200.tree.Tree
    sys.modules['ete3'] = ete3
    return saved


def _restore(saved):
    for k, v in saved.items():
        if v is None:
            sys.modules.pop(k, None)
        else:
            sys.modules[k] = v


MODULES = ['ACRParameterOptimisationMPPAEFTTest', 'ACRParameterOptimisationMPPAF81Test', 'ACRParameterOptimisationMPPAJCTest',
           'ACRStateAcctranTest', 'ACRStateCopyTest', 'ACRStateDeltranTest', 'ACRStateDownpassTest', 'ACRStateJointEFTTest',
           'ACRStateJointF81Test', 'ACRStateJointJCTest', 'ACRStateMAPEFTTest', 'ACRStateMAPF81Test', 'ACRStateMAPJCTest',
           'ACRStateMPPAEFTTest', 'ACRStateMPPAF81Test', 'ACRStateMPPAJCTest', 'CUSTOM_RATESTest', 'DateConversionTest',
           'HKYF81Test', 'HKYJCTest', 'JTTTest', 'MRANDJCTest', 'PijTest']


@pytest.fixture(scope='module')
def reference_tests(tmp_path_factory):
    src = _reference_tests_dir()
    if src is None:
        pytest.skip('no copy of the reference tests (run oracle/stage_reference.py in the build container)')
    dst = tmp_path_factory.mktemp('reference_tests')
    work = os.path.join(str(dst), 'tests')
    shutil.copytree(src, work)
    saved = _alias_pastml()
    sys.path.insert(0, work)
    cwd = os.getcwd()
    os.chdir(work)
    yield work
    os.chdir(cwd)
    sys.path.remove(work)
    for name in MODULES:
        sys.modules.pop(name, None)
    _restore(saved)


@pytest.mark.parametrize('module', MODULES)
def test_reference_unit_tests_pass_on_this_engine(reference_tests, module):
    mod = importlib.import_module(module)
    suite = unittest.defaultTestLoader.loadTestsFromModule(mod)
    assert suite.countTestCases() > 0
    result = unittest.TestResult()
    suite.run(result)
    problems = ['{}: {}'.format(t.id(), tb.strip().splitlines()[-1]) for t, tb in result.failures + result.errors]
    assert not problems, '\n'.join(problems)

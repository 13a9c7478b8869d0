"""TEST INFRASTRUCTURE ONLY -- reference outcomes for the host-side options of the fit that the main fixtures do not
exercise: smoothing (tau in the parameter vector, ``acr.py:185-187, 545``), frequency smoothing
(``models/__init__.py:331-335``), re-optimisation from a parameter file and parameter files of HKY
(``models/__init__.py:191-258``, ``HKYModel.py:132-189``), DENV2 under MAP / JOINT / MPPA x JC / EFT / F81, and the
whole batch of HIV1C drug-resistance columns (K = 2) at once.  Run in the build container:

    python oracle/make_golden_options.py [case ...]     ->  tests/golden/options/<case>.npz

Every case stores what `acr()` of the UNMODIFIED reference returned (log-likelihoods, parameters, state masks) plus
the ensemble of optima the reference reaches when one ulp of noise is injected into its own objective (the
yardstick for "same optimum": L-BFGS-B differentiates by forward differences with h = 1e-8).
"""
import json
import os
import sys
import time

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path = [p for p in sys.path if os.path.abspath(p or '.') != HERE]
sys.path.insert(0, REPO)

from oracle import ref_env  # noqa: E402

ref_env.install()

from pastml.acr import acr  # noqa: E402
import pastml.ml as ref_ml  # noqa: E402
from pastml_b200.tree import read_tree, flatten_forest, name_tree  # noqa: E402
from oracle.make_golden import table_forest, annotate, masks_of  # noqa: E402

OUT_DIR = os.path.join(REPO, 'tests', 'golden', 'options')
REF = ref_env.REFERENCE_ROOT
ALBANIA = (os.path.join(REF, 'tests/data/Albanian.tree.152tax.tre'), os.path.join(REF, 'tests/data/data.txt'))


def _result_fields(res, states):
    mdl = res.get('model')
    out = {}
    for key in ('log_likelihood', 'log_likelihood_restricted_JOINT', 'log_likelihood_restricted_MAP',
                'log_likelihood_restricted_MPPA', 'num_unresolved_nodes', 'num_states_per_node_avg', 'steps'):
        if key in res:
            out[key] = np.float64(res[key])
    if mdl is not None:
        out['sf'] = np.float64(mdl.sf)
        out['tau'] = np.float64(mdl.tau)
        out['pi'] = np.array(getattr(mdl, 'frequencies', np.ones(len(states)) / len(states)), dtype=np.float64)
        out['kappa'] = np.float64(getattr(mdl, 'kappa', 0.))
    return out


def _with_noise(fn, seeds, noise=1e-15):
    """Outcomes of fn() when the reference's objective carries one ulp of multiplicative noise."""
    orig = ref_ml.get_bottom_up_loglikelihood
    outs = []
    for seed in seeds:
        rng = np.random.default_rng(seed)

        def noisy(*a, **k):
            return orig(*a, **k) * (1. + noise * rng.standard_normal())

        ref_ml.get_bottom_up_loglikelihood = noisy
        try:
            outs.append(fn())
        finally:
            ref_ml.get_bottom_up_loglikelihood = orig
    return outs


def single(name, make, column, states, model, ensemble=6, **kwargs):
    """One column, one acr() call with the given options."""
    t0 = time.time()
    states = np.array(sorted(states)) if model != 'HKY' else np.array(list('ACGT'))

    def run():
        forest, node2states = make()
        annotate(forest, node2states, column)
        np.random.seed(0)
        res = acr(forest, columns=[column], column2states={column: states}, model=model, threads=1, **kwargs)
        return forest, res

    forest0, node2states0 = make()   # the inputs, before acr() overwrites the feature with its predictions
    annotate(forest0, node2states0, column)
    annot_mask, has_state = masks_of(flatten_forest(forest0), states, column)
    forest, res = run()
    flat = flatten_forest(forest)
    out = dict(parent=flat.parent, dist=flat.dist, names=np.array(flat.names), annot_mask=annot_mask, has_state=has_state,
               states=states)
    out.update(_result_fields(res[-1], states))
    out['state_mask'] = masks_of(flat, states, column)[0]
    if ensemble:
        ens = _with_noise(lambda: _result_fields(run()[1][-1], states), range(1, ensemble + 1))
        for key in ('log_likelihood', 'sf', 'tau', 'kappa'):
            out['ens_' + key] = np.array([e[key] for e in ens])
        out['ens_pi'] = np.array([e['pi'] for e in ens])
    pfile = (kwargs.get('column2parameters') or {}).get(column)
    if pfile:
        out['parameter_file_text'] = np.array(open(pfile).read())
    meta = dict(name=name, column=column, model=model, options={k: v for k, v in kwargs.items() if not isinstance(v, dict)},
                reference_seconds=round(time.time() - t0, 1))
    out['meta'] = np.array(json.dumps(meta, default=str))
    os.makedirs(OUT_DIR, exist_ok=True)
    np.savez_compressed(os.path.join(OUT_DIR, name + '.npz'), **out)
    print('{}: lh={:.10f} sf={:.9g} tau={:.6g} ({:.0f}s)'.format(name, float(out['log_likelihood']), float(out['sf']),
                                                                  float(out['tau']), time.time() - t0))


def case_smoothing():
    make, states = table_forest(*ALBANIA, 'Country')
    single('albania_f81_smoothing', make, 'Country', states, 'F81', prediction_method='MPPA', tau=None)


def case_frequency_smoothing():
    make, states = table_forest(*ALBANIA, 'Country')
    single('albania_f81_frequency_smoothing', make, 'Country', states, 'F81', prediction_method='MPPA', frequency_smoothing=True)


def case_reoptimise():
    make, states = table_forest(*ALBANIA, 'Country')
    pfile = os.path.join(REF, 'tests/data/params.character_Country.method_MPPA.model_JC.tab')
    single('albania_jc_reoptimise', make, 'Country', states, 'JC', prediction_method='MPPA',
           column2parameters={'Country': pfile}, reoptimise=True)
    single('albania_jc_parameter_file', make, 'Country', states, 'JC', ensemble=0, prediction_method='MPPA',
           column2parameters={'Country': pfile})


def case_hky_parameter_file():
    base = os.path.join(REF, 'tests/data/tree.152taxa.sf_0.5.A_0.6.C_0.15.G_0.2.T_0.05')
    df = pd.read_csv(base + '.pastml.tab', sep='\t', index_col=0)
    tips = {t.name for t in read_tree(base + '.nwk')}
    node2states = {str(k): {v} for k, v in df['Real'].items() if str(k) in tips and not pd.isna(v)}
    make = lambda: ([read_tree(base + '.nwk')], node2states)   # noqa: E731
    pfile = os.path.join(OUT_DIR, 'hky_parameters.tab')
    os.makedirs(OUT_DIR, exist_ok=True)
    with open(pfile, 'w') as f:   # the format Model.save_parameters writes (parameter <tab> value)
        f.write('parameter\tvalue\nscaling_factor\t0.62\nkappa\t3.1\nA\t0.55\nC\t0.17\nG\t0.2\nT\t0.08\n')
    single('sim152_hky_parameter_file', make, 'nt', list('ACGT'), 'HKY', ensemble=0, prediction_method='MPPA',
           column2parameters={'nt': pfile})
    single('sim152_hky_reoptimise', make, 'nt', list('ACGT'), 'HKY', prediction_method='MPPA',
           column2parameters={'nt': pfile}, reoptimise=True)


def case_denv2_methods():
    tree = os.path.join(REF, 'examples/DENV2/data/real/raxml_tree.dated.nwk')
    table = os.path.join(REF, 'examples/DENV2/data/real/metadata.tab')
    for column in ('Location', 'genotype'):
        make, states = table_forest(tree, table, column, sep='\t', index_col=0, do_name=True)
        for model in ('JC', 'EFT'):
            for method in ('MAP', 'JOINT', 'MPPA'):
                single('denv2_{}_{}_{}'.format(column.lower(), model.lower(), method.lower()), make, column, states, model,
                       ensemble=4 if method == 'MPPA' else 0, prediction_method=method)


def case_hiv1c_drms(max_columns=None):
    """All the drug-resistance columns of the HIV1C example (binary characters) in ONE acr() call, JC."""
    t0 = time.time()
    tree = os.path.join(REF, 'examples/HIV1C/data/best/pastml_phyml_tree.nwk')
    table = os.path.join(REF, 'examples/HIV1C/data/metadata.tab')
    df = pd.read_csv(table, sep='\t', index_col=0, header=0)
    df.index = df.index.map(str)
    columns = [c for c in df.columns if (c.startswith('RT:') or c.startswith('PR:')) and df[c].nunique(dropna=True) == 2]
    if max_columns:
        columns = columns[:max_columns]
    root = read_tree(open(tree).read().replace('\n', ''))
    for n in root.traverse():
        if n.dist < 0:
            n.dist = 0
    tips = {t.name for t in root}
    sub = df.loc[[i for i in df.index if i in tips], columns].astype(object)
    sub = sub.where(sub.notna(), None)
    sub.columns = [c.replace(':', '_') for c in columns]
    np.random.seed(0)
    res = acr(root, df=sub.map(lambda v: None if v is None else str(v)), prediction_method='MPPA', model='JC', threads=1)
    flat = flatten_forest([root])
    out = dict(parent=flat.parent, dist=flat.dist, names=np.array(flat.names), columns=np.array(list(sub.columns)))
    lhs, sfs, masks, annot, has, states_all = [], [], [], [], [], []
    for r in res:
        col = r['character']
        states = np.array(r['states'])
        m, h = None, None
        lhs.append(r['log_likelihood'])
        sfs.append(r['model'].sf)
        states_all.append('|'.join(states))
        masks.append(masks_of(flat, states, col)[0])
    # tip annotations as masks (the feature was overwritten by the prediction on the tips: rebuild from the table)
    for col in sub.columns:
        states = sorted({str(v) for v in sub[col] if v is not None})
        s2i = {s: i for i, s in enumerate(states)}
        a = np.zeros(flat.n_nodes, dtype=np.uint64)
        hs = np.zeros(flat.n_nodes, dtype=np.uint8)
        lookup = sub[col].to_dict()
        for i, name in enumerate(flat.names):
            if name in lookup:
                hs[i] = 1
                if lookup[name] is not None:
                    a[i] = np.uint64(1 << s2i[str(lookup[name])])
        annot.append(a)
        has.append(hs)
    out.update(log_likelihood=np.array(lhs), sf=np.array(sfs), mppa_mask=np.array(masks).astype(np.uint8),
               annot_mask=np.array(annot).astype(np.uint8), has_state=np.array(has), states=np.array(states_all))
    out['meta'] = np.array(json.dumps(dict(name='hiv1c_drms_jc', n_columns=len(columns), reference_seconds=round(time.time() - t0, 1))))
    np.savez_compressed(os.path.join(OUT_DIR, 'hiv1c_drms_jc.npz'), **out)
    print('hiv1c_drms_jc: {} columns ({:.0f}s)'.format(len(columns), time.time() - t0))


CASES = {'smoothing': case_smoothing, 'frequency_smoothing': case_frequency_smoothing, 'reoptimise': case_reoptimise,
         'hky_parameter_file': case_hky_parameter_file, 'denv2_methods': case_denv2_methods, 'hiv1c_drms': case_hiv1c_drms}

if __name__ == '__main__':
    import warnings
    warnings.filterwarnings('ignore')
    for name in (sys.argv[1:] or list(CASES)):
        CASES[name]()

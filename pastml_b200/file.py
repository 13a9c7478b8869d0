"""Names of the files ``pastml_pipeline`` writes into its work directory.

The names are part of pastml's output contract (``/root/reference/pastml/file.py:7-103``: users and downstream
scripts look for ``params.character_<c>.method_<m>.model_<model>.tab`` etc.), so they are reproduced exactly; they are
kept in one table here instead of a function per file.
"""
import os

from pastml_b200 import col_name2cat, get_personalized_feature_name

_TEMPLATES = {
    'work_dir': '{tree}_pastml',
    'combined_states': 'combined_ancestral_states.tab',
    'named_tree': 'named.tree_{tree}.nwk',
    'ml_params': 'params.character_{state}.method_{method}.model_{model}.tab',
    'mp_params': 'params.character_{state}.method_{method}.tab',
    'marginal_probs': 'marginal_probabilities.character_{state}.model_{model}.tab',
}


def _resolved(column, method):
    """Meta-methods (ML, MP, ALL) write their files under the default method they stand for, with the column renamed
    to ``<column>_<method>`` (the feature the reconstruction is stored under)."""
    from pastml_b200.ml import is_meta_ml, get_default_ml_method
    from pastml_b200.parsimony import is_meta_mp, get_default_mp_method
    column = col_name2cat(column)
    default = get_default_ml_method() if is_meta_ml(method) else get_default_mp_method() if is_meta_mp(method) else None
    if default is None:
        return column, method
    return get_personalized_feature_name(column, default), default


def get_pastml_parameter_file(method, model, column):
    from pastml_b200.ml import is_ml
    column, method = _resolved(column, method)
    return _TEMPLATES['ml_params' if is_ml(method) else 'mp_params'].format(state=column, method=method, model=model)


def get_pastml_marginal_prob_file(method, model, column):
    from pastml_b200.ml import is_marginal
    if not is_marginal(method):
        return None
    column, _ = _resolved(column, method)
    return _TEMPLATES['marginal_probs'].format(state=column, model=model)


def get_combined_ancestral_state_file():
    return _TEMPLATES['combined_states']


def get_named_tree_file(tree):
    return _TEMPLATES['named_tree'].format(tree=os.path.splitext(os.path.basename(tree))[0])


def get_pastml_work_dir(tree):
    return _TEMPLATES['work_dir'].format(tree=os.path.splitext(tree)[0])

"""pastml_b200 -- a B200-native engine for PastML's ancestral character reconstruction hot path.

The package mirrors the slice of pastml's Python API that fronts the numerical path
(``acr`` / ``ml_acr`` / ``parsimonious_acr`` / the model classes; result-dictionary keys below are pastml's,
``/root/reference/pastml/__init__.py:6-15``) and runs it on hand-written sm_100a CUDA kernels behind a C ABI
(``include/pastml_b200.h``).  There is no CPU fallback.
"""

PASTML_B200_VERSION = '0.1.0'
# the pastml release whose behaviour is reproduced
PASTML_VERSION = '1.9.50'

METHOD = 'method'
STATES = 'states'
CHARACTER = 'character'

NUM_SCENARIOS = 'num_scenarios'
NUM_UNRESOLVED_NODES = 'num_unresolved_nodes'
NUM_STATES_PER_NODE = 'num_states_per_node_avg'
PERC_UNRESOLVED = 'percentage_of_unresolved_nodes'
NUM_NODES = 'num_nodes'
NUM_TIPS = 'num_tips'


def get_personalized_feature_name(character, feature):
    """``<character>_<feature>``: pastml's per-character namespace for node features."""
    return '{}_{}'.format(character, feature)


def col_name2cat(column):
    """Keeps letters, digits and underscores of a column name (spaces become underscores)."""
    return ''.join(ch for ch in column.replace(' ', '_') if ch.isalnum() or ch == '_')


def value2list(n, value, default_value):
    """Broadcasts a per-column setting to n columns (``/root/reference/pastml/__init__.py:81-94``)."""
    if value is None:
        value = default_value
    if not isinstance(value, list):
        return [value] * n
    if len(value) == 1:
        return value * n
    return value + [default_value] * (n - len(value))


def datetime2numeric(d):
    """Calendar date -> decimal year (the day's start as a fraction of its year, leap years included)."""
    from datetime import datetime
    start, end = datetime(d.year, 1, 1), datetime(d.year + 1, 1, 1)
    return d.year + (datetime(d.year, d.month, d.day) - start).days / (end - start).days


def numeric2datetime(x):
    """Decimal year -> calendar date (inverse of :func:`datetime2numeric`, rounding to the nearest day)."""
    from datetime import datetime, timedelta
    year = int(x)
    start, end = datetime(year, 1, 1), datetime(year + 1, 1, 1)
    return start + timedelta(days=round((x - year) * (end - start).days))


def _set_up_pastml_logger(verbose):
    import logging
    logger = logging.getLogger('pastml')
    logger.setLevel(level=logging.DEBUG if verbose else logging.ERROR)
    logger.propagate = False
    if not logger.hasHandlers():
        handler = logging.StreamHandler()
        handler.setFormatter(logging.Formatter('%(name)s:%(levelname)s:%(asctime)s %(message)s', datefmt='%H:%M:%S'))
        logger.addHandler(handler)
    return logger

"""Import-path compatibility with ``pastml.models.HKYModel``: everything lives in :mod:`pastml_b200.models`."""
from pastml_b200.models import HKYModel, HKY, HKY_STATES, KAPPA  # noqa: F401

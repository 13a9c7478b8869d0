"""Import-path compatibility with ``pastml.models.HKYModel``: everything lives in :mod:`pastml_b200.models`."""
from pastml_b200.models import HKYModel, HKY, HKY_STATES, KAPPA  # noqa: F401

# indices of the nucleotides in HKY_STATES (``/root/reference/pastml/models/HKYModel.py:9``)
A, C, G, T = range(4)

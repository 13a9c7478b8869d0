"""Import-path compatibility with ``pastml.models.JTTModel``: everything lives in :mod:`pastml_b200.models`."""
from pastml_b200.models import JTTModel, JTT, JTT_STATES, JTT_FREQUENCIES, JTT_RATE_MATRIX  # noqa: F401

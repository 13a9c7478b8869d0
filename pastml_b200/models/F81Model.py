"""Import-path compatibility with ``pastml.models.F81Model``: everything lives in :mod:`pastml_b200.models`."""
from pastml_b200.models import F81Model, F81  # noqa: F401

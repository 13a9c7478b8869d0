"""Import-path compatibility with ``pastml.models.JCModel``: everything lives in :mod:`pastml_b200.models`."""
from pastml_b200.models import JCModel, JC  # noqa: F401

"""Import-path compatibility with ``pastml.models.EFTModel``: everything lives in :mod:`pastml_b200.models`."""
from pastml_b200.models import EFTModel, EFT  # noqa: F401

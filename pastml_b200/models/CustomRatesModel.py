"""Import-path compatibility with ``pastml.models.CustomRatesModel``: everything lives in :mod:`pastml_b200.models`."""
from pastml_b200.models import CustomRatesModel, CUSTOM_RATES, load_custom_rates  # noqa: F401

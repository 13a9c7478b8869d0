"""Import-path compatibility with ``pastml.models.generator``: everything lives in :mod:`pastml_b200.models`."""
from pastml_b200.models import get_diagonalisation, get_normalised_generator, get_pij_matrix, save_matrix  # noqa: F401

"""CUSTOM_RATES: a user-supplied symmetric rate matrix (``pastml/models/CustomRatesModel.py``)."""
import numpy as np
import pandas as pd  # noqa: F401

from pastml_b200.models import Model, ModelWithFrequencies, CUSTOM_RATES, ENGINE_EIG, _log  # noqa: F401
from pastml_b200.models.generator import get_diagonalisation, get_pij_matrix, get_normalised_generator  # noqa: F401


def load_custom_rates(infile):
    """Reads a symmetric rate matrix (first line: '# ' + state names) and returns (sorted states, matrix)."""
    rate_matrix = np.loadtxt(infile, dtype=np.float64, comments='#', delimiter=' ')
    if rate_matrix.ndim != 2 or rate_matrix.shape[0] != rate_matrix.shape[1]:
        raise ValueError('The input rate matrix must be squared, but yours is {}.'
                         .format('x'.join(str(_) for _ in rate_matrix.shape)))
    if not np.all(rate_matrix == rate_matrix.transpose()):
        raise ValueError('The input rate matrix must be symmetric, but yours is not.')
    np.fill_diagonal(rate_matrix, 0)
    n = len(rate_matrix)
    if np.count_nonzero(rate_matrix) != n * (n - 1):
        _log.warning('The rate matrix contains zero rates (apart from the diagonal).')
    with open(infile, 'r') as f:
        header = f.readline()
    if not header.startswith('#'):
        raise ValueError('The rate matrix file should start with state names, separated by whitespaces and '
                         'preceded by # .')
    states = np.array(header.strip('#').strip('\n').strip().split(' '), dtype=str)
    if len(states) != n:
        raise ValueError('The number of specified state names ({}) does not correspond to the rate matrix '
                         'dimensions ({}x{}).'.format(len(states), *rate_matrix.shape))
    order = np.argsort(states)
    return states[order], np.array(rate_matrix)[:, order][order, :]


class CustomRatesModel(ModelWithFrequencies):
    """General time-reversible model from a user rate matrix; P(t) through the eigen-decomposition."""

    engine_kind = ENGINE_EIG

    def __init__(self, forest_stats, sf=None, frequencies=None, rate_matrix_file=None, states=None, rate_matrix=None,
                 tau=0, optimise_tau=False, frequency_smoothing=False, parameter_file=None, reoptimise=False,
                 **kwargs):
        ModelWithFrequencies.__init__(self, states=states, forest_stats=forest_stats, sf=sf, tau=tau,
                                      frequencies=frequencies, optimise_tau=optimise_tau,
                                      frequency_smoothing=frequency_smoothing, reoptimise=reoptimise,
                                      parameter_file=parameter_file, **kwargs)
        self.name = CUSTOM_RATES
        if rate_matrix_file is None and (rate_matrix is None or states is None):
            raise ValueError('Either the rate matrix file or the rate matrix plus the states must be specified for '
                             '{} model'.format(CUSTOM_RATES))
        if rate_matrix_file is None:
            self._rate_matrix = rate_matrix
        else:
            self._states, self._rate_matrix = load_custom_rates(rate_matrix_file)
        self.D_DIAGONAL, self.A, self.A_INV = get_diagonalisation(self.frequencies, self._rate_matrix)

    @property
    def rate_matrix(self):
        return self._rate_matrix

    @rate_matrix.setter
    def rate_matrix(self, rate_matrix):
        raise NotImplementedError('The rate matrix is preset and cannot be changed.')

    @property
    def frequencies(self):
        return self._frequencies

    @frequencies.setter
    def frequencies(self, frequencies):
        self._set_frequencies(frequencies)
        self.D_DIAGONAL, self.A, self.A_INV = get_diagonalisation(frequencies, self.rate_matrix)

    def get_Pij_t(self, t, *args, **kwargs):
        return get_pij_matrix(self.transform_t(t), self.D_DIAGONAL, self.A, self.A_INV)

    def _engine_fields(self):
        return dict(pi=self.frequencies, d=np.real(self.D_DIAGONAL), A=np.real(self.A), Ainv=np.real(self.A_INV))

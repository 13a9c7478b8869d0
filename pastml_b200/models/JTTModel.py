"""JTT amino-acid model (``pastml/models/JTTModel.py``)."""
import os

import numpy as np

from pastml_b200.models import Model, ModelWithFrequencies, JTT, _log  # noqa: F401
from pastml_b200.models.CustomRatesModel import CustomRatesModel


def _load_jtt():
    data = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'data', 'jtt.npz'))
    return np.array([str(_) for _ in data['states']]), data['frequencies'], data['rate_matrix']


JTT_STATES, JTT_FREQUENCIES, JTT_RATE_MATRIX = _load_jtt()


class JTTModel(CustomRatesModel):
    """Jones-Taylor-Thornton amino-acid model: fixed alphabet, rates and frequencies."""

    def __init__(self, forest_stats, sf=None, tau=0, optimise_tau=False, parameter_file=None, reoptimise=False,
                 **kwargs):
        kwargs['states'] = JTT_STATES
        kwargs.pop('frequency_smoothing', None)
        CustomRatesModel.__init__(self, forest_stats=forest_stats, sf=sf, frequencies=JTT_FREQUENCIES,
                                  rate_matrix=JTT_RATE_MATRIX, parameter_file=parameter_file, reoptimise=reoptimise,
                                  frequency_smoothing=False, tau=tau, optimise_tau=optimise_tau, **kwargs)
        self._optimise_frequencies = False
        self.name = JTT

    @property
    def states(self):
        return self._states

    @states.setter
    def states(self, states):
        raise NotImplementedError('The JTT states are preset and cannot be changed.')

    def parse_parameters(self, params, reoptimise=False):
        return Model.parse_parameters(self, params, reoptimise)

"""Jukes-Cantor: F81 with equal, fixed frequencies (``pastml/models/JCModel.py``)."""
import numpy as np

from pastml_b200.models import Model, ModelWithFrequencies, JC, _log  # noqa: F401
from pastml_b200.models.F81Model import F81Model


class JCModel(F81Model):
    """F81 with all-equal, fixed frequencies."""

    def __init__(self, states, forest_stats, sf=None, tau=0, optimise_tau=False, parameter_file=None,
                 reoptimise=False, **kwargs):
        kwargs['frequency_smoothing'] = False
        F81Model.__init__(self, states=states, forest_stats=forest_stats, sf=sf, tau=tau, optimise_tau=optimise_tau,
                          frequencies=np.ones(len(states), dtype=np.float64) / len(states), reoptimise=reoptimise,
                          parameter_file=parameter_file, **kwargs)
        self._optimise_frequencies = False
        self.name = JC

    def parse_parameters(self, params, reoptimise=False):
        return Model.parse_parameters(self, params, reoptimise)

    def _print_parameters(self):
        return '{}\tfrequencies\tall equal to {:g}\t(fixed)\n'.format(Model._print_parameters(self),
                                                                      1 / len(self.states))

"""EFT: F81 with the frequencies fixed to the observed tip frequencies (``pastml/models/EFTModel.py``)."""
import numpy as np  # noqa: F401

from pastml_b200.models import Model, ModelWithFrequencies, EFT, _log  # noqa: F401
from pastml_b200.models.F81Model import F81Model


class EFTModel(F81Model):
    """F81 with frequencies fixed to the ones observed at the tips."""

    def __init__(self, states, forest_stats, observed_frequencies, sf=None, tau=0, optimise_tau=False,
                 parameter_file=None, reoptimise=False, **kwargs):
        F81Model.__init__(self, states=states, forest_stats=forest_stats, sf=sf, tau=tau, optimise_tau=optimise_tau,
                          frequencies=observed_frequencies, reoptimise=reoptimise, parameter_file=parameter_file,
                          **kwargs)
        self._optimise_frequencies = False
        self._frequency_smoothing = False
        self.name = EFT

    def parse_parameters(self, params, reoptimise=False):
        return Model.parse_parameters(self, params, reoptimise)

    def _print_parameters(self):
        return '{}\tfrequencies:\tobserved in the tree\t(fixed)\n'.format(Model._print_parameters(self))

"""F81: P_ij = (1 - e) pi_j + delta_ij e with e = exp(-mu t'), mu = 1 / (1 - sum pi^2) (``pastml/models/F81Model.py``)."""
import numpy as np

from pastml_b200.models import ModelWithFrequencies, F81, ENGINE_F81, _log  # noqa: F401


class F81Model(ModelWithFrequencies):
    """P_ij(t) = pi_j (1 - e^{-mu t}) + [i == j] e^{-mu t},  mu = 1 / (1 - sum pi^2)."""

    engine_kind = ENGINE_F81

    def __init__(self, states, forest_stats, sf=None, frequencies=None, tau=0, frequency_smoothing=False,
                 optimise_tau=False, parameter_file=None, reoptimise=False, **kwargs):
        ModelWithFrequencies.__init__(self, states=states, forest_stats=forest_stats, sf=sf, tau=tau,
                                      frequencies=frequencies, optimise_tau=optimise_tau,
                                      frequency_smoothing=frequency_smoothing, reoptimise=reoptimise,
                                      parameter_file=parameter_file, **kwargs)
        self.name = F81

    def get_mu(self):
        return 1. / (1. - self.frequencies.dot(self.frequencies))

    def get_Pij_t(self, t, *args, **kwargs):
        t = self.transform_t(t)
        with np.errstate(divide='ignore'):
            mu = self.get_mu()
        decay = 0. if mu == np.inf else np.exp(-mu * t)
        return (1 - decay) * self.frequencies + np.eye(len(self.frequencies)) * decay

    def _engine_fields(self):
        with np.errstate(divide='ignore'):
            return dict(pi=self.frequencies, mu=self.get_mu())

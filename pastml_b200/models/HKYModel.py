"""HKY85 (``pastml/models/HKYModel.py``)."""
import numpy as np

from pastml_b200.models import (Model, ModelWithFrequencies, HKY, KAPPA, ENGINE_HKY, _log, _read_parameters,  # noqa: F401
                                _positive_float)

# indices of the nucleotides in HKY_STATES (``/root/reference/pastml/models/HKYModel.py:9``)
A, C, G, T = range(4)

HKY_STATES = np.array(['A', 'C', 'G', 'T'])


class HKYModel(ModelWithFrequencies):
    """HKY85 on the fixed alphabet A, C, G, T with a transition/transversion ratio kappa."""

    engine_kind = ENGINE_HKY

    def __init__(self, forest_stats, sf=None, frequencies=None, kappa=4, tau=0, frequency_smoothing=False,
                 optimise_tau=False, parameter_file=None, reoptimise=False, **kwargs):
        self._kappa = None
        self._optimise_kappa = True
        kwargs['states'] = HKY_STATES
        ModelWithFrequencies.__init__(self, forest_stats=forest_stats, sf=sf, tau=tau, optimise_tau=optimise_tau,
                                      frequencies=frequencies, frequency_smoothing=frequency_smoothing,
                                      reoptimise=reoptimise, parameter_file=parameter_file, **kwargs)
        if self._kappa is None:
            self._kappa = kappa
        self.name = HKY

    @property
    def kappa(self):
        return self._kappa

    @kappa.setter
    def kappa(self, kappa):
        if not self._optimise_kappa:
            raise NotImplementedError('The kappa value is preset and cannot be changed.')
        self._kappa = kappa

    @property
    def states(self):
        return self._states

    @states.setter
    def states(self, states):
        raise NotImplementedError('The HKY model is only implemented for nucleotides: '
                                  'the states are A, C, G, T and cannot be reset')

    def get_Pij_t(self, t, *args, **kwargs):
        """Closed-form HKY85 transition matrix (rows/columns in A, C, G, T order)."""
        t = self.transform_t(t)
        a, c, g, tt = self.frequencies
        pur, pyr = a + g, c + tt
        k = self.kappa
        beta = .5 / (pur * pyr + k * (a * g + c * tt))
        e_all = np.exp(-beta * t)
        e_pyr = np.exp(-beta * t * (1. + pyr * (k - 1.))) / pyr
        e_pur = np.exp(-beta * t * (1. + pur * (k - 1.))) / pur
        s_pyr = (pyr + pur * e_all) / pyr
        s_pur = (pur + pyr * e_all) / pur
        p = np.ones((4, 4), dtype=np.float64) * (1 - e_all)
        p *= self.frequencies
        A_, C_, G_, T_ = 0, 1, 2, 3
        p[T_, T_], p[T_, C_] = tt * s_pyr + c * e_pyr, c * s_pyr - c * e_pyr
        p[C_, T_], p[C_, C_] = tt * s_pyr - tt * e_pyr, c * s_pyr + tt * e_pyr
        p[A_, A_], p[A_, G_] = a * s_pur + g * e_pur, g * s_pur - g * e_pur
        p[G_, A_], p[G_, G_] = a * s_pur - a * e_pur, g * s_pur + a * e_pur
        return p

    def get_num_params(self):
        return ModelWithFrequencies.get_num_params(self) + (1 if self._optimise_kappa else 0)

    def set_params_from_optimised(self, ps, **kwargs):
        if self.extra_params_fixed():
            Model.set_params_from_optimised(self, ps, **kwargs)
            return
        ModelWithFrequencies.set_params_from_optimised(self, ps, **kwargs)
        if self._optimise_kappa:
            self.kappa = ps[ModelWithFrequencies.get_num_params(self)]

    def get_optimised_parameters(self):
        if self.extra_params_fixed():
            return Model.get_optimised_parameters(self)
        return np.hstack((ModelWithFrequencies.get_optimised_parameters(self),
                          [self.kappa] if self._optimise_kappa else []))

    def get_bounds(self):
        if self.extra_params_fixed():
            return Model.get_bounds(self)
        return np.array((*ModelWithFrequencies.get_bounds(self),
                         *([np.array([1e-6, 20.])] if self._optimise_kappa else [])))

    def parse_parameters(self, params, reoptimise=False):
        params = ModelWithFrequencies.parse_parameters(self, params, reoptimise)
        if KAPPA in params:
            self._kappa = _positive_float(params[KAPPA], 'Kappa')
            if self._kappa is not None:
                self._optimise_kappa = reoptimise
        return params

    def freeze(self):
        ModelWithFrequencies.freeze(self)
        self._optimise_kappa = False

    def save_parameters(self, filehandle):
        ModelWithFrequencies.save_parameters(self, filehandle)
        filehandle.write('{}\t{:g}\n'.format(KAPPA, self.kappa))

    def _print_parameters(self):
        return '{}\tkappa\t{:.6f}\t{}\n'.format(ModelWithFrequencies._print_parameters(self), self.kappa,
                                               '(optimised)' if self._optimise_kappa else '(fixed)')

    def _engine_fields(self):
        return dict(pi=self.frequencies, kappa=float(self.kappa))


# ---------------------------------------------------------------------- general reversible models

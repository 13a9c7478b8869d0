"""Character-evolution models: parameter objects, optimiser-vector packing and ``P(t)``.

Host-side mirror of ``pastml.models`` (``/root/reference/pastml/models/__init__.py`` and the per-model modules): same
class names, constructor arguments, attributes and parameter-vector conventions, so that objects returned in ACR
results (``result['model']``) look like pastml's.  ``get_Pij_t`` is kept for callers and tests; the likelihood
sweeps never call it -- they receive one packed parameter block per (parameter set, column) lane
(:meth:`Model.engine_block`, layout in ``include/pastml_b200.h``) and evaluate P(t) on the device.

Optimiser vector (``get_optimised_parameters`` / ``set_params_from_optimised`` / ``get_bounds``), in order:
``[sf] [tau] [pi_0/pi_{K-1} .. pi_{K-2}/pi_{K-1} | frequency-smoothing s] [kappa]``, each present only when free.
"""
import logging
import os

import numpy as np
import pandas as pd

from pastml_b200 import NUM_NODES, NUM_TIPS

MODEL = 'model'
CHANGES_PER_AVG_BRANCH = 'state_changes_per_avg_branch'
SCALING_FACTOR = 'scaling_factor'
SMOOTHING_FACTOR = 'smoothing_factor'
FREQUENCIES = 'frequencies'
KAPPA = 'kappa'

F81, JC, EFT, HKY, JTT, CUSTOM_RATES = 'F81', 'JC', 'EFT', 'HKY', 'JTT', 'CUSTOM_RATES'

# model families of the engine (include/pastml_b200.h)
ENGINE_F81, ENGINE_HKY, ENGINE_EIG = 0, 1, 2

_log = logging.getLogger('pastml')


def _read_parameters(params):
    """dict, or path of a two-column tab-delimited table (name, 'value') -> {str: value}."""
    if not isinstance(params, (str, dict)):
        raise ValueError('Parameters must be specified either as a dict or as a path to a csv file, not as {}!'
                         .format(type(params)))
    if isinstance(params, str):
        if not os.path.exists(params):
            raise ValueError('The specified parameter file ({}) does not exist.'.format(params))
        try:
            table = pd.read_csv(params, header=0, index_col=0, sep='\t')
            params = table.to_dict()['value']
        except Exception:
            raise ValueError('The specified parameter file {} is malformed, should be a tab-delimited file with two '
                             'columns, the first one containing parameter names, and the second, named "value", '
                             'containing parameter values.'.format(params))
    return {str(k.encode('ASCII', 'replace').decode()): v for k, v in params.items()}


def _positive_float(value, what, allow_zero=False):
    try:
        value = np.float64(value)
    except Exception:
        _log.error('{} ({}) given in parameters is not float, ignoring it.'.format(what, value))
        return None
    if value < 0 or (value == 0 and not allow_zero):
        _log.error('{} cannot be negative, ignoring the value given in parameters ({}).'.format(what, value))
        return None
    return value


class Model(object):
    """Scaling factor + smoothing factor: what every model has (``transform_t``)."""

    engine_kind = None

    def __init__(self, states, forest_stats, sf=None, tau=0, optimise_tau=False, parameter_file=None,
                 reoptimise=False, character=None, **kwargs):
        self._name = None
        self._states = np.sort(states)
        self._forest_stats = forest_stats
        self._optimise_tau = optimise_tau
        self._optimise_sf = True
        self._sf = None
        self._tau = None
        self._character = character
        self.parse_parameters(parameter_file, reoptimise)
        if self._sf is None:
            self._sf = sf if sf is not None else 1. / forest_stats.avg_nonzero_brlen
        if self._tau is None:
            self._tau = tau if tau else 0
        self.calc_tau_factor()
        self._extra_params_fixed = False

    # ---------------------------------------------------------------- plain properties
    name = property(lambda self: self._name, lambda self, v: setattr(self, '_name', v))
    states = property(lambda self: self._states, lambda self, v: setattr(self, '_states', v))

    @property
    def forest_stats(self):
        return self._forest_stats

    @forest_stats.setter
    def forest_stats(self, forest_stats):
        self._forest_stats = forest_stats
        self.calc_tau_factor()

    @property
    def sf(self):
        return self._sf

    @sf.setter
    def sf(self, sf):
        if not self._optimise_sf:
            raise NotImplementedError('The scaling factor is preset and cannot be changed.')
        self._sf = sf

    @property
    def tau(self):
        return self._tau

    @tau.setter
    def tau(self, tau):
        if not self._optimise_tau:
            raise NotImplementedError('Tau is preset and cannot be changed.')
        self._tau = tau
        self.calc_tau_factor()

    def calc_tau_factor(self):
        """Branches get +tau and are renormalised to keep the forest length (``models/__init__.py:39-42``)."""
        fs = self._forest_stats
        self._tau_factor = fs.forest_length / (fs.forest_length + self._tau * (fs.num_nodes - 1)) if self._tau else 1

    def transform_t(self, t):
        return (t + self.tau) * self._tau_factor * self.sf

    def get_Pij_t(self, t, *args, **kwargs):
        raise NotImplementedError('Please implement this method in the Model subclass')

    # ---------------------------------------------------------------- optimiser vector
    def _basic_count(self):
        return int(bool(self._optimise_sf)) + int(bool(self._optimise_tau))

    def get_num_params(self):
        return self._basic_count()

    def get_optimised_parameters(self):
        return np.hstack(([self.sf] if self._optimise_sf else [], [self.tau] if self._optimise_tau else []))

    def set_params_from_optimised(self, ps, **kwargs):
        if self._optimise_sf:
            self.sf = ps[0]
        if self._optimise_tau:
            self.tau = ps[1 if self._optimise_sf else 0]

    def _basic_bounds(self):
        avg = self.forest_stats.avg_nonzero_brlen
        bounds = []
        if self._optimise_sf:
            bounds.append(np.array([0.001 / avg, 10. / avg]))
        if self._optimise_tau:
            bounds.append(np.array([0, avg]))
        return bounds

    def get_bounds(self):
        return np.array(self._basic_bounds(), np.float64)

    def basic_params_fixed(self):
        return not self._optimise_tau and not self._optimise_sf

    def extra_params_fixed(self):
        return self._extra_params_fixed

    def fix_extra_params(self):
        self._extra_params_fixed = True

    def unfix_extra_params(self):
        self._extra_params_fixed = False

    def freeze(self):
        self._optimise_sf = False
        self._optimise_tau = False

    # ---------------------------------------------------------------- parameter input / output
    def parse_parameters(self, params, reoptimise=False):
        """scaling_factor / smoothing_factor entries of a parameter dict or file; given values are kept fixed unless
        ``reoptimise`` (``models/__init__.py:191-258``)."""
        if params is None:
            return {}
        params = _read_parameters(params)
        if SCALING_FACTOR in params:
            self._sf = _positive_float(params[SCALING_FACTOR], 'Scaling factor')
            if self._sf is not None:
                self._optimise_sf = reoptimise
        if SMOOTHING_FACTOR in params:
            self._tau = _positive_float(params[SMOOTHING_FACTOR], 'Smoothing factor', allow_zero=True)
        return params

    def save_parameters(self, filehandle):
        fs = self.forest_stats
        for key, value in ((MODEL, self.name), (NUM_NODES, fs.num_nodes), (NUM_TIPS, fs.num_tips),
                           (SCALING_FACTOR, self.sf), (CHANGES_PER_AVG_BRANCH, self.sf * fs.avg_nonzero_brlen),
                           (SMOOTHING_FACTOR, self.tau)):
            filehandle.write('{}\t{}\n'.format(key, value))

    def _print_basic_parameters(self):
        return '\tscaling factor:\t{:.6f}, i.e. {:.6f} changes per avg branch\t{}\n\tsmoothing factor:\t{:.6f}\t{}\n' \
            .format(self.sf, self.forest_stats.avg_nonzero_brlen * self.sf,
                    '(optimised)' if self._optimise_sf else '(fixed)', self.tau,
                    '(optimised)' if self._optimise_tau else '(fixed)')

    def _print_parameters(self):
        return self._print_basic_parameters()

    def __str__(self):
        return 'Model {} for character {} with parameter values:\n{}'.format(self.name, self._character,
                                                                            self._print_parameters())

    # ---------------------------------------------------------------- engine
    def _engine_fields(self):
        raise NotImplementedError

    def engine_block(self):
        """This model's parameter block for one engine lane."""
        from pastml_b200.engine import pack_params
        return pack_params(self.engine_kind, len(self.states), float(self.sf), float(self.tau),
                           float(self._tau_factor), **self._engine_fields())

    def engine_blocks(self, points):
        """Parameter blocks for several optimiser vectors (the perturbed points of one finite-difference gradient): what
        ``set_params_from_optimised(ps); engine_block()`` gives for each of them in turn, the model being left at the last
        point.  Subclasses may build all the blocks at once (:class:`F81Model`): with one optimiser thread per column
        under the GIL, this per-point Python work was 70 % of the fit's host time."""
        blocks = []
        for ps in points:
            self.set_params_from_optimised(ps)
            blocks.append(self.engine_block())
        return blocks


class ModelWithFrequencies(Model):
    """Adds equilibrium frequencies, either optimised (K-1 ratios to the last one) or smoothed by one scalar."""

    def __init__(self, states, forest_stats, sf=None, frequencies=None, tau=0, optimise_tau=False,
                 frequency_smoothing=False, parameter_file=None, reoptimise=False, **kwargs):
        self._frequencies = None
        self._optimise_frequencies = not frequency_smoothing
        self._frequency_smoothing = frequency_smoothing
        Model.__init__(self, states, forest_stats=forest_stats, sf=sf, tau=tau, optimise_tau=optimise_tau,
                       reoptimise=reoptimise, parameter_file=parameter_file, **kwargs)
        if self._frequencies is None:
            self._frequencies = frequencies if frequencies is not None \
                else np.ones(len(states), dtype=np.float64) / len(states)

    @property
    def frequencies(self):
        return self._frequencies

    @frequencies.setter
    def frequencies(self, frequencies):
        self._set_frequencies(frequencies)

    def _set_frequencies(self, frequencies):
        if not (self._optimise_frequencies or self._frequency_smoothing):
            raise NotImplementedError('The frequencies are preset and cannot be changed.')
        self._frequencies = frequencies

    def _extra_count(self):
        if self._optimise_frequencies:
            return len(self.frequencies) - 1
        return 1 if self._frequency_smoothing else 0

    def get_num_params(self):
        return self._basic_count() + self._extra_count()

    def extra_params_fixed(self):
        return self._extra_params_fixed or self._basic_count() == self.get_num_params()

    def basic_params_fixed(self):
        return not self._basic_count()

    def set_params_from_optimised(self, ps, **kwargs):
        Model.set_params_from_optimised(self, ps, **kwargs)
        if self.extra_params_fixed():
            return
        first = self._basic_count()
        if self._optimise_frequencies:
            freqs = np.hstack((ps[first: first + len(self.frequencies) - 1], [1.]))
            freqs /= freqs.sum()
            self.frequencies = freqs
        elif self._frequency_smoothing:
            # cumulative on purpose: the reference smooths the *current* frequencies (models/__init__.py:331-335)
            freqs = self.frequencies * self.forest_stats.num_tips
            freqs += ps[first]
            freqs /= freqs.sum()
            self.frequencies = freqs

    def get_optimised_parameters(self):
        basic = Model.get_optimised_parameters(self)
        if self.extra_params_fixed():
            return basic
        if self._optimise_frequencies:
            return np.hstack((basic, self.frequencies[:-1] / self.frequencies[-1]))
        return np.hstack((basic, [0] if self._frequency_smoothing else []))

    def _frequency_bounds(self):
        extras = []
        if self._optimise_frequencies:
            extras += [np.array([1e-6, 10e6], np.float64)] * (len(self.frequencies) - 1)
        if self._frequency_smoothing:
            extras.append(np.array([0, self.forest_stats.num_nodes]))
        return extras

    def get_bounds(self):
        if self.extra_params_fixed():
            return Model.get_bounds(self)
        return np.array((*self._basic_bounds(), *self._frequency_bounds()))

    def parse_parameters(self, params, reoptimise=False):
        """Besides the basic parameters, frequencies keyed by state name (``models/__init__.py:365-414``)."""
        params = Model.parse_parameters(self, params, reoptimise)
        known = set(self.states) & set(params.keys())
        if not known:
            return params
        unknown = [s for s in self.states if s not in params]
        if unknown and not reoptimise:
            _log.error('Frequencies for some of the states ({}) are missing, ignoring the specified frequencies.'
                       .format(', '.join(unknown)))
            return params
        self._frequencies = np.array([params[s] if s in params else 0 for s in self.states])
        try:
            self._frequencies = self._frequencies.astype(np.float64)
        except Exception:
            _log.error('Could not convert the frequencies given in parameters ({}) to float, ignoring them.'
                       .format(self._frequencies))
            self._frequencies = None
            return params
        if np.round(self._frequencies.sum() - 1, 2) != 0 and not reoptimise:
            _log.error('Frequencies given in parameters ({}) do not sum up to one ({}), ignoring them.'
                       .format(self._frequencies, self._frequencies.sum()))
            self._frequencies = None
        elif np.any(self._frequencies < 0) and not reoptimise:
            _log.error('Some of the frequencies given in parameters ({}) are negative, ignoring them.'
                       .format(self._frequencies))
            self._frequencies = None
        else:
            if unknown:
                _log.error('Frequencies for some of the states ({}) are missing from parameters.'
                           .format(', '.join(unknown)))
            # the given values are used as they are (the reference computes a clamped, renormalised copy and
            # drops it, models/__init__.py:407-408)
            self._optimise_frequencies = reoptimise and not self._frequency_smoothing
        return params

    def freeze(self):
        Model.freeze(self)
        self._optimise_frequencies = False
        self._frequency_smoothing = False

    def save_parameters(self, filehandle):
        Model.save_parameters(self, filehandle)
        for state, frequency in zip(self.states, self.frequencies):
            filehandle.write('{}\t{}\n'.format(state, frequency))

    def _print_parameters(self):
        tag = '(optimised)' if self._optimise_frequencies else '(smoothed)' if self._frequency_smoothing else '(fixed)'
        rows = '\n'.join('\t\t{}:\t{:g}'.format(s, f) for s, f in zip(self.states, self.frequencies))
        return '{}\tfrequencies\t{}\n{}\n'.format(Model._print_parameters(self), tag, rows)


def __getattr__(name):
    """The model classes live in their own modules, as in pastml (``from pastml.models.F81Model import F81Model``);
    ``pastml_b200.models.model2class`` and the class names are resolved on first use to keep the imports acyclic."""
    if name == 'model2class':
        from pastml_b200.models.F81Model import F81Model
        from pastml_b200.models.JCModel import JCModel
        from pastml_b200.models.EFTModel import EFTModel
        from pastml_b200.models.HKYModel import HKYModel
        from pastml_b200.models.CustomRatesModel import CustomRatesModel
        from pastml_b200.models.JTTModel import JTTModel
        table = {F81: F81Model, JC: JCModel, CUSTOM_RATES: CustomRatesModel, HKY: HKYModel, JTT: JTTModel, EFT: EFTModel}
        globals()['model2class'] = table
        return table
    for module, names in (('HKYModel', ('HKY_STATES',)), ('JTTModel', ('JTT_STATES', 'JTT_FREQUENCIES', 'JTT_RATE_MATRIX')),
                          ('generator', ('save_matrix', 'get_diagonalisation', 'get_normalised_generator', 'get_pij_matrix')),
                          ('CustomRatesModel', ('load_custom_rates',))):
        if name in names:
            import importlib
            return getattr(importlib.import_module('pastml_b200.models.' + module), name)
    raise AttributeError('module {!r} has no attribute {!r}'.format(__name__, name))

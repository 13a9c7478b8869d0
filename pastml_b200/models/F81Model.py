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

    def engine_blocks(self, points):
        """All the blocks of a finite-difference batch in a handful of array operations -- the same floating-point
        operations, row by row, as the point-by-point route (row sums of a C-contiguous array use numpy's pairwise
        summation per row like the 1-D sum; the dot product of ``get_mu`` stays a per-row BLAS call); checked bit for
        bit in tests/test_host_cpu.py.  Smoothing (tau or the frequency-smoothing scalar in the vector) takes the
        point-by-point route."""
        points = list(points)
        if not points or self._optimise_tau or self._frequency_smoothing or type(self)._engine_fields is not F81Model._engine_fields:
            return ModelWithFrequencies.engine_blocks(self, points)
        from pastml_b200.engine import param_stride
        pts = np.asarray(points, dtype=np.float64)
        n, K = len(pts), len(self.states)
        out = np.zeros((n, param_stride(self.engine_kind, K)), dtype=np.float64)
        out[:, 0] = pts[:, 0] if self._optimise_sf else float(self.sf)
        out[:, 1] = float(self.tau)
        out[:, 2] = float(self._tau_factor)
        if not self.extra_params_fixed() and self._optimise_frequencies:
            first = self._basic_count()
            freqs = np.empty((n, K), dtype=np.float64)
            freqs[:, :K - 1] = pts[:, first: first + K - 1]
            freqs[:, K - 1] = 1.
            freqs /= freqs.sum(axis=1)[:, np.newaxis]
        else:
            freqs = np.broadcast_to(np.asarray(self.frequencies, dtype=np.float64), (n, K))
        with np.errstate(divide='ignore'):
            out[:, 3] = [np.float64(1.) / (1. - f.dot(f)) for f in freqs]
        out[:, 4:4 + K] = freqs
        self.set_params_from_optimised(points[-1])   # the state the point-by-point loop leaves behind
        return list(out)

"""Rate matrices of the eigen-decomposition models (CUSTOM_RATES, JTT): the generator normalised to one expected change
per unit of time, its spectral form and P(t) -- the host side of ``EigLane`` in ``csrc/pml_kernels.cuh``.  Same
floating-point operations in the same order as ``pastml/models/generator.py`` (the decomposition is uploaded to the
device as it comes out of here, so the log-likelihoods inherit its bits)."""
import numpy as np
import pandas as pd  # noqa: F401


def _rates_or_equal(n_states, rate_matrix):
    """The exchangeability matrix; every pair of distinct states exchanges at rate 1 when none is given."""
    if rate_matrix is not None:
        return rate_matrix
    return np.ones(shape=(n_states, n_states), dtype=np.float64) - np.eye(n_states)


def get_normalised_generator(frequencies, rate_matrix=None):
    """Q with q_ij = r_ij pi_j (i != j), zero row sums, and -sum_i pi_i q_ii = 1."""
    q = _rates_or_equal(len(frequencies), rate_matrix) * frequencies
    q -= np.diag(q.sum(axis=1))
    expected_changes = -q.diagonal().dot(frequencies)
    q /= expected_changes
    return q


def get_diagonalisation(frequencies, rate_matrix=None):
    """(d, A, A^-1) with A diag(d) A^-1 = Q, from ``numpy.linalg.eig`` (``generator.py:16-30``)."""
    eigenvalues, right = np.linalg.eig(get_normalised_generator(frequencies, rate_matrix))
    return eigenvalues, right, np.linalg.inv(right)


def get_pij_matrix(t, diag, A, A_inv):
    """P(t) = A diag(exp(d t)) A^-1."""
    decay = np.diag(np.exp(diag * t))
    return A.dot(decay).dot(A_inv)


def save_matrix(states, matrix, outfile):
    """Text form of a rate matrix: '# state names' header, then one row per line."""
    np.savetxt(outfile, matrix, header=' '.join(states), fmt='%.18e', delimiter=' ')

"""Rate matrices: normalised generator, eigen-decomposition, P(t) (``pastml/models/generator.py``)."""
import numpy as np
import pandas as pd  # noqa: F401


def save_matrix(states, matrix, outfile):
    """Writes a rate matrix with a '# state names' header line."""
    np.savetxt(outfile, matrix, delimiter=' ', fmt='%.18e', header=' '.join(states))


def get_normalised_generator(frequencies, rate_matrix=None):
    """Q_ij = r_ij pi_j off the diagonal, rows summing to zero, scaled to one expected change per time unit."""
    if rate_matrix is None:
        n = len(frequencies)
        rate_matrix = np.ones(shape=(n, n), dtype=np.float64) - np.eye(n)
    generator = rate_matrix * frequencies
    generator -= np.diag(generator.sum(axis=1))
    generator /= -generator.diagonal().dot(frequencies)
    return generator


def get_diagonalisation(frequencies, rate_matrix=None):
    """(d, A, A^-1) with A diag(d) A^-1 = the normalised generator (``numpy.linalg.eig`` like the reference,
    ``generator.py:16-30``)."""
    d, A = np.linalg.eig(get_normalised_generator(frequencies, rate_matrix))
    return d, A, np.linalg.inv(A)


def get_pij_matrix(t, diag, A, A_inv):
    return A.dot(np.diag(np.exp(diag * t))).dot(A_inv)

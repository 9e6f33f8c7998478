"""Single-matrix inverse operators: the `svd` / `chol` seam of the reference
(cosmogp/Gaussian_process.py:6-9, cosmogp/inv_matrix.py).

`cholesky_inverse` runs on the GPU (blocked FP64 Cholesky, triangular inverse and
L^-T L^-1 on the DMMA pipe).  `svd_inverse` is kept, on the host with the same scipy
calls, ONLY as the reference check north_star asks for -- nothing in the batched
hot path calls it.
"""
import numpy as np
from scipy import linalg


def svd_inverse(matrix, return_logdet=False):
    """cosmogp/inv_matrix.py:4-18 (host reference check): pseudo-inverse keeping
    singular values above 1e-15; logdet over the kept ones.  Does not print."""
    u, s, v = linalg.svd(matrix)
    keep = s > 10 ** -15
    inv = np.dot(v.T[:, keep], np.dot(np.diag(1. / s[keep]), u.T[keep]))
    if return_logdet:
        return inv, np.sum(np.log(s[keep]))
    return inv


def cholesky_inverse(matrix, return_logdet=False):
    """cosmogp/inv_matrix.py:21-31 on the device: K^-1 = L^-T L^-1, logdet = sum 2 log L_ii.
    Raises numpy.linalg.LinAlgError when the matrix is not positive definite."""
    from . import dense
    return dense.cholesky_inverse(matrix, return_logdet=return_logdet)

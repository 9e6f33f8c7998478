"""Kernel callables with the reference signatures (cosmogp/kernel.py): the
`kernel(x, hyperparameter, new_x=None, nugget=0., floor=0., y_err=None)` seam of
cosmogp/Gaussian_process.py:136-154.  The matrices are built on the device by a
streaming kernel; the batched hot path never calls these (it generates covariance
tiles on chip), they exist for users of the operator seam and for large objects.
"""
import numpy as np


def init_rbf(x, y):
    """Initial [sigma, l] guess, cosmogp/kernel.py:6-22, including its quirk (Q6): the
    loop overwrites L_min / L_max / sigma, so only the LAST object's extent and scatter
    are used, while the point counts of all objects enter d.  O(B) host scalars."""
    number_point = np.array([len(xi) for xi in x], dtype=float)
    last_x, last_y = np.asarray(x[len(y) - 1]), np.asarray(y[len(y) - 1])
    L_min, L_max, sigma = np.min(last_x), np.max(last_x), np.std(last_y)
    d = np.mean(np.sqrt((L_max - L_min) ** 2 / number_point))
    L = np.mean(L_max - L_min)
    return np.mean(sigma), np.mean([d, L])


def rbf_kernel_1d(x, hyperparameter, new_x=None, nugget=0., floor=0.00, y_err=None):
    """cosmogp/kernel.py:25-77.  new_x None: (N,N) auto-covariance with
    y_err^2+floor^2+nugget^2 on the diagonal; else (len(new_x), N) cross-covariance."""
    from . import dense
    return dense.covariance(x, hyperparameter, 1, new_x=new_x, nugget=nugget, floor=floor, y_err=y_err)


def rbf_kernel_2d(x, hyperparameter, new_x=None, nugget=0., floor=0.00, y_err=None, flags=0):
    """cosmogp/kernel.py:80-155 at HEAD: the auto-covariance carries NO sigma^2 (unit
    diagonal + noise) while the cross-covariance does (quirk Q2); pass
    flags=CGP_AMP_ON_AUTOCOV for the corrected form.  Does not print."""
    from . import dense
    return dense.covariance(x, hyperparameter, 2, new_x=new_x, nugget=nugget, floor=floor, y_err=y_err, flags=flags)

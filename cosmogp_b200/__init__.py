"""cosmogp_b200: the Gaussian-process hot path of PFLeget/cosmogp on B200 (sm_100a).

Same public names as cosmogp/__init__.py:12-25.  All arithmetic of the hot path runs in
libcosmogp_b200.so (hand-written CUDA behind a C ABI, include/cosmogp_b200.h); there is
no CPU fallback -- without the library or a GPU the compute calls raise.
"""
from .inv_matrix import svd_inverse
from .inv_matrix import cholesky_inverse

from .mean import return_mean

from .gp import Gaussian_process
from .gp import gaussian_process
from .gp import gaussian_process_nobject

from .kernel import init_rbf
from .kernel import rbf_kernel_1d
from .kernel import rbf_kernel_2d

from .pull import build_pull

from .batch import DeviceBatch, pack_csr

__all__ = ["svd_inverse", "cholesky_inverse", "return_mean", "Gaussian_process", "gaussian_process",
           "gaussian_process_nobject", "init_rbf", "rbf_kernel_1d", "rbf_kernel_2d", "build_pull",
           "DeviceBatch", "pack_csr"]

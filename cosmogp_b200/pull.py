"""Pull and residual of the GP interpolation: drop-in for cosmogp/pull.py build_pull.

The reference refits the GP N times per object on N-1 points (pull.py:66-90, O(N^4));
here one factorisation per object gives every leave-one-out prediction and variance in
closed form from K^-1, all objects in one launch.
"""
import numpy as np

from . import _lib
from . import mean as _mean
from .batch import DeviceBatch, RaggedView, pack_csr


class build_pull:

    def __init__(self, y, x, hyperparameters, nugget=0.,
                 y_err=None, y_mean=None, x_axis_mean=None,
                 kernel='RBF1D'):
        """Same arguments and attributes as cosmogp/pull.py:10-40.  `hyperparameters` (and `nugget`) may
        also be per-object arrays of shape (n_object, n_hyp) / (n_object,): each object is then
        pulled with its own fit, as in the reference's per-object notebook loop."""
        self.y = y
        self.x = x
        self.hyperparameters = hyperparameters

        self.y_err = y_err
        self.y_mean = y_mean
        self.x_axis_mean = x_axis_mean

        self.kernel = kernel
        self.nugget = nugget

        self.n_object = len(y)

        empty = np.zeros(1, dtype=np.int64)
        self._pull = RaggedView(np.zeros(0), empty)       # per object, list-like (views, no 10^6-element list)
        self.pull = np.zeros(0)          # flat, grows with every compute_pull call like the reference lists
        self.residual = np.zeros(0)
        self.prediction = RaggedView(np.zeros(0), empty)

        self.pull_average = None
        self.pull_std = None
        self.flags = 0

    def compute_pull(self, diff=None, svd_method=True, substract_mean=False):
        """Pulls and residuals (pull.py:43-102).  Mean handling, per object sn:
          no y_mean, substract_mean False : plain leave-one-out of y              (mode A)
          y_mean, diff None               : offset re-estimated on the kept points (mode B)
          y_mean, diff given              : fixed offset diff[sn]                  (mode C)
          no y_mean, substract_mean True  : the mean of the FIRST object becomes the
                                            template for all (pull.py:71-73, quirk Q8)  (mode D)
        pred_var = |cov_tt| and the pull denominator counts nugget^2 twice (pull.py:90-93)."""
        assert self.kernel in ['RBF1D', 'RBF2D'], '%s is not in implemented kernel' % (self.kernel)
        dim = 1 if self.kernel == 'RBF1D' else 2
        x_flat, off = pack_csr(self.x, dim)
        y_flat, _ = pack_csr(self.y, 1)
        ye_flat = pack_csr(self.y_err, 1)[0] if self.y_err is not None else None

        if self.y_mean is None and substract_mean and self.n_object:
            self.y_mean = np.ones_like(self.y[0]) * np.mean(self.y[0])
            self.x_axis_mean = self.x[0]

        template = None
        mode = _lib.CGP_LOO_PLAIN
        if self.y_mean is not None:
            template = (_mean.interpolate_mean_1d if dim == 1 else _mean.interpolate_mean_2d)(
                self.x_axis_mean, self.y_mean, x_flat)
            if diff is None:
                mode = _lib.CGP_LOO_RECENTER
            else:
                template = template + np.repeat(np.asarray(diff, dtype=float), np.diff(off))

        batch = DeviceBatch(x_flat, y_flat, off, y0=template, y_err=ye_flat, dim=dim)
        pred, pvar, pull, resid, info = batch.loo(self.hyperparameters, self.nugget, mode=mode, flags=self.flags)
        bad = np.nonzero(info)[0]
        if len(bad):
            raise np.linalg.LinAlgError("covariance of object %d is not positive definite" % int(bad[0]))

        self._pull = self._pull.extended(pull, off)
        self.prediction = self.prediction.extended(pred, off)
        self.prediction_variance = pvar
        self.pull = pull if not len(self.pull) else np.concatenate([self.pull, pull])
        self.residual = resid if not len(self.residual) else np.concatenate([self.residual, resid])

        # scipy.stats.norm.fit (pull.py:102) = sample mean and population standard deviation
        self.pull_average = float(np.mean(self.pull))
        self.pull_std = float(np.sqrt(np.mean((self.pull - self.pull_average) ** 2)))

    def plot_result(self, binning=60):
        """Histogram of the pulls with the fitted normal law (pull.py:105-140)."""
        import pylab as plt
        from scipy.stats import norm as normal
        plt.figure()
        plt.hist(self.pull, bins=binning, density=True)
        xmin, xmax = plt.xlim()
        _max = max([abs(xmin), abs(xmax)])
        plt.xlim(-_max, _max)
        xaxis = np.linspace(-_max, _max, 100)
        plt.plot(xaxis, normal.pdf(xaxis, self.pull_average, self.pull_std), 'r', linewidth=3)
        plt.title(r"Fit results: $\mu$ = $ %.2f \pm %.2f $, $\sigma$ = $ %.2f \pm %.2f $" % (
            self.pull_average, self.pull_std / np.sqrt(len(self.pull)),
            self.pull_std, self.pull_std / np.sqrt(2 * len(self.pull))))
        plt.ylabel('Number of points (normed)')
        plt.xlabel('Pull')

"""Mean function / detrending on the host (cosmogp/mean.py:7-104).

Outside the CUDA hot path: it runs once per object at
construction, not per likelihood evaluation.  The arithmetic stays in the same scipy
routines the reference calls (FITPACK via InterpolatedUnivariateSpline / bisplrep), so
y0 is bit-identical; what changes is that a template shared by all objects is fitted
ONCE and evaluated for every epoch of every object in one call, instead of rebuilding
the spline 10^5 times (mean.py:28-31).
"""
import numpy as np
import scipy.interpolate as inter


def interpolate_mean_1d(old_binning, mean_function, new_binning):
    """mean.py:7-33: cubic interpolating spline of the template on a new grid."""
    return inter.InterpolatedUnivariateSpline(old_binning, mean_function)(new_binning)


def interpolate_mean_2d(old_binning, mean_function, new_binning):
    """mean.py:36-67: bivariate spline (bisplrep, task=1), evaluated point by point."""
    old_binning = np.asarray(old_binning)
    new_binning = np.asarray(new_binning)
    tck = inter.bisplrep(old_binning[:, 0], old_binning[:, 1], mean_function, task=1)
    return np.array([inter.bisplev(p[0], p[1], tck) for p in new_binning])


def _is_1d(x):
    return type(x[0]) is np.float64          # the reference's dispatch (mean.py:81,95; quirk Q10)


def return_mean(y, x, new_x=None, mean_y=None, mean_xaxis=None, diff=None):
    """mean.py:70-104 for one object: template interpolated at x plus `diff`
    (default: mean of y - template); returned at x, or at new_x when given."""
    if mean_y is not None:
        assert mean_xaxis is not None, 'you should provide an x axis for the average'
        assert len(mean_y) == len(mean_xaxis), 'mean_y and mean_xaxis should have the same len'
        interp = interpolate_mean_1d if _is_1d(x) else interpolate_mean_2d
        shape = interp(mean_xaxis, mean_y, x)
    else:
        shape = 0
    if diff is None:
        diff = np.mean(y - shape)
    y0 = shape + diff
    if new_x is None:
        return y0
    if mean_y is None:
        return y0
    return interp(mean_xaxis, mean_y, new_x) + diff


def _segment_means(values, off):
    sizes = np.diff(off)
    b = len(sizes)
    if b and (sizes == sizes[0]).all() and sizes[0] > 0:
        return values.reshape(b, sizes[0]).mean(axis=1)
    if b <= 20000:
        return np.array([np.mean(values[off[i]:off[i + 1]]) for i in range(b)])
    safe = np.minimum(off[:-1], max(len(values) - 1, 0))
    sums = np.add.reduceat(values, safe) if len(values) else np.zeros(b)
    return np.where(sizes > 0, sums / np.maximum(sizes, 1), np.nan)


def batched_mean(x_flat, y_flat, off, dim, mean_y, mean_xaxis, diff):
    """All objects at once -> (y0 flat, diff per object float64[B]).
    `diff` is the reference's per-object array: entries may be None (estimate)."""
    b = len(off) - 1
    if mean_y is not None:
        assert mean_xaxis is not None, 'you should provide an x axis for the average'
        assert len(mean_y) == len(mean_xaxis), 'mean_y and mean_xaxis should have the same len'
        if dim == 1:
            shape = interpolate_mean_1d(mean_xaxis, mean_y, x_flat) if len(x_flat) else np.zeros(0)
        else:
            shape = interpolate_mean_2d(mean_xaxis, mean_y, x_flat)
    else:
        shape = np.zeros(len(y_flat))
    d = np.empty(b)
    given = np.array([v is not None for v in diff], dtype=bool) if diff is not None else np.zeros(b, dtype=bool)
    if not given.all():
        d[:] = _segment_means(y_flat - shape, off)
    if given.any():
        d[given] = np.array([float(diff[i]) for i in np.nonzero(given)[0]])
    return shape + np.repeat(d, np.diff(off)), d


def template_on_grid(grid, dim, mean_y, mean_xaxis):
    """Template evaluated on a prediction grid (zeros when there is no template)."""
    if mean_y is None:
        return None
    interp = interpolate_mean_1d if dim == 1 else interpolate_mean_2d
    return interp(mean_xaxis, mean_y, grid)

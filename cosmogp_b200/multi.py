"""A batch of GP objects sharded over the GPUs of one box, inside ONE process (cgp_ctx_* of the C ABI).

What `DeviceBatch` is for one GPU: the data of a reference `Gaussian_process` (cosmogp/Gaussian_process.py:156-186)
uploaded once -- here cut into contiguous ranges of objects balanced by sum N^3, one range resident per GPU.  Objects
never interact except through the scalar sum of their log-likelihoods (:205-213), so nothing is exchanged on the data
path; per-object outputs come back either over every GPU's own PCIe link (gather=False) or through a final NCCL
gather on GPU 0 over NVLink (gather=True).  numpy in, numpy out; no torchrun, no per-rank copies of the inputs.
"""
import ctypes as C
import glob
import os

import numpy as np

from . import _lib


def _nccl_library():
    """the libnccl.so.2 PyTorch ships (already mapped into the process once torch is imported)"""
    try:
        import nvidia.nccl as pkg
        for root in list(getattr(pkg, "__path__", [])):
            hits = glob.glob(os.path.join(root, "lib", "libnccl.so*"))
            if hits:
                return sorted(hits)[0]
    except Exception:
        pass
    return None


class ShardedBatch:
    """x: flat (sumN,) or (sumN, 2); y, y0, y_err: flat (sumN,) host arrays (y0 / y_err may be None); off: int64 (B+1,).
    devices: None / 'all' (every visible GPU), an int (the first k GPUs) or a list of device ids."""

    def __init__(self, x, y, off, y0=None, y_err=None, dim=1, devices=None):
        _lib.require_device()
        L = _lib.lib()
        path = _nccl_library()
        if path:
            L.cgp_set_nccl_library(path.encode())
        ids = None
        if devices is None or devices == "all":
            n = 0
        elif isinstance(devices, (int, np.integer)):
            n = int(devices)
        else:
            ids = np.ascontiguousarray(list(devices), dtype=np.int32)
            n = len(ids)
        self._ctx = C.c_void_p()
        _lib.check(L.cgp_ctx_create(n, None if ids is None else ids.ctypes.data, C.byref(self._ctx)), "cgp_ctx_create")
        nd, nccl = C.c_int(0), C.c_int(0)
        L.cgp_ctx_info(self._ctx, C.byref(nd), C.byref(nccl))
        self.n_devices, self.have_nccl = nd.value, bool(nccl.value)
        self.dim = int(dim)
        self.off_host = np.ascontiguousarray(off, dtype=np.int64)
        self.n_obj = len(self.off_host) - 1
        self.n_pts = int(self.off_host[-1]) if self.n_obj > 0 else 0
        self.max_n = int(np.diff(self.off_host).max()) if self.n_obj > 0 else 0
        f = _lib.f64
        x, y, y0, y_err = f(x), f(y), f(y0), f(y_err)
        self._batch = C.c_void_p()
        _lib.check(L.cgp_ctx_batch_create(self._ctx, self.n_obj, _lib.hptr(self.off_host), self.dim, _lib.hptr(x), _lib.hptr(y),
                                          _lib.hptr(y0), _lib.hptr(y_err), C.byref(self._batch)), "cgp_ctx_batch_create")
        self.h2d_bytes = 8 * (self.n_obj + 1 + sum(a.size for a in (x, y, y0, y_err) if a is not None))
        self.d2h_bytes = 0
        starts = np.zeros(self.n_devices + 1, dtype=np.int64)
        L.cgp_ctx_batch_ranges(self._batch, starts.ctypes.data)
        self.ranges = [(int(starts[d]), int(starts[d + 1])) for d in range(self.n_devices)]
        self._tot = C.c_double(0.0)
        self._last = None

    def close(self):
        L = _lib.lib()
        if getattr(self, "_batch", None) is not None and self._batch.value:
            L.cgp_ctx_batch_destroy(self._batch); self._batch = C.c_void_p()
        if getattr(self, "_ctx", None) is not None and self._ctx.value:
            L.cgp_ctx_destroy(self._ctx); self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _hyp(self, hyp):
        h = np.ascontiguousarray(np.asarray(hyp, dtype=np.float64).ravel())
        need = 2 if self.dim == 1 else 4
        assert len(h) == need, "expected %d hyperparameters, got %d" % (need, len(h))
        return h

    # ---- the hot path, same names as DeviceBatch
    def log_likelihood_total(self, hyp, nugget=0.0, floor=0.0, flags=0):
        """-> (sum over all objects, number of non-positive-definite objects): 16 bytes per GPU cross PCIe."""
        h = self._hyp(hyp)
        rc = _lib.lib().cgp_ctx_batch_ll(self._batch, _lib.hptr(h), float(nugget), float(floor), int(flags),
                                         C.byref(self._tot), None, None)
        _lib.check(rc, "cgp_ctx_batch_ll")
        self.d2h_bytes += 16 * self.n_devices
        self._last = (h, float(nugget), float(floor), int(flags))
        return float(self._tot.value), int(rc)

    def _ll_full(self):
        if self._last is None:
            raise RuntimeError("no likelihood has been evaluated on this batch yet (call log_likelihood_total first)")
        h, nugget, floor, flags = self._last
        ll = np.empty(max(self.n_obj, 1)); info = np.zeros(max(self.n_obj, 1), dtype=np.int32)
        _lib.check(_lib.lib().cgp_ctx_batch_ll(self._batch, _lib.hptr(h), nugget, floor, flags, C.byref(self._tot),
                                               ll.ctypes.data, info.ctypes.data), "cgp_ctx_batch_ll")
        return ll[:self.n_obj], info[:self.n_obj]

    def ll_host(self):
        return self._ll_full()[0]

    def info_host(self):
        return self._ll_full()[1]

    def predict(self, hyp, nugget, grid, goff=None, new_y0=None, want_var=True, floor=0.0, flags=0, mean_template=None,
                gather=False, with_ll=False):
        """Shared grid only.  -> mean (B, M), var (B, M) or None, info (B,) [, ll (B,) with with_ll=True]."""
        assert goff is None, "the multi-GPU batch predicts on a shared grid (new_binning given)"
        h = self._hyp(hyp)
        g = np.ascontiguousarray(grid, dtype=np.float64)
        m = int(g.shape[0])
        ny0 = None
        if mean_template is not None:
            tmpl, diff = mean_template
            ny0 = np.concatenate([np.asarray(tmpl, dtype=np.float64).ravel(), np.asarray(diff, dtype=np.float64).ravel()])
            assert ny0.size == m + self.n_obj
            flags = int(flags) | _lib.CGP_MEAN_TEMPLATE
        elif new_y0 is not None:
            ny0 = np.ascontiguousarray(new_y0, dtype=np.float64).reshape(self.n_obj, m)
        if self.dim == 1:
            flags = int(flags) | _lib.CGP_GRID_UNIFORM
        mean = np.empty((max(self.n_obj, 1), m)); var = np.empty((max(self.n_obj, 1), m)) if want_var else None
        info = np.zeros(max(self.n_obj, 1), dtype=np.int32)
        ll = np.empty(max(self.n_obj, 1)) if with_ll else None
        rc = _lib.lib().cgp_ctx_batch_predict(self._batch, _lib.hptr(h), float(nugget), float(floor), int(flags), _lib.hptr(g), m,
                                              _lib.hptr(ny0), _lib.hptr(ll), _lib.hptr(mean), _lib.hptr(var), info.ctypes.data,
                                              1 if gather else 0)
        _lib.check(rc, "cgp_ctx_batch_predict")
        self.d2h_bytes += 8 * self.n_obj * m * (2 if want_var else 1)
        out = (mean[:self.n_obj], var[:self.n_obj] if want_var else None, info[:self.n_obj])
        return out + (ll[:self.n_obj],) if with_ll else out

    def loo(self, hyp, nugget, mode=_lib.CGP_LOO_PLAIN, floor=0.0, flags=0, gather=False):
        """Closed-form leave-one-out of every object -> pred, pred_var, pull, resid (flat host arrays), info,
        (sum of pulls, sum of squared pulls)."""
        h = self._hyp(hyp)
        outs = [np.empty(max(self.n_pts, 1)) for _ in range(4)]
        info = np.zeros(max(self.n_obj, 1), dtype=np.int32)
        mom = np.zeros(2)
        rc = _lib.lib().cgp_ctx_batch_loo(self._batch, _lib.hptr(h), float(nugget), float(floor), int(flags), int(mode),
                                          *[o.ctypes.data for o in outs], info.ctypes.data, mom.ctypes.data, 1 if gather else 0)
        _lib.check(rc, "cgp_ctx_batch_loo")
        self.d2h_bytes += 4 * 8 * self.n_pts
        return [o[:self.n_pts] for o in outs] + [info[:self.n_obj], (float(mom[0]), float(mom[1]))]


def shard_ranges(off, n_parts):
    """contiguous object ranges with near-equal sum N^3 (cgp_shard_ranges; host only, no GPU needed)"""
    off = np.ascontiguousarray(off, dtype=np.int64)
    starts = np.zeros(n_parts + 1, dtype=np.int64)
    _lib.check(_lib.lib().cgp_shard_ranges(len(off) - 1, off.ctypes.data, int(n_parts), starts.ctypes.data), "cgp_shard_ranges")
    return [(int(starts[i]), int(starts[i + 1])) for i in range(n_parts)]

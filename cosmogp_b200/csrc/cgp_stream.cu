// Host-resident batches, pipelined: the chunk loop of "upload -> LL -> factor -> grid -> download"
// over several CUDA streams, written against the public C ABI.  One call per step from the host
// language; every copy and launch inside is a native call (a Python loop over the same 500 calls
// costs more host time than the kernels take).
//
// Replaces, for a batch that lives in host memory, the two per-object loops of the reference:
// Gaussian_process.compute_log_likelihood (cosmogp/Gaussian_process.py:205-213) and get_prediction
// (:270-361), evaluated at the same hyperparameters.
#include "../../include/cosmogp_b200.h"
#include "cgp_internal.h"

#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

struct cgp_streamer {
  int dev = 0, dim = 1, n_pts = 0, n_streams = 0;
  int64_t chunk = 0, m_grid = 0;
  bool two_kernel = false;
  struct Slot {
    cudaStream_t st = nullptr;
    // x owns the input block (rows x | y | y_err | y0), mean the output block (rows mean | var); y, ye, y0, var point into them
    double *x = nullptr, *y = nullptr, *y0 = nullptr, *ye = nullptr, *ny0 = nullptr, *mean = nullptr, *var = nullptr, *ws = nullptr;
    cudaEvent_t up_done = nullptr, kern_done = nullptr, dn_done = nullptr;
  };
  std::vector<Slot> slots;
  int64_t* off = nullptr;      // iota * n_pts, shared by all slots
  double* grid = nullptr;
  // all uploads go through ONE stream and all downloads through another, in chunk order: copies issued from
  // several streams share the link, so the first (small) chunk would arrive no sooner than the ones behind it
  cudaStream_t up = nullptr, dn = nullptr;
  double* spl_t = nullptr; double* spl_c = nullptr; int spl_n = 0;     // mean template as a cubic B-spline (optional)
  // Few and large copies: with both directions of the link busy every additional copy costs ~20 us of transfer time
  // (tools/microbench_wc.cu: 144 MB each way take 3.0 ms as one copy, 3.7 ms as 13, 4.2 ms as 52).  So the per-object
  // scalars live in arrays for the WHOLE batch (the log-likelihoods and info flags leave in one copy each at the end of
  // a run, the mean template and its per-object offsets arrive in one copy at the start), and the per-point inputs /
  // per-grid-point outputs of a chunk travel as ONE two-dimensional copy each when the caller's arrays are rows of one
  // pinned block (x, y, y_err [, y0] and mean, var at a constant pitch: what StreamedEvaluator allocates).
  double* ll_all = nullptr; int* info_all = nullptr; double* tmpl_all = nullptr; int64_t cap_all = 0;
};

namespace {

int sfail(int code, const char* what, cudaError_t e = cudaSuccess) {
  if (e != cudaSuccess) return cgp::fail(code, "%s: CUDA error %d (%s)", what, (int)e, cudaGetErrorString(e));
  return cgp::fail(code, "%s", what);
}

struct DeviceGuard {
  int prev = -1; bool switched = false;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) switched = cudaSetDevice(dev) == cudaSuccess;
  }
  ~DeviceGuard() { if (switched) cudaSetDevice(prev); }
};

__global__ void iota_scaled_kernel(int64_t* out, int64_t n, int64_t scale) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = i * scale;
}

// Chunk sizes for one run.  Only the first upload and the last download are exposed, so the chunks
// ramp up from 2048 objects (x1.6: the next upload hides behind the current kernels), stay at the buffer
// capacity in the middle (full-size launches run at full efficiency) and ramp down at the end (x2).
std::vector<int64_t> chunk_schedule(int64_t n_obj, int64_t cap, bool ramp) {
  std::vector<int64_t> sz;
  const int64_t lo = 2048;
  std::vector<int64_t> up, down;
  int64_t ramps = 0;
  if (ramp && cap > 2 * lo) {
    for (int64_t c = lo; c < cap; c = c * 8 / 5) { up.push_back(c); ramps += c; }
    for (int64_t c = lo; c < cap; c *= 2) { down.push_back(c); ramps += c; }
    while (ramps > n_obj / 2 && (!up.empty() || !down.empty())) {      // small batches: shorter ramps
      std::vector<int64_t>& v = (!up.empty() && (down.empty() || up.back() >= down.back())) ? up : down;
      ramps -= v.back(); v.pop_back();
    }
  }
  for (int64_t c : up) sz.push_back(c);
  const int64_t mid = n_obj - ramps;
  if (mid > 0) {
    const int64_t k = (mid + cap - 1) / cap, each = (mid + k - 1) / k;
    for (int64_t i = 0, left = mid; i < k; ++i) { const int64_t c = left < each ? left : each; sz.push_back(c); left -= c; }
  }
  for (size_t i = down.size(); i-- > 0;) sz.push_back(down[i]);
  return sz;
}

template <class T> cudaError_t dalloc(T** p, size_t count) { return cudaMalloc((void**)p, (count ? count : 1) * sizeof(T)); }

void release(cgp_streamer* s) {
  if (!s) return;
  DeviceGuard g(s->dev);
  for (auto& k : s->slots) {
    if (k.st) cudaStreamSynchronize(k.st);
    for (double* p : {k.x, k.ny0, k.mean, k.ws}) if (p) cudaFree(p);
    if (k.st) cudaStreamDestroy(k.st);
    for (cudaEvent_t ev : {k.up_done, k.kern_done, k.dn_done}) if (ev) cudaEventDestroy(ev);
  }
  if (s->up) cudaStreamDestroy(s->up);
  if (s->dn) cudaStreamDestroy(s->dn);
  if (s->off) cudaFree(s->off);
  if (s->grid) cudaFree(s->grid);
  if (s->spl_t) cudaFree(s->spl_t);
  if (s->spl_c) cudaFree(s->spl_c);
  if (s->ll_all) cudaFree(s->ll_all);
  if (s->info_all) cudaFree(s->info_all);
  if (s->tmpl_all) cudaFree(s->tmpl_all);
  delete s;
}

}  // namespace

extern "C" {

int64_t cgp_streamer_schedule(int64_t n_obj, int64_t chunk_objects, int n_pts, int64_t* sizes, int64_t max_sizes) {
  if (n_obj < 0 || chunk_objects <= 0 || n_pts <= 0) return -1;
  const std::vector<int64_t> sz = chunk_schedule(n_obj, chunk_objects, n_pts <= 64 && chunk_objects >= 2048);
  for (size_t i = 0; i < sz.size() && (int64_t)i < max_sizes && sizes; ++i) sizes[i] = sz[i];
  return (int64_t)sz.size();
}

int cgp_streamer_create(int64_t chunk_objects, int n_pts, int64_t m_grid, int dim, int n_streams, cgp_streamer** out) {
  if (!out) return sfail(-1, "cgp_streamer_create: out is NULL");
  *out = nullptr;
  if (chunk_objects <= 0 || n_pts <= 0 || n_pts > CGP_SMALL_MAX_N || m_grid < 0 || (dim != 1 && dim != 2) || n_streams <= 0 ||
      n_streams > 64)
    return sfail(-1, "cgp_streamer_create: bad argument (objects of 1..224 points, dim 1 or 2, 1..64 streams)");
  cgp_streamer* s = new (std::nothrow) cgp_streamer;
  if (!s) return sfail(-2, "cgp_streamer_create: out of host memory");
  cudaError_t e = cudaGetDevice(&s->dev);
  s->dim = dim; s->n_pts = n_pts; s->n_streams = n_streams; s->chunk = chunk_objects; s->m_grid = m_grid;
  s->two_kernel = n_pts <= 64 && chunk_objects >= 2048;     // the split cgp_predict_batched_dev makes internally
  s->slots.resize((size_t)n_streams);
  const size_t c = (size_t)chunk_objects, n = (size_t)n_pts, m = (size_t)m_grid;
  if (e == cudaSuccess) e = dalloc(&s->off, c + 1);
  if (e == cudaSuccess) e = dalloc(&s->grid, m * dim);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s->up, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s->dn, cudaStreamNonBlocking);
  for (auto& k : s->slots) {
    if (e != cudaSuccess) break;
    e = cudaStreamCreateWithFlags(&k.st, cudaStreamNonBlocking);
    for (cudaEvent_t* ev : {&k.up_done, &k.kern_done, &k.dn_done})
      if (e == cudaSuccess) e = cudaEventCreateWithFlags(ev, cudaEventDisableTiming);
    if (e == cudaSuccess) e = dalloc(&k.x, c * n * (dim + 3));   // rows x | y | y_err | y0 at a pitch of c*n doubles (dim 1)
    if (e == cudaSuccess) { k.y = k.x + c * n * dim; k.ye = k.y + c * n; k.y0 = k.ye + c * n; }
    if (e == cudaSuccess) e = dalloc(&k.ny0, c * m + m);        // rows per object (mean function given as (B, M))
    if (e == cudaSuccess) e = dalloc(&k.mean, 2 * c * m);       // rows mean | var at a pitch of c*m doubles
    if (e == cudaSuccess) k.var = k.mean + c * m;
    if (e == cudaSuccess && s->two_kernel) e = dalloc(&k.ws, c * (size_t)cgp_factor_ws_doubles(n_pts));
  }
  if (e == cudaSuccess) {
    iota_scaled_kernel<<<(unsigned)((c + 1 + 255) / 256), 256, 0, s->slots[0].st>>>(s->off, (int64_t)c + 1, (int64_t)n);
    e = cudaStreamSynchronize(s->slots[0].st);
  }
  if (e != cudaSuccess) { release(s); return sfail(-100 - (int)e, "cgp_streamer_create", e); }
  *out = s;
  return 0;
}

void cgp_streamer_destroy(cgp_streamer* s) { release(s); }

int cgp_streamer_set_mean_spline(cgp_streamer* s, const double* t, const double* c, int n_knots) {
  if (!s) return sfail(-1, "cgp_streamer_set_mean_spline: NULL streamer");
  DeviceGuard guard(s->dev);
  if (s->spl_t) { cudaFree(s->spl_t); s->spl_t = nullptr; }
  if (s->spl_c) { cudaFree(s->spl_c); s->spl_c = nullptr; }
  s->spl_n = 0;
  if (!t || !c || n_knots == 0) return 0;                     // cleared
  if (s->dim != 1 || n_knots < 8) return sfail(-1, "cgp_streamer_set_mean_spline: cubic spline of a 1D template (>= 8 knots) expected");
  cudaError_t e = dalloc(&s->spl_t, (size_t)n_knots);
  if (e == cudaSuccess) e = dalloc(&s->spl_c, (size_t)n_knots);
  if (e == cudaSuccess) e = cudaMemcpy(s->spl_t, t, sizeof(double) * n_knots, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(s->spl_c, c, sizeof(double) * n_knots, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) return sfail(-100 - (int)e, "cgp_streamer_set_mean_spline", e);
  s->spl_n = n_knots;
  return 0;
}

int cgp_streamer_run(cgp_streamer* s, int64_t n_obj,
                     const double* x, const double* y, const double* y0, const double* y_err,
                     const double* hyp, double nugget, double floor, unsigned flags,
                     const double* xnew, const double* new_y0,
                     double* ll, double* mean, double* var, int* info, double* ll_sum,
                     int64_t* h2d_bytes, int64_t* d2h_bytes) {
  if (!s || n_obj < 0 || (n_obj && (!x || !y || !hyp || !ll || !info)) || (s->m_grid && n_obj && (!xnew || !mean)))
    return sfail(-1, "cgp_streamer_run: NULL argument");
  DeviceGuard guard(s->dev);
  const size_t n = (size_t)s->n_pts, m = (size_t)s->m_grid, dim = (size_t)s->dim;
  const bool tmpl = (flags & CGP_MEAN_TEMPLATE) && new_y0;
  const unsigned kflags = flags & ~(CGP_MEAN_TEMPLATE | CGP_GRID_UNIFORM);
  // uniformly spaced grid: verified here on the host copy, once per run (the hint costs nothing when it fails)
  const int uniform = (flags & CGP_GRID_UNIFORM) && s->dim == 1 && s->two_kernel && s->m_grid >= 2 &&
                      cgp::uniform_grid_ok(xnew, s->m_grid, hyp);
  const unsigned pflags = flags & ~CGP_GRID_UNIFORM;
  int64_t up = 0, down = 0;
  int rc = 0, bad = 0;
  cudaError_t e = cudaSuccess;
  auto h2d = [&](void* d, const void* h, size_t bytes, cudaStream_t st) {
    if (e == cudaSuccess && bytes) { e = cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, st); up += (int64_t)bytes; }
  };
  auto d2h = [&](void* h, const void* d, size_t bytes, cudaStream_t st) {
    if (e == cudaSuccess && bytes) { e = cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, st); down += (int64_t)bytes; }
  };
  // whole-batch arrays for the per-object scalars (see cgp_streamer): grown on demand
  if (n_obj > s->cap_all) {
    for (void* p : {(void*)s->ll_all, (void*)s->info_all, (void*)s->tmpl_all}) if (p) cudaFree(p);
    s->ll_all = nullptr; s->info_all = nullptr; s->tmpl_all = nullptr; s->cap_all = 0;
    e = dalloc(&s->ll_all, (size_t)n_obj);
    if (e == cudaSuccess) e = dalloc(&s->info_all, (size_t)n_obj);
    if (e == cudaSuccess) e = dalloc(&s->tmpl_all, (size_t)n_obj + m);
    if (e != cudaSuccess) return sfail(-100 - (int)e, "cgp_streamer_run (per-object arrays)", e);
    s->cap_all = n_obj;
  }
  if (m && n_obj) {
    h2d(s->grid, xnew, m * dim * 8, s->up);          // ahead of every chunk on the upload stream
    if (tmpl) h2d(s->tmpl_all, new_y0, (m + (size_t)n_obj) * 8, s->up);      // template + every object's offset: one copy per run
  }
  // the caller's per-point arrays as rows of one pinned block (x | y | y_err [| y0] at a constant pitch, dim 1)?  Then a
  // chunk of all of them is ONE two-dimensional copy; the same for the outputs (mean | var)
  const ptrdiff_t pitch_in = y - x;
  const bool in2d = dim == 1 && y_err && pitch_in >= (ptrdiff_t)(n_obj * (int64_t)n) && y_err - y == pitch_in &&
                    (!y0 || y0 - y_err == pitch_in);
  const int rows_in = y0 ? 4 : 3;
  const ptrdiff_t pitch_out = (m && var) ? var - mean : 0;
  const bool out2d = m && var && pitch_out >= (ptrdiff_t)(n_obj * (int64_t)m);
  const size_t cap = (size_t)s->chunk;
  const std::vector<int64_t> sizes = chunk_schedule(n_obj, s->chunk, s->two_kernel);
  // CGP_STREAM_TRACE=1: per-chunk timeline (upload / kernels / download, ms since the start of the run) on stderr
  static const bool trace = getenv("CGP_STREAM_TRACE") != nullptr;
  std::vector<cudaEvent_t> ev;
  if (trace) {
    ev.resize(4 * sizes.size() + 1);
    for (auto& v : ev) cudaEventCreate(&v);
    cudaEventRecord(ev.back(), s->up);
  }
  auto mark = [&](int64_t chunk, int which, cudaStream_t st) { if (trace) cudaEventRecord(ev[(size_t)(4 * chunk + which)], st); };
  int64_t chunk_index = 0, a = 0;
  cudaEvent_t last_kern = nullptr;
  for (; chunk_index < (int64_t)sizes.size() && e == cudaSuccess && rc >= 0; a += sizes[(size_t)chunk_index], ++chunk_index) {
    const size_t nb = (size_t)sizes[(size_t)chunk_index];
    cgp_streamer::Slot& k = s->slots[(size_t)(chunk_index % s->n_streams)];
    const cudaStream_t cs = k.st;                    // kernels of consecutive chunks on different streams: tails overlap
    const bool reused = chunk_index >= s->n_streams;  // the slot's previous chunk must be done with the buffers
    if (reused) e = cudaStreamWaitEvent(s->up, k.kern_done, 0);
    mark(chunk_index, 0, s->up);
    if (in2d) {
      if (e == cudaSuccess) {
        e = cudaMemcpy2DAsync(k.x, cap * n * 8, x + (size_t)a * n, (size_t)pitch_in * 8, nb * n * 8, (size_t)rows_in,
                              cudaMemcpyHostToDevice, s->up);
        up += (int64_t)(rows_in * nb * n * 8);
      }
    } else {
      h2d(k.x, x + (size_t)a * n * dim, nb * n * dim * 8, s->up);
      h2d(k.y, y + (size_t)a * n, nb * n * 8, s->up);
      if (y0) h2d(k.y0, y0 + (size_t)a * n, nb * n * 8, s->up);
      if (y_err) h2d(k.ye, y_err + (size_t)a * n, nb * n * 8, s->up);
    }
    if (m && new_y0 && !tmpl) h2d(k.ny0, new_y0 + (size_t)a * m, nb * m * 8, s->up);
    if (e != cudaSuccess) break;
    mark(chunk_index, 1, s->up);
    if ((e = cudaEventRecord(k.up_done, s->up)) != cudaSuccess) break;
    if ((e = cudaStreamWaitEvent(cs, k.up_done, 0)) != cudaSuccess) break;
    if (reused && (e = cudaStreamWaitEvent(cs, k.dn_done, 0)) != cudaSuccess) break;   // outputs still downloading
    const double* dy0 = y0 ? k.y0 : nullptr;
    const double* doffs = tmpl ? s->tmpl_all + m + a : nullptr;     // this chunk's offsets of the template mean
    if (!y0 && s->spl_n && tmpl) {                   // mean at the epochs = spline(x) + the object's offset, on the device
      rc = cgp_spline_mean_dev(s->spl_t, s->spl_c, s->spl_n, k.x, (int64_t)(nb * n), s->off, (int64_t)nb, doffs, k.y0, cs);
      if (rc < 0) break;
      dy0 = k.y0;
    }
    const double* dye = y_err ? k.ye : nullptr;
    double* dvar = var ? k.var : nullptr;
    double* dll = s->ll_all + a; int* dinfo = s->info_all + a;
    const bool fused_ll = m && s->two_kernel;        // the factor kernel emits the likelihood too: one factorisation
    if (!fused_ll) {
      rc = cgp_ll_batched_dev((int64_t)nb, s->off, s->n_pts, s->dim, k.x, k.y, dy0, dye, hyp, nugget, floor, kflags,
                              dll, dinfo, cs);
      if (rc < 0) break;
    }
    if (m) {
      if (s->two_kernel) {
        rc = cgp_factor_batched_dev((int64_t)nb, s->off, s->n_pts, s->dim, k.x, k.y, dy0, dye, hyp, nugget, floor, kflags,
                                    k.ws, dll, dinfo, cs);
        if (rc < 0) break;
        rc = cgp::predict_factored((int64_t)nb, s->off, s->n_pts, s->dim, k.x, hyp, nugget, pflags, k.ws, dinfo,
                                   s->grid, nullptr, (int64_t)m, tmpl ? s->tmpl_all : (new_y0 ? k.ny0 : nullptr), k.mean, dvar,
                                   uniform, cs, doffs);
      } else {
        const double* dny0 = nullptr;
        if (new_y0 && tmpl) {                        // the public entry point takes [template | offsets] in one array
          cudaMemcpyAsync(k.ny0, s->tmpl_all, m * 8, cudaMemcpyDeviceToDevice, cs);
          cudaMemcpyAsync(k.ny0 + m, doffs, nb * 8, cudaMemcpyDeviceToDevice, cs);
          dny0 = k.ny0;
        } else if (new_y0) dny0 = k.ny0;
        rc = cgp_predict_batched_dev((int64_t)nb, s->off, s->n_pts, s->dim, k.x, k.y, dy0, dye, hyp, nugget, floor, pflags,
                                     s->grid, nullptr, (int64_t)m, dny0, k.mean, dvar, dinfo, cs);
      }
      if (rc < 0) break;
    }
    mark(chunk_index, 2, cs);
    if ((e = cudaEventRecord(k.kern_done, cs)) != cudaSuccess) break;
    last_kern = k.kern_done;
    if ((e = cudaStreamWaitEvent(s->dn, k.kern_done, 0)) != cudaSuccess) break;
    if (m) {
      if (out2d) {
        e = cudaMemcpy2DAsync(mean + (size_t)a * m, (size_t)pitch_out * 8, k.mean, cap * m * 8, nb * m * 8, 2,
                              cudaMemcpyDeviceToHost, s->dn);
        down += (int64_t)(2 * nb * m * 8);
      } else {
        d2h(mean + (size_t)a * m, k.mean, nb * m * 8, s->dn);
        if (var) d2h(var + (size_t)a * m, k.var, nb * m * 8, s->dn);
      }
    }
    mark(chunk_index, 3, s->dn);
    if (e == cudaSuccess) e = cudaEventRecord(k.dn_done, s->dn);
  }
  // the per-object scalars of the whole batch: one copy each, behind the last chunk's kernels (the download stream has
  // waited for every chunk's kernels in order)
  if (e == cudaSuccess && rc >= 0 && n_obj) {
    if (!m && last_kern) e = cudaStreamWaitEvent(s->dn, last_kern, 0);
    d2h(ll, s->ll_all, (size_t)n_obj * 8, s->dn);
    d2h(info, s->info_all, (size_t)n_obj * 4, s->dn);
  }
  for (cudaStream_t st : {s->up, s->dn}) {        // always drain, also after an error
    cudaError_t se = cudaStreamSynchronize(st);
    if (e == cudaSuccess) e = se;
  }
  for (auto& k : s->slots) {
    cudaError_t se = cudaStreamSynchronize(k.st);
    if (e == cudaSuccess) e = se;
  }
  if (trace) {
    if (rc >= 0 && e == cudaSuccess) {
      fprintf(stderr, "chunk objects  upload[start end]  kernels[end]  download[end]  (ms)\n");
      for (size_t c = 0; c < sizes.size(); ++c) {
        float t[4] = {0, 0, 0, 0};
        for (int w = 0; w < 4; ++w) cudaEventElapsedTime(&t[w], ev.back(), ev[4 * c + w]);
        fprintf(stderr, "%5zu %7lld  %7.3f %7.3f  %7.3f  %7.3f\n", c, (long long)sizes[c], t[0], t[1], t[2], t[3]);
      }
    }
    for (auto& v : ev) cudaEventDestroy(v);
  }
  if (rc < 0) return rc;                          // message already recorded by the failing entry point
  if (e != cudaSuccess) return sfail(-100 - (int)e, "cgp_streamer_run", e);
  double total = 0.0;                             // fixed left-to-right order: reproducible
  for (int64_t i = 0; i < n_obj; ++i) { total += ll[i]; bad += info[i] != 0; }
  if (ll_sum) *ll_sum = total;
  if (h2d_bytes) *h2d_bytes = up;
  if (d2h_bytes) *d2h_bytes = down;
  return bad;
}

}  // extern "C"

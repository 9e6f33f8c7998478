// Per-object maximum-likelihood fits entirely on the device.
//
// The reference fits one object at a time: gaussian_process(y[i], x[i]).find_hyperparameters()
// in a Python loop (docs/notebook/1D_kernel_example_with_noise.ipynb cell 13;
// cosmogp/Gaussian_process.py:216-253 -> scipy.optimize.fmin).  Here every object runs scipy's
// Nelder-Mead (scipy/optimize/_optimize.py::_minimize_neldermead, non-adaptive) as a per-object
// state machine: one thread owns one simplex, the objective of all pending trial points is ONE
// batched likelihood launch (per-object hyperparameters, work list and its length read from device
// memory), and the host only enqueues launches -- it reads one counter every few iterations to
// learn when every object has converged.  The arithmetic follows numpy operation by operation
// (no FMA contraction, NaN-propagating maxima, stable sort), so each object takes exactly the
// decisions scipy would take for it: cosmogp_b200/fit.py is the host statement of the same rules.
#include "cgp_internal.h"

#include <cmath>
#include <cstring>

namespace cgp {

constexpr int NM_MAXP = 5;                 // 2D: sigma, lx, ly, lxy (+ nugget)

struct NmState {
  int64_t n_obj;
  int n, n_hyp;                            // parameters per object; the first n_hyp are kernel hyperparameters
  double* S;                               // [B][n+1][n] simplex, sorted by F between iterations
  double* F;                               // [B][n+1]
  double* fxr;                             // [B] objective at the reflected point
  int* cls;                                // [B] 0 expansion, 1 accept reflection, 2 outside, 3 inside contraction
  int* fc; int* it;                        // [B] evaluation / iteration counters
  double* hyp; double* nug;                // trial point read by the likelihood kernel: [B][n_hyp], [B]
  const double* ll; const int* info;       // likelihood kernel outputs, by object id
  double xatol, fatol; int maxiter, maxfun;
  double* x_out; double* f_out; int* it_out; int* fc_out;
};

__device__ __forceinline__ double nm_objective(const NmState& s, int64_t b) {
  const double f = -s.ll[b];
  return (s.info[b] != 0 || !isfinite(f)) ? INFINITY : f;      // not positive definite -> +inf
}
__device__ __forceinline__ void nm_put_trial(const NmState& s, int64_t b, const double* p) {
  for (int k = 0; k < s.n_hyp; ++k) s.hyp[b * s.n_hyp + k] = p[k];
  if (s.n > s.n_hyp) s.nug[b] = p[s.n_hyp];
}
// np.add.reduce(sim[:-1], 0) / N: sequential over the vertices
__device__ __forceinline__ void nm_centroid(const NmState& s, const double* S, double* xbar) {
  for (int c = 0; c < s.n; ++c) {
    double acc = S[c];
    for (int v = 1; v < s.n; ++v) acc = __dadd_rn(acc, S[v * s.n + c]);
    xbar[c] = acc / (double)s.n;
  }
}
// a*xbar + b*worst with numpy's two roundings per product (no contraction)
__device__ __forceinline__ double nm_comb(double a, double xb, double b, double w) {
  return __dadd_rn(__dmul_rn(a, xb), __dmul_rn(b, w));
}
__device__ __forceinline__ double nm_nanmax(double m, double v) { return (v > m || v != v) ? v : m; }
__device__ __forceinline__ int64_t nm_item(const int* list, int64_t k) { return list ? (int64_t)list[k] : k; }

// stable insertion sort of the simplex by objective value (np.argsort(kind="stable"); F holds no NaN)
__device__ void nm_sort(const NmState& s, int64_t b) {
  double* S = s.S + b * (s.n + 1) * s.n;
  double* F = s.F + b * (s.n + 1);
  for (int i = 1; i <= s.n; ++i) {
    const double f = F[i];
    double row[NM_MAXP];
    for (int c = 0; c < s.n; ++c) row[c] = S[i * s.n + c];
    int j = i - 1;
    while (j >= 0 && F[j] > f) {
      F[j + 1] = F[j];
      for (int c = 0; c < s.n; ++c) S[(j + 1) * s.n + c] = S[j * s.n + c];
      --j;
    }
    F[j + 1] = f;
    for (int c = 0; c < s.n; ++c) S[(j + 1) * s.n + c] = row[c];
  }
}

// initial simplex: 5 % steps, 0.00025 for zero coordinates
__global__ void nm_init_kernel(NmState s, const double* __restrict__ x0) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= s.n_obj) return;
  double* S = s.S + b * (s.n + 1) * s.n;
  for (int v = 0; v <= s.n; ++v)
    for (int c = 0; c < s.n; ++c) {
      double val = x0[b * s.n + c];
      if (v == c + 1) val = (val != 0.0) ? __dmul_rn(1.05, val) : 0.00025;
      S[v * s.n + c] = val;
    }
  s.fc[b] = 0; s.it[b] = 1;
}

__global__ void nm_set_vertex_kernel(NmState s, const int* __restrict__ list, const int* __restrict__ cnt, int v) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= (cnt ? (int64_t)*cnt : s.n_obj)) return;
  const int64_t b = nm_item(list, k);
  nm_put_trial(s, b, s.S + b * (s.n + 1) * s.n + v * s.n);
}
__global__ void nm_get_vertex_kernel(NmState s, const int* __restrict__ list, const int* __restrict__ cnt, int v) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= (cnt ? (int64_t)*cnt : s.n_obj)) return;
  const int64_t b = nm_item(list, k);
  s.F[b * (s.n + 1) + v] = nm_objective(s, b);
  s.fc[b] += 1;
}
__global__ void nm_sort_kernel(NmState s, const int* __restrict__ list, const int* __restrict__ cnt, int bump) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= (cnt ? (int64_t)*cnt : s.n_obj)) return;
  const int64_t b = nm_item(list, k);
  nm_sort(s, b);
  s.it[b] += bump;
}

// Top of an iteration: retire converged / exhausted objects, reflect the others.
__global__ void nm_reflect_kernel(NmState s, const int* __restrict__ list_in, const int* __restrict__ cnt_in,
                                  int* __restrict__ list_out, int* __restrict__ cnt_out) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= (cnt_in ? (int64_t)*cnt_in : s.n_obj)) return;
  const int64_t b = nm_item(list_in, k);
  const double* S = s.S + b * (s.n + 1) * s.n;
  const double* F = s.F + b * (s.n + 1);
  double spread = 0.0, fspread = 0.0;
  bool first = true;
  for (int v = 1; v <= s.n; ++v) {
    for (int c = 0; c < s.n; ++c) {
      const double d = fabs(__dsub_rn(S[v * s.n + c], S[c]));
      spread = first ? d : nm_nanmax(spread, d);
      first = false;
    }
    const double df = fabs(__dsub_rn(F[0], F[v]));
    fspread = (v == 1) ? df : nm_nanmax(fspread, df);
  }
  const bool keep = s.fc[b] < s.maxfun && s.it[b] < s.maxiter && !(spread <= s.xatol && fspread <= s.fatol);
  if (!keep) {
    for (int c = 0; c < s.n; ++c) s.x_out[b * s.n + c] = S[c];
    s.f_out[b] = F[0]; s.it_out[b] = s.it[b]; s.fc_out[b] = s.fc[b];
    return;
  }
  list_out[atomicAdd(cnt_out, 1)] = (int)b;
  double xbar[NM_MAXP], xr[NM_MAXP];
  nm_centroid(s, S, xbar);
  for (int c = 0; c < s.n; ++c) xr[c] = nm_comb(2.0, xbar[c], -1.0, S[s.n * s.n + c]);   // (1+rho) xbar - rho worst
  nm_put_trial(s, b, xr);
}

// After f(xr): classify, and queue the second trial point (expansion or contraction).
__global__ void nm_second_kernel(NmState s, const int* __restrict__ list, const int* __restrict__ cnt,
                                 int* __restrict__ list2, int* __restrict__ cnt2) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= (int64_t)*cnt) return;
  const int64_t b = list[k];
  const double* S = s.S + b * (s.n + 1) * s.n;
  const double* F = s.F + b * (s.n + 1);
  const double fxr = nm_objective(s, b);
  s.fxr[b] = fxr; s.fc[b] += 1;
  int cls;
  if (fxr < F[0]) cls = 0;
  else if (fxr < F[s.n - 1]) cls = 1;
  else if (fxr < F[s.n]) cls = 2;
  else cls = 3;
  s.cls[b] = cls;
  if (cls == 1) return;
  double xbar[NM_MAXP], x2[NM_MAXP];
  nm_centroid(s, S, xbar);
  const double ca = cls == 0 ? 3.0 : (cls == 2 ? 1.5 : 0.5);       // 1+rho*chi | 1+psi*rho | 1-psi
  const double cb = cls == 0 ? -2.0 : (cls == 2 ? -0.5 : 0.5);     // -rho*chi  | -psi*rho  | psi
  for (int c = 0; c < s.n; ++c) x2[c] = nm_comb(ca, xbar[c], cb, S[s.n * s.n + c]);
  list2[atomicAdd(cnt2, 1)] = (int)b;
  nm_put_trial(s, b, x2);
}

// After the second evaluation: accept a point (and sort), or shrink and queue the n re-evaluations.
__global__ void nm_update_kernel(NmState s, const int* __restrict__ list, const int* __restrict__ cnt,
                                 int* __restrict__ list_shrink, int* __restrict__ cnt_shrink) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= (int64_t)*cnt) return;
  const int64_t b = list[k];
  double* S = s.S + b * (s.n + 1) * s.n;
  double* F = s.F + b * (s.n + 1);
  const int cls = s.cls[b];
  const double fxr = s.fxr[b];
  double xbar[NM_MAXP];
  nm_centroid(s, S, xbar);
  double* worst = S + s.n * s.n;
  bool take2 = false, take_r = cls == 1, shrink = false;
  double f2 = 0.0;
  if (cls != 1) {
    f2 = nm_objective(s, b);
    s.fc[b] += 1;
    if (cls == 0) { take2 = f2 < fxr; take_r = !take2; }
    else if (cls == 2) { take2 = f2 <= fxr; shrink = !take2; }
    else { take2 = f2 < F[s.n]; shrink = !take2; }
  }
  if (take2) {                               // the trial point is still in hyp / nug, but rebuild it bit for bit
    const double ca = cls == 0 ? 3.0 : (cls == 2 ? 1.5 : 0.5);
    const double cb = cls == 0 ? -2.0 : (cls == 2 ? -0.5 : 0.5);
    for (int c = 0; c < s.n; ++c) worst[c] = nm_comb(ca, xbar[c], cb, worst[c]);
    F[s.n] = f2;
  } else if (take_r) {
    for (int c = 0; c < s.n; ++c) worst[c] = nm_comb(2.0, xbar[c], -1.0, worst[c]);
    F[s.n] = fxr;
  }
  if (shrink) {
    for (int v = 1; v <= s.n; ++v)
      for (int c = 0; c < s.n; ++c)
        S[v * s.n + c] = __dadd_rn(S[c], __dmul_rn(0.5, __dsub_rn(S[v * s.n + c], S[c])));
    list_shrink[atomicAdd(cnt_shrink, 1)] = (int)b;
    return;                                  // sorted (and counted) after its re-evaluations
  }
  nm_sort(s, b);
  s.it[b] += 1;
}

static inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

// Returns a cudaError_t as int (0 = ok).  All pointers are device pointers; x0 / x_out are (n_obj, n_par).
int fit_nelder_mead(int64_t n_obj, const int64_t* off, int max_n, int dim,
                    const double* x, const double* y, const double* y0, const double* y_err,
                    const double* x0, int n_par, double nugget, double floor, unsigned flags,
                    double xatol, double fatol, int maxiter, int maxfun,
                    double* x_out, double* f_out, int* it_out, int* fc_out, cudaStream_t st) {
  if (n_obj == 0) return 0;
  const int n_hyp = dim == 1 ? 2 : 4;
  const int n = n_par;
  const size_t B = (size_t)n_obj;
  // one stream-ordered workspace
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t at = o; o += align256(bytes); return at; };
  const size_t oS = take(B * (n + 1) * n * 8), oF = take(B * (n + 1) * 8), oFxr = take(B * 8), oHyp = take(B * n_hyp * 8),
               oNug = take(B * 8), oLl = take(B * 8), oInfo = take(B * 4), oCls = take(B * 4), oFc = take(B * 4),
               oIt = take(B * 4), oLa = take(B * 4), oLb = take(B * 4), oL2 = take(B * 4), oLs = take(B * 4),
               oCnt = take(8 * 4);
  char* ws = nullptr;
  cudaError_t ce = cudaMallocAsync((void**)&ws, o, st);
  if (ce != cudaSuccess) return (int)ce;
  int* h_cnt = nullptr;
  if ((ce = cudaMallocHost((void**)&h_cnt, sizeof(int))) != cudaSuccess) { cudaFreeAsync(ws, st); return (int)ce; }

  NmState s; memset(&s, 0, sizeof s);
  s.n_obj = n_obj; s.n = n; s.n_hyp = n_hyp;
  s.S = (double*)(ws + oS); s.F = (double*)(ws + oF); s.fxr = (double*)(ws + oFxr);
  s.hyp = (double*)(ws + oHyp); s.nug = (double*)(ws + oNug);
  double* ll = (double*)(ws + oLl); int* info = (int*)(ws + oInfo);
  s.ll = ll; s.info = info;
  s.cls = (int*)(ws + oCls); s.fc = (int*)(ws + oFc); s.it = (int*)(ws + oIt);
  int* lists[2] = {(int*)(ws + oLa), (int*)(ws + oLb)};
  int* list2 = (int*)(ws + oL2); int* list_s = (int*)(ws + oLs);
  int* cnt = (int*)(ws + oCnt);              // [0],[1]: active lists (ping-pong); [2]: second trial; [3]: shrink
  s.xatol = xatol; s.fatol = fatol; s.maxiter = maxiter; s.maxfun = maxfun;
  s.x_out = x_out; s.f_out = f_out; s.it_out = it_out; s.fc_out = fc_out;

  SmallArgs a; memset(&a, 0, sizeof a);
  a.n_obj = n_obj; a.off = off; a.x = x; a.y = y; a.y0 = y0; a.yerr = y_err; a.ll = ll; a.info = info;
  a.hyp_obj = s.hyp; a.n_hyp = n_hyp; a.nugget_obj = n > n_hyp ? s.nug : nullptr; a.nugget_shared = nugget;
  a.floor_shared = floor; a.flags = flags; a.compact_io = 0; a.cov = Cov();
  int rc = 0;
  auto likelihood = [&](const int* list, const int* count) {
    if (rc) return;
    a.order = list; a.n_obj_dev = count;
    rc = launch_small(TASK_LL, dim, max_n, a, st);
  };
  const unsigned nt = 128, nb = (unsigned)((B + nt - 1) / nt);

  nm_init_kernel<<<nb, nt, 0, st>>>(s, x0);
  for (int v = 0; v <= n; ++v) {
    nm_set_vertex_kernel<<<nb, nt, 0, st>>>(s, nullptr, nullptr, v);
    likelihood(nullptr, nullptr);
    nm_get_vertex_kernel<<<nb, nt, 0, st>>>(s, nullptr, nullptr, v);
  }
  nm_sort_kernel<<<nb, nt, 0, st>>>(s, nullptr, nullptr, 0);
  count_launch(2 * (n + 1) + 2);

  const int check_every = 4;
  for (int iter = 0; iter <= maxiter + check_every && !rc; ++iter) {
    const int cur = iter & 1;
    const int* list_in = iter ? lists[cur ^ 1] : nullptr;
    const int* cnt_in = iter ? cnt + (cur ^ 1) : nullptr;
    int* list_a = lists[cur]; int* cnt_a = cnt + cur;
    cudaMemsetAsync(cnt_a, 0, sizeof(int), st);
    cudaMemsetAsync(cnt + 2, 0, 2 * sizeof(int), st);
    nm_reflect_kernel<<<nb, nt, 0, st>>>(s, list_in, cnt_in, list_a, cnt_a);
    likelihood(list_a, cnt_a);
    nm_second_kernel<<<nb, nt, 0, st>>>(s, list_a, cnt_a, list2, cnt + 2);
    likelihood(list2, cnt + 2);
    nm_update_kernel<<<nb, nt, 0, st>>>(s, list_a, cnt_a, list_s, cnt + 3);
    for (int v = 1; v <= n; ++v) {           // shrunk simplices (rare): re-evaluate vertices 1..n
      nm_set_vertex_kernel<<<nb, nt, 0, st>>>(s, list_s, cnt + 3, v);
      likelihood(list_s, cnt + 3);
      nm_get_vertex_kernel<<<nb, nt, 0, st>>>(s, list_s, cnt + 3, v);
    }
    nm_sort_kernel<<<nb, nt, 0, st>>>(s, list_s, cnt + 3, 1);
    count_launch(4 + 2 * n);
    if (iter % check_every == check_every - 1) {
      cudaMemcpyAsync(h_cnt, cnt_a, sizeof(int), cudaMemcpyDeviceToHost, st);
      if ((ce = cudaStreamSynchronize(st)) != cudaSuccess) { rc = (int)ce; break; }
      if (*h_cnt == 0) break;
    }
  }
  if (!rc) rc = (int)cudaGetLastError();
  cudaFreeAsync(ws, st);
  cudaError_t se = cudaStreamSynchronize(st);
  cudaFreeHost(h_cnt);
  if (!rc && se != cudaSuccess) rc = (int)se;
  return rc;
}

}  // namespace cgp

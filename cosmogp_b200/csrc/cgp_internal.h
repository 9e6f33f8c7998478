// Internal declarations shared by the kernel translation units and the C ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cgp {

// Covariance function, prepared on the host from hyp / nugget / floor / flags.
//   K(a,b) = amp * exp(-q/2);  1D: q = (a-b)^2 / l^2 (cosmogp/kernel.py:71-72)
//   2D: q = d^T M d with M the inverse metric (cosmogp/kernel.py:127-130).
//   h00, h01, h11 hold -M00/2, -M01, -M11/2, so that -q/2 = dx^2 h00 + dx dy h01 + dy^2 h11.
struct Cov {
  double amp_auto;     // auto-covariance amplitude: sigma^2 (1D) or 1 (2D at HEAD, kernel.py:146-148)
  double amp_cross;    // cross-covariance amplitude: sigma^2 (kernel.py:72,141)
  double h00, h01, h11;
  double noise_const;  // floor^2 + nugget^2, added to y_err^2 on the diagonal (kernel.py:75,151)
  double nugget2;      // nugget^2: K(x*,x*) carries it but not y_err^2 (Gaussian_process.py:357)
};

enum Task { TASK_LL = 0, TASK_PREDICT = 1, TASK_LOO = 2, TASK_MATRICES = 3 };

struct SmallArgs {
  int64_t n_obj;
  const int64_t* off;      // CSR [n_obj+1]
  const int* order;        // optional processing order (object ids), may be null
  const double* x;         // dim doubles per point
  const double* y;
  const double* y0;        // may be null
  const double* yerr;      // may be null
  Cov cov;
  int* info;               // [n_obj]
  // TASK_LL
  double* ll;              // [n_obj]
  // TASK_PREDICT
  const double* xnew;
  const int64_t* goff;     // null -> shared grid of m_shared points
  int64_t m_shared;
  const double* new_y0;    // may be null
  double* mean;
  double* var;             // may be null
  int split;               // CTAs per object (each takes every split-th block of 8 grid points)
  // TASK_LOO
  int loo_mode;
  double* pred; double* pvar; double* pull; double* resid;
  // TASK_MATRICES
  const int64_t* moff;
  double* kmat; double* kinv;
};

// Launchers (cgp_small.cu).  max_n = largest object in the batch.  Return cudaError_t as int.
int launch_small(Task task, int dim, int max_n, const SmallArgs& a, cudaStream_t stream);
size_t small_smem_bytes(Task task, int dim, int nb);

// FP64 ceiling probes (cgp_small.cu)
int measure_fp64_peak(int kind, double* tflops);

void count_launch(int n = 1);

}  // namespace cgp

// Single-process multi-GPU context: the objects of a batch are sharded over the GPUs of one box
// (contiguous ranges balanced by sum N^3), every GPU keeps its shard resident and runs the same
// batched kernels as the one-GPU path; nothing is exchanged on the data path.  The only traffic
// between GPUs is the final gather of the per-object outputs on a root GPU over NVLink (NCCL
// point-to-point, ncclSend / ncclRecv inside one group) and, for callers that want it, an
// all-reduce of per-GPU scalars.  Replaces the reference's single-threaded loops over objects
// (cosmogp/Gaussian_process.py:205-213, :304-335, cosmogp/pull.py:66-94) for callers that hand over
// ONE set of host arrays and expect ONE set back -- no torchrun, no per-rank copies of the inputs.
//
// NCCL is resolved at run time (dlopen of the libnccl.so.2 the process already carries -- PyTorch
// bundles one -- or of the path given to cgp_set_nccl_library): the library itself links only cudart.
#include "../../include/cosmogp_b200.h"
#include "cgp_internal.h"

#include <dlfcn.h>

#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace cgp {
namespace {

constexpr int MAX_DEV = 16;

// ---- the handful of NCCL entry points used here (ABI of nccl.h 2.x; declared locally so that the build needs no header)
typedef struct ncclComm* ncclComm_t;
typedef int ncclResult_t;
constexpr int kNcclFloat64 = 8, kNcclSum = 0;
struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok() const { return handle && CommInitAll && CommDestroy && GroupStart && GroupEnd && Send && Recv && AllReduce; }
};
std::string g_nccl_path;
NcclApi g_nccl;

bool load_nccl(std::string* why) {
  if (g_nccl.ok()) return true;
  const char* env = getenv("CGP_NCCL_LIB");
  const char* cands[] = {g_nccl_path.empty() ? nullptr : g_nccl_path.c_str(), env, "libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* c : cands) {
    if (!c || !*c) continue;
    h = dlopen(c, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) {
    const char* err = dlerror();          // a second call would return NULL: the message is cleared by the first
    if (why) *why = std::string("libnccl.so.2 not found (") + (err ? err : "") + ")";
    return false;
  }
  g_nccl.handle = h;
#define CGP_SYM(field, name) g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(h, name))
  CGP_SYM(CommInitAll, "ncclCommInitAll"); CGP_SYM(CommDestroy, "ncclCommDestroy");
  CGP_SYM(GroupStart, "ncclGroupStart"); CGP_SYM(GroupEnd, "ncclGroupEnd");
  CGP_SYM(Send, "ncclSend"); CGP_SYM(Recv, "ncclRecv"); CGP_SYM(AllReduce, "ncclAllReduce");
  CGP_SYM(GetErrorString, "ncclGetErrorString");
#undef CGP_SYM
  if (!g_nccl.ok()) { if (why) *why = "libnccl lacks a required symbol"; return false; }
  return true;
}

struct Ctx {
  int n_dev = 0;
  int dev[MAX_DEV];
  cudaStream_t st[MAX_DEV];
  ncclComm_t comm[MAX_DEV];
  bool have_nccl = false;
};

struct Shard {                       // one device's resident slice of a batch
  int64_t o0 = 0, o1 = 0;            // objects [o0, o1)
  int64_t p0 = 0, p1 = 0;            // points  [p0, p1)
  int max_n = 0;
  int64_t* off = nullptr;            // local CSR (starts at 0)
  double *x = nullptr, *y = nullptr, *y0 = nullptr, *ye = nullptr;
  double* ll = nullptr; int* info = nullptr; double* tot = nullptr;      // per-object outputs kept on the device
  double* tot_host = nullptr;        // pinned, 2 doubles
};

struct Batch {
  Ctx* ctx = nullptr;
  int64_t n_obj = 0, n_pts = 0; int dim = 1;
  std::vector<int64_t> starts;       // n_dev + 1 object boundaries
  Shard sh[MAX_DEV];
};

// run f(d) for every device of the context, one host thread each (uploads / downloads from pageable
// host memory block their caller: one thread per link keeps all links busy); collects the first error.
template <class F>
int for_each_device(Ctx* c, F f, const char* who) {
  std::vector<int> rc((size_t)c->n_dev, 0);
  std::vector<std::string> msg((size_t)c->n_dev);
  auto body = [&](int d) {
    if (cudaSetDevice(c->dev[d]) != cudaSuccess) { rc[d] = -100; msg[d] = "cudaSetDevice failed"; return; }
    rc[d] = f(d);
    if (rc[d] < 0) msg[d] = cgp_last_error();
  };
  if (c->n_dev == 1) body(0);
  else {
    std::vector<std::thread> th;
    for (int d = 0; d < c->n_dev; ++d) th.emplace_back(body, d);
    for (auto& t : th) t.join();
  }
  int bad = 0;
  for (int d = 0; d < c->n_dev; ++d) {
    if (rc[d] < 0) return fail(rc[d], "%s (device %d): %s", who, c->dev[d], msg[d].c_str());
    bad += rc[d];
  }
  return bad;
}

int cu(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return 0;
  return fail(-100 - (int)e, "%s: CUDA error %d (%s)", what, (int)e, cudaGetErrorString(e));
}
int nc(ncclResult_t r, const char* what) {
  if (r == 0) return 0;
  return fail(-300 - r, "%s: NCCL error %d (%s)", what, r, g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
}

template <class T>
int up(T** dst, const T* src, size_t n, cudaStream_t st) {
  *dst = nullptr;
  if (!src) return 0;
  int rc = cu(cudaMallocAsync((void**)dst, (n ? n : 1) * sizeof(T), st), "upload (allocation)");
  if (rc) return rc;
  return cu(cudaMemcpyAsync(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice, st), "upload");
}

// gather per-device slices (device pointers, `counts` doubles each) on the root device: recv is a root-device
// buffer of sum(counts) doubles.  One NCCL group of point-to-point transfers over NVLink; the root's own
// slice is a device-to-device copy.
int gather_on_root(Ctx* c, double* const* send, const int64_t* counts, double* recv, int root) {
  if (c->n_dev == 1) {
    if (send[0] != recv && counts[0])
      return cu(cudaMemcpyAsync(recv, send[0], counts[0] * sizeof(double), cudaMemcpyDeviceToDevice, c->st[0]), "gather (copy)");
    return 0;
  }
  if (!c->have_nccl) return fail(-3, "gather over NVLink needs NCCL, which could not be loaded for this context");
  int rc = nc(g_nccl.GroupStart(), "ncclGroupStart");
  int64_t o = 0;
  for (int d = 0; d < c->n_dev && !rc; ++d) {
    if (counts[d]) {
      if (d == root) {
        cudaSetDevice(c->dev[root]);
        rc = cu(cudaMemcpyAsync(recv + o, send[d], counts[d] * sizeof(double), cudaMemcpyDeviceToDevice, c->st[root]), "gather (root slice)");
      } else {
        rc = nc(g_nccl.Send(send[d], (size_t)counts[d], kNcclFloat64, root, c->comm[d], c->st[d]), "ncclSend");
        if (!rc) rc = nc(g_nccl.Recv(recv + o, (size_t)counts[d], kNcclFloat64, d, c->comm[root], c->st[root]), "ncclRecv");
      }
    }
    o += counts[d];
  }
  int rc2 = nc(g_nccl.GroupEnd(), "ncclGroupEnd");
  return rc ? rc : rc2;
}

}  // namespace
}  // namespace cgp

using namespace cgp;

extern "C" {

int cgp_set_nccl_library(const char* path) {
  g_nccl_path = path ? path : "";
  return 0;
}

// contiguous object ranges with near-equal sum of N^3 (the factorisation cost): starts[n_parts + 1]
int cgp_shard_ranges(int64_t n_obj, const int64_t* off, int n_parts, int64_t* starts) {
  if (n_obj < 0 || n_parts < 1 || !starts || (n_obj && !off)) return fail(-1, "cgp_shard_ranges: bad arguments");
  double total = 0.0;
  for (int64_t i = 0; i < n_obj; ++i) { const double n = (double)(off[i + 1] - off[i]); total += n * n * n + 1.0; }
  starts[0] = 0;
  double acc = 0.0; int part = 1;
  for (int64_t i = 0; i < n_obj && part < n_parts; ++i) {
    const double n = (double)(off[i + 1] - off[i]);
    acc += n * n * n + 1.0;
    while (part < n_parts && acc >= total * part / n_parts) starts[part++] = i + 1;
  }
  for (; part < n_parts; ++part) starts[part] = n_obj;
  starts[n_parts] = n_obj;
  return 0;
}

int cgp_ctx_create(int n_dev, const int* dev_ids, void** out) {
  if (!out) return fail(-1, "cgp_ctx_create: out is NULL");
  *out = nullptr;
  int avail = 0;
  int rc = cu(cudaGetDeviceCount(&avail), "cgp_ctx_create");
  if (rc) return rc;
  if (n_dev <= 0) n_dev = avail;
  if (n_dev < 1 || n_dev > MAX_DEV || n_dev > avail) return fail(-1, "cgp_ctx_create: %d devices requested, %d visible", n_dev, avail);
  Ctx* c = new Ctx();
  c->n_dev = n_dev;
  int prev = 0; cudaGetDevice(&prev);
  for (int d = 0; d < n_dev; ++d) {
    c->dev[d] = dev_ids ? dev_ids[d] : d;
    c->st[d] = nullptr; c->comm[d] = nullptr;
    if (c->dev[d] < 0 || c->dev[d] >= avail) { delete c; return fail(-1, "cgp_ctx_create: device id %d out of range", c->dev[d]); }
  }
  for (int d = 0; d < n_dev && !rc; ++d) {
    rc = cu(cudaSetDevice(c->dev[d]), "cgp_ctx_create (set device)");
    if (!rc) rc = cu(cudaStreamCreateWithFlags(&c->st[d], cudaStreamNonBlocking), "cgp_ctx_create (stream)");
    cudaMemPool_t pool;
    if (!rc && cudaDeviceGetDefaultMemPool(&pool, c->dev[d]) == cudaSuccess) {
      unsigned long long keep = ~0ULL;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
  }
  if (!rc && n_dev > 1) {
    std::string why;
    if (load_nccl(&why)) {
      rc = nc(g_nccl.CommInitAll(c->comm, n_dev, c->dev), "ncclCommInitAll");
      c->have_nccl = rc == 0;
    }
    // without NCCL the context still shards and downloads per device; cgp_ctx_gather_f64 / gather=1 then fail loudly
  }
  cudaSetDevice(prev);
  if (rc) { cgp_ctx_destroy(c); return rc; }
  *out = c;
  return 0;
}

void cgp_ctx_destroy(void* ctx) {
  Ctx* c = (Ctx*)ctx;
  if (!c) return;
  int prev = 0; cudaGetDevice(&prev);
  for (int d = 0; d < c->n_dev; ++d) {
    if (c->have_nccl && c->comm[d]) g_nccl.CommDestroy(c->comm[d]);
    if (c->st[d]) { cudaSetDevice(c->dev[d]); cudaStreamSynchronize(c->st[d]); cudaStreamDestroy(c->st[d]); }
  }
  cudaSetDevice(prev);
  delete c;
}

int cgp_ctx_info(void* ctx, int* n_dev, int* have_nccl) {
  Ctx* c = (Ctx*)ctx;
  if (!c) return fail(-1, "cgp_ctx_info: ctx is NULL");
  if (n_dev) *n_dev = c->n_dev;
  if (have_nccl) *have_nccl = c->have_nccl ? 1 : 0;
  return 0;
}

// ---- thin NCCL wrappers on per-device buffers
int cgp_ctx_gather_f64(void* ctx, double* const* send_dev, const int64_t* counts, double* recv_root_dev, int root) {
  Ctx* c = (Ctx*)ctx;
  if (!c || !send_dev || !counts || !recv_root_dev || root < 0 || root >= c->n_dev) return fail(-1, "cgp_ctx_gather_f64: bad arguments");
  int prev = 0; cudaGetDevice(&prev);
  int rc = gather_on_root(c, send_dev, counts, recv_root_dev, root);
  for (int d = 0; d < c->n_dev && !rc; ++d) { cudaSetDevice(c->dev[d]); rc = cu(cudaStreamSynchronize(c->st[d]), "cgp_ctx_gather_f64"); }
  cudaSetDevice(prev);
  return rc;
}

int cgp_ctx_allreduce_sum_f64(void* ctx, double* const* buf_dev, int64_t count) {
  Ctx* c = (Ctx*)ctx;
  if (!c || !buf_dev || count < 0) return fail(-1, "cgp_ctx_allreduce_sum_f64: bad arguments");
  if (c->n_dev == 1 || count == 0) return 0;
  if (!c->have_nccl) return fail(-3, "cgp_ctx_allreduce_sum_f64 needs NCCL, which could not be loaded for this context");
  int prev = 0; cudaGetDevice(&prev);
  int rc = nc(g_nccl.GroupStart(), "ncclGroupStart");
  for (int d = 0; d < c->n_dev && !rc; ++d)
    rc = nc(g_nccl.AllReduce(buf_dev[d], buf_dev[d], (size_t)count, kNcclFloat64, kNcclSum, c->comm[d], c->st[d]), "ncclAllReduce");
  int rc2 = nc(g_nccl.GroupEnd(), "ncclGroupEnd");
  if (!rc) rc = rc2;
  for (int d = 0; d < c->n_dev && !rc; ++d) { cudaSetDevice(c->dev[d]); rc = cu(cudaStreamSynchronize(c->st[d]), "cgp_ctx_allreduce_sum_f64"); }
  cudaSetDevice(prev);
  return rc;
}

// ---- a batch sharded over the devices of the context, resident until destroyed
void cgp_ctx_batch_destroy(void* batch) {
  Batch* b = (Batch*)batch;
  if (!b) return;
  int prev = 0; cudaGetDevice(&prev);
  for (int d = 0; d < b->ctx->n_dev; ++d) {
    Shard& s = b->sh[d];
    cudaSetDevice(b->ctx->dev[d]);
    cudaStream_t st = b->ctx->st[d];
    void* ptrs[] = {s.off, s.x, s.y, s.y0, s.ye, s.ll, s.info, s.tot};
    for (void* p : ptrs) if (p) cudaFreeAsync(p, st);
    if (s.tot_host) cudaFreeHost(s.tot_host);
    cudaStreamSynchronize(st);
  }
  cudaSetDevice(prev);
  delete b;
}

int cgp_ctx_batch_create(void* ctx, int64_t n_obj, const int64_t* off, int dim,
                         const double* x, const double* y, const double* y0, const double* y_err, void** out) {
  Ctx* c = (Ctx*)ctx;
  if (!c || !out || n_obj < 0 || (n_obj && (!off || !x || !y)) || (dim != 1 && dim != 2))
    return fail(-1, "cgp_ctx_batch_create: bad arguments");
  *out = nullptr;
  Batch* b = new Batch();
  b->ctx = c; b->n_obj = n_obj; b->dim = dim; b->n_pts = n_obj ? off[n_obj] : 0;
  b->starts.resize((size_t)c->n_dev + 1);
  int rc = cgp_shard_ranges(n_obj, off, c->n_dev, b->starts.data());
  if (rc) { delete b; return rc; }
  int prev = 0; cudaGetDevice(&prev);
  rc = for_each_device(c, [&](int d) -> int {
    Shard& s = b->sh[d];
    cudaStream_t st = c->st[d];
    s.o0 = b->starts[d]; s.o1 = b->starts[d + 1];
    const int64_t nb = s.o1 - s.o0;
    s.p0 = n_obj ? off[s.o0] : 0; s.p1 = n_obj ? off[s.o1] : 0;
    const size_t np = (size_t)(s.p1 - s.p0);
    std::vector<int64_t> loc((size_t)nb + 1);
    for (int64_t i = 0; i <= nb; ++i) loc[(size_t)i] = off[s.o0 + i] - s.p0;
    for (int64_t i = 0; i < nb; ++i) { const int n = (int)(loc[(size_t)i + 1] - loc[(size_t)i]); if (n > s.max_n) s.max_n = n; }
    int e = up(&s.off, loc.data(), (size_t)nb + 1, st);
    if (!e) e = up(&s.x, x + s.p0 * dim, np * dim, st);
    if (!e) e = up(&s.y, y + s.p0, np, st);
    if (!e && y0) e = up(&s.y0, y0 + s.p0, np, st);
    if (!e && y_err) e = up(&s.ye, y_err + s.p0, np, st);
    if (!e) e = cu(cudaMallocAsync((void**)&s.ll, (nb ? nb : 1) * sizeof(double), st), "batch (ll)");
    if (!e) e = cu(cudaMallocAsync((void**)&s.info, (nb ? nb : 1) * sizeof(int), st), "batch (info)");
    if (!e) e = cu(cudaMallocAsync((void**)&s.tot, 2 * sizeof(double), st), "batch (totals)");
    if (!e) e = cu(cudaMallocHost((void**)&s.tot_host, 2 * sizeof(double)), "batch (pinned totals)");
    if (!e) e = cu(cudaStreamSynchronize(st), "cgp_ctx_batch_create");          // `loc` must outlive the copy
    return e;
  }, "cgp_ctx_batch_create");
  cudaSetDevice(prev);
  if (rc < 0) { cgp_ctx_batch_destroy(b); return rc; }
  *out = b;
  return 0;
}

int cgp_ctx_batch_ranges(void* batch, int64_t* starts) {
  Batch* b = (Batch*)batch;
  if (!b || !starts) return fail(-1, "cgp_ctx_batch_ranges: bad arguments");
  for (size_t i = 0; i < b->starts.size(); ++i) starts[i] = b->starts[i];
  return 0;
}

// One likelihood evaluation of the whole batch: every device evaluates and reduces its shard; 16 bytes
// per device come back and are added in device order (deterministic).  ll_obj / info (host, nullable)
// receive the per-object values.  Launches on all devices are issued from the calling thread (they are
// asynchronous); only the bulk downloads, if asked for, use one thread per device.
int cgp_ctx_batch_ll(void* batch, const double* hyp, double nugget, double floor, unsigned flags,
                     double* ll_sum, double* ll_obj, int* info) {
  Batch* b = (Batch*)batch;
  if (!b || !hyp) return fail(-1, "cgp_ctx_batch_ll: bad arguments");
  Ctx* c = b->ctx;
  int prev = 0; cudaGetDevice(&prev);
  int rc = 0;
  for (int d = 0; d < c->n_dev && !rc; ++d) {
    Shard& s = b->sh[d];
    cudaSetDevice(c->dev[d]);
    s.tot_host[0] = 0.0; s.tot_host[1] = 0.0;
    if (s.o1 == s.o0) continue;
    rc = cgp_ll_total_dev(s.o1 - s.o0, s.off, s.max_n, b->dim, s.x, s.y, s.y0, s.ye, hyp, nugget, floor, flags,
                          s.ll, s.info, s.tot, nullptr, c->st[d]);
    if (!rc) rc = cu(cudaMemcpyAsync(s.tot_host, s.tot, 2 * sizeof(double), cudaMemcpyDeviceToHost, c->st[d]), "cgp_ctx_batch_ll (read back)");
  }
  double total = 0.0, bad = 0.0;
  for (int d = 0; d < c->n_dev && !rc; ++d) {
    cudaSetDevice(c->dev[d]);
    rc = cu(cudaStreamSynchronize(c->st[d]), "cgp_ctx_batch_ll");
    total += b->sh[d].tot_host[0]; bad += b->sh[d].tot_host[1];
  }
  if (!rc && (ll_obj || info)) {
    rc = for_each_device(c, [&](int d) -> int {
      Shard& s = b->sh[d];
      const size_t nb = (size_t)(s.o1 - s.o0);
      int e = 0;
      if (ll_obj && nb) e = cu(cudaMemcpyAsync(ll_obj + s.o0, s.ll, nb * sizeof(double), cudaMemcpyDeviceToHost, c->st[d]), "ll download");
      if (!e && info && nb) e = cu(cudaMemcpyAsync(info + s.o0, s.info, nb * sizeof(int), cudaMemcpyDeviceToHost, c->st[d]), "info download");
      if (!e) e = cu(cudaStreamSynchronize(c->st[d]), "cgp_ctx_batch_ll (download)");
      return e;
    }, "cgp_ctx_batch_ll");
    if (rc > 0) rc = 0;
  }
  cudaSetDevice(prev);
  if (rc < 0) return rc;
  if (ll_sum) *ll_sum = total;
  return bad > 2147483647.0 ? 2147483647 : (int)bad;
}

// shared working code of predict / step and pulls: per device run `launch` into device outputs of `width` doubles
// per unit (units = objects x m for predictions, points for pulls), then bring `n_out` arrays to the host --
// gather = 0: each device copies its slice straight into the host arrays over its own PCIe link;
// gather = 1: the slices are first gathered on device 0 over NVLink (NCCL send / recv), then leave through one link.
static int collect(Batch* b, int n_arr, double* const host[], double* dev_out[][8], const int64_t unit0[], const int64_t units[],
                   int* info_host, int gather) {
  Ctx* c = b->ctx;
  int rc = 0;
  if (gather && c->n_dev > 1) {
    int64_t total = 0;
    for (int d = 0; d < c->n_dev; ++d) total += units[d];
    cudaSetDevice(c->dev[0]);
    for (int a = 0; a < n_arr && !rc; ++a) {
      if (!host[a]) continue;
      double* root = nullptr;
      rc = cu(cudaMallocAsync((void**)&root, (total ? total : 1) * sizeof(double), c->st[0]), "gather (root buffer)");
      if (rc) break;
      // the root buffer must exist before the peers' sends are matched: allocation is stream ordered on st[0], as is the recv
      double* send[MAX_DEV]; int64_t cnt[MAX_DEV];
      for (int d = 0; d < c->n_dev; ++d) { send[d] = dev_out[d][a]; cnt[d] = units[d]; }
      rc = gather_on_root(c, send, cnt, root, 0);
      cudaSetDevice(c->dev[0]);
      if (!rc) rc = cu(cudaMemcpyAsync(host[a], root, total * sizeof(double), cudaMemcpyDeviceToHost, c->st[0]), "gather (download)");
      cudaFreeAsync(root, c->st[0]);
    }
    for (int d = 0; d < c->n_dev && !rc; ++d) { cudaSetDevice(c->dev[d]); rc = cu(cudaStreamSynchronize(c->st[d]), "gather"); }
    if (rc) return rc;
  }
  rc = for_each_device(c, [&](int d) -> int {
    Shard& s = b->sh[d];
    int e = 0;
    if (!(gather && c->n_dev > 1))
      for (int a = 0; a < n_arr && !e; ++a)
        if (host[a] && units[d])
          e = cu(cudaMemcpyAsync(host[a] + unit0[d], dev_out[d][a], units[d] * sizeof(double), cudaMemcpyDeviceToHost, c->st[d]), "download");
    if (!e && info_host && s.o1 > s.o0)
      e = cu(cudaMemcpyAsync(info_host + s.o0, s.info, (size_t)(s.o1 - s.o0) * sizeof(int), cudaMemcpyDeviceToHost, c->st[d]), "info download");
    if (!e) e = cu(cudaStreamSynchronize(c->st[d]), "download");
    return e;
  }, "collect");
  return rc;
}

// Prediction (and, with ll_obj, the likelihood of the same factorisation: the "step") of the whole batch on a
// shared grid.  new_y0: NULL, or with CGP_MEAN_TEMPLATE the packed [template (m) | offsets (n_obj)], or (n_obj, m).
int cgp_ctx_batch_predict(void* batch, const double* hyp, double nugget, double floor, unsigned flags,
                          const double* xnew, int64_t m, const double* new_y0,
                          double* ll_obj, double* mean, double* var, int* info, int gather) {
  Batch* b = (Batch*)batch;
  if (!b || !hyp || !xnew || !mean || m < 1) return fail(-1, "cgp_ctx_batch_predict: bad arguments");
  Ctx* c = b->ctx;
  int prev = 0; cudaGetDevice(&prev);
  double* outs[MAX_DEV][8]; memset(outs, 0, sizeof outs);
  double* grids[MAX_DEV] = {nullptr}; double* ny0s[MAX_DEV] = {nullptr};
  int64_t unit0[MAX_DEV], units[MAX_DEV];
  const bool tmpl = (flags & CGP_MEAN_TEMPLATE) && new_y0;
  int rc = for_each_device(c, [&](int d) -> int {
    Shard& s = b->sh[d];
    cudaStream_t st = c->st[d];
    const int64_t nb = s.o1 - s.o0;
    unit0[d] = s.o0 * m; units[d] = nb * m;
    if (!nb) return 0;
    int e = up(&grids[d], xnew, (size_t)m * b->dim, st);
    std::vector<double> packed;
    if (!e && new_y0) {
      if (tmpl) {                                          // template + this shard's offsets
        packed.assign(new_y0, new_y0 + m);
        packed.insert(packed.end(), new_y0 + m + s.o0, new_y0 + m + s.o1);
        e = up(&ny0s[d], packed.data(), packed.size(), st);
      } else e = up(&ny0s[d], new_y0 + s.o0 * m, (size_t)(nb * m), st);
    }
    if (!e) e = cu(cudaMallocAsync((void**)&outs[d][0], nb * m * sizeof(double), st), "predict (mean)");
    if (!e && var) e = cu(cudaMallocAsync((void**)&outs[d][1], nb * m * sizeof(double), st), "predict (var)");
    if (e) return e;
    if (ll_obj) e = cgp_step_batched_dev(nb, s.off, s.max_n, b->dim, s.x, s.y, s.y0, s.ye, hyp, nugget, floor, flags,
                                         grids[d], nullptr, m, ny0s[d], s.ll, outs[d][0], outs[d][1], s.info, st);
    else e = cgp_predict_batched_dev(nb, s.off, s.max_n, b->dim, s.x, s.y, s.y0, s.ye, hyp, nugget, floor, flags,
                                     grids[d], nullptr, m, ny0s[d], outs[d][0], outs[d][1], s.info, st);
    if (e) return e;
    if (ll_obj) e = cu(cudaMemcpyAsync(ll_obj + s.o0, s.ll, nb * sizeof(double), cudaMemcpyDeviceToHost, st), "ll download");
    if (!e) e = cu(cudaStreamSynchronize(st), "cgp_ctx_batch_predict");     // `packed` must outlive its copy; kernels done
    return e;
  }, "cgp_ctx_batch_predict");
  if (rc >= 0) {
    double* host[2] = {mean, var};
    rc = collect(b, 2, host, outs, unit0, units, info, gather);
  }
  for (int d = 0; d < c->n_dev; ++d) {
    cudaSetDevice(c->dev[d]);
    void* ptrs[] = {grids[d], ny0s[d], outs[d][0], outs[d][1]};
    for (void* p : ptrs) if (p) cudaFreeAsync(p, c->st[d]);
  }
  cudaSetDevice(prev);
  if (rc < 0) return rc;
  int bad = 0;
  if (info) for (int64_t i = 0; i < b->n_obj; ++i) bad += info[i] != 0;
  return bad;
}

// Closed-form leave-one-out pulls of the whole batch (the batch's y0 is the template mean `m` of cgp_loo_batched_dev).
// moments (nullable, 2 doubles): sum of the pulls and sum of their squares over all devices (norm.fit, pull.py:102),
// reduced per device, then either added on the host in device order (gather = 0) or all-reduced over NVLink (gather = 1).
int cgp_ctx_batch_loo(void* batch, const double* hyp, double nugget, double floor, unsigned flags, int mode,
                      double* pred, double* pred_var, double* pull, double* resid, int* info, double* moments, int gather) {
  Batch* b = (Batch*)batch;
  if (!b || !hyp) return fail(-1, "cgp_ctx_batch_loo: bad arguments");
  Ctx* c = b->ctx;
  int prev = 0; cudaGetDevice(&prev);
  double* outs[MAX_DEV][8]; memset(outs, 0, sizeof outs);
  double* mom[MAX_DEV] = {nullptr};
  int64_t unit0[MAX_DEV], units[MAX_DEV];
  double* host[4] = {pred, pred_var, pull, resid};
  std::vector<double> mom_host((size_t)2 * c->n_dev, 0.0);
  int rc = for_each_device(c, [&](int d) -> int {
    Shard& s = b->sh[d];
    cudaStream_t st = c->st[d];
    const int64_t nb = s.o1 - s.o0, np = s.p1 - s.p0;
    unit0[d] = s.p0; units[d] = np;
    int e = cu(cudaMallocAsync((void**)&mom[d], 2 * sizeof(double), st), "loo (moments)");
    if (!e) e = cu(cudaMemsetAsync(mom[d], 0, 2 * sizeof(double), st), "loo (moments)");
    if (e || !nb) return e;
    for (int a = 0; a < 4 && !e; ++a)
      if (host[a] || (a == 2 && moments)) e = cu(cudaMallocAsync((void**)&outs[d][a], (np ? np : 1) * sizeof(double), st), "loo (outputs)");
    if (e) return e;
    e = cgp_loo_batched_dev(nb, s.off, s.max_n, b->dim, s.x, s.y, s.y0, s.ye, hyp, nugget, floor, flags, mode,
                            outs[d][0], outs[d][1], outs[d][2], outs[d][3], s.info, st);
    if (!e && moments) e = cgp_moments_dev(outs[d][2], np, 0.0, mom[d], st);
    if (!e && moments && !(gather && c->n_dev > 1))
      e = cu(cudaMemcpyAsync(&mom_host[2 * (size_t)d], mom[d], 2 * sizeof(double), cudaMemcpyDeviceToHost, st), "loo (moments download)");
    if (!e) e = cu(cudaStreamSynchronize(st), "cgp_ctx_batch_loo");
    return e;
  }, "cgp_ctx_batch_loo");
  if (rc >= 0 && moments) {
    if (gather && c->n_dev > 1) {
      rc = cgp_ctx_allreduce_sum_f64(c, mom, 2);
      cudaSetDevice(c->dev[0]);
      if (!rc) rc = cu(cudaMemcpy(moments, mom[0], 2 * sizeof(double), cudaMemcpyDeviceToHost), "loo (moments)");
    } else {
      moments[0] = moments[1] = 0.0;
      for (int d = 0; d < c->n_dev; ++d) { moments[0] += mom_host[2 * (size_t)d]; moments[1] += mom_host[2 * (size_t)d + 1]; }
    }
  }
  if (rc >= 0) rc = collect(b, 4, host, outs, unit0, units, info, gather);
  for (int d = 0; d < c->n_dev; ++d) {
    cudaSetDevice(c->dev[d]);
    void* ptrs[] = {outs[d][0], outs[d][1], outs[d][2], outs[d][3], mom[d]};
    for (void* p : ptrs) if (p) cudaFreeAsync(p, c->st[d]);
  }
  cudaSetDevice(prev);
  if (rc < 0) return rc;
  int bad = 0;
  if (info) for (int64_t i = 0; i < b->n_obj; ++i) bad += info[i] != 0;
  return bad;
}

}  // extern "C"

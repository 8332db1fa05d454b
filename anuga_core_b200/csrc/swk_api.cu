// swk_api.cu - host side of libswk.so: the C ABI declared in include/swk.h.
//
// Build (see __graft_entry__.build / anuga_core_b200/build.py):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false \
//        -Xcompiler -fPIC,-ffp-contract=off -shared -o libswk.so swk_api.cu
//
// No CPU fallback: every entry point needs a CUDA device that can run sm_100a code.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/swk.h"
#include "swk_kernels.cuh"

using namespace swk;

// ----------------------------------------------------------------------------
// error handling
// ----------------------------------------------------------------------------
static thread_local std::string g_err;

static int fail(int code, const std::string &msg)
{
  g_err = msg;
  return code;
}

#define CK(call)                                                                        \
  do {                                                                                  \
    cudaError_t e_ = (call);                                                            \
    if (e_ != cudaSuccess)                                                              \
      return fail(SWK_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));    \
  } while (0)

#define CKV(call)                                                                       \
  do {                                                                                  \
    int r_ = (call);                                                                    \
    if (r_ != SWK_OK) return r_;                                                        \
  } while (0)

// value table layout (doubles): boundary values [SEG_MAX][3 substeps][3], then [OP_MAX]{rate, factor}
constexpr int SEG_MAX = 4097, OP_MAX = 1024;
constexpr size_t VALS_OPS = (size_t)SEG_MAX * 9, VALS_TOTAL = VALS_OPS + 2 * (size_t)OP_MAX;
static inline int nblk(long long n) { return (int)((n + BLOCK - 1) / BLOCK); }
static inline double u2d_host(unsigned long long u) { double x; memcpy(&x, &u, 8); return x; }

// ----------------------------------------------------------------------------
// NCCL, loaded at run time so that single-GPU use has no NCCL dependency
// ----------------------------------------------------------------------------
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclInt64 = 4, ncclUint64 = 5, ncclFloat64 = 8 };
enum { ncclSumOp = 0, ncclMaxOp = 2, ncclMinOp = 3 };
struct NcclApi {
  void *lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;

static int load_nccl()
{
  if (g_nccl.lib) return SWK_OK;
  const char *env = getenv("SWK_NCCL_LIB");
  const char *names[] = {env, "libnccl.so.2", "libnccl.so"};
  void *lib = nullptr;
  for (const char *n : names) {
    if (!n) continue;
    lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (lib) break;
  }
  if (!lib) return fail(SWK_ERR_NCCL, "cannot dlopen libnccl.so.2 (set SWK_NCCL_LIB)");
#define SYM(field, name)                                                        \
  *(void **)(&g_nccl.field) = dlsym(lib, name);                                 \
  if (!g_nccl.field) return fail(SWK_ERR_NCCL, std::string("missing symbol ") + name);
  SYM(GetUniqueId, "ncclGetUniqueId");
  SYM(CommInitRank, "ncclCommInitRank");
  SYM(CommDestroy, "ncclCommDestroy");
  SYM(AllReduce, "ncclAllReduce");
  SYM(Send, "ncclSend");
  SYM(Recv, "ncclRecv");
  SYM(GroupStart, "ncclGroupStart");
  SYM(GroupEnd, "ncclGroupEnd");
  SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
  g_nccl.lib = lib;
  return SWK_OK;
}

#define NK(call)                                                                        \
  do {                                                                                  \
    ncclResult_t r_ = (call);                                                           \
    if (r_ != 0)                                                                        \
      return fail(SWK_ERR_NCCL, std::string(#call) + ": " + g_nccl.GetErrorString(r_)); \
  } while (0)

// ----------------------------------------------------------------------------
// the handle
// ----------------------------------------------------------------------------
struct RateOp {
  double rate, factor;
  bool dynamic = false;        // rate / factor change every step: read from the device value table, never fused
  double *d_rate_array = nullptr;
  int *d_indices = nullptr;
  int n = 0;
  int all_nonneg = 1;
  double *d_partial = nullptr;
  int nblocks = 0;
};

// A registered set of triangles that the host layer gathers / scatters every step (the rows of an inlet):
// device ids and a page-locked staging buffer live with the handle, so a transfer is one kernel + one copy.
struct CellSet {
  int n = 0;
  int *d_ids = nullptr;
  double *d_buf = nullptr;       // 4*n doubles
  double *h_buf = nullptr;       // page-locked, 4*n doubles
  cudaEvent_t ev = nullptr;      // last H2D copy out of h_buf
  bool inflight = false;
};

struct Peer {
  int rank;
  int n_send = 0, n_recv = 0;
  int *d_send_ids = nullptr, *d_recv_ids = nullptr;
  double *d_send_buf = nullptr, *d_recv_buf = nullptr;
};

// a process-level NCCL communicator (one per GPU process): shared by the device time loop of
// every domain attached to it and by the host-level collectives of the Python layer
struct swk_comm {
  ncclComm_t nc = nullptr;
  int rank = 0, nranks = 1, device = 0;
  cudaStream_t stream = nullptr;      // host-level collectives (swk_comm_allreduce)
  void *d_buf = nullptr;
  size_t cap = 0;
};

struct swk_domain {
  int device = 0;
  cudaStream_t stream = nullptr;
  int64_t N = 0, M = 0, NP = 0;
  swk_params P;
  Consts K;
  TimeParams TP;
  Dev D;
  bool has_riverwalls = false;
  bool per_call = false;

  std::vector<int> new2old, old2new;
  int *d_new2old = nullptr;
  int *d_ident_b = nullptr;   // identity for boundary arrays

  // owned device memory
  d4 *cq = nullptr, *eq = nullptr, *xg = nullptr, *fg = nullptr, *bq = nullptr;
  i4 *connA = nullptr, *connB = nullptr;
  double *eu = nullptr, *bk = nullptr, *eta = nullptr, *max_speed = nullptr, *vcoord = nullptr, *wind = nullptr;
  double *xbed = nullptr;      // per-call layer: [3][NP] edge beds + [NP] centroid heights as the caller gave them
  unsigned char *zflag = nullptr;
  int *rw_counter = nullptr, *rw_rowIndex = nullptr, *rw_list = nullptr;
  int n_rw_list = 0;           // triangles with at least one riverwall edge (device ids, ascending)
  std::vector<int> h_rw_list;
  double *rw_elevation = nullptr, *rw_hydraulic = nullptr;
  Clock *d_clock = nullptr, *h_clock = nullptr;
  double *staging = nullptr;     // 3*N doubles
  size_t staging_n = 0;
  double *acct_val = nullptr;
  int *pos_b = nullptr, *acct_keys = nullptr, *acct_keys_pos = nullptr;
  int n_acct = 0, n_acct_keys = 0;
  double full_area = 0.0;

  // boundary
  int *b_cell = nullptr, *b_edge = nullptr, *b_seg = nullptr;
  std::vector<int> h_b_seg;
  std::vector<int> seg_kind;
  int *d_seg_kind = nullptr;
  int seg_cap = 0;
  // time-space tables of the table kinds: one device buffer per segment
  std::vector<double *> seg_tab;      // device pointers (null: none)
  std::vector<int> seg_np, seg_nframes;
  std::vector<int> h_b_point;         // [M]
  double **d_seg_tab = nullptr;
  int *d_seg_np = nullptr, *d_b_point = nullptr;
  bool seg_dirty = true;       // kinds / edge->segment map changed (structure; uploaded with a sync)
  // value table: boundary values [segment][substep][3], then {rate, factor} of every Rate_operator.
  // Page-locked host copy + device copy, refreshed by one async copy per step when dirty.
  double *h_vals = nullptr, *d_vals = nullptr;
  bool vals_dirty = true;
  cudaEvent_t ev_vals = nullptr;
  bool vals_inflight = false;

  std::vector<RateOp> rate_ops;
  std::vector<CellSet> cell_sets;
  int *d_ghost_full = nullptr, *d_ghost_ghost = nullptr;
  int n_ghost_copy = 0;

  // multi-GPU
  ncclComm_t comm = nullptr;
  swk_comm *comm_obj = nullptr;
  bool comm_owned = false;
  bool nccl_warm = false;      // connections to the halo peers are set up (outside any graph capture)
  int rank = 0, nranks = 1;
  std::vector<Peer> peers;
  int n_full = 0;              // full triangles occupy device ids [0, n_full) (0: not a prefix)
  int halo_front = 0;          // every halo-source (send) triangle has a device id < halo_front
  cudaStream_t comm_stream = nullptr;
  cudaEvent_t ev_update = nullptr, ev_halo = nullptr;
  bool overlap = true;         // SWK_NO_OVERLAP=1: halo exchange in-stream

  int64_t launches = 0;
  int64_t launches_mark = 0;

  // one whole timestep captured as a CUDA graph: small meshes are launch-bound (a 40k-triangle
  // step is ~10 kernels of a few microseconds each)
  cudaGraph_t graph[3] = {nullptr, nullptr, nullptr};          // whole step, first half, second half
  cudaGraphExec_t graph_exec[3] = {nullptr, nullptr, nullptr};
  bool graph_valid = false;
  bool use_graph = true;
  bool capturing = false;
  int64_t launches_per_graph[3] = {0, 0, 0};

  // optional per-kernel event timing (bench.py roofline)
  bool timing = false;
  std::vector<cudaEvent_t> ev_pool;
  std::vector<int> ev_kind;
  size_t ev_used = 0;
  double k_ms[4] = {0, 0, 0, 0};
  int64_t k_n[4] = {0, 0, 0, 0};
};

struct TimedScope {
  swk_domain *d;
  bool on = false;
  size_t slot = 0;
  TimedScope(swk_domain *d_, int kind) : d(d_)
  {
    if (!d->timing || d->ev_used + 2 > d->ev_pool.size()) return;
    on = true;
    slot = d->ev_used;
    d->ev_used += 2;
    d->ev_kind[slot / 2] = kind;
    cudaEventRecord(d->ev_pool[slot], d->stream);
  }
  ~TimedScope()
  {
    if (on) cudaEventRecord(d->ev_pool[slot + 1], d->stream);
  }
};

static void make_consts(swk_domain *d)
{
  const swk_params &p = d->P;
  Consts &K = d->K;
  K.epsilon = p.epsilon;
  K.g = p.g;
  K.mah = p.minimum_allowed_height;
  K.beta_w = p.beta_w; K.beta_w_dry = p.beta_w_dry;
  K.beta_uh = p.beta_uh; K.beta_uh_dry = p.beta_uh_dry;
  K.beta_vh = p.beta_vh; K.beta_vh_dry = p.beta_vh_dry;
  K.evolve_max_timestep = p.evolve_max_timestep;
  K.vel2 = p.extrapolate_velocity_second_order ? 1 : 0;
  K.low_froude = (int)p.low_froude;
  K.protect = 1;
  K.pad = 0;
  K.sqrt_g = pow(p.g, 0.5);             // Python's gravity**0.5 (boundaries.py:802): the host's pow
  TimeParams &T = d->TP;
  T.CFL = p.CFL;
  T.evolve_max_timestep = p.evolve_max_timestep;
  T.evolve_min_timestep = p.evolve_min_timestep;
  T.fixed_flux_timestep = p.fixed_flux_timestep;
  T.epsilon = p.epsilon;
  T.max_smallsteps = (int)p.max_smallsteps;
  T.default_order = (int)p.default_order;
  T.method = (int)p.timestepping_method;
}

static int check_params(const swk_params *p)
{
  if (!p) return fail(SWK_ERR_ARG, "params is NULL");
  if (p->timestepping_method < 1 || p->timestepping_method > 3)
    return fail(SWK_ERR_ARG, "timestepping_method must be 1 (euler), 2 (rk2) or 3 (rk3)");
  if (p->maximum_allowed_speed != 0.0)
    return fail(SWK_ERR_UNSUPPORTED, "maximum_allowed_speed != 0 is not part of the DE path");
  return SWK_OK;
}

template <typename T>
static int dalloc(T **p, size_t n)
{
  *p = nullptr;
  if (n == 0) n = 1;
  cudaError_t e = cudaMalloc((void **)p, n * sizeof(T));
  if (e != cudaSuccess) return fail(SWK_ERR_CUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e));
  return SWK_OK;
}

template <typename T>
static int upload(T *dst, const std::vector<T> &v)
{
  if (v.empty()) return SWK_OK;
  CK(cudaMemcpy(dst, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return SWK_OK;
}

extern "C" const char *swk_last_error(void) { return g_err.c_str(); }
extern "C" int swk_abi_version(void) { return SWK_ABI_VERSION; }

extern "C" int swk_device_count(int *count)
{
  if (!count) return fail(SWK_ERR_ARG, "count is NULL");
  *count = 0;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) return fail(SWK_ERR_CUDA, std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e));
  int ok = 0;
  for (int i = 0; i < n; i++) {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, i) == cudaSuccess && prop.major == 10) ok++;
  }
  *count = ok;
  if (ok == 0) return fail(SWK_ERR_CUDA, "no sm_100 device visible");
  return SWK_OK;
}

static int select_device(int device)
{
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return fail(SWK_ERR_CUDA, std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "count is 0") +
                                  " (libswk has no CPU fallback)");
  if (device < 0 || device >= n) return fail(SWK_ERR_ARG, "device index out of range");
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(SWK_ERR_CUDA, std::string("device ") + prop.name + " is not sm_100: libswk contains sm_100a code only");
  CK(cudaSetDevice(device));
  return SWK_OK;
}

// ----------------------------------------------------------------------------
// host-side set-up helper: neighbour structure by sorting the directed edges
// ----------------------------------------------------------------------------
struct EdgeKey { uint64_t key; uint32_t idx; uint32_t pad; };

// LSD radix sort of 16-byte (key, idx) records on the low `bits` bits of the key
static void radix_sort_keys(std::vector<EdgeKey> &a, int bits)
{
  std::vector<EdgeKey> b(a.size());
  const int RB = 11, R = 1 << RB;
  std::vector<size_t> count(R);
  for (int shift = 0; shift < bits; shift += RB) {
    std::fill(count.begin(), count.end(), 0);
    for (const EdgeKey &x : a) count[(x.key >> shift) & (R - 1)]++;
    size_t sum = 0;
    for (int r = 0; r < R; r++) { const size_t c = count[r]; count[r] = sum; sum += c; }
    for (const EdgeKey &x : a) b[count[(x.key >> shift) & (R - 1)]++] = x;
    a.swap(b);
  }
}

extern "C" int swk_mesh_geometry(int64_t N, int64_t nn, const double *nodes, const int64_t *tri, int64_t inscribed,
                                 double *V, double *areas, double *normals, double *edgelengths, double *centroids,
                                 double *radii, double *E, int64_t *first_degenerate)
{
  if (N < 0 || nn <= 0 || !nodes || !tri || !V || !areas || !normals || !edgelengths || !centroids || !radii || !E)
    return fail(SWK_ERR_ARG, "bad argument");
  int64_t bad = -1;
  for (int64_t k = 0; k < N; k++) {
    double x[3], y[3];
    for (int v = 0; v < 3; v++) {
      const int64_t p = tri[3 * k + v];
      if (p < 0 || p >= nn) return fail(SWK_ERR_ARG, "triangle refers to a node >= number_of_nodes");
      x[v] = nodes[2 * p];
      y[v] = nodes[2 * p + 1];
      V[6 * k + 2 * v] = x[v];
      V[6 * k + 2 * v + 1] = y[v];
    }
    const double a = -((x[1] * y[0] - x[0] * y[1]) + (x[2] * y[1] - x[1] * y[2]) + (x[0] * y[2] - x[2] * y[0])) / 2.0;
    areas[k] = a;
    if (!(a > 0.0) && bad < 0) bad = k;
    for (int e = 0; e < 3; e++) {           // edge e runs from vertex e+1 to vertex e+2
      const int p = (e + 1) % 3, q = (e + 2) % 3;
      double xn = x[q] - x[p], yn = y[q] - y[p];
      const double l = sqrt(xn * xn + yn * yn);
      xn = xn / l;
      yn = yn / l;
      normals[6 * k + 2 * e] = yn;
      normals[6 * k + 2 * e + 1] = -xn;
      edgelengths[3 * k + e] = l;
      E[6 * k + 2 * e] = 0.5 * (x[p] + x[q]);
      E[6 * k + 2 * e + 1] = 0.5 * (y[p] + y[q]);
    }
    const double cx = (x[0] + x[1] + x[2]) / 3, cy = (y[0] + y[1] + y[2]) / 3;
    centroids[2 * k] = cx;
    centroids[2 * k + 1] = cy;
    if (!inscribed) {
      double d[3];
      for (int e = 0; e < 3; e++) {
        const int p = (e + 1) % 3, q = (e + 2) % 3;
        const double xm = (x[p] + x[q]) / 2, ym = (y[p] + y[q]) / 2;
        const double dx = cx - xm, dy = cy - ym;
        d[e] = sqrt(dx * dx + dy * dy);
      }
      const double d01 = d[0] < d[1] ? d[0] : d[1];
      radii[k] = d01 < d[2] ? d01 : d[2];
    } else {
      // perimeter from the vertex distances in the order |v0v1| + |v1v2| + |v2v0|
      const double s0 = sqrt((x[0] - x[1]) * (x[0] - x[1]) + (y[0] - y[1]) * (y[0] - y[1]));
      const double s1 = sqrt((x[1] - x[2]) * (x[1] - x[2]) + (y[1] - y[2]) * (y[1] - y[2]));
      const double s2 = sqrt((x[2] - x[0]) * (x[2] - x[0]) + (y[2] - y[0]) * (y[2] - y[0]));
      radii[k] = 2.0 * a / (s0 + s1 + s2);
    }
  }
  if (first_degenerate) *first_degenerate = bad;
  return SWK_OK;
}

extern "C" int swk_build_neighbour_structure(int64_t N, int64_t nn, const int64_t *tri, int64_t *neighbours,
                                             int64_t *neighbour_edges, int64_t *number_of_boundaries)
{
  if (N < 0 || nn <= 0 || !tri || !neighbours || !neighbour_edges || !number_of_boundaries)
    return fail(SWK_ERR_ARG, "bad argument");
  if ((double)nn * (double)nn >= 9.0e18) return fail(SWK_ERR_ARG, "too many nodes");
  // directed edge (a, b) of triangle k, edge e (opposite vertex e): key a*nn+b; its twin is (b, a)
  static const int A[3] = {1, 2, 0}, B[3] = {2, 0, 1};
  std::vector<EdgeKey> fwd((size_t)3 * N), rev((size_t)3 * N);
  for (int64_t k = 0; k < N; k++)
    for (int e = 0; e < 3; e++) {
      const uint64_t a = (uint64_t)tri[3 * k + A[e]], b = (uint64_t)tri[3 * k + B[e]];
      if (a >= (uint64_t)nn || b >= (uint64_t)nn) return fail(SWK_ERR_ARG, "triangle refers to a node >= number_of_nodes");
      fwd[3 * k + e] = {a * (uint64_t)nn + b, (uint32_t)(3 * k + e), 0};
      rev[3 * k + e] = {b * (uint64_t)nn + a, (uint32_t)(3 * k + e), 0};
    }
  int bits = 1;
  while (bits < 63 && ((uint64_t)1 << bits) < (uint64_t)nn * (uint64_t)nn) bits++;
  radix_sort_keys(fwd, bits);
  radix_sort_keys(rev, bits);
  for (size_t j = 1; j < fwd.size(); j++)
    if (fwd[j].key == fwd[j - 1].key)
      return fail(SWK_ERR_ARG, "Edge " + std::to_string(fwd[j].idx % 3) + " of triangle " +
                                   std::to_string(fwd[j].idx / 3) + " is duplicating an edge of another triangle");
  for (int64_t k = 0; k < N; k++) number_of_boundaries[k] = 3;
  for (int64_t j = 0; j < 3 * N; j++) { neighbours[j] = -1; neighbour_edges[j] = -1; }
  size_t f = 0;
  for (size_t r = 0; r < rev.size(); r++) {          // both sorted: one linear merge
    while (f < fwd.size() && fwd[f].key < rev[r].key) f++;
    if (f < fwd.size() && fwd[f].key == rev[r].key) {
      const uint32_t me = rev[r].idx, other = fwd[f].idx;
      neighbours[me] = other / 3;
      neighbour_edges[me] = other % 3;
      number_of_boundaries[me / 3]--;
    }
  }
  return SWK_OK;
}

// ----------------------------------------------------------------------------
// create / destroy
// ----------------------------------------------------------------------------
extern "C" int swk_destroy(swk_domain *d)
{
  if (!d) return SWK_OK;
  cudaSetDevice(d->device);
  if (d->stream) cudaStreamSynchronize(d->stream);
  void *ptrs[] = {d->cq, d->eq, d->xg, d->fg, d->bq, d->connA, d->connB, d->eu, d->bk, d->eta, d->max_speed,
                  d->vcoord, d->wind, d->xbed, d->zflag, d->rw_counter, d->rw_rowIndex, d->rw_list, d->rw_elevation, d->rw_hydraulic,
                  d->d_clock, d->staging, d->acct_val, d->pos_b, d->acct_keys, d->acct_keys_pos, d->b_cell, d->b_edge, d->b_seg, d->d_seg_kind,
                  d->d_vals, d->d_new2old, d->d_ghost_full, d->d_ghost_ghost, d->d_ident_b};
  for (void *p : ptrs)
    if (p) cudaFree(p);
  for (auto &op : d->rate_ops) {
    if (op.d_rate_array) cudaFree(op.d_rate_array);
    if (op.d_indices) cudaFree(op.d_indices);
    if (op.d_partial) cudaFree(op.d_partial);
  }
  for (double *t : d->seg_tab)
    if (t) cudaFree(t);
  for (auto &cs : d->cell_sets) {
    if (cs.d_ids) cudaFree(cs.d_ids);
    if (cs.d_buf) cudaFree(cs.d_buf);
    if (cs.h_buf) cudaFreeHost(cs.h_buf);
    if (cs.ev) cudaEventDestroy(cs.ev);
  }
  if (d->d_seg_tab) cudaFree(d->d_seg_tab);
  if (d->d_seg_np) cudaFree(d->d_seg_np);
  if (d->d_b_point) cudaFree(d->d_b_point);
  for (auto &pe : d->peers) {
    if (pe.d_send_ids) cudaFree(pe.d_send_ids);
    if (pe.d_recv_ids) cudaFree(pe.d_recv_ids);
    if (pe.d_send_buf) cudaFree(pe.d_send_buf);
    if (pe.d_recv_buf) cudaFree(pe.d_recv_buf);
  }
  if (d->comm_obj && d->comm_owned) swk_comm_destroy(d->comm_obj);
  for (int w = 0; w < 3; w++) {
    if (d->graph_exec[w]) cudaGraphExecDestroy(d->graph_exec[w]);
    if (d->graph[w]) cudaGraphDestroy(d->graph[w]);
  }
  if (d->h_clock) cudaFreeHost(d->h_clock);
  if (d->h_vals) cudaFreeHost(d->h_vals);
  if (d->ev_vals) cudaEventDestroy(d->ev_vals);
  if (d->ev_update) cudaEventDestroy(d->ev_update);
  if (d->ev_halo) cudaEventDestroy(d->ev_halo);
  if (d->comm_stream) cudaStreamDestroy(d->comm_stream);
  if (d->stream) cudaStreamDestroy(d->stream);
  delete d;
  return SWK_OK;
}

static int create_impl(const swk_mesh *m, const swk_params *p, int device, swk_domain *d)
{
  d->device = device;
  d->P = *p;
  make_consts(d);
  const int64_t N = d->N = m->number_of_elements;
  const int64_t M = d->M = m->boundary_length;
  if (N <= 0) return fail(SWK_ERR_ARG, "number_of_elements must be positive");
  if (N >= (1LL << 29)) return fail(SWK_ERR_ARG, "number_of_elements must be < 2^29 per device");
  const int64_t NP = d->NP = (N + 63) / 64 * 64;
  if (!m->neighbours || !m->neighbour_edges || !m->surrogate_neighbours || !m->number_of_boundaries ||
      !m->tri_full_flag || !m->normals || !m->edgelengths || !m->radii || !m->areas ||
      !m->centroid_coordinates || !m->edge_coordinates || (M > 0 && (!m->boundary_cells || !m->boundary_edges)))
    return fail(SWK_ERR_ARG, "mesh has NULL arrays");
  CK(cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking));
  {
    const char *env = getenv("SWK_NO_GRAPH");
    d->use_graph = !(env && env[0] == '1');
  }

  // permutation
  d->new2old.resize(N);
  d->old2new.resize(N);
  if (m->permutation) {
    std::vector<char> seen(N, 0);
    for (int64_t k = 0; k < N; k++) {
      const int64_t o = m->permutation[k];
      if (o < 0 || o >= N || seen[o]) return fail(SWK_ERR_ARG, "permutation is not a bijection");
      seen[o] = 1;
      d->new2old[k] = (int)o;
      d->old2new[o] = (int)k;
    }
  } else {
    for (int64_t k = 0; k < N; k++) d->new2old[k] = d->old2new[k] = (int)k;
  }
  const std::vector<int> &n2o = d->new2old, &o2n = d->old2new;
  {   // full triangles first, ghosts after (the reference's local numbering; kept by the reordering)?
    int64_t nf = 0;
    while (nf < N && m->tri_full_flag[n2o[nf]] == 1) nf++;
    bool prefix = true;
    for (int64_t k = nf; k < N; k++)
      if (m->tri_full_flag[n2o[k]] == 1) { prefix = false; break; }
    d->n_full = prefix ? (int)nf : 0;
  }

  // riverwalls?
  d->has_riverwalls = false;
  if (m->edge_flux_type && m->number_of_riverwall_edges > 0)
    for (int64_t j = 0; j < 3 * N; j++)
      if (m->edge_flux_type[j] == 1) { d->has_riverwalls = true; break; }

  // ---- static records, built in device order with the reference's arithmetic ----
  std::vector<i4> connA(NP), connB(NP);
  std::vector<d4> xg(3 * NP), fg(3 * NP);
  const double *cc = m->centroid_coordinates, *ec = m->edge_coordinates;
  for (int64_t k = 0; k < NP; k++) {
    if (k >= N) {   // padding: self-referencing dry cell, never executed (k >= N guards) but keep it sane
      connA[k] = {(int)0, (int)0, (int)0, 3};
      connB[k] = {-1, -1, -1, 0};
      continue;
    }
    const int64_t o = n2o[k];
    const double x = cc[2 * o], y = cc[2 * o + 1];
    double dxv[3], dyv[3];
    for (int i = 0; i < 3; i++) {
      dxv[i] = ec[6 * o + 2 * i] - x;           // sw_domain_openmp.c:1445-1450
      dyv[i] = ec[6 * o + 2 * i + 1] - y;
    }
    const int64_t s0 = m->surrogate_neighbours[3 * o], s1 = m->surrogate_neighbours[3 * o + 1],
                  s2 = m->surrogate_neighbours[3 * o + 2];
    const int nb = (int)m->number_of_boundaries[o];
    double dx1 = 0, dx2 = 0, dy1 = 0, dy2 = 0, inv_area2 = 0;
    int which = 0;
    if (nb <= 1) {
      const double x0 = cc[2 * s0], y0 = cc[2 * s0 + 1];
      const double x1 = cc[2 * s1], y1 = cc[2 * s1 + 1];
      const double x2 = cc[2 * s2], y2 = cc[2 * s2 + 1];
      dx1 = x1 - x0; dx2 = x2 - x0; dy1 = y1 - y0; dy2 = y2 - y0;    // :1477-1480
      const double area2 = dy2 * dx1 - dy1 * dx2;                    // :1484
      inv_area2 = 1.0 / area2;                                       // :1553
    } else if (nb == 2) {
      const int64_t ss[3] = {s0, s1, s2};
      which = 2;
      for (int i = 0; i < 3; i++)
        if (ss[i] != o) { which = i; break; }                        // :1656-1674
      const int64_t kn = ss[which];
      dx1 = cc[2 * kn] - x;                                          // :1684-1696
      dy1 = cc[2 * kn + 1] - y;
      const double dist2 = dx1 * dx1 + dy1 * dy1;
      dx2 = 1.0 / dist2;
      dy2 = dx2 * dy1;
      dx2 *= dx1;
    }
    const int fullbit = (m->tri_full_flag[o] == 1) ? 1 : 0;
    connA[k] = {o2n[s0], o2n[s1], o2n[s2], (nb & 3) | (which << 2) | (fullbit << 4)};
    xg[k] = {dxv[0], dxv[1], dxv[2], dyv[0]};
    xg[NP + k] = {dyv[1], dyv[2], dx1, dx2};
    xg[2 * NP + k] = {dy1, dy2, inv_area2, m->areas[o]};

    const double *nr = m->normals + 6 * o;
    const double *el = m->edgelengths + 3 * o;
    fg[k] = {nr[0], nr[1], el[0], 1.0 / m->areas[o]};                        // inv_area: :714
    fg[NP + k] = {nr[2], nr[3], el[1], m->radii[o]};
    fg[2 * NP + k] = {nr[4], nr[5], el[2], m->areas[o]};

    int pk[3];
    int flags = (m->tri_full_flag[o] == 1) ? 1 : 0;
    for (int i = 0; i < 3; i++) {
      const int64_t n = m->neighbours[3 * o + i];
      if (n < 0) {
        if (-n - 1 >= M) return fail(SWK_ERR_ARG, "neighbours holds a boundary index >= boundary_length");
        pk[i] = (int)n;
      } else {
        pk[i] = (o2n[n] << 2) | (int)m->neighbour_edges[3 * o + i];
      }
      if (d->has_riverwalls && m->edge_flux_type[3 * o + i] == 1) flags |= (2 << i);
      if (n >= 0 && m->tri_full_flag[o] == 1 && m->tri_full_flag[n] == 0) flags |= (16 << i);
    }
    connB[k] = {pk[0], pk[1], pk[2], flags};
  }
  // boundary-flux accounting edges in the reference's (k, i) order (:696): one slot each
  std::vector<int> pos_b(std::max<int64_t>(M, 1), -1);
  std::vector<std::pair<int, int>> ghost_keys;     // (device key, slot)
  {
    int slot = 0;
    double area_sum = 0.0;
    for (int64_t o = 0; o < N; o++) {
      if (m->tri_full_flag[o] != 1) continue;
      area_sum += m->areas[o];
      for (int i = 0; i < 3; i++) {
        const int64_t n = m->neighbours[3 * o + i];
        if (n < 0) pos_b[-n - 1] = slot++;
        else if (m->tri_full_flag[n] == 0) ghost_keys.push_back({(o2n[o] << 2) | i, slot++});
      }
    }
    d->n_acct = slot;
    d->full_area = area_sum;
  }
  std::sort(ghost_keys.begin(), ghost_keys.end());
  std::vector<int> gk(ghost_keys.size()), gp(ghost_keys.size());
  for (size_t j = 0; j < ghost_keys.size(); j++) { gk[j] = ghost_keys[j].first; gp[j] = ghost_keys[j].second; }
  d->n_acct_keys = (int)gk.size();
  CKV(dalloc(&d->pos_b, pos_b.size())); CKV(upload(d->pos_b, pos_b));
  CKV(dalloc(&d->acct_keys, gk.size())); CKV(upload(d->acct_keys, gk));
  CKV(dalloc(&d->acct_keys_pos, gp.size())); CKV(upload(d->acct_keys_pos, gp));
  CKV(dalloc(&d->acct_val, (size_t)d->n_acct));
  CK(cudaMemset(d->acct_val, 0, std::max(d->n_acct, 1) * sizeof(double)));

  CKV(dalloc(&d->connA, NP)); CKV(upload(d->connA, connA));
  CKV(dalloc(&d->connB, NP)); CKV(upload(d->connB, connB));
  CKV(dalloc(&d->xg, 3 * NP)); CKV(upload(d->xg, xg));
  CKV(dalloc(&d->fg, 3 * NP)); CKV(upload(d->fg, fg));
  CKV(dalloc(&d->d_new2old, N)); CKV(upload(d->d_new2old, d->new2old));
  {
    std::vector<i4>().swap(connA); std::vector<i4>().swap(connB);
    std::vector<d4>().swap(xg); std::vector<d4>().swap(fg);
  }

  if (p->use_sloped_mannings) {
    if (!m->vertex_coordinates) return fail(SWK_ERR_ARG, "sloped Manning needs vertex_coordinates");
    std::vector<double> vc(6 * NP, 0.0);
    for (int64_t k = 0; k < N; k++)
      for (int j = 0; j < 6; j++) vc[j * NP + k] = m->vertex_coordinates[6 * n2o[k] + j];
    CKV(dalloc(&d->vcoord, 6 * NP)); CKV(upload(d->vcoord, vc));
  }
  if (d->has_riverwalls) {
    std::vector<int> rc(3 * NP, 0);
    int64_t maxc = 0;
    for (int64_t k = 0; k < N; k++)
      for (int i = 0; i < 3; i++) {
        rc[i * NP + k] = (int)m->edge_river_wall_counter[3 * n2o[k] + i];
        maxc = std::max<int64_t>(maxc, rc[i * NP + k]);
      }
    if (maxc > m->number_of_riverwall_edges) return fail(SWK_ERR_ARG, "edge_river_wall_counter exceeds number_of_riverwall_edges");
    CKV(dalloc(&d->rw_counter, 3 * NP)); CKV(upload(d->rw_counter, rc));
    std::vector<int> wl;
    for (int64_t k = 0; k < N; k++)
      if (rc[k] || rc[NP + k] || rc[2 * NP + k]) wl.push_back((int)k);
    d->n_rw_list = (int)wl.size();
    CKV(dalloc(&d->rw_list, wl.size())); CKV(upload(d->rw_list, wl));
    d->h_rw_list = wl;
    const int64_t nrw = m->number_of_riverwall_edges;
    std::vector<double> re(m->riverwall_elevation, m->riverwall_elevation + nrw);
    std::vector<int> ri(nrw);
    int64_t maxrow = 0;
    for (int64_t j = 0; j < nrw; j++) { ri[j] = (int)m->riverwall_rowIndex[j]; maxrow = std::max<int64_t>(maxrow, ri[j]); }
    const int64_t ncol = m->ncol_riverwall_hydraulic_properties;
    std::vector<double> rh(m->riverwall_hydraulic_properties, m->riverwall_hydraulic_properties + (maxrow + 1) * ncol);
    CKV(dalloc(&d->rw_elevation, nrw)); CKV(upload(d->rw_elevation, re));
    CKV(dalloc(&d->rw_rowIndex, nrw)); CKV(upload(d->rw_rowIndex, ri));
    CKV(dalloc(&d->rw_hydraulic, rh.size())); CKV(upload(d->rw_hydraulic, rh));
  }

  // boundary index arrays
  {
    std::vector<int> bc(M), be(M), ident(std::max<int64_t>(M, 1));
    for (int64_t j = 0; j < M; j++) {
      const int64_t vol = m->boundary_cells[j];
      if (vol < 0 || vol >= N) return fail(SWK_ERR_ARG, "boundary_cells out of range");
      bc[j] = o2n[vol];
      be[j] = (int)m->boundary_edges[j];
      ident[j] = (int)j;
    }
    d->h_b_seg.assign(M, -1);
    CKV(dalloc(&d->b_cell, M)); CKV(upload(d->b_cell, bc));
    CKV(dalloc(&d->b_edge, M)); CKV(upload(d->b_edge, be));
    CKV(dalloc(&d->b_seg, M)); CKV(upload(d->b_seg, d->h_b_seg));
    CKV(dalloc(&d->d_ident_b, M)); CKV(upload(d->d_ident_b, ident));
  }

  // dynamic arrays
  CKV(dalloc(&d->cq, NP)); CK(cudaMemset(d->cq, 0, NP * sizeof(d4)));
  CKV(dalloc(&d->eq, 3 * NP)); CK(cudaMemset(d->eq, 0, 3 * NP * sizeof(d4)));
  CKV(dalloc(&d->bq, M)); CK(cudaMemset(d->bq, 0, std::max<int64_t>(M, 1) * sizeof(d4)));
  CKV(dalloc(&d->eu, 3 * NP)); CK(cudaMemset(d->eu, 0, 3 * NP * sizeof(double)));
  CKV(dalloc(&d->bk, 3 * NP)); CK(cudaMemset(d->bk, 0, 3 * NP * sizeof(double)));
  CKV(dalloc(&d->eta, NP)); CK(cudaMemset(d->eta, 0, NP * sizeof(double)));
  CKV(dalloc(&d->max_speed, NP)); CK(cudaMemset(d->max_speed, 0, NP * sizeof(double)));
  CKV(dalloc(&d->zflag, NP)); CK(cudaMemset(d->zflag, 0, NP));
  CKV(dalloc(&d->d_clock, 1)); CK(cudaMemset(d->d_clock, 0, sizeof(Clock)));
  CK(cudaHostAlloc((void **)&d->h_clock, sizeof(Clock), cudaHostAllocDefault));
  memset(d->h_clock, 0, sizeof(Clock));
  d->h_clock->order = (int)p->default_order;
  d->h_clock->finaltime = -1.0;
  d->h_clock->yieldtime = 0.0;
  d->h_clock->recorded_min_timestep = p->evolve_max_timestep;
  d->h_clock->recorded_max_timestep = p->evolve_min_timestep;
  d->h_clock->dt_min_bits = 0x7FF0000000000000ULL;
  CK(cudaMemcpy(d->d_clock, d->h_clock, sizeof(Clock), cudaMemcpyHostToDevice));
  CK(cudaHostAlloc((void **)&d->h_vals, VALS_TOTAL * sizeof(double), cudaHostAllocDefault));
  memset(d->h_vals, 0, VALS_TOTAL * sizeof(double));
  CKV(dalloc(&d->d_vals, VALS_TOTAL)); CK(cudaMemset(d->d_vals, 0, VALS_TOTAL * sizeof(double)));
  CK(cudaEventCreateWithFlags(&d->ev_vals, cudaEventDisableTiming));

  Dev &D = d->D;
  D.N = (int)N; D.NP = (int)NP; D.M = (int)M;
  D.cq = d->cq; D.eq = d->eq; D.xg = d->xg; D.fg = d->fg;
  D.connA = d->connA; D.connB = d->connB;
  D.eu = d->eu; D.bk = d->bk; D.eta = d->eta; D.zflag = d->zflag; D.max_speed = d->max_speed;
  D.bq = d->bq; D.vcoord = d->vcoord; D.wind = nullptr; D.clock = d->d_clock;
  D.bed_e_x = nullptr; D.hc_x = nullptr;
  D.acct_val = d->acct_val; D.pos_b = d->pos_b; D.acct_keys = d->acct_keys; D.acct_keys_pos = d->acct_keys_pos;
  D.n_acct = d->n_acct; D.n_acct_keys = d->n_acct_keys;
  D.rw_counter = d->rw_counter; D.rw_elevation = d->rw_elevation; D.rw_rowIndex = d->rw_rowIndex;
  D.rw_hydraulic = d->rw_hydraulic; D.rw_ncol = (int)m->ncol_riverwall_hydraulic_properties;
  CK(cudaDeviceSynchronize());
  return SWK_OK;
}

extern "C" int swk_create(const swk_mesh *mesh, const swk_params *params, int device, swk_domain **out)
{
  if (!mesh || !out) return fail(SWK_ERR_ARG, "mesh/out is NULL");
  *out = nullptr;
  CKV(check_params(params));
  CKV(select_device(device));
  swk_domain *d = new swk_domain();
  int r = create_impl(mesh, params, device, d);
  if (r != SWK_OK) {
    std::string keep = g_err;
    swk_destroy(d);
    g_err = keep;
    return r;
  }
  *out = d;
  return SWK_OK;
}

extern "C" int swk_set_params(swk_domain *d, const swk_params *params)
{
  if (!d) return fail(SWK_ERR_ARG, "handle is NULL");
  CKV(check_params(params));
  if (params->use_sloped_mannings && !d->vcoord)
    return fail(SWK_ERR_ARG, "use_sloped_mannings must be chosen at swk_create (vertex coordinates are uploaded there)");
  d->P = *params;
  make_consts(d);
  d->graph_valid = false;
  return SWK_OK;
}

// ----------------------------------------------------------------------------
// quantity transfer
// ----------------------------------------------------------------------------
static int ensure_staging(swk_domain *d, size_t n)
{
  if (d->staging_n >= n) return SWK_OK;
  if (d->staging) cudaFree(d->staging);
  d->staging = nullptr;
  d->staging_n = 0;
  CKV(dalloc(&d->staging, n));
  d->staging_n = n;
  return SWK_OK;
}

#define LAUNCH(d, kernel, grid, block, ...)                          \
  do {                                                               \
    kernel<<<(grid), (block), 0, (d)->stream>>>(__VA_ARGS__);        \
    (d)->launches++;                                                 \
  } while (0)

static int pull_clock(swk_domain *d);
static int push_clock(swk_domain *d);

static int sync_check(swk_domain *d)
{
  CK(cudaStreamSynchronize(d->stream));
  CK(cudaGetLastError());
  return SWK_OK;
}

extern "C" int swk_set_quantity(swk_domain *d, int q, const double *host, int64_t n)
{
  if (!d || !host) return fail(SWK_ERR_ARG, "NULL argument");
  CK(cudaSetDevice(d->device));
  const int N = (int)d->N, NP = (int)d->NP, M = (int)d->M;
  int64_t expect = N;
  if ((q >= 10 && q < 30)) expect = 3LL * N;
  if (q >= 30 && q < 40) expect = M;
  if (n != expect) return fail(SWK_ERR_ARG, "swk_set_quantity: wrong element count");
  if (n == 0) return SWK_OK;
  CKV(ensure_staging(d, (size_t)std::max<int64_t>(3LL * N, M)));
  CK(cudaMemcpyAsync(d->staging, host, n * sizeof(double), cudaMemcpyHostToDevice, d->stream));
  switch (q) {
    case SWK_Q_STAGE_C: LAUNCH(d, k_scatter_component, nblk(N), BLOCK, d->cq, NP, N, 1, 0, d->staging, d->d_new2old); break;
    case SWK_Q_XMOM_C: LAUNCH(d, k_scatter_component, nblk(N), BLOCK, d->cq, NP, N, 1, 1, d->staging, d->d_new2old); break;
    case SWK_Q_YMOM_C: LAUNCH(d, k_scatter_component, nblk(N), BLOCK, d->cq, NP, N, 1, 2, d->staging, d->d_new2old); break;
    case SWK_Q_ELEVATION_C: LAUNCH(d, k_scatter_component, nblk(N), BLOCK, d->cq, NP, N, 1, 3, d->staging, d->d_new2old); break;
    case SWK_Q_FRICTION_C: LAUNCH(d, k_scatter_plain, nblk(N), BLOCK, d->eta, NP, N, 1, d->staging, d->d_new2old); break;
    case SWK_Q_STAGE_E: LAUNCH(d, k_scatter_component, nblk(3LL * N), BLOCK, d->eq, NP, N, 3, 0, d->staging, d->d_new2old); break;
    case SWK_Q_HEIGHT_E: LAUNCH(d, k_scatter_component, nblk(3LL * N), BLOCK, d->eq, NP, N, 3, 1, d->staging, d->d_new2old); break;
    case SWK_Q_XMOM_E: LAUNCH(d, k_scatter_component, nblk(3LL * N), BLOCK, d->eq, NP, N, 3, 2, d->staging, d->d_new2old); break;
    case SWK_Q_YMOM_E: LAUNCH(d, k_scatter_component, nblk(3LL * N), BLOCK, d->eq, NP, N, 3, 3, d->staging, d->d_new2old); break;
    case SWK_Q_STAGE_B: LAUNCH(d, k_scatter_component, nblk(M), BLOCK, d->bq, M, M, 1, 0, d->staging, d->d_ident_b); break;
    case SWK_Q_XMOM_B: LAUNCH(d, k_scatter_component, nblk(M), BLOCK, d->bq, M, M, 1, 1, d->staging, d->d_ident_b); break;
    case SWK_Q_YMOM_B: LAUNCH(d, k_scatter_component, nblk(M), BLOCK, d->bq, M, M, 1, 2, d->staging, d->d_ident_b); break;
    case SWK_Q_STAGE_EU: LAUNCH(d, k_scatter_plain, nblk(N), BLOCK, d->eu, NP, N, 1, d->staging, d->d_new2old); break;
    case SWK_Q_XMOM_EU: LAUNCH(d, k_scatter_plain, nblk(N), BLOCK, d->eu + NP, NP, N, 1, d->staging, d->d_new2old); break;
    case SWK_Q_YMOM_EU: LAUNCH(d, k_scatter_plain, nblk(N), BLOCK, d->eu + 2 * NP, NP, N, 1, d->staging, d->d_new2old); break;
    case SWK_Q_STAGE_BACKUP: LAUNCH(d, k_scatter_plain, nblk(N), BLOCK, d->bk, NP, N, 1, d->staging, d->d_new2old); break;
    case SWK_Q_XMOM_BACKUP: LAUNCH(d, k_scatter_plain, nblk(N), BLOCK, d->bk + NP, NP, N, 1, d->staging, d->d_new2old); break;
    case SWK_Q_YMOM_BACKUP: LAUNCH(d, k_scatter_plain, nblk(N), BLOCK, d->bk + 2 * NP, NP, N, 1, d->staging, d->d_new2old); break;
    default: return fail(SWK_ERR_ARG, "swk_set_quantity: quantity id is not settable");
  }
  return sync_check(d);
}

extern "C" int swk_get_quantity(swk_domain *d, int q, double *host, int64_t n)
{
  if (!d || !host) return fail(SWK_ERR_ARG, "NULL argument");
  CK(cudaSetDevice(d->device));
  const int N = (int)d->N, NP = (int)d->NP, M = (int)d->M;
  int64_t expect = N;
  if ((q >= 10 && q < 30)) expect = 3LL * N;
  if (q >= 30 && q < 40) expect = M;
  if (n != expect) return fail(SWK_ERR_ARG, "swk_get_quantity: wrong element count");
  if (n == 0) return SWK_OK;
  CKV(ensure_staging(d, (size_t)std::max<int64_t>(3LL * N, M)));
  switch (q) {
    case SWK_Q_STAGE_C: LAUNCH(d, k_gather_component, nblk(N), BLOCK, d->cq, NP, N, 1, 0, d->staging, d->d_new2old); break;
    case SWK_Q_XMOM_C: LAUNCH(d, k_gather_component, nblk(N), BLOCK, d->cq, NP, N, 1, 1, d->staging, d->d_new2old); break;
    case SWK_Q_YMOM_C: LAUNCH(d, k_gather_component, nblk(N), BLOCK, d->cq, NP, N, 1, 2, d->staging, d->d_new2old); break;
    case SWK_Q_ELEVATION_C: LAUNCH(d, k_gather_component, nblk(N), BLOCK, d->cq, NP, N, 1, 3, d->staging, d->d_new2old); break;
    case SWK_Q_FRICTION_C: LAUNCH(d, k_gather_plain, nblk(N), BLOCK, d->eta, NP, N, 1, d->staging, d->d_new2old); break;
    case SWK_Q_HEIGHT_C: LAUNCH(d, k_gather_derived, nblk(N), BLOCK, d->D, d->K, 0, d->staging, d->d_new2old); break;
    case SWK_Q_STAGE_E: LAUNCH(d, k_gather_component, nblk(3LL * N), BLOCK, d->eq, NP, N, 3, 0, d->staging, d->d_new2old); break;
    case SWK_Q_HEIGHT_E: LAUNCH(d, k_gather_component, nblk(3LL * N), BLOCK, d->eq, NP, N, 3, 1, d->staging, d->d_new2old); break;
    case SWK_Q_XMOM_E: LAUNCH(d, k_gather_component, nblk(3LL * N), BLOCK, d->eq, NP, N, 3, 2, d->staging, d->d_new2old); break;
    case SWK_Q_YMOM_E: LAUNCH(d, k_gather_component, nblk(3LL * N), BLOCK, d->eq, NP, N, 3, 3, d->staging, d->d_new2old); break;
    case SWK_Q_ELEVATION_E: LAUNCH(d, k_gather_derived, nblk(N), BLOCK, d->D, d->K, 1, d->staging, d->d_new2old); break;
    case SWK_Q_STAGE_V: LAUNCH(d, k_gather_derived, nblk(N), BLOCK, d->D, d->K, 2, d->staging, d->d_new2old); break;
    case SWK_Q_HEIGHT_V: LAUNCH(d, k_gather_derived, nblk(N), BLOCK, d->D, d->K, 3, d->staging, d->d_new2old); break;
    case SWK_Q_XMOM_V: LAUNCH(d, k_gather_derived, nblk(N), BLOCK, d->D, d->K, 4, d->staging, d->d_new2old); break;
    case SWK_Q_YMOM_V: LAUNCH(d, k_gather_derived, nblk(N), BLOCK, d->D, d->K, 5, d->staging, d->d_new2old); break;
    case SWK_Q_ELEVATION_V: LAUNCH(d, k_gather_derived, nblk(N), BLOCK, d->D, d->K, 6, d->staging, d->d_new2old); break;
    case SWK_Q_STAGE_B: LAUNCH(d, k_gather_component, nblk(M), BLOCK, d->bq, M, M, 1, 0, d->staging, d->d_ident_b); break;
    case SWK_Q_XMOM_B: LAUNCH(d, k_gather_component, nblk(M), BLOCK, d->bq, M, M, 1, 1, d->staging, d->d_ident_b); break;
    case SWK_Q_YMOM_B: LAUNCH(d, k_gather_component, nblk(M), BLOCK, d->bq, M, M, 1, 2, d->staging, d->d_ident_b); break;
    case SWK_Q_STAGE_EU: LAUNCH(d, k_gather_plain, nblk(N), BLOCK, d->eu, NP, N, 1, d->staging, d->d_new2old); break;
    case SWK_Q_XMOM_EU: LAUNCH(d, k_gather_plain, nblk(N), BLOCK, d->eu + NP, NP, N, 1, d->staging, d->d_new2old); break;
    case SWK_Q_YMOM_EU: LAUNCH(d, k_gather_plain, nblk(N), BLOCK, d->eu + 2 * NP, NP, N, 1, d->staging, d->d_new2old); break;
    case SWK_Q_MAX_SPEED: LAUNCH(d, k_gather_plain, nblk(N), BLOCK, d->max_speed, NP, N, 1, d->staging, d->d_new2old); break;
    case SWK_Q_STAGE_BACKUP: LAUNCH(d, k_gather_plain, nblk(N), BLOCK, d->bk, NP, N, 1, d->staging, d->d_new2old); break;
    case SWK_Q_XMOM_BACKUP: LAUNCH(d, k_gather_plain, nblk(N), BLOCK, d->bk + NP, NP, N, 1, d->staging, d->d_new2old); break;
    case SWK_Q_YMOM_BACKUP: LAUNCH(d, k_gather_plain, nblk(N), BLOCK, d->bk + 2 * NP, NP, N, 1, d->staging, d->d_new2old); break;
    default: return fail(SWK_ERR_ARG, "swk_get_quantity: unknown quantity id");
  }
  CK(cudaMemcpyAsync(host, d->staging, n * sizeof(double), cudaMemcpyDeviceToHost, d->stream));
  return sync_check(d);
}

// ----------------------------------------------------------------------------
// boundaries, operators, ghosts
// ----------------------------------------------------------------------------
// the host is about to change h_vals: a copy of the previous contents that is still queued must go first
static int vals_writable(swk_domain *d)
{
  if (d->vals_inflight) {
    CK(cudaEventSynchronize(d->ev_vals));
    d->vals_inflight = false;
  }
  return SWK_OK;
}

static int push_values(swk_domain *d)
{
  if (!d->vals_dirty) return SWK_OK;
  const size_t ns = d->seg_kind.size(), no = d->rate_ops.size();
  if (ns > 0) CK(cudaMemcpyAsync(d->d_vals, d->h_vals, 9 * ns * sizeof(double), cudaMemcpyHostToDevice, d->stream));
  if (no > 0)
    CK(cudaMemcpyAsync(d->d_vals + VALS_OPS, d->h_vals + VALS_OPS, 2 * no * sizeof(double), cudaMemcpyHostToDevice,
                       d->stream));
  CK(cudaEventRecord(d->ev_vals, d->stream));
  d->vals_inflight = true;
  d->vals_dirty = false;
  return SWK_OK;
}

static int push_segments(swk_domain *d)
{
  if (d->seg_dirty) {
    const int ns = (int)d->seg_kind.size();
    if (ns > d->seg_cap) {
      if (d->d_seg_kind) cudaFree(d->d_seg_kind);
      if (d->d_seg_tab) cudaFree(d->d_seg_tab);
      if (d->d_seg_np) cudaFree(d->d_seg_np);
      d->seg_cap = std::max(16, 2 * ns);
      CKV(dalloc(&d->d_seg_kind, d->seg_cap));
      CKV(dalloc(&d->d_seg_tab, d->seg_cap));
      CKV(dalloc(&d->d_seg_np, d->seg_cap));
    }
    d->seg_tab.resize(ns, nullptr);
    d->seg_np.resize(ns, 0);
    d->seg_nframes.resize(ns, 0);
    if (ns > 0) {
      CK(cudaMemcpyAsync(d->d_seg_kind, d->seg_kind.data(), ns * sizeof(int), cudaMemcpyHostToDevice, d->stream));
      CK(cudaMemcpyAsync(d->d_seg_tab, d->seg_tab.data(), ns * sizeof(double *), cudaMemcpyHostToDevice, d->stream));
      CK(cudaMemcpyAsync(d->d_seg_np, d->seg_np.data(), ns * sizeof(int), cudaMemcpyHostToDevice, d->stream));
    }
    if (d->M > 0) {
      CK(cudaMemcpyAsync(d->b_seg, d->h_b_seg.data(), d->M * sizeof(int), cudaMemcpyHostToDevice, d->stream));
      if (!d->d_b_point) CKV(dalloc(&d->d_b_point, d->M));
      d->h_b_point.resize(d->M, 0);
      CK(cudaMemcpyAsync(d->d_b_point, d->h_b_point.data(), d->M * sizeof(int), cudaMemcpyHostToDevice, d->stream));
    }
    CK(cudaStreamSynchronize(d->stream));   // host vectors may change after return
    d->seg_dirty = false;
  }
  return push_values(d);
}

extern "C" int swk_set_boundary_segment(swk_domain *d, int segment, int kind, const int64_t *ids, int64_t n_ids,
                                        const double values[3])
{
  if (!d) return fail(SWK_ERR_ARG, "handle is NULL");
  if (segment < 0 || segment >= SEG_MAX) return fail(SWK_ERR_ARG, "segment id out of range");
  if (kind < SWK_BC_NONE || kind > SWK_BC_DIRICHLET_DISCHARGE) return fail(SWK_ERR_ARG, "unknown boundary kind");
  if ((int)d->seg_kind.size() <= segment) d->seg_kind.resize(segment + 1, 0);
  d->seg_kind[segment] = kind;
  CKV(vals_writable(d));
  for (int sub = 0; sub < 3; sub++)
    for (int j = 0; j < 3; j++) d->h_vals[(3 * segment + sub) * 3 + j] = values ? values[j] : 0.0;
  d->vals_dirty = true;
  d->h_b_point.resize(d->M, 0);
  for (int64_t j = 0; j < n_ids; j++) {
    if (ids[j] < 0 || ids[j] >= d->M) return fail(SWK_ERR_ARG, "boundary id out of range");
    d->h_b_seg[ids[j]] = segment;
    d->h_b_point[ids[j]] = (int)j;          // position in the segment: row of its time-space table
  }
  d->seg_dirty = true;
  d->graph_valid = false;
  return SWK_OK;
}

extern "C" int swk_set_boundary_values(swk_domain *d, int segment, const double values[3])
{
  if (!d || !values) return fail(SWK_ERR_ARG, "NULL argument");
  if (segment < 0 || segment >= (int)d->seg_kind.size()) return fail(SWK_ERR_ARG, "unknown segment");
  CKV(vals_writable(d));
  for (int sub = 0; sub < 3; sub++)
    for (int j = 0; j < 3; j++) d->h_vals[(3 * segment + sub) * 3 + j] = values[j];
  d->vals_dirty = true;                // values live in device tables: the graph stays valid
  return SWK_OK;
}

// Time-space table of a segment bound to SWK_BC_TIME_SPACE_TABLE[_MEAN_STAGE]: frames[n_frames][n_points][3]
// (stage, xmomentum, ymomentum), point j belonging to the j-th boundary index given to
// swk_set_boundary_segment.  Uploaded once and kept resident; the boundary values of the segment are then
// {ratio, frame index, mean_stage} (swk_set_boundary_values*).
extern "C" int swk_set_boundary_table(swk_domain *d, int segment, int64_t n_frames, int64_t n_points,
                                      const double *frames)
{
  if (!d || !frames) return fail(SWK_ERR_ARG, "NULL argument");
  if (segment < 0 || segment >= (int)d->seg_kind.size()) return fail(SWK_ERR_ARG, "unknown segment");
  if (n_frames < 1 || n_points < 0) return fail(SWK_ERR_ARG, "bad table shape");
  CK(cudaSetDevice(d->device));
  CKV(sync_check(d));
  d->seg_tab.resize(d->seg_kind.size(), nullptr);
  d->seg_np.resize(d->seg_kind.size(), 0);
  d->seg_nframes.resize(d->seg_kind.size(), 0);
  if (d->seg_tab[segment]) { cudaFree(d->seg_tab[segment]); d->seg_tab[segment] = nullptr; }
  const size_t n = (size_t)n_frames * (size_t)n_points * 3;
  CKV(dalloc(&d->seg_tab[segment], n));
  if (n > 0) CK(cudaMemcpy(d->seg_tab[segment], frames, n * sizeof(double), cudaMemcpyHostToDevice));
  d->seg_np[segment] = (int)n_points;
  d->seg_nframes[segment] = (int)n_frames;
  d->seg_dirty = true;
  d->graph_valid = false;
  return SWK_OK;
}

// rewrite one frame of a segment's table (Time_space_boundary: one frame per RK substep, every step)
extern "C" int swk_set_boundary_table_frame(swk_domain *d, int segment, int64_t frame, const double *values)
{
  if (!d || !values) return fail(SWK_ERR_ARG, "NULL argument");
  if (segment < 0 || segment >= (int)d->seg_tab.size() || !d->seg_tab[segment])
    return fail(SWK_ERR_ARG, "the segment has no table");
  if (frame < 0 || frame >= d->seg_nframes[segment]) return fail(SWK_ERR_ARG, "frame out of range");
  CK(cudaSetDevice(d->device));
  const size_t n = (size_t)d->seg_np[segment] * 3;
  // (pageable source: the copy has left the host buffer when the call returns)
  CK(cudaMemcpyAsync(d->seg_tab[segment] + (size_t)frame * n, values, n * sizeof(double), cudaMemcpyHostToDevice,
                     d->stream));
  CK(cudaStreamSynchronize(d->stream));
  return SWK_OK;
}

// values of one RK substep only (0: at the start of the step, 1: second flux evaluation, 2: third)
extern "C" int swk_set_boundary_values_substep(swk_domain *d, int segment, int substep, const double values[3])
{
  if (!d || !values) return fail(SWK_ERR_ARG, "NULL argument");
  if (segment < 0 || segment >= (int)d->seg_kind.size()) return fail(SWK_ERR_ARG, "unknown segment");
  if (substep < 0 || substep > 2) return fail(SWK_ERR_ARG, "substep must be 0, 1 or 2");
  CKV(vals_writable(d));
  for (int j = 0; j < 3; j++) d->h_vals[(3 * segment + substep) * 3 + j] = values[j];
  d->vals_dirty = true;
  return SWK_OK;
}

extern "C" int swk_add_rate_operator(swk_domain *d, double rate, double factor, const double *rate_array,
                                     const int64_t *indices, int64_t n_indices, int *op_id)
{
  if (!d) return fail(SWK_ERR_ARG, "handle is NULL");
  CK(cudaSetDevice(d->device));
  RateOp op;
  op.rate = rate;
  op.factor = factor;
  op.all_nonneg = 1;
  const int64_t N = d->N;
  if (indices) {
    std::vector<int> idx(n_indices);
    for (int64_t j = 0; j < n_indices; j++) {
      if (indices[j] < 0 || indices[j] >= N) return fail(SWK_ERR_ARG, "rate operator index out of range");
      idx[j] = d->old2new[indices[j]];
    }
    op.n = (int)n_indices;
    CKV(dalloc(&op.d_indices, idx.size())); CKV(upload(op.d_indices, idx));
  } else {
    op.n = (int)N;
  }
  if (rate_array) {   // (N,) in caller order; num.all(rate >= 0) is evaluated on the selected entries
    std::vector<double> ra(d->NP, 0.0);
    for (int64_t k = 0; k < N; k++) ra[k] = rate_array[d->new2old[k]];
    if (indices) { for (int64_t j = 0; j < n_indices; j++) if (!(rate_array[indices[j]] >= 0.0)) op.all_nonneg = 0; }
    else { for (int64_t k = 0; k < N; k++) if (!(rate_array[k] >= 0.0)) op.all_nonneg = 0; }
    CKV(dalloc(&op.d_rate_array, ra.size())); CKV(upload(op.d_rate_array, ra));
  } else {
    op.all_nonneg = (rate >= 0.0) ? 1 : 0;
  }
  op.nblocks = nblk(op.n);
  CKV(dalloc(&op.d_partial, op.nblocks));
  if ((int)d->rate_ops.size() >= OP_MAX) return fail(SWK_ERR_ARG, "too many rate operators");
  CKV(vals_writable(d));
  d->h_vals[VALS_OPS + 2 * d->rate_ops.size()] = rate;
  d->h_vals[VALS_OPS + 2 * d->rate_ops.size() + 1] = factor;
  d->vals_dirty = true;
  d->rate_ops.push_back(op);
  d->graph_valid = false;
  if (op_id) *op_id = (int)d->rate_ops.size() - 1;
  return SWK_OK;
}

// Explicit forcing that does not depend on the state (shallow_water/forcing.py: Wind_stress :80-215 adds S*u, S*v
// to the momentum updates; General_forcing / Rainfall / Inflow :215-640 add a rate to the stage update over a
// region): three (N,) arrays in the caller's triangle order (stage may be NULL = zero), added to the explicit
// updates inside the update kernels.  All NULL switches it off.
extern "C" int swk_set_explicit_forcing(swk_domain *d, const double *stage_force, const double *xmom_force,
                                        const double *ymom_force, int64_t n)
{
  if (!d) return fail(SWK_ERR_ARG, "handle is NULL");
  CK(cudaSetDevice(d->device));
  CKV(sync_check(d));
  const bool was_on = d->D.wind != nullptr;
  if (!stage_force && !xmom_force && !ymom_force) {
    d->D.wind = nullptr;
    if (was_on) d->graph_valid = false;
    return SWK_OK;
  }
  if (n != d->N) return fail(SWK_ERR_ARG, "forcing arrays must have one entry per triangle");
  if (!d->wind) CKV(dalloc(&d->wind, 3 * d->NP));
  std::vector<double> w(3 * d->NP, 0.0);
  for (int64_t k = 0; k < d->N; k++) {
    const int64_t o = d->new2old[k];
    if (stage_force) w[k] = stage_force[o];
    if (xmom_force) w[d->NP + k] = xmom_force[o];
    if (ymom_force) w[2 * d->NP + k] = ymom_force[o];
  }
  CKV(upload(d->wind, w));
  d->D.wind = d->wind;
  if (!was_on) d->graph_valid = false;      // the pointer is a kernel argument
  return SWK_OK;
}

// Drop every registered Rate_operator (the host layer registers the current set again): an adapter
// that mirrors a reference domain rebuilds its operator list at every evolve() call.
extern "C" int swk_clear_rate_operators(swk_domain *d)
{
  if (!d) return fail(SWK_ERR_ARG, "handle is NULL");
  CK(cudaSetDevice(d->device));
  CKV(sync_check(d));
  for (auto &op : d->rate_ops) {
    if (op.d_rate_array) cudaFree(op.d_rate_array);
    if (op.d_indices) cudaFree(op.d_indices);
    if (op.d_partial) cudaFree(op.d_partial);
  }
  d->rate_ops.clear();
  d->graph_valid = false;
  return SWK_OK;
}

extern "C" int swk_set_rate(swk_domain *d, int op_id, double rate, double factor)
{
  if (!d || op_id < 0 || op_id >= (int)d->rate_ops.size()) return fail(SWK_ERR_ARG, "unknown rate operator");
  RateOp &op = d->rate_ops[op_id];
  op.rate = rate;
  op.factor = factor;
  if (!op.d_rate_array) op.all_nonneg = (rate >= 0.0) ? 1 : 0;
  CKV(vals_writable(d));
  d->h_vals[VALS_OPS + 2 * op_id] = rate;
  d->h_vals[VALS_OPS + 2 * op_id + 1] = factor;
  d->vals_dirty = true;
  if (!op.dynamic) d->graph_valid = false;     // a static operator's scalars are baked into the launches
  return SWK_OK;
}

// Mark an operator whose rate / factor are functions of time: its scalars are then read from the device
// value table at every step (refreshed by swk_set_rate), it is applied by its own kernel, and changing
// them does not invalidate the captured step.
extern "C" int swk_set_rate_dynamic(swk_domain *d, int op_id, int dynamic)
{
  if (!d || op_id < 0 || op_id >= (int)d->rate_ops.size()) return fail(SWK_ERR_ARG, "unknown rate operator");
  d->rate_ops[op_id].dynamic = dynamic != 0;
  d->graph_valid = false;
  return SWK_OK;
}

static int cell_ids_to_device(swk_domain *d, const int64_t *ids, int64_t n, int **d_ids)
{
  std::vector<int> v(n);
  for (int64_t j = 0; j < n; j++) {
    if (ids[j] < 0 || ids[j] >= d->N) return fail(SWK_ERR_ARG, "triangle id out of range");
    v[j] = d->old2new[ids[j]];
  }
  CKV(dalloc(d_ids, (size_t)n));
  return upload(*d_ids, v);
}

extern "C" int swk_gather_centroids(swk_domain *d, const int64_t *ids, int64_t n, double *out)
{
  if (!d || !ids || !out || n < 0) return fail(SWK_ERR_ARG, "bad argument");
  if (n == 0) return SWK_OK;
  CK(cudaSetDevice(d->device));
  int *d_ids = nullptr;
  CKV(cell_ids_to_device(d, ids, n, &d_ids));
  CKV(ensure_staging(d, (size_t)std::max<int64_t>(4 * n, 3LL * d->N)));
  LAUNCH(d, k_gather_cells, nblk(n), BLOCK, d->D, d_ids, (int)n, d->staging);
  cudaError_t e = cudaMemcpyAsync(out, d->staging, 4 * n * sizeof(double), cudaMemcpyDeviceToHost, d->stream);
  int rc = (e == cudaSuccess) ? sync_check(d) : fail(SWK_ERR_CUDA, cudaGetErrorString(e));
  cudaFree(d_ids);
  return rc;
}

extern "C" int swk_scatter_centroids(swk_domain *d, const int64_t *ids, int64_t n, const double *in)
{
  if (!d || !ids || !in || n < 0) return fail(SWK_ERR_ARG, "bad argument");
  if (n == 0) return SWK_OK;
  CK(cudaSetDevice(d->device));
  int *d_ids = nullptr;
  CKV(cell_ids_to_device(d, ids, n, &d_ids));
  CKV(ensure_staging(d, (size_t)std::max<int64_t>(4 * n, 3LL * d->N)));
  cudaError_t e = cudaMemcpyAsync(d->staging, in, 3 * n * sizeof(double), cudaMemcpyHostToDevice, d->stream);
  if (e == cudaSuccess) LAUNCH(d, k_scatter_cells, nblk(n), BLOCK, d->D, d_ids, (int)n, d->staging);
  int rc = (e == cudaSuccess) ? sync_check(d) : fail(SWK_ERR_CUDA, cudaGetErrorString(e));
  cudaFree(d_ids);
  return rc;
}

extern "C" int swk_scatter_bed(swk_domain *d, const int64_t *ids, int64_t n, const double *in)
{
  if (!d || !ids || !in || n < 0) return fail(SWK_ERR_ARG, "bad argument");
  if (n == 0) return SWK_OK;
  CK(cudaSetDevice(d->device));
  int *d_ids = nullptr;
  CKV(cell_ids_to_device(d, ids, n, &d_ids));
  CKV(ensure_staging(d, (size_t)std::max<int64_t>(4 * n, 3LL * d->N)));
  cudaError_t e = cudaMemcpyAsync(d->staging, in, n * sizeof(double), cudaMemcpyHostToDevice, d->stream);
  if (e == cudaSuccess) LAUNCH(d, k_scatter_bed, nblk(n), BLOCK, d->D, d_ids, (int)n, d->staging);
  int rc = (e == cudaSuccess) ? sync_check(d) : fail(SWK_ERR_CUDA, cudaGetErrorString(e));
  cudaFree(d_ids);
  return rc;
}

__global__ void k_add_volume(Clock *c, double v) { c->fractional_step_volume_integral += v; }

extern "C" int swk_add_fractional_step_volume(swk_domain *d, double volume)
{
  if (!d) return fail(SWK_ERR_ARG, "handle is NULL");
  CK(cudaSetDevice(d->device));
  LAUNCH(d, k_add_volume, 1, 1, d->d_clock, volume);       // in stream order, no host round trip
  return SWK_OK;
}

// ---- registered cell sets: what Inlet.fetch / Inlet.commit move every step (structures/inlet.py:135-330) ----
extern "C" int swk_register_cells(swk_domain *d, const int64_t *ids, int64_t n, int *set_id)
{
  if (!d || (!ids && n > 0) || n < 0 || !set_id) return fail(SWK_ERR_ARG, "bad argument");
  CK(cudaSetDevice(d->device));
  CellSet cs;
  cs.n = (int)n;
  if (n > 0) {
    CKV(cell_ids_to_device(d, ids, n, &cs.d_ids));
    CKV(dalloc(&cs.d_buf, 4 * (size_t)n));
    CK(cudaHostAlloc((void **)&cs.h_buf, 4 * (size_t)n * sizeof(double), cudaHostAllocDefault));
  }
  CK(cudaEventCreateWithFlags(&cs.ev, cudaEventDisableTiming));
  d->cell_sets.push_back(cs);
  *set_id = (int)d->cell_sets.size() - 1;
  return SWK_OK;
}

static int cell_set_writable(CellSet &cs)
{
  if (cs.inflight) {
    CK(cudaEventSynchronize(cs.ev));
    cs.inflight = false;
  }
  return SWK_OK;
}

// out: (n, 4) stage, xmomentum, ymomentum, elevation of the set's triangles, in the order they were registered
extern "C" int swk_gather_set(swk_domain *d, int set_id, double *out)
{
  if (!d || set_id < 0 || set_id >= (int)d->cell_sets.size() || !out) return fail(SWK_ERR_ARG, "bad argument");
  CellSet &cs = d->cell_sets[set_id];
  if (cs.n == 0) return SWK_OK;
  CK(cudaSetDevice(d->device));
  CKV(cell_set_writable(cs));
  LAUNCH(d, k_gather_cells, nblk(cs.n), BLOCK, d->D, cs.d_ids, cs.n, cs.d_buf);
  CK(cudaMemcpyAsync(cs.h_buf, cs.d_buf, 4 * (size_t)cs.n * sizeof(double), cudaMemcpyDeviceToHost, d->stream));
  CKV(sync_check(d));
  memcpy(out, cs.h_buf, 4 * (size_t)cs.n * sizeof(double));
  return SWK_OK;
}

// in: (n, 3) new stage, xmomentum, ymomentum; queued in stream order, returns without waiting
extern "C" int swk_scatter_set(swk_domain *d, int set_id, const double *in)
{
  if (!d || set_id < 0 || set_id >= (int)d->cell_sets.size() || !in) return fail(SWK_ERR_ARG, "bad argument");
  CellSet &cs = d->cell_sets[set_id];
  if (cs.n == 0) return SWK_OK;
  CK(cudaSetDevice(d->device));
  CKV(cell_set_writable(cs));
  memcpy(cs.h_buf, in, 3 * (size_t)cs.n * sizeof(double));
  CK(cudaMemcpyAsync(cs.d_buf, cs.h_buf, 3 * (size_t)cs.n * sizeof(double), cudaMemcpyHostToDevice, d->stream));
  CK(cudaEventRecord(cs.ev, d->stream));
  cs.inflight = true;
  LAUNCH(d, k_scatter_cells, nblk(cs.n), BLOCK, d->D, cs.d_ids, cs.n, cs.d_buf);
  return SWK_OK;
}



extern "C" int swk_set_local_ghost_copy(swk_domain *d, const int64_t *full_ids, const int64_t *ghost_ids, int64_t n)
{
  if (!d) return fail(SWK_ERR_ARG, "handle is NULL");
  CK(cudaSetDevice(d->device));
  if (d->d_ghost_full) cudaFree(d->d_ghost_full);
  if (d->d_ghost_ghost) cudaFree(d->d_ghost_ghost);
  d->d_ghost_full = d->d_ghost_ghost = nullptr;
  d->n_ghost_copy = 0;
  if (n <= 0) return SWK_OK;
  std::vector<int> f(n), g(n);
  for (int64_t j = 0; j < n; j++) {
    if (full_ids[j] < 0 || full_ids[j] >= d->N || ghost_ids[j] < 0 || ghost_ids[j] >= d->N)
      return fail(SWK_ERR_ARG, "ghost copy id out of range");
    f[j] = d->old2new[full_ids[j]];
    g[j] = d->old2new[ghost_ids[j]];
  }
  CKV(dalloc(&d->d_ghost_full, n)); CKV(upload(d->d_ghost_full, f));
  CKV(dalloc(&d->d_ghost_ghost, n)); CKV(upload(d->d_ghost_ghost, g));
  d->n_ghost_copy = (int)n;
  d->graph_valid = false;
  return SWK_OK;
}

// ----------------------------------------------------------------------------
// clock helpers
// ----------------------------------------------------------------------------
static int pull_clock(swk_domain *d)
{
  CK(cudaMemcpyAsync(d->h_clock, d->d_clock, sizeof(Clock), cudaMemcpyDeviceToHost, d->stream));
  return sync_check(d);
}

static int push_clock(swk_domain *d)
{
  CK(cudaMemcpyAsync(d->d_clock, d->h_clock, sizeof(Clock), cudaMemcpyHostToDevice, d->stream));
  return sync_check(d);
}

extern "C" int swk_set_time(swk_domain *d, double relative_time)
{
  if (!d) return fail(SWK_ERR_ARG, "handle is NULL");
  CK(cudaSetDevice(d->device));
  CKV(pull_clock(d));
  d->h_clock->time = relative_time;
  d->h_clock->step_start_time = relative_time;
  return push_clock(d);
}

extern "C" int swk_reset_yield_statistics(swk_domain *d)
{
  if (!d) return fail(SWK_ERR_ARG, "handle is NULL");
  CK(cudaSetDevice(d->device));
  CKV(pull_clock(d));
  d->h_clock->recorded_min_timestep = d->P.evolve_max_timestep;
  d->h_clock->recorded_max_timestep = d->P.evolve_min_timestep;
  d->h_clock->number_of_steps = 0;
  d->h_clock->number_of_first_order_steps = 0;
  CKV(push_clock(d));
  CK(cudaMemsetAsync(d->max_speed, 0, d->NP * sizeof(double), d->stream));
  return sync_check(d);
}

// ----------------------------------------------------------------------------
// launch helpers for the individual passes
// ----------------------------------------------------------------------------
static void launch_extrapolate(swk_domain *d, const Consts &K)
{
  TimedScope ts(d, 0);
  LAUNCH(d, k_extrapolate, nblk(d->N), BLOCK, d->D, K);
}

static int launch_boundary(swk_domain *d, int substep = 0)
{
  if (!d->capturing) CKV(push_segments(d));     // (table copies are not part of a captured step)
  if (d->M == 0) return SWK_OK;
  Segments S;
  S.b_cell = d->b_cell; S.b_edge = d->b_edge; S.b_seg = d->b_seg;
  S.seg_kind = d->d_seg_kind; S.seg_val = d->d_vals; S.substep = substep;
  S.seg_tab = d->d_seg_tab; S.seg_np = d->d_seg_np; S.b_point = d->d_b_point;
  if (d->seg_kind.empty()) return SWK_OK;
  LAUNCH(d, k_boundary_values, nblk(d->M), BLOCK, d->D, S, d->K, (int)d->P.centroid_transmissive_bc);
  return SWK_OK;
}

// Triangles the flux / update kernels work on: with a communicator the ghosts are refreshed from
// their owners after every update, so only the full triangles [0, n_full) are evaluated.
static int n_active(const swk_domain *d)
{
  return (d->comm && d->n_full > 0) ? d->n_full : (int)d->N;
}

static void launch_flux(swk_domain *d, int first, int write_speed)
{
  TimedScope ts(d, 1);
  const int n = n_active(d);
  if (d->D.bed_e_x) {        // per-call layer with edge beds / centroid heights given explicitly
    if (d->has_riverwalls) LAUNCH(d, (k_flux<true, true>), nblk(n), BLOCK, d->D, d->K, first, write_speed, 0, n);
    else LAUNCH(d, (k_flux<false, true>), nblk(n), BLOCK, d->D, d->K, first, write_speed, 0, n);
    return;
  }
  if (d->has_riverwalls) LAUNCH(d, k_flux<true>, nblk(n), BLOCK, d->D, d->K, first, write_speed, 0, n);
  else LAUNCH(d, k_flux<false>, nblk(n), BLOCK, d->D, d->K, first, write_speed, 0, n);
}

static void launch_bflux(swk_domain *d, int substep)
{
  LAUNCH(d, k_boundary_flux_sum, 1, 1024, d->D, substep);
}

// A single scalar, non-negative, all-cells Rate_operator is folded into the last update kernel
// of the step (no extra pass over the centroids); anything else runs as its own kernel.
static bool rain_is_fusable(const swk_domain *d)
{
  if (d->rate_ops.size() != 1) return false;
  const RateOp &op = d->rate_ops[0];
  return !op.d_indices && !op.d_rate_array && op.all_nonneg && !op.dynamic;
}

static UpdateArgs update_args(swk_domain *d, int do_backup, int do_saxpy, double a, double b, double divide_by,
                              bool last_of_step = false)
{
  UpdateArgs U;
  U.a = a; U.b = b; U.divide_by = divide_by;
  U.g = d->P.g;
  U.do_backup = do_backup; U.do_saxpy = do_saxpy;
  U.sloped = d->P.use_sloped_mannings ? 1 : 0;
  U.do_rain = 0; U.rain_rate = 0.0; U.rain_factor = 0.0;
  if (last_of_step && rain_is_fusable(d)) {
    U.do_rain = 1;
    U.rain_rate = d->rate_ops[0].rate;
    U.rain_factor = d->rate_ops[0].factor;
  }
  return U;
}

static void launch_rate_ops(swk_domain *d)
{
  if (!d->capturing) push_values(d);             // dynamic operators read {rate, factor} from the device table
  for (auto &op : d->rate_ops) {
    const double *rf = op.dynamic ? d->d_vals + VALS_OPS + 2 * (&op - &d->rate_ops[0]) : nullptr;
    LAUNCH(d, k_rate_operator, op.nblocks, BLOCK, d->D, op.rate, op.factor, op.d_rate_array, op.d_indices, op.n,
           op.all_nonneg, op.d_partial, rf);
    LAUNCH(d, k_rate_finish, 1, 1024, d->d_clock, op.d_partial, op.nblocks);
  }
}

// NCCL halo exchange on `stream` (parallel_generic_communications.py:159-248): pack the
// full_send lists, grouped ncclSend/ncclRecv with every peer, unpack into the ghosts
static int launch_exchange(swk_domain *d, cudaStream_t stream)
{
  if (!d->comm || d->peers.empty()) return SWK_OK;
  for (auto &pe : d->peers)
    if (pe.n_send > 0) {
      k_halo_pack<<<nblk(pe.n_send), BLOCK, 0, stream>>>(d->D, pe.d_send_ids, pe.n_send, pe.d_send_buf);
      d->launches++;
    }
  NK(g_nccl.GroupStart());
  for (auto &pe : d->peers) {
    if (pe.n_recv > 0) NK(g_nccl.Recv(pe.d_recv_buf, 3 * (size_t)pe.n_recv, ncclFloat64, pe.rank, d->comm, stream));
    if (pe.n_send > 0) NK(g_nccl.Send(pe.d_send_buf, 3 * (size_t)pe.n_send, ncclFloat64, pe.rank, d->comm, stream));
  }
  NK(g_nccl.GroupEnd());
  for (auto &pe : d->peers)
    if (pe.n_recv > 0) {
      k_halo_unpack<<<nblk(pe.n_recv), BLOCK, 0, stream>>>(d->D, pe.d_recv_ids, pe.n_recv, pe.d_recv_buf);
      d->launches++;
    }
  return SWK_OK;
}

// ghost update in-stream: local copy and/or NCCL halo exchange
static int launch_ghosts(swk_domain *d)
{
  if (d->n_ghost_copy > 0)
    LAUNCH(d, k_ghost_copy, nblk(d->n_ghost_copy), BLOCK, d->D, d->d_ghost_full, d->d_ghost_ghost, d->n_ghost_copy);
  return launch_exchange(d, d->stream);
}

// Run an update-type kernel f(k0, k1) over the active triangles and, if a ghost update follows it,
// overlap the halo exchange with the bulk of the kernel: the halo-source triangles [0, halo_front)
// are updated first, their values travel on the communication stream while [halo_front, n) is
// still being computed, and the main stream only waits for the unpack before the next
// extrapolation (north_star: "overlapped with interior flux computation").
template <typename F>
static int update_with_exchange(swk_domain *d, bool exchange_after, F f)
{
  const int n = n_active(d);
  const bool overlap = exchange_after && d->comm && !d->peers.empty() && d->n_full > 0 &&
                       d->halo_front > 0 && d->halo_front < n && d->overlap;
  if (!overlap) {
    f(0, n);
    if (exchange_after) CKV(launch_ghosts(d));
    return SWK_OK;
  }
  f(0, d->halo_front);
  CK(cudaEventRecord(d->ev_update, d->stream));
  CK(cudaStreamWaitEvent(d->comm_stream, d->ev_update, 0));
  CKV(launch_exchange(d, d->comm_stream));
  CK(cudaEventRecord(d->ev_halo, d->comm_stream));
  f(d->halo_front, n);
  if (d->n_ghost_copy > 0)
    LAUNCH(d, k_ghost_copy, nblk(d->n_ghost_copy), BLOCK, d->D, d->d_ghost_full, d->d_ghost_ghost, d->n_ghost_copy);
  CK(cudaStreamWaitEvent(d->stream, d->ev_halo, 0));
  return SWK_OK;
}

// global dt: min over ranks of the flux kernel's local min (parallel_generic_communications.py:67).
// The running minimum lives in the clock as the uint64 image of a positive double, which is
// monotone, so one in-place unsigned min-allreduce of that word is the exact double minimum.
static int launch_dt_allreduce(swk_domain *d)
{
  if (!d->comm || d->nranks == 1) return SWK_OK;
  void *word = (char *)d->d_clock + offsetof(Clock, dt_min_bits);
  NK(g_nccl.AllReduce(word, word, 1, ncclUint64, ncclMinOp, d->comm, d->stream));
  return SWK_OK;
}

// ---- one timestep, in two halves ------------------------------------------------------------
// first half : A, boundary values, B1 (+ boundary-flux sum), global dt, update_timestep
//              -> the clock holds the timestep
// second half: B2 [+ ghost update], the later RK substeps (A, boundary, fused B [+ ghost update]),
//              fractional steps, k_finish_step
// A resident run launches both back to back (one graph); a run with time-dependent boundary values or
// rates stops between the halves, once per step, so that the host can evaluate its functions at the
// substep times t + dt (and t + dt/2), which exist only now (swk_step_first / swk_step_rest).
static int launch_first_half(swk_domain *d)
{
  launch_extrapolate(d, d->K);
  CKV(launch_boundary(d, 0));
  launch_flux(d, 1, (int)d->P.track_max_speed);
  CKV(launch_dt_allreduce(d));
  LAUNCH(d, k_update_timestep, 1, 1024, d->D, d->TP, 1);
  return SWK_OK;
}

// B2 of substep 0 [+ ghost update]
static int launch_first_update(swk_domain *d, int do_backup, bool last_of_step, bool exchange_after)
{
  const UpdateArgs U = update_args(d, do_backup, 0, 1.0, 0.0, 1.0, last_of_step);
  return update_with_exchange(d, exchange_after, [&](int k0, int k1) {
    TimedScope ts(d, 2);
    LAUNCH(d, k_update, nblk(k1 - k0), BLOCK, d->D, d->K, U, -1.0, k0, k1);
  });
}

// a later RK substep with the RK combination folded in [+ ghost update]
static int launch_later_substep(swk_domain *d, int substep, double a, double b, double divide_by,
                                bool last_of_step, bool exchange_after)
{
  launch_extrapolate(d, d->K);
  CKV(launch_boundary(d, substep));
  const UpdateArgs U = update_args(d, 0, 1, a, b, divide_by, last_of_step);
  if (d->has_riverwalls) {
    // fused everywhere except on the wall triangles, which follow in k_update_list; the halo exchange
    // (if any) starts after both, so a wall triangle that is a halo source travels with its new state
    const int n = n_active(d);
    {
      TimedScope ts(d, 3);
      LAUNCH(d, k_flux_update<true>, nblk(n), BLOCK, d->D, d->K, U, 0, n);
    }
    // (ghost triangles are not evaluated under a communicator: only the wall triangles below n)
    const int nw = (int)(std::lower_bound(d->h_rw_list.begin(), d->h_rw_list.end(), n) - d->h_rw_list.begin());
    if (nw > 0) LAUNCH(d, k_update_list, nblk(nw), BLOCK, d->D, d->K, U, d->rw_list, nw);
    if (exchange_after) CKV(launch_ghosts(d));
  } else {
    CKV(update_with_exchange(d, exchange_after, [&](int k0, int k1) {
      TimedScope ts(d, 3);
      LAUNCH(d, k_flux_update<false>, nblk(k1 - k0), BLOCK, d->D, d->K, U, k0, k1);
    }));
  }
  if (!last_of_step) launch_bflux(d, substep);     // the last substep's sum rides in k_finish_step
  return SWK_OK;
}

static int launch_second_half(swk_domain *d)
{
  // step_start_time / dt_min_bits are (re)set by the previous k_finish_step (or by the host before
  // the first step of a call).  Ghost updates (generic_domain.py:1857, 2013-2015, 2096, 2135) ride with
  // the update kernel that precedes them; the one after the step can do so only when no separate
  // fractional-step kernel still changes the state.
  const int method = (int)d->P.timestepping_method;
  const bool ops_fused = d->rate_ops.empty() || rain_is_fusable(d);
  if (method == 1) {
    CKV(launch_first_update(d, 0, true, ops_fused));
  } else if (method == 2) {
    CKV(launch_first_update(d, 1, false, d->P.ghost_layer_width < 4));
    CKV(launch_later_substep(d, 1, 0.5, 0.5, 1.0, true, ops_fused));
  } else {
    CKV(launch_first_update(d, 1, false, true));
    CKV(launch_later_substep(d, 1, 0.25, 0.75, 1.0, false, true));
    CKV(launch_later_substep(d, 2, 2.0, 1.0, 3.0, true, ops_fused));
  }
  FusedRain R;
  R.on = 0; R.rate = 0.0; R.factor = 0.0; R.full_area = d->full_area;
  if (rain_is_fusable(d)) {                         // apply_fractional_steps (:1849)
    R.on = 1;
    R.rate = d->rate_ops[0].rate;
    R.factor = d->rate_ops[0].factor;
  } else {
    launch_rate_ops(d);
  }
  LAUNCH(d, k_finish_step, 1, 1024, d->D, d->TP, method == 1 ? -1 : method - 1, R);
  if (!ops_fused) CKV(launch_ghosts(d));            // :1857
  return SWK_OK;
}

// one full timestep = one iteration of _evolve_base's while loop (generic_domain.py:1835-1862)
static int launch_step(swk_domain *d)
{
  CKV(launch_first_half(d));
  return launch_second_half(d);
}

// Capture a launch sequence once (which: 0 whole step, 1 first half, 2 second half); every parameter
// baked into the kernel nodes (scalars, boundary kinds, static rain rate, halo lists) invalidates the
// graphs through d->graph_valid = false.  Boundary VALUES and dynamic rates live in device tables.
static int ensure_graph(swk_domain *d, int which)
{
  if (!d->graph_valid) {
    for (int w = 0; w < 3; w++) {
      if (d->graph_exec[w]) { cudaGraphExecDestroy(d->graph_exec[w]); d->graph_exec[w] = nullptr; }
      if (d->graph[w]) { cudaGraphDestroy(d->graph[w]); d->graph[w] = nullptr; }
    }
    d->graph_valid = true;
  }
  if (d->graph_exec[which]) return SWK_OK;
  CKV(push_segments(d));                       // host->device table copies must not be captured
  const int64_t l0 = d->launches;
  CK(cudaStreamBeginCapture(d->stream, cudaStreamCaptureModeThreadLocal));
  d->capturing = true;
  int rc = (which == 0) ? launch_step(d) : ((which == 1) ? launch_first_half(d) : launch_second_half(d));
  d->capturing = false;
  cudaError_t e = cudaStreamEndCapture(d->stream, &d->graph[which]);
  if (rc != SWK_OK) return rc;
  if (e != cudaSuccess) return fail(SWK_ERR_CUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e));
  d->launches_per_graph[which] = d->launches - l0;
  d->launches = l0;
  CK(cudaGraphInstantiate(&d->graph_exec[which], d->graph[which], 0));
  return SWK_OK;
}

// With a communicator the NCCL calls (dt min-allreduce on the main stream, halo send/recv on the
// communication stream, forked and joined by events) are captured into the step's graph as well,
// so that a multi-GPU step is one cudaGraphLaunch per rank: no host work between the 2-3 points
// per step at which the ranks wait for each other.
static int warm_nccl(swk_domain *d);
static bool graph_ok(const swk_domain *d)
{
  return d->use_graph && !d->timing;
}

// which: 0 whole step, 1 first half, 2 second half
static int run_part(swk_domain *d, int which)
{
  if (d->comm && !d->nccl_warm) CKV(warm_nccl(d));
  if (graph_ok(d)) {
    CKV(ensure_graph(d, which));
    CKV(push_segments(d));                     // value tables (async copy when dirty); cheap when clean
    CK(cudaGraphLaunch(d->graph_exec[which], d->stream));
    d->launches += d->launches_per_graph[which];
    return SWK_OK;
  }
  CKV(push_segments(d));
  return (which == 0) ? launch_step(d) : ((which == 1) ? launch_first_half(d) : launch_second_half(d));
}

static int run_one_step(swk_domain *d) { return run_part(d, 0); }

static int status_from_stop(int stop)
{
  switch (stop) {
    case -3: return fail(SWK_ERR_DENOMINATOR, "semi-implicit update: denominator <= 0 (quantity.c:806)");
    case -4: return fail(SWK_ERR_SMALLSTEP, "Too small timestep reached even after max_smallsteps steps of 1 order scheme");
    case -5: return fail(SWK_ERR_OVERSHOOT, "time overshot finaltime");
    default: return fail(SWK_ERR_CUDA, "device reported an unknown error state");
  }
}

static void fill_result(swk_domain *d, swk_evolve_result *r, int64_t launches0)
{
  if (!r) return;
  const Clock *c = d->h_clock;
  r->time = c->time;
  r->timestep = c->dt;
  r->flux_timestep = c->flux_dt;
  r->recorded_min_timestep = c->recorded_min_timestep;
  r->recorded_max_timestep = c->recorded_max_timestep;
  r->boundary_flux_integral = c->boundary_flux_integral;
  r->fractional_step_volume_integral = c->fractional_step_volume_integral;
  r->mass_error = c->mass_error;
  for (int i = 0; i < 3; i++) r->boundary_flux_sum[i] = c->boundary_flux_sum[i];
  r->number_of_steps = c->number_of_steps;
  r->number_of_first_order_steps = c->number_of_first_order_steps;
  r->total_steps = c->total_steps;
  r->negative_cells = c->negative_cells;
  r->stop_reason = (c->stop == 1) ? 1 : ((c->stop == 2) ? 2 : 0);
  r->kernel_launches = d->launches - launches0;
}

static int yield_epilogue(swk_domain *d, int reason, swk_evolve_result *result, int64_t launches0);

extern "C" int swk_evolve(swk_domain *d, double relative_yieldtime, double relative_finaltime, int64_t max_steps,
                          swk_evolve_result *result)
{
  if (!d) return fail(SWK_ERR_ARG, "handle is NULL");
  CK(cudaSetDevice(d->device));
  const int64_t launches0 = d->launches;
  CKV(pull_clock(d));
  Clock *c = d->h_clock;
  if (c->stop < 0) return status_from_stop(c->stop);
  c->yieldtime = relative_yieldtime;
  c->finaltime = relative_finaltime;
  c->step_budget = (max_steps > 0) ? c->total_steps + max_steps : 0;
  c->stop = 0;
  c->step_start_time = c->time;
  c->dt_min_bits = 0x54B249AD2594C37DULL;          // bits of 1.0e+100
  CKV(push_clock(d));

  int64_t batch = 1;
  for (;;) {
    for (int64_t b = 0; b < batch; b++) CKV(run_one_step(d));
    CKV(pull_clock(d));
    if (c->stop != 0) break;
    // estimate how many more steps fit before the next stop (the device clips dt itself;
    // surplus launches see clock->stop and return immediately)
    double target = relative_yieldtime;
    if (relative_finaltime >= 0.0 && relative_finaltime < target) target = relative_finaltime;
    double est = (c->dt > 0.0) ? (target - c->time) / c->dt : 1.0;
    if (!(est >= 1.0)) est = 1.0;
    batch = (int64_t)std::min(est + 1.0, 64.0);
    if (c->step_budget > 0) batch = std::min<int64_t>(batch, std::max<int64_t>(1, c->step_budget - c->total_steps));
  }
  if (c->stop < 0) return status_from_stop(c->stop);
  return yield_epilogue(d, c->stop, result, launches0);
}

// ---- host-paced run: one host visit per timestep, between its two halves -------------------------
// For boundary values / rates that are Python functions of time.  Everything stays resident and fused
// (the halves replay as CUDA graphs); the host learns (t, dt) after the first half - one small D2H read -
// evaluates its functions at the substep times, writes the value tables (swk_set_boundary_values_substep,
// swk_set_rate; one small async H2D copy) and releases the second half.  Sequence per evolve segment:
//   swk_step_begin;  { swk_step_first -> stop? break : set values; swk_step_rest }*;  swk_step_end
extern "C" int swk_step_begin(swk_domain *d, double relative_yieldtime, double relative_finaltime)
{
  if (!d) return fail(SWK_ERR_ARG, "handle is NULL");
  CK(cudaSetDevice(d->device));
  d->launches_mark = d->launches;
  CKV(pull_clock(d));
  Clock *c = d->h_clock;
  if (c->stop < 0) return status_from_stop(c->stop);
  c->yieldtime = relative_yieldtime;
  c->finaltime = relative_finaltime;
  c->step_budget = 0;
  c->stop = 0;
  c->step_start_time = c->time;
  c->dt_min_bits = 0x54B249AD2594C37DULL;          // bits of 1.0e+100
  return push_clock(d);
}

extern "C" int swk_step_first(swk_domain *d, swk_evolve_result *result)
{
  if (!d) return fail(SWK_ERR_ARG, "handle is NULL");
  CK(cudaSetDevice(d->device));
  CKV(run_part(d, 1));
  CKV(pull_clock(d));
  Clock *c = d->h_clock;
  if (c->stop < 0) return status_from_stop(c->stop);
  fill_result(d, result, d->launches_mark);
  if (result) result->time = (c->stop != 0) ? c->time : c->step_start_time;
  return SWK_OK;
}

extern "C" int swk_step_rest(swk_domain *d)
{
  if (!d) return fail(SWK_ERR_ARG, "handle is NULL");
  CK(cudaSetDevice(d->device));
  return run_part(d, 2);
}

// the yield: distribute_to_vertices_and_edges + update_boundary (generic_domain.py:1884-1885, 1899-1900)
static int yield_epilogue(swk_domain *d, int reason, swk_evolve_result *result, int64_t launches0)
{
  Clock *c = d->h_clock;
  if (reason == 1 || reason == 2) {
    c->stop = 0;
    CKV(push_clock(d));
    launch_extrapolate(d, d->K);
    LAUNCH(d, k_materialize_centroids, nblk(d->N), BLOCK, d->D, d->K, 0);
    CKV(launch_boundary(d));
    CKV(pull_clock(d));
  }
  c->stop = reason;
  fill_result(d, result, launches0);
  c->stop = 0;
  return push_clock(d);
}

extern "C" int swk_step_end(swk_domain *d, swk_evolve_result *result)
{
  if (!d) return fail(SWK_ERR_ARG, "handle is NULL");
  CK(cudaSetDevice(d->device));
  CKV(pull_clock(d));
  Clock *c = d->h_clock;
  if (c->stop < 0) return status_from_stop(c->stop);
  return yield_epilogue(d, c->stop, result, d->launches_mark);
}

extern "C" int swk_run_steps(swk_domain *d, int64_t n_steps, int per_kernel, float *elapsed_ms)
{
  if (!d || n_steps < 0) return fail(SWK_ERR_ARG, "bad argument");
  CK(cudaSetDevice(d->device));
  CKV(pull_clock(d));
  Clock *c = d->h_clock;
  if (c->stop < 0) return status_from_stop(c->stop);
  c->yieldtime = 1.0e300;
  c->finaltime = -1.0;
  c->step_budget = 0;
  c->stop = 0;
  c->step_start_time = c->time;
  c->dt_min_bits = 0x54B249AD2594C37DULL;          // bits of 1.0e+100
  CKV(push_clock(d));
  d->timing = false;
  if (per_kernel) {
    // up to 3 extrapolations + flux + 2 x (update | fused) launches per substep with the halo overlap
    const size_t want = (size_t)std::min<int64_t>(n_steps, 512) * 12 * 2;
    while (d->ev_pool.size() < want) {
      cudaEvent_t e;
      CK(cudaEventCreate(&e));
      d->ev_pool.push_back(e);
    }
    d->ev_kind.assign(d->ev_pool.size() / 2, 0);
    d->ev_used = 0;
    for (int i = 0; i < 4; i++) { d->k_ms[i] = 0; d->k_n[i] = 0; }
    d->timing = true;
  }
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  CK(cudaStreamSynchronize(d->stream));
  CK(cudaEventRecord(e0, d->stream));
  int rc = SWK_OK;
  for (int64_t s = 0; s < n_steps && rc == SWK_OK; s++) rc = run_one_step(d);
  CK(cudaEventRecord(e1, d->stream));
  d->timing = false;
  CK(cudaEventSynchronize(e1));
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (elapsed_ms) *elapsed_ms = ms;
  if (per_kernel) {
    for (size_t j = 0; j + 1 < d->ev_used; j += 2) {
      float kms = 0.f;
      if (cudaEventElapsedTime(&kms, d->ev_pool[j], d->ev_pool[j + 1]) == cudaSuccess) {
        d->k_ms[d->ev_kind[j / 2]] += kms;
        d->k_n[d->ev_kind[j / 2]] += 1;
      }
    }
  }
  if (rc != SWK_OK) return rc;
  CKV(pull_clock(d));
  if (c->stop < 0) return status_from_stop(c->stop);
  return SWK_OK;
}

extern "C" int swk_kernel_timing(swk_domain *d, double total_ms[4], int64_t launches[4])
{
  if (!d || !total_ms || !launches) return fail(SWK_ERR_ARG, "NULL argument");
  for (int i = 0; i < 4; i++) { total_ms[i] = d->k_ms[i]; launches[i] = d->k_n[i]; }
  return SWK_OK;
}

extern "C" int swk_stream(swk_domain *d, void **cuda_stream_out)
{
  if (!d || !cuda_stream_out) return fail(SWK_ERR_ARG, "NULL argument");
  *cuda_stream_out = (void *)d->stream;
  return SWK_OK;
}

extern "C" int swk_synchronize(swk_domain *d)
{
  if (!d) return fail(SWK_ERR_ARG, "handle is NULL");
  CK(cudaSetDevice(d->device));
  return sync_check(d);
}

extern "C" int swk_pin_host_buffer(swk_domain *d, void *host, size_t bytes)
{
  if (!d || !host) return fail(SWK_ERR_ARG, "NULL argument");
  CK(cudaSetDevice(d->device));
  cudaError_t e = cudaHostRegister(host, bytes, cudaHostRegisterDefault);
  if (e == cudaErrorHostMemoryAlreadyRegistered) { cudaGetLastError(); return SWK_OK; }
  if (e != cudaSuccess) return fail(SWK_ERR_CUDA, std::string("cudaHostRegister: ") + cudaGetErrorString(e));
  return SWK_OK;
}

extern "C" int swk_unpin_host_buffer(swk_domain *d, void *host)
{
  if (!d || !host) return fail(SWK_ERR_ARG, "NULL argument");
  CK(cudaSetDevice(d->device));
  cudaError_t e = cudaHostUnregister(host);
  if (e != cudaSuccess) { cudaGetLastError(); }
  return SWK_OK;
}

extern "C" int swk_kernel_launch_count(swk_domain *d, int64_t *count)
{
  if (!d || !count) return fail(SWK_ERR_ARG, "NULL argument");
  *count = d->launches;
  return SWK_OK;
}

extern "C" int swk_bytes_per_triangle_step(swk_domain *d, double *algorithmic, double *layout)
{
  if (!d) return fail(SWK_ERR_ARG, "handle is NULL");
  const int method = (int)d->P.timestepping_method;
  // DE0 556 B, DE1 1076 B, DE2 1572 B per triangle-step
  const double alg[4] = {0, 556.0, 1076.0, 1572.0};
  // this library's records: pass A 241, B1 264, B2 121, fused B 297
  const double A = 32 + 16 + 96 + 96 + 1, B1 = 96 + 16 + 96 + 32 + 24, B2 = 24 + 32 + 8 + 1 + 32, B = 96 + 16 + 96 + 32 + 8 + 1 + 24 + 32;
  double lay = A + B1 + B2 + (method >= 2 ? 24 : 0);
  for (int s = 1; s < method; s++) lay += A + B;
  if (algorithmic) *algorithmic = alg[method];
  if (layout) *layout = lay;
  return SWK_OK;
}

// ----------------------------------------------------------------------------
// individual steps on resident data (host-stepped mode and per-call layer)
// ----------------------------------------------------------------------------
extern "C" int swk_protect(swk_domain *d, double *mass_error)
{
  if (!d) return fail(SWK_ERR_ARG, "handle is NULL");
  CK(cudaSetDevice(d->device));
  CKV(pull_clock(d));
  const double before = d->h_clock->mass_error;
  LAUNCH(d, k_protect_mass, nblk(d->N), BLOCK, d->D, d->K);
  LAUNCH(d, k_materialize_centroids, nblk(d->N), BLOCK, d->D, d->K, 1);
  CKV(pull_clock(d));
  if (mass_error) *mass_error = d->h_clock->mass_error - before;
  return SWK_OK;
}

static int extrapolate_impl(swk_domain *d, int protect)
{
  Consts K = d->K;
  K.protect = protect;
  launch_extrapolate(d, K);
  LAUNCH(d, k_materialize_centroids, nblk(d->N), BLOCK, d->D, K, 0);
  return sync_check(d);
}

extern "C" int swk_extrapolate_second_order_edge_sw(swk_domain *d)
{
  if (!d) return fail(SWK_ERR_ARG, "handle is NULL");
  CK(cudaSetDevice(d->device));
  return extrapolate_impl(d, 0);
}

extern "C" int swk_distribute_to_vertices_and_edges(swk_domain *d)
{
  if (!d) return fail(SWK_ERR_ARG, "handle is NULL");
  CK(cudaSetDevice(d->device));
  return extrapolate_impl(d, 1);
}

extern "C" int swk_update_boundary(swk_domain *d)
{
  if (!d) return fail(SWK_ERR_ARG, "handle is NULL");
  CK(cudaSetDevice(d->device));
  CKV(launch_boundary(d));
  return sync_check(d);
}

extern "C" int swk_compute_fluxes(swk_domain *d, int substep, double *flux_timestep)
{
  if (!d) return fail(SWK_ERR_ARG, "handle is NULL");
  if (substep < 0 || substep > 2) return fail(SWK_ERR_ARG, "substep must be 0, 1 or 2");
  CK(cudaSetDevice(d->device));
  const int first = (substep == 0);
  if (first) LAUNCH(d, k_begin_step, 1, 1, d->d_clock);
  launch_flux(d, first, 1);
  launch_bflux(d, substep);
  if (first) CKV(launch_dt_allreduce(d));
  CKV(pull_clock(d));
  if (flux_timestep) *flux_timestep = first ? u2d_host(d->h_clock->dt_min_bits) : d->P.evolve_max_timestep;
  return SWK_OK;
}

extern "C" int swk_update_conserved_quantities(swk_domain *d, double timestep, int64_t *num_negative)
{
  if (!d) return fail(SWK_ERR_ARG, "handle is NULL");
  CK(cudaSetDevice(d->device));
  CKV(pull_clock(d));
  const long long before = d->h_clock->negative_cells;
  LAUNCH(d, k_update, nblk(d->N), BLOCK, d->D, d->K, update_args(d, 0, 0, 1.0, 0.0, 1.0), timestep, 0, (int)d->N);
  CKV(pull_clock(d));
  if (d->h_clock->stop < 0) {
    const int st = d->h_clock->stop;
    d->h_clock->stop = 0;
    CKV(push_clock(d));
    return status_from_stop(st);
  }
  if (num_negative) *num_negative = d->h_clock->negative_cells - before;
  return SWK_OK;
}

extern "C" int swk_backup_conserved_quantities(swk_domain *d)
{
  if (!d) return fail(SWK_ERR_ARG, "handle is NULL");
  CK(cudaSetDevice(d->device));
  LAUNCH(d, k_backup, nblk(d->N), BLOCK, d->D);
  return sync_check(d);
}

extern "C" int swk_saxpy_conserved_quantities(swk_domain *d, double a, double b, double divide_by)
{
  if (!d) return fail(SWK_ERR_ARG, "handle is NULL");
  if (divide_by == 0.0) return fail(SWK_ERR_ARG, "divide_by must be non-zero");
  CK(cudaSetDevice(d->device));
  LAUNCH(d, k_saxpy, nblk(d->N), BLOCK, d->D, a, b, divide_by);
  return sync_check(d);
}

extern "C" int swk_apply_fractional_steps(swk_domain *d, double timestep)
{
  if (!d) return fail(SWK_ERR_ARG, "handle is NULL");
  CK(cudaSetDevice(d->device));
  LAUNCH(d, k_set_dt, 1, 1, d->d_clock, timestep);
  LAUNCH(d, k_bfi_update, 1, 1, d->d_clock, d->TP);
  launch_rate_ops(d);
  return sync_check(d);
}

extern "C" int swk_get_statistics(swk_domain *d, swk_evolve_result *result)
{
  if (!d || !result) return fail(SWK_ERR_ARG, "NULL argument");
  CK(cudaSetDevice(d->device));
  CKV(pull_clock(d));
  fill_result(d, result, d->launches);
  return SWK_OK;
}

extern "C" int swk_update_ghosts(swk_domain *d)
{
  if (!d) return fail(SWK_ERR_ARG, "handle is NULL");
  CK(cudaSetDevice(d->device));
  CKV(launch_ghosts(d));
  return sync_check(d);
}

// update_ghosts queued in stream order (the next synchronising call covers it)
extern "C" int swk_update_ghosts_async(swk_domain *d)
{
  if (!d) return fail(SWK_ERR_ARG, "handle is NULL");
  CK(cudaSetDevice(d->device));
  if (d->comm && !d->nccl_warm) CKV(warm_nccl(d));
  return launch_ghosts(d);
}

// ----------------------------------------------------------------------------
// multi-GPU plumbing
// ----------------------------------------------------------------------------
extern "C" int swk_nccl_unique_id(void *id128)
{
  if (!id128) return fail(SWK_ERR_ARG, "id128 is NULL");
  CKV(load_nccl());
  ncclUniqueId id;
  NK(g_nccl.GetUniqueId(&id));
  memcpy(id128, &id, sizeof(id));
  return SWK_OK;
}

extern "C" int swk_comm_create(const void *id128, int rank, int nranks, int device, swk_comm **out)
{
  if (!id128 || !out) return fail(SWK_ERR_ARG, "NULL argument");
  *out = nullptr;
  CKV(select_device(device));
  CKV(load_nccl());
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  swk_comm *c = new swk_comm();
  c->rank = rank; c->nranks = nranks; c->device = device;
  ncclResult_t r = g_nccl.CommInitRank(&c->nc, nranks, id, rank);
  if (r != 0) {
    delete c;
    return fail(SWK_ERR_NCCL, std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(r));
  }
  CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  *out = c;
  return SWK_OK;
}

extern "C" int swk_comm_destroy(swk_comm *c)
{
  if (!c) return SWK_OK;
  cudaSetDevice(c->device);
  if (c->stream) { cudaStreamSynchronize(c->stream); cudaStreamDestroy(c->stream); }
  if (c->d_buf) cudaFree(c->d_buf);
  if (c->nc && g_nccl.CommDestroy) g_nccl.CommDestroy(c->nc);
  delete c;
  return SWK_OK;
}

// Host-level collective on a small host buffer (barriers, timing maxima, the bit-exact merges of
// the structure operators): staged through a device scratch, synchronous.  The caller must not
// have a device time loop of an attached domain in flight (swk_evolve / swk_run_steps return
// synchronised, so this holds whenever the host has control).
extern "C" int swk_comm_allreduce(swk_comm *c, void *host_buf, int64_t n, int dtype, int op)
{
  if (!c || (!host_buf && n > 0) || n < 0) return fail(SWK_ERR_ARG, "bad argument");
  if (dtype != SWK_F64 && dtype != SWK_I64) return fail(SWK_ERR_ARG, "dtype must be SWK_F64 or SWK_I64");
  if (op < SWK_SUM || op > SWK_MAX) return fail(SWK_ERR_ARG, "op must be SWK_SUM, SWK_MIN or SWK_MAX");
  CK(cudaSetDevice(c->device));
  const size_t bytes = (size_t)std::max<int64_t>(n, 1) * 8;
  if (bytes > c->cap) {
    if (c->d_buf) cudaFree(c->d_buf);
    c->d_buf = nullptr;
    c->cap = 0;
    CK(cudaMalloc(&c->d_buf, bytes));
    c->cap = bytes;
  }
  if (n == 0) return SWK_OK;
  CK(cudaMemcpyAsync(c->d_buf, host_buf, (size_t)n * 8, cudaMemcpyHostToDevice, c->stream));
  const int nccl_op = (op == SWK_SUM) ? ncclSumOp : ((op == SWK_MIN) ? ncclMinOp : ncclMaxOp);
  NK(g_nccl.AllReduce(c->d_buf, c->d_buf, (size_t)n, dtype == SWK_F64 ? ncclFloat64 : ncclInt64, nccl_op, c->nc,
                      c->stream));
  CK(cudaMemcpyAsync(host_buf, c->d_buf, (size_t)n * 8, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return SWK_OK;
}

extern "C" int swk_comm_attach(swk_domain *d, swk_comm *c)
{
  if (!d || !c) return fail(SWK_ERR_ARG, "NULL argument");
  if (d->comm) return fail(SWK_ERR_ARG, "the domain already has a communicator");
  if (c->device != d->device) return fail(SWK_ERR_ARG, "communicator and domain live on different devices");
  CK(cudaSetDevice(d->device));
  d->comm = c->nc;
  d->comm_obj = c;
  d->rank = c->rank;
  d->nranks = c->nranks;
  d->graph_valid = false;
  d->nccl_warm = false;
  CK(cudaStreamCreateWithFlags(&d->comm_stream, cudaStreamNonBlocking));
  CK(cudaEventCreateWithFlags(&d->ev_update, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&d->ev_halo, cudaEventDisableTiming));
  {
    const char *env = getenv("SWK_NO_OVERLAP");
    d->overlap = !(env && env[0] == '1');
  }
  return SWK_OK;
}

extern "C" int swk_comm_init(swk_domain *d, const void *id128, int rank, int nranks)
{
  if (!d || !id128) return fail(SWK_ERR_ARG, "NULL argument");
  swk_comm *c = nullptr;
  CKV(swk_comm_create(id128, rank, nranks, d->device, &c));
  int r = swk_comm_attach(d, c);
  if (r != SWK_OK) { swk_comm_destroy(c); return r; }
  d->comm_owned = true;
  return SWK_OK;
}

// First use of a peer connection makes NCCL set the transport up (host work, allocations): do it
// once outside any stream capture - one empty-handed round of the step's collectives (the halo
// buffers travel without pack / unpack, so no state is touched).
static int warm_nccl(swk_domain *d)
{
  d->nccl_warm = true;
  if (!d->comm) return SWK_OK;
  if (!d->peers.empty()) {
    NK(g_nccl.GroupStart());
    for (auto &pe : d->peers) {
      if (pe.n_recv > 0) NK(g_nccl.Recv(pe.d_recv_buf, 3 * (size_t)pe.n_recv, ncclFloat64, pe.rank, d->comm, d->comm_stream));
      if (pe.n_send > 0) NK(g_nccl.Send(pe.d_send_buf, 3 * (size_t)pe.n_send, ncclFloat64, pe.rank, d->comm, d->comm_stream));
    }
    NK(g_nccl.GroupEnd());
    CK(cudaStreamSynchronize(d->comm_stream));
  }
  if (d->nranks > 1) {
    unsigned long long *w = nullptr;
    CK(cudaMalloc((void **)&w, 8));
    CK(cudaMemsetAsync(w, 0, 8, d->stream));
    NK(g_nccl.AllReduce(w, w, 1, ncclUint64, ncclMinOp, d->comm, d->stream));
    CK(cudaStreamSynchronize(d->stream));
    cudaFree(w);
  }
  return SWK_OK;
}

extern "C" int swk_set_halo(swk_domain *d, int n_peers, const int *peer_ranks, const int64_t *send_counts,
                            const int64_t *const *send_ids, const int64_t *recv_counts,
                            const int64_t *const *recv_ids)
{
  if (!d) return fail(SWK_ERR_ARG, "handle is NULL");
  CK(cudaSetDevice(d->device));
  for (auto &pe : d->peers) {
    if (pe.d_send_ids) cudaFree(pe.d_send_ids);
    if (pe.d_recv_ids) cudaFree(pe.d_recv_ids);
    if (pe.d_send_buf) cudaFree(pe.d_send_buf);
    if (pe.d_recv_buf) cudaFree(pe.d_recv_buf);
  }
  d->peers.clear();
  d->halo_front = 0;
  d->graph_valid = false;
  d->nccl_warm = false;
  for (int q = 0; q < n_peers; q++) {
    Peer pe;
    pe.rank = peer_ranks[q];
    pe.n_send = (int)send_counts[q];
    pe.n_recv = (int)recv_counts[q];
    std::vector<int> s(pe.n_send), r(pe.n_recv);
    for (int j = 0; j < pe.n_send; j++) {
      if (send_ids[q][j] < 0 || send_ids[q][j] >= d->N) return fail(SWK_ERR_ARG, "halo send id out of range");
      s[j] = d->old2new[send_ids[q][j]];
    }
    for (int j = 0; j < pe.n_recv; j++) {
      if (recv_ids[q][j] < 0 || recv_ids[q][j] >= d->N) return fail(SWK_ERR_ARG, "halo recv id out of range");
      r[j] = d->old2new[recv_ids[q][j]];
    }
    CKV(dalloc(&pe.d_send_ids, s.size())); CKV(upload(pe.d_send_ids, s));
    CKV(dalloc(&pe.d_recv_ids, r.size())); CKV(upload(pe.d_recv_ids, r));
    CKV(dalloc(&pe.d_send_buf, 3 * s.size()));
    CKV(dalloc(&pe.d_recv_buf, 3 * r.size()));
    for (int v : s) d->halo_front = std::max(d->halo_front, v + 1);
    d->peers.push_back(pe);
  }
  if (d->n_full > 0 && d->halo_front > d->n_full) d->halo_front = 0;   // a send id among the ghosts: no overlap
  return SWK_OK;
}

// ============================================================================
// (2) PER-CALL LAYER: host arrays in, host arrays out
// ============================================================================
extern "C" int swk_call_open(const swk_host_view *v, int device, swk_domain **out)
{
  if (!v) return fail(SWK_ERR_ARG, "view is NULL");
  CKV(swk_create(&v->mesh, &v->params, device, out));
  (*out)->per_call = true;
  (*out)->K.protect = 0;      // each call is exactly one reference function
  return SWK_OK;
}

extern "C" int swk_call_close(swk_domain *d) { return swk_destroy(d); }

static int refresh_per_call(swk_domain *d, const swk_host_view *v)
{
  if (!d || !v) return fail(SWK_ERR_ARG, "NULL argument");
  if (v->mesh.number_of_elements != d->N || v->mesh.boundary_length != d->M)
    return fail(SWK_ERR_ARG, "host view does not match the opened context");
  CK(cudaSetDevice(d->device));
  CKV(check_params(&v->params));
  d->P = v->params;
  make_consts(d);
  d->K.protect = 0;
  return SWK_OK;
}

extern "C" int swk_call_compute_fluxes_ext_central(swk_domain *d, const swk_host_view *v, double timestep,
                                                   int substep, double *flux_timestep)
{
  CKV(refresh_per_call(d, v));
  if (substep < 0 || substep >= v->params.timestepping_method) return fail(SWK_ERR_ARG, "substep out of range");
  const int64_t N = d->N, M = d->M;
  // The device edge record carries (stage, height, xmom, ymom); bed_edge = stage_edge - height_edge and
  // height_centroid = max(stage - bed, 0) are what extrapolate always leaves behind
  // (sw_domain_openmp.c:1376-1378, 1863-1865) and what the kernels recompute.  The reference's entry point
  // takes ANY arrays: when the caller's do not satisfy those identities, they are uploaded as they are and
  // the flux kernel variant that reads them is launched.
  bool consistent = true;
  for (int64_t j = 0; j < 3 * N && consistent; j++)
    if (v->bed_edge_values[j] != v->stage_edge_values[j] - v->height_edge_values[j]) consistent = false;
  for (int64_t k = 0; k < N && consistent; k++)
    if (v->height_centroid_values[k] != fmax(v->stage_centroid_values[k] - v->bed_centroid_values[k], 0.0)) consistent = false;
  d->D.bed_e_x = nullptr;
  d->D.hc_x = nullptr;
  if (!consistent) {
    const int64_t NP = d->NP;
    if (!d->xbed) CKV(dalloc(&d->xbed, 4 * NP));
    std::vector<double> x(4 * NP, 0.0);
    for (int64_t k = 0; k < N; k++) {
      const int64_t o = d->new2old[k];
      for (int i = 0; i < 3; i++) x[i * NP + k] = v->bed_edge_values[3 * o + i];
      x[3 * NP + k] = v->height_centroid_values[o];
    }
    CKV(upload(d->xbed, x));
    d->D.bed_e_x = d->xbed;
    d->D.hc_x = d->xbed + 3 * NP;
  }
  CKV(swk_set_quantity(d, SWK_Q_STAGE_C, v->stage_centroid_values, N));
  CKV(swk_set_quantity(d, SWK_Q_ELEVATION_C, v->bed_centroid_values, N));
  CKV(swk_set_quantity(d, SWK_Q_STAGE_E, v->stage_edge_values, 3 * N));
  CKV(swk_set_quantity(d, SWK_Q_HEIGHT_E, v->height_edge_values, 3 * N));
  CKV(swk_set_quantity(d, SWK_Q_XMOM_E, v->xmom_edge_values, 3 * N));
  CKV(swk_set_quantity(d, SWK_Q_YMOM_E, v->ymom_edge_values, 3 * N));
  if (M > 0) {
    CKV(swk_set_quantity(d, SWK_Q_STAGE_B, v->stage_boundary_values, M));
    CKV(swk_set_quantity(d, SWK_Q_XMOM_B, v->xmom_boundary_values, M));
    CKV(swk_set_quantity(d, SWK_Q_YMOM_B, v->ymom_boundary_values, M));
  }
  double ft = 0.0;
  CKV(swk_compute_fluxes(d, substep, &ft));
  CKV(swk_get_quantity(d, SWK_Q_STAGE_EU, v->stage_explicit_update, N));
  CKV(swk_get_quantity(d, SWK_Q_XMOM_EU, v->xmom_explicit_update, N));
  CKV(swk_get_quantity(d, SWK_Q_YMOM_EU, v->ymom_explicit_update, N));
  if (substep == 0 && v->max_speed) CKV(swk_get_quantity(d, SWK_Q_MAX_SPEED, v->max_speed, N));
  if (v->boundary_flux_sum) v->boundary_flux_sum[substep] = d->h_clock->boundary_flux_sum[substep];
  if (flux_timestep) *flux_timestep = (substep == 0) ? ft : timestep;   // :768-771
  return SWK_OK;
}

extern "C" int swk_call_extrapolate_second_order_edge_sw(swk_domain *d, const swk_host_view *v)
{
  CKV(refresh_per_call(d, v));
  const int64_t N = d->N;
  CKV(swk_set_quantity(d, SWK_Q_STAGE_C, v->stage_centroid_values, N));
  CKV(swk_set_quantity(d, SWK_Q_XMOM_C, v->xmom_centroid_values, N));
  CKV(swk_set_quantity(d, SWK_Q_YMOM_C, v->ymom_centroid_values, N));
  CKV(swk_set_quantity(d, SWK_Q_ELEVATION_C, v->bed_centroid_values, N));
  CKV(extrapolate_impl(d, 0));
  CKV(swk_get_quantity(d, SWK_Q_XMOM_C, v->xmom_centroid_values, N));
  CKV(swk_get_quantity(d, SWK_Q_YMOM_C, v->ymom_centroid_values, N));
  if (v->height_centroid_values) CKV(swk_get_quantity(d, SWK_Q_HEIGHT_C, v->height_centroid_values, N));
  struct { int q; double *p; } outs[] = {
      {SWK_Q_STAGE_E, v->stage_edge_values}, {SWK_Q_XMOM_E, v->xmom_edge_values},
      {SWK_Q_YMOM_E, v->ymom_edge_values}, {SWK_Q_HEIGHT_E, v->height_edge_values},
      {SWK_Q_ELEVATION_E, v->bed_edge_values}, {SWK_Q_STAGE_V, v->stage_vertex_values},
      {SWK_Q_XMOM_V, v->xmom_vertex_values}, {SWK_Q_YMOM_V, v->ymom_vertex_values},
      {SWK_Q_HEIGHT_V, v->height_vertex_values}, {SWK_Q_ELEVATION_V, v->bed_vertex_values}};
  for (auto &o : outs)
    if (o.p) CKV(swk_get_quantity(d, o.q, o.p, 3 * N));
  return SWK_OK;
}

extern "C" int swk_call_protect_new(swk_domain *d, const swk_host_view *v, double *mass_error)
{
  CKV(refresh_per_call(d, v));
  const int64_t N = d->N;
  CKV(swk_set_quantity(d, SWK_Q_STAGE_C, v->stage_centroid_values, N));
  CKV(swk_set_quantity(d, SWK_Q_XMOM_C, v->xmom_centroid_values, N));
  CKV(swk_set_quantity(d, SWK_Q_ELEVATION_C, v->bed_centroid_values, N));
  // vertex fix-up of lifted cells (sw_domain_openmp.c:1158-1160) uses the pre-call stage
  if (v->stage_vertex_values)
    for (int64_t k = 0; k < N; k++)
      if (v->stage_centroid_values[k] < v->bed_centroid_values[k])
        v->stage_vertex_values[3 * k] = v->stage_vertex_values[3 * k + 1] = v->stage_vertex_values[3 * k + 2] =
            v->bed_centroid_values[k];
  CKV(swk_protect(d, mass_error));
  CKV(swk_get_quantity(d, SWK_Q_STAGE_C, v->stage_centroid_values, N));
  CKV(swk_get_quantity(d, SWK_Q_XMOM_C, v->xmom_centroid_values, N));
  return SWK_OK;
}

extern "C" int swk_call_fix_negative_cells(swk_domain *d, const swk_host_view *v, int64_t *count)
{
  CKV(refresh_per_call(d, v));
  const int64_t N = d->N;
  CKV(swk_set_quantity(d, SWK_Q_STAGE_C, v->stage_centroid_values, N));
  CKV(swk_set_quantity(d, SWK_Q_XMOM_C, v->xmom_centroid_values, N));
  CKV(swk_set_quantity(d, SWK_Q_YMOM_C, v->ymom_centroid_values, N));
  CKV(swk_set_quantity(d, SWK_Q_ELEVATION_C, v->bed_centroid_values, N));
  CKV(pull_clock(d));
  const long long before = d->h_clock->negative_cells;
  LAUNCH(d, k_fix_negative, nblk(N), BLOCK, d->D);
  CKV(pull_clock(d));
  if (count) *count = d->h_clock->negative_cells - before;
  CKV(swk_get_quantity(d, SWK_Q_STAGE_C, v->stage_centroid_values, N));
  CKV(swk_get_quantity(d, SWK_Q_XMOM_C, v->xmom_centroid_values, N));
  CKV(swk_get_quantity(d, SWK_Q_YMOM_C, v->ymom_centroid_values, N));
  return SWK_OK;
}

// ---- stateless elementwise entry points ------------------------------------------
struct DevBuf {
  double *p = nullptr;
  ~DevBuf() { if (p) cudaFree(p); }
  int put(const double *h, size_t n)
  {
    CKV(dalloc(&p, n));
    if (h) CK(cudaMemcpy(p, h, n * sizeof(double), cudaMemcpyHostToDevice));
    return SWK_OK;
  }
  int get(double *h, size_t n)
  {
    CK(cudaMemcpy(h, p, n * sizeof(double), cudaMemcpyDeviceToHost));
    return SWK_OK;
  }
};

static int manning_call(int device, int sloped, double g, double eps, int64_t N, const double *x, const double *w,
                        const double *zv, const double *uh, const double *vh, const double *eta,
                        double *xmom_update, double *ymom_update)
{
  if (N < 0 || !w || !zv || !uh || !vh || !eta || !xmom_update || !ymom_update || (sloped && !x))
    return fail(SWK_ERR_ARG, "NULL argument");
  CKV(select_device(device));
  if (N == 0) return SWK_OK;
  DevBuf dx, dw, dz, du, dv, de, dxu, dyu;
  if (sloped) CKV(dx.put(x, 6 * N));
  CKV(dw.put(w, N)); CKV(dz.put(zv, sloped ? 3 * N : N)); CKV(du.put(uh, N)); CKV(dv.put(vh, N));
  CKV(de.put(eta, N)); CKV(dxu.put(xmom_update, N)); CKV(dyu.put(ymom_update, N));
  k_manning_plain<<<nblk(N), BLOCK>>>(g, eps, (int)N, sloped, dx.p, dw.p, dz.p, du.p, dv.p, de.p, dxu.p, dyu.p);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  CKV(dxu.get(xmom_update, N));
  CKV(dyu.get(ymom_update, N));
  return SWK_OK;
}

extern "C" int swk_call_manning_friction_flat(int device, double g, double eps, int64_t N, const double *w,
                                              const double *zv, const double *uh, const double *vh,
                                              const double *eta, double *xmom_update, double *ymom_update)
{
  return manning_call(device, 0, g, eps, N, nullptr, w, zv, uh, vh, eta, xmom_update, ymom_update);
}

extern "C" int swk_call_manning_friction_sloped(int device, double g, double eps, int64_t N, const double *x,
                                                const double *w, const double *zv, const double *uh,
                                                const double *vh, const double *eta, double *xmom_update,
                                                double *ymom_update)
{
  return manning_call(device, 1, g, eps, N, x, w, zv, uh, vh, eta, xmom_update, ymom_update);
}

extern "C" int swk_call_update(int device, int64_t N, double timestep, double *centroid_values,
                               const double *explicit_update, double *semi_implicit_update)
{
  if (N < 0 || !centroid_values || !explicit_update || !semi_implicit_update) return fail(SWK_ERR_ARG, "NULL argument");
  CKV(select_device(device));
  if (N == 0) return SWK_OK;
  DevBuf c, e, s;
  CKV(c.put(centroid_values, N)); CKV(e.put(explicit_update, N)); CKV(s.put(semi_implicit_update, N));
  int *derr = nullptr;
  CKV(dalloc(&derr, 1));
  CK(cudaMemset(derr, 0, sizeof(int)));
  k_update_plain<<<nblk(N), BLOCK>>>((int)N, timestep, c.p, e.p, s.p, derr);
  int herr = 0;
  cudaError_t ce = cudaMemcpy(&herr, derr, sizeof(int), cudaMemcpyDeviceToHost);
  cudaFree(derr);
  if (ce != cudaSuccess) return fail(SWK_ERR_CUDA, cudaGetErrorString(ce));
  if (herr) return fail(SWK_ERR_DENOMINATOR, "semi-implicit update: denominator <= 0 (quantity.c:806)");
  CKV(c.get(centroid_values, N));
  CKV(s.get(semi_implicit_update, N));
  return SWK_OK;
}

extern "C" int swk_call_backup_centroid_values(int device, int64_t N, const double *centroid_values,
                                               double *centroid_backup_values)
{
  if (N < 0 || !centroid_values || !centroid_backup_values) return fail(SWK_ERR_ARG, "NULL argument");
  CKV(select_device(device));
  if (N == 0) return SWK_OK;
  // a copy through the device keeps the call on the same data path as the other entry points
  DevBuf c;
  CKV(c.put(centroid_values, N));
  CKV(c.get(centroid_backup_values, N));
  return SWK_OK;
}

extern "C" int swk_call_saxpy_centroid_values(int device, int64_t N, double a, double b, double *centroid_values,
                                              const double *centroid_backup_values)
{
  if (N < 0 || !centroid_values || !centroid_backup_values) return fail(SWK_ERR_ARG, "NULL argument");
  CKV(select_device(device));
  if (N == 0) return SWK_OK;
  DevBuf c, bk;
  CKV(c.put(centroid_values, N)); CKV(bk.put(centroid_backup_values, N));
  k_saxpy_plain<<<nblk(N), BLOCK>>>((int)N, a, b, c.p, bk.p);
  CK(cudaGetLastError());
  CKV(c.get(centroid_values, N));
  return SWK_OK;
}

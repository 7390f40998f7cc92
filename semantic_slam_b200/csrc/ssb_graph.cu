// ssb_graph.cu — host side of the graph hot path behind the C-ABI of include/ssb.h:
// graph container mirroring ps_graph_slam::GraphSLAM (graph_slam.cpp:40-239), CSR edge-table
// construction, and the Levenberg-Marquardt driver that restates g2o's
// OptimizationAlgorithmLevenberg::solve on top of the sm_100a kernels in
// ssb_graph_kernels.cuh (fused linearise / Schur / block-Jacobi PCG / update / chi2).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <condition_variable>
#include <limits>
#include <map>
#include <mutex>
#include <sstream>
#include <string>
#include <vector>

#include <dlfcn.h>
#include <unistd.h>

#include "../../include/ssb.h"
#include "ssb_graph_kernels.cuh"
#include "ssb_peer.cuh"
#include "ssb_pcg_flow.cuh"
#include "ssb_marg_direct.cuh"

// ---- NCCL, resolved at run time (the library the process already loaded — torch's — else libnccl.so.2).
// Minimal declarations of the stable C ABI; nothing links against libnccl at build time.
extern "C" {
typedef struct ncclComm* ssb_ncclComm_t;
typedef struct {
  char internal[128];
} ssb_ncclUniqueId;
}
namespace ssb {
struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(ssb_ncclUniqueId*) = nullptr;
  int (*CommInitRank)(ssb_ncclComm_t*, int, ssb_ncclUniqueId, int) = nullptr;
  int (*CommDestroy)(ssb_ncclComm_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ssb_ncclComm_t, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, ssb_ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok = false;
};
constexpr int kNcclFloat64 = 8;  // ncclDouble
constexpr int kNcclSum = 0;
static NcclApi& nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api;
  tried = true;
  api.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
  if (!api.handle) api.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
  if (!api.handle) api.handle = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
  if (!api.handle) return api;
  api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.handle, "ncclGetUniqueId");
  api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.handle, "ncclCommInitRank");
  api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.handle, "ncclCommDestroy");
  api.AllReduce = (decltype(api.AllReduce))dlsym(api.handle, "ncclAllReduce");
  api.AllGather = (decltype(api.AllGather))dlsym(api.handle, "ncclAllGather");
  api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.handle, "ncclGetErrorString");
  api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce && api.AllGather && api.GetErrorString;
  return api;
}
#define SSB_NCCL_CHECK(expr)                                                                        \
  do {                                                                                              \
    int _r = (expr);                                                                                \
    if (_r != 0) {                                                                                  \
      ::ssb::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr, nccl_api().GetErrorString(_r)); \
      return SSB_ERR_COMM;                                                                          \
    }                                                                                               \
  } while (0)
}  // namespace ssb

namespace ssb {

thread_local std::string g_last_error;
void set_error(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
}

template <class T>
struct DBuf {
  T* p = nullptr;
  size_t cap = 0;
  bool view = false;   // non-owning window into the peer arena of a sharded graph (ssb_peer.cuh)
  void set_view(void* q, size_t n) {
    if (p && !view) cudaFree(p);
    p = (T*)q;
    cap = n;
    view = true;
  }
  int ensure(size_t n) {
    if (n <= cap && p) return SSB_OK;
    if (view) {
      set_error("internal: peer-arena view too small (%zu > %zu elements)", n, cap);
      return SSB_ERR_INVALID;
    }
    // grow with slack: a graph that gains a keyframe per tick must not reallocate every buffer every tick
    size_t want = std::max<size_t>(n, 1);
    if (p) want = std::max(want, cap + cap / 2);
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    cudaError_t e = cudaMalloc((void**)&p, want * sizeof(T));
    if (e != cudaSuccess) {
      set_error("cudaMalloc(%zu bytes) failed: %s", want * sizeof(T), cudaGetErrorString(e));
      return SSB_ERR_CUDA;
    }
    cap = want;
    return SSB_OK;
  }
  ~DBuf() {
    if (p && !view) cudaFree(p);
  }
};

// page-locked host staging buffer (tables built in place, then uploaded by a truly asynchronous copy)
template <class T>
struct PinnedBuf {
  T* p = nullptr;
  size_t cap = 0;
  int ensure(size_t n) {
    if (n <= cap && p) return SSB_OK;
    size_t want = std::max<size_t>(n, 1);
    if (p) want = std::max(want, cap + cap / 2);
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
    cudaError_t e = cudaHostAlloc((void**)&p, want * sizeof(T), cudaHostAllocDefault);
    if (e != cudaSuccess) {
      set_error("cudaHostAlloc(%zu bytes) failed: %s", want * sizeof(T), cudaGetErrorString(e));
      return SSB_ERR_CUDA;
    }
    cap = want;
    return SSB_OK;
  }
  ~PinnedBuf() {
    if (p) cudaFreeHost(p);
  }
};

static inline double wall_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

enum { VK_SE3 = 0, VK_XYZ = 1, VK_PLANE = 2 };
enum { EK_PP = 0, EK_PL = 1, EK_LL = 2 };

struct HostVertex {
  int kind;
  int idx;  // index into poses / landmarks
  bool fixed;
  int hidx;
};
struct HostEdgeRef {
  int kind;
  int idx;
};
struct LLEdge {
  int a, b;
  double z[3];
  double info[9];
};

// ---- sharded graphs: host-side bookkeeping (device side: ssb_peer.cuh) -------------------------------------------
// Who owns what.  Computed from the full host graph, which every rank holds, so every rank derives the SAME plan for
// ALL ranks without talking to anybody: a rank knows under which local index each of its keyframes / landmark parts
// lives on every neighbour and can aim its pushes there.
struct RankLocal {
  int ps = 0, pe = 0;                 // own keyframes: global pose indices [ps, pe)
  std::vector<int> l2g_pose;          // local pose -> global pose: own first (global order), then ghosts (global order)
  std::vector<int> l2g_lm;            // local landmark -> global: owned first, then the other touched ones
  int n_owned_lm = 0;
  std::vector<int> partbase;          // [local landmarks + 1] first v-cell part of every local landmark
  std::vector<int> g2l_pose, g2l_lm;  // global -> local (-1: not present on this rank)
};
struct ShardPlan {
  int world = 1, nb = 0, Np = 0, Nl = 0;
  std::vector<RankLocal> R;
  std::vector<int> lm_deg;            // [Nl] observations per landmark
  std::vector<int> lm_owner;          // [Nl] rank that eliminates the landmark (owner of its first observer)
};
// byte offsets inside a rank's peer arena (a function of that rank's local sizes => computable by everybody)
struct ArenaLayout {
  size_t flags, red, cells, lines, x, z, v, slots, pose_full, lm_full, gcent, grows, glines, total;
  size_t n_cells, n_ucells, n_lines, n_x, n_v;
};
struct PeerBlob {                     // what the ranks tell each other when an arena was (re)allocated
  int pid, device;
  void* ptr;
  unsigned long long bytes, serial;
  cudaIpcMemHandle_t handle;
};
struct MrCtx {                        // member of the INNER (shard) handle
  int world = 1, rank = 0, nb = 0;
  int n_own = 0, n_owned_lm = 0;
  std::vector<int> sortkey;           // local pose -> global pose index: the L-order is by GLOBAL keyframe index, so
                                      // every rank (and the unsharded run) sums a landmark's edges in the same order
  std::vector<unsigned char> lm_owned;
  // peer arena
  unsigned char* arena = nullptr;
  size_t arena_cap = 0;
  unsigned long long arena_serial = 0;
  std::vector<unsigned char*> retired;        // replaced arenas, freed once every peer has unmapped them
  unsigned char* peer_base[SSB_MAX_WORLD] = {};
  PeerBlob peer_blob[SSB_MAX_WORLD] = {};
  bool peer_mapped[SSB_MAX_WORLD] = {};
  ArenaLayout lay[SSB_MAX_WORLD] = {};
  PeerDev P{};
  PeerGather PG{};
  unsigned long long epoch = 0;               // k_peer_exchange calls so far (identical on every rank)
  // push tables (absolute pointers into the peers' arenas)
  DBuf<int> d_upush_rowptr, d_vpush_rowptr;
  DBuf<uint4*> d_upush_cell, d_vpush_cell;
  DBuf<double*> d_upush_x;
  FlowPeer FP{};
  StreamPeer SP{};
  GlobDev GD{};                               // rank-level coarse level of the preconditioner
  bool glob = false;
  DBuf<int> d_pose_rank;
  DBuf<double> d_Bg, d_Gg, d_gpart;
  DBuf<float> d_Aginv;
  DBuf<double*> d_gcent_all, d_grows_all;
  DBuf<double*> d_zpush_z, d_svpush_v;
  DBuf<int> d_svpush_rowptr;
  // chi2 over the edges this rank owns (a shared edge is counted once)
  DBuf<PLEdge> d_pl_own;
  DBuf<PPEdge> d_pp_own;
  DBuf<double> d_zd_own;
  int n_pl_own = 0, n_pp_own = 0;
  DBuf<int> d_err;
  double* h_red = nullptr;                    // pinned: [world][PEER_RED_N]
  int* h_err = nullptr;                       // pinned
  DBuf<int> d_own_l2g;                        // [n_own] and [n_owned_lm]: targets of the final estimate gather
  DBuf<int> d_ownlm_l2g;
};

}  // namespace ssb

using namespace ssb;

struct ssb_graph {
  ssb_graph_opts opts;
  // host graph (authoritative for structure; estimates mirrored, see est_on_device)
  std::vector<HostVertex> V;
  std::vector<HostEdgeRef> E;
  std::vector<Pose> poses;      // estimates, pose index order
  std::vector<double> lms;      // 4 per landmark vertex: xyz + pad, or the 4 normalised plane coefficients
  std::vector<unsigned char> lm_kind;  // per landmark vertex: 0 = VertexPointXYZ, 1 = VertexPlane
  std::vector<double> pl_zd;    // per pose-landmark edge (creation order): 4th coefficient of a measured plane
  int n_plane_vertices = 0;
  std::vector<int> pose_vid, lm_vid;
  std::vector<PPEdge> pp;       // creation order
  std::vector<PLEdge> pl;       // creation order (host); device copy is L-order
  std::vector<int> pl_full_info_sym;  // unused
  std::vector<LLEdge> ll;
  std::vector<int> plL_of_edge;  // creation index -> L-order position
  // landmarks promoted into the reduced system (they carry a landmark-landmark edge, ssb_math.cuh pp_edge_linearize):
  // the tables prepare() works on when `ll` is not empty — keyframes + pseudo-keyframes, their edges as pose-pose entries
  std::vector<Pose> eff_poses;
  std::vector<PPEdge> eff_pp;
  std::vector<PLEdge> eff_pl;
  std::vector<double> eff_zd;
  std::vector<int> prom_lm;            // promoted landmark ids (index k -> pseudo-keyframe poses.size() + k)
  std::vector<unsigned char> eff_kind; // per effective keyframe: 1 = promoted landmark
  DBuf<unsigned char> d_pose_kind;
  bool structure_dirty = true;
  bool host_est_dirty = true;    // host estimates changed since last upload
  bool device_est_newer = false; // device estimates not yet copied back
  bool have_system = false;      // a linearised system (H, b) is resident
  bool have_snapshot = false;
  int device = 0;
  int num_sms = 0;
  int pcg_grid = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t stream2 = nullptr;   // auxiliary stream for independent small kernels (forked from / joined to `stream`)
  cudaStream_t stream3 = nullptr;   // the coarse-matrix inversion of a damped trial (k_coarse_invert)
  cudaEvent_t ev_fork = nullptr, ev_mid = nullptr, ev_join = nullptr, ev_join3 = nullptr;
  bool coarse_ready = false;        // k_coarse_invert ran for the system / lambda of the coming k_pcg_flow launches
  size_t cinv_smem = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  std::vector<cudaEvent_t> ev_pool;  // pairs around every k_pcg launch of the current optimize
  size_t ev_used = 0;
  // device buffers
  DBuf<Pose> d_pose, d_pose_bak, d_pose_snap;
  DBuf<double> d_lm, d_lm_bak, d_lm_snap;
  DBuf<unsigned char> d_pose_fixed, d_lm_fixed, d_lm_kind, d_lm_owned, d_blob_stage;
  DBuf<double> d_pl_zd;
  DBuf<PLEdge> d_pl;
  PinnedBuf<PLEdge> h_plL;   // L-ordered pose-landmark edges, built in place by prepare()
  PinnedBuf<double> h_zdL;
  DBuf<PPEdge> d_pp;
  DBuf<int> d_lm_rowptr, d_pose_pl_rowptr, d_pose_pl_idx, d_pose_pp_rowptr, d_pose_pp_idx, d_pose_pp_other, d_plP_lm;
  DBuf<double> d_Hpp, d_bp, d_Hoff, d_Hll, d_bl, d_HplL, d_HplP, d_HllInv, d_Dinv, d_g;
  DBuf<double> d_x, d_r, d_z, d_p0, d_p1, d_q, d_v, d_dl, d_part, d_scalars, d_tmp;
  DBuf<int> d_iscalars, d_ainv_ok;
  // landmark marginals, direct form (ssb_marg_direct.cuh)
  DBuf<double> md_Bsub, md_Ginv, md_Esub, md_Y, md_T, md_Row, md_ColT, md_Pinv, md_out;
  DBuf<int> md_k0, md_status, md_lidx;
  PinnedBuf<int> md_hstatus;
  // coarse level
  DBuf<double> d_Bmat, d_Grun, d_panel, d_B1mat, d_D1inv, d_ainv, d_ctacen;
  bool ainv_valid = false;   // d_ainv holds the rows of a previously inverted coarse matrix
  int solves_since_refresh = 0;
  DBuf<int> d_run_lm, d_run_group, d_run_e0, d_lm_run_rowptr, d_grp_run_rowptr, d_grp_runs;
  DBuf<int> d_run1_lm, d_run1_e0, d_agg_run_rowptr, d_agg_runs, d_grp_first_agg, d_grp_seg_rowptr, d_grp_seg_r0, d_grp_seg_m, d_grp_seg_off, d_run1_agg;
  int grp_max_runs = 0;
  DBuf<double> d_Grun1, d_D1raw, d_GrpInv;
  DBuf<BarSlot> d_slots;
  CoarseDev Cz;
  size_t pcg_smem = 0, pcgw_smem = 0;
  DBuf<int> d_ft_ulm_rowptr, d_ft_ulm, d_ft_pl_loc, d_ft_upp_rowptr, d_ft_upp, d_ft_pp_loc, d_ft_pp_src, d_ft_ext_rowptr, d_ft_ext, d_ft_gj_order, d_ft_gj_mask, d_ft_part_lm, d_ft_part_e0, d_ft_part_e1, d_ft_lm_partbase;
  int flow_n_parts = 0;
  FlowTabs FT;
  DBuf<uint4> d_ucell, d_lines, d_gj;
  DBuf<unsigned long long> d_trace;
  DBuf<double2> d_hlpark;   // tagged cells of the data-flow PCG kernel (ssb_pcg_flow.cuh)
  unsigned flow_seq = 0;                   // launch counter -> tag base (seq << 16)
  bool use_flow = true;                    // the grid is 148 CTAs (B200): k_pcg_flow is instantiated for that size
  bool fast_ok = false;   // the graph fits the on-chip resident PCG kernel
  bool allow_fast = true;
  double* h_scalars = nullptr;  // pinned: 8 doubles
  int* h_iscalars = nullptr;    // pinned: 4 ints
  DevGraph G;
  std::vector<double> history;  // 6 per iteration
  long long launches = 0;
  int comm_rank = 0, comm_world = 1;
  ssb_ncclComm_t comm = nullptr;
  // ---- one graph sharded over several ranks (ssb_peer.cuh, "sharded graphs" below) ----
  int local_group = 0;           // 1: the ranks are host threads of this process (ssb_graph_attach_local)
  std::string local_key;
  ssb_graph* shard = nullptr;    // outer handle: this rank's local subgraph (own + ghost keyframes), a handle of its own
  MrCtx* mr = nullptr;           // inner (shard) handle: what it needs to know about the other ranks
  ShardPlan* plan = nullptr;     // outer handle: who owns what, identical on every rank
  std::vector<Pose> snap_poses;  // outer handle: host copy of the estimates at ssb_graph_snapshot
  std::vector<double> snap_lms;
  // ---- landmark marginals of graphs that fill only part of the chip (marginals_replicated) ----
  ssb_graph* marg_rep = nullptr;            // k copies of this graph side by side: one PCG launch = k columns
  int marg_rep_k = 0;
  unsigned long long structure_serial = 0;  // bumped whenever prepare() rebuilt the tables
  unsigned long long marg_rep_serial = 0;   // structure_serial the replicated graph was built from
  bool linpoint_in_bak = false;             // d_pose_bak / d_lm_bak hold the estimates the resident system was linearised at
  // the replicated (shadow) handle itself: copy j lives on the CTAs [j rep_ctas, (j + 1) rep_ctas) of the PCG grid
  int rep_ctas = 0, rep_count = 0;
  int force_C = 0;                          // keyframes per CTA (0: derived from the graph size)
  int part_edges = 64;                      // a landmark's edges are cut into parts of at most this many (one warp each); 32 = no
                                            // second batch of u cells per part (the copies multiply the parts a CTA gets)
  int marg_plan[4] = {1, 0, 0, 0};          // cached MargRepPlan of this structure (copies, CTAs per copy, C, stride)
  unsigned long long marg_plan_serial = ~0ULL;
  int marg_plan_env = -2;
  // sharded graphs: the marginals are computed on an unsharded copy of the graph on this rank's GPU (marginals_sharded)
  ssb_graph* marg_full = nullptr;
  unsigned long long marg_full_rev = ~0ULL;
  unsigned long long host_rev = 0;          // bumped by every call that changes the structure of the host graph
};

// k_pcg_flow is instantiated for the grids it is launched with: 148 CTAs (one per B200 SM); a sharded graph whose ranks
// are host threads sharing one GPU ("virtual shards", ssb_graph_attach_local) runs 74 or 37 CTAs per rank
static void* flow_kernel(int grid, bool mr, bool rep = false) {
  if (!mr) return grid == 148 ? (rep ? (void*)k_pcg_flow<148, false, true> : (void*)k_pcg_flow<148, false, false>) : nullptr;
  switch (grid) {
    case 148: return (void*)k_pcg_flow<148, true>;
    case 74: return (void*)k_pcg_flow<74, true>;
    case 37: return (void*)k_pcg_flow<37, true>;
    default: return nullptr;
  }
}
static int check_vertex(const ssb_graph* g, int id, int kind) {
  return g && id >= 0 && id < (int)g->V.size() && g->V[id].kind == kind;
}

static void pose_from_34(const double* T, Pose& P) {
  double R[9] = {T[0], T[1], T[2], T[4], T[5], T[6], T[8], T[9], T[10]};
  R_to_quat(R, P.q);
  P.t[0] = T[3];
  P.t[1] = T[7];
  P.t[2] = T[11];
  P.pad = 0.0;
}
static void pose_to_34(const Pose& P, double* T) {
  double R[9];
  quat_to_R(P.q, R);
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) T[4 * r + c] = R[3 * r + c];
    T[4 * r + 3] = P.t[r];
  }
}

extern "C" {

const char* ssb_last_error(void) { return g_last_error.c_str(); }
#define SSB_STR2(x) #x
#define SSB_STR(x) SSB_STR2(x)
const char* ssb_build_info(void) {
  return "semantic_slam_b200 libssb: sm_100a, CUDA " SSB_STR(__CUDACC_VER_MAJOR__) "." SSB_STR(__CUDACC_VER_MINOR__);
}

void ssb_graph_default_opts(ssb_graph_opts* o) {
  if (!o) return;
  std::memset(o, 0, sizeof(*o));
  o->device = -1;
  o->verbose = 0;
  o->max_pcg_iters = 20000;
  o->pcg_tol = 1e-8;
  o->preconditioner = 0;
  o->coarse_group = 32;
}

ssb_graph* ssb_graph_create(const ssb_graph_opts* opts) {
  ssb_graph* g = new ssb_graph();
  if (opts)
    g->opts = *opts;
  else
    ssb_graph_default_opts(&g->opts);
  if (g->opts.max_pcg_iters <= 0) g->opts.max_pcg_iters = 20000;
  if (!(g->opts.pcg_tol > 0)) g->opts.pcg_tol = 1e-8;
  int dev = g->opts.device;
  cudaError_t e;
  if (dev < 0) {
    e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
      set_error("no CUDA device: %s (the CUDA back-end has no CPU fallback)", cudaGetErrorString(e));
      delete g;
      return nullptr;
    }
  }
  e = cudaSetDevice(dev);
  if (e != cudaSuccess) {
    set_error("cudaSetDevice(%d) failed: %s (the CUDA back-end has no CPU fallback)", dev, cudaGetErrorString(e));
    delete g;
    return nullptr;
  }
  g->device = dev;
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, dev);
  if (e != cudaSuccess) {
    set_error("cudaGetDeviceProperties failed: %s", cudaGetErrorString(e));
    delete g;
    return nullptr;
  }
  g->num_sms = prop.multiProcessorCount;
  if (cudaStreamCreateWithFlags(&g->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&g->stream2, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&g->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&g->ev_mid, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&g->ev_join, cudaEventDisableTiming) != cudaSuccess ||
      cudaStreamCreateWithFlags(&g->stream3, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&g->ev_join3, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreate(&g->ev0) != cudaSuccess || cudaEventCreate(&g->ev1) != cudaSuccess ||
      cudaMallocHost((void**)&g->h_scalars, 32 * sizeof(double)) != cudaSuccess ||
      cudaMallocHost((void**)&g->h_iscalars, 4 * sizeof(int)) != cudaSuccess) {
    set_error("CUDA resource creation failed: %s", cudaGetErrorString(cudaGetLastError()));
    delete g;
    return nullptr;
  }
  g->pcg_grid = std::min(g->num_sms, PCG_THREADS);  // one persistent CTA per SM
  const bool shard_handle = g->opts.reserved[2] > 0;   // internal: a shard of a larger graph, reserved[2] = CTAs of this rank
  if (shard_handle) g->pcg_grid = std::min(g->pcg_grid, g->opts.reserved[2]);
  g->pcg_smem = (size_t)(PCG_PART + 14 * 6 * g->pcg_grid + (PCG_THREADS / 36) * 36) * sizeof(double);
  g->allow_fast = g->opts.reserved[0] == 0;
  g->use_flow = flow_kernel(g->pcg_grid, shard_handle) != nullptr;
  g->pcgw_smem = pcg_flow_smem_doubles(g->pcg_grid) * sizeof(double);
  int nb = 0;
  void* sk = shard_handle ? (void*)k_pcg<true> : (void*)k_pcg<false>;
  e = cudaFuncSetAttribute(sk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g->pcg_smem);
  if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, sk, PCG_THREADS, g->pcg_smem);
  if (e != cudaSuccess || nb < 1) {
    set_error("k_pcg cannot be made resident (occupancy %d, %zu B smem): %s", nb, g->pcg_smem, cudaGetErrorString(e));
    delete g;
    return nullptr;
  }
  if (g->use_flow) {
    void* fk = flow_kernel(g->pcg_grid, shard_handle);
    e = cudaFuncSetAttribute(fk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g->pcgw_smem);
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fk, PCGF_THREADS, g->pcgw_smem);
    if (e == cudaSuccess && !shard_handle && flow_kernel(g->pcg_grid, false, true))   // several right-hand sides per launch (K5)
      e = cudaFuncSetAttribute(flow_kernel(g->pcg_grid, false, true), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g->pcgw_smem);
  }
  if (e != cudaSuccess || nb < 1) {
    set_error("k_pcg_flow cannot be made resident (occupancy %d, %zu B smem): %s", nb, g->pcgw_smem, cudaGetErrorString(e));
    delete g;
    return nullptr;
  }
  return g;
}

void ssb_graph_destroy(ssb_graph* g) {
  if (!g) return;
  cudaSetDevice(g->device);
  if (g->shard) {
    ssb_graph_destroy(g->shard);
    g->shard = nullptr;
  }
  if (g->marg_rep) {
    ssb_graph_destroy(g->marg_rep);
    g->marg_rep = nullptr;
  }
  if (g->marg_full) {
    ssb_graph_destroy(g->marg_full);
    g->marg_full = nullptr;
  }
  delete g->plan;
  g->plan = nullptr;
  if (g->mr) {
    MrCtx* mr = g->mr;
    if (g->stream) cudaStreamSynchronize(g->stream);
    // windows into the arena are not owned by their DBufs
    for (int r = 0; r < SSB_MAX_WORLD; ++r)
      if (mr->peer_mapped[r] && mr->peer_blob[r].pid != (int)getpid()) cudaIpcCloseMemHandle(mr->peer_base[r]);
    for (unsigned char* q : mr->retired) cudaFree(q);
    if (mr->arena) cudaFree(mr->arena);
    if (mr->h_red) cudaFreeHost(mr->h_red);
    if (mr->h_err) cudaFreeHost(mr->h_err);
    delete mr;
    g->mr = nullptr;
  }
  if (g->stream) cudaStreamSynchronize(g->stream);
  for (cudaEvent_t e : g->ev_pool) cudaEventDestroy(e);
  if (g->stream2) cudaStreamSynchronize(g->stream2);
  if (g->stream3) cudaStreamSynchronize(g->stream3);
  if (g->ev_join3) cudaEventDestroy(g->ev_join3);
  if (g->stream3) cudaStreamDestroy(g->stream3);
  if (g->ev_fork) cudaEventDestroy(g->ev_fork);
  if (g->ev_mid) cudaEventDestroy(g->ev_mid);
  if (g->ev_join) cudaEventDestroy(g->ev_join);
  if (g->stream2) cudaStreamDestroy(g->stream2);
  if (g->ev0) cudaEventDestroy(g->ev0);
  if (g->ev1) cudaEventDestroy(g->ev1);
  if (g->h_scalars) cudaFreeHost(g->h_scalars);
  if (g->h_iscalars) cudaFreeHost(g->h_iscalars);
  if (g->comm && nccl_api().ok) nccl_api().CommDestroy(g->comm);
  if (g->stream) cudaStreamDestroy(g->stream);
  delete g;
}

void* ssb_graph_stream(ssb_graph* g) { return g ? (void*)g->stream : nullptr; }

static int sync_estimates_to_host(ssb_graph* g);

int ssb_graph_add_se3_node(ssb_graph* g, const double T34[12]) {
  if (!g || !T34) return SSB_ERR_INVALID;
  if (sync_estimates_to_host(g) != SSB_OK) return SSB_ERR_CUDA;
  HostVertex v;
  v.kind = VK_SE3;
  v.idx = (int)g->poses.size();
  v.fixed = g->V.empty();  // graph_slam.cpp:109-111
  v.hidx = -1;
  Pose P;
  pose_from_34(T34, P);
  g->poses.push_back(P);
  g->pose_vid.push_back((int)g->V.size());
  g->V.push_back(v);
  g->structure_dirty = true;
  g->host_rev++;
  return (int)g->V.size() - 1;
}

int ssb_graph_add_point_xyz_node(ssb_graph* g, const double xyz[3]) {
  if (!g || !xyz) return SSB_ERR_INVALID;
  if (sync_estimates_to_host(g) != SSB_OK) return SSB_ERR_CUDA;
  HostVertex v;
  v.kind = VK_XYZ;
  v.idx = (int)(g->lms.size() / 4);
  v.fixed = false;
  v.hidx = -1;
  g->lms.push_back(xyz[0]);
  g->lms.push_back(xyz[1]);
  g->lms.push_back(xyz[2]);
  g->lms.push_back(0.0);
  g->lm_kind.push_back(0);
  g->lm_vid.push_back((int)g->V.size());
  g->V.push_back(v);
  g->structure_dirty = true;
  g->host_rev++;
  return (int)g->V.size() - 1;
}

int ssb_graph_add_se3_edge(ssb_graph* g, int v1, int v2, const double Z34[12], const double info[36]) {
  if (!g || !Z34 || !info || !check_vertex(g, v1, VK_SE3) || !check_vertex(g, v2, VK_SE3)) {
    set_error("add_se3_edge: invalid vertex ids %d, %d", v1, v2);
    return SSB_ERR_INVALID;
  }
  PPEdge e;
  std::memset(&e, 0, sizeof(e));
  e.i = g->V[v1].idx;
  e.j = g->V[v2].idx;
  Pose Z;
  pose_from_34(Z34, Z);
  for (int k = 0; k < 3; ++k) e.zt[k] = Z.t[k];
  for (int k = 0; k < 4; ++k) e.zq[k] = Z.q[k];
  int k = 0;
  for (int r = 0; r < 6; ++r)
    for (int c = r; c < 6; ++c) e.info[k++] = 0.5 * (info[6 * r + c] + info[6 * c + r]);
  g->pp.push_back(e);
  g->E.push_back({EK_PP, (int)g->pp.size() - 1});
  g->structure_dirty = true;
  g->host_rev++;
  return (int)g->E.size() - 1;
}

int ssb_graph_add_se3_point_xyz_edge(ssb_graph* g, int v_se3, int v_xyz, const double xyz[3], const double info[9]) {
  if (!g || !xyz || !info || !check_vertex(g, v_se3, VK_SE3) || !check_vertex(g, v_xyz, VK_XYZ)) {
    set_error("add_se3_point_xyz_edge: invalid vertex ids %d, %d", v_se3, v_xyz);
    return SSB_ERR_INVALID;
  }
  PLEdge e;
  e.p = g->V[v_se3].idx;
  e.l = g->V[v_xyz].idx;
  for (int k = 0; k < 3; ++k) e.z[k] = xyz[k];
  e.info[0] = info[0];
  e.info[1] = 0.5 * (info[1] + info[3]);
  e.info[2] = 0.5 * (info[2] + info[6]);
  e.info[3] = info[4];
  e.info[4] = 0.5 * (info[5] + info[7]);
  e.info[5] = info[8];
  g->pl.push_back(e);
  g->pl_zd.push_back(0.0);
  g->E.push_back({EK_PL, (int)g->pl.size() - 1});
  g->structure_dirty = true;
  g->host_rev++;
  return (int)g->E.size() - 1;
}

int ssb_graph_add_point_xyz_point_xyz_edge(ssb_graph* g, int v1, int v2, const double xyz[3], const double info[9]) {
  if (!g || !xyz || !info || !check_vertex(g, v1, VK_XYZ) || !check_vertex(g, v2, VK_XYZ)) {
    set_error("add_point_xyz_point_xyz_edge: invalid vertex ids %d, %d", v1, v2);
    return SSB_ERR_INVALID;
  }
  LLEdge e;
  e.a = g->V[v1].idx;
  e.b = g->V[v2].idx;
  std::memcpy(e.z, xyz, sizeof(e.z));
  std::memcpy(e.info, info, sizeof(e.info));
  g->ll.push_back(e);
  g->E.push_back({EK_LL, (int)g->ll.size() - 1});
  g->structure_dirty = true;
  g->host_rev++;
  return (int)g->E.size() - 1;
}

// VertexPlane / EdgeSE3Plane: the plane API the reference keeps commented out (graph_slam.hpp:44,74-75,
// graph_slam.cpp:117-125) with its own edge type include/g2o/edge_se3_plane.hpp.  A plane vertex is a 3-DoF
// landmark for the Schur back-end, exactly like a point landmark.
int ssb_graph_add_plane_node(ssb_graph* g, const double coeffs[4]) {
  if (!g || !coeffs) return SSB_ERR_INVALID;
  if (sync_estimates_to_host(g) != SSB_OK) return SSB_ERR_CUDA;
  double c[4] = {coeffs[0], coeffs[1], coeffs[2], coeffs[3]};
  const double n = std::sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
  if (!(n > 0.0) || !std::isfinite(n)) {
    set_error("add_plane_node: degenerate plane normal");
    return SSB_ERR_INVALID;
  }
  plane_normalize(c);  // Plane3D(const Vector4D&)
  HostVertex v;
  v.kind = VK_PLANE;
  v.idx = (int)(g->lms.size() / 4);
  v.fixed = false;
  v.hidx = -1;
  for (int k = 0; k < 4; ++k) g->lms.push_back(c[k]);
  g->lm_kind.push_back(1);
  g->n_plane_vertices++;
  g->lm_vid.push_back((int)g->V.size());
  g->V.push_back(v);
  g->structure_dirty = true;
  g->host_rev++;
  return (int)g->V.size() - 1;
}

int ssb_graph_add_se3_plane_edge(ssb_graph* g, int v_se3, int v_plane, const double plane[4], const double info[9]) {
  if (!g || !plane || !info || !check_vertex(g, v_se3, VK_SE3) || !check_vertex(g, v_plane, VK_PLANE)) {
    set_error("add_se3_plane_edge: invalid vertex ids %d, %d", v_se3, v_plane);
    return SSB_ERR_INVALID;
  }
  double c[4] = {plane[0], plane[1], plane[2], plane[3]};
  const double n = std::sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
  if (!(n > 0.0) || !std::isfinite(n)) {
    set_error("add_se3_plane_edge: degenerate plane normal");
    return SSB_ERR_INVALID;
  }
  plane_normalize(c);  // setMeasurement(Plane3D(v))
  PLEdge e;
  e.p = g->V[v_se3].idx;
  e.l = g->V[v_plane].idx;
  for (int k = 0; k < 3; ++k) e.z[k] = c[k];
  e.info[0] = info[0];
  e.info[1] = 0.5 * (info[1] + info[3]);
  e.info[2] = 0.5 * (info[2] + info[6]);
  e.info[3] = info[4];
  e.info[4] = 0.5 * (info[5] + info[7]);
  e.info[5] = info[8];
  g->pl.push_back(e);
  g->pl_zd.push_back(c[3]);
  g->E.push_back({EK_PL, (int)g->pl.size() - 1});
  g->structure_dirty = true;
  g->host_rev++;
  return (int)g->E.size() - 1;
}

int ssb_graph_num_vertices(const ssb_graph* g) { return g ? (int)g->V.size() : SSB_ERR_INVALID; }
int ssb_graph_num_edges(const ssb_graph* g) { return g ? (int)g->E.size() : SSB_ERR_INVALID; }

}  // extern "C"

// --------------------------------------------------------------------------------------------
// device state management
// --------------------------------------------------------------------------------------------
static int sync_estimates_to_host(ssb_graph* g) {
  if (!g->device_est_newer) return SSB_OK;
  SSB_CUDA_CHECK(cudaSetDevice(g->device));
  const size_t Np = g->poses.size(), Nl = g->lms.size() / 4;
  const bool prom = !g->ll.empty() && g->eff_poses.size() == Np + g->prom_lm.size() && !g->prom_lm.empty();
  if (prom)
    SSB_CUDA_CHECK(cudaMemcpyAsync(g->eff_poses.data(), g->d_pose.p, g->eff_poses.size() * sizeof(Pose), cudaMemcpyDeviceToHost, g->stream));
  else if (Np)
    SSB_CUDA_CHECK(cudaMemcpyAsync(g->poses.data(), g->d_pose.p, Np * sizeof(Pose), cudaMemcpyDeviceToHost, g->stream));
  if (Nl) SSB_CUDA_CHECK(cudaMemcpyAsync(g->lms.data(), g->d_lm.p, Nl * 4 * sizeof(double), cudaMemcpyDeviceToHost, g->stream));
  SSB_CUDA_CHECK(cudaStreamSynchronize(g->stream));
  if (prom) {   // promoted landmarks live among the keyframes on the device
    std::copy(g->eff_poses.begin(), g->eff_poses.begin() + Np, g->poses.begin());
    for (size_t k = 0; k < g->prom_lm.size(); ++k)
      for (int c = 0; c < 3; ++c) g->lms[4 * (size_t)g->prom_lm[k] + c] = g->eff_poses[Np + k].t[c];
  }
  g->device_est_newer = false;
  return SSB_OK;
}

#define SSB_TRY(x)            \
  do {                        \
    int _r = (x);             \
    if (_r != SSB_OK) return _r; \
  } while (0)

// Nested-dissection pivot order for the block Gauss-Jordan of the coarse matrix A_c (k_pcg_flow prologue).
// Aggregates are adjacent when a pose-pose edge links them or they observe a common landmark.  With separators
// eliminated last, rows of one subdomain have an exactly-zero block in the pivot column of another subdomain,
// so they skip those steps, and independent subdomains are inverted concurrently: the critical path of the
// data-flow elimination shrinks from N steps to the depth of the elimination tree.
static void nd_order_rec(const std::vector<std::vector<int>>& adj, std::vector<int> nodes, std::vector<char>& in,
                         std::vector<int>& order) {
  if (nodes.size() <= 2) {
    for (int v : nodes) order.push_back(v);
    return;
  }
  for (int v : nodes) in[v] = 1;
  auto bfs = [&](int src, std::vector<int>& level) {
    // levels over the induced subgraph; returns the last node reached
    for (int v : nodes) level[v] = -1;
    std::vector<int> q{src};
    level[src] = 0;
    for (size_t h = 0; h < q.size(); ++h)
      for (int w : adj[q[h]])
        if (in[w] && level[w] < 0) {
          level[w] = level[q[h]] + 1;
          q.push_back(w);
        }
    return q;
  };
  std::vector<int> level(adj.size(), -1);
  std::vector<int> reach = bfs(nodes[0], level);
  if (reach.size() < nodes.size()) {
    // disconnected: order every component on its own
    std::vector<int> rest;
    for (int v : nodes)
      if (level[v] < 0) rest.push_back(v);
    for (int v : nodes) in[v] = 0;
    nd_order_rec(adj, reach, in, order);
    nd_order_rec(adj, rest, in, order);
    return;
  }
  reach = bfs(reach.back(), level);   // pseudo-peripheral start
  reach = bfs(reach.back(), level);
  const int depth = level[reach.back()];
  for (int v : nodes) in[v] = 0;
  if (depth < 2) {  // (nearly) complete graph: no useful separator
    for (int v : nodes) order.push_back(v);
    return;
  }
  // separator = the BFS level that splits the node count most evenly
  std::vector<int> cnt(depth + 1, 0);
  for (int v : nodes) cnt[level[v]]++;
  int best = 1, acc = cnt[0];
  long best_score = -1;
  for (int L = 1; L < depth; ++L) {
    const long a = acc, b = (long)nodes.size() - acc - cnt[L];
    const long score = std::min(a, b) * 4 - cnt[L];
    if (score > best_score) {
      best_score = score;
      best = L;
    }
    acc += cnt[L];
  }
  std::vector<int> A, B, S;
  for (int v : nodes) (level[v] < best ? A : level[v] > best ? B : S).push_back(v);
  nd_order_rec(adj, A, in, order);
  nd_order_rec(adj, B, in, order);
  for (int v : S) order.push_back(v);
}

// 3x3 information (upper triangle, 6 values) into the upper-left corner of a 6x6 upper triangle (21 values)
static void pack_info3_into6(const double* u3, double* u21) {
  for (int k = 0; k < 21; ++k) u21[k] = 0.0;
  u21[0] = u3[0];
  u21[1] = u3[1];
  u21[2] = u3[2];
  u21[6] = u3[3];
  u21[7] = u3[4];
  u21[11] = u3[5];
}
static void refresh_promoted_estimates(ssb_graph* g) {
  const size_t Np = g->poses.size();
  g->eff_poses.resize(Np + g->prom_lm.size());
  std::copy(g->poses.begin(), g->poses.end(), g->eff_poses.begin());
  for (size_t k = 0; k < g->prom_lm.size(); ++k) {
    Pose P;
    std::memset(&P, 0, sizeof(P));
    for (int c = 0; c < 3; ++c) P.t[c] = g->lms[4 * (size_t)g->prom_lm[k] + c];
    P.q[3] = 1.0;
    g->eff_poses[Np + k] = P;
  }
}
// landmark-landmark edges (GraphSLAM::add_point_xyz_point_xyz_edge, graph_slam.cpp:168-180): every landmark they touch
// leaves the point-wise Schur elimination and joins the reduced system as a pseudo-keyframe
static int build_promoted(ssb_graph* g) {
  const int Np = (int)g->poses.size(), Nl = (int)(g->lms.size() / 4);
  std::vector<int> prom_idx(std::max(Nl, 1), -1);
  g->prom_lm.clear();
  for (auto& e : g->ll)
    for (int l : {e.a, e.b}) {
      if (g->lm_kind[l]) {
        set_error("landmark-landmark edges between plane vertices are not defined (EdgePointXYZ joins VertexPointXYZ)");
        return SSB_ERR_INVALID;
      }
      if (prom_idx[l] < 0) {
        prom_idx[l] = (int)g->prom_lm.size();
        g->prom_lm.push_back(l);
      }
    }
  refresh_promoted_estimates(g);
  g->eff_kind.assign(Np + g->prom_lm.size(), 0);
  for (size_t k = 0; k < g->prom_lm.size(); ++k) g->eff_kind[Np + k] = 1;
  g->eff_pp = g->pp;
  g->eff_pl.clear();
  g->eff_zd.clear();
  for (size_t k = 0; k < g->pl.size(); ++k) {
    const PLEdge& e = g->pl[k];
    if (prom_idx[e.l] < 0) {
      g->eff_pl.push_back(e);
      g->eff_zd.push_back(g->pl_zd[k]);
      continue;
    }
    PPEdge q;
    std::memset(&q, 0, sizeof(q));
    q.i = e.p;
    q.j = Np + prom_idx[e.l];
    for (int c = 0; c < 3; ++c) q.zt[c] = e.z[c];
    q.zq[3] = 1.0;
    pack_info3_into6(e.info, q.info);
    q.pad = 1.0;
    g->eff_pp.push_back(q);
  }
  for (auto& e : g->ll) {
    PPEdge q;
    std::memset(&q, 0, sizeof(q));
    q.i = Np + prom_idx[e.a];
    q.j = Np + prom_idx[e.b];
    for (int c = 0; c < 3; ++c) q.zt[c] = e.z[c];
    q.zq[3] = 1.0;
    const double u3[6] = {e.info[0], 0.5 * (e.info[1] + e.info[3]), 0.5 * (e.info[2] + e.info[6]), e.info[4],
                          0.5 * (e.info[5] + e.info[7]), e.info[8]};
    pack_info3_into6(u3, q.info);
    q.pad = 2.0;
    g->eff_pp.push_back(q);
  }
  return SSB_OK;
}

// Build CSR edge tables (initializeOptimization + buildStructure analogue) and upload everything.
static int prepare(ssb_graph* g) {
  SSB_CUDA_CHECK(cudaSetDevice(g->device));
  const bool prom = !g->ll.empty();
  if (prom && g->structure_dirty) {
    SSB_TRY(sync_estimates_to_host(g));
    SSB_TRY(build_promoted(g));
  }
  const std::vector<Pose>& POSES = prom ? g->eff_poses : g->poses;
  const std::vector<PPEdge>& PP = prom ? g->eff_pp : g->pp;
  const std::vector<PLEdge>& PL = prom ? g->eff_pl : g->pl;
  const std::vector<double>& ZD = prom ? g->eff_zd : g->pl_zd;
  const int Np = (int)POSES.size(), Nl = (int)(g->lms.size() / 4);
  const int El = (int)PL.size(), Epp = (int)PP.size();
  // a shard of a larger graph: own keyframes [0, n_own), then ghosts; owned landmarks [0, n_owned_lm), then ghosts
  const MrCtx* mr = g->mr;
  const int n_own = mr ? mr->n_own : Np;
  const int n_owned_lm = mr ? mr->n_owned_lm : Nl;
  if (g->structure_dirty) {
    SSB_TRY(sync_estimates_to_host(g));
    // hessian indices (buildIndexMapping: id order, fixed = -1)
    int h = 0;
    for (auto& v : g->V) v.hidx = v.fixed ? -1 : h++;
    std::vector<unsigned char> pfix(std::max(Np, 1), 0), lfix(std::max(Nl, 1), 0);
    for (auto& v : g->V) {
      if (v.kind == VK_SE3)
        pfix[v.idx] = v.fixed;
      else
        lfix[v.idx] = v.fixed;
    }
    for (size_t k = 0; k < (prom ? g->prom_lm.size() : 0); ++k) pfix[g->poses.size() + k] = lfix[g->prom_lm[k]];
    const bool prep_timing = std::getenv("SSB_PREP_TIMING") != nullptr;
    double tp0 = wall_ms();
    auto tick = [&](const char* what) {
      if (prep_timing) {
        const double t = wall_ms();
        std::fprintf(stderr, "[ssb prepare] %-28s %.3f ms\n", what, t - tp0);
        tp0 = t;
      }
    };
    // L-order (stable counting sort by landmark)
    std::vector<int> lm_rowptr(Nl + 1, 0);
    for (auto& e : PL) lm_rowptr[e.l + 1]++;
    for (int l = 0; l < Nl; ++l) lm_rowptr[l + 1] += lm_rowptr[l];
    SSB_TRY(g->h_plL.ensure(std::max(El, 1)));
    SSB_TRY(g->h_zdL.ensure(std::max(El, 1)));
    SSB_TRY(g->d_pl.ensure(El));
    if (g->n_plane_vertices) SSB_TRY(g->d_pl_zd.ensure(El));
    PLEdge* plL = g->h_plL.p;
    double* zdL = g->h_zdL.p;
    g->plL_of_edge.assign(El, 0);
    {
      // L-order: by landmark, then by pose index, then by creation order.  Edges normally arrive with non-decreasing
      // pose index per landmark (keyframes are created in time order): one stable counting-sort pass by landmark
      // is enough then; otherwise two passes (least significant key first)
      // (a shard sorts by the GLOBAL keyframe index of the pose, so that every rank sums a landmark's edges in the
      // order of the unsharded run)
      std::vector<int> ord(El), tmp;
      auto key = [&](int p) { return mr ? mr->sortkey[p] : p; };
      int nkey = Np;
      if (mr)
        for (int i = 0; i < Np; ++i) nkey = std::max(nkey, mr->sortkey[i] + 1);
      bool pose_sorted = true;
      {
        std::vector<int> lastp(std::max(Nl, 1), -1);
        for (int k = 0; k < El; ++k) {
          const PLEdge& e = PL[k];
          if (key(e.p) < lastp[e.l]) {
            pose_sorted = false;
            break;
          }
          lastp[e.l] = key(e.p);
        }
      }
      std::vector<int> cntl(lm_rowptr.begin(), lm_rowptr.end() - 1);
      if (pose_sorted) {
        for (int k = 0; k < El; ++k) ord[cntl[PL[k].l]++] = k;
      } else {
        tmp.resize(El);
        std::vector<int> cntp(nkey + 1, 0);
        for (int k = 0; k < El; ++k) cntp[key(PL[k].p) + 1]++;
        for (int i = 0; i < nkey; ++i) cntp[i + 1] += cntp[i];
        for (int k = 0; k < El; ++k) tmp[cntp[key(PL[k].p)]++] = k;
        for (int q = 0; q < El; ++q) ord[cntl[PL[tmp[q]].l]++] = tmp[q];
      }
      const bool planes = g->n_plane_vertices != 0;
      for (int pos = 0; pos < El; ++pos) {
        plL[pos] = PL[ord[pos]];
        if (planes) zdL[pos] = ZD[ord[pos]];
        g->plL_of_edge[ord[pos]] = pos;
      }
    }
    // the big table goes out now (page-locked source: the copy overlaps the rest of the host work)
    if (El) SSB_CUDA_CHECK(cudaMemcpyAsync(g->d_pl.p, plL, (size_t)El * sizeof(PLEdge), cudaMemcpyHostToDevice, g->stream));
    if (El && g->n_plane_vertices)
      SSB_CUDA_CHECK(cudaMemcpyAsync(g->d_pl_zd.p, zdL, (size_t)El * sizeof(double), cudaMemcpyHostToDevice, g->stream));
    tick("hessian index + L-order sort");
    // coarse aggregates: one per persistent CTA, contiguous pose ranges of C poses (multiple of 5)
    const int nblk = g->pcg_grid;
    int Cc = (n_own + nblk - 1) / nblk;
    Cc = std::max(5, ((Cc + 4) / 5) * 5);
    if (g->force_C > 0) Cc = g->force_C;   // replicated graph: every copy starts on a CTA boundary
    std::vector<int> run_lm, run_group, run_e0, lm_run_rowptr(Nl + 1, 0);
    run_lm.reserve(El / 4 + 16);
    run_group.reserve(El / 4 + 16);
    run_e0.reserve(El / 4 + 17);
    for (int l = 0; l < Nl; ++l) {
      lm_run_rowptr[l] = (int)run_lm.size();
      int prev = -1;
      for (int e = lm_rowptr[l]; e < lm_rowptr[l + 1]; ++e) {
        // ghost keyframes carry no basis: their edges ride inside the previous run and contribute exact zeros
        if (plL[e].p >= n_own) continue;
        const int grp = plL[e].p / Cc;
        if (grp != prev) {
          run_lm.push_back(l);
          run_group.push_back(grp);
          run_e0.push_back(e);
          prev = grp;
        }
      }
    }
    lm_run_rowptr[Nl] = (int)run_lm.size();
    run_e0.push_back(El);
    const int n_runs = (int)run_lm.size();
    std::vector<int> grp_run_rowptr(nblk + 1, 0), grp_runs(std::max(n_runs, 1));
    for (int r = 0; r < n_runs; ++r) grp_run_rowptr[run_group[r] + 1]++;
    for (int b = 0; b < nblk; ++b) grp_run_rowptr[b + 1] += grp_run_rowptr[b];
    {
      std::vector<int> f(grp_run_rowptr.begin(), grp_run_rowptr.end() - 1);
      for (int r = 0; r < n_runs; ++r) grp_runs[f[run_group[r]]++] = r;
    }
    // the same run structure for the middle level: (landmark, 5-pose aggregate) runs, grouped by aggregate
    const int n_agg = (Np + 4) / 5;
    std::vector<int> run1_lm, run1_agg, run1_e0, agg_run_rowptr(n_agg + 1, 0);
    run1_lm.reserve(El / 2 + 16);
    run1_agg.reserve(El / 2 + 16);
    run1_e0.reserve(El / 2 + 17);
    for (int l = 0; l < Nl; ++l) {
      int prev = -1;
      for (int e = lm_rowptr[l]; e < lm_rowptr[l + 1]; ++e) {
        if (plL[e].p >= n_own) continue;
        const int ag = plL[e].p / 5;
        if (ag != prev) {
          run1_lm.push_back(l);
          run1_agg.push_back(ag);
          run1_e0.push_back(e);
          prev = ag;
        }
      }
    }
    run1_e0.push_back(El);
    const int n_runs1 = (int)run1_lm.size();
    std::vector<int> agg_runs(std::max(n_runs1, 1));
    for (int r = 0; r < n_runs1; ++r) agg_run_rowptr[run1_agg[r] + 1]++;
    for (int a = 0; a < n_agg; ++a) agg_run_rowptr[a + 1] += agg_run_rowptr[a];
    {
      std::vector<int> f(agg_run_rowptr.begin(), agg_run_rowptr.end() - 1);
      for (int r = 0; r < n_runs1; ++r) agg_runs[f[run1_agg[r]]++] = r;
    }
    // preconditioner 3: two groups of 5-pose aggregates per CTA, coupled exactly; landmark segments = >= 2 consecutive
    // runs of one landmark whose aggregates fall into the same group (runs are ordered by landmark, then aggregate)
    const int apc = Cc / 5;                     // aggregates per CTA
    const int n_groups = 2 * nblk;
    const int n_agg_own = (n_own + 4) / 5;
    std::vector<int> grp_first_agg(n_groups + 1, 0);
    for (int b = 0; b < nblk; ++b) {
      const int ga0 = std::min(n_agg_own, b * apc), ga1 = std::min(n_agg_own, ga0 + apc);
      grp_first_agg[2 * b] = ga0;
      grp_first_agg[2 * b + 1] = ga0 + (ga1 - ga0 + 1) / 2;
    }
    grp_first_agg[n_groups] = std::min(n_agg_own, nblk * apc);
    auto group_of = [&](int ag) {
      const int b = std::min(nblk - 1, ag / apc);
      return 2 * b + (ag >= grp_first_agg[2 * b + 1] ? 1 : 0);
    };
    std::vector<int> seg_grp, seg_r0, seg_m;
    for (int r = 0; r < n_runs1;) {
      int e = r + 1;
      const int gr = group_of(run1_agg[r]);
      while (e < n_runs1 && run1_lm[e] == run1_lm[r] && group_of(run1_agg[e]) == gr) ++e;
      if (e - r >= 2) {
        seg_grp.push_back(gr);
        seg_r0.push_back(r);
        seg_m.push_back(e - r);
      }
      r = e;
    }
    std::vector<int> grp_seg_rowptr(n_groups + 1, 0), grp_seg_r0(std::max<size_t>(seg_r0.size(), 1)), grp_seg_m(std::max<size_t>(seg_r0.size(), 1)),
        grp_seg_off(std::max<size_t>(seg_r0.size(), 1), 0);
    for (int gsg : seg_grp) grp_seg_rowptr[gsg + 1]++;
    for (int q = 0; q < n_groups; ++q) grp_seg_rowptr[q + 1] += grp_seg_rowptr[q];
    {
      std::vector<int> f(grp_seg_rowptr.begin(), grp_seg_rowptr.end() - 1);
      for (size_t q = 0; q < seg_grp.size(); ++q) {
        grp_seg_r0[f[seg_grp[q]]] = seg_r0[q];
        grp_seg_m[f[seg_grp[q]]++] = seg_m[q];
      }
    }
    int grp_max_runs = 0, grp_max_seg = 0;
    for (int q = 0; q < n_groups; ++q) {
      int o = 0;
      for (int sgi = grp_seg_rowptr[q]; sgi < grp_seg_rowptr[q + 1]; ++sgi) {
        grp_seg_off[sgi] = o;
        o += grp_seg_m[sgi];
      }
      grp_max_runs = std::max(grp_max_runs, o);
      grp_max_seg = std::max(grp_max_seg, grp_seg_rowptr[q + 1] - grp_seg_rowptr[q]);
    }
    g->grp_max_runs = grp_max_runs;
    // pose-major index over L-order positions
    std::vector<int> ppl_rowptr(Np + 1, 0);
    for (int k = 0; k < El; ++k) ppl_rowptr[plL[k].p + 1]++;
    for (int i = 0; i < Np; ++i) ppl_rowptr[i + 1] += ppl_rowptr[i];
    std::vector<int> fill2(ppl_rowptr.begin(), ppl_rowptr.end() - 1);
    std::vector<int> ppl_idx(std::max(El, 1));
    for (int k = 0; k < El; ++k) ppl_idx[fill2[plL[k].p]++] = k;
    // pose-pose incidence
    std::vector<int> ppp_rowptr(Np + 1, 0);
    for (auto& e : PP) {
      ppp_rowptr[e.i + 1]++;
      ppp_rowptr[e.j + 1]++;
    }
    for (int i = 0; i < Np; ++i) ppp_rowptr[i + 1] += ppp_rowptr[i];
    std::vector<int> fill3(ppp_rowptr.begin(), ppp_rowptr.end() - 1);
    std::vector<int> ppp_idx(std::max(2 * Epp, 1)), ppp_other(std::max(2 * Epp, 1));
    for (int k = 0; k < Epp; ++k) {
      ppp_other[fill3[PP[k].i]] = PP[k].j;
      ppp_idx[fill3[PP[k].i]++] = (k << 1) | 0;
      ppp_other[fill3[PP[k].j]] = PP[k].i;
      ppp_idx[fill3[PP[k].j]++] = (k << 1) | 1;
    }
    tick("run lists + CSR");
    // allocate + upload
    SSB_TRY(g->d_pose.ensure(Np));
    SSB_TRY(g->d_pose_bak.ensure(Np));
    SSB_TRY(g->d_lm.ensure((size_t)4 * Nl));
    SSB_TRY(g->d_lm_bak.ensure((size_t)4 * Nl));
    SSB_TRY(g->d_pose_fixed.ensure(Np));
    SSB_TRY(g->d_lm_fixed.ensure(Nl));
    if (g->n_plane_vertices) {
      SSB_TRY(g->d_lm_kind.ensure(Nl));
      SSB_TRY(g->d_pl_zd.ensure(El));
    }
    SSB_TRY(g->d_pl.ensure(El));
    SSB_TRY(g->d_pp.ensure(Epp));
    SSB_TRY(g->d_lm_rowptr.ensure(Nl + 1));
    SSB_TRY(g->d_pose_pl_rowptr.ensure(Np + 1));
    SSB_TRY(g->d_pose_pl_idx.ensure(El));
    SSB_TRY(g->d_pose_pp_rowptr.ensure(Np + 1));
    SSB_TRY(g->d_pose_pp_idx.ensure((size_t)2 * Epp));
    SSB_TRY(g->d_pose_pp_other.ensure((size_t)2 * Epp));
    SSB_TRY(g->d_plP_lm.ensure(El));
    SSB_TRY(g->d_Hpp.ensure((size_t)36 * Np));
    SSB_TRY(g->d_bp.ensure((size_t)6 * Np));
    SSB_TRY(g->d_Hoff.ensure((size_t)36 * Epp));
    SSB_TRY(g->d_Hll.ensure((size_t)6 * Nl));
    SSB_TRY(g->d_bl.ensure((size_t)3 * Nl));
    SSB_TRY(g->d_HplL.ensure((size_t)18 * El));
    SSB_TRY(g->d_HplP.ensure((size_t)18 * El));
    SSB_TRY(g->d_HllInv.ensure((size_t)6 * Nl));
    SSB_TRY(g->d_Dinv.ensure((size_t)36 * Np));
    SSB_TRY(g->d_g.ensure((size_t)6 * Np));
    SSB_TRY(g->d_x.ensure((size_t)6 * (Np + 64)));
    SSB_TRY(g->d_r.ensure((size_t)6 * Np));
    SSB_TRY(g->d_z.ensure((size_t)6 * Np));
    SSB_TRY(g->d_p0.ensure((size_t)6 * (Np + 64)));
    SSB_TRY(g->d_p1.ensure((size_t)6 * Np));
    SSB_TRY(g->d_q.ensure((size_t)6 * Np));
    SSB_TRY(g->d_v.ensure((size_t)3 * (Nl + 64)));
    SSB_TRY(g->d_dl.ensure((size_t)3 * Nl));
    const size_t nb_bs = ((size_t)Np + (size_t)32 * Nl + 127) / 128 + 1;
    SSB_TRY(g->d_part.ensure(std::max<size_t>(3 * PART_STRIDE, nb_bs) + 4096));
    SSB_TRY(g->d_scalars.ensure(32));
    SSB_TRY(g->d_iscalars.ensure(4));
    SSB_TRY(g->d_ainv_ok.ensure(1));
    SSB_TRY(g->d_tmp.ensure(128));
    const int ncoarse = 6 * nblk;
    {
      // does the graph fit the on-chip resident kernel?  (<= 80 poses per CTA, bounded incidence lists,
      // one landmark per warp, <= 64 edges per landmark, bounded overflow per CTA)
      // landmark "parts": a landmark's L-order edges are cut into runs of <= 64 edges, one warp each; the parts of
      // a landmark publish W_l * (partial sum) separately and the consumers add them, so the landmark degree is
      // unbounded
      std::vector<int> part_lm, part_e0, part_e1, lm_partbase(Nl + 1, 0);
      for (int l = 0; l < Nl; ++l) {
        lm_partbase[l] = (int)part_lm.size();
        for (int e = lm_rowptr[l]; e < lm_rowptr[l + 1]; e += g->part_edges) {
          part_lm.push_back(l);
          part_e0.push_back(e);
          part_e1.push_back(std::min(e + g->part_edges, lm_rowptr[l + 1]));
        }
      }
      lm_partbase[Nl] = (int)part_lm.size();
      // a shard computes the parts of the landmarks it owns (they come first); the parts of the others arrive as cells
      const int n_parts_all = (int)part_lm.size();
      const int n_parts = lm_partbase[n_owned_lm];
      bool ok = g->allow_fast && g->use_flow && Cc <= 5 * (PCGF_THREADS / 32) && n_parts <= nblk * (PCGF_THREADS / 32) &&
                6 * (size_t)(Np + 64) + 3 * (size_t)n_parts_all < (1u << 24);   // cell indices are staged in 24 bits
      if (ok) {
        std::vector<int> ov(nblk, 0);
        for (int q = 0; q < n_parts_all && ok; ++q) {
          if (lm_partbase[part_lm[q] + 1] - lm_partbase[part_lm[q]] > 127) ok = false;   // part count rides in the top bits of an int
          if (q < n_parts) ov[q % nblk] += std::max(0, part_e1[q] - part_e0[q] - 32);
        }
        for (int b = 0; b < nblk && ok; ++b) {
          const int q0 = std::min(n_own, b * Cc), q1 = std::min(n_own, q0 + Cc);
          if (ov[b] > PCGF_MAXOV || ppl_rowptr[q1] - ppl_rowptr[q0] > PCGF_MAXPL || ppp_rowptr[q1] - ppp_rowptr[q0] > PCGF_MAXPP)
            ok = false;
        }
      }
      g->fast_ok = ok;
      tick("buffer allocation");
      // gather tables of the data-flow kernel: per CTA the distinct landmarks / pose-pose edges / external
      // neighbour poses its keyframe range touches, and per incidence the slot inside those lists
      std::vector<int> ulm_rowptr(nblk + 1, 0), ulm, pl_loc(std::max(El, 1), 0), upp_rowptr(nblk + 1, 0), upp,
          pp_loc(std::max(2 * Epp, 1), 0), pp_src(std::max(2 * Epp, 1), 0), ext_rowptr(nblk + 1, 0), ext;
      if (ok) {
        std::vector<int> lm_stamp(std::max(Nl, 1), -1), lm_slot(std::max(Nl, 1), 0), e_stamp(std::max(Epp, 1), -1),
            e_slot(std::max(Epp, 1), 0);
        ulm.reserve(El / 2 + 16);
        upp.reserve(Epp + nblk + 16);
        ext.reserve(4 * (size_t)nblk + 16);
        for (int b = 0; b < nblk && ok; ++b) {
          const int q0 = std::min(n_own, b * Cc), q1 = std::min(n_own, q0 + Cc);
          const size_t u0 = ulm.size(), p0e = upp.size(), x0 = ext.size();
          for (int kk = ppl_rowptr[q0]; kk < ppl_rowptr[q1]; ++kk) {
            const int l = plL[ppl_idx[kk]].l;
            if (lm_stamp[l] != b) {
              lm_stamp[l] = b;
              lm_slot[l] = (int)(ulm.size() - u0);
              ulm.push_back(l);
            }
            pl_loc[kk] = lm_slot[l];
          }
          for (int i = q0; i < q1; ++i)
            for (int kk = ppp_rowptr[i]; kk < ppp_rowptr[i + 1]; ++kk) {
              const int code = ppp_idx[kk], e = code >> 1, role = code & 1;
              if (e_stamp[e] != b) {
                e_stamp[e] = b;
                e_slot[e] = (int)(upp.size() - p0e);
                upp.push_back(e);
              }
              pp_loc[kk] = e_slot[e];
              const int o = role == 0 ? PP[e].j : PP[e].i;
              int src;
              if (o >= q0 && o < q1) {
                src = o - q0;
              } else {
                size_t xi = x0;
                while (xi < ext.size() && ext[xi] != o) ++xi;
                if (xi == ext.size()) ext.push_back(o);
                src = PCGW_POSES + (int)(xi - x0);
              }
              pp_src[kk] = src | (role ? (int)0x80000000u : 0);
            }
          ulm_rowptr[b + 1] = (int)ulm.size();
          upp_rowptr[b + 1] = (int)upp.size();
          ext_rowptr[b + 1] = (int)ext.size();
          if ((int)(ulm.size() - u0) > PCGW_MAXU || (int)(upp.size() - p0e) > PCGW_MAXPPE || (int)(ext.size() - x0) > PCGW_MAXEXT)
            ok = false;
        }
        g->fast_ok = ok;
      }
      if (ok) {
        cudaStream_t s2 = g->stream;
        auto up = [&](DBuf<int>& d, const std::vector<int>& h) -> int {
          SSB_TRY(d.ensure(std::max<size_t>(h.size(), 1)));
          if (!h.empty()) SSB_CUDA_CHECK(cudaMemcpyAsync(d.p, h.data(), h.size() * sizeof(int), cudaMemcpyHostToDevice, s2));
          return SSB_OK;
        };
        SSB_TRY(up(g->d_ft_ulm_rowptr, ulm_rowptr));
        SSB_TRY(up(g->d_ft_ulm, ulm));
        SSB_TRY(up(g->d_ft_pl_loc, pl_loc));
        SSB_TRY(up(g->d_ft_upp_rowptr, upp_rowptr));
        SSB_TRY(up(g->d_ft_upp, upp));
        SSB_TRY(up(g->d_ft_pp_loc, pp_loc));
        SSB_TRY(up(g->d_ft_pp_src, pp_src));
        SSB_TRY(up(g->d_ft_ext_rowptr, ext_rowptr));
        SSB_TRY(up(g->d_ft_ext, ext));
        SSB_TRY(up(g->d_ft_part_lm, part_lm));
        SSB_TRY(up(g->d_ft_part_e0, part_e0));
        SSB_TRY(up(g->d_ft_part_e1, part_e1));
        SSB_TRY(up(g->d_ft_lm_partbase, lm_partbase));
        g->flow_n_parts = n_parts;
        // pivot order of the coarse Gauss-Jordan
        std::vector<std::vector<int>> cadj(nblk);
        auto link = [&](int a, int b) {
          if (a == b || a < 0 || b < 0 || a >= nblk || b >= nblk) return;
          cadj[a].push_back(b);
          cadj[b].push_back(a);
        };
        for (auto& e : PP)
          if (e.i < n_own && e.j < n_own) link(e.i / Cc, e.j / Cc);
        for (int l = 0; l < Nl; ++l)
          for (int ra = lm_run_rowptr[l]; ra < lm_run_rowptr[l + 1]; ++ra)
            for (int rb = ra + 1; rb < lm_run_rowptr[l + 1]; ++rb) link(run_group[ra], run_group[rb]);
        for (auto& v : cadj) {
          std::sort(v.begin(), v.end());
          v.erase(std::unique(v.begin(), v.end()), v.end());
        }
        std::vector<int> all(nblk), gj_order;
        for (int b = 0; b < nblk; ++b) all[b] = b;
        std::vector<char> in(nblk, 0);
        nd_order_rec(cadj, all, in, gj_order);
        SSB_TRY(up(g->d_ft_gj_order, gj_order));
        // symbolic Gauss-Jordan: which column blocks of the pivot panel are structurally non-zero at every step
        // (bit b of gj_mask[step] = block b); receivers neither load nor update the zero blocks
        {
          const int W = (nblk + 31) / 32;
          std::vector<std::vector<unsigned>> pat(nblk, std::vector<unsigned>(W, 0u));
          for (int b = 0; b < nblk; ++b) {
            pat[b][b >> 5] |= 1u << (b & 31);
            for (int c : cadj[b]) pat[b][c >> 5] |= 1u << (c & 31);
          }
          std::vector<int> gj_mask((size_t)nblk * W, 0);
          for (int step = 0; step < nblk; ++step) {
            const int k = gj_order[step];
            for (int w = 0; w < W; ++w) gj_mask[(size_t)step * W + w] = (int)pat[k][w];
            for (int r = 0; r < nblk; ++r)
              if (r != k && ((pat[r][k >> 5] >> (k & 31)) & 1u))
                for (int w = 0; w < W; ++w) pat[r][w] |= pat[k][w];
          }
          SSB_TRY(up(g->d_ft_gj_mask, gj_mask));
        }
        SSB_CUDA_CHECK(cudaStreamSynchronize(s2));
        g->FT = FlowTabs{g->d_ft_ulm_rowptr.p, g->d_ft_ulm.p, g->d_ft_pl_loc.p, g->d_ft_upp_rowptr.p, g->d_ft_upp.p,
                         g->d_ft_pp_loc.p, g->d_ft_pp_src.p, g->d_ft_ext_rowptr.p, g->d_ft_ext.p, g->d_ft_gj_order.p, (const unsigned*)g->d_ft_gj_mask.p,
                         g->d_ft_part_lm.p, g->d_ft_part_e0.p, g->d_ft_part_e1.p, g->d_ft_lm_partbase.p, n_parts};
      }
    }
    SSB_TRY(g->d_ctacen.ensure((size_t)3 * nblk));
    SSB_TRY(g->d_Bmat.ensure((size_t)36 * Np));
    if (mr && Np) SSB_CUDA_CHECK(cudaMemsetAsync(g->d_Bmat.p, 0, (size_t)36 * Np * sizeof(double), g->stream));   // ghosts: B = 0
    SSB_TRY(g->d_B1mat.ensure((size_t)36 * Np));
    SSB_TRY(g->d_D1inv.ensure((size_t)36 * ((Np + 4) / 5 + 1)));
    SSB_TRY(g->d_Grun.ensure((size_t)18 * n_runs));
    SSB_TRY(g->d_panel.ensure((size_t)2 * (6 * ncoarse + 8)));
    SSB_TRY(g->d_run_lm.ensure(n_runs));
    SSB_TRY(g->d_run_group.ensure(n_runs));
    SSB_TRY(g->d_run_e0.ensure(n_runs + 1));
    SSB_TRY(g->d_lm_run_rowptr.ensure(Nl + 1));
    SSB_TRY(g->d_grp_run_rowptr.ensure(nblk + 1));
    SSB_TRY(g->d_grp_runs.ensure(n_runs));
    SSB_TRY(g->d_run1_lm.ensure(n_runs1));
    SSB_TRY(g->d_run1_e0.ensure(n_runs1 + 1));
    SSB_TRY(g->d_agg_run_rowptr.ensure(n_agg + 1));
    SSB_TRY(g->d_agg_runs.ensure(n_runs1));
    SSB_TRY(g->d_Grun1.ensure((size_t)18 * n_runs1));
    SSB_TRY(g->d_D1raw.ensure((size_t)36 * (n_agg + 1)));
    SSB_TRY(g->d_GrpInv.ensure((size_t)n_groups * GRP_PACK));
    SSB_TRY(g->d_grp_first_agg.ensure(n_groups + 1));
    SSB_TRY(g->d_grp_seg_rowptr.ensure(n_groups + 1));
    SSB_TRY(g->d_grp_seg_r0.ensure(grp_seg_r0.size()));
    SSB_TRY(g->d_grp_seg_m.ensure(grp_seg_m.size()));
    SSB_TRY(g->d_grp_seg_off.ensure(grp_seg_off.size()));
    SSB_TRY(g->d_run1_agg.ensure(std::max(n_runs1, 1)));
    SSB_TRY(g->d_slots.ensure((size_t)2 * nblk + 1));
    SSB_TRY(g->d_ainv.ensure((size_t)nblk * 6 * ncoarse));
    // u cells, then v cells (one triple per landmark part; parts <= Nl + El / 64)
    SSB_TRY(g->d_ucell.ensure((size_t)6 * (Np + 64) + (size_t)3 * ((size_t)Nl + El / g->part_edges + 64)));
    SSB_TRY(g->d_lines.ensure((size_t)2 * nblk * 8));
    SSB_TRY(g->d_hlpark.ensure((size_t)nblk * (PCGF_THREADS / 32) * 9 * 32));
    SSB_TRY(g->d_trace.ensure(8 * 256));
    SSB_CUDA_CHECK(cudaMemsetAsync(g->d_ucell.p, 0, g->d_ucell.cap * sizeof(uint4), g->stream));
    SSB_CUDA_CHECK(cudaMemsetAsync(g->d_lines.p, 0, g->d_lines.cap * sizeof(uint4), g->stream));
    if (g->use_flow && g->opts.preconditioner >= 1) {
      SSB_TRY(g->d_gj.ensure((size_t)nblk * (36 * (size_t)nblk + 8)));
      SSB_CUDA_CHECK(cudaMemsetAsync(g->d_gj.p, 0, g->d_gj.cap * sizeof(uint4), g->stream));
    }
    g->flow_seq = (unsigned)std::max(0, g->opts.reserved[3]);   // test hook: start close to the tag wrap-around
    g->ainv_valid = false;
    g->coarse_ready = false;
    cudaStream_t s = g->stream;
    if (n_runs) {
      SSB_CUDA_CHECK(cudaMemcpyAsync(g->d_run_lm.p, run_lm.data(), n_runs * sizeof(int), cudaMemcpyHostToDevice, s));
      SSB_CUDA_CHECK(cudaMemcpyAsync(g->d_run_group.p, run_group.data(), n_runs * sizeof(int), cudaMemcpyHostToDevice, s));
      SSB_CUDA_CHECK(cudaMemcpyAsync(g->d_grp_runs.p, grp_runs.data(), n_runs * sizeof(int), cudaMemcpyHostToDevice, s));
    }
    SSB_CUDA_CHECK(cudaMemcpyAsync(g->d_run_e0.p, run_e0.data(), (n_runs + 1) * sizeof(int), cudaMemcpyHostToDevice, s));
    if (n_runs1) {
      SSB_CUDA_CHECK(cudaMemcpyAsync(g->d_run1_lm.p, run1_lm.data(), n_runs1 * sizeof(int), cudaMemcpyHostToDevice, s));
      SSB_CUDA_CHECK(cudaMemcpyAsync(g->d_agg_runs.p, agg_runs.data(), n_runs1 * sizeof(int), cudaMemcpyHostToDevice, s));
    }
    SSB_CUDA_CHECK(cudaMemcpyAsync(g->d_run1_e0.p, run1_e0.data(), (n_runs1 + 1) * sizeof(int), cudaMemcpyHostToDevice, s));
    SSB_CUDA_CHECK(cudaMemcpyAsync(g->d_agg_run_rowptr.p, agg_run_rowptr.data(), (n_agg + 1) * sizeof(int), cudaMemcpyHostToDevice, s));
    SSB_CUDA_CHECK(cudaMemcpyAsync(g->d_grp_first_agg.p, grp_first_agg.data(), (n_groups + 1) * sizeof(int), cudaMemcpyHostToDevice, s));
    SSB_CUDA_CHECK(cudaMemcpyAsync(g->d_grp_seg_rowptr.p, grp_seg_rowptr.data(), (n_groups + 1) * sizeof(int), cudaMemcpyHostToDevice, s));
    SSB_CUDA_CHECK(cudaMemcpyAsync(g->d_grp_seg_r0.p, grp_seg_r0.data(), grp_seg_r0.size() * sizeof(int), cudaMemcpyHostToDevice, s));
    SSB_CUDA_CHECK(cudaMemcpyAsync(g->d_grp_seg_m.p, grp_seg_m.data(), grp_seg_m.size() * sizeof(int), cudaMemcpyHostToDevice, s));
    SSB_CUDA_CHECK(cudaMemcpyAsync(g->d_grp_seg_off.p, grp_seg_off.data(), grp_seg_off.size() * sizeof(int), cudaMemcpyHostToDevice, s));
    if (n_runs1) SSB_CUDA_CHECK(cudaMemcpyAsync(g->d_run1_agg.p, run1_agg.data(), n_runs1 * sizeof(int), cudaMemcpyHostToDevice, s));
    SSB_CUDA_CHECK(cudaMemcpyAsync(g->d_lm_run_rowptr.p, lm_run_rowptr.data(), (Nl + 1) * sizeof(int), cudaMemcpyHostToDevice, s));
    SSB_CUDA_CHECK(cudaMemcpyAsync(g->d_grp_run_rowptr.p, grp_run_rowptr.data(), (nblk + 1) * sizeof(int), cudaMemcpyHostToDevice, s));
    {
      CoarseDev& Cz = g->Cz;
      Cz.enabled = g->opts.preconditioner >= 1 ? 1 : 0;
      Cz.sub_enabled = g->opts.preconditioner >= 2 ? 1 : 0;
      Cz.B1mat = g->d_B1mat.p;
      Cz.D1inv = g->d_D1inv.p;
      Cz.C = Cc;
      Cz.nc = ncoarse;
      Cz.Bmat = g->d_Bmat.p;
      Cz.cen = g->d_ctacen.p;
      Cz.Grun = g->d_Grun.p;
      Cz.run_lm = g->d_run_lm.p;
      Cz.run_group = g->d_run_group.p;
      Cz.run_e0 = g->d_run_e0.p;
      Cz.lm_run_rowptr = g->d_lm_run_rowptr.p;
      Cz.grp_run_rowptr = g->d_grp_run_rowptr.p;
      Cz.grp_runs = g->d_grp_runs.p;
      Cz.n_runs = n_runs;
      Cz.panel = g->d_panel.p;
      Cz.reuse_inverse = 0;
      Cz.Grun1 = g->d_Grun1.p;
      Cz.run1_lm = g->d_run1_lm.p;
      Cz.run1_e0 = g->d_run1_e0.p;
      Cz.agg_run_rowptr = g->d_agg_run_rowptr.p;
      Cz.agg_runs = g->d_agg_runs.p;
      Cz.n_runs1 = n_runs1;
      // preconditioner 3 needs <= GRP_MAXA aggregates per group (<= 80 poses per CTA) and the on-chip kernel
      // ... and every group's landmark segments staged in the shared memory of k_grp_invert
      Cz.grp_enabled = (g->opts.preconditioner >= 3 && (apc + 1) / 2 <= GRP_MAXA && g->comm_world == 1 &&
                        grp_max_seg <= GRP_MAXSEG && grp_max_runs <= GRP_MAXRUN && (size_t)grp_max_runs * 18 * sizeof(double) <= 96 * 1024)
                           ? 1 : 0;
      Cz.grp_seg_off = g->d_grp_seg_off.p;
      Cz.run1_agg = g->d_run1_agg.p;
      Cz.D1raw = g->d_D1raw.p;
      Cz.GrpInv = g->d_GrpInv.p;
      Cz.grp_first_agg = g->d_grp_first_agg.p;
      Cz.grp_seg_rowptr = g->d_grp_seg_rowptr.p;
      Cz.grp_seg_r0 = g->d_grp_seg_r0.p;
      Cz.grp_seg_m = g->d_grp_seg_m.p;
      Cz.n_groups = n_groups;
      Cz.ainv_store = g->d_ainv.p;
      Cz.ainv_ok = g->d_ainv_ok.p;
    }
    SSB_CUDA_CHECK(cudaMemsetAsync(g->d_scalars.p, 0, 32 * sizeof(double), s));
    SSB_CUDA_CHECK(cudaMemsetAsync(g->d_iscalars.p, 0, 4 * sizeof(int), s));
    if (Np) SSB_CUDA_CHECK(cudaMemcpyAsync(g->d_pose_fixed.p, pfix.data(), Np, cudaMemcpyHostToDevice, s));
    if (Nl) SSB_CUDA_CHECK(cudaMemcpyAsync(g->d_lm_fixed.p, lfix.data(), Nl, cudaMemcpyHostToDevice, s));
    if (g->n_plane_vertices) {
      SSB_CUDA_CHECK(cudaMemcpyAsync(g->d_lm_kind.p, g->lm_kind.data(), Nl, cudaMemcpyHostToDevice, s));
    }
    if (Epp) SSB_CUDA_CHECK(cudaMemcpyAsync(g->d_pp.p, PP.data(), (size_t)Epp * sizeof(PPEdge), cudaMemcpyHostToDevice, s));
    SSB_CUDA_CHECK(cudaMemcpyAsync(g->d_lm_rowptr.p, lm_rowptr.data(), (Nl + 1) * sizeof(int), cudaMemcpyHostToDevice, s));
    SSB_CUDA_CHECK(cudaMemcpyAsync(g->d_pose_pl_rowptr.p, ppl_rowptr.data(), (Np + 1) * sizeof(int), cudaMemcpyHostToDevice, s));
    if (El) SSB_CUDA_CHECK(cudaMemcpyAsync(g->d_pose_pl_idx.p, ppl_idx.data(), (size_t)El * sizeof(int), cudaMemcpyHostToDevice, s));
    SSB_CUDA_CHECK(cudaMemcpyAsync(g->d_pose_pp_rowptr.p, ppp_rowptr.data(), (Np + 1) * sizeof(int), cudaMemcpyHostToDevice, s));
    if (Epp) SSB_CUDA_CHECK(cudaMemcpyAsync(g->d_pose_pp_idx.p, ppp_idx.data(), (size_t)2 * Epp * sizeof(int), cudaMemcpyHostToDevice, s));
    if (Epp) SSB_CUDA_CHECK(cudaMemcpyAsync(g->d_pose_pp_other.p, ppp_other.data(), (size_t)2 * Epp * sizeof(int), cudaMemcpyHostToDevice, s));
    tick("flow tables + ND order + enqueue uploads");
    SSB_CUDA_CHECK(cudaStreamSynchronize(s));  // host vectors above go out of scope
    tick("upload sync");
    g->structure_dirty = false;
    g->structure_serial++;
    g->host_est_dirty = true;
    g->have_system = false;
    g->have_snapshot = false;
    DevGraph& G = g->G;
    G.Np = Np;
    G.Nl = Nl;
    G.El = El;
    G.Epp = Epp;
    G.pose = g->d_pose.p;
    G.lm = g->d_lm.p;
    G.pose_fixed = g->d_pose_fixed.p;
    G.lm_fixed = g->d_lm_fixed.p;
    G.lm_kind = g->n_plane_vertices ? g->d_lm_kind.p : nullptr;
    G.pl_zd = g->n_plane_vertices ? g->d_pl_zd.p : nullptr;
    G.pl = g->d_pl.p;
    G.pp = g->d_pp.p;
    G.lm_rowptr = g->d_lm_rowptr.p;
    G.pose_pl_rowptr = g->d_pose_pl_rowptr.p;
    G.pose_pl_idx = g->d_pose_pl_idx.p;
    G.pose_pp_rowptr = g->d_pose_pp_rowptr.p;
    G.pose_pp_idx = g->d_pose_pp_idx.p;
    G.pose_pp_other = g->d_pose_pp_other.p;
    G.Hpp = g->d_Hpp.p;
    G.bp = g->d_bp.p;
    G.Hoff = g->d_Hoff.p;
    G.Hll = g->d_Hll.p;
    G.bl = g->d_bl.p;
    G.HplL = g->d_HplL.p;
    G.HplP = g->d_HplP.p;
    G.plP_lm = g->d_plP_lm.p;
    G.HllInv = g->d_HllInv.p;
    G.Dinv = g->d_Dinv.p;
    G.g = g->d_g.p;
    G.x = g->d_x.p;
    G.r = g->d_r.p;
    G.z = g->d_z.p;
    G.p0 = g->d_p0.p;
    G.p1 = g->d_p1.p;
    G.q = g->d_q.p;
    G.v = g->d_v.p;
    G.dl = g->d_dl.p;
    G.part = g->d_part.p;
    G.scalars = g->d_scalars.p;
    G.iscalars = g->d_iscalars.p;
    G.Np_own = n_own;
    G.lm_owned = nullptr;
    G.pose_kind = nullptr;
    if (prom) {
      SSB_TRY(g->d_pose_kind.ensure(std::max(Np, 1)));
      SSB_CUDA_CHECK(cudaMemcpyAsync(g->d_pose_kind.p, g->eff_kind.data(), Np, cudaMemcpyHostToDevice, g->stream));
      SSB_CUDA_CHECK(cudaStreamSynchronize(g->stream));
      G.pose_kind = g->d_pose_kind.p;
    }
    if (mr) {
      SSB_TRY(g->d_lm_owned.ensure(std::max(Nl, 1)));
      if (Nl) {
        SSB_CUDA_CHECK(cudaMemcpyAsync(g->d_lm_owned.p, mr->lm_owned.data(), Nl, cudaMemcpyHostToDevice, g->stream));
        SSB_CUDA_CHECK(cudaStreamSynchronize(g->stream));
      }
      G.lm_owned = g->d_lm_owned.p;
    }
  }
  if (g->host_est_dirty) {
    cudaStream_t s = g->stream;
    if (prom) refresh_promoted_estimates(g);
    if (Np) SSB_CUDA_CHECK(cudaMemcpyAsync(g->d_pose.p, POSES.data(), (size_t)Np * sizeof(Pose), cudaMemcpyHostToDevice, s));
    if (Nl) SSB_CUDA_CHECK(cudaMemcpyAsync(g->d_lm.p, g->lms.data(), (size_t)4 * Nl * sizeof(double), cudaMemcpyHostToDevice, s));
    SSB_CUDA_CHECK(cudaStreamSynchronize(s));
    g->host_est_dirty = false;
    g->device_est_newer = false;
    g->have_system = false;
  }
  return SSB_OK;
}

// ---- kernel launch helpers -----------------------------------------------------------------
static int peer_exchange(ssb_graph* g, bool with_record);
static int launch_chi2(ssb_graph* g) {
  DevGraph G = g->G;
  if (g->mr) {   // a shard sums the edges it owns; the partial sums are folded by read_scalars_sharded
    G.pl = g->mr->d_pl_own.p;
    G.El = g->mr->n_pl_own;
    G.pp = g->mr->d_pp_own.p;
    G.Epp = g->mr->n_pp_own;
    if (G.pl_zd) G.pl_zd = g->mr->d_zd_own.p;
  }
  const int nt = (G.El + CHI2_PL_TILE - 1) / CHI2_PL_TILE + (G.Epp + CHI2_PP_TILE - 1) / CHI2_PP_TILE;
  const int grid = std::max(1, std::min(nt, 4 * g->num_sms));
  k_chi2<<<grid, CHI2_THREADS, 0, g->stream>>>(G, G.part + 3 * PART_STRIDE, G.scalars + 0, G.iscalars + 2);
  g->launches++;
  SSB_CUDA_CHECK(cudaGetLastError());
  return SSB_OK;
}
// The small per-iteration kernels under-fill the GPU (a few thousand threads each), so independent ones run
// concurrently on an auxiliary stream forked from / joined to the handle's stream with events:
//   s : lin_landmarks -> [E1] -> sub_basis -> sub_runs            s2: lin_poses -> (wait E1) coarse_basis -> coarse_runs
//   s : prep_landmarks -> [E1] -> sub_assemble -> grp_invert       s2: (wait E1) prep_poses
static int launch_linearize(ssb_graph* g) {
  DevGraph& G = g->G;
  cudaStream_t s = g->stream, s2 = g->stream2;
  SSB_CUDA_CHECK(cudaEventRecord(g->ev_fork, s));
  SSB_CUDA_CHECK(cudaStreamWaitEvent(s2, g->ev_fork, 0));
  if (G.Nl) {
    k_lin_landmarks<<<(32 * G.Nl + 127) / 128, 128, 0, s>>>(G);
    g->launches++;
  }
  SSB_CUDA_CHECK(cudaEventRecord(g->ev_mid, s));
  if (G.Np) {
    k_lin_poses<<<(G.Np + 31) / 32, 64, 0, s2>>>(G);
    g->launches++;
  }
  if (g->Cz.sub_enabled && G.Np) {
    k_sub_basis<<<(G.Np + 127) / 128, 128, 0, s>>>(G, g->Cz);
    g->launches++;
    if (g->Cz.n_runs1) {
      k_sub_runs<<<(18 * g->Cz.n_runs1 + 127) / 128, 128, 0, s>>>(G, g->Cz);
      g->launches++;
    }
  }
  if (g->Cz.enabled) {
    SSB_CUDA_CHECK(cudaStreamWaitEvent(s2, g->ev_mid, 0));   // coarse_runs reads HplL written by lin_landmarks
    k_coarse_basis<<<g->pcg_grid, 256, 0, s2>>>(G, g->Cz);
    g->launches++;
    if (g->Cz.n_runs) {
      k_coarse_runs<<<(72 * g->Cz.n_runs + 127) / 128, 128, 0, s2>>>(G, g->Cz);
      g->launches++;
    }
  }
  SSB_CUDA_CHECK(cudaEventRecord(g->ev_join, s2));
  SSB_CUDA_CHECK(cudaStreamWaitEvent(s, g->ev_join, 0));
  if (g->mr && g->mr->glob && g->fast_ok) {
    // rank-level coarse level, per linearisation: everybody's centroid, then the prolongation blocks about the owner's
    // centroid (ghosts included) and the per-(landmark, rank) products
    MrCtx* mr = g->mr;
    k_g_centroid<<<1, 1024, 0, s>>>(G, mr->world, mr->rank, mr->d_gcent_all.p);
    SSB_TRY(peer_exchange(g, false));
    k_g_basis<<<(G.Np + 127) / 128, 128, 0, s>>>(G, mr->GD);
    if (G.Nl) k_g_runs<<<(18 * mr->world * G.Nl + 127) / 128, 128, 0, s>>>(G, mr->GD);
    g->launches += 3;
  }
  SSB_CUDA_CHECK(cudaGetLastError());
  g->have_system = true;
  g->linpoint_in_bak = false;
  return SSB_OK;
}
// separate_coarse: invert the coarse matrix in k_coarse_invert (third stream) instead of inside k_pcg_flow.  Measured
// on cfg2: the stand-alone kernel must leave registers for its neighbours (64 per thread) and takes 259 us against
// 181 us in-kernel, which cancels the 57 us of overlap with k_sub_assemble -> k_grp_invert, so the LM loop keeps the
// in-kernel inversion; the landmark marginals solve 3 systems per landmark with ONE matrix and reuse the stored rows.
static int launch_prep(ssb_graph* g, double lambda, bool separate_coarse = false) {
  DevGraph& G = g->G;
  cudaStream_t s = g->stream, s2 = g->stream2;
  if (G.Nl) {
    k_prep_landmarks<<<(G.Nl + 127) / 128, 128, 0, s>>>(G, lambda);
    g->launches++;
  }
  const bool fork = g->Cz.sub_enabled && g->comm_world == 1;
  // the coarse matrix of this trial is assembled and inverted beside the other per-trial kernels (third stream)
  const bool cinv = separate_coarse && fork && g->Cz.enabled && g->fast_ok && g->use_flow && g->opts.reserved[1] <= 1 && g->d_gj.p &&
                    g->pcg_grid == 148 && !g->mr;
  g->coarse_ready = false;
  if (fork) {
    SSB_CUDA_CHECK(cudaEventRecord(g->ev_fork, s));
    SSB_CUDA_CHECK(cudaStreamWaitEvent(s2, g->ev_fork, 0));
  }
  if (cinv) {
    cudaStream_t s3 = g->stream3;
    SSB_CUDA_CHECK(cudaStreamWaitEvent(s3, g->ev_fork, 0));
    if (!g->cinv_smem) {
      g->cinv_smem = ((size_t)6 * 6 * g->pcg_grid + (CINV_THREADS / 36) * 36) * sizeof(double);
      SSB_CUDA_CHECK(cudaFuncSetAttribute(k_coarse_invert<148>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g->cinv_smem));
    }
    // same tag family as the cells of the coming k_pcg_flow launch (flow_seq + 1): unique per launch
    const unsigned tag = ((g->flow_seq + 1u) << 16) + 1u;
    // its 148 CTAs wait on each other's panels: launched cooperatively (the runtime refuses the launch unless the whole
    // grid can be resident) — the kernels running beside it never wait on anything, so it always gets its SMs
    {
      uint4* gjp = g->d_gj.p;
      const int* gjo = g->FT.gj_order;
      const unsigned* gjm = g->FT.gj_mask;
      void* cargs[] = {(void*)&G, (void*)&g->Cz, (void*)&gjp, (void*)&tag, (void*)&gjo, (void*)&gjm, (void*)&lambda};
      SSB_CUDA_CHECK(cudaLaunchCooperativeKernel((void*)k_coarse_invert<148>, dim3(g->pcg_grid), dim3(CINV_THREADS), cargs, g->cinv_smem, s3));
    }
    g->launches++;
    SSB_CUDA_CHECK(cudaEventRecord(g->ev_join3, s3));
    g->coarse_ready = true;
  }
  k_prep_poses<<<(G.Np_own + 63) / 64, 64, 0, fork ? s2 : s>>>(G, lambda);
  g->launches++;
  if (fork) {
    k_sub_assemble<<<(G.Np_own + 4) / 5, SUBA_THREADS, 0, s>>>(G, g->Cz, lambda, (g->Cz.grp_enabled && g->fast_ok) ? 0 : 1);
    g->launches++;
    if (g->Cz.grp_enabled && g->fast_ok) {
      const size_t dsm = (size_t)std::max(g->grp_max_runs, 1) * 18 * sizeof(double);
      if (dsm > 40 * 1024)
        SSB_CUDA_CHECK(cudaFuncSetAttribute(k_grp_invert, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsm));
      k_grp_invert<<<g->Cz.n_groups, GRP_THREADS, dsm, s>>>(G, g->Cz, g->grp_max_runs);
      g->launches++;
    }
    SSB_CUDA_CHECK(cudaEventRecord(g->ev_join, s2));
    SSB_CUDA_CHECK(cudaStreamWaitEvent(s, g->ev_join, 0));
  }
  if (cinv) SSB_CUDA_CHECK(cudaStreamWaitEvent(s, g->ev_join3, 0));
  if (g->mr && g->mr->glob && g->fast_ok) {
    // rank-level coarse level, per damped trial: my 6 rows of A_g -> everybody; invert the 6W x 6W matrix redundantly
    MrCtx* mr = g->mr;
    k_g_rows<<<G_ROW_BLOCKS, G_ROW_THREADS, 0, s>>>(G, mr->GD, lambda);
    k_g_fold<<<1, G_ROW_THREADS, 0, s>>>(mr->GD, G_ROW_BLOCKS, mr->d_grows_all.p);
    SSB_TRY(peer_exchange(g, false));
    k_g_invert<<<1, 256, 0, s>>>(mr->GD, 1.0f);
    g->launches += 3;
  }
  SSB_CUDA_CHECK(cudaGetLastError());
  return SSB_OK;
}
static int launch_pcg(ssb_graph* g, double lambda);
static int peer_exchange(ssb_graph* g, bool with_record);
static int launch_solve(ssb_graph* g, double lambda, int apply) {
  DevGraph& G = g->G;
  cudaStream_t s = g->stream;
  SSB_TRY(launch_prep(g, lambda));
  SSB_TRY(launch_pcg(g, lambda));
  // sharded: the solution of the ghost keyframes was pushed by their owners at the end of the PCG kernel
  if (g->mr) SSB_TRY(peer_exchange(g, false));
  if (apply) {
    k_backsub_update<<<(G.Np + 32 * G.Nl + 127) / 128, 128, 0, s>>>(G, lambda, g->d_pose_bak.p, g->d_lm_bak.p);
    g->launches++;
    g->linpoint_in_bak = true;   // the backup taken before the update = the linearisation point of this iteration
  }
  SSB_CUDA_CHECK(cudaGetLastError());
  return SSB_OK;
}
static int launch_pcg(ssb_graph* g, double lambda) {
  DevGraph& G = g->G;
  cudaStream_t s = g->stream;
  double tol2 = g->opts.pcg_tol * g->opts.pcg_tol;
  int maxit = g->opts.max_pcg_iters;
  // coarse inverse refresh policy: opts.reserved[1] = n > 0 re-inverts A_c only every n-th solve and
  // reuses the stored (stale but SPD) inverse in between — the PCG solution is unaffected, only its
  // iteration count can change.  Default 1 = always fresh.
  {
    const int every = std::max(1, g->opts.reserved[1]);
    const bool reuse = g->Cz.enabled && g->ainv_valid && every > 1 && (g->solves_since_refresh % every) != 0;
    g->Cz.reuse_inverse = g->coarse_ready ? 2 : (reuse ? 1 : 0);
    if (!reuse) g->solves_since_refresh = 0;
    g->solves_since_refresh++;
    if (g->Cz.enabled) g->ainv_valid = true;
  }
  BarSlot* slots = g->d_slots.p;
  SSB_CUDA_CHECK(cudaMemsetAsync(slots, 0, ((size_t)2 * g->pcg_grid + 1) * sizeof(BarSlot), s));
  StreamPeer SP{};
  if (g->mr) SP = g->mr->SP;
  void* args[] = {(void*)&G, (void*)&g->Cz, (void*)&slots, (void*)&lambda, (void*)&tol2, (void*)&maxit, (void*)&SP};
  if (g->ev_used + 2 > g->ev_pool.size()) {
    for (int k = 0; k < 64; ++k) {
      cudaEvent_t e;
      SSB_CUDA_CHECK(cudaEventCreate(&e));
      g->ev_pool.push_back(e);
    }
  }
  SSB_CUDA_CHECK(cudaEventRecord(g->ev_pool[g->ev_used], s));
  if (g->fast_ok) {
    // tags = (seq << 16) + iteration: unique per launch, so the cell buffers are never cleared between solves
    if (++g->flow_seq >= 0xFFFFu) {
      if (g->mr) SSB_TRY(peer_exchange(g, false));   // nobody may still be pushing cells of the old tag family
      SSB_CUDA_CHECK(cudaMemsetAsync(g->d_ucell.p, 0, g->d_ucell.cap * sizeof(uint4), s));
      SSB_CUDA_CHECK(cudaMemsetAsync(g->d_lines.p, 0, g->d_lines.cap * sizeof(uint4), s));
      if (g->d_gj.p) SSB_CUDA_CHECK(cudaMemsetAsync(g->d_gj.p, 0, g->d_gj.cap * sizeof(uint4), s));
      if (g->mr) SSB_TRY(peer_exchange(g, false));   // ... nor start before everybody has cleared
      g->flow_seq = 1;
    }
    FlowBufs F{g->d_ucell.p, g->d_ucell.p + (size_t)6 * (G.Np + 64), g->d_lines.p, g->d_gj.p, g->flow_seq << 16, g->d_hlpark.p, g->d_trace.p,
               g->rep_ctas, g->rep_count};
    int maxit_f = std::min(maxit, 60000);
    FlowPeer FP{};
    if (g->mr) FP = g->mr->FP;
    void* fargs[] = {(void*)&G, (void*)&g->Cz, (void*)&slots, (void*)&F, (void*)&g->FT, (void*)&lambda, (void*)&tol2, (void*)&maxit_f, (void*)&FP};
    // shards sharing one GPU: two cooperative launches do not overlap, and the kernel needs co-residency with the
    // OTHER ranks' grids, not cg::grid.sync — launch it as a plain kernel (all the grids together fit the SMs)
    if (g->mr && g->pcg_grid < g->num_sms) {
      SSB_CUDA_CHECK(cudaLaunchKernel(flow_kernel(g->pcg_grid, true), dim3(g->pcg_grid), dim3(PCGF_THREADS), fargs, g->pcgw_smem, s));
      // ... and nothing may be queued BEHIND a kernel that waits for another rank of this process: streams share
      // hardware queues (CUDA_DEVICE_MAX_CONNECTIONS), and a queue whose head depends on the waiting kernel would
      // hold back the other rank's launches (false dependency => deadlock)
      SSB_CUDA_CHECK(cudaStreamSynchronize(s));
    } else
      SSB_CUDA_CHECK(cudaLaunchCooperativeKernel(flow_kernel(g->pcg_grid, g->mr != nullptr, g->rep_count > 1), dim3(g->pcg_grid), dim3(PCGF_THREADS),
                                                 fargs, g->pcgw_smem, s));
  } else if (g->mr) {
    // the cross-rank barrier slots of the streaming kernel start every launch from zero on every rank
    SSB_CUDA_CHECK(cudaMemsetAsync(g->mr->arena + g->mr->lay[g->mr->rank].slots, 0, ((size_t)2 * g->mr->world * g->pcg_grid + 1) * sizeof(BarSlot), s));
    SSB_TRY(peer_exchange(g, false));
    if (g->pcg_grid < g->num_sms) {   // shards sharing one GPU: see the data-flow kernel above
      SSB_CUDA_CHECK(cudaLaunchKernel((void*)k_pcg<true>, dim3(g->pcg_grid), dim3(PCG_THREADS), args, g->pcg_smem, s));
      SSB_CUDA_CHECK(cudaStreamSynchronize(s));
    } else
      SSB_CUDA_CHECK(cudaLaunchCooperativeKernel((void*)k_pcg<true>, dim3(g->pcg_grid), dim3(PCG_THREADS), args, g->pcg_smem, s));
  } else
    SSB_CUDA_CHECK(cudaLaunchCooperativeKernel((void*)k_pcg<false>, dim3(g->pcg_grid), dim3(PCG_THREADS), args, g->pcg_smem, s));
  SSB_CUDA_CHECK(cudaEventRecord(g->ev_pool[g->ev_used + 1], s));
  g->ev_used += 2;
  g->launches++;
  SSB_CUDA_CHECK(cudaGetLastError());
  return SSB_OK;
}
// ============================================================================================================
// Sharded graphs: ONE graph over several ranks by contiguous keyframe range.
//
// Every rank holds the full host graph (the caller replays the same add_* calls everywhere) and derives the same
// ShardPlan.  Rank r then builds its LOCAL SUBGRAPH as a handle of its own (g->shard):
//   keyframes : its own range [ps, pe) first, then "ghosts" = keyframes of other ranks that observe a landmark one of
//               ours observes, or are pose-pose neighbours of ours;
//   landmarks : every landmark one of its keyframes observes ("touched"), with ALL edges of those landmarks;
//               a landmark is eliminated ("owned") by the rank that owns its first observer.
// All per-iteration kernels run unchanged on that subgraph: a touched landmark has all its edges here, so H_ll, b_l,
// (H_ll + lambda I)^-1 and the back-substitution are computed redundantly and bit-identically by every rank that needs
// them — nothing of the linearisation is exchanged.  Ghost keyframes have no basis (B = 0) in the aggregate levels,
// which makes the preconditioner of rank r the one of S restricted to its own keyframes (block-Jacobi across ranks).
// What crosses NVLink (written by the producer straight into the consumer's arena, ssb_peer.cuh):
//   per PCG iteration : u of own keyframes that are ghosts elsewhere (6 cells each), v of owned landmark parts touched
//                       elsewhere (3 cells each), 2 cells (r'u, w'u) per CTA to every other rank;
//   per damped trial  : the solution x of those keyframes; one 64-byte record per rank (chi2, scale, ...);
//   per optimize()    : the final estimates of own keyframes / owned landmarks to every rank (full replica).
// ============================================================================================================
static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static ArenaLayout arena_layout(int world, int nb, int NpL, int NlL, int ElL, int NpG, int NlG) {
  ArenaLayout L{};
  size_t o = 0;
  L.flags = o;
  o = align_up(o + (size_t)world * sizeof(unsigned long long), 256);
  L.red = o;
  o = align_up(o + (size_t)world * PEER_RED_N * sizeof(double), 256);
  L.n_ucells = (size_t)6 * (NpL + 64);
  L.n_cells = L.n_ucells + (size_t)3 * ((size_t)NlL + ElL / 64 + 64);
  L.cells = o;
  o = align_up(o + L.n_cells * sizeof(uint4), 256);
  L.n_lines = (size_t)2 * world * nb * 8;
  L.lines = o;
  o = align_up(o + L.n_lines * sizeof(uint4), 256);
  L.n_x = (size_t)6 * (NpL + 64);
  L.x = o;
  o = align_up(o + L.n_x * sizeof(double), 256);
  L.z = o;
  o = align_up(o + L.n_x * sizeof(double), 256);
  L.n_v = (size_t)3 * (NlL + 64);
  L.v = o;
  o = align_up(o + L.n_v * sizeof(double), 256);
  L.slots = o;
  o = align_up(o + ((size_t)2 * world * nb + 1) * sizeof(BarSlot), 256);
  L.pose_full = o;
  o = align_up(o + (size_t)std::max(NpG, 1) * sizeof(Pose), 256);
  L.lm_full = o;
  o = align_up(o + (size_t)std::max(NlG, 1) * 4 * sizeof(double), 256);
  L.gcent = o;
  o = align_up(o + (size_t)world * 4 * sizeof(double), 256);
  L.grows = o;
  o = align_up(o + (size_t)world * 36 * world * sizeof(double), 256);
  L.glines = o;
  o = align_up(o + L.n_lines * sizeof(uint4), 256);
  L.total = o;
  return L;
}

struct EdgeIdx {
  int a, b;
};
static void build_plan_raw(int Np, int Nl, const std::vector<EdgeIdx>& pl, const std::vector<EdgeIdx>& pp, int world, int nb, ShardPlan& P);
static void build_plan(const ssb_graph* g, int world, int nb, ShardPlan& P) {
  std::vector<EdgeIdx> pl(g->pl.size()), pp(g->pp.size());
  for (size_t k = 0; k < g->pl.size(); ++k) pl[k] = {g->pl[k].p, g->pl[k].l};
  for (size_t k = 0; k < g->pp.size(); ++k) pp[k] = {g->pp[k].i, g->pp[k].j};
  build_plan_raw((int)g->poses.size(), (int)(g->lms.size() / 4), pl, pp, world, nb, P);
}
// pl: (keyframe, landmark) index pairs; pp: (keyframe i, keyframe j) index pairs — indices within their kind
static void build_plan_raw(int Np, int Nl, const std::vector<EdgeIdx>& pl, const std::vector<EdgeIdx>& pp, int world, int nb, ShardPlan& P) {
  P.world = world;
  P.nb = nb;
  P.Np = Np;
  P.Nl = Nl;
  P.R.assign(world, RankLocal());
  const int per = std::max(1, (Np + world - 1) / world);
  auto rank_of = [&](int p) { return std::min(world - 1, p / per); };
  for (int r = 0; r < world; ++r) {
    P.R[r].ps = std::min(Np, r * per);
    P.R[r].pe = r == world - 1 ? Np : std::min(Np, (r + 1) * per);
  }
  P.lm_deg.assign(Nl, 0);
  std::vector<int> first(Nl, Np);
  for (auto& e : pl) {
    P.lm_deg[e.b]++;
    first[e.b] = std::min(first[e.b], e.a);
  }
  P.lm_owner.assign(Nl, 0);
  for (int l = 0; l < Nl; ++l) P.lm_owner[l] = first[l] < Np ? rank_of(first[l]) : 0;
  // touched[r][l]
  std::vector<std::vector<unsigned char>> touched(world, std::vector<unsigned char>(std::max(Nl, 1), 0));
  for (auto& e : pl) touched[rank_of(e.a)][e.b] = 1;
  for (int l = 0; l < Nl; ++l) touched[P.lm_owner[l]][l] = 1;   // (a landmark without edges stays with rank 0)
  for (int r = 0; r < world; ++r) {
    RankLocal& R = P.R[r];
    std::vector<unsigned char> need(std::max(Np, 1), 0);
    for (auto& e : pl)
      if (touched[r][e.b]) need[e.a] = 1;
    for (auto& e : pp) {
      const bool io = e.a >= R.ps && e.a < R.pe, jo = e.b >= R.ps && e.b < R.pe;
      if (io || jo) need[e.a] = need[e.b] = 1;
    }
    R.l2g_pose.clear();
    for (int p = R.ps; p < R.pe; ++p) R.l2g_pose.push_back(p);
    for (int p = 0; p < Np; ++p)
      if (need[p] && (p < R.ps || p >= R.pe)) R.l2g_pose.push_back(p);
    R.g2l_pose.assign(std::max(Np, 1), -1);
    for (size_t k = 0; k < R.l2g_pose.size(); ++k) R.g2l_pose[R.l2g_pose[k]] = (int)k;
    R.l2g_lm.clear();
    for (int l = 0; l < Nl; ++l)
      if (touched[r][l] && P.lm_owner[l] == r) R.l2g_lm.push_back(l);
    R.n_owned_lm = (int)R.l2g_lm.size();
    for (int l = 0; l < Nl; ++l)
      if (touched[r][l] && P.lm_owner[l] != r) R.l2g_lm.push_back(l);
    R.g2l_lm.assign(std::max(Nl, 1), -1);
    R.partbase.assign(R.l2g_lm.size() + 1, 0);
    for (size_t k = 0; k < R.l2g_lm.size(); ++k) {
      R.g2l_lm[R.l2g_lm[k]] = (int)k;
      R.partbase[k + 1] = R.partbase[k] + (P.lm_deg[R.l2g_lm[k]] + 63) / 64;
    }
  }
}

// ---- rank <-> rank exchange of small host blobs (arena handles): NCCL between processes, a registry between threads
struct LocalGroup {
  std::mutex m;
  std::condition_variable cv;
  int arrived = 0;
  unsigned long long gen = 0;
  PeerBlob in[SSB_MAX_WORLD], out[SSB_MAX_WORLD];
};
static std::mutex g_groups_mutex;
static std::map<std::string, LocalGroup*> g_groups;
static int exchange_blobs(ssb_graph* g, const PeerBlob& mine, PeerBlob* all) {
  const int world = g->comm_world, rank = g->comm_rank;
  if (g->local_group) {
    LocalGroup* grp;
    {
      std::lock_guard<std::mutex> lk(g_groups_mutex);
      LocalGroup*& slot = g_groups[g->local_key];
      if (!slot) slot = new LocalGroup();
      grp = slot;
    }
    std::unique_lock<std::mutex> lk(grp->m);
    grp->in[rank] = mine;
    const unsigned long long gen = grp->gen;
    if (++grp->arrived == world) {
      grp->arrived = 0;
      for (int r = 0; r < world; ++r) grp->out[r] = grp->in[r];
      grp->gen++;
      grp->cv.notify_all();
    } else if (!grp->cv.wait_for(lk, std::chrono::seconds(120), [&] { return grp->gen != gen; })) {
      grp->arrived--;
      set_error("sharded graph: the other ranks of group '%s' did not reach the same call within 120 s", g->local_key.c_str());
      return SSB_ERR_COMM;
    }
    for (int r = 0; r < world; ++r) all[r] = grp->out[r];
    return SSB_OK;
  }
  NcclApi& N = nccl_api();
  if (!g->comm || !N.ok) {
    set_error("sharded graph: no communicator attached");
    return SSB_ERR_COMM;
  }
  MrCtx* mr = g->shard->mr;
  DBuf<unsigned char>& stage = g->shard->d_blob_stage;
  SSB_TRY(stage.ensure((size_t)(world + 1) * sizeof(PeerBlob)));
  cudaStream_t s = g->shard->stream;
  SSB_CUDA_CHECK(cudaMemcpyAsync(stage.p + (size_t)world * sizeof(PeerBlob), &mine, sizeof(PeerBlob), cudaMemcpyHostToDevice, s));
  SSB_NCCL_CHECK(N.AllGather(stage.p + (size_t)world * sizeof(PeerBlob), stage.p, sizeof(PeerBlob), /*ncclInt8*/ 0, g->comm, s));
  SSB_CUDA_CHECK(cudaMemcpyAsync(all, stage.p, (size_t)world * sizeof(PeerBlob), cudaMemcpyDeviceToHost, s));
  SSB_CUDA_CHECK(cudaStreamSynchronize(s));
  (void)mr;
  return SSB_OK;
}

// (re)allocate this rank's arena if needed, tell everybody, map everybody's
static int map_arenas(ssb_graph* g, const ShardPlan& plan, const std::vector<int>& ElL) {
  ssb_graph* in = g->shard;
  MrCtx* mr = in->mr;
  const int world = plan.world, rank = g->comm_rank;
  for (int r = 0; r < world; ++r)
    mr->lay[r] = arena_layout(world, plan.nb, (int)plan.R[r].l2g_pose.size(), (int)plan.R[r].l2g_lm.size(), ElL[r], plan.Np, plan.Nl);
  const size_t need = mr->lay[rank].total;
  // arenas replaced two exchanges ago are unmapped everywhere by now
  if (mr->retired.size() > 1) {
    cudaFree(mr->retired.front());
    mr->retired.erase(mr->retired.begin());
  }
  if (need > mr->arena_cap || !mr->arena) {
    if (mr->arena) mr->retired.push_back(mr->arena);
    const size_t want = need + need / 2 + (1u << 20);
    mr->arena = nullptr;
    SSB_CUDA_CHECK(cudaMalloc((void**)&mr->arena, want));
    mr->arena_cap = want;
    mr->arena_serial++;
  }
  // flags / records / cells start from zero whenever the structure changed (epochs restart with them)
  SSB_CUDA_CHECK(cudaMemsetAsync(mr->arena, 0, mr->lay[rank].total, in->stream));
  SSB_CUDA_CHECK(cudaStreamSynchronize(in->stream));
  mr->epoch = 0;
  PeerBlob mine{};
  mine.pid = (int)getpid();
  mine.device = in->device;
  mine.ptr = mr->arena;
  mine.bytes = mr->arena_cap;
  mine.serial = mr->arena_serial;
  if (!g->local_group) SSB_CUDA_CHECK(cudaIpcGetMemHandle(&mine.handle, mr->arena));
  PeerBlob all[SSB_MAX_WORLD];
  SSB_TRY(exchange_blobs(g, mine, all));
  for (int r = 0; r < world; ++r) {
    if (r == rank) {
      mr->peer_base[r] = mr->arena;
      continue;
    }
    const PeerBlob& b = all[r];
    if (b.bytes < mr->lay[r].total) {
      set_error("sharded graph: rank %d holds a different graph (arena %llu < %zu bytes)", r, b.bytes, mr->lay[r].total);
      return SSB_ERR_COMM;
    }
    const bool same = mr->peer_mapped[r] && mr->peer_blob[r].pid == b.pid && mr->peer_blob[r].serial == b.serial && mr->peer_blob[r].ptr == b.ptr;
    if (same) continue;
    if (mr->peer_mapped[r] && mr->peer_blob[r].pid != (int)getpid()) cudaIpcCloseMemHandle(mr->peer_base[r]);
    mr->peer_mapped[r] = false;
    if (b.pid == (int)getpid()) {
      if (b.device != in->device) {
        int can = 0;
        SSB_CUDA_CHECK(cudaDeviceCanAccessPeer(&can, in->device, b.device));
        if (!can) {
          set_error("sharded graph: device %d cannot access device %d", in->device, b.device);
          return SSB_ERR_COMM;
        }
        cudaError_t e = cudaDeviceEnablePeerAccess(b.device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) SSB_CUDA_CHECK(e);
        cudaGetLastError();
      }
      mr->peer_base[r] = (unsigned char*)b.ptr;
    } else {
      void* q = nullptr;
      SSB_CUDA_CHECK(cudaIpcOpenMemHandle(&q, b.handle, cudaIpcMemLazyEnablePeerAccess));
      mr->peer_base[r] = (unsigned char*)q;
    }
    mr->peer_blob[r] = b;
    mr->peer_mapped[r] = true;
  }
  PeerDev& P = mr->P;
  P.world = world;
  P.rank = rank;
  for (int r = 0; r < world; ++r) {
    unsigned char* base = mr->peer_base[r];
    P.flags[r] = (unsigned long long*)(base + mr->lay[r].flags);
    P.red[r] = (double*)(base + mr->lay[r].red);
    P.lines[r] = (uint4*)(base + mr->lay[r].lines);
    P.slots[r] = base + mr->lay[r].slots;
    mr->PG.pose_full[r] = base + mr->lay[r].pose_full;
    mr->PG.lm_full[r] = (double*)(base + mr->lay[r].lm_full);
    mr->FP.lines[r] = P.lines[r];
    mr->FP.glines[r] = (uint4*)(base + mr->lay[r].glines);
  }
  mr->FP.gcent = (const double*)(mr->arena + mr->lay[rank].gcent);
  mr->PG.world = world;
  mr->FP.world = world;
  mr->FP.rank = rank;
  return SSB_OK;
}

// barrier (+ small all-gather of the reduction results when with_record) across the ranks, on the handle's stream
static int peer_exchange(ssb_graph* in, bool with_record) {
  MrCtx* mr = in->mr;
  ++mr->epoch;
  k_peer_exchange<<<1, 32, 0, in->stream>>>(mr->P, with_record ? in->G.scalars : nullptr, in->G.iscalars, mr->epoch, mr->d_err.p);
  in->launches++;
  SSB_CUDA_CHECK(cudaGetLastError());
  if (in->pcg_grid < in->num_sms) SSB_CUDA_CHECK(cudaStreamSynchronize(in->stream));   // shards sharing one GPU: see launch_pcg
  return SSB_OK;
}
// the sharded counterpart of read_scalars: exchange the per-rank records and fold them in rank order on the host
// (identical bits on every rank => identical accept / reject decisions)
static int read_scalars_sharded(ssb_graph* in) {
  MrCtx* mr = in->mr;
  SSB_TRY(peer_exchange(in, true));
  SSB_CUDA_CHECK(cudaMemcpyAsync(mr->h_red, mr->P.red[mr->rank], (size_t)mr->world * PEER_RED_N * sizeof(double), cudaMemcpyDeviceToHost, in->stream));
  SSB_CUDA_CHECK(cudaMemcpyAsync(mr->h_err, mr->d_err.p, sizeof(int), cudaMemcpyDeviceToHost, in->stream));
  SSB_CUDA_CHECK(cudaStreamSynchronize(in->stream));
  if (*mr->h_err) {
    set_error("sharded graph: rank %d did not arrive at the exchange (timeout)", *mr->h_err - 1);
    return SSB_ERR_COMM;
  }
  double chi = 0.0, scale = 0.0, md = 0.0;
  int status = 0;
  for (int r = 0; r < mr->world; ++r) {
    const double* rec = mr->h_red + (size_t)r * PEER_RED_N;
    chi += rec[PR_CHI2];
    scale += rec[PR_SCALE];
    md = std::max(md, rec[PR_MAXDIAG]);
    status = std::max(status, (int)rec[PR_PCG_STATUS]);
  }
  const double* me = mr->h_red + (size_t)mr->rank * PEER_RED_N;
  in->h_scalars[0] = chi;
  in->h_scalars[1] = scale;
  in->h_scalars[2] = md;
  in->h_scalars[3] = me[PR_GAMMA];
  in->h_scalars[4] = me[PR_GAMMA0];
  in->h_iscalars[0] = (int)me[PR_PCG_ITERS];
  in->h_iscalars[1] = status;
  return SSB_OK;
}

// own estimates -> the full replica of every rank
__global__ void k_gather_push(DevGraph G, PeerGather PG, const int* own_l2g, int n_own, const int* ownlm_l2g, int n_owned_lm) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n_own) {
    const Pose X = G.pose[t];
    for (int r = 0; r < PG.world; ++r) ((Pose*)PG.pose_full[r])[own_l2g[t]] = X;
  } else if (t < n_own + n_owned_lm) {
    const int l = t - n_own;
    const double4 v = reinterpret_cast<const double4*>(G.lm)[l];
    for (int r = 0; r < PG.world; ++r) reinterpret_cast<double4*>(PG.lm_full[r])[ownlm_l2g[l]] = v;
  }
}

static int prepare(ssb_graph* g);
// outer handle: (re)build the plan, this rank's subgraph handle, the arenas and the push tables
static int prepare_sharded(ssb_graph* g) {
  SSB_CUDA_CHECK(cudaSetDevice(g->device));
  if (!g->ll.empty()) {
    set_error("landmark-landmark (EdgePointXYZ) edges are not supported on a sharded graph");
    return SSB_ERR_INVALID;
  }
  const int world = g->comm_world, rank = g->comm_rank;
  if (!g->shard) {
    ssb_graph_opts o = g->opts;
    o.device = g->device;
    o.reserved[2] = g->opts.reserved[2] > 0 ? g->opts.reserved[2] : g->pcg_grid;
    g->shard = ssb_graph_create(&o);
    if (!g->shard) return SSB_ERR_CUDA;
    g->shard->mr = new MrCtx();
    MrCtx* mr = g->shard->mr;
    mr->world = world;
    mr->rank = rank;
    mr->nb = g->shard->pcg_grid;
    SSB_CUDA_CHECK(cudaMallocHost((void**)&mr->h_red, (size_t)SSB_MAX_WORLD * PEER_RED_N * sizeof(double)));
    SSB_CUDA_CHECK(cudaMallocHost((void**)&mr->h_err, sizeof(int)));
    SSB_TRY(mr->d_err.ensure(1));
    SSB_CUDA_CHECK(cudaMemset(mr->d_err.p, 0, sizeof(int)));
  }
  ssb_graph* in = g->shard;
  MrCtx* mr = in->mr;
  if (g->structure_dirty) {
    const int Np = (int)g->poses.size(), Nl = (int)(g->lms.size() / 4);
    if (Np < world) {
      set_error("sharded graph: %d keyframes cannot be split over %d ranks", Np, world);
      return SSB_ERR_INVALID;
    }
    if (!g->plan) g->plan = new ShardPlan();
    ShardPlan& plan = *g->plan;
    build_plan(g, world, mr->nb, plan);
    const RankLocal& R = plan.R[rank];
    // hessian indices of the full graph (ssb_graph_hessian_index, marginals)
    int h = 0;
    for (auto& v : g->V) v.hidx = v.fixed ? -1 : h++;
    // ---- the local subgraph ----
    const int NpL = (int)R.l2g_pose.size(), NlL = (int)R.l2g_lm.size();
    in->V.clear();
    in->E.clear();
    in->poses.resize(NpL);
    in->lms.resize((size_t)4 * NlL);
    in->lm_kind.assign(NlL, 0);
    in->pose_vid.clear();
    in->lm_vid.clear();
    in->pl.clear();
    in->pl_zd.clear();
    in->pp.clear();
    in->n_plane_vertices = 0;
    for (int k = 0; k < NpL; ++k) {
      const int gp = R.l2g_pose[k];
      HostVertex v{VK_SE3, k, g->V[g->pose_vid[gp]].fixed, -1};
      in->pose_vid.push_back((int)in->V.size());
      in->V.push_back(v);
    }
    for (int k = 0; k < NlL; ++k) {
      const int gl = R.l2g_lm[k];
      HostVertex v{g->lm_kind[gl] ? VK_PLANE : VK_XYZ, k, g->V[g->lm_vid[gl]].fixed, -1};
      in->lm_kind[k] = g->lm_kind[gl];
      if (g->lm_kind[gl]) in->n_plane_vertices++;
      in->lm_vid.push_back((int)in->V.size());
      in->V.push_back(v);
    }
    std::vector<int> ElL(world, 0);
    for (auto& e : g->pl)
      for (int r = 0; r < world; ++r)
        if (plan.R[r].g2l_lm[e.l] >= 0) ElL[r]++;
    std::vector<PLEdge> pl_own;
    std::vector<double> zd_own;
    for (size_t k = 0; k < g->pl.size(); ++k) {
      const PLEdge& e = g->pl[k];
      const int ll = R.g2l_lm[e.l];
      if (ll < 0) continue;
      PLEdge le = e;
      le.p = R.g2l_pose[e.p];
      le.l = ll;
      in->pl.push_back(le);
      in->pl_zd.push_back(g->pl_zd[k]);
      in->E.push_back({EK_PL, (int)in->pl.size() - 1});
      if (e.p >= R.ps && e.p < R.pe) {
        pl_own.push_back(le);
        zd_own.push_back(g->pl_zd[k]);
      }
    }
    std::vector<PPEdge> pp_own;
    for (auto& e : g->pp) {
      const bool io = e.i >= R.ps && e.i < R.pe, jo = e.j >= R.ps && e.j < R.pe;
      if (!io && !jo) continue;
      PPEdge le = e;
      le.i = R.g2l_pose[e.i];
      le.j = R.g2l_pose[e.j];
      in->pp.push_back(le);
      in->E.push_back({EK_PP, (int)in->pp.size() - 1});
      if (io) pp_own.push_back(le);   // a boundary edge is counted by the owner of its first vertex
    }
    mr->n_own = R.pe - R.ps;
    mr->n_owned_lm = R.n_owned_lm;
    mr->sortkey = R.l2g_pose;
    mr->lm_owned.assign(std::max(NlL, 1), 0);
    for (int k = 0; k < R.n_owned_lm; ++k) mr->lm_owned[k] = 1;
    in->structure_dirty = true;
    // ---- arenas: the buffers a neighbour writes into are windows of the arena ----
    SSB_TRY(map_arenas(g, plan, ElL));
    const ArenaLayout& L = mr->lay[rank];
    in->d_ucell.set_view(mr->arena + L.cells, L.n_cells);
    in->d_lines.set_view(mr->arena + L.lines, L.n_lines);
    in->d_x.set_view(mr->arena + L.x, L.n_x);
    in->d_z.set_view(mr->arena + L.z, L.n_x);
    in->d_v.set_view(mr->arena + L.v, L.n_v);
    // ---- push tables ----
    {
      std::vector<int> urow(mr->n_own + 1, 0);
      std::vector<uint4*> ucellp;
      std::vector<double*> uxp, uzp, svp;
      std::vector<int> svrow(R.n_owned_lm + 1, 0);
      for (int k = 0; k < mr->n_own; ++k) {
        const int gp = R.l2g_pose[k];
        for (int r = 0; r < world; ++r) {
          if (r == rank) continue;
          const int li = plan.R[r].g2l_pose[gp];
          if (li < 0) continue;
          ucellp.push_back((uint4*)(mr->peer_base[r] + mr->lay[r].cells) + 6 * (size_t)li);
          uxp.push_back((double*)(mr->peer_base[r] + mr->lay[r].x) + 6 * (size_t)li);
          uzp.push_back((double*)(mr->peer_base[r] + mr->lay[r].z) + 6 * (size_t)li);
        }
        urow[k + 1] = (int)ucellp.size();
      }
      const int n_own_parts = R.partbase[R.n_owned_lm];
      std::vector<int> vrow(n_own_parts + 1, 0);
      std::vector<uint4*> vcellp;
      for (int k = 0; k < R.n_owned_lm; ++k) {
        const int gl = R.l2g_lm[k];
        for (int r = 0; r < world; ++r) {
          if (r == rank) continue;
          const int ll = plan.R[r].g2l_lm[gl];
          if (ll >= 0) svp.push_back((double*)(mr->peer_base[r] + mr->lay[r].v) + 3 * (size_t)ll);
        }
        svrow[k + 1] = (int)svp.size();
        for (int q = R.partbase[k]; q < R.partbase[k + 1]; ++q) {
          for (int r = 0; r < world; ++r) {
            if (r == rank) continue;
            const int ll = plan.R[r].g2l_lm[gl];
            if (ll < 0) continue;
            const size_t cell = mr->lay[r].n_ucells + (size_t)3 * (plan.R[r].partbase[ll] + (q - R.partbase[k]));
            vcellp.push_back((uint4*)(mr->peer_base[r] + mr->lay[r].cells) + cell);
          }
          vrow[q + 1] = (int)vcellp.size();
        }
      }
      cudaStream_t s = in->stream;
      SSB_TRY(mr->d_upush_rowptr.ensure(urow.size()));
      SSB_TRY(mr->d_upush_cell.ensure(std::max<size_t>(ucellp.size(), 1)));
      SSB_TRY(mr->d_upush_x.ensure(std::max<size_t>(uxp.size(), 1)));
      SSB_TRY(mr->d_vpush_rowptr.ensure(vrow.size()));
      SSB_TRY(mr->d_vpush_cell.ensure(std::max<size_t>(vcellp.size(), 1)));
      SSB_CUDA_CHECK(cudaMemcpyAsync(mr->d_upush_rowptr.p, urow.data(), urow.size() * sizeof(int), cudaMemcpyHostToDevice, s));
      if (!ucellp.empty()) {
        SSB_CUDA_CHECK(cudaMemcpyAsync(mr->d_upush_cell.p, ucellp.data(), ucellp.size() * sizeof(uint4*), cudaMemcpyHostToDevice, s));
        SSB_CUDA_CHECK(cudaMemcpyAsync(mr->d_upush_x.p, uxp.data(), uxp.size() * sizeof(double*), cudaMemcpyHostToDevice, s));
      }
      SSB_TRY(mr->d_zpush_z.ensure(std::max<size_t>(uzp.size(), 1)));
      SSB_TRY(mr->d_svpush_v.ensure(std::max<size_t>(svp.size(), 1)));
      SSB_TRY(mr->d_svpush_rowptr.ensure(svrow.size()));
      if (!uzp.empty()) SSB_CUDA_CHECK(cudaMemcpyAsync(mr->d_zpush_z.p, uzp.data(), uzp.size() * sizeof(double*), cudaMemcpyHostToDevice, s));
      if (!svp.empty()) SSB_CUDA_CHECK(cudaMemcpyAsync(mr->d_svpush_v.p, svp.data(), svp.size() * sizeof(double*), cudaMemcpyHostToDevice, s));
      SSB_CUDA_CHECK(cudaMemcpyAsync(mr->d_svpush_rowptr.p, svrow.data(), svrow.size() * sizeof(int), cudaMemcpyHostToDevice, s));
      SSB_CUDA_CHECK(cudaMemcpyAsync(mr->d_vpush_rowptr.p, vrow.data(), vrow.size() * sizeof(int), cudaMemcpyHostToDevice, s));
      if (!vcellp.empty())
        SSB_CUDA_CHECK(cudaMemcpyAsync(mr->d_vpush_cell.p, vcellp.data(), vcellp.size() * sizeof(uint4*), cudaMemcpyHostToDevice, s));
      // chi2 tables and gather maps
      mr->n_pl_own = (int)pl_own.size();
      mr->n_pp_own = (int)pp_own.size();
      SSB_TRY(mr->d_pl_own.ensure(std::max<size_t>(pl_own.size(), 1)));
      SSB_TRY(mr->d_zd_own.ensure(std::max<size_t>(zd_own.size(), 1)));
      SSB_TRY(mr->d_pp_own.ensure(std::max<size_t>(pp_own.size(), 1)));
      if (!pl_own.empty()) {
        SSB_CUDA_CHECK(cudaMemcpyAsync(mr->d_pl_own.p, pl_own.data(), pl_own.size() * sizeof(PLEdge), cudaMemcpyHostToDevice, s));
        SSB_CUDA_CHECK(cudaMemcpyAsync(mr->d_zd_own.p, zd_own.data(), zd_own.size() * sizeof(double), cudaMemcpyHostToDevice, s));
      }
      if (!pp_own.empty())
        SSB_CUDA_CHECK(cudaMemcpyAsync(mr->d_pp_own.p, pp_own.data(), pp_own.size() * sizeof(PPEdge), cudaMemcpyHostToDevice, s));
      SSB_TRY(mr->d_own_l2g.ensure(std::max(mr->n_own, 1)));
      SSB_TRY(mr->d_ownlm_l2g.ensure(std::max(R.n_owned_lm, 1)));
      if (mr->n_own) SSB_CUDA_CHECK(cudaMemcpyAsync(mr->d_own_l2g.p, R.l2g_pose.data(), mr->n_own * sizeof(int), cudaMemcpyHostToDevice, s));
      if (R.n_owned_lm)
        SSB_CUDA_CHECK(cudaMemcpyAsync(mr->d_ownlm_l2g.p, R.l2g_lm.data(), R.n_owned_lm * sizeof(int), cudaMemcpyHostToDevice, s));
      SSB_CUDA_CHECK(cudaStreamSynchronize(s));
      mr->FP.upush_rowptr = mr->d_upush_rowptr.p;
      mr->FP.upush_cell = mr->d_upush_cell.p;
      mr->FP.upush_x = mr->d_upush_x.p;
      mr->FP.vpush_rowptr = mr->d_vpush_rowptr.p;
      mr->FP.vpush_cell = mr->d_vpush_cell.p;
      mr->SP.world = world;
      mr->SP.rank = rank;
      mr->SP.zpush_rowptr = mr->d_upush_rowptr.p;
      mr->SP.zpush_z = mr->d_zpush_z.p;
      mr->SP.zpush_x = mr->d_upush_x.p;
      mr->SP.n_own_lm = R.n_owned_lm;
      mr->SP.vpush_rowptr = mr->d_svpush_rowptr.p;
      mr->SP.vpush_v = mr->d_svpush_v.p;
      for (int r = 0; r < world; ++r) mr->SP.slots[r] = mr->P.slots[r];
      // rank-level coarse level: owner of every local keyframe, pointers to every rank's centroid / row-block slots
      {
        const int per = std::max(1, (plan.Np + world - 1) / world);
        std::vector<int> prank(NpL);
        for (int k = 0; k < NpL; ++k) prank[k] = std::min(world - 1, R.l2g_pose[k] / per);
        std::vector<double*> gc(world), gr(world);
        for (int r = 0; r < world; ++r) {
          gc[r] = (double*)(mr->peer_base[r] + mr->lay[r].gcent);
          gr[r] = (double*)(mr->peer_base[r] + mr->lay[r].grows);
        }
        SSB_TRY(mr->d_pose_rank.ensure(std::max(NpL, 1)));
        SSB_TRY(mr->d_Bg.ensure((size_t)36 * std::max(NpL, 1)));
        SSB_TRY(mr->d_Gg.ensure((size_t)18 * world * std::max(NlL, 1)));
        SSB_TRY(mr->d_gpart.ensure((size_t)G_ROW_BLOCKS * 36 * world));
        SSB_TRY(mr->d_Aginv.ensure((size_t)36 * world));
        SSB_TRY(mr->d_gcent_all.ensure(world));
        SSB_TRY(mr->d_grows_all.ensure(world));
        if (NpL) SSB_CUDA_CHECK(cudaMemcpyAsync(mr->d_pose_rank.p, prank.data(), NpL * sizeof(int), cudaMemcpyHostToDevice, s));
        SSB_CUDA_CHECK(cudaMemcpyAsync(mr->d_gcent_all.p, gc.data(), world * sizeof(double*), cudaMemcpyHostToDevice, s));
        SSB_CUDA_CHECK(cudaMemcpyAsync(mr->d_grows_all.p, gr.data(), world * sizeof(double*), cudaMemcpyHostToDevice, s));
        SSB_CUDA_CHECK(cudaMemsetAsync(mr->d_Aginv.p, 0, (size_t)36 * world * sizeof(float), s));
        SSB_CUDA_CHECK(cudaStreamSynchronize(s));
        GlobDev& GD = mr->GD;
        GD.world = world;
        GD.rank = rank;
        GD.pose_rank = mr->d_pose_rank.p;
        GD.gcent = (double*)(mr->arena + L.gcent);
        GD.Bg = mr->d_Bg.p;
        GD.Gg = mr->d_Gg.p;
        GD.part = mr->d_gpart.p;
        GD.grows = (double*)(mr->arena + L.grows);
        GD.Aginv = mr->d_Aginv.p;
        mr->glob = in->opts.preconditioner >= 1 && std::getenv("SSB_NO_GLOBAL_LEVEL") == nullptr;
        mr->FP.glob = mr->glob ? 1 : 0;
        mr->FP.Aginv = mr->d_Aginv.p;
      }
    }
    g->structure_dirty = false;
    g->host_est_dirty = true;
  }
  if (g->host_est_dirty) {
    const RankLocal& R = g->plan->R[rank];
    for (size_t k = 0; k < R.l2g_pose.size(); ++k) in->poses[k] = g->poses[R.l2g_pose[k]];
    for (size_t k = 0; k < R.l2g_lm.size(); ++k) std::memcpy(&in->lms[4 * k], &g->lms[4 * (size_t)R.l2g_lm[k]], 4 * sizeof(double));
    in->host_est_dirty = true;
    in->device_est_newer = false;
    g->host_est_dirty = false;
    g->device_est_newer = false;
  }
  const bool rebuilt = in->structure_dirty;
  SSB_TRY(prepare(in));
  // Meet on the host once the tables are built: (1) every rank must run the SAME PCG kernel — the on-chip data-flow
  // kernel only if every shard fits it, else the streaming kernel everywhere; (2) ranks sharing one device: cudaMalloc /
  // cudaFree synchronise the whole device, so nobody may start a kernel that waits for a peer while another rank is
  // still allocating.
  if (rebuilt) {
    PeerBlob mine{}, all[SSB_MAX_WORLD];
    mine.serial = in->fast_ok ? 1 : 0;
    SSB_TRY(exchange_blobs(g, mine, all));
    for (int r = 0; r < world; ++r)
      if (!all[r].serial) in->fast_ok = false;
  }
  return SSB_OK;
}

// after an LM run on the shard: every rank pushes the estimates it owns into everybody's full replica; read it back
static int gather_estimates_sharded(ssb_graph* g) {
  ssb_graph* in = g->shard;
  MrCtx* mr = in->mr;
  const int n = mr->n_own + mr->n_owned_lm;
  if (n) {
    k_gather_push<<<(n + 127) / 128, 128, 0, in->stream>>>(in->G, mr->PG, mr->d_own_l2g.p, mr->n_own, mr->d_ownlm_l2g.p, mr->n_owned_lm);
    in->launches++;
  }
  SSB_TRY(peer_exchange(in, false));
  const ArenaLayout& L = mr->lay[mr->rank];
  const size_t Np = g->poses.size(), Nl = g->lms.size() / 4;
  if (Np) SSB_CUDA_CHECK(cudaMemcpyAsync(g->poses.data(), mr->arena + L.pose_full, Np * sizeof(Pose), cudaMemcpyDeviceToHost, in->stream));
  if (Nl) SSB_CUDA_CHECK(cudaMemcpyAsync(g->lms.data(), mr->arena + L.lm_full, Nl * 4 * sizeof(double), cudaMemcpyDeviceToHost, in->stream));
  SSB_CUDA_CHECK(cudaMemcpyAsync(mr->h_err, mr->d_err.p, sizeof(int), cudaMemcpyDeviceToHost, in->stream));
  SSB_CUDA_CHECK(cudaStreamSynchronize(in->stream));
  if (*mr->h_err) {
    set_error("sharded graph: rank %d did not arrive at the final gather (timeout)", *mr->h_err - 1);
    return SSB_ERR_COMM;
  }
  in->device_est_newer = false;   // the shard's host copy is not used; the outer handle holds the estimates
  g->device_est_newer = false;
  g->host_est_dirty = false;
  return SSB_OK;
}

// Host-side deadline for everything queued on the handle's stream.  The persistent PCG kernel waits on cells written by
// its own CTAs; if one is ever lost the kernel spins forever and a plain cudaStreamSynchronize would hang the caller with
// it.  Polling the stream costs nothing on the device (the in-kernel counter costs 4 %, ssb_pcg_flow.cuh): after
// SSB_HOST_DEADLINE_S seconds the call returns an error — the process must then exit to release the GPU.
static int stream_wait(ssb_graph* g) {
  static const double limit_s = [] {
    const char* e = std::getenv("SSB_HOST_DEADLINE_S");
    return e ? std::max(1.0, std::atof(e)) : 120.0;
  }();
  const double t0 = wall_ms();
  unsigned spins = 0;
  for (;;) {
    const cudaError_t q = cudaStreamQuery(g->stream);
    if (q == cudaSuccess) return SSB_OK;
    if (q != cudaErrorNotReady) SSB_CUDA_CHECK(q);
    if ((++spins & 1023u) == 0 && wall_ms() - t0 > 1e3 * limit_s) {
      set_error("the device did not finish within %.0f s (a persistent kernel is waiting for data that never arrived); exit the process to release the GPU", limit_s);
      return SSB_ERR_CUDA;
    }
  }
}
static int read_scalars(ssb_graph* g) {
  if (g->mr) return read_scalars_sharded(g);
  SSB_CUDA_CHECK(cudaMemcpyAsync(g->h_scalars, g->d_scalars.p, 32 * sizeof(double), cudaMemcpyDeviceToHost, g->stream));
  SSB_CUDA_CHECK(cudaMemcpyAsync(g->h_iscalars, g->d_iscalars.p, 4 * sizeof(int), cudaMemcpyDeviceToHost, g->stream));
  SSB_TRY(stream_wait(g));
  return SSB_OK;
}

// The LM loop of SparseOptimizer::optimize + OptimizationAlgorithmLevenberg::solve on device state.
static int lm_loop(ssb_graph* g, int max_iterations, ssb_lm_stats* st) {
  DevGraph& G = g->G;
  cudaStream_t s = g->stream;
  g->history.clear();
  g->ev_used = 0;
  SSB_CUDA_CHECK(cudaEventRecord(g->ev0, s));
  static const bool dbg = std::getenv("SSB_SHARD_DEBUG") != nullptr;
  const int dbg_rank = g->mr ? g->mr->rank : 0;
  SSB_TRY(launch_chi2(g));
  SSB_TRY(read_scalars(g));
  double currentChi = g->h_scalars[0];
  st->chi2_initial = currentChi;
  if (dbg) std::fprintf(stderr, "[ssb lm r%d] chi2_0 = %.6f\n", dbg_rank, currentChi);
  double lambda = 0.0, ni = 2.0;
  bool ok = true;
  int it = 0;
  for (it = 0; it < max_iterations && ok; ++it) {
    SSB_TRY(launch_linearize(g));
    if (it == 0) {
      // computeLambdaInit: tau * max diag, tau = 1e-5
      SSB_CUDA_CHECK(cudaMemsetAsync(G.scalars + 2, 0, sizeof(double), s));
      k_maxdiag<<<(G.Np + G.Nl + 255) / 256, 256, 0, s>>>(G);
      g->launches++;
      SSB_TRY(read_scalars(g));
      lambda = 1e-5 * g->h_scalars[2];
      ni = 2.0;
    }
    double rho = 0.0;
    int qmax = 0, pcg_its = 0;
    const double chi_before = currentChi;
    do {
      if (dbg) std::fprintf(stderr, "[ssb lm r%d] it %d trial %d lambda %.3e: solve\n", dbg_rank, it, qmax, lambda);
      SSB_TRY(launch_solve(g, lambda, 1));  // push() + setLambda + solve + update
      SSB_TRY(launch_chi2(g));              // computeActiveErrors + activeRobustChi2
      SSB_TRY(read_scalars(g));
      if (dbg)
        std::fprintf(stderr, "[ssb lm r%d] it %d trial %d: pcg %d its status %d chi2 %.6f scale %.6e\n", dbg_rank, it, qmax, g->h_iscalars[0],
                     g->h_iscalars[1], g->h_scalars[0], g->h_scalars[1]);
      st->total_trials++;
      pcg_its += g->h_iscalars[0];
      double tempChi = g->h_scalars[0];
      const bool ok2 = g->h_iscalars[1] == 0;
      if (!ok2) tempChi = std::numeric_limits<double>::max();
      rho = currentChi - tempChi;
      double scale = g->h_scalars[1] + 1e-3;
      rho /= scale;
      if (rho > 0 && std::isfinite(tempChi)) {
        double alpha = 1.0 - std::pow(2 * rho - 1, 3);
        alpha = std::min(alpha, 2.0 / 3.0);
        double scaleFactor = std::max(1.0 / 3.0, alpha);
        lambda *= scaleFactor;
        ni = 2;
        currentChi = tempChi;
        // discardTop(): nothing to do
      } else {
        lambda *= ni;
        ni *= 2;
        // pop(): restore the backup
        const int n = std::max(G.Np, G.Nl);
        k_copy_state<<<(n + 255) / 256, 256, 0, s>>>(G.pose, g->d_pose_bak.p, G.Np, G.lm, g->d_lm_bak.p, G.Nl);
        g->launches++;
      }
      qmax++;
    } while (rho < 0 && qmax < 10);
    st->total_pcg_iters += pcg_its;
    g->history.push_back(chi_before);
    g->history.push_back(currentChi);
    g->history.push_back(lambda);
    g->history.push_back(rho);
    g->history.push_back((double)qmax);
    g->history.push_back((double)pcg_its);
    if (qmax == 10 || rho == 0) {
      st->terminated = 1;
      ok = false;
    }
  }
  SSB_CUDA_CHECK(cudaEventRecord(g->ev1, s));
  SSB_CUDA_CHECK(cudaStreamSynchronize(s));
  float ms = 0;
  SSB_CUDA_CHECK(cudaEventElapsedTime(&ms, g->ev0, g->ev1));
  st->ms_device = ms;
  st->ms_pcg = 0.0;
  for (size_t k = 0; k + 1 < g->ev_used; k += 2) {
    float t = 0;
    SSB_CUDA_CHECK(cudaEventElapsedTime(&t, g->ev_pool[k], g->ev_pool[k + 1]));
    st->ms_pcg += t;
  }
  st->iterations = it;
  st->chi2_final = currentChi;
  st->lambda_final = lambda;
  g->device_est_newer = true;
  return SSB_OK;
}

// ---- landmark marginals, direct form (ssb_marg_direct.cuh): applies when every pose-pose edge joins consecutive keyframes ----
namespace {
struct MdCudaLauncher {
  cudaStream_t s;
  long long* launches;
  template <class K, class... A>
  void operator()(K kern, int gx, int gy, int block, A... args) {
    kern<<<dim3((unsigned)gx, (unsigned)gy), block, 0, s>>>(args...);
    ++*launches;
  }
  void zero(void* p, size_t n) { cudaMemsetAsync(p, 0, n, s); }
};
}  // namespace
// returns 1 (done), 0 (the factorisation met a non-positive pivot: the caller falls back to the iterative path, which reports
// the failure in its own terms), -100 (not applicable), or a negative error
static int marginals_direct(ssb_graph* g, const int* vids, int n, double* out9n) {
  if (const char* e = std::getenv("SSB_MARG_DIRECT"))
    if (std::atoi(e) == 0) return -100;
  if (g->mr || !g->ll.empty() || g->rep_count > 1) return -100;
  const DevGraph& G = g->G;
  if (G.Np < 1 || G.Nl < 1 || G.El < 1 || G.pose_kind != nullptr) return -100;
  for (const PPEdge& e : g->pp)
    if (e.i - e.j != 1 && e.j - e.i != 1) return -100;   // a loop closure: H_pp is not block tridiagonal
  const ssb_md::MdDims d = ssb_md::md_dims(G.Np, G.Nl);
  {
    double limit_gb = 16.0;
    if (const char* e = std::getenv("SSB_MARG_DIRECT_MAX_GB")) limit_gb = std::atof(e);
    const double need = 8.0 * ((double)d.K * d.ld + (double)d.ld * d.ld + 2.0 * ssb_md::TB * d.ld + 108.0 * d.Np);
    if (need > limit_gb * 1e9) return -100;
    if (need > 256e6) {   // only a large system is worth a cudaMemGetInfo (a driver round trip per call otherwise: once per frame)
      size_t free_b = 0, total_b = 0;
      if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) return -100;
      const double held = 8.0 * (double)(g->md_Y.cap + g->md_T.cap);
      if (need > 0.8 * ((double)free_b + held)) return -100;
    }
  }
  if (!g->have_system) SSB_TRY(launch_linearize(g));   // else: the system built by the last optimize (g2o semantics)
  SSB_TRY(g->md_Bsub.ensure((size_t)36 * d.Np));
  SSB_TRY(g->md_Ginv.ensure((size_t)36 * d.Np));
  SSB_TRY(g->md_Esub.ensure((size_t)36 * d.Np));
  SSB_TRY(g->md_Y.ensure((size_t)d.K * d.ld));
  SSB_TRY(g->md_T.ensure((size_t)d.ld * d.ld));
  SSB_TRY(g->md_Row.ensure((size_t)ssb_md::TB * d.ld));
  SSB_TRY(g->md_ColT.ensure((size_t)ssb_md::TB * d.ld));
  SSB_TRY(g->md_Pinv.ensure((size_t)ssb_md::TB * ssb_md::TB));
  SSB_TRY(g->md_out.ensure((size_t)9 * n));
  SSB_TRY(g->md_k0.ensure(d.nt));
  SSB_TRY(g->md_status.ensure(2));
  SSB_TRY(g->md_lidx.ensure(n));
  SSB_TRY(g->md_hstatus.ensure(2));
  std::vector<int> lidx(n);
  for (int k = 0; k < n; ++k) lidx[k] = g->V[vids[k]].idx;
  cudaStream_t s = g->stream;
  SSB_CUDA_CHECK(cudaMemcpyAsync(g->md_lidx.p, lidx.data(), (size_t)n * sizeof(int), cudaMemcpyHostToDevice, s));
  ssb_md::MdBuffers b{};
  b.pose_pp_rowptr = G.pose_pp_rowptr;
  b.pose_pp_idx = G.pose_pp_idx;
  b.pose_pp_other = G.pose_pp_other;
  b.Hoff = G.Hoff;
  b.Hpp = G.Hpp;
  b.Hll = G.Hll;
  b.HplL = G.HplL;
  b.edge_pose = reinterpret_cast<const int*>(G.pl);   // PLEdge::p is the first int of the 80-byte record
  b.edge_stride = (int)(sizeof(PLEdge) / sizeof(int));
  b.lm_rowptr = G.lm_rowptr;
  b.lidx = g->md_lidx.p;
  b.n_req = n;
  b.Bsub = g->md_Bsub.p;
  b.Ginv = g->md_Ginv.p;
  b.Esub = g->md_Esub.p;
  b.Y = g->md_Y.p;
  b.T = g->md_T.p;
  b.Row = g->md_Row.p;
  b.ColT = g->md_ColT.p;
  b.Pinv = g->md_Pinv.p;
  b.tile_k0 = g->md_k0.p;
  b.status = g->md_status.p;
  b.out9n = g->md_out.p;
  MdCudaLauncher L{s, &g->launches};
  ssb_md::md_run(L, d, b);
  SSB_CUDA_CHECK(cudaGetLastError());
  SSB_CUDA_CHECK(cudaMemcpyAsync(out9n, g->md_out.p, (size_t)9 * n * sizeof(double), cudaMemcpyDeviceToHost, s));
  SSB_CUDA_CHECK(cudaMemcpyAsync(g->md_hstatus.p, g->md_status.p, 2 * sizeof(int), cudaMemcpyDeviceToHost, s));
  SSB_TRY(stream_wait(g));
  if (std::getenv("SSB_MARG_DEBUG"))
    std::fprintf(stderr, "[ssb marginals] direct form: %d keyframes, %d landmarks, T %d x %d (%d tiles), Y %d rows, status %d %d\n", d.Np, d.Nl,
                 d.ld, d.ld, d.nt, d.K, g->md_hstatus.p[0], g->md_hstatus.p[1]);
  if (g->md_hstatus.p[0] != 0 || g->md_hstatus.p[1] != 0) return 0;
  for (int k = 0; k < 9 * n; ++k)
    if (!std::isfinite(out9n[k])) return 0;
  return 1;
}

extern "C" {

static bool is_sharded(const ssb_graph* g) { return g->comm_world > 1; }

int ssb_graph_prepare(ssb_graph* g) {
  if (!g) return SSB_ERR_INVALID;
  return is_sharded(g) ? prepare_sharded(g) : prepare(g);
}

// sharded LM: the loop runs on this rank's subgraph handle; every host decision is taken from the same exchanged
// records on every rank, so the ranks stay in lock-step without a leader
static int optimize_sharded(ssb_graph* g, int max_iterations, ssb_lm_stats* st, bool do_prepare) {
  const double t0 = wall_ms();
  if (do_prepare) SSB_TRY(prepare_sharded(g));
  if (!g->shard || g->structure_dirty) {
    set_error("optimize_resident: call ssb_graph_prepare first");
    return SSB_ERR_INVALID;
  }
  st->ms_prepare = wall_ms() - t0;
  ssb_graph* in = g->shard;
  SSB_CUDA_CHECK(cudaSetDevice(in->device));
  const long long l0 = in->launches;
  SSB_TRY(lm_loop(in, max_iterations, st));
  g->history = in->history;
  SSB_TRY(gather_estimates_sharded(g));
  st->ms_total = wall_ms() - t0;
  st->kernel_launches = in->launches - l0;
  return SSB_OK;
}

int ssb_graph_optimize(ssb_graph* g, int max_iterations, ssb_lm_stats* stats) {
  if (!g) return SSB_ERR_INVALID;
  ssb_lm_stats st;
  std::memset(&st, 0, sizeof(st));
  if (g->E.size() < 10) {  // graph_slam.cpp:184-186
    if (stats) *stats = st;
    return 0;
  }
  if (is_sharded(g)) {
    SSB_TRY(optimize_sharded(g, max_iterations, &st, true));
    if (stats) *stats = st;
    return 1;
  }
  const double t0 = wall_ms();
  const long long l0 = g->launches;
  SSB_TRY(prepare(g));
  st.ms_prepare = wall_ms() - t0;
  int r = lm_loop(g, max_iterations, &st);
  if (r != SSB_OK) return r;
  SSB_TRY(sync_estimates_to_host(g));
  st.ms_total = wall_ms() - t0;
  st.kernel_launches = g->launches - l0;
  if (stats) *stats = st;
  return 1;
}

// device-resident variant: requires ssb_graph_prepare; neither uploads nor downloads estimates
int ssb_graph_optimize_resident(ssb_graph* g, int max_iterations, ssb_lm_stats* stats) {
  if (!g) return SSB_ERR_INVALID;
  if (g->structure_dirty) {
    set_error("optimize_resident: call ssb_graph_prepare first");
    return SSB_ERR_INVALID;
  }
  ssb_lm_stats st;
  std::memset(&st, 0, sizeof(st));
  if (g->E.size() < 10) {
    if (stats) *stats = st;
    return 0;
  }
  SSB_CUDA_CHECK(cudaSetDevice(g->device));
  if (is_sharded(g)) {
    SSB_TRY(optimize_sharded(g, max_iterations, &st, false));
    if (stats) *stats = st;
    return 1;
  }
  const double t0 = wall_ms();
  const long long l0 = g->launches;
  int r = lm_loop(g, max_iterations, &st);
  if (r != SSB_OK) return r;
  st.ms_total = wall_ms() - t0;
  st.kernel_launches = g->launches - l0;
  if (stats) *stats = st;
  return 1;
}

// debug: globaltimer stamps of one k_pcg_flow iteration (only with -DSSB_FLOW_TRACE)
int ssb_graph_debug_trace(ssb_graph* g, unsigned long long* out, int n) {
  if (!g || !out || !g->d_trace.p) return SSB_ERR_INVALID;
  SSB_CUDA_CHECK(cudaMemcpy(out, g->d_trace.p, (size_t)n * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  return SSB_OK;
}

// debug: cycle counters of k_pcg accumulated since prepare (only with -DSSB_PCG_TIMERS)
int ssb_graph_debug_timers(ssb_graph* g, double out8[8]) {
  if (!g || !out8) return SSB_ERR_INVALID;
  SSB_TRY(read_scalars(g));
  for (int k = 0; k < 8; ++k) out8[k] = g->h_scalars[8 + k];
  return SSB_OK;
}

int ssb_graph_get_history(ssb_graph* g, double* out6n, int cap) {
  if (!g) return SSB_ERR_INVALID;
  int n = (int)(g->history.size() / 6);
  int m = std::min(n, cap);
  if (out6n && m > 0) std::memcpy(out6n, g->history.data(), (size_t)m * 6 * sizeof(double));
  return n;
}

int ssb_graph_chi2(ssb_graph* g, double* chi2_out) {
  if (!g || !chi2_out) return SSB_ERR_INVALID;
  if (is_sharded(g)) {   // collective: every rank must call
    SSB_TRY(prepare_sharded(g));
    SSB_TRY(launch_chi2(g->shard));
    SSB_TRY(read_scalars(g->shard));
    *chi2_out = g->shard->h_scalars[0];
    return SSB_OK;
  }
  SSB_TRY(prepare(g));
  SSB_TRY(launch_chi2(g));
  SSB_TRY(read_scalars(g));
  *chi2_out = g->h_scalars[0];
  return SSB_OK;
}

int ssb_graph_get_se3(ssb_graph* g, int vid, double T34[12]) {
  if (!check_vertex(g, vid, VK_SE3) || !T34) return SSB_ERR_INVALID;
  SSB_TRY(sync_estimates_to_host(g));
  pose_to_34(g->poses[g->V[vid].idx], T34);
  return SSB_OK;
}
int ssb_graph_get_point_xyz(ssb_graph* g, int vid, double xyz[3]) {
  if (!check_vertex(g, vid, VK_XYZ) || !xyz) return SSB_ERR_INVALID;
  SSB_TRY(sync_estimates_to_host(g));
  std::memcpy(xyz, &g->lms[4 * (size_t)g->V[vid].idx], 3 * sizeof(double));
  return SSB_OK;
}
int ssb_graph_set_se3(ssb_graph* g, int vid, const double T34[12]) {
  if (!check_vertex(g, vid, VK_SE3) || !T34) return SSB_ERR_INVALID;
  SSB_TRY(sync_estimates_to_host(g));
  pose_from_34(T34, g->poses[g->V[vid].idx]);
  g->host_est_dirty = true;
  return SSB_OK;
}
int ssb_graph_get_plane(ssb_graph* g, int vid, double coeffs[4]) {
  if (!check_vertex(g, vid, VK_PLANE) || !coeffs) return SSB_ERR_INVALID;
  SSB_TRY(sync_estimates_to_host(g));
  std::memcpy(coeffs, &g->lms[4 * (size_t)g->V[vid].idx], 4 * sizeof(double));
  return SSB_OK;
}
int ssb_graph_set_plane(ssb_graph* g, int vid, const double coeffs[4]) {
  if (!check_vertex(g, vid, VK_PLANE) || !coeffs) return SSB_ERR_INVALID;
  SSB_TRY(sync_estimates_to_host(g));
  double c[4] = {coeffs[0], coeffs[1], coeffs[2], coeffs[3]};
  plane_normalize(c);
  std::memcpy(&g->lms[4 * (size_t)g->V[vid].idx], c, 4 * sizeof(double));
  g->host_est_dirty = true;
  return SSB_OK;
}
int ssb_graph_set_point_xyz(ssb_graph* g, int vid, const double xyz[3]) {
  if (!check_vertex(g, vid, VK_XYZ) || !xyz) return SSB_ERR_INVALID;
  SSB_TRY(sync_estimates_to_host(g));
  std::memcpy(&g->lms[4 * (size_t)g->V[vid].idx], xyz, 3 * sizeof(double));
  g->host_est_dirty = true;
  return SSB_OK;
}
int ssb_graph_set_fixed(ssb_graph* g, int vid, int fixed) {
  if (!g || vid < 0 || vid >= (int)g->V.size()) return SSB_ERR_INVALID;
  g->V[vid].fixed = fixed != 0;
  g->structure_dirty = true;
  g->host_rev++;
  return SSB_OK;
}
int ssb_graph_hessian_index(ssb_graph* g, int vid) {
  if (!g || vid < 0 || vid >= (int)g->V.size()) return SSB_ERR_INVALID;
  if (!g->structure_dirty) return g->V[vid].hidx;   // assigned by prepare (buildIndexMapping): O(1) after an optimize
  int h = 0;
  for (int k = 0; k < vid; ++k)
    if (!g->V[k].fixed) ++h;
  return g->V[vid].fixed ? -1 : h;
}
int ssb_graph_get_all(ssb_graph* g, double* se3_out, double* xyz_out) {
  if (!g) return SSB_ERR_INVALID;
  SSB_TRY(sync_estimates_to_host(g));
  if (se3_out)
    for (size_t i = 0; i < g->poses.size(); ++i) pose_to_34(g->poses[i], se3_out + 12 * i);
  if (xyz_out)
    for (size_t l = 0; l < g->lms.size() / 4; ++l) std::memcpy(xyz_out + 3 * l, &g->lms[4 * l], 3 * sizeof(double));
  return SSB_OK;
}
// bulk setter (id order within each kind), counterpart of ssb_graph_get_all
int ssb_graph_set_all(ssb_graph* g, const double* se3_in, const double* xyz_in) {
  if (!g) return SSB_ERR_INVALID;
  SSB_TRY(sync_estimates_to_host(g));
  if (se3_in)
    for (size_t i = 0; i < g->poses.size(); ++i) pose_from_34(se3_in + 12 * i, g->poses[i]);
  if (xyz_in)
    for (size_t l = 0; l < g->lms.size() / 4; ++l) std::memcpy(&g->lms[4 * l], xyz_in + 3 * l, 3 * sizeof(double));
  g->host_est_dirty = true;
  return SSB_OK;
}
// force the next optimize() to rebuild and re-upload the edge tables (what the reference's
// initializeOptimization does on every call, graph_slam.cpp:199)
int ssb_graph_invalidate(ssb_graph* g) {
  if (!g) return SSB_ERR_INVALID;
  SSB_TRY(sync_estimates_to_host(g));
  g->structure_dirty = true;
  g->host_rev++;
  return SSB_OK;
}

int ssb_graph_snapshot(ssb_graph* g) {
  if (!g) return SSB_ERR_INVALID;
  if (is_sharded(g)) {
    SSB_TRY(prepare_sharded(g));
    g->snap_poses = g->poses;
    g->snap_lms = g->lms;
    g->have_snapshot = true;
    return ssb_graph_snapshot(g->shard);
  }
  SSB_TRY(prepare(g));
  SSB_TRY(g->d_pose_snap.ensure(g->G.Np));
  SSB_TRY(g->d_lm_snap.ensure((size_t)4 * g->G.Nl));
  const int n = std::max(g->G.Np, g->G.Nl);
  k_copy_state<<<(n + 255) / 256, 256, 0, g->stream>>>(g->d_pose_snap.p, g->G.pose, g->G.Np, g->d_lm_snap.p, g->G.lm, g->G.Nl);
  g->launches++;
  SSB_CUDA_CHECK(cudaStreamSynchronize(g->stream));
  g->have_snapshot = true;
  return SSB_OK;
}
int ssb_graph_restore(ssb_graph* g) {
  if (!g || !g->have_snapshot || g->structure_dirty) {
    set_error("restore: no snapshot for the current structure");
    return SSB_ERR_INVALID;
  }
  SSB_CUDA_CHECK(cudaSetDevice(g->device));
  if (is_sharded(g)) {
    g->poses = g->snap_poses;
    g->lms = g->snap_lms;
    SSB_TRY(ssb_graph_restore(g->shard));
    g->shard->device_est_newer = false;
    return SSB_OK;
  }
  const int n = std::max(g->G.Np, g->G.Nl);
  k_copy_state<<<(n + 255) / 256, 256, 0, g->stream>>>(g->G.pose, g->d_pose_snap.p, g->G.Np, g->G.lm, g->d_lm_snap.p, g->G.Nl);
  g->launches++;
  SSB_CUDA_CHECK(cudaGetLastError());
  g->device_est_newer = true;
  g->host_est_dirty = false;
  g->have_system = false;
  return SSB_OK;
}

int ssb_graph_edge_linearize(ssb_graph* g, int eid, double* err, double* Ji, double* Jj) {
  if (!g || eid < 0 || eid >= (int)g->E.size() || !err || !Ji || !Jj || is_sharded(g) || !g->ll.empty()) return SSB_ERR_INVALID;
  SSB_TRY(prepare(g));
  HostEdgeRef r = g->E[eid];
  if (r.kind == EK_LL) return SSB_ERR_INVALID;
  int idx = r.kind == EK_PP ? r.idx : g->plL_of_edge[r.idx];
  k_edge_linearize<<<1, 32, 0, g->stream>>>(g->G, r.kind == EK_PP ? 0 : 1, idx, g->d_tmp.p);
  g->launches++;
  double out[78];
  SSB_CUDA_CHECK(cudaMemcpyAsync(out, g->d_tmp.p, sizeof(out), cudaMemcpyDeviceToHost, g->stream));
  SSB_CUDA_CHECK(cudaStreamSynchronize(g->stream));
  if (r.kind == EK_PP) {
    std::memcpy(err, out, 6 * sizeof(double));
    std::memcpy(Ji, out + 6, 36 * sizeof(double));
    std::memcpy(Jj, out + 42, 36 * sizeof(double));
  } else {
    std::memcpy(err, out, 3 * sizeof(double));
    std::memcpy(Ji, out + 6, 18 * sizeof(double));
    std::memcpy(Jj, out + 42, 9 * sizeof(double));
  }
  return SSB_OK;
}

int ssb_graph_solve_once(ssb_graph* g, double lambda, double* x, int x_len) {
  if (!g || !x) return SSB_ERR_INVALID;
  if (is_sharded(g)) {
    set_error("solve_once: test hook of the unsharded back-end");
    return SSB_ERR_INVALID;
  }
  SSB_TRY(prepare(g));
  SSB_TRY(launch_linearize(g));
  SSB_TRY(launch_solve(g, lambda, 0));
  // dl without applying: reuse the back-substitution on a scratch copy of the state
  {
    const int n = std::max(g->G.Np, g->G.Nl);
    SSB_TRY(g->d_pose_snap.ensure(g->G.Np));
    SSB_TRY(g->d_lm_snap.ensure((size_t)4 * g->G.Nl));
    g->have_snapshot = false;
    k_copy_state<<<(n + 255) / 256, 256, 0, g->stream>>>(g->d_pose_snap.p, g->G.pose, g->G.Np, g->d_lm_snap.p, g->G.lm, g->G.Nl);
    k_backsub_update<<<(g->G.Np + 32 * g->G.Nl + 127) / 128, 128, 0, g->stream>>>(g->G, lambda, g->d_pose_bak.p, g->d_lm_bak.p);
    k_copy_state<<<(n + 255) / 256, 256, 0, g->stream>>>(g->G.pose, g->d_pose_snap.p, g->G.Np, g->G.lm, g->d_lm_snap.p, g->G.Nl);
    g->launches += 3;
    g->linpoint_in_bak = true;
  }
  std::vector<double> dp((size_t)6 * g->G.Np + 1), dl((size_t)3 * g->G.Nl + 1);
  SSB_CUDA_CHECK(cudaMemcpyAsync(dp.data(), g->G.x, (size_t)6 * g->G.Np * sizeof(double), cudaMemcpyDeviceToHost, g->stream));
  if (g->G.Nl)
    SSB_CUDA_CHECK(cudaMemcpyAsync(dl.data(), g->G.dl, (size_t)3 * g->G.Nl * sizeof(double), cudaMemcpyDeviceToHost, g->stream));
  SSB_TRY(read_scalars(g));
  int need = 0;
  for (auto& v : g->V)
    if (!v.fixed) need += v.kind == VK_SE3 ? 6 : 3;
  if (x_len < need) {
    set_error("solve_once: x_len %d < %d", x_len, need);
    return SSB_ERR_INVALID;
  }
  int o = 0;
  std::vector<int> prom_idx(g->lms.size() / 4 + 1, -1);
  if (!g->ll.empty())
    for (size_t k = 0; k < g->prom_lm.size(); ++k) prom_idx[g->prom_lm[k]] = (int)k;
  for (auto& v : g->V) {
    if (v.fixed) continue;
    if (v.kind == VK_SE3) {
      std::memcpy(x + o, &dp[6 * (size_t)v.idx], 6 * sizeof(double));
      o += 6;
    } else if (prom_idx[v.idx] >= 0) {   // a promoted landmark is solved for among the keyframes
      std::memcpy(x + o, &dp[6 * (g->poses.size() + (size_t)prom_idx[v.idx])], 3 * sizeof(double));
      o += 3;
    } else {
      std::memcpy(x + o, &dl[3 * (size_t)v.idx], 3 * sizeof(double));
      o += 3;
    }
  }
  if (g->h_iscalars[1] != 0) {
    set_error("PCG breakdown (status %d)", g->h_iscalars[1]);
    return SSB_ERR_NUMERIC;
  }
  return g->h_iscalars[0];
}

// ---- K5 with several right-hand sides per launch ------------------------------------------------------------------
// One damped solve is latency-bound and costs about the same for 1 000 as for 10 000 keyframes (the data-flow kernel
// holds up to 80 keyframes per CTA, a small graph fills a fraction of that).  The 3 n columns of the landmark marginals
// share ONE matrix, so a graph of Np keyframes is laid out k times side by side in a shadow handle (copy j: vertices / edges
// shifted, keyframes padded with fixed, edge-less ones up to a CTA boundary so that copy j owns the CTAs [j R, (j + 1) R)) and
// k_pcg_flow<148, false, true> runs one conjugate-gradient recurrence per copy: every column converges as if solved alone,
// one launch delivers k columns.  (Sharing the CG scalars between the copies does not work: one polynomial then has to
// serve every right-hand side and the iteration count grows with k — profiles/r02_marg_replicas/.)
// The shadow is rebuilt when the structure changed and takes its estimates (the linearisation point of the system the
// last optimize left behind, g2o's computeMarginals semantics) device-to-device.
struct MargRepPlan {
  int k, R, C, stride;   // copies, CTAs per copy, keyframes per CTA, keyframes per copy incl. padding
};
static MargRepPlan marg_replica_plan_uncached(const ssb_graph* g, int env);
static MargRepPlan marg_replica_plan(ssb_graph* g) {
  const char* ev = std::getenv("SSB_MARG_REPLICAS");   // 0 / 1: one column per launch (A/B measurements and tests)
  const int env = ev ? std::atoi(ev) : -1;
  if (g->marg_plan_serial != g->structure_serial || g->marg_plan_env != env) {
    const MargRepPlan P = marg_replica_plan_uncached(g, env);
    g->marg_plan[0] = P.k;
    g->marg_plan[1] = P.R;
    g->marg_plan[2] = P.C;
    g->marg_plan[3] = P.stride;
    g->marg_plan_serial = g->structure_serial;
    g->marg_plan_env = env;
  }
  return MargRepPlan{g->marg_plan[0], g->marg_plan[1], g->marg_plan[2], g->marg_plan[3]};
}
static MargRepPlan marg_replica_plan_uncached(const ssb_graph* g, int env) {
  MargRepPlan P{1, 0, 0, 0};
  const int Np = (int)g->poses.size(), Nl = (int)(g->lms.size() / 4);
  if (env == 0 || env == 1 || Np < 1 || g->pcg_grid != 148 || !g->use_flow || !g->allow_fast || g->mr) return P;
  int kmax = std::min({SSB_MARG_MAX_REP, PCGF_MAXREP, 148 / ((Np + PCGW_POSES - 1) / PCGW_POSES)});
  if (env > 1) kmax = std::min(kmax, env);
  if (kmax < 2) return P;
  // what a CTA of the on-chip kernel can hold (the limits prepare() checks), evaluated for k copies before building them:
  // pose-landmark / pose-pose incidences of its keyframe range, and — the landmark parts of ALL copies are dealt
  // round-robin over the 148 CTAs — the edges beyond the first 32 of its landmark parts
  std::vector<int> cpl(Np + 1, 0), cpp(Np + 1, 0), deg(std::max(Nl, 1), 0), ovp;
  for (const PLEdge& e : g->pl) {
    cpl[e.p + 1]++;
    deg[e.l]++;
  }
  for (const PPEdge& e : g->pp) {
    cpp[e.i + 1]++;
    cpp[e.j + 1]++;
  }
  for (int i = 0; i < Np; ++i) {
    cpl[i + 1] += cpl[i];
    cpp[i + 1] += cpp[i];
  }
  for (int l = 0; l < Nl; ++l)
    for (int e0 = 0; e0 < deg[l]; e0 += 32) ovp.push_back(0);   // the shadow cuts parts of <= 32 edges: no overflow batch
  std::vector<std::pair<int, int>> byp;   // (keyframe, landmark), keyframe-major: distinct landmarks of a keyframe range
  byp.reserve(g->pl.size());
  for (const PLEdge& e : g->pl) byp.push_back({e.p, e.l});
  std::sort(byp.begin(), byp.end());
  std::vector<int> stamp(std::max(Nl, 1), -1);
  for (int k = kmax; k >= 2; --k) {
    const int R = 148 / k;
    const int C = std::max(5, (((Np + R - 1) / R + 4) / 5) * 5);
    if (C > PCGW_POSES || (long long)k * (long long)ovp.size() > 148LL * (PCGF_THREADS / 32)) continue;
    bool ok = true;
    for (int b = 0; b < R && ok; ++b) {
      const int q0 = std::min(Np, b * C), q1 = std::min(Np, q0 + C);
      int distinct = 0;
      for (int t = cpl[q0]; t < cpl[q1]; ++t)
        if (stamp[byp[t].second] != k * 256 + b) {
          stamp[byp[t].second] = k * 256 + b;
          ++distinct;
        }
      // (distinct pose-pose edges / external neighbours: bounded through the incidences; prepare() has the last word)
      if (cpl[q1] - cpl[q0] > PCGF_MAXPL || distinct > PCGW_MAXU || cpp[q1] - cpp[q0] > PCGF_MAXPP || (cpp[q1] - cpp[q0]) / 2 + 2 > PCGW_MAXPPE)
        ok = false;
    }
    if (ok) {
      std::vector<int> ov(148, 0);
      size_t q = 0;
      for (int j = 0; j < k; ++j)
        for (int o : ovp) ov[q++ % 148] += o;
      for (int b = 0; b < 148; ++b)
        if (ov[b] > PCGF_MAXOV) ok = false;
    }
    if (!ok) continue;
    P.k = k;
    P.R = R;
    P.C = C;
    P.stride = R * C;
    return P;
  }
  return P;
}
static int build_marg_replica(ssb_graph* g, const MargRepPlan& P) {
  const int k = P.k;
  if (g->marg_rep && g->marg_rep_k == k && g->marg_rep_serial == g->structure_serial) return SSB_OK;
  if (!g->marg_rep) {
    ssb_graph_opts o = g->opts;
    o.device = g->device;
    g->marg_rep = ssb_graph_create(&o);
    if (!g->marg_rep) return SSB_ERR_CUDA;
  }
  ssb_graph* r = g->marg_rep;
  const int Np = (int)g->poses.size(), Nl = (int)(g->lms.size() / 4), nV = (int)g->V.size();
  const int nPP = (int)g->pp.size(), nPL = (int)g->pl.size();
  const int pad = P.stride - Np, nV1 = nV + pad;
  r->V.clear();
  r->E.clear();
  r->poses.clear();
  r->lms.clear();
  r->lm_kind.clear();
  r->pl_zd.clear();
  r->pose_vid.clear();
  r->lm_vid.clear();
  r->pp.clear();
  r->pl.clear();
  r->V.reserve((size_t)k * nV1);
  r->E.reserve((size_t)k * g->E.size());
  r->pp.reserve((size_t)k * nPP);
  r->pl.reserve((size_t)k * nPL);
  Pose ident;
  std::memset(&ident, 0, sizeof(ident));
  ident.q[3] = 1.0;
  for (int j = 0; j < k; ++j) {
    for (const HostVertex& v0 : g->V) {
      HostVertex v = v0;
      v.idx += j * (v.kind == VK_SE3 ? P.stride : Nl);
      v.hidx = -1;
      r->V.push_back(v);
    }
    for (int v : g->pose_vid) r->pose_vid.push_back(v + j * nV1);
    for (int q = 0; q < pad; ++q) {   // fixed keyframes without edges: identity diagonal block, zero right-hand side
      HostVertex v;
      v.kind = VK_SE3;
      v.idx = j * P.stride + Np + q;
      v.fixed = true;
      v.hidx = -1;
      r->pose_vid.push_back((int)r->V.size());
      r->V.push_back(v);
    }
    for (const HostEdgeRef& e0 : g->E) {
      HostEdgeRef e = e0;
      e.idx += j * (e.kind == EK_PP ? nPP : nPL);
      r->E.push_back(e);
    }
    r->poses.insert(r->poses.end(), g->poses.begin(), g->poses.end());
    r->poses.insert(r->poses.end(), (size_t)pad, ident);
    r->lms.insert(r->lms.end(), g->lms.begin(), g->lms.end());
    r->lm_kind.insert(r->lm_kind.end(), g->lm_kind.begin(), g->lm_kind.end());
    r->pl_zd.insert(r->pl_zd.end(), g->pl_zd.begin(), g->pl_zd.end());
    for (int v : g->lm_vid) r->lm_vid.push_back(v + j * nV1);
    for (const PPEdge& e0 : g->pp) {
      PPEdge e = e0;
      e.i += j * P.stride;
      e.j += j * P.stride;
      r->pp.push_back(e);
    }
    for (const PLEdge& e0 : g->pl) {
      PLEdge e = e0;
      e.p += j * P.stride;
      e.l += j * Nl;
      r->pl.push_back(e);
    }
  }
  r->n_plane_vertices = k * g->n_plane_vertices;
  r->rep_ctas = P.R;
  r->rep_count = k;
  r->force_C = P.C;
  r->part_edges = 32;
  r->structure_dirty = true;
  r->host_est_dirty = true;
  r->device_est_newer = false;
  g->marg_rep_k = k;
  g->marg_rep_serial = g->structure_serial;
  return SSB_OK;
}
// returns 1 / 0 like ssb_graph_landmark_marginals, or -100 when the replicated graph does not fit the on-chip kernel
// (the caller then solves one column per launch)
static int marginals_replicated(ssb_graph* g, const int* vids, int n, double* out9n, const MargRepPlan& P) {
  const int k = P.k;
  SSB_TRY(build_marg_replica(g, P));
  ssb_graph* r = g->marg_rep;
  SSB_TRY(prepare(r));
  if (!r->fast_ok || !r->use_flow) {   // a limit the plan did not foresee: one column per launch for this structure
    g->marg_plan[0] = 1;
    return -100;
  }
  const int Np = g->G.Np, Nl = g->G.Nl;
  cudaStream_t s = r->stream;
  SSB_CUDA_CHECK(cudaStreamSynchronize(g->stream));
  // the estimates the resident system of g was linearised at (see linpoint_in_bak), into every copy
  const bool from_bak = g->have_system && g->linpoint_in_bak;
  const Pose* src_pose = from_bak ? g->d_pose_bak.p : g->G.pose;
  const double* src_lm = from_bak ? g->d_lm_bak.p : g->G.lm;
  for (int j = 0; j < k; ++j) {
    SSB_CUDA_CHECK(cudaMemcpyAsync(r->G.pose + (size_t)j * P.stride, src_pose, (size_t)Np * sizeof(Pose), cudaMemcpyDeviceToDevice, s));
    if (Nl) SSB_CUDA_CHECK(cudaMemcpyAsync(r->G.lm + (size_t)4 * j * Nl, src_lm, (size_t)4 * Nl * sizeof(double), cudaMemcpyDeviceToDevice, s));
  }
  r->have_system = false;
  const long long l0 = r->launches;
  SSB_TRY(launch_linearize(r));
  SSB_TRY(launch_prep(r, 0.0, std::getenv("SSB_MARG_INKERNEL") == nullptr));
  SSB_TRY(r->d_tmp.ensure((size_t)std::max(128, 9 * n + 8)));
  double* status = r->d_tmp.p + 9 * (size_t)n;
  SSB_CUDA_CHECK(cudaMemsetAsync(status, 0, sizeof(double), s));
  const size_t ev_keep = r->ev_used;
  const int ncol = 3 * n;
  const bool dbg = std::getenv("SSB_MARG_DEBUG") != nullptr;
  for (int c0 = 0; c0 < ncol; c0 += k) {
    MargCols mc;
    for (int j = 0; j < SSB_MARG_MAX_REP; ++j) {
      const int col = c0 + j;
      const bool live = j < k && col < ncol;
      mc.l[j] = live ? g->V[vids[col / 3]].idx : -1;
      mc.c[j] = live ? col % 3 : 0;
      mc.o[j] = live ? col / 3 : 0;
    }
    k_marg_rhs_rep<<<k, 256, 0, s>>>(r->G, mc, P.stride, Nl);
    r->ev_used = ev_keep;  // do not grow the timing-event pool
    SSB_TRY(launch_pcg(r, 0.0));
    k_marg_out_rep<<<k, 64, 0, s>>>(r->G, mc, Nl, r->d_tmp.p, status);
    r->launches += 2;
    if (dbg) {
      SSB_TRY(read_scalars(r));
      std::fprintf(stderr, "[ssb marginals] %d copies of %d keyframes (%d CTAs x %d keyframes each), columns %d..: %d PCG iterations, status %d\n", k, Np,
                   P.R, P.C, c0, r->h_iscalars[0], r->h_iscalars[1]);
    }
  }
  r->ev_used = ev_keep;
  double st = 0.0;
  SSB_CUDA_CHECK(cudaMemcpyAsync(out9n, r->d_tmp.p, (size_t)9 * n * sizeof(double), cudaMemcpyDeviceToHost, s));
  SSB_CUDA_CHECK(cudaMemcpyAsync(&st, status, sizeof(double), cudaMemcpyDeviceToHost, s));
  SSB_TRY(read_scalars(r));
  g->launches += r->launches - l0;
  if (st != 0.0 || r->h_iscalars[1] != 0) {
    set_error("landmark_marginals: PCG breakdown (status of the last launch %d: 1 = non-positive curvature, 2 = bad right-hand side)", r->h_iscalars[1]);
    return 0;
  }
  return 1;
}

// Sharded graphs: every rank holds the full host graph and, after an optimize, the final estimates of ALL vertices (the
// gather at the end of optimize_sharded), so the marginals are computed per rank on an UNSHARDED copy of the graph on its
// own GPU — no exchange, every rank gets identical bits.  The copy is linearised at the final estimates (g2o's
// computeMarginals uses the system of the last iteration: identical whenever the LM terminated on rejected trials, one
// accepted step behind otherwise).  Capacity: the graph must fit one GPU (it does up to the streaming kernel's limits).
static int marginals_sharded(ssb_graph* g, const int* vids, int n, double* out9n) {
  if (!g->marg_full) {
    ssb_graph_opts o = g->opts;
    o.device = g->device;
    o.reserved[2] = 0;
    g->marg_full = ssb_graph_create(&o);
    if (!g->marg_full) return SSB_ERR_CUDA;
  }
  ssb_graph* f = g->marg_full;
  if (g->marg_full_rev != g->host_rev) {
    f->V = g->V;
    f->E = g->E;
    f->lm_kind = g->lm_kind;
    f->pl_zd = g->pl_zd;
    f->n_plane_vertices = g->n_plane_vertices;
    f->pose_vid = g->pose_vid;
    f->lm_vid = g->lm_vid;
    f->pp = g->pp;
    f->pl = g->pl;
    f->ll = g->ll;
    f->structure_dirty = true;
    g->marg_full_rev = g->host_rev;
  }
  f->poses = g->poses;
  f->lms = g->lms;
  f->host_est_dirty = true;
  f->device_est_newer = false;
  return ssb_graph_landmark_marginals(f, vids, n, out9n);
}

int ssb_graph_landmark_marginals(ssb_graph* g, const int* vids, int n, double* out9n) {
  if (!g || (n > 0 && (!vids || !out9n)) || n < 0) return SSB_ERR_INVALID;
  for (int k = 0; k < n; ++k)
    if (!(check_vertex(g, vids[k], VK_XYZ) || check_vertex(g, vids[k], VK_PLANE)) || g->V[vids[k]].fixed) {
      set_error("landmark_marginals: vertex %d is not a free XYZ vertex", vids[k]);
      return SSB_ERR_INVALID;
    }
  if (g->E.size() < 10) return 0;  // nothing was ever optimised (graph_slam.cpp:184-186)
  if (!g->ll.empty()) {
    set_error("landmark_marginals: not available on a graph with landmark-landmark edges");
    return 0;
  }
  if (is_sharded(g)) return marginals_sharded(g, vids, n, out9n);
  SSB_TRY(prepare(g));
  if (n == 0) return 1;
  {
    const int rd = marginals_direct(g, vids, n, out9n);
    if (rd == 1) return 1;
    if (rd < 0 && rd != -100) return rd;
  }
  {
    const MargRepPlan P = marg_replica_plan(g);
    if (P.k >= 2) {
      const int rr = marginals_replicated(g, vids, n, out9n, P);
      if (rr != -100) return rr;
    }
  }
  if (!g->have_system) SSB_TRY(launch_linearize(g));  // else: the system built by the last optimize (g2o semantics)
  SSB_TRY(launch_prep(g, 0.0, std::getenv("SSB_MARG_INKERNEL") == nullptr));   // env: A/B switch for measurements
  SSB_TRY(g->d_tmp.ensure((size_t)std::max(128, 9 * n)));
  const size_t ev_keep = g->ev_used;
  for (int k = 0; k < n; ++k) {
    const int l = g->V[vids[k]].idx;
    for (int c = 0; c < 3; ++c) {
      k_marg_rhs<<<1, 256, 0, g->stream>>>(g->G, l, c);
      g->ev_used = ev_keep;  // do not grow the timing-event pool
      SSB_TRY(launch_pcg(g, 0.0));
      k_marg_out<<<1, 64, 0, g->stream>>>(g->G, l, c, g->d_tmp.p + 9 * (size_t)k);
      g->launches += 2;
    }
  }
  g->ev_used = ev_keep;
  SSB_CUDA_CHECK(cudaMemcpyAsync(out9n, g->d_tmp.p, (size_t)9 * n * sizeof(double), cudaMemcpyDeviceToHost, g->stream));
  SSB_TRY(read_scalars(g));
  if (g->h_iscalars[1] != 0) {
    set_error("landmark_marginals: PCG breakdown");
    return 0;
  }
  return 1;
}

// ---- g2o text format (graph_slam.cpp:236-239 -> OptimizableGraph::save) ---------------------
int ssb_graph_save_g2o(ssb_graph* g, const char* path) {
  if (!g || !path) return SSB_ERR_INVALID;
  SSB_TRY(sync_estimates_to_host(g));
  FILE* f = std::fopen(path, "w");
  if (!f) {
    set_error("cannot open %s for writing", path);
    return SSB_ERR_INVALID;
  }
  std::fprintf(f, "PARAMS_SE3OFFSET 0 0 0 0 0 0 0 1\n");
  for (size_t id = 0; id < g->V.size(); ++id) {
    const HostVertex& v = g->V[id];
    if (v.kind == VK_SE3) {
      const Pose& P = g->poses[v.idx];
      std::fprintf(f, "VERTEX_SE3:QUAT %zu %.17g %.17g %.17g %.17g %.17g %.17g %.17g\n", id, P.t[0], P.t[1], P.t[2], P.q[0],
                   P.q[1], P.q[2], P.q[3]);
    } else if (v.kind == VK_PLANE) {
      // g2o VertexPlane::write: 4 coefficients followed by the display colour
      const double* p = &g->lms[4 * (size_t)v.idx];
      std::fprintf(f, "VERTEX_PLANE %zu %.17g %.17g %.17g %.17g 0 0 0\n", id, p[0], p[1], p[2], p[3]);
    } else {
      const double* p = &g->lms[4 * (size_t)v.idx];
      std::fprintf(f, "VERTEX_TRACKXYZ %zu %.17g %.17g %.17g\n", id, p[0], p[1], p[2]);
    }
    if (v.fixed) std::fprintf(f, "FIX %zu\n", id);
  }
  for (auto& r : g->E) {
    if (r.kind == EK_PP) {
      const PPEdge& e = g->pp[r.idx];
      std::fprintf(f, "EDGE_SE3:QUAT %d %d %.17g %.17g %.17g %.17g %.17g %.17g %.17g", g->pose_vid[e.i], g->pose_vid[e.j],
                   e.zt[0], e.zt[1], e.zt[2], e.zq[0], e.zq[1], e.zq[2], e.zq[3]);
      for (int k = 0; k < 21; ++k) std::fprintf(f, " %.17g", e.info[k]);
      std::fprintf(f, "\n");
    } else if (r.kind == EK_PL && g->lm_kind[g->pl[r.idx].l]) {
      // EdgeSE3Plane::write (include/g2o/edge_se3_plane.hpp:39-46): 4 coefficients + upper triangle
      const PLEdge& e = g->pl[r.idx];
      std::fprintf(f, "EDGE_SE3_PLANE %d %d %.17g %.17g %.17g %.17g", g->pose_vid[e.p], g->lm_vid[e.l], e.z[0], e.z[1], e.z[2],
                   g->pl_zd[r.idx]);
      for (int k = 0; k < 6; ++k) std::fprintf(f, " %.17g", e.info[k]);
      std::fprintf(f, "\n");
    } else if (r.kind == EK_PL) {
      const PLEdge& e = g->pl[r.idx];
      std::fprintf(f, "EDGE_SE3_TRACKXYZ %d %d 0 %.17g %.17g %.17g", g->pose_vid[e.p], g->lm_vid[e.l], e.z[0], e.z[1], e.z[2]);
      for (int k = 0; k < 6; ++k) std::fprintf(f, " %.17g", e.info[k]);
      std::fprintf(f, "\n");
    } else {
      const LLEdge& e = g->ll[r.idx];
      std::fprintf(f, "EDGE_POINT_XYZ %d %d %.17g %.17g %.17g", g->lm_vid[e.a], g->lm_vid[e.b], e.z[0], e.z[1], e.z[2]);
      const int ut[6] = {0, 1, 2, 4, 5, 8};
      for (int k = 0; k < 6; ++k) std::fprintf(f, " %.17g", e.info[ut[k]]);
      std::fprintf(f, "\n");
    }
  }
  std::fclose(f);
  return SSB_OK;
}

int ssb_graph_load_g2o(ssb_graph* g, const char* path) {
  if (!g || !path) return SSB_ERR_INVALID;
  if (!g->V.empty()) {
    set_error("load_g2o: graph must be empty");
    return SSB_ERR_INVALID;
  }
  std::ifstream in(path);
  if (!in) {
    set_error("cannot open %s", path);
    return SSB_ERR_INVALID;
  }
  std::string line;
  std::vector<int> fixes;
  while (std::getline(in, line)) {
    std::istringstream ss(line);
    std::string tag;
    if (!(ss >> tag)) continue;
    if (tag == "VERTEX_SE3:QUAT") {
      int id;
      double t[3], q[4];
      ss >> id >> t[0] >> t[1] >> t[2] >> q[0] >> q[1] >> q[2] >> q[3];
      double R[9], T[12];
      double qq[4] = {q[0], q[1], q[2], q[3]};
      quat_normalize(qq);
      quat_to_R(qq, R);
      for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) T[4 * r + c] = R[3 * r + c];
        T[4 * r + 3] = t[r];
      }
      int got = ssb_graph_add_se3_node(g, T);
      if (got != id) {
        set_error("load_g2o: vertex ids must be consecutive (got %d, expected %d)", id, got);
        return SSB_ERR_INVALID;
      }
    } else if (tag == "VERTEX_TRACKXYZ") {
      int id;
      double p[3];
      ss >> id >> p[0] >> p[1] >> p[2];
      int got = ssb_graph_add_point_xyz_node(g, p);
      if (got != id) {
        set_error("load_g2o: vertex ids must be consecutive (got %d, expected %d)", id, got);
        return SSB_ERR_INVALID;
      }
    } else if (tag == "VERTEX_PLANE") {
      int id;
      double c[4];
      ss >> id >> c[0] >> c[1] >> c[2] >> c[3];
      int got = ssb_graph_add_plane_node(g, c);
      if (got != id) {
        set_error("load_g2o: vertex ids must be consecutive (got %d, expected %d)", id, got);
        return SSB_ERR_INVALID;
      }
    } else if (tag == "EDGE_SE3_PLANE") {
      int a, b;
      double c[4], u[6], info[9];
      ss >> a >> b >> c[0] >> c[1] >> c[2] >> c[3];
      for (int k = 0; k < 6; ++k) ss >> u[k];
      expand_sym3(u, info);
      if (ssb_graph_add_se3_plane_edge(g, a, b, c, info) < 0) return SSB_ERR_INVALID;
    } else if (tag == "FIX") {
      int id;
      while (ss >> id) fixes.push_back(id);
    } else if (tag == "EDGE_SE3:QUAT") {
      int a, b;
      double t[3], q[4], u[21];
      ss >> a >> b >> t[0] >> t[1] >> t[2] >> q[0] >> q[1] >> q[2] >> q[3];
      for (int k = 0; k < 21; ++k) ss >> u[k];
      double R[9], T[12], info[36];
      quat_normalize(q);
      quat_to_R(q, R);
      for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) T[4 * r + c] = R[3 * r + c];
        T[4 * r + 3] = t[r];
      }
      expand_sym6(u, info);
      if (ssb_graph_add_se3_edge(g, a, b, T, info) < 0) return SSB_ERR_INVALID;
    } else if (tag == "EDGE_SE3_TRACKXYZ") {
      int a, b, pid;
      double z[3], u[6], info[9];
      ss >> a >> b >> pid >> z[0] >> z[1] >> z[2];
      for (int k = 0; k < 6; ++k) ss >> u[k];
      expand_sym3(u, info);
      if (ssb_graph_add_se3_point_xyz_edge(g, a, b, z, info) < 0) return SSB_ERR_INVALID;
    } else if (tag == "EDGE_POINT_XYZ") {
      int a, b;
      double z[3], u[6], info[9];
      ss >> a >> b >> z[0] >> z[1] >> z[2];
      for (int k = 0; k < 6; ++k) ss >> u[k];
      expand_sym3(u, info);
      if (ssb_graph_add_point_xyz_point_xyz_edge(g, a, b, z, info) < 0) return SSB_ERR_INVALID;
    }
  }
  // g2o's reader applies FIX lines after loading; the first-vertex rule of add_se3_node stays
  for (size_t id = 0; id < g->V.size(); ++id) g->V[id].fixed = false;
  for (int id : fixes)
    if (id >= 0 && id < (int)g->V.size()) g->V[id].fixed = true;
  g->structure_dirty = true;
  g->host_rev++;
  return SSB_OK;
}

int ssb_comm_unique_id(unsigned char id_out[128]) {
  if (!id_out) return SSB_ERR_INVALID;
  NcclApi& N = nccl_api();
  if (!N.ok) {
    set_error("multi-GPU: libnccl.so.2 could not be loaded");
    return SSB_ERR_COMM;
  }
  ssb_ncclUniqueId id;
  SSB_NCCL_CHECK(N.GetUniqueId(&id));
  std::memcpy(id_out, id.internal, 128);
  return SSB_OK;
}
int ssb_graph_attach_comm(ssb_graph* g, int rank, int world, const unsigned char unique_id[128]) {
  if (!g || world < 1 || world > SSB_MAX_WORLD || rank < 0 || rank >= world) return SSB_ERR_INVALID;
  g->local_group = 0;
  g->structure_dirty = true;
  g->host_rev++;
  NcclApi& N = nccl_api();
  if (g->comm && N.ok) {
    N.CommDestroy(g->comm);
    g->comm = nullptr;
  }
  g->comm_rank = 0;
  g->comm_world = 1;
  if (world == 1) return SSB_OK;
  if (!unique_id) return SSB_ERR_INVALID;
  if (!N.ok) {
    set_error("multi-GPU: libnccl.so.2 could not be loaded");
    return SSB_ERR_COMM;
  }
  SSB_CUDA_CHECK(cudaSetDevice(g->device));
  ssb_ncclUniqueId id;
  std::memcpy(id.internal, unique_id, 128);
  SSB_NCCL_CHECK(N.CommInitRank(&g->comm, world, id, rank));
  g->comm_rank = rank;
  g->comm_world = world;
  return SSB_OK;
}
// ranks = host threads of THIS process (one handle each; same or different devices): the arena pointers are handed
// over through an in-process registry keyed by `group_key`.  Several shards may share one GPU ("virtual shards":
// cta_per_rank = 74 or 37 so that all shards' persistent kernels are resident together) — this is how the sharded
// protocol is exercised on a single-GPU box.
int ssb_graph_attach_local(ssb_graph* g, int rank, int world, const char* group_key, int cta_per_rank) {
  if (!g || world < 1 || world > SSB_MAX_WORLD || rank < 0 || rank >= world || !group_key) return SSB_ERR_INVALID;
  if (g->shard) {
    set_error("attach_local: the graph is already sharded");
    return SSB_ERR_INVALID;
  }
  g->comm_rank = world > 1 ? rank : 0;
  g->comm_world = world;
  g->local_group = world > 1;
  g->local_key = group_key;
  g->opts.reserved[2] = cta_per_rank > 0 ? cta_per_rank : 0;
  g->structure_dirty = true;
  g->host_rev++;
  return SSB_OK;
}
// own keyframe range [out[0], out[1]) of `rank` and, with a graph, its local sizes: out[2] = keyframes incl. ghosts,
// out[3] = landmarks touched (host-only helper)
int ssb_shard_ranges(int n_poses, int n_landmarks, int world, int rank, int out4[4]) {
  if (world < 1 || rank < 0 || rank >= world || !out4 || n_poses < 0 || n_landmarks < 0) return SSB_ERR_INVALID;
  const int per = std::max(1, (n_poses + world - 1) / world);
  out4[0] = std::min(n_poses, rank * per);
  out4[1] = rank == world - 1 ? n_poses : std::min(n_poses, (rank + 1) * per);
  out4[2] = out4[3] = 0;
  return SSB_OK;
}
// The sharding plan from bare index lists (host-only, no GPU, no handle): what ssb_graph_prepare derives on every rank.
// pl_pose / pl_lm: keyframe and landmark index of every pose-landmark edge; pp_i / pp_j: keyframes of every pose-pose
// edge.  For `rank`: out[0..1] own keyframe range, out[2] local keyframes, out[3] owned / out[4] touched landmarks,
// out[5] local pose-landmark edges, out[6] (keyframe, rank) pushes of u per PCG iteration, out[7] (landmark part, rank)
// pushes of v.  ghost_out (may be NULL, capacity n_poses): global indices of the ghost keyframes, returns their count in
// out[2] - (out[1] - out[0]).  push_to (may be NULL, world ints): number of own keyframes pushed to every rank.
int ssb_shard_plan(int n_poses, int n_landmarks, const int* pl_pose, const int* pl_lm, int n_pl, const int* pp_i, const int* pp_j,
                   int n_pp, int world, int rank, int out8[8], int* ghost_out, int* push_to) {
  if (world < 1 || world > SSB_MAX_WORLD || rank < 0 || rank >= world || !out8 || n_poses < 0 || n_landmarks < 0 || n_pl < 0 || n_pp < 0)
    return SSB_ERR_INVALID;
  std::vector<EdgeIdx> pl(n_pl), pp(n_pp);
  for (int k = 0; k < n_pl; ++k) {
    if (pl_pose[k] < 0 || pl_pose[k] >= n_poses || pl_lm[k] < 0 || pl_lm[k] >= n_landmarks) return SSB_ERR_INVALID;
    pl[k] = {pl_pose[k], pl_lm[k]};
  }
  for (int k = 0; k < n_pp; ++k) {
    if (pp_i[k] < 0 || pp_i[k] >= n_poses || pp_j[k] < 0 || pp_j[k] >= n_poses) return SSB_ERR_INVALID;
    pp[k] = {pp_i[k], pp_j[k]};
  }
  ShardPlan plan;
  build_plan_raw(n_poses, n_landmarks, pl, pp, world, 148, plan);
  const RankLocal& R = plan.R[rank];
  out8[0] = R.ps;
  out8[1] = R.pe;
  out8[2] = (int)R.l2g_pose.size();
  out8[3] = R.n_owned_lm;
  out8[4] = (int)R.l2g_lm.size();
  int el = 0;
  for (auto& e : pl)
    if (R.g2l_lm[e.b] >= 0) ++el;
  out8[5] = el;
  int upush = 0, vpush = 0;
  if (push_to)
    for (int r = 0; r < world; ++r) push_to[r] = 0;
  for (int gp = R.ps; gp < R.pe; ++gp)
    for (int r = 0; r < world; ++r)
      if (r != rank && plan.R[r].g2l_pose[gp] >= 0) {
        ++upush;
        if (push_to) push_to[r]++;
      }
  for (int k = 0; k < R.n_owned_lm; ++k)
    for (int r = 0; r < world; ++r)
      if (r != rank && plan.R[r].g2l_lm[R.l2g_lm[k]] >= 0) vpush += R.partbase[k + 1] - R.partbase[k];
  out8[6] = upush;
  out8[7] = vpush;
  if (ghost_out)
    for (size_t k = R.pe - R.ps; k < R.l2g_pose.size(); ++k) ghost_out[k - (R.pe - R.ps)] = R.l2g_pose[k];
  return SSB_OK;
}
// the sharding plan of the current graph as every rank computes it (host-only): for rank `rank`, out[0..1] = own
// keyframe range, out[2] = local keyframes (own + ghosts), out[3] = owned landmarks, out[4] = touched landmarks,
// out[5] = local pose-landmark edges
int ssb_graph_shard_info(ssb_graph* g, int world, int rank, int out6[6]) {
  if (!g || world < 1 || world > SSB_MAX_WORLD || rank < 0 || rank >= world || !out6) return SSB_ERR_INVALID;
  ShardPlan plan;
  build_plan(g, world, 148, plan);
  const RankLocal& R = plan.R[rank];
  out6[0] = R.ps;
  out6[1] = R.pe;
  out6[2] = (int)R.l2g_pose.size();
  out6[3] = R.n_owned_lm;
  out6[4] = (int)R.l2g_lm.size();
  int el = 0;
  for (auto& e : g->pl)
    if (R.g2l_lm[e.l] >= 0) ++el;
  out6[5] = el;
  return SSB_OK;
}

}  // extern "C"

// pamg_api.cu -- handle, device memory and the C ABI of libpamg_cuda.so (see include/pamg.h).
// Host orchestration mirrors the contained procedures of Semi_implicit_iterative
// (transport_tri_semi.F90:407-889) and its V-cycle loop (:319-379).  No CPU fallback: every compute
// entry fails with PAMG_ERR_CUDA when no device is present.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "pamg_internal.h"
#include "pamg_kernels.cuh"
#include "pamg_stream.cuh"
#include "pamg_unstr.cuh"

using namespace pamg;

// ------------------------------------------------------------------ NCCL, bound lazily with dlopen
// (libnccl.so.2; the few entry points used for the halo exchange and the norm all-reduce)
namespace {
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclUint8 = 1, ncclInt32 = 2, ncclFloat64 = 8, ncclSum = 0, ncclMax = 2, ncclMin = 3, ncclSuccess = 0 };
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(ncclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  bool load() {
    if (lib) return true;
    lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) return false;
    GetUniqueId = (decltype(GetUniqueId))dlsym(lib, "ncclGetUniqueId");
    CommInitRank = (decltype(CommInitRank))dlsym(lib, "ncclCommInitRank");
    CommDestroy = (decltype(CommDestroy))dlsym(lib, "ncclCommDestroy");
    GroupStart = (decltype(GroupStart))dlsym(lib, "ncclGroupStart");
    GroupEnd = (decltype(GroupEnd))dlsym(lib, "ncclGroupEnd");
    Send = (decltype(Send))dlsym(lib, "ncclSend");
    Recv = (decltype(Recv))dlsym(lib, "ncclRecv");
    AllReduce = (decltype(AllReduce))dlsym(lib, "ncclAllReduce");
    return GetUniqueId && CommInitRank && CommDestroy && GroupStart && GroupEnd && Send && Recv && AllReduce;
  }
};
NcclApi g_nccl;
}  // namespace

struct pamg_handle;
namespace { void p2p_close(pamg_handle* h); }

// ------------------------------------------------------------------ handle
struct LevelDev {
  int s = 0, S = 0;
  long long C = 0, nelem = 0, ndof = 0;
  double* T[2] = {nullptr, nullptr};  // T[cur] = TNONLIN (the iterate); T[cur^1] = TNEW unless aliased
  int cur = 0;
  bool tnew_alias = true;             // TNEW == TNONLIN logically (no separate copy materialised)
  double *told = nullptr, *rhs = nullptr, *res = nullptr;
  double* ovlb[2] = {nullptr, nullptr};       // halo strips, double-buffered: (nstrips + nsend) * 3S doubles each
  int ovl_cur = 0;                            // ovlb[ovl_cur] holds the strips of the current iterate when strips_valid
  bool strips_valid = false;
  double* ovl_old = nullptr;                  // told strips (update_overlaps as written only)
  double* pc = nullptr;               // [U][NPC]
  int2* items = nullptr; int nitems = 0;  // work list of the row-streaming kernel (levels with s >= STREAM_MIN_S)
  bool rhs_valid = false;             // level 1: RHS matches TOLD
  double* spare = nullptr;            // third field buffer of pamg_smooth_host (level 1, allocated on first use)
};

struct pamg_handle {
  pamg_params p;
  int device = 0, nsm = 148;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[16] = {};
  long long launches = 0;
  std::string err;
  HaloPlan plan;
  int U = 0;  // local parents
  double* xg = nullptr;
  int32_t *strip_of = nullptr, *dst_strip = nullptr, *rev = nullptr, *hmap = nullptr;
  std::vector<LevelDev> lev;
  double* partial = nullptr; int npartial = 0; int last_partials = 0;
  double* out3 = nullptr;         // device
  double* out3_host = nullptr;    // pinned
  double* scratch = nullptr; size_t scratch_bytes = 0;  // L2 flush
  double* stage = nullptr; size_t stage_bytes = 0;      // pinned staging for host-buffer entry points
  int kernel_mode = 4;  // 4 window kernel (default; 1-D TMA tile ring, all neighbours from shared memory), 1 pipelined 1-D TMA tiles,
                        // 2 row-streaming, 3 branch-free direct, 0 direct loads; PAMG_KERNEL=win|tma1d|stream|direct2|direct
  int* counters = nullptr;
  bool capturing = false;   // stream capture in progress: no synchronisation, no per-launch error polling
  bool use_graph = true;    // replay the V-cycle as a CUDA graph from the second cycle on (PAMG_GRAPH=0 disables)
  struct VcGraph { long long key; cudaGraphExec_t exec; long long launches; };
  std::vector<VcGraph> vc_graphs;   // a few cached V-cycle graphs (solver / sweep counts / buffer parity)
  bool graph_nccl = true;       // try to capture NCCL calls into the V-cycle graph (PAMG_GRAPH_NCCL=0 disables)
  bool split_boundary = false;  // PAMG_SPLIT=1: tile kernel + k_boundary_fix for the children on parent faces (measured: no gain for Jacobi)
  bool fused_halo = false;  // PAMG_FUSED_HALO=1: sweeps write the next sweep's strips themselves (measured slower: the extra work
                            // of the few children on parent faces delays the per-tile barrier; profiles/README.md)
  bool gs_tma = true;   // coloured GS pass through the TMA tile kernel (PAMG_GS=direct selects the direct kernel)
  bool win_producer = true;  // window kernel with a producer warp (k_element_win2); PAMG_WIN=barrier: k_element_win
  bool gs_fused = true; // both colours in one pass (k_gs_win); PAMG_GS=twopass keeps the two in-place passes
  // per-kernel timing (element kernels only)
  bool profiling = false;
  std::vector<cudaEvent_t> pev;  // pairs
  int pev_used = 0;
  // distributed
  ncclComm_t comm = nullptr; int nranks = 1, rank = 0;
  // coarse-level agglomeration: levels >= agg_level are solved on part 0 for the whole mesh
  pamg_handle* agg = nullptr;       // part 0 only: handle over ALL parents whose level 1 is my level agg_level
  int agg_level = 0;                // 0 = none
  int level_offset = 0;             // agg handle: its level l is level l + level_offset of the owner
  bool shared_stream = false;
  std::vector<int32_t> part_first;  // copy of the partition table
  int U_global = 0;
  // unstructured
  UnstrDev un;
  int un_use_dir = 0;
  // pamg_smooth_host: copies on their own streams so that the upload of call k+1 overlaps the download of call k
  cudaStream_t up_stream = nullptr, down_stream = nullptr;
  cudaEvent_t ev_up = nullptr, ev_comp = nullptr, ev_down[2] = {nullptr, nullptr};
  unsigned pipe_calls = 0;             // calls since the last pamg_sync
  // halo exchange by direct stores into peer memory (CUDA IPC over NVLink); PAMG_P2P=0 keeps ncclSend/ncclRecv
  bool p2p_enabled = true, p2p_ready = false, p2p_failed = false;
  bool p2p_fuse = true;                // cut-face values go to the peers from inside k_halo (PAMG_P2P_FUSE=0: separate kernel)
  unsigned long long* p2p_sync = nullptr;   // exchange number, block counter, error word (local)
  uint4* p2p_stage = nullptr;               // flagged receive staging, 2 parities x p2p_stage_words (IPC-exported)
  long long p2p_stage_words = 0;
  struct P2PPeer { int slot_at_peer = -1, strip_begin_at_peer = 0; long long recv_strips_at_peer = 0; uint4* stage = nullptr; };
  std::vector<P2PPeer> p2p_peers;      // same order as plan.peers
  std::vector<void*> p2p_opened;       // IPC mappings to close
  unsigned long long p2p_timeout_ns = 60000000000ull;   // PAMG_P2P_TIMEOUT_S: how long a halo kernel polls for a peer before it gives up
};

namespace {

int fail(pamg_handle* h, int code, const std::string& msg) {
  if (h) h->err = msg;
  return code;
}
#define CK(call)                                                                                 \
  do {                                                                                           \
    cudaError_t e_ = (call);                                                                     \
    if (e_ != cudaSuccess)                                                                       \
      return fail(h, PAMG_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));         \
  } while (0)

inline int grid_for(const pamg_handle* h, long long n) {
  long long need = (n + TPB - 1) / TPB;
  long long cap = (long long)h->nsm * 8;  // 8 resident CTAs of 256 threads per SM
  return (int)std::max(1ll, std::min(need, cap));
}

bool valid_level(const pamg_handle* h, int level) { return h && level >= 1 && level <= (int)h->lev.size(); }

double* tnew_ptr(LevelDev& L) { return L.tnew_alias ? L.T[L.cur] : L.T[L.cur ^ 1]; }

// make TNEW a real, separate copy of what it logically holds
int materialise_tnew(pamg_handle* h, LevelDev& L) {
  if (!L.tnew_alias) return PAMG_OK;
  CK(cudaMemcpyAsync(L.T[L.cur ^ 1], L.T[L.cur], L.ndof * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  L.tnew_alias = false;
  return PAMG_OK;
}

double* field_ptr(pamg_handle* h, int field, int level, bool for_write, int* rc) {
  *rc = PAMG_OK;
  LevelDev& L = h->lev[level - 1];
  switch (field) {
    case PAMG_TNEW:
      if (for_write) { *rc = materialise_tnew(h, L); L.strips_valid = false; return L.T[L.cur ^ 1]; }
      return tnew_ptr(L);
    case PAMG_TNONLIN:
      if (for_write) { *rc = materialise_tnew(h, L); L.strips_valid = false; }
      return L.T[L.cur];
    case PAMG_TOLD: if (for_write) L.rhs_valid = false; return L.told;
    case PAMG_RHS: if (for_write) L.rhs_valid = true; return L.rhs;
    case PAMG_RES: return L.res;
  }
  *rc = PAMG_ERR_ARG;
  return nullptr;
}

// ---- per-parent geometry in closed form (tri_det_nlx ShapFun.F90:1414-1454; det_snlx_all :1554-1590;
//      level scaling :1678-1683,1751-1780; get_d_center Msh2Tri.F90:358-383; add_diffusion_surf
//      matrices.F90:84-110).  One row of NPC coefficients per parent per level.
void parent_coefficients(const pamg_params& p, const double* Xall, const int32_t* neig, int g /*global parent*/,
                         int s, double* pc) {
  const double* X = Xall + (size_t)g * 6;
  const double x1 = X[0], y1 = X[1], x2 = X[2], y2 = X[3], x3 = X[4], y3 = X[5];
  const double A = x1 - x3, B = y1 - y3, Cc = x2 - x3, D = y2 - y3;
  const double detj = A * D - B * Cc;
  const double area = 0.5 * std::fabs(detj);
  const double g1[2] = {D / detj, -Cc / detj}, g2[2] = {-B / detj, A / detj};
  const double g3[2] = {-(g1[0] + g2[0]), -(g1[1] + g2[1])};
  const double* G[3] = {g1, g2, g3};
  const double two_s = std::ldexp(1.0, s), four_s = std::ldexp(1.0, 2 * s);
  pc[PC_CM] = area / (12.0 * four_s * p.dt);
  auto dot = [](const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1]; };
  pc[PC_K11] = p.k * area * dot(g1, g1); pc[PC_K12] = p.k * area * dot(g1, g2); pc[PC_K13] = p.k * area * dot(g1, g3);
  pc[PC_K22] = p.k * area * dot(g2, g2); pc[PC_K23] = p.k * area * dot(g2, g3); pc[PC_K33] = p.k * area * dot(g3, g3);
  const double uu[2] = {p.u_x, p.u_y};
  for (int i = 0; i < 3; ++i) pc[PC_ADV + i] = area * dot(G[i], uu) / (3.0 * two_s);
  const double cx = (x1 + x2 + x3) / 3.0, cy = (y1 + y2 + y3) / 3.0;
  // child faces: f1 = nodes (1,3) on side 1, f2 = (3,2) on side 3, f3 = (2,1) on side 2
  const int fa[3] = {0, 2, 1}, fb[3] = {2, 1, 0}, mface[3] = {0, 2, 1};
  const double V1[2] = {A, B}, V2[2] = {Cc, D};
  const double ci[3][2] = {{1.0 / 3, -2.0 / 3}, {-2.0 / 3, 1.0 / 3}, {1.0 / 3, 1.0 / 3}};  // centroid offsets of child 2's neighbours
  for (int f = 0; f < 3; ++f) {
    const double ax = X[2 * fa[f]], ay = X[2 * fa[f] + 1], bx = X[2 * fb[f]], by = X[2 * fb[f] + 1];
    const double ex = bx - ax, ey = by - ay, L = std::sqrt(ex * ex + ey * ey);
    double nx = ey / L, ny = -ex / L;
    const double mx = 0.5 * (ax + bx), my = 0.5 * (ay + by);
    if (nx * (mx - cx) + ny * (my - cy) < 0) { nx = -nx; ny = -ny; }
    const double lhalf = 0.5 * L;
    pc[PC_FL + f] = (uu[0] * nx + uu[1] * ny) * lhalf / (3.0 * two_s);
    const double dix = ci[f][0] * V1[0] + ci[f][1] * V2[0], diy = ci[f][0] * V1[1] + ci[f][1] * V2[1];
    const double dcI = std::sqrt(dix * dix + diy * diy);
    pc[PC_PENI + f] = p.k * lhalf / (3.0 * dcI);
    const int q = neig[(size_t)g * 3 + mface[f]];
    double dcX;
    if (q != 0) {
      const double* Y = Xall + (size_t)(q - 1) * 6;
      const double qx = (Y[0] + Y[2] + Y[4]) / 3.0, qy = (Y[1] + Y[3] + Y[5]) / 3.0;
      dcX = std::sqrt((cx - qx) * (cx - qx) + (cy - qy) * (cy - qy));
    } else {
      dcX = std::sqrt((cx - mx) * (cx - mx) + (cy - my) * (cy - my));
    }
    pc[PC_PENX + f] = p.k * lhalf / (3.0 * dcX);
  }
  // omega / D of children with all three faces inside the parent (get_diagonal :481-486): node 1 sits on
  // faces 1,3; node 2 on faces 2,3; node 3 on faces 1,2.  The penalty diagonal exists only with the face block.
  const double f2 = p.face_terms ? 2.0 : 0.0;
  pc[PC_W + 0] = p.omega / (4.0 * pc[PC_CM] + pc[PC_K11] + f2 * (pc[PC_PENI + 0] + pc[PC_PENI + 2]));
  pc[PC_W + 1] = p.omega / (4.0 * pc[PC_CM] + pc[PC_K22] + f2 * (pc[PC_PENI + 1] + pc[PC_PENI + 2]));
  pc[PC_W + 2] = p.omega / (4.0 * pc[PC_CM] + pc[PC_K33] + f2 * (pc[PC_PENI + 0] + pc[PC_PENI + 1]));
  pc[22] = 0.0; pc[23] = 0.0;
  // folded operator of interior children, one set per orientation (struct Folded in pamg_kernels.cuh)
  for (int o = 0; o < 2; ++o) {
    const double sg = o == 0 ? 1.0 : -1.0;
    double* F = pc + PC_FOLD + 16 * o;
    double A[3][3];
    const double K[3][3] = {{pc[PC_K11], pc[PC_K12], pc[PC_K13]}, {pc[PC_K12], pc[PC_K22], pc[PC_K23]},
                            {pc[PC_K13], pc[PC_K23], pc[PC_K33]}};
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j)
        A[i][j] = pc[PC_CM] * ((i == j ? 1.0 : 0.0) + 1.0) - sg * pc[PC_ADV + i] + K[i][j];
    const int fa2[3] = {0, 2, 1}, fb2[3] = {2, 1, 0};   // face nodes (a,b): f1 (1,3), f2 (3,2), f3 (2,1)
    double c[3];
    for (int f = 0; f < 3; ++f) {
      const double pen = p.face_terms ? pc[PC_PENI + f] : 0.0;
      const double fl = p.face_terms ? sg * pc[PC_FL + f] : 0.0;
      const bool in = fl < 0.0;
      const double own = pen + (in ? 0.0 : fl);
      c[f] = -pen + (in ? fl : 0.0);
      const int ia = fa2[f], ib = fb2[f];
      A[ia][ia] += 2.0 * own; A[ia][ib] += own; A[ib][ia] += own; A[ib][ib] += 2.0 * own;
    }
    F[0] = A[0][0]; F[1] = A[0][1]; F[2] = A[0][2]; F[3] = A[1][0]; F[4] = A[1][1]; F[5] = A[1][2];
    F[6] = A[2][0]; F[7] = A[2][1]; F[8] = A[2][2]; F[9] = c[0]; F[10] = c[1]; F[11] = c[2];
    F[12] = pc[PC_W + 0]; F[13] = pc[PC_W + 1]; F[14] = pc[PC_W + 2]; F[15] = 0.0;
  }
  // children with faces on the parent boundary: penalty change per face and omega / D per face mask
  for (int f = 0; f < 3; ++f) pc[PC_DPEN + f] = p.face_terms ? pc[PC_PENX + f] - pc[PC_PENI + f] : 0.0;
  pc[PC_DPEN + 3] = 0.0;
  for (int mask = 0; mask < 8; ++mask) {
    double pen[3];
    for (int f = 0; f < 3; ++f) pen[f] = p.face_terms ? ((mask >> f) & 1 ? pc[PC_PENX + f] : pc[PC_PENI + f]) : 0.0;
    pc[PC_WB + mask * 3 + 0] = p.omega / (4.0 * pc[PC_CM] + pc[PC_K11] + 2.0 * (pen[0] + pen[2]));
    pc[PC_WB + mask * 3 + 1] = p.omega / (4.0 * pc[PC_CM] + pc[PC_K22] + 2.0 * (pen[1] + pen[2]));
    pc[PC_WB + mask * 3 + 2] = p.omega / (4.0 * pc[PC_CM] + pc[PC_K33] + 2.0 * (pen[0] + pen[1]));
  }
  for (int i = PC_WB + 24; i < NPC; ++i) pc[i] = 0.0;
}

int launch_halo(pamg_handle* h, int level, int what = 0);

void p2p_close(pamg_handle* h) {
  for (void* p : h->p2p_opened) cudaIpcCloseMemHandle(p);
  h->p2p_opened.clear();
  if (h->p2p_sync) { cudaFree(h->p2p_sync); h->p2p_sync = nullptr; }
  if (h->p2p_stage) { cudaFree(h->p2p_stage); h->p2p_stage = nullptr; }
  h->p2p_ready = false; h->p2p_failed = false;
  for (auto& pp : h->p2p_peers) pp.stage = nullptr;
}

// collective over all ranks (called at the first exchange, never during stream capture): allocate the flagged
// staging buffer, exchange its CUDA IPC handle, map the peers' buffers, agree on the outcome
int p2p_setup(pamg_handle* h) {
  const int R = h->nranks;
  int ok = 1;
  if ((int)h->plan.peers.size() > P2P_MAXP) ok = 0;
  for (const auto& pp : h->p2p_peers) if (pp.slot_at_peer < 0) ok = 0;
  long long strips = 0;
  for (const auto& pr : h->plan.peers) strips = std::max(strips, (long long)pr.strip_begin + pr.nfaces);
  h->p2p_stage_words = strips * 3 * h->lev[0].S;
  const size_t stage_bytes = (size_t)std::max(1ll, 2 * h->p2p_stage_words) * sizeof(uint4);
  if (cudaMalloc(&h->p2p_sync, P2P_WORDS * sizeof(unsigned long long)) != cudaSuccess) { h->p2p_sync = nullptr; ok = 0; }
  else CK(cudaMemset(h->p2p_sync, 0, P2P_WORDS * sizeof(unsigned long long)));
  if (cudaMalloc(&h->p2p_stage, stage_bytes) != cudaSuccess) { h->p2p_stage = nullptr; ok = 0; }
  else CK(cudaMemset(h->p2p_stage, 0, stage_bytes));
  cudaIpcMemHandle_t mine;
  std::memset(&mine, 0, sizeof(mine));
  if (ok && cudaIpcGetMemHandle(&mine, h->p2p_stage) != cudaSuccess) { ok = 0; cudaGetLastError(); }
  // all-gather of the handles with grouped send / recv (R <= 8)
  const size_t hb = sizeof(cudaIpcMemHandle_t);
  std::vector<cudaIpcMemHandle_t> all(R);
  unsigned char *d_mine = nullptr, *d_all = nullptr;
  int* d_ok = nullptr;
  CK(cudaMalloc(&d_mine, hb)); CK(cudaMalloc(&d_all, hb * R)); CK(cudaMalloc(&d_ok, sizeof(int)));
  CK(cudaMemcpy(d_mine, &mine, hb, cudaMemcpyHostToDevice));
  g_nccl.GroupStart();
  for (int r = 0; r < R; ++r) {
    g_nccl.Send(d_mine, hb, ncclUint8, r, h->comm, h->stream);
    g_nccl.Recv(d_all + hb * r, hb, ncclUint8, r, h->comm, h->stream);
  }
  if (g_nccl.GroupEnd() != ncclSuccess) return fail(h, PAMG_ERR_CUDA, "handle exchange failed");
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaMemcpy(all.data(), d_all, hb * R, cudaMemcpyDeviceToHost));
  if (ok) {
    for (size_t i = 0; i < h->plan.peers.size(); ++i) {
      void* ptr = nullptr;
      if (cudaIpcOpenMemHandle(&ptr, all[h->plan.peers[i].part], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; cudaGetLastError(); break; }
      h->p2p_opened.push_back(ptr);
      h->p2p_peers[i].stage = (uint4*)ptr;
    }
  }
  // every rank must take the same path
  CK(cudaMemcpy(d_ok, &ok, sizeof(int), cudaMemcpyHostToDevice));
  if (g_nccl.AllReduce(d_ok, d_ok, 1, ncclInt32, ncclMin, h->comm, h->stream) != ncclSuccess) return fail(h, PAMG_ERR_CUDA, "ncclAllReduce failed");
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaMemcpy(&ok, d_ok, sizeof(int), cudaMemcpyDeviceToHost));
  cudaFree(d_mine); cudaFree(d_all); cudaFree(d_ok);
  if (ok) h->p2p_ready = true; else h->p2p_failed = true;
  return PAMG_OK;
}

int p2p_args(pamg_handle* h, LevelDev& L, double* ovl, P2PArgs& a) {
  const long long S3 = 3ll * L.S;
  const int base = h->plan.peers[0].send_begin;
  a.send = ovl + ((size_t)h->plan.nstrips + base) * S3;
  a.strips = ovl; a.stage = h->p2p_stage; a.stage_words = h->p2p_stage_words;
  a.sync = h->p2p_sync; a.npeers = (int)h->plan.peers.size(); a.timeout_ns = h->p2p_timeout_ns;
  long long so = 0, ro = 0;
  for (int i = 0; i < a.npeers; ++i) {
    const auto& pr = h->plan.peers[i];
    if ((long long)(pr.send_begin - base) * S3 != so) return fail(h, PAMG_ERR_STATE, "send slots are not contiguous per peer");
    a.remote[i] = h->p2p_peers[i].stage + (long long)h->p2p_peers[i].strip_begin_at_peer * S3;
    a.rstride[i] = h->p2p_peers[i].recv_strips_at_peer * 3 * h->lev[0].S;   // the peer's own p2p_stage_words
    a.soff[i] = so; a.roff[i] = ro; a.rbeg[i] = (long long)pr.strip_begin * S3;
    so += pr.nfaces * S3; ro += pr.nfaces * S3;
  }
  a.soff[a.npeers] = so; a.roff[a.npeers] = ro;
  a.send_base = base;
  return PAMG_OK;
}

int p2p_exchange(pamg_handle* h, LevelDev& L, double* ovl) {
  P2PArgs a;
  int rc = p2p_args(h, L, ovl, a);
  if (rc) return rc;
  const long long so = a.soff[a.npeers];
  static const int maxgrid = getenv("PAMG_P2P_GRID") ? std::max(1, atoi(getenv("PAMG_P2P_GRID"))) : 32;
  const int grid = (int)std::max(1ll, std::min((so + 2 * TPB - 1) / (2 * TPB), (long long)maxgrid));
  k_p2p_exchange<<<grid, TPB, 0, h->stream>>>(a);
  h->launches++;
  CK(cudaGetLastError());
  return PAMG_OK;
}

// exchange of the cut-face strips (one process per GPU): the send slots follow the local strips in the
// strip space; the receive range of a peer is a contiguous range of my own strips (pamg_plan.cpp)
int exchange_halo(pamg_handle* h, LevelDev& L, double* ovl) {
  if (h->plan.peers.empty()) return PAMG_OK;
  if (!h->comm) return fail(h, PAMG_ERR_STATE, "partitioned mesh but pamg_comm_init was not called");
  if (h->p2p_enabled && !h->p2p_ready && !h->p2p_failed && !h->capturing && h->level_offset == 0) {
    int rc = p2p_setup(h);
    if (rc) return rc;
  }
  if (h->p2p_ready) return p2p_exchange(h, L, ovl);
  const size_t S3 = (size_t)3 * L.S;
  g_nccl.GroupStart();
  for (const auto& pr : h->plan.peers) {
    const size_t n = (size_t)pr.nfaces * S3;
    g_nccl.Send(ovl + ((size_t)h->plan.nstrips + pr.send_begin) * S3, n, ncclFloat64, pr.part, h->comm, h->stream);
    g_nccl.Recv(ovl + (size_t)pr.strip_begin * S3, n, ncclFloat64, pr.part, h->comm, h->stream);
  }
  if (g_nccl.GroupEnd() != ncclSuccess) return fail(h, PAMG_ERR_CUDA, "ncclGroupEnd failed in halo exchange");
  return PAMG_OK;
}

// what: 0 = update_overlaps as written (every face, tnew and told strips); 1 = Dirichlet faces only (static data,
// written once per level into both strip buffers); 3 = every face between parents.  The sweeps themselves refresh
// the strips of the next sweep (strips_write in the kernels), so this kernel only runs when the field was changed
// by something else than a sweep (upload, fill, prolongation) or through the pamg_update_overlaps entry.
int launch_halo(pamg_handle* h, int level, int what) {
  LevelDev& L = h->lev[level - 1];
  HaloArgs a;
  a.tnew = tnew_ptr(L); a.told = L.told; a.ovl = L.ovlb[L.ovl_cur]; a.ovl_old = L.ovl_old; a.xg = h->xg;
  a.dst_strip = h->dst_strip; a.rev = h->rev; a.strip_of = h->strip_of;
  a.bc_scale = (h->p.coarse_bc_zero && level + h->level_offset > 1) ? 0.0 : 1.0;
  a.U = h->U; a.s = L.s; a.with_old = (what == 0) ? 1 : 0; a.what = what; a.nstrips = h->plan.nstrips;
  const long long n = (long long)h->U * 3 * L.S;
  a.x.npeers = 0;
  const bool cut = what != 1 && !h->plan.peers.empty();
  if (cut && h->comm && h->p2p_enabled && !h->p2p_ready && !h->p2p_failed && !h->capturing && h->level_offset == 0) {
    int rc = p2p_setup(h);      // collective, first exchange only
    if (rc) return rc;
  }
  const bool fused_x = cut && h->p2p_ready && h->p2p_fuse;
  if (fused_x) { int rc = p2p_args(h, L, L.ovlb[L.ovl_cur], a.x); if (rc) return rc; }
  // with the exchange fused in, blocks poll for remote data after their own work: keep the grid within one wave
  int hgrid = grid_for(h, n);
  if (fused_x) {
    static int halo_resident = 0;
    if (halo_resident == 0) {
      CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&halo_resident, k_halo, TPB, 0));
      if (halo_resident < 1) halo_resident = 1;
    }
    hgrid = std::min(hgrid, h->nsm * std::min(halo_resident, 4));
  }
  k_halo<<<hgrid, TPB, 0, h->stream>>>(a);
  h->launches++;
  CK(cudaGetLastError());
  if (what == 1) return PAMG_OK;
  L.strips_valid = true;
  if (fused_x) return PAMG_OK;
  return exchange_halo(h, L, L.ovlb[L.ovl_cur]);
}

// strips of the current iterate, refreshed only if something other than a sweep touched the field
int ensure_strips(pamg_handle* h, int level) {
  LevelDev& L = h->lev[level - 1];
  if (L.strips_valid || !h->p.face_terms) return PAMG_OK;
  return launch_halo(h, level, 3);
}

int launch_build_rhs(pamg_handle* h) {
  LevelDev& L = h->lev[0];
  RhsArgs a;
  a.told = L.told; a.rhs = L.rhs; a.pc = L.pc; a.xg = h->xg; a.dt = h->p.dt; a.source_coef = h->p.source_coef;
  a.nelem = L.nelem; a.s = L.s; a.literal_source = h->p.literal_source;
  k_build_rhs<<<grid_for(h, L.nelem), TPB, 0, h->stream>>>(a);
  h->launches++;
  CK(cudaGetLastError());
  L.rhs_valid = true;
  return PAMG_OK;
}

template <int MODE>
int launch_element(pamg_handle* h, LevelDev& L, const double* Tin, double* Tout, int colour, int grid, bool write_strips = false) {
  ElemArgs a;
  a.Tin = Tin; a.Tout = Tout; a.rhs = L.rhs; a.ovl = L.ovlb[L.ovl_cur]; a.pc = L.pc; a.strip_of = h->strip_of; a.hmap = h->hmap;
  a.ovl_next = (write_strips && h->p.face_terms) ? L.ovlb[L.ovl_cur ^ 1] : nullptr; a.dst_strip = h->dst_strip; a.rev = h->rev;
  a.partial = h->partial; a.omega = h->p.omega; a.rsign = (double)h->p.residual_sign; a.nelem = L.nelem; a.s = L.s;
  a.colour = colour;
  a.split_boundary = 0; a.partial_off = 0;
  const bool prof = h->profiling && h->pev_used + 2 <= (int)h->pev.size();
  if (prof) CK(cudaEventRecord(h->pev[h->pev_used], h->stream));
  if (MODE != MODE_GS && h->kernel_mode == 2 && L.nitems > 0) {
    // row-streaming kernel: TMA-prefetched ring of row segments, all neighbours from shared memory
    StreamArgs sa;
    sa.e = a; sa.items = L.items; sa.nitems = L.nitems; sa.counters = h->counters;
    const int sgrid = std::min(L.nitems, std::min(grid, h->nsm * 5));
    if (h->p.face_terms) k_stream<MODE, true><<<sgrid, SW, 0, h->stream>>>(sa);
    else k_stream<MODE, false><<<sgrid, SW, 0, h->stream>>>(sa);
    if (MODE == MODE_RESID) h->last_partials = sgrid;
  } else if (MODE != MODE_GS && h->kernel_mode == 4 && h->win_producer && L.s >= 6 && L.s <= 8) {
    // window kernel with a producer warp (no CTA-wide barrier between tiles)
    auto kern = h->p.face_terms ? k_element_win2<MODE, true> : k_element_win2<MODE, false>;
    static int resident_win2[2] = {0, 0};
    int& resident = resident_win2[h->p.face_terms ? 1 : 0];
    if (resident == 0) {
      CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WIN_SMEM_BYTES));
      CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, kern, WIN2_THREADS, WIN_SMEM_BYTES));
      if (resident < 1) resident = 1;
    }
    const int tgrid = (int)std::max(1ll, std::min(L.nelem / TPB, (long long)h->nsm * resident));
    kern<<<tgrid, WIN2_THREADS, WIN_SMEM_BYTES, h->stream>>>(a);
    if (MODE == MODE_RESID) h->last_partials = tgrid;
  } else if ((MODE != MODE_GS || h->gs_tma) && h->kernel_mode == 4 && L.C >= TPB && L.s <= 8) {
    // (a vertical neighbour is up to 2^(s+1) children away: the 8-tile ring covers s <= 8)
    // window kernel: ring of 8 field tiles in shared memory, every neighbour value read from it
    auto kern = h->p.face_terms ? k_element_win<MODE, true> : k_element_win<MODE, false>;
    static int resident_win[2] = {0, 0};
    int& resident = resident_win[h->p.face_terms ? 1 : 0];
    if (resident == 0) {
      CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WIN_SMEM_BYTES));
      CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, kern, TPB, WIN_SMEM_BYTES));
      if (resident < 1) resident = 1;
    }
    const int tgrid = (int)std::max(1ll, std::min(L.nelem / TPB, (long long)h->nsm * resident));
    kern<<<tgrid, TPB, WIN_SMEM_BYTES, h->stream>>>(a);
    if (MODE == MODE_RESID) h->last_partials = tgrid;
  } else if ((MODE != MODE_GS || h->gs_tma) && (h->kernel_mode == 1 || h->kernel_mode == 2 || h->kernel_mode == 4) && L.C >= TPB) {
    // 1-D TMA tiles: contiguous 6 KB spans through shared memory (pamg_kernels.cuh)
    // contiguous tile ranges per CTA: exactly one wave of resident CTAs (occupancy from the runtime)
    auto kern = h->p.face_terms ? k_element_tma<MODE, true> : k_element_tma<MODE, false>;
    static int resident_by_face[2] = {0, 0};   // per instantiation
    int& resident = resident_by_face[h->p.face_terms ? 1 : 0];
    if (resident == 0) {
      CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TMA_SMEM_BYTES));
      CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, kern, TPB, TMA_SMEM_BYTES));
      if (resident < 1) resident = 1;
    }
    const int tgrid = (int)std::max(1ll, std::min((L.nelem + TPB - 1) / TPB, (long long)h->nsm * resident));
    const bool split = h->split_boundary && h->p.face_terms;
    a.split_boundary = split ? 1 : 0;
    kern<<<tgrid, TPB, TMA_SMEM_BYTES, h->stream>>>(a);
    int fgrid = 0;
    if (split && !(MODE == MODE_GS && colour == 0)) {     // every child on a parent face is an "up" child
      // children on parent faces: separate small launch (their halo look-ups would stall whole tiles)
      h->launches++;
      CK(cudaGetLastError());
      ElemArgs f = a;
      f.split_boundary = 0; f.partial_off = tgrid;
      fgrid = grid_for(h, (long long)h->U * 3 * L.S);
      k_boundary_fix<MODE><<<fgrid, TPB, 0, h->stream>>>(f, h->U);
    }
    if (MODE == MODE_RESID) h->last_partials = tgrid + fgrid;
  } else if (h->kernel_mode != 0) {
    // branch-free direct kernel (all loads of a child in flight at once); also the coloured GS pass
    if (MODE == MODE_RESID) h->last_partials = grid;
    if (h->p.face_terms) k_element_direct2<MODE, true><<<grid, TPB, 0, h->stream>>>(a);
    else k_element_direct2<MODE, false><<<grid, TPB, 0, h->stream>>>(a);
  } else {
    if (MODE == MODE_RESID) h->last_partials = grid;
    if (h->p.face_terms) k_element<MODE, true><<<grid, TPB, 0, h->stream>>>(a);
    else k_element<MODE, false><<<grid, TPB, 0, h->stream>>>(a);
  }
  if (prof) { CK(cudaEventRecord(h->pev[h->pev_used + 1], h->stream)); h->pev_used += 2; }
  h->launches++;
  CK(cudaGetLastError());
  return PAMG_OK;
}

// coloured Gauss-Seidel sweep in one pass (k_gs_win), out of place
bool gs_fused_ok(const pamg_handle* h, const LevelDev& L) {
  return h->gs_fused && h->kernel_mode == 4 && h->p.face_terms && L.C >= TPB && L.s <= 8;
}

int launch_gs_fused(pamg_handle* h, LevelDev& L, const double* Tin, double* Tout) {
  ElemArgs a;
  a.Tin = Tin; a.Tout = Tout; a.rhs = L.rhs; a.ovl = L.ovlb[L.ovl_cur]; a.pc = L.pc; a.strip_of = h->strip_of; a.hmap = h->hmap;
  a.ovl_next = nullptr; a.dst_strip = h->dst_strip; a.rev = h->rev;
  a.partial = h->partial; a.omega = h->p.omega; a.rsign = (double)h->p.residual_sign; a.nelem = L.nelem; a.s = L.s;
  a.colour = 1; a.split_boundary = 0; a.partial_off = 0;
  const bool producer = h->win_producer && L.s >= 6;
  static int resident2[2] = {0, 0};
  int& resident = resident2[producer ? 1 : 0];
  if (resident == 0) {
    if (producer) {
      CK(cudaFuncSetAttribute(k_gs_win2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GSW_SMEM_BYTES));
      CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, k_gs_win2, WIN2_THREADS, GSW_SMEM_BYTES));
    } else {
      CK(cudaFuncSetAttribute(k_gs_win, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GSW_SMEM_BYTES));
      CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, k_gs_win, TPB, GSW_SMEM_BYTES));
    }
    if (resident < 1) resident = 1;
  }
  const bool prof = h->profiling && h->pev_used + 2 <= (int)h->pev.size();
  if (prof) CK(cudaEventRecord(h->pev[h->pev_used], h->stream));
  const int tgrid = (int)std::max(1ll, std::min(L.nelem / TPB, (long long)h->nsm * resident));
  if (producer) k_gs_win2<<<tgrid, WIN2_THREADS, GSW_SMEM_BYTES, h->stream>>>(a);
  else k_gs_win<<<tgrid, TPB, GSW_SMEM_BYTES, h->stream>>>(a);
  if (prof) { CK(cudaEventRecord(h->pev[h->pev_used + 1], h->stream)); h->pev_used += 2; }
  h->launches++;
  CK(cudaGetLastError());
  return PAMG_OK;
}

int do_smooth(pamg_handle* h, int level, int solver, int nsweeps) {
  LevelDev& L = h->lev[level - 1];
  if (level == 1 && !L.rhs_valid) { int rc = launch_build_rhs(h); if (rc) return rc; }
  const int grid = grid_for(h, L.nelem);
  // optional (PAMG_FUSED_HALO=1): the default kernel families write the next sweep's strips themselves; otherwise
  // one k_halo launch per sweep refreshes the strips of the faces between parents (Dirichlet strips are static)
  const bool fused = h->fused_halo && (h->kernel_mode == 1 || h->kernel_mode == 3);
  for (int sw = 0; sw < nsweeps; ++sw) {
    // tnew <- tnew_nonlin (:550) is the buffer swap below for Jacobi; halo from it (:555)
    L.tnew_alias = true;
    if (!fused) L.strips_valid = false;
    int rc = ensure_strips(h, level);
    if (rc) return rc;
    if (solver == 1 || solver == 2) {
      rc = (solver == 1) ? launch_element<MODE_JACOBI>(h, L, L.T[L.cur], L.T[L.cur ^ 1], 0, grid, fused)
                         : launch_element<MODE_RICH>(h, L, L.T[L.cur], L.T[L.cur ^ 1], 0, grid, fused);
      if (rc) return rc;
      L.cur ^= 1;
      L.tnew_alias = false;  // the old buffer now holds the start-of-sweep field = tracer%tnew
    } else if (solver == 3 || solver == 4) {
      // two-colour ordering of the reference's Gauss-Seidel sweep: all down children, then all up children;
      // values across parent faces stay lagged through the halo strips exactly as at :647-655.
      if (gs_fused_ok(h, L)) {
        // both colours in one pass over memory, written to the other buffer (which then holds tracer%tnew, as for Jacobi)
        rc = launch_gs_fused(h, L, L.T[L.cur], L.T[L.cur ^ 1]);
        if (rc) return rc;
        L.cur ^= 1;
        L.tnew_alias = false;
      } else {
        if (sw == nsweeps - 1 && h->p.keep_tnew_gs) { rc = materialise_tnew(h, L); if (rc) return rc; }  // keep tracer%tnew observable
        rc = launch_element<MODE_GS>(h, L, L.T[L.cur], L.T[L.cur], 0, grid, false);
        if (rc) return rc;
        rc = launch_element<MODE_GS>(h, L, L.T[L.cur], L.T[L.cur], 1, grid, fused);   // all children on parent faces are "up"
        if (rc) return rc;
      }
    } else {
      return fail(h, PAMG_ERR_ARG, "solver must be 1 (Jacobi), 2 (Richardson) or 3 (Gauss-Seidel)");
    }
    if (fused && h->p.face_terms) {
      L.ovl_cur ^= 1;                 // the sweep wrote the strips of the new iterate (incl. the send slots)
      rc = exchange_halo(h, L, L.ovlb[L.ovl_cur]);
      if (rc) return rc;
      L.strips_valid = true;
    } else {
      L.strips_valid = false;
    }
  }
  return PAMG_OK;
}

int do_residual(pamg_handle* h, int level, double* l2, double* linf, double* smax) {
  LevelDev& L = h->lev[level - 1];
  if (level == 1 && !L.rhs_valid) { int rc = launch_build_rhs(h); if (rc) return rc; }
  const int grid = grid_for(h, L.nelem);
  if (2 * grid > h->npartial) return fail(h, PAMG_ERR_STATE, "partial buffer too small");
  int rc = launch_element<MODE_RESID>(h, L, tnew_ptr(L), L.res, 0, grid);
  if (rc) return rc;
  if (l2 || linf || smax) {
    k_reduce_partials<<<1, 1024, 0, h->stream>>>(h->partial, h->last_partials, h->out3);
    h->launches++;
    CK(cudaGetLastError());
    if (h->comm && h->nranks > 1) {
      // global norms: sum of squares, max |r|, max r
      g_nccl.GroupStart();
      g_nccl.AllReduce(h->out3, h->out3, 1, ncclFloat64, ncclSum, h->comm, h->stream);
      g_nccl.AllReduce(h->out3 + 1, h->out3 + 1, 2, ncclFloat64, ncclMax, h->comm, h->stream);
      if (g_nccl.GroupEnd() != ncclSuccess) return fail(h, PAMG_ERR_CUDA, "ncclAllReduce failed");
    }
    CK(cudaMemcpyAsync(h->out3_host, h->out3, 3 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (h->capturing) return PAMG_OK;      // the caller synchronises after the graph launch
    CK(cudaStreamSynchronize(h->stream));
    if (l2) *l2 = std::sqrt(h->out3_host[0]);
    if (linf) *linf = h->out3_host[1];
    if (smax) *smax = h->out3_host[2];
  }
  return PAMG_OK;
}

int do_restrict(pamg_handle* h, int fine_level) {
  if (fine_level >= (int)h->lev.size()) return PAMG_OK;  // splitting.F90:18
  LevelDev& F = h->lev[fine_level - 1];
  LevelDev& Cc = h->lev[fine_level];
  XferArgs a;
  a.src = F.res; a.dst = Cc.rhs; a.ncoarse = Cc.nelem; a.sc = Cc.s; a.mode = h->p.transfer;
  k_restrict<<<grid_for(h, Cc.nelem), TPB, 0, h->stream>>>(a);
  h->launches++;
  CK(cudaGetLastError());
  Cc.rhs_valid = true;
  return PAMG_OK;
}

int do_prolong(pamg_handle* h, int fine_level, bool keep_tnew = true) {
  if (fine_level >= (int)h->lev.size()) return fail(h, PAMG_ERR_ARG, "no coarser level to prolong from");
  LevelDev& F = h->lev[fine_level - 1];
  LevelDev& Cc = h->lev[fine_level];
  XferArgs a;
  a.ncoarse = Cc.nelem; a.sc = Cc.s; a.mode = h->p.transfer;
  if (h->p.transfer == 0) {
    // as written: tracer(ilevel)%tnew += f(tracer(ilevel+1)%tnew), splitting.F90:59-88
    int rc = materialise_tnew(h, F);
    if (rc) return rc;
    a.src = tnew_ptr(Cc); a.dst = F.T[F.cur ^ 1];
    k_prolong_literal<<<grid_for(h, Cc.nelem), TPB, 0, h->stream>>>(a);
  } else {
    if (keep_tnew) {                  // the correction goes to the iterate only: keep tracer%tnew as it was
      int rc = materialise_tnew(h, F);
      if (rc) return rc;
    } else {
      F.tnew_alias = true;            // inside the V-cycle nothing reads the pre-correction field
    }
    a.src = Cc.T[Cc.cur]; a.dst = F.T[F.cur];
    F.strips_valid = false;           // the iterate changes outside a sweep
    k_prolong_p1<<<grid_for(h, F.nelem), TPB, 0, h->stream>>>(a);
  }
  h->launches++;
  CK(cudaGetLastError());
  return PAMG_OK;
}

int do_fill(pamg_handle* h, double* p, long long n, double v) {
  if (v == 0.0) { CK(cudaMemsetAsync(p, 0, n * sizeof(double), h->stream)); return PAMG_OK; }
  k_fill<<<grid_for(h, n), TPB, 0, h->stream>>>(p, n, v);
  h->launches++;
  CK(cudaGetLastError());
  return PAMG_OK;
}

int vcycle_rec(pamg_handle* h, int level, int solver, int nu1, int nu2, int ncoarse);

// Coarse-level agglomeration: below agg_level every kernel is launch-latency bound and every sweep
// would pay an NCCL exchange, so the restricted right-hand side of all parts is gathered on part 0, the remaining
// levels of the V-cycle run there on the whole mesh (no exchange at all), and the correction is scattered back.
int agg_coarse_solve(pamg_handle* h, int solver, int nu1, int nu2, int ncoarse) {
  LevelDev& Lc = h->lev[h->agg_level - 1];
  const size_t per_parent = (size_t)3 * Lc.C;
  if (!h->comm) return fail(h, PAMG_ERR_STATE, "agglomeration needs pamg_comm_init");
  g_nccl.GroupStart();
  if (h->rank == 0) {
    LevelDev& A = h->agg->lev[0];
    for (int p = 1; p < h->nranks; ++p)
      g_nccl.Recv(A.rhs + per_parent * h->part_first[p], per_parent * (h->part_first[p + 1] - h->part_first[p]), ncclFloat64, p,
                  h->comm, h->stream);
  } else {
    g_nccl.Send(Lc.rhs, per_parent * h->U, ncclFloat64, 0, h->comm, h->stream);
  }
  if (g_nccl.GroupEnd() != ncclSuccess) return fail(h, PAMG_ERR_CUDA, "ncclGroupEnd failed in coarse gather");
  if (h->rank == 0) {
    pamg_handle* g = h->agg;
    LevelDev& A = g->lev[0];
    CK(cudaMemcpyAsync(A.rhs, Lc.rhs, per_parent * h->U * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    CK(cudaMemsetAsync(A.T[A.cur], 0, A.ndof * sizeof(double), h->stream));
    A.tnew_alias = true; A.strips_valid = false; A.rhs_valid = true;
    const long long l0 = g->launches;
    int rc = vcycle_rec(g, 1, solver, nu1, nu2, ncoarse);
    h->launches += g->launches - l0;
    if (rc) return fail(h, rc, g->err);
    CK(cudaMemcpyAsync(Lc.T[Lc.cur], A.T[A.cur], per_parent * h->U * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  }
  g_nccl.GroupStart();
  if (h->rank == 0) {
    LevelDev& A = h->agg->lev[0];
    for (int p = 1; p < h->nranks; ++p)
      g_nccl.Send(A.T[A.cur] + per_parent * h->part_first[p], per_parent * (h->part_first[p + 1] - h->part_first[p]), ncclFloat64,
                  p, h->comm, h->stream);
  } else {
    g_nccl.Recv(Lc.T[Lc.cur], per_parent * h->U, ncclFloat64, 0, h->comm, h->stream);
  }
  if (g_nccl.GroupEnd() != ncclSuccess) return fail(h, PAMG_ERR_CUDA, "ncclGroupEnd failed in coarse scatter");
  Lc.tnew_alias = true; Lc.strips_valid = false;
  return PAMG_OK;
}

// Jacobi / Richardson sweeps ping-pong between the two T buffers; a cycle that ends on the other buffer would not
// be replayable as a CUDA graph (pointers are baked in), so the iterate is moved back when the sweep count is odd.
int normalise_parity(pamg_handle* h, LevelDev& L, int cur0) {
  if (L.cur == cur0) return PAMG_OK;
  CK(cudaMemcpyAsync(L.T[L.cur ^ 1], L.T[L.cur], L.ndof * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  L.cur ^= 1;
  L.tnew_alias = true;
  return PAMG_OK;
}

int vcycle_rec(pamg_handle* h, int level, int solver, int nu1, int nu2, int ncoarse) {
  const int Lmax = (int)h->lev.size();
  int rc;
  LevelDev& L = h->lev[level - 1];
  const int cur0 = L.cur;
  if (level == Lmax) {
    if ((rc = do_smooth(h, level, solver, ncoarse))) return rc;
    return normalise_parity(h, L, cur0);
  }
  if ((rc = do_smooth(h, level, solver, nu1))) return rc;
  L.tnew_alias = true;                                   // tnew = tnew_nonlin
  if ((rc = ensure_strips(h, level))) return rc;
  if ((rc = do_residual(h, level, nullptr, nullptr, nullptr))) return rc;
  if ((rc = do_restrict(h, level))) return rc;
  LevelDev& Cc = h->lev[level];
  if ((rc = do_fill(h, Cc.T[Cc.cur], Cc.ndof, 0.0))) return rc;
  Cc.tnew_alias = true;
  Cc.strips_valid = false;
  if (h->agg_level && level + 1 == h->agg_level) {
    if ((rc = agg_coarse_solve(h, solver, nu1, nu2, ncoarse))) return rc;
  } else {
    if ((rc = vcycle_rec(h, level + 1, solver, nu1, nu2, ncoarse))) return rc;
  }
  if ((rc = do_prolong(h, level, false))) return rc;
  if ((rc = do_smooth(h, level, solver, nu2))) return rc;
  return normalise_parity(h, L, cur0);
}

// one V-cycle followed by the residual norms of level 1 (the copy to pinned memory is queued, not awaited)
int vcycle_body(pamg_handle* h, int solver, int nu1, int nu2, int ncoarse) {
  int rc;
  if ((rc = vcycle_rec(h, 1, solver, nu1, nu2, ncoarse))) return rc;
  LevelDev& L = h->lev[0];
  L.tnew_alias = true;
  if ((rc = ensure_strips(h, 1))) return rc;
  double dummy;
  const bool cap = h->capturing;
  h->capturing = true;    // do_residual: queue the read-back only
  rc = do_residual(h, 1, &dummy, nullptr, nullptr);
  h->capturing = cap;
  return rc;
}

void free_levels(pamg_handle* h) {
  for (auto& g : h->vc_graphs) cudaGraphExecDestroy(g.exec);
  h->vc_graphs.clear();
  for (auto& L : h->lev) {
    cudaFree(L.T[0]); cudaFree(L.T[1]); cudaFree(L.spare); cudaFree(L.told); cudaFree(L.rhs); cudaFree(L.res);
    cudaFree(L.ovlb[0]); cudaFree(L.ovlb[1]); cudaFree(L.ovl_old); cudaFree(L.pc); cudaFree(L.items);
  }
  h->lev.clear();
  cudaFree(h->xg); cudaFree(h->strip_of); cudaFree(h->dst_strip); cudaFree(h->rev); cudaFree(h->hmap);
  cudaFree(h->partial); cudaFree(h->counters); h->counters = nullptr;
  h->xg = nullptr; h->strip_of = h->dst_strip = h->rev = h->hmap = nullptr; h->partial = nullptr;
}

int ensure_stage(pamg_handle* h, size_t bytes) {
  if (h->stage_bytes >= bytes) return PAMG_OK;
  if (h->stage) cudaFreeHost(h->stage);
  h->stage = nullptr; h->stage_bytes = 0;
  CK(cudaMallocHost(&h->stage, bytes));
  h->stage_bytes = bytes;
  return PAMG_OK;
}

}  // namespace

// ================================================================== C ABI
extern "C" {

const char* pamg_version(void) { return "pamg-b200 0.1 (sm_100a)"; }

int pamg_device_count(int* n) {
  if (!n) return PAMG_ERR_ARG;
  int c = 0;
  cudaError_t e = cudaGetDeviceCount(&c);
  if (e != cudaSuccess || c < 1) { *n = 0; return PAMG_ERR_CUDA; }
  *n = c;
  return PAMG_OK;
}

void pamg_default_params(pamg_params* p, int literal_head) {
  if (!p) return;
  std::memset(p, 0, sizeof(*p));
  p->n_split = 1; p->multi_levels = 1; p->n_smooth = 4; p->n_multigrid = 2; p->n_coarse_smooth = 15; p->solver = 3;
  p->theta = 1.0; p->k = 1.0; p->omega = 0.8; p->u_x = 0.0; p->u_y = 0.0;
  if (literal_head) {  // main.F90:46-47, transport_tri_semi.F90:117-140 as checked in
    p->dt = 1.25e-5; p->face_terms = 0; p->literal_source = 1; p->transfer = 0; p->residual_sign = 1;
    p->halo_rule = 0; p->coarse_bc_zero = 0; p->source_coef = -2.0; p->keep_tnew_gs = 1;
  } else {
    p->dt = 1e-3; p->face_terms = 1; p->literal_source = 0; p->transfer = 1; p->residual_sign = -1;
    p->halo_rule = 1; p->coarse_bc_zero = 1; p->source_coef = 2.0; p->keep_tnew_gs = 0;
  }
}

int pamg_create(const pamg_params* p, int device, pamg_handle** out) {
  if (!p || !out) return PAMG_ERR_ARG;
  *out = nullptr;
  if (p->n_split < 1 || p->n_split > 13 || p->multi_levels < 1 || p->multi_levels > p->n_split) return PAMG_ERR_ARG;  // :120-123
  if (p->theta != 1.0) return PAMG_ERR_UNSUPPORTED;
  if (!(p->dt > 0.0)) return PAMG_ERR_ARG;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) return PAMG_ERR_CUDA;  // no CPU fallback
  if (device < 0 || device >= ndev) return PAMG_ERR_ARG;
  pamg_handle* h = new pamg_handle();
  h->p = *p; h->device = device;
  {
    const char* e = getenv("PAMG_KERNEL");
    if (e && !strcmp(e, "direct")) h->kernel_mode = 0;
    else if (e && !strcmp(e, "tma1d")) h->kernel_mode = 1;
    else if (e && !strcmp(e, "stream")) h->kernel_mode = 2;
    else if (e && !strcmp(e, "direct2")) h->kernel_mode = 3;
    else if (e && !strcmp(e, "win")) h->kernel_mode = 4;
    const char* wp = getenv("PAMG_WIN");
    if (wp && !strcmp(wp, "producer")) h->win_producer = true;
    if (wp && !strcmp(wp, "barrier")) h->win_producer = false;
    const char* pp = getenv("PAMG_P2P");
    if (pp && pp[0] == '0') h->p2p_enabled = false;
    const char* pt = getenv("PAMG_P2P_TIMEOUT_S");
    if (pt && atof(pt) > 0.0) h->p2p_timeout_ns = (unsigned long long)(atof(pt) * 1e9);
    const char* pf = getenv("PAMG_P2P_FUSE");
    if (pf && pf[0] == '0') h->p2p_fuse = false;
    const char* gr = getenv("PAMG_GRAPH");
    if (gr && gr[0] == '0') h->use_graph = false;
    const char* gn = getenv("PAMG_GRAPH_NCCL");
    if (gn && gn[0] == '0') h->graph_nccl = false;
    const char* sp = getenv("PAMG_SPLIT");
    if (sp && sp[0] == '1') h->split_boundary = true;
    const char* fh = getenv("PAMG_FUSED_HALO");
    if (fh && fh[0] == '1') h->fused_halo = true;
    const char* g = getenv("PAMG_GS");
    if (g && !strcmp(g, "direct")) { h->gs_tma = false; h->gs_fused = false; }
    if (g && !strcmp(g, "twopass")) h->gs_fused = false;
  }
  if (cudaSetDevice(device) != cudaSuccess) { delete h; return PAMG_ERR_CUDA; }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) h->nsm = prop.multiProcessorCount;
  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) { delete h; return PAMG_ERR_CUDA; }
  for (auto& e : h->ev) cudaEventCreate(&e);
  cudaMalloc(&h->out3, 3 * sizeof(double));
  cudaMallocHost(&h->out3_host, 3 * sizeof(double));
  *out = h;
  return PAMG_OK;
}

void pamg_destroy(pamg_handle* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  // graphs that contain NCCL nodes must go before the communicator
  for (auto& g : h->vc_graphs) cudaGraphExecDestroy(g.exec);
  h->vc_graphs.clear();
  if (h->stream) cudaStreamSynchronize(h->stream);
  if (h->agg) { pamg_destroy(h->agg); h->agg = nullptr; }
  if (h->comm) g_nccl.CommDestroy(h->comm);
  p2p_close(h);
  free_levels(h);
  unstr_free(h->un);
  cudaFree(h->out3); cudaFreeHost(h->out3_host); cudaFree(h->scratch);
  if (h->stage) cudaFreeHost(h->stage);
  for (auto& e : h->ev) if (e) cudaEventDestroy(e);
  for (auto& e : h->pev) if (e) cudaEventDestroy(e);
  if (h->up_stream) { cudaStreamSynchronize(h->up_stream); cudaStreamSynchronize(h->down_stream); cudaStreamDestroy(h->up_stream); cudaStreamDestroy(h->down_stream); cudaEventDestroy(h->ev_up); cudaEventDestroy(h->ev_comp); cudaEventDestroy(h->ev_down[0]); cudaEventDestroy(h->ev_down[1]); }
  if (h->stream && !h->shared_stream) cudaStreamDestroy(h->stream);
  delete h;
}

const char* pamg_last_error(const pamg_handle* h) { return h ? h->err.c_str() : "null handle"; }

int pamg_set_parents_partition(pamg_handle* h, int U_global, const double* X, const int32_t* neig,
                               const int32_t* fneig, const int32_t* dir, int nparts, const int32_t* part_first,
                               int my_part) {
  if (!h) return PAMG_ERR_ARG;
  CK(cudaSetDevice(h->device));
  int rc = build_halo_plan(U_global, X, neig, fneig, dir, h->p.halo_rule, nparts, part_first, my_part, h->plan);
  if (rc) return fail(h, rc, "inconsistent parent arrays (X / Neig / fNeig)");
  free_levels(h);
  if (h->agg) { pamg_destroy(h->agg); h->agg = nullptr; }
  h->agg_level = 0;
  // where my cut-face strips live in every peer's strip space (the peer's own plan, rebuilt here on the host)
  p2p_close(h);
  h->p2p_peers.assign(h->plan.peers.size(), pamg_handle::P2PPeer());
  for (size_t i = 0; i < h->plan.peers.size(); ++i) {
    HaloPlan theirs;
    if (build_halo_plan(U_global, X, neig, fneig, dir, h->p.halo_rule, nparts, part_first, h->plan.peers[i].part, theirs)) continue;
    for (size_t j = 0; j < theirs.peers.size(); ++j)
      h->p2p_peers[i].recv_strips_at_peer = std::max(h->p2p_peers[i].recv_strips_at_peer, (long long)theirs.peers[j].strip_begin + theirs.peers[j].nfaces);
    for (size_t j = 0; j < theirs.peers.size(); ++j)
      if (theirs.peers[j].part == my_part && theirs.peers[j].nfaces == h->plan.peers[i].nfaces) {
        h->p2p_peers[i].slot_at_peer = (int)j;
        h->p2p_peers[i].strip_begin_at_peer = theirs.peers[j].strip_begin;
      }
  }
  const int U = h->plan.U_local, first = h->plan.first;
  if (U < 1) return fail(h, PAMG_ERR_ARG, "empty partition");
  h->U = U; h->U_global = U_global;
  h->part_first.assign(nparts + 1, 0);
  if (part_first) for (int i = 0; i <= nparts; ++i) h->part_first[i] = part_first[i]; else h->part_first[1] = U_global;
  std::vector<double> xg((size_t)U * 6);
  for (int u = 0; u < U; ++u) {
    const double* P = X + (size_t)(first + u) * 6;
    double* o = &xg[(size_t)u * 6];
    o[0] = P[4]; o[1] = P[5]; o[2] = P[0] - P[4]; o[3] = P[1] - P[5]; o[4] = P[2] - P[4]; o[5] = P[3] - P[5];
  }
  CK(cudaMalloc(&h->xg, xg.size() * sizeof(double)));
  CK(cudaMemcpy(h->xg, xg.data(), xg.size() * sizeof(double), cudaMemcpyHostToDevice));
  const size_t n3 = (size_t)U * 3 * sizeof(int32_t);
  CK(cudaMalloc(&h->strip_of, n3)); CK(cudaMalloc(&h->dst_strip, n3)); CK(cudaMalloc(&h->rev, n3)); CK(cudaMalloc(&h->hmap, n3));
  CK(cudaMemcpy(h->strip_of, h->plan.strip_of.data(), n3, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(h->dst_strip, h->plan.dst_strip.data(), n3, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(h->rev, h->plan.rev.data(), n3, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(h->hmap, h->plan.hmap.data(), n3, cudaMemcpyHostToDevice));
  h->lev.resize(h->p.multi_levels);
  std::vector<double> pc((size_t)U * NPC);
  for (int il = 0; il < h->p.multi_levels; ++il) {
    LevelDev& L = h->lev[il];
    L.s = h->p.n_split - il; L.S = 1 << L.s; L.C = 1ll << (2 * L.s);
    L.nelem = L.C * U; L.ndof = 3 * L.nelem;
    const size_t fb = (size_t)L.ndof * sizeof(double);
    CK(cudaMalloc(&L.T[0], fb)); CK(cudaMalloc(&L.T[1], fb)); CK(cudaMalloc(&L.told, fb));
    CK(cudaMalloc(&L.rhs, fb)); CK(cudaMalloc(&L.res, fb));
    CK(cudaMemsetAsync(L.T[0], 0, fb, h->stream)); CK(cudaMemsetAsync(L.T[1], 0, fb, h->stream));
    CK(cudaMemsetAsync(L.told, 0, fb, h->stream)); CK(cudaMemsetAsync(L.rhs, 0, fb, h->stream));
    CK(cudaMemsetAsync(L.res, 0, fb, h->stream));
    const size_t ob = (size_t)(h->plan.nstrips + h->plan.nsend) * 3 * L.S * sizeof(double);
    CK(cudaMalloc(&L.ovlb[0], ob)); CK(cudaMalloc(&L.ovlb[1], ob)); CK(cudaMalloc(&L.ovl_old, ob));
    CK(cudaMemsetAsync(L.ovlb[0], 0, ob, h->stream)); CK(cudaMemsetAsync(L.ovlb[1], 0, ob, h->stream));
    CK(cudaMemsetAsync(L.ovl_old, 0, ob, h->stream));  // :207
    L.ovl_cur = 0; L.strips_valid = false;
    for (int u = 0; u < U; ++u) parent_coefficients(h->p, X, neig, first + u, L.s, &pc[(size_t)u * NPC]);
    CK(cudaMalloc(&L.pc, pc.size() * sizeof(double)));
    CK(cudaMemcpy(L.pc, pc.data(), pc.size() * sizeof(double), cudaMemcpyHostToDevice));
    L.cur = 0; L.tnew_alias = true; L.rhs_valid = (il != 0);
    // work list of the row-streaming kernel: (parent, row chunk, column strip) cut out of the triangle of
    // children in (r, x = ipos + r - 1) coordinates; largest items first (dynamic scheduling)
    if (L.s >= STREAM_MIN_S) {
      const int b = 2 << L.s;
      struct It { int u, r0, x0, n; };
      std::vector<It> its;
      for (int u = 0; u < U; ++u)
        for (int r0 = 1; r0 <= L.S; r0 += SR)
          for (int x0 = 1; x0 <= b - 1; x0 += SW) {
            const int x1 = x0 + SW - 1;
            const int rend = std::min(std::min(r0 + SR - 1, L.S), std::min(x1, b - x0));
            if (rend < r0) continue;
            int n = 0;
            for (int r = r0; r <= rend; ++r) n += std::min(x1, b - r) - std::max(x0, r) + 1;
            its.push_back(It{u, r0, x0, n});
          }
      std::stable_sort(its.begin(), its.end(), [](const It& p, const It& q) { return p.n > q.n; });
      std::vector<int2> packed(its.size());
      for (size_t i = 0; i < its.size(); ++i) packed[i] = make_int2(its[i].u, (its[i].r0 << 16) | its[i].x0);
      L.nitems = (int)packed.size();
      CK(cudaMalloc(&L.items, packed.size() * sizeof(int2)));
      CK(cudaMemcpy(L.items, packed.data(), packed.size() * sizeof(int2), cudaMemcpyHostToDevice));
    }
  }
  CK(cudaMalloc(&h->counters, 2 * sizeof(int)));
  CK(cudaMemset(h->counters, 0, 2 * sizeof(int)));
  // Dirichlet data sin(x+y) on domain-boundary faces never changes: fill those strips once per level
  for (int il = 1; il <= h->p.multi_levels; ++il)
    for (int bsel = 0; bsel < 2; ++bsel) {
      h->lev[il - 1].ovl_cur = bsel;
      int rc2 = launch_halo(h, il, 1);
      if (rc2) return rc2;
    }
  for (auto& Lv : h->lev) Lv.ovl_cur = 0;
  // coarse-level agglomeration on part 0: first level whose GLOBAL size is small enough to be launch-bound
  if (nparts > 1 && h->level_offset == 0) {
    const char* e = getenv("PAMG_AGG_ELEMS");
    // elements of the whole mesh on that level; 0 disables.  With the halo exchange at ~1.5 us per sweep the break-even
    // moved down: 2^21 cost 2.6 ms per solve at 4 GPUs against 2^17 (profiles/README.md)
    const long long limit = e ? atoll(e) : (1ll << 17);
    int lvl = 0;
    for (int il = 2; il <= h->p.multi_levels; ++il)
      if (limit > 0 && (long long)U_global * h->lev[il - 1].C <= limit) { lvl = il; break; }
    if (lvl) {
      h->agg_level = lvl;
      if (my_part == 0) {
        pamg_params ap = h->p;
        ap.n_split = h->lev[lvl - 1].s;
        ap.multi_levels = h->p.multi_levels - lvl + 1;
        pamg_handle* g = new pamg_handle();
        g->p = ap; g->device = h->device; g->nsm = h->nsm; g->kernel_mode = h->kernel_mode; g->gs_tma = h->gs_tma; g->gs_fused = h->gs_fused; g->win_producer = h->win_producer;
        g->stream = h->stream; g->shared_stream = true; g->level_offset = lvl - 1;
        for (auto& ev : g->ev) cudaEventCreate(&ev);
        cudaMalloc(&g->out3, 3 * sizeof(double));
        cudaMallocHost(&g->out3_host, 3 * sizeof(double));
        h->agg = g;
        int rc3 = pamg_set_parents_partition(g, U_global, X, neig, fneig, dir, 1, nullptr, 0);
        if (rc3) return fail(h, rc3, std::string("agglomerated coarse problem: ") + g->err);
      }
    }
  }
  h->npartial = h->nsm * 16;
  CK(cudaMalloc(&h->partial, (size_t)h->npartial * 3 * sizeof(double)));
  CK(cudaStreamSynchronize(h->stream));
  return PAMG_OK;
}

int pamg_set_parents(pamg_handle* h, int U, const double* X, const int32_t* neig, const int32_t* fneig,
                     const int32_t* dir) {
  return pamg_set_parents_partition(h, U, X, neig, fneig, dir, 1, nullptr, 0);
}

int pamg_ndof(const pamg_handle* h, int level, int64_t* ndof) {
  if (!valid_level(h, level) || !ndof) return PAMG_ERR_ARG;
  *ndof = h->lev[level - 1].ndof;
  return PAMG_OK;
}

int pamg_upload_field(pamg_handle* h, int field, int level, const double* host) {
  if (!valid_level(h, level) || !host) return PAMG_ERR_ARG;
  CK(cudaSetDevice(h->device));
  int rc;
  double* d = field_ptr(h, field, level, true, &rc);
  if (rc) return rc;
  CK(cudaMemcpyAsync(d, host, h->lev[level - 1].ndof * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return PAMG_OK;
}

int pamg_download_field(pamg_handle* h, int field, int level, double* host) {
  if (!valid_level(h, level) || !host) return PAMG_ERR_ARG;
  CK(cudaSetDevice(h->device));
  int rc;
  double* d = field_ptr(h, field, level, false, &rc);
  if (rc) return rc;
  CK(cudaMemcpyAsync(host, d, h->lev[level - 1].ndof * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return PAMG_OK;
}

int pamg_fill_field(pamg_handle* h, int field, int level, double value) {
  if (!valid_level(h, level)) return PAMG_ERR_ARG;
  CK(cudaSetDevice(h->device));
  int rc;
  double* d = field_ptr(h, field, level, true, &rc);
  if (rc) return rc;
  return do_fill(h, d, h->lev[level - 1].ndof, value);
}

int pamg_copy_field(pamg_handle* h, int level, int dst_field, int src_field) {
  if (!valid_level(h, level)) return PAMG_ERR_ARG;
  CK(cudaSetDevice(h->device));
  LevelDev& L = h->lev[level - 1];
  if (dst_field == src_field) return PAMG_OK;
  if (dst_field == PAMG_TNEW && src_field == PAMG_TNONLIN) { L.tnew_alias = true; return PAMG_OK; }      // :550
  if (dst_field == PAMG_TNONLIN && src_field == PAMG_TNEW) {                                             // :327
    if (!L.tnew_alias) { L.cur ^= 1; L.tnew_alias = true; L.strips_valid = false; }
    return PAMG_OK;
  }
  int rc;
  const double* s = field_ptr(h, src_field, level, false, &rc);
  if (rc) return rc;
  double* d = field_ptr(h, dst_field, level, true, &rc);
  if (rc) return rc;
  CK(cudaMemcpyAsync(d, s, L.ndof * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  return PAMG_OK;
}

int pamg_download_overlap(pamg_handle* h, int level, int old, double* host) {
  if (!valid_level(h, level) || !host) return PAMG_ERR_ARG;
  CK(cudaSetDevice(h->device));
  LevelDev& L = h->lev[level - 1];
  const size_t S3 = (size_t)3 * L.S;
  std::vector<double> tmp((size_t)h->plan.nstrips * S3);
  CK(cudaMemcpyAsync(tmp.data(), old ? L.ovl_old : L.ovlb[L.ovl_cur], tmp.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  for (int lf = 0; lf < h->U * 3; ++lf)
    std::memcpy(host + (size_t)lf * S3, &tmp[(size_t)h->plan.strip_of[lf] * S3], S3 * sizeof(double));
  return PAMG_OK;
}

int pamg_device_ptr(pamg_handle* h, int field, int level, void** dptr) {
  if (!valid_level(h, level) || !dptr) return PAMG_ERR_ARG;
  int rc;
  *dptr = field_ptr(h, field, level, false, &rc);
  return rc;
}

int pamg_update_overlaps(pamg_handle* h, int level) {
  if (!valid_level(h, level)) return PAMG_ERR_ARG;
  CK(cudaSetDevice(h->device));
  return launch_halo(h, level);
}

int pamg_build_rhs(pamg_handle* h) {
  if (!valid_level(h, 1)) return PAMG_ERR_STATE;
  CK(cudaSetDevice(h->device));
  return launch_build_rhs(h);
}

int pamg_smooth(pamg_handle* h, int level, int solver, int nsweeps) {
  if (!valid_level(h, level) || nsweeps < 0) return PAMG_ERR_ARG;
  CK(cudaSetDevice(h->device));
  return do_smooth(h, level, solver, nsweeps);
}

int pamg_residual(pamg_handle* h, int level, double* l2, double* linf) {
  if (!valid_level(h, level)) return PAMG_ERR_ARG;
  CK(cudaSetDevice(h->device));
  return do_residual(h, level, l2, linf, nullptr);
}

int pamg_convergence(pamg_handle* h, int level, double* conv) {
  if (!valid_level(h, level) || !conv) return PAMG_ERR_ARG;
  CK(cudaSetDevice(h->device));
  double l2, li;
  return do_residual(h, level, &l2, &li, conv);
}

int pamg_restrict(pamg_handle* h, int fine_level) {
  if (!valid_level(h, fine_level)) return PAMG_ERR_ARG;
  CK(cudaSetDevice(h->device));
  return do_restrict(h, fine_level);
}

int pamg_prolong(pamg_handle* h, int fine_level) {
  if (!valid_level(h, fine_level)) return PAMG_ERR_ARG;
  CK(cudaSetDevice(h->device));
  return do_prolong(h, fine_level);
}

int pamg_vcycle_solve(pamg_handle* h, int solver, int nu1, int nu2, int ncoarse, int max_cycles, double tol,
                      int* cycles, double* hist) {
  if (!valid_level(h, 1) || max_cycles < 0) return PAMG_ERR_ARG;
  CK(cudaSetDevice(h->device));
  LevelDev& L = h->lev[0];
  int rc;
  L.tnew_alias = true;
  L.strips_valid = false;
  if ((rc = ensure_strips(h, 1))) return rc;
  double r0 = 0, r = 0;
  if ((rc = do_residual(h, 1, &r0, nullptr, nullptr))) return rc;
  if (hist) hist[0] = r0;
  if (cycles) *cycles = 0;
  if (r0 == 0.0) return PAMG_OK;
  // cycle 1 runs eagerly (it also sets the kernels' attributes); from cycle 2 on the identical launch sequence
  // (~170 launches, most of them on launch-bound coarse levels) is replayed as one CUDA graph
  const long long key0 = ((((long long)solver * 64 + nu1) * 64 + nu2) * 64 + ncoarse);
  // with NCCL in the cycle (halo exchange, coarse gather/scatter) the capture is attempted once; if the library
  // refuses, the handle falls back to eager launches for good
  const bool graph_ok = h->use_graph && !h->profiling && (!h->comm || h->graph_nccl);
  for (int c = 1; c <= max_cycles; ++c) {
    if (c >= 2 && graph_ok) {
      long long key = key0;                      // the graph bakes in which of the two T buffers holds the iterate
      for (auto& Lv : h->lev) key = key * 2 + Lv.cur;
      pamg_handle::VcGraph* vg = nullptr;
      for (auto& g : h->vc_graphs) if (g.key == key) vg = &g;
      if (!vg) {
        if (h->vc_graphs.size() >= 8) { cudaGraphExecDestroy(h->vc_graphs.front().exec); h->vc_graphs.erase(h->vc_graphs.begin()); }
        cudaGraph_t graph = nullptr;
        const long long l0 = h->launches;
        CK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
        h->capturing = true;
        rc = vcycle_body(h, solver, nu1, nu2, ncoarse);
        h->capturing = false;
        cudaError_t ce = cudaStreamEndCapture(h->stream, &graph);
        pamg_handle::VcGraph ng{key, nullptr, h->launches - l0};
        h->launches = l0;
        if (!rc && ce == cudaSuccess) {
          ce = cudaGraphInstantiate(&ng.exec, graph, 0);
        }
        if (graph) cudaGraphDestroy(graph);
        if (rc || ce != cudaSuccess) {
          if (!h->comm) {
            if (rc) return rc;
            return fail(h, PAMG_ERR_CUDA, std::string("CUDA graph capture of the V-cycle: ") + cudaGetErrorString(ce));
          }
          // NCCL inside the capture was refused: run eagerly from now on (host-side state is periodic per cycle)
          (void)cudaGetLastError();
          h->graph_nccl = false;
          if ((rc = vcycle_body(h, solver, nu1, nu2, ncoarse))) return rc;
          goto cycle_done;
        }
        h->vc_graphs.push_back(ng);
        vg = &h->vc_graphs.back();
      }
      CK(cudaGraphLaunch(vg->exec, h->stream));
      h->launches += vg->launches;
    } else {
      if ((rc = vcycle_body(h, solver, nu1, nu2, ncoarse))) return rc;
    }
  cycle_done:
    CK(cudaStreamSynchronize(h->stream));
    r = std::sqrt(h->out3_host[0]);
    if (hist) hist[c] = r;
    if (cycles) *cycles = c;
    if (r / r0 <= tol) return PAMG_OK;
  }
  if (cycles) *cycles = max_cycles + 1;
  return PAMG_OK;
}

int pamg_literal_timestep(pamg_handle* h, int solver, int n_multigrid, int n_smooth) {
  if (!valid_level(h, 1)) return PAMG_ERR_STATE;
  CK(cudaSetDevice(h->device));
  const int ML = (int)h->lev.size();
  int rc;
  if ((rc = pamg_copy_field(h, 1, PAMG_TOLD, PAMG_TNEW))) return rc;      // :316
  if ((rc = pamg_copy_field(h, 1, PAMG_TNONLIN, PAMG_TNEW))) return rc;   // :317
  for (int mg = 0; mg < n_multigrid; ++mg) {
    for (int il = 1; il <= ML; ++il) {
      if ((rc = pamg_copy_field(h, il, PAMG_TNONLIN, PAMG_TNEW))) return rc;   // :327
      if ((rc = do_smooth(h, il, solver, n_smooth))) return rc;                // :331
      if ((rc = do_restrict(h, il))) return rc;                                // :336
      if ((rc = do_residual(h, il, nullptr, nullptr, nullptr))) return rc;     // :338
    }
    if ((rc = pamg_copy_field(h, ML, PAMG_TNONLIN, PAMG_TNEW))) return rc;     // :348
    for (int i = 0; i < h->p.n_coarse_smooth; ++i)
      if ((rc = do_smooth(h, ML, solver, n_smooth))) return rc;                // :351-352
    for (int il = ML - 1; il >= 1; --il) {
      if ((rc = pamg_copy_field(h, il, PAMG_TNONLIN, PAMG_TNEW))) return rc;   // :367
      if ((rc = do_prolong(h, il))) return rc;                                 // :370
      if ((rc = do_smooth(h, il, solver, n_smooth))) return rc;                // :376
    }
  }
  return PAMG_OK;
}

int pamg_timestep_host(pamg_handle* h, const double* tnew_in, double* tnew_out, int max_cycles, double tol,
                       int* cycles, double* relres) {
  if (!valid_level(h, 1) || !tnew_in || !tnew_out) return PAMG_ERR_ARG;
  CK(cudaSetDevice(h->device));
  LevelDev& L = h->lev[0];
  const size_t bytes = (size_t)L.ndof * sizeof(double);
  // told = tnew ; tnew_nonlin = tnew (transport_tri_semi.F90:316-317)
  CK(cudaMemcpyAsync(L.T[L.cur], tnew_in, bytes, cudaMemcpyHostToDevice, h->stream));
  L.tnew_alias = true;
  L.strips_valid = false;
  CK(cudaMemcpyAsync(L.told, L.T[L.cur], bytes, cudaMemcpyDeviceToDevice, h->stream));
  L.rhs_valid = false;
  std::vector<double> hist((size_t)max_cycles + 2, 0.0);
  int cyc = 0;
  int rc = pamg_vcycle_solve(h, h->p.solver, h->p.n_smooth, h->p.n_smooth, h->p.n_coarse_smooth, max_cycles, tol, &cyc,
                             hist.data());
  if (rc) return rc;
  if (cycles) *cycles = cyc;
  if (relres) *relres = hist[0] > 0 ? hist[std::min(cyc, max_cycles)] / hist[0] : 0.0;
  CK(cudaMemcpyAsync(tnew_out, L.T[L.cur], bytes, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return PAMG_OK;
}

// smoother with HOST buffers, pipelined across calls: upload(k+1) runs while download(k) is still in flight (PCIe is full
// duplex); tnew_out of call k is complete after the next call that reuses it has returned, or after pamg_sync
int pamg_smooth_host(pamg_handle* h, int solver, int nsweeps, const double* tnew_in, double* tnew_out) {
  if (!valid_level(h, 1) || !tnew_in || !tnew_out || nsweeps < 1) return PAMG_ERR_ARG;
  CK(cudaSetDevice(h->device));
  LevelDev& L = h->lev[0];
  const size_t bytes = (size_t)L.ndof * sizeof(double);
  if (!h->up_stream) {
    CK(cudaStreamCreateWithFlags(&h->up_stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&h->down_stream, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&h->ev_up, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&h->ev_comp, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&h->ev_down[0], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&h->ev_down[1], cudaEventDisableTiming));
  }
  if (!L.spare) CK(cudaMalloc(&L.spare, bytes));
  // Three field buffers rotate: the result of call k is downloaded out of one while call k+1 sweeps between the
  // other two and call k+2 uploads into the first again - so the upload only waits for the download issued two calls ago.
  if (h->pipe_calls >= 2) CK(cudaStreamWaitEvent(h->up_stream, h->ev_down[h->pipe_calls & 1], 0));
  else { CK(cudaEventRecord(h->ev_comp, h->stream)); CK(cudaStreamWaitEvent(h->up_stream, h->ev_comp, 0)); }
  CK(cudaMemcpyAsync(L.spare, tnew_in, bytes, cudaMemcpyHostToDevice, h->up_stream));
  CK(cudaEventRecord(h->ev_up, h->up_stream));
  CK(cudaStreamWaitEvent(h->stream, h->ev_up, 0));
  // cached V-cycle graphs have the old buffer addresses baked in
  if (!h->vc_graphs.empty()) { CK(cudaStreamSynchronize(h->stream)); for (auto& g : h->vc_graphs) cudaGraphExecDestroy(g.exec); h->vc_graphs.clear(); }
  std::swap(L.T[L.cur], L.spare);   // the uploaded field becomes the iterate; the previous result stays in `spare` for its download
  L.tnew_alias = true;              // tnew_nonlin = tnew = the uploaded field (transport_tri_semi.F90:317)
  L.strips_valid = false;
  int rc = do_smooth(h, 1, solver, nsweeps);
  if (rc) return rc;
  CK(cudaEventRecord(h->ev_comp, h->stream));
  CK(cudaStreamWaitEvent(h->down_stream, h->ev_comp, 0));
  CK(cudaMemcpyAsync(tnew_out, L.T[L.cur], bytes, cudaMemcpyDeviceToHost, h->down_stream));
  CK(cudaEventRecord(h->ev_down[h->pipe_calls & 1], h->down_stream));
  h->pipe_calls++;
  return PAMG_OK;
}

// ---- distributed ------------------------------------------------------------------------------
int pamg_comm_unique_id(char* id128) {
  if (!id128) return PAMG_ERR_ARG;
  if (!g_nccl.load()) return PAMG_ERR_STATE;
  ncclUniqueId id;
  if (g_nccl.GetUniqueId(&id) != ncclSuccess) return PAMG_ERR_CUDA;
  std::memcpy(id128, id.internal, 128);
  return PAMG_OK;
}

int pamg_comm_init(pamg_handle* h, const char* id128, int nranks, int rank) {
  if (!h || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return PAMG_ERR_ARG;
  if (!g_nccl.load()) return fail(h, PAMG_ERR_STATE, "libnccl.so.2 could not be loaded");
  CK(cudaSetDevice(h->device));
  ncclUniqueId id;
  std::memcpy(id.internal, id128, 128);
  if (g_nccl.CommInitRank(&h->comm, nranks, id, rank) != ncclSuccess) return fail(h, PAMG_ERR_CUDA, "ncclCommInitRank failed");
  h->nranks = nranks; h->rank = rank;
  return PAMG_OK;
}

int pamg_halo_peer_count(const pamg_handle* h, int* npeers) {
  if (!h || !npeers) return PAMG_ERR_ARG;
  *npeers = (int)h->plan.peers.size();
  return PAMG_OK;
}

int pamg_halo_peer_info(const pamg_handle* h, int idx, int* peer_part, int* nfaces) {
  if (!h || idx < 0 || idx >= (int)h->plan.peers.size()) return PAMG_ERR_ARG;
  if (peer_part) *peer_part = h->plan.peers[idx].part;
  if (nfaces) *nfaces = h->plan.peers[idx].nfaces;
  return PAMG_OK;
}

// ---- unstructured explicit step ------------------------------------------------------------------
int pamg_set_unstructured(pamg_handle* h, int E, const double* X, const int32_t* neig, const int32_t* fneig) {
  if (!h || E < 1 || !X || !neig || !fneig) return PAMG_ERR_ARG;
  CK(cudaSetDevice(h->device));
  std::string e;
  int rc = unstr_setup(h->un, E, X, neig, fneig, h->stream, e);
  if (rc) return fail(h, rc, e);
  return PAMG_OK;
}

int pamg_unstr_upload(pamg_handle* h, const double* tnew) {
  if (!h || !tnew || h->un.E < 1) return PAMG_ERR_ARG;
  CK(cudaSetDevice(h->device));
  CK(cudaMemcpyAsync(h->un.T[h->un.cur], tnew, (size_t)h->un.E * 3 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return PAMG_OK;
}

int pamg_unstr_download(pamg_handle* h, double* tnew) {
  if (!h || !tnew || h->un.E < 1) return PAMG_ERR_ARG;
  CK(cudaSetDevice(h->device));
  CK(cudaMemcpyAsync(tnew, h->un.T[h->un.cur], (size_t)h->un.E * 3 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return PAMG_OK;
}

int pamg_explicit_step(pamg_handle* h, double dt, double u_x, double u_y, double t_bc, int ntime, int nits,
                       int njac_its, int use_exact_minv, int use_dir) {
  if (!h || h->un.E < 1 || ntime < 0 || nits < 1 || njac_its < 0) return PAMG_ERR_ARG;
  CK(cudaSetDevice(h->device));
  std::string e;
  long long nl = 0;
  int rc = unstr_step(h->un, dt, u_x, u_y, t_bc, ntime, nits, njac_its, use_exact_minv, use_dir, h->nsm, h->stream, nl, e);
  h->launches += nl;
  if (rc) return fail(h, rc, e);
  return PAMG_OK;
}

// ---- unstructured implicit operator in block-CSR (unstr_implicit, transport_tri_unstr.F90:214-387) --------
int pamg_implicit_assemble(pamg_handle* h, double dt, double u_x, double u_y, int use_dir) {
  if (!h || h->un.E < 1 || !(dt > 0.0)) return PAMG_ERR_ARG;
  if (use_dir < 0) use_dir = h->un_use_dir;   // internal: re-assemble with the previous pairing rule
  h->un_use_dir = use_dir; h->un.with_stab = false;
  CK(cudaSetDevice(h->device));
  std::string e;
  long long nl = 0;
  int rc = implicit_assemble(h->un, dt, u_x, u_y, use_dir, h->nsm, h->stream, nl, e);
  h->launches += nl;
  if (rc) return fail(h, rc, e);
  return PAMG_OK;
}

int pamg_implicit_get_bsr(pamg_handle* h, double* val, int32_t* col) {
  if (!h || h->un.E < 1 || (!val && !col)) return PAMG_ERR_ARG;
  if (!h->un.assembled) return fail(h, PAMG_ERR_STATE, "pamg_implicit_assemble has not been called");
  CK(cudaSetDevice(h->device));
  const size_t E = (size_t)h->un.E;
  if (val) CK(cudaMemcpyAsync(val, h->un.bsr_val, E * 36 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  if (col) CK(cudaMemcpyAsync(col, h->un.bsr_col, E * 4 * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return PAMG_OK;
}

int pamg_implicit_apply(pamg_handle* h, const double* x, double* y) {
  if (!h || h->un.E < 1 || !x || !y) return PAMG_ERR_ARG;
  if (!h->un.assembled) return fail(h, PAMG_ERR_STATE, "pamg_implicit_assemble has not been called");
  CK(cudaSetDevice(h->device));
  const size_t nb = (size_t)h->un.E * 3 * sizeof(double);
  double* W = h->un.work;
  CK(cudaMemcpyAsync(W, x, nb, cudaMemcpyHostToDevice, h->stream));
  const int grid = std::max(1, std::min((h->un.E + TPB - 1) / TPB, h->nsm * 8));
  k_bsr_spmv<<<grid, TPB, 0, h->stream>>>(h->un.bsr_val, h->un.bsr_col, W, nullptr, W + (size_t)h->un.E * 3, h->un.E, 0);
  h->launches++;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(y, W + (size_t)h->un.E * 3, nb, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return PAMG_OK;
}

int pamg_implicit_set_stab(pamg_handle* h, int with_stab) {
  if (!h || h->un.E < 1) return PAMG_ERR_ARG;
  if (!h->un.assembled) return fail(h, PAMG_ERR_STATE, "pamg_implicit_assemble has not been called");
  CK(cudaSetDevice(h->device));
  if (h->un.with_stab && !with_stab)   // back to the plain operator: restore the diagonal blocks and their inverses
    return pamg_implicit_assemble(h, h->un.dt, h->un.ux, h->un.uy, -1);
  h->un.with_stab = with_stab != 0;
  return PAMG_OK;
}

int pamg_unstr_stab(pamg_handle* h, const double* told, double dt, double u_x, double u_y, double* diff_coe, double* stab) {
  if (!h || h->un.E < 1 || !told || !(dt > 0.0) || (!diff_coe && !stab)) return PAMG_ERR_ARG;
  CK(cudaSetDevice(h->device));
  const size_t E = (size_t)h->un.E;
  double *d_old = nullptr, *d_out = nullptr;
  CK(cudaMalloc(&d_old, E * 3 * sizeof(double)));
  if (cudaMalloc(&d_out, E * 12 * sizeof(double)) != cudaSuccess) { cudaFree(d_old); return fail(h, PAMG_ERR_CUDA, "cudaMalloc"); }
  cudaMemcpyAsync(d_old, told, E * 3 * sizeof(double), cudaMemcpyHostToDevice, h->stream);
  StabArgs sa;
  sa.X = h->un.X; sa.tnew = h->un.T[h->un.cur]; sa.told = d_old; sa.diff_coe = d_out + E * 9; sa.stab = d_out; sa.diag0 = nullptr;
  sa.val = nullptr; sa.dinv = nullptr; sa.dt = dt; sa.ux = u_x; sa.uy = u_y; sa.E = h->un.E; sa.mode = 0;
  const int grid = std::max(1, std::min((h->un.E + TPB - 1) / TPB, h->nsm * 8));
  k_unstr_stab<<<grid, TPB, 0, h->stream>>>(sa);
  h->launches++;
  cudaError_t e1 = cudaGetLastError();
  if (stab) cudaMemcpyAsync(stab, d_out, E * 9 * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
  if (diff_coe) cudaMemcpyAsync(diff_coe, d_out + E * 9, E * 3 * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
  cudaError_t e2 = cudaStreamSynchronize(h->stream);
  cudaFree(d_old); cudaFree(d_out);
  if (e1 != cudaSuccess || e2 != cudaSuccess) return fail(h, PAMG_ERR_CUDA, cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
  return PAMG_OK;
}

int pamg_implicit_step(pamg_handle* h, int ntime, int nits, double tol, int max_iters, int* iters_total, double* relres) {
  if (!h || h->un.E < 1 || ntime < 0 || nits < 1 || !(tol > 0.0) || max_iters < 1) return PAMG_ERR_ARG;
  if (!h->un.assembled) return fail(h, PAMG_ERR_STATE, "pamg_implicit_assemble has not been called");
  CK(cudaSetDevice(h->device));
  std::string e;
  long long nl = 0;
  int rc = implicit_step(h->un, ntime, nits, tol, max_iters, iters_total, relres, h->nsm, h->stream, nl, e);
  h->launches += nl;
  if (rc) return fail(h, rc, e);
  return PAMG_OK;
}

// ---- trans_rec front-end (transport_rect.F90:7) -------------------------------------------------------------
int pamg_trans_rec(pamg_handle* h, double CFL, int no_ele_row, int no_ele_col, double x_length, double y_length, double u_x,
                   double u_y, double time, int nits, int njac_its, int direct_solver, int volume_term, double* x_all,
                   double* tnew, int* ntime) {
  if (!h || !tnew || no_ele_row < 1 || no_ele_col < 1 || !(CFL > 0.0) || !(x_length > 0.0) || !(y_length > 0.0) || time < 0.0 ||
      nits < 1 || njac_its < 0)
    return PAMG_ERR_ARG;
  CK(cudaSetDevice(h->device));
  std::string e;
  long long nl = 0;
  int rc = rect_run(CFL, no_ele_row, no_ele_col, x_length, y_length, u_x, u_y, time, nits, njac_its, direct_solver, volume_term,
                    x_all, tnew, ntime, h->nsm, h->stream, nl, e);
  h->launches += nl;
  if (rc) return fail(h, rc, e);
  return PAMG_OK;
}

int pamg_apply_local_minv(pamg_handle* h, int n, int batch, const double* M, const double* rhs, double* x,
                          double* Minv, int32_t* status) {
  if (!h || batch < 1 || !M || !(n == 3 || n == 4 || n == 6)) return PAMG_ERR_ARG;
  if (!rhs != !x) return PAMG_ERR_ARG;
  CK(cudaSetDevice(h->device));
  std::string e;
  long long nl = 0;
  int rc = local_minv(n, batch, M, rhs, x, Minv, status, h->nsm, h->stream, nl, e);
  h->launches += nl;
  if (rc) return fail(h, rc, e);
  return PAMG_OK;
}

// ---- output (get_vtu, get_vtk_files.F90:10-140; get_error transport_tri_semi.F90:531-540) ---------------------
namespace {
int output_fields(pamg_handle* h, double* x_all, double* analytical, double* error) {
  LevelDev& L = h->lev[0];
  const size_t n = (size_t)L.nelem;
  double* d = nullptr;
  CK(cudaMalloc(&d, n * 12 * sizeof(double)));
  OutArgs a;
  a.xg = h->xg; a.T = tnew_ptr(L); a.x_all = x_all ? d : nullptr; a.analytical = analytical ? d + n * 6 : nullptr;
  a.error = error ? d + n * 9 : nullptr; a.nelem = L.nelem; a.s = L.s;
  k_output_fields<<<grid_for(h, L.nelem), TPB, 0, h->stream>>>(a);
  h->launches++;
  cudaError_t e = cudaGetLastError();
  if (x_all) cudaMemcpyAsync(x_all, d, n * 6 * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
  if (analytical) cudaMemcpyAsync(analytical, d + n * 6, n * 3 * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
  if (error) cudaMemcpyAsync(error, d + n * 9, n * 3 * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
  cudaError_t e2 = cudaStreamSynchronize(h->stream);
  cudaFree(d);
  if (e != cudaSuccess || e2 != cudaSuccess) return fail(h, PAMG_ERR_CUDA, cudaGetErrorString(e != cudaSuccess ? e : e2));
  return PAMG_OK;
}
}  // namespace

int pamg_output_fields(pamg_handle* h, double* x_all, double* analytical, double* error) {
  if (!h || h->lev.empty() || (!x_all && !analytical && !error)) return PAMG_ERR_ARG;
  CK(cudaSetDevice(h->device));
  return output_fields(h, x_all, analytical, error);
}

int pamg_write_vtu(pamg_handle* h, const char* path, const char* solve_for, int binary) {
  if (!h || h->lev.empty() || !path || !solve_for) return PAMG_ERR_ARG;
  CK(cudaSetDevice(h->device));
  LevelDev& L = h->lev[0];
  const size_t n = (size_t)L.nelem;
  std::vector<double> X(n * 6), T(n * 3), An(n * 3), Er(n * 3);
  int rc = output_fields(h, X.data(), An.data(), Er.data());
  if (rc) return rc;
  CK(cudaMemcpy(T.data(), tnew_ptr(L), n * 3 * sizeof(double), cudaMemcpyDeviceToHost));
  FILE* f = fopen(path, binary ? "wb" : "w");
  if (!f) return fail(h, PAMG_ERR_IO, std::string("cannot open ") + path);
  const unsigned long long npts = 3ull * n;
  fprintf(f, "<VTKFile type=\"UnstructuredGrid\" version=\"0.1\" byte_order=\"LittleEndian\"%s>\n", binary ? " header_type=\"UInt64\"" : "");
  fprintf(f, "  <UnstructuredGrid>\n    <Piece NumberOfPoints=\"%llu\" NumberOfCells=\"%llu\">\n", npts, (unsigned long long)n);
  fprintf(f, "      <PointData Scalars=\"scalars\">\n");
  if (!binary) {
    // the reference's layout: one value per line, F12.10 for the unknown, F10.7 for error / analytical, F10.3 coordinates
    const double* arr[3] = {T.data(), Er.data(), An.data()};
    const char* names[3] = {solve_for, "error", "analytical"};
    for (int k = 0; k < 3; ++k) {
      fprintf(f, "        <DataArray type=\"Float32\" Name=\"%s\" Format=\"ascii\">\n", names[k]);
      for (size_t i = 0; i < n * 3; ++i) fprintf(f, k == 0 ? "          %12.10f  \n" : "          %10.7f  \n", arr[k][i]);
      fprintf(f, "        </DataArray>\n");
    }
    fprintf(f, "      </PointData>\n      <Points>\n        <DataArray type=\"Float32\" NumberOfComponents=\"3\" Format=\"ascii\">\n");
    for (size_t i = 0; i < n * 3; ++i) fprintf(f, "          %.3f %.3f %.3f   \n", X[2 * i], X[2 * i + 1], 0.0);
    fprintf(f, "        </DataArray>\n      </Points>\n      <Cells>\n        <DataArray type=\"Int32\" Name=\"connectivity\" Format=\"ascii\">\n");
    for (size_t e = 0; e < n; ++e) fprintf(f, "          %zu %zu %zu\n", 3 * e, 3 * e + 1, 3 * e + 2);
    fprintf(f, "        </DataArray>\n        <DataArray type=\"Int32\" Name=\"offsets\" Format=\"ascii\">\n          ");
    for (size_t e = 1; e <= n; ++e) fprintf(f, e == 1 ? "%zu" : "  %zu", 3 * e);
    fprintf(f, "\n        </DataArray>\n        <DataArray type=\"Int32\" Name=\"types\" Format=\"ascii\">\n          ");
    for (size_t e = 0; e < n; ++e) fprintf(f, e == 0 ? "5" : " 5");                  // cell_type 5 = VTK_TRIANGLE (main.F90)
    fprintf(f, "\n        </DataArray>\n      </Cells>\n    </Piece>\n  </UnstructuredGrid>\n</VTKFile>\n");
  } else {
    // raw appended data, Float64 fields: [UInt64 byte count][payload] per array, offsets relative to the '_' marker
    unsigned long long off = 0;
    const char* names[3] = {solve_for, "error", "analytical"};
    for (int k = 0; k < 3; ++k) {
      fprintf(f, "        <DataArray type=\"Float64\" Name=\"%s\" format=\"appended\" offset=\"%llu\"/>\n", names[k], off);
      off += 8 + npts * 8;
    }
    fprintf(f, "      </PointData>\n      <Points>\n        <DataArray type=\"Float64\" NumberOfComponents=\"3\" format=\"appended\" offset=\"%llu\"/>\n      </Points>\n", off);
    off += 8 + npts * 24;
    fprintf(f, "      <Cells>\n        <DataArray type=\"Int64\" Name=\"connectivity\" format=\"appended\" offset=\"%llu\"/>\n", off);
    off += 8 + npts * 8;
    fprintf(f, "        <DataArray type=\"Int64\" Name=\"offsets\" format=\"appended\" offset=\"%llu\"/>\n", off);
    off += 8 + n * 8;
    fprintf(f, "        <DataArray type=\"UInt8\" Name=\"types\" format=\"appended\" offset=\"%llu\"/>\n", off);
    fprintf(f, "      </Cells>\n    </Piece>\n  </UnstructuredGrid>\n  <AppendedData encoding=\"raw\">\n   _");
    auto block = [&](const void* p, unsigned long long bytes) { fwrite(&bytes, 8, 1, f); fwrite(p, 1, bytes, f); };
    block(T.data(), npts * 8); block(Er.data(), npts * 8); block(An.data(), npts * 8);
    std::vector<double> P3(npts * 3);
    for (size_t i = 0; i < npts; ++i) { P3[3 * i] = X[2 * i]; P3[3 * i + 1] = X[2 * i + 1]; P3[3 * i + 2] = 0.0; }
    block(P3.data(), npts * 24);
    std::vector<long long> conn(npts), offs(n);
    for (size_t i = 0; i < npts; ++i) conn[i] = (long long)i;
    for (size_t e = 0; e < n; ++e) offs[e] = 3ll * (long long)(e + 1);
    block(conn.data(), npts * 8); block(offs.data(), n * 8);
    std::vector<unsigned char> types(n, 5);
    block(types.data(), n);
    fprintf(f, "\n  </AppendedData>\n</VTKFile>\n");
  }
  if (fclose(f) != 0) return fail(h, PAMG_ERR_IO, std::string("write failed: ") + path);
  return PAMG_OK;
}

// ---- timing helpers --------------------------------------------------------------------------------
int pamg_sync(pamg_handle* h) {
  if (!h) return PAMG_ERR_ARG;
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));
  if (h->up_stream) { CK(cudaStreamSynchronize(h->up_stream)); CK(cudaStreamSynchronize(h->down_stream)); h->pipe_calls = 0; }
  if (h->p2p_ready) {      // a halo exchange that gave up waiting for a peer raised the error word instead of hanging
    unsigned long long err = 0;
    CK(cudaMemcpy(&err, h->p2p_sync + P2P_ERR, sizeof(err), cudaMemcpyDeviceToHost));
    if (err) return fail(h, PAMG_ERR_CUDA, "halo exchange timed out waiting for a peer GPU");
  }
  return PAMG_OK;
}

int pamg_event_record(pamg_handle* h, int slot) {
  if (!h || slot < 0 || slot >= 16) return PAMG_ERR_ARG;
  CK(cudaEventRecord(h->ev[slot], h->stream));
  return PAMG_OK;
}

int pamg_event_elapsed_ms(pamg_handle* h, int a, int b, float* ms) {
  if (!h || !ms || a < 0 || a >= 16 || b < 0 || b >= 16) return PAMG_ERR_ARG;
  CK(cudaEventSynchronize(h->ev[b]));
  CK(cudaEventElapsedTime(ms, h->ev[a], h->ev[b]));
  return PAMG_OK;
}

int pamg_launch_count(const pamg_handle* h, int64_t* n) {
  if (!h || !n) return PAMG_ERR_ARG;
  *n = h->launches;
  return PAMG_OK;
}

int pamg_profile(pamg_handle* h, int on) {
  if (!h) return PAMG_ERR_ARG;
  CK(cudaSetDevice(h->device));
  if (on && h->pev.empty()) {
    h->pev.resize(2048);
    for (auto& e : h->pev) CK(cudaEventCreate(&e));
  }
  h->profiling = on != 0;
  h->pev_used = 0;
  return PAMG_OK;
}

int pamg_profile_read(pamg_handle* h, double* total_ms, int* launches) {
  if (!h || !total_ms || !launches) return PAMG_ERR_ARG;
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));
  double tot = 0.0;
  for (int i = 0; i + 1 < h->pev_used; i += 2) {
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, h->pev[i], h->pev[i + 1]));
    tot += ms;
  }
  *total_ms = tot;
  *launches = h->pev_used / 2;
  return PAMG_OK;
}

int pamg_host_alloc(void** p, int64_t bytes) {
  if (!p || bytes <= 0) return PAMG_ERR_ARG;
  return cudaMallocHost(p, (size_t)bytes) == cudaSuccess ? PAMG_OK : PAMG_ERR_CUDA;
}

int pamg_host_free(void* p) { return cudaFreeHost(p) == cudaSuccess ? PAMG_OK : PAMG_ERR_CUDA; }

int pamg_flush_l2(pamg_handle* h) {
  if (!h) return PAMG_ERR_ARG;
  CK(cudaSetDevice(h->device));
  if (!h->scratch) {
    h->scratch_bytes = (size_t)256 << 20;
    CK(cudaMalloc(&h->scratch, h->scratch_bytes));
  }
  k_fill<<<grid_for(h, (long long)(h->scratch_bytes / 8)), TPB, 0, h->stream>>>(h->scratch, (long long)(h->scratch_bytes / 8), 1.0);
  CK(cudaGetLastError());
  return PAMG_OK;
}

}  // extern "C"

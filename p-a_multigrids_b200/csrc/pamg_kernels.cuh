// pamg_kernels.cuh -- sm_100a kernels of the semi-structured P1 DG hot path.
//
// All kernels are HBM-bandwidth bound fp64 streaming stencils (24 B per DOF update for a
// Jacobi sweep / residual evaluation); there is no dense contraction, so no tensor-core work.  Design:
//   * one thread per child triangle, 256 threads per CTA, grid-stride over tiles with the grid sized as a
//     multiple of the SM count;
//   * children of a parent are numbered row by row (Msh2Tri.F90:42-58); rows r and 2^s+1-r together hold
//     exactly 2^(s+1) children, so a flat index maps to (row, position) with shifts only - no integer sqrt
//     and no ragged tiles.  Consecutive threads touch consecutive 24-byte records => coalesced streams;
//   * neighbour values are re-reads of the same stream (adjacent rows) served by L1/L2, halo strips and the
//     20 per-parent coefficients come through the read-only path;
//   * geometry is per parent (ShapFun.F90:1661-1684,1737-1783) and folded on the host into closed-form
//     coefficients, so the kernels carry ~100 flops per child.
#pragma once
#include <cstdint>

namespace pamg {

constexpr int NPC = 20;  // doubles per parent per level
// per-parent per-level coefficient slots
enum { PC_CM = 0, PC_K11 = 1, PC_K12, PC_K13, PC_K22, PC_K23, PC_K33, PC_ADV = 7, PC_FL = 10, PC_PENI = 13, PC_PENX = 16 };

constexpr int TPB = 256;

enum { MODE_JACOBI = 0, MODE_RESID = 1, MODE_GS = 2, MODE_RICH = 3 };

struct ElemArgs {
  const double* Tin;      // field the sweep reads (may alias Tout for the in-place coloured pass)
  double* Tout;           // Jacobi/GS/Richardson: new iterate; residual: RES
  const double* rhs;
  const double* ovl;      // halo strips [strip][S][3]
  const double* pc;       // [U][NPC]
  const int32_t* strip_of;  // [U*3]
  const int32_t* hmap;      // [U*3]
  double* partial;        // residual: [nblocks][3] = sum r^2, max |r|, max r
  double omega;
  double rsign;
  long long nelem;        // U * C
  int s;                  // split of this level
  int colour;             // GS: 0 = down children, 1 = up children
};

// flat child index t in [0, 4^s) -> row r (1-based), position ipos (1-based), element id ele (1-based)
__device__ __forceinline__ void child_from_flat(int t, int s, int& r, int& ipos, int& ele, int& len) {
  const int b = 2 << s;            // 2^(s+1)
  const int S = 1 << s;
  const int p = t >> (s + 1);
  const int q = t & (b - 1);
  const int lenA = b - 1 - 2 * p;  // length of row p+1
  if (q < lenA) { r = p + 1; ipos = q + 1; len = lenA; }
  else { r = S - p; ipos = q - lenA + 1; len = 2 * p + 1; }
  ele = 1 + (r - 1) * (b + 1 - r) + ipos - 1;
}

template <int MODE, bool FACE>
__global__ void __launch_bounds__(TPB) k_element(ElemArgs a) {
  const int s = a.s;
  const int twos = 2 * s;
  const int b = 2 << s;
  const int S = 1 << s;
  const long long Cmask = (1ll << twos) - 1;
  double acc_sum = 0.0, acc_abs = 0.0, acc_max = 0.0;

  for (long long gid = (long long)blockIdx.x * TPB + threadIdx.x; gid < a.nelem; gid += (long long)gridDim.x * TPB) {
    const int u = (int)(gid >> twos);
    const int t = (int)(gid & Cmask);
    int r, ipos, ele, len;
    child_from_flat(t, s, r, ipos, ele, len);
    const bool up = ipos & 1;
    if (MODE == MODE_GS && (int)up != a.colour) continue;

    const long long pbase = ((long long)u << twos);       // first child of the parent
    const long long base = (pbase + ele - 1) * 3;
    // read-only (non-coherent) path unless the pass updates the field in place (coloured GS)
    auto ldT = [&](long long i) -> double { return MODE == MODE_GS ? a.Tin[i] : __ldg(a.Tin + i); };
    const double T1 = ldT(base), T2 = ldT(base + 1), T3 = ldT(base + 2);
    const double* __restrict__ pc = a.pc + (size_t)u * NPC;
    const double sg = up ? 1.0 : -1.0;

    // ---- volume terms: (1/dt) M T - S T + K T   (get_A_x, transport_tri_semi.F90:412-448, theta = 1)
    const double cm = __ldg(pc + PC_CM);
    const double sumT = T1 + T2 + T3;
    const double k11 = __ldg(pc + PC_K11), k12 = __ldg(pc + PC_K12), k13 = __ldg(pc + PC_K13);
    const double k22 = __ldg(pc + PC_K22), k23 = __ldg(pc + PC_K23), k33 = __ldg(pc + PC_K33);
    const double adv = sg * sumT;
    double mass1 = cm * (T1 + sumT), mass2 = cm * (T2 + sumT), mass3 = cm * (T3 + sumT);
    double st1 = __ldg(pc + PC_ADV + 0) * adv, st2 = __ldg(pc + PC_ADV + 1) * adv, st3 = __ldg(pc + PC_ADV + 2) * adv;
    double ax1 = mass1 - st1 + (k11 * T1 + k12 * T2 + k13 * T3);
    double ax2 = mass2 - st2 + (k12 * T1 + k22 * T2 + k23 * T3);
    double ax3 = mass3 - st3 + (k13 * T1 + k23 * T2 + k33 * T3);
    // get_diagonal (:481-486): ml/dt + K_ii (+ penalty diagonal below); ml = A/3 = 4 * A/12
    double d1 = 4.0 * cm + k11, d2 = 4.0 * cm + k22, d3 = 4.0 * cm + k33;
    double fx1 = 0.0, fx2 = 0.0, fx3 = 0.0;  // upwind flux, kept apart for Richardson (:516)

    if (FACE) {
      // neighbour values at the nodes coincident with my face nodes (a,b) of child faces
      //   f1:(1,3)  f2:(3,2)  f3:(2,1)      (transport_tri_semi.F90:142-147)
      double n1a, n1b, n2a, n2b, n3a, n3b;
      double pen1, pen2, pen3;
      if (!up) {
        // down child: f1 -> child above, f2 -> ele+1, f3 -> ele-1 (splitting.F90:766); never on a parent face
        const long long o1 = (pbase + (ele + b - 2 * r) - 1) * 3;
        const long long o2 = base + 3, o3 = base - 3;
        n1a = ldT(o1 + 2); n1b = ldT(o1 + 0);   // my node 1 <-> its node 3, my node 3 <-> its node 1
        n2a = ldT(o2 + 1); n2b = ldT(o2 + 2);   // my 3 <-> its 2, my 2 <-> its 3
        n3a = ldT(o3 + 0); n3b = ldT(o3 + 1);   // my 2 <-> its 1, my 1 <-> its 2
        pen1 = __ldg(pc + PC_PENI + 0); pen2 = __ldg(pc + PC_PENI + 1); pen3 = __ldg(pc + PC_PENI + 2);
      } else {
        const size_t S3 = (size_t)3 * S;
        if (r > 1) {
          const long long o1 = (pbase + (ele - b - 2 + 2 * r) - 1) * 3;
          n1a = ldT(o1 + 2); n1b = ldT(o1 + 0);
          pen1 = __ldg(pc + PC_PENI + 0);
        } else {  // parent face 1, slot ipos/2+1 (:629-631)
          const int hm = __ldg(a.hmap + u * 3 + 0);
          const double* e = a.ovl + (size_t)__ldg(a.strip_of + u * 3 + 0) * S3 + (size_t)(ipos >> 1) * 3;
          n1a = __ldg(e + (hm & 3)); n1b = __ldg(e + (hm >> 2));
          pen1 = __ldg(pc + PC_PENX + 0);
        }
        if (ipos > 1) {
          n2a = ldT(base - 3 + 1); n2b = ldT(base - 3 + 2);
          pen2 = __ldg(pc + PC_PENI + 1);
        } else {  // parent face 3, slot irow (:632-634)
          const int hm = __ldg(a.hmap + u * 3 + 2);
          const double* e = a.ovl + (size_t)__ldg(a.strip_of + u * 3 + 2) * S3 + (size_t)(r - 1) * 3;
          n2a = __ldg(e + (hm & 3)); n2b = __ldg(e + (hm >> 2));
          pen2 = __ldg(pc + PC_PENX + 1);
        }
        if (ipos < len) {
          n3a = ldT(base + 3 + 0); n3b = ldT(base + 3 + 1);
          pen3 = __ldg(pc + PC_PENI + 2);
        } else {  // parent face 2, slot irow (:635-637)
          const int hm = __ldg(a.hmap + u * 3 + 1);
          const double* e = a.ovl + (size_t)__ldg(a.strip_of + u * 3 + 1) * S3 + (size_t)(r - 1) * 3;
          n3a = __ldg(e + (hm & 3)); n3b = __ldg(e + (hm >> 2));
          pen3 = __ldg(pc + PC_PENX + 2);
        }
      }
      // penalty diffusion (k/dx) int sn_i (T - T2)  (matrices.F90:113-115, get_diff_surf_stencl :468-477):
      // face mass (L/6)[[2,1],[1,2]] folded into pen = k (L/2) / (3 dx)
      {
        const double da = T1 - n1a, db = T3 - n1b;   // face 1: a = node 1, b = node 3
        ax1 += pen1 * (2.0 * da + db); ax3 += pen1 * (da + 2.0 * db);
        d1 += 2.0 * pen1; d3 += 2.0 * pen1;
      }
      {
        const double da = T3 - n2a, db = T2 - n2b;   // face 2: a = node 3, b = node 2
        ax3 += pen2 * (2.0 * da + db); ax2 += pen2 * (da + 2.0 * db);
        d3 += 2.0 * pen2; d2 += 2.0 * pen2;
      }
      {
        const double da = T2 - n3a, db = T1 - n3b;   // face 3: a = node 2, b = node 1
        ax2 += pen3 * (2.0 * da + db); ax1 += pen3 * (da + 2.0 * db);
        d2 += 2.0 * pen3; d1 += 2.0 * pen3;
      }
      // upwind flux: income = 1 when n.u < 0 (transport_tri_unstr.F90:729-738)
      {
        const double fl = sg * __ldg(pc + PC_FL + 0);
        const bool in = fl < 0.0;
        const double wa = in ? n1a : T1, wb = in ? n1b : T3;
        fx1 += fl * (2.0 * wa + wb); fx3 += fl * (wa + 2.0 * wb);
      }
      {
        const double fl = sg * __ldg(pc + PC_FL + 1);
        const bool in = fl < 0.0;
        const double wa = in ? n2a : T3, wb = in ? n2b : T2;
        fx3 += fl * (2.0 * wa + wb); fx2 += fl * (wa + 2.0 * wb);
      }
      {
        const double fl = sg * __ldg(pc + PC_FL + 2);
        const bool in = fl < 0.0;
        const double wa = in ? n3a : T2, wb = in ? n3b : T1;
        fx2 += fl * (2.0 * wa + wb); fx1 += fl * (wa + 2.0 * wb);
      }
      ax1 += fx1; ax2 += fx2; ax3 += fx3;
    }

    const double b1 = __ldg(a.rhs + base), b2 = __ldg(a.rhs + base + 1), b3 = __ldg(a.rhs + base + 2);
    if (MODE == MODE_RESID) {
      const double r1 = a.rsign * (ax1 - b1), r2 = a.rsign * (ax2 - b2), r3 = a.rsign * (ax3 - b3);  // :869
      a.Tout[base] = r1; a.Tout[base + 1] = r2; a.Tout[base + 2] = r3;
      acc_sum += r1 * r1 + r2 * r2 + r3 * r3;
      acc_abs = fmax(acc_abs, fmax(fabs(r1), fmax(fabs(r2), fabs(r3))));
      acc_max = fmax(acc_max, fmax(r1, fmax(r2, r3)));
    } else if (MODE == MODE_RICH) {
      // solve_Richardson (:511-518): omega * (b - (mass - stiff + flux))
      a.Tout[base] = T1 + a.omega * (b1 - (mass1 - st1 + fx1));
      a.Tout[base + 1] = T2 + a.omega * (b2 - (mass2 - st2 + fx2));
      a.Tout[base + 2] = T3 + a.omega * (b3 - (mass3 - st3 + fx3));
    } else {
      // solve_Jacobi (:491-497) / solve_Gauss_Seidel (:501-507)
      a.Tout[base] = T1 + a.omega / d1 * (b1 - ax1);
      a.Tout[base + 1] = T2 + a.omega / d2 * (b2 - ax2);
      a.Tout[base + 2] = T3 + a.omega / d3 * (b3 - ax3);
    }
  }

  if (MODE == MODE_RESID) {
    // warp-shuffle reduction, then one partial per CTA (deterministic two-stage reduction)
    for (int o = 16; o > 0; o >>= 1) {
      acc_sum += __shfl_xor_sync(0xffffffffu, acc_sum, o);
      acc_abs = fmax(acc_abs, __shfl_xor_sync(0xffffffffu, acc_abs, o));
      acc_max = fmax(acc_max, __shfl_xor_sync(0xffffffffu, acc_max, o));
    }
    __shared__ double sh[3][TPB / 32];
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { sh[0][w] = acc_sum; sh[1][w] = acc_abs; sh[2][w] = acc_max; }
    __syncthreads();
    if (threadIdx.x == 0) {
      double s0 = 0, s1 = 0, s2 = 0;
      for (int i = 0; i < TPB / 32; ++i) { s0 += sh[0][i]; s1 = fmax(s1, sh[1][i]); s2 = fmax(s2, sh[2][i]); }
      a.partial[(size_t)blockIdx.x * 3 + 0] = s0;
      a.partial[(size_t)blockIdx.x * 3 + 1] = s1;
      a.partial[(size_t)blockIdx.x * 3 + 2] = s2;
    }
  }
}

// second stage of the norm reduction: one CTA
__global__ void __launch_bounds__(1024) k_reduce_partials(const double* partial, int n, double* out3) {
  double s0 = 0, s1 = 0, s2 = 0;
  for (int i = threadIdx.x; i < n; i += 1024) {
    s0 += partial[(size_t)i * 3]; s1 = fmax(s1, partial[(size_t)i * 3 + 1]); s2 = fmax(s2, partial[(size_t)i * 3 + 2]);
  }
  for (int o = 16; o > 0; o >>= 1) {
    s0 += __shfl_xor_sync(0xffffffffu, s0, o);
    s1 = fmax(s1, __shfl_xor_sync(0xffffffffu, s1, o));
    s2 = fmax(s2, __shfl_xor_sync(0xffffffffu, s2, o));
  }
  __shared__ double sh[3][32];
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { sh[0][w] = s0; sh[1][w] = s1; sh[2][w] = s2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    s0 = 0; s1 = 0; s2 = 0;
    for (int i = 0; i < 32; ++i) { s0 += sh[0][i]; s1 = fmax(s1, sh[1][i]); s2 = fmax(s2, sh[2][i]); }
    out3[0] = s0; out3[1] = s1; out3[2] = s2;
  }
}

// ------------------------------------------------------------------------------------------------
// update_overlaps (splitting.F90:1210-1397): one thread per (parent, side, position)
struct HaloArgs {
  const double* tnew; const double* told;
  double* ovl; double* ovl_old;           // strip space: local strips then send slots
  const double* xg;                        // [U][6] X3, X1-X3, X2-X3
  const int32_t* dst_strip; const int32_t* rev; const int32_t* strip_of;
  double bc_scale;
  int U, s, with_old;
};

__device__ __forceinline__ void child_nodes(const double* __restrict__ xg, int s, int r, int ipos, double x[3][2]) {
  // get_splitting, Msh2Tri.F90:79-106 (same operation order as the reference: divide first)
  const double inv = (double)(1 << s);
  for (int d = 0; d < 2; ++d) {
    const double o = __ldg(xg + d), v1 = __ldg(xg + 2 + d) / inv, v2 = __ldg(xg + 4 + d) / inv;
    if (ipos & 1) {
      x[2][d] = o + (r - 1) * v2 + (ipos / 2) * v1;
      x[1][d] = o + r * v2 + (ipos / 2) * v1;
      x[0][d] = o + (r - 1) * v2 + v1 * (ipos / 2 + 1);
    } else {
      x[0][d] = o + r * v2 + v1 * (ipos / 2 - 1);
      x[1][d] = o + (r - 1) * v2 + v1 * (ipos / 2);
      x[2][d] = o + r * v2 + v1 * (ipos / 2);
    }
  }
}

__global__ void __launch_bounds__(TPB) k_halo(HaloArgs a) {
  const int S = 1 << a.s, b = 2 << a.s;
  const long long n = (long long)a.U * 3 * S;
  for (long long tid = (long long)blockIdx.x * TPB + threadIdx.x; tid < n; tid += (long long)gridDim.x * TPB) {
    const int i = (int)(tid & (S - 1));           // position - 1
    const int lf = (int)(tid >> a.s);             // u*3 + mf
    const int u = lf / 3, mf = lf - 3 * u;
    const int pos = i + 1;
    int r, ipos;
    if (mf == 0) { r = 1; ipos = 2 * pos - 1; }                  // surf_ele(:,1): odd children of row 1
    else if (mf == 2) { r = pos; ipos = 1; }                     // surf_ele(:,3): first child of each row
    else { r = pos; ipos = b + 1 - 2 * pos; }                    // surf_ele(:,2): last child of each row
    const int ele = 1 + (r - 1) * (b + 1 - r) + ipos - 1;
    const int dst = __ldg(a.dst_strip + lf);
    const size_t S3 = (size_t)3 * S;
    if (dst < 0) {
      // Dirichlet data sin(x+y) at the two face nodes (:1246-1252,1287-1293,1344-1350)
      double x[3][2];
      child_nodes(a.xg + (size_t)u * 6, a.s, r, ipos, x);
      const int na = (mf == 2) ? 1 : 0, nb = (mf == 0) ? 2 : (mf == 1 ? 1 : 2);
      const double ta = a.bc_scale * sin(x[na][0] + x[na][1]);
      const double tb = a.bc_scale * sin(x[nb][0] + x[nb][1]);
      const size_t o = (size_t)__ldg(a.strip_of + lf) * S3 + (size_t)i * 3;
      a.ovl[o + na] = ta; a.ovl[o + nb] = tb;
      if (a.with_old) { a.ovl_old[o + na] = ta; a.ovl_old[o + nb] = tb; }
    } else {
      const int slot = __ldg(a.rev + lf) ? (S - pos) : (pos - 1);
      const size_t o = (size_t)dst * S3 + (size_t)slot * 3;
      const size_t src = (((size_t)u << (2 * a.s)) + ele - 1) * 3;
      a.ovl[o] = a.tnew[src]; a.ovl[o + 1] = a.tnew[src + 1]; a.ovl[o + 2] = a.tnew[src + 2];
      if (a.with_old) { a.ovl_old[o] = a.told[src]; a.ovl_old[o + 1] = a.told[src + 1]; a.ovl_old[o + 2] = a.told[src + 2]; }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// level-1 RHS: source reset (:593) + get_RHS (:452-464), theta = 1
struct RhsArgs {
  const double* told; double* rhs; const double* pc; const double* xg;
  double dt, source_coef;
  long long nelem;
  int s, literal_source;
};

__global__ void __launch_bounds__(TPB) k_build_rhs(RhsArgs a) {
  const int s = a.s, twos = 2 * s;
  const long long Cmask = (1ll << twos) - 1;
  for (long long gid = (long long)blockIdx.x * TPB + threadIdx.x; gid < a.nelem; gid += (long long)gridDim.x * TPB) {
    const int u = (int)(gid >> twos);
    int r, ipos, ele, len;
    child_from_flat((int)(gid & Cmask), s, r, ipos, ele, len);
    const long long base = (((long long)u << twos) + ele - 1) * 3;
    double x[3][2];
    child_nodes(a.xg + (size_t)u * 6, s, r, ipos, x);
    const double cm = __ldg(a.pc + (size_t)u * NPC + PC_CM);
    const double m12 = cm * a.dt;   // A_child / 12
    double s1 = a.source_coef * sin(x[0][0] + x[0][1]);
    double s2 = a.source_coef * sin(x[1][0] + x[1][1]);
    double s3 = a.source_coef * sin(x[2][0] + x[2][1]);
    double q1, q2, q3;
    if (a.literal_source) {  // in-place: src(i) = M(i,:) . src with src already partly overwritten
      q1 = m12 * (2.0 * s1 + s2 + s3);
      q2 = m12 * (q1 + 2.0 * s2 + s3);
      q3 = m12 * (q1 + q2 + 2.0 * s3);
    } else {
      const double ss = s1 + s2 + s3;
      q1 = m12 * (s1 + ss); q2 = m12 * (s2 + ss); q3 = m12 * (s3 + ss);
    }
    const double o1 = a.told[base], o2 = a.told[base + 1], o3 = a.told[base + 2];
    const double so = o1 + o2 + o3;
    a.rhs[base] = cm * (o1 + so) + q1;
    a.rhs[base + 1] = cm * (o2 + so) + q2;
    a.rhs[base + 2] = cm * (o3 + so) + q3;
  }
}

// ------------------------------------------------------------------------------------------------
// restrictor (splitting.F90:10-32) / its INTENDED form (transpose of P1 interpolation).
// One thread per COARSE child; fine ids from the closed form of element_conversion (:105-139).
__device__ __forceinline__ void fine_children(int sc, int r, int ipos, int fin[4]) {
  const int bf = 4 << sc;  // 2^(sf+1), sf = sc+1
  auto start_f = [&](int rf) { return 1 + (rf - 1) * (bf + 1 - rf); };
  if (ipos & 1) {
    fin[0] = start_f(2 * r - 1) + 2 * ipos - 2;
    fin[1] = fin[0] + 1; fin[2] = fin[0] + 2;
    fin[3] = start_f(2 * r) + 2 * ipos - 2;
  } else {
    fin[2] = start_f(2 * r) + 2 * ipos - 3;
    fin[1] = fin[2] + 1; fin[0] = fin[2] + 2;
    fin[3] = start_f(2 * r - 1) + 2 * ipos - 1;
  }
}

struct XferArgs {
  const double* src; double* dst;
  long long ncoarse;   // U * 4^sc
  int sc;              // coarse split
  int mode;            // 0 literal, 1 intended
};

__global__ void __launch_bounds__(TPB) k_restrict(XferArgs a) {
  const int sc = a.sc, twoc = 2 * sc, twof = twoc + 2;
  const long long Cmask = (1ll << twoc) - 1;
  for (long long gid = (long long)blockIdx.x * TPB + threadIdx.x; gid < a.ncoarse; gid += (long long)gridDim.x * TPB) {
    const int u = (int)(gid >> twoc);
    int r, ipos, ele, len;
    child_from_flat((int)(gid & Cmask), sc, r, ipos, ele, len);
    int fin[4];
    fine_children(sc, r, ipos, fin);
    const long long fb = ((long long)u << twof);
    double R[4][3];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const long long o = (fb + fin[k] - 1) * 3;
      R[k][0] = __ldg(a.src + o); R[k][1] = __ldg(a.src + o + 1); R[k][2] = __ldg(a.src + o + 2);
    }
    const long long oc = (((long long)u << twoc) + ele - 1) * 3;
    if (a.mode == 0) {
      a.dst[oc + 0] = (R[2][0] + R[2][1] + R[2][2]) / 3.0;   // splitting.F90:26
      a.dst[oc + 1] = (R[3][0] + R[3][1] + R[3][2]) / 3.0;   // :27
      a.dst[oc + 2] = (R[0][0] + R[0][1] + R[0][2]) / 3.0;   // :28
    } else {
      // P^T with P from splitting.F90:59-88: fin1 holds coarse node 3, fin3 node 1, fin4 node 2,
      // fin2 is the inverted centre child with nodes mid(2,3), mid(1,3), mid(1,2)
      a.dst[oc + 0] = R[2][0] + 0.5 * (R[0][0] + R[1][1] + R[1][2] + R[2][1] + R[2][2] + R[3][0]);
      a.dst[oc + 1] = R[3][1] + 0.5 * (R[0][1] + R[1][0] + R[1][2] + R[2][1] + R[3][0] + R[3][2]);
      a.dst[oc + 2] = R[0][2] + 0.5 * (R[0][0] + R[0][1] + R[1][0] + R[1][1] + R[2][2] + R[3][2]);
    }
  }
}

// prolongator (splitting.F90:38-91) as written: one thread per coarse child (the mixing of totals and
// corrections for the centre child, SURVEY B-8, only touches that child's own four fine children)
__global__ void __launch_bounds__(TPB) k_prolong_literal(XferArgs a) {
  const int sc = a.sc, twoc = 2 * sc, twof = twoc + 2;
  const long long Cmask = (1ll << twoc) - 1;
  for (long long gid = (long long)blockIdx.x * TPB + threadIdx.x; gid < a.ncoarse; gid += (long long)gridDim.x * TPB) {
    const int u = (int)(gid >> twoc);
    int r, ipos, ele, len;
    child_from_flat((int)(gid & Cmask), sc, r, ipos, ele, len);
    int fin[4];
    fine_children(sc, r, ipos, fin);
    const long long oc = (((long long)u << twoc) + ele - 1) * 3;
    const double c1 = a.src[oc], c2 = a.src[oc + 1], c3 = a.src[oc + 2];
    const long long fb = ((long long)u << twof);
    double* f1 = a.dst + (fb + fin[0] - 1) * 3;
    double* f2 = a.dst + (fb + fin[1] - 1) * 3;
    double* f3 = a.dst + (fb + fin[2] - 1) * 3;
    double* f4 = a.dst + (fb + fin[3] - 1) * 3;
    f1[0] += 0.5 * c3 + 0.5 * c1; f1[1] += 0.5 * c2 + 0.5 * c3; f1[2] += c3;
    f2[0] += f1[1]; f2[1] += f1[0]; f2[2] += 0.5 * c1 + 0.5 * c2;
    f3[0] += c1; f3[1] += f2[2]; f3[2] += f2[1];
    f4[0] += f2[2]; f4[1] += c2; f4[2] += f2[0];
  }
}

// INTENDED prolongation: P1 interpolation of the coarse correction, one thread per FINE child so that
// the read-modify-write of the fine field is a coalesced stream.
__global__ void __launch_bounds__(TPB) k_prolong_p1(XferArgs a) {
  const int sf = a.sc + 1, twof = 2 * sf, twoc = twof - 2;
  const long long nfine = a.ncoarse * 4;
  const long long Fmask = (1ll << twof) - 1;
  const int bc = 2 << a.sc;
  for (long long gid = (long long)blockIdx.x * TPB + threadIdx.x; gid < nfine; gid += (long long)gridDim.x * TPB) {
    const int u = (int)(gid >> twof);
    int rf, ipf, elef, len;
    child_from_flat((int)(gid & Fmask), sf, rf, ipf, elef, len);
    const int rc = (rf + 1) >> 1;
    const int m = ipf & 3;
    int k, ipc;
    if (rf & 1) {
      if (m == 1) { k = 0; ipc = (ipf + 1) >> 1; }
      else if (m == 2) { k = 1; ipc = ipf >> 1; }
      else if (m == 3) { k = 2; ipc = (ipf - 1) >> 1; }
      else { k = 3; ipc = ipf >> 1; }
    } else {
      if (m == 1) { k = 3; ipc = (ipf + 1) >> 1; }
      else if (m == 2) { k = 2; ipc = (ipf + 2) >> 1; }
      else if (m == 3) { k = 1; ipc = (ipf + 1) >> 1; }
      else { k = 0; ipc = ipf >> 1; }
    }
    const int elec = 1 + (rc - 1) * (bc + 1 - rc) + ipc - 1;
    const long long oc = (((long long)u << twoc) + elec - 1) * 3;
    const double c1 = __ldg(a.src + oc), c2 = __ldg(a.src + oc + 1), c3 = __ldg(a.src + oc + 2);
    double e1, e2, e3;
    if (k == 0) { e1 = 0.5 * (c1 + c3); e2 = 0.5 * (c2 + c3); e3 = c3; }
    else if (k == 1) { e1 = 0.5 * (c2 + c3); e2 = 0.5 * (c1 + c3); e3 = 0.5 * (c1 + c2); }
    else if (k == 2) { e1 = c1; e2 = 0.5 * (c1 + c2); e3 = 0.5 * (c1 + c3); }
    else { e1 = 0.5 * (c1 + c2); e2 = c2; e3 = 0.5 * (c2 + c3); }
    const long long of = (((long long)u << twof) + elef - 1) * 3;
    a.dst[of] += e1; a.dst[of + 1] += e2; a.dst[of + 2] += e3;
  }
}

__global__ void __launch_bounds__(TPB) k_fill(double* p, long long n, double v) {
  for (long long i = (long long)blockIdx.x * TPB + threadIdx.x; i < n; i += (long long)gridDim.x * TPB) p[i] = v;
}

}  // namespace pamg

// pamg_stream.cuh -- row-streaming element kernel (Jacobi / Richardson / residual) for sm_100a.
//
// In (r, x) coordinates with x = ipos + r - 1 the children of a parent form a regular grid: row r holds
// x in [r, b - r] (b = 2^(s+1)); left/right neighbours are x -/+ 1 and the vertical neighbour of an up child
// (row r-1) or a down child (row r+1) has the SAME x (splitting.F90:741-774 rewritten in these coordinates).
// A work item is a strip of SW columns x a chunk of SR rows of one parent.  The CTA (SW threads, one per
// column) marches up the rows; every row segment is one contiguous span of memory, fetched exactly once per
// item by a 1-D bulk copy (cp.async.bulk -> UBLKCP, mbarrier completion) into a ring of row buffers that is
// kept PF rows ahead of the compute.  All six neighbour values come from shared memory; results leave through
// bulk stores.  HBM traffic is the algorithmic 24 B/DOF, L2->SM traffic ~ (1 + 2/SR + 2/SW) * 24 + 24 B/DOF
// instead of 96 B/DOF for the direct kernel, and nothing strides through L1 any more.
#pragma once
#include "pamg_kernels.cuh"

namespace pamg {

constexpr int SW = 128;        // columns per strip = threads per CTA
constexpr int SR = 32;         // rows per chunk
constexpr int PF = 4;          // prefetch distance of T rows
constexpr int PFB = 3;         // prefetch distance of rhs rows
constexpr int NRT = 8;         // ring of T row buffers (>= PF + 3)
constexpr int NRB = 4;         // ring of rhs row buffers (>= PFB + 1)
constexpr int ROWBUF = 3 * (SW + 4);   // doubles per T row buffer (strip + halo + alignment slack)
constexpr int BBUF = 3 * (SW + 2);
constexpr int STREAM_MIN_S = 6;   // rows must be at least one strip long; coarser levels use the direct kernel

struct StreamArgs {
  ElemArgs e;
  const int2* items;   // (parent, r0 << 16 | x0), largest first
  int nitems;
  int* counters;       // [0] next item, [1] CTAs that ran out of work (the last one resets both)
};

// span of row q needed by a strip [x0, x1]: columns [max(x0-1,q), min(x1+1,b-q)], as 0-based child indices of
// the parent, start rounded down to even and count rounded up to even (16-byte granularity of bulk copies)
__device__ __forceinline__ void row_span(int q, int x0, int x1, int b, int S, int halo, int& ea, int& cnt) {
  cnt = 0; ea = 0;
  if (q < 1 || q > S) return;
  const int xa = max(x0 - halo, q), xb = min(x1 + halo, b - q);
  if (xb < xa) return;
  const int e0 = (q - 1) * (b + 1 - q) - q;   // child index of column x is e0 + x
  ea = (e0 + xa) & ~1;
  cnt = ((e0 + xb) - ea + 2) & ~1;
}

template <int MODE, bool FACE>
__global__ void __launch_bounds__(SW) k_stream(StreamArgs sa) {
  const ElemArgs& a = sa.e;
  __shared__ __align__(128) double sT[NRT][ROWBUF];
  __shared__ __align__(128) double sB[NRB][BBUF];
  __shared__ __align__(128) double sO[2][3 * (SW + 2)];
  __shared__ __align__(8) uint64_t barT[NRT];
  __shared__ __align__(8) uint64_t barB[NRB];
  __shared__ int eaT[NRT];
  __shared__ int s_item;
  const int s = a.s, twos = 2 * s, b = 2 << s, S = 1 << s;
  const int tid = threadIdx.x;
  double acc_sum = 0.0, acc_abs = 0.0, acc_max = 0.0;
  if (tid == 0) {
    for (int i = 0; i < NRT; ++i) mbar_init(&barT[i], 1);
    for (int i = 0; i < NRB; ++i) mbar_init(&barB[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  uint32_t phT = 0, phB = 0;   // per-slot phase bits, tracked identically by every thread
  int obuf = 0;

  for (;;) {
    // dynamic scheduling: items differ in size (the strips are cut out of a triangle)
    if (tid == 0) s_item = atomicAdd(&sa.counters[0], 1);
    __syncthreads();
    const int it = s_item;
    if (it >= sa.nitems) break;
    const int2 item = sa.items[it];
    const int u = item.x;
    const int r0 = item.y >> 16, x0 = item.y & 0xffff;
    const int x1 = x0 + SW - 1;
    const int rend = min(min(r0 + SR - 1, S), min(x1, b - x0));
    const long long pbase = ((long long)u << twos);
    const double* __restrict__ pc = a.pc + (size_t)u * NPC;
    const int qa = max(1, r0 - 1), qb = min(S, rend + 1);

    auto issueT = [&](int q) {
      int ea, cnt;
      row_span(q, x0, x1, b, S, 1, ea, cnt);
      if (cnt == 0) return;
      const int sl = q % NRT;
      eaT[sl] = ea;
      mbar_expect_tx(&barT[sl], (uint32_t)cnt * 24u);
      tma_load_1d(sT[sl], a.Tin + (pbase + ea) * 3, (uint32_t)cnt * 24u, &barT[sl]);
    };
    auto issueB = [&](int r) {
      int ea, cnt;
      row_span(r, x0, x1, b, S, 0, ea, cnt);
      if (cnt == 0) return;
      const int sl = r % NRB;
      mbar_expect_tx(&barB[sl], (uint32_t)cnt * 24u);
      tma_load_1d(sB[sl], a.rhs + (pbase + ea) * 3, (uint32_t)cnt * 24u, &barB[sl]);
    };
    auto waitT = [&](int q) {
      int ea, cnt;
      row_span(q, x0, x1, b, S, 1, ea, cnt);
      if (cnt == 0) return;
      const int sl = q % NRT;
      mbar_wait(&barT[sl], (phT >> sl) & 1u);
      phT ^= 1u << sl;
    };

    if (tid == 0) {
      for (int q = qa; q <= min(qb, r0 + PF); ++q) issueT(q);
      for (int r = r0; r <= min(rend, r0 + PFB - 1); ++r) issueB(r);
    }
    if (qa < r0) waitT(qa);
    waitT(r0);

    for (int r = r0; r <= rend; ++r) {
      if (tid == 0) {
        if (r + 1 + PF <= qb) issueT(r + 1 + PF);
        if (r + PFB <= rend) issueB(r + PFB);
      }
      if (r + 1 <= qb) waitT(r + 1);
      int eb, cntb;
      row_span(r, x0, x1, b, S, 0, eb, cntb);
      {
        const int sl = r % NRB;
        mbar_wait(&barB[sl], (phB >> sl) & 1u);
        phB ^= 1u << sl;
      }
      const int xlo = max(x0, r), xhi = min(x1, b - r);
      const int x = x0 + tid;
      const bool active = (x >= xlo) && (x <= xhi);
      const int e0r = (r - 1) * (b + 1 - r) - r;     // child index of column x in row r is e0r + x
      const int elo = e0r + xlo, nrow = xhi - xlo + 1;
      const int head = elo & 1;                       // first child not 16-byte aligned in global memory
      double* so = sO[obuf];
      if (active) {
        const int ipos = x - r + 1, len = b + 1 - 2 * r;
        const bool up = ipos & 1;
        const double* t = sT[r % NRT] + (size_t)(e0r + x - eaT[r % NRT]) * 3;
        const double T1 = t[0], T2 = t[1], T3 = t[2];
        FaceIn fi;
        if (FACE) {
          if (!up) {
            const int q = r + 1, e0q = (q - 1) * (b + 1 - q) - q;
            const double* v = sT[q % NRT] + (size_t)(e0q + x - eaT[q % NRT]) * 3;
            fi.n1a = v[2]; fi.n1b = v[0];
            fi.n2a = t[3 + 1]; fi.n2b = t[3 + 2];
            fi.n3a = t[-3 + 0]; fi.n3b = t[-3 + 1];
            fi.pen1 = __ldg(pc + PC_PENI + 0); fi.pen2 = __ldg(pc + PC_PENI + 1); fi.pen3 = __ldg(pc + PC_PENI + 2);
          } else {
            if (r > 1) {
              const int q = r - 1, e0q = (q - 1) * (b + 1 - q) - q;
              const double* v = sT[q % NRT] + (size_t)(e0q + x - eaT[q % NRT]) * 3;
              fi.n1a = v[2]; fi.n1b = v[0];
              fi.pen1 = __ldg(pc + PC_PENI + 0);
            } else { halo_pair(a, u, 0, ipos >> 1, S, fi.n1a, fi.n1b); fi.pen1 = __ldg(pc + PC_PENX + 0); }
            if (ipos > 1) { fi.n2a = t[-3 + 1]; fi.n2b = t[-3 + 2]; fi.pen2 = __ldg(pc + PC_PENI + 1); }
            else { halo_pair(a, u, 2, r - 1, S, fi.n2a, fi.n2b); fi.pen2 = __ldg(pc + PC_PENX + 1); }
            if (ipos < len) { fi.n3a = t[3 + 0]; fi.n3b = t[3 + 1]; fi.pen3 = __ldg(pc + PC_PENI + 2); }
            else { halo_pair(a, u, 1, r - 1, S, fi.n3a, fi.n3b); fi.pen3 = __ldg(pc + PC_PENX + 2); }
          }
        }
        const double* bb = sB[r % NRB] + (size_t)(e0r + x - eb) * 3;
        double o1, o2, o3;
        elem_apply<MODE, FACE>(pc, up, T1, T2, T3, fi, bb[0], bb[1], bb[2], a.omega, a.rsign, o1, o2, o3);
        // staged at index j + head so that the 16-byte aligned part of the row starts on a 16-byte boundary
        const int j = x - xlo + head;
        so[j * 3] = o1; so[j * 3 + 1] = o2; so[j * 3 + 2] = o3;
        if (MODE == MODE_RESID) {
          acc_sum += o1 * o1 + o2 * o2 + o3 * o3;
          acc_abs = fmax(acc_abs, fmax(fabs(o1), fmax(fabs(o2), fabs(o3))));
          acc_max = fmax(acc_max, fmax(o1, fmax(o2, o3)));
        }
        // row ends that a bulk store cannot cover (16-byte granularity): plain stores
        const int nmid = (nrow - head) & ~1;
        const int jj = x - xlo;
        if (jj < head || jj >= head + nmid) {
          double* g = a.Tout + (pbase + e0r + x) * 3;
          g[0] = o1; g[1] = o2; g[2] = o3;
        }
      }
      if (tid == 0) tma_store_wait_read();   // the store issued one row ago has drained the other staging buffer
      fence_async_smem();
      __syncthreads();
      if (tid == 0) {
        const int nmid = (nrow - head) & ~1;
        if (nmid > 0) {
          tma_store_1d(a.Tout + (pbase + elo + head) * 3, so + (size_t)(2 * head) * 3, (uint32_t)nmid * 24u);
          tma_store_commit();
        }
      }
      obuf ^= 1;
    }
  }
  if (tid == 0) {
    tma_store_wait_all();
    if (atomicAdd(&sa.counters[1], 1) == (int)gridDim.x - 1) {   // every CTA has stopped fetching work
      sa.counters[0] = 0; sa.counters[1] = 0;
      __threadfence();
    }
  }
  if (MODE == MODE_RESID) {
    for (int o = 16; o > 0; o >>= 1) {
      acc_sum += __shfl_xor_sync(0xffffffffu, acc_sum, o);
      acc_abs = fmax(acc_abs, __shfl_xor_sync(0xffffffffu, acc_abs, o));
      acc_max = fmax(acc_max, __shfl_xor_sync(0xffffffffu, acc_max, o));
    }
    __shared__ double sh[3][SW / 32];
    const int w = tid >> 5, l = tid & 31;
    if (l == 0) { sh[0][w] = acc_sum; sh[1][w] = acc_abs; sh[2][w] = acc_max; }
    __syncthreads();
    if (tid == 0) {
      double s0 = 0, s1 = 0, s2 = 0;
      for (int i = 0; i < SW / 32; ++i) { s0 += sh[0][i]; s1 = fmax(s1, sh[1][i]); s2 = fmax(s2, sh[2][i]); }
      a.partial[(size_t)blockIdx.x * 3 + 0] = s0;
      a.partial[(size_t)blockIdx.x * 3 + 1] = s1;
      a.partial[(size_t)blockIdx.x * 3 + 2] = s2;
    }
  }
}

}  // namespace pamg

// pamg_unstr.cuh -- fully unstructured explicit P1 DG step (unstr_explicit,
// transport_tri_unstr.F90:588-795) and the batched element-local inverse (FINDInv,
// matrix_inversion.F90:50-148 == matrices.F90:1618-1716) as sm_100a kernels.
//
// One thread per element.  Geometry (tri_det_nlx ShapFun.F90:1414-1454, det_snlx_all :1554-1590) is
// recomputed in registers from the 6 vertex coordinates - cheaper than streaming 19 stored doubles.
// Neighbour values are gathered through L2.  The 3x3 / 4x4 / 6x6 local systems live entirely in
// registers (fully unrolled Gauss-Jordan); tensor cores are pointless for blocks this small.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <string>
#include <vector>

#include "pamg.h"
#include "pamg_kernels.cuh"

namespace pamg {

struct UnstrDev {
  int E = 0;
  double* X = nullptr;        // [E][3][2]
  int32_t* neig = nullptr;    // [E][3] 1-based, 0 = boundary
  int32_t* nside = nullptr;   // [E][3] fNeig | swap<<2 (swap: geometric pairing differs from get_unstr_sn2)
  double* T[2] = {nullptr, nullptr};
  double* told = nullptr;
  int cur = 0;
};

struct UnstrArgs {
  const double* X; const int32_t* neig; const int32_t* nside;
  const double* Tin; const double* told; double* Tout;
  double dt, ux, uy, t_bc;
  int E, njac, exact, use_dir;
};

__global__ void __launch_bounds__(TPB) k_unstr_explicit(UnstrArgs a) {
  const double al = 0.78867513459481288, be = 0.21132486540518712;  // sn_orig, ShapFun.F90:1100-1111
  // weights of int sn_c * trace over the face: (2/3, 1/3) -- exact products of the 2-point Gauss rule
  const double w2 = al * al + be * be, w1 = 2.0 * al * be;
  for (int e = blockIdx.x * TPB + threadIdx.x; e < a.E; e += gridDim.x * TPB) {
    const double* __restrict__ X = a.X + (size_t)e * 6;
    const double x1 = __ldg(X), y1 = __ldg(X + 1), x2 = __ldg(X + 2), y2 = __ldg(X + 3), x3 = __ldg(X + 4), y3 = __ldg(X + 5);
    const double A = x1 - x3, B = y1 - y3, C = x2 - x3, D = y2 - y3;
    const double detj = A * D - B * C;
    const double area = 0.5 * fabs(detj);
    const double gx[3] = {D / detj, -B / detj, -(D / detj) - (-B / detj)};
    const double gy[3] = {-C / detj, A / detj, -(-C / detj) - (A / detj)};
    const double T[3] = {__ldg(a.Tin + (size_t)e * 3), __ldg(a.Tin + (size_t)e * 3 + 1), __ldg(a.Tin + (size_t)e * 3 + 2)};
    const double To[3] = {__ldg(a.told + (size_t)e * 3), __ldg(a.told + (size_t)e * 3 + 1), __ldg(a.told + (size_t)e * 3 + 2)};
    const double sumT = T[0] + T[1] + T[2];
    double rhs[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) rhs[i] = (gx[i] * a.ux + gy[i] * a.uy) * (area / 3.0) * sumT;  // :668-672
    const double cx = (x1 + x2 + x3) / 3.0, cy = (y1 + y2 + y3) / 3.0;
    const double px[3] = {x1, x2, x3}, py[3] = {y1, y2, y3};
    // gmsh faces: 1 = nodes (1,3), 2 = (2,1), 3 = (3,2)  (ShapFun_unstruc.F90:160-188)
    const int L1[3] = {0, 1, 2}, L2[3] = {2, 0, 1};
#pragma unroll
    for (int f = 0; f < 3; ++f) {
      const int l1 = L1[f], l2 = L2[f];
      const double ex = px[l2] - px[l1], ey = py[l2] - py[l1];
      const double len = sqrt(ex * ex + ey * ey);
      double nx = ey / len, ny = -ex / len;
      const double mx = 0.5 * (px[l1] + px[l2]) - cx, my = 0.5 * (py[l1] + py[l2]) - cy;
      if (nx * mx + ny * my < 0.0) { nx = -nx; ny = -ny; }
      const double sdet = 0.5 * len;
      const int q = __ldg(a.neig + (size_t)e * 3 + f);
      const int enc = __ldg(a.nside + (size_t)e * 3 + f);
      const int ns = enc & 3;
      double T2a = 0.0, T2b = 0.0, has2 = 0.0;   // neighbour values paired with sn_orig(:,1), sn_orig(:,2)
      if (ns >= 1) {
        // get_unstr_sn2 (ShapFun_unstruc.F90:205-222): Nside 1 -> nodes (3,1), 2 -> (1,2), 3 -> (2,3)
        int m1 = (ns == 1) ? 2 : (ns == 2 ? 0 : 1), m2 = (ns == 1) ? 0 : (ns == 2 ? 1 : 2);
        if (a.use_dir && (enc >> 2)) { const int t = m1; m1 = m2; m2 = t; }
        has2 = 1.0;
        if (q != 0) { T2a = __ldg(a.Tin + (size_t)(q - 1) * 3 + m1); T2b = __ldg(a.Tin + (size_t)(q - 1) * 3 + m2); }
        else { T2a = a.t_bc; T2b = a.t_bc; }
      }
      const double unn = nx * a.ux + ny * a.uy;
      const double un = 0.5 * (unn + has2 * unn);       // n . (u + u2)/2 ; u2 = 0 when sn2 = 0 (Nside = 0)
      const bool in = signbit(-un) == 0;                // income = 0.5 + 0.5*sign(1, -un)  (:731)
      double ca, cb;
      if (in) { const double f2 = sdet * has2 * unn; ca = f2 * (w2 * T2a + w1 * T2b); cb = f2 * (w1 * T2a + w2 * T2b); }
      else { const double f1 = sdet * unn; ca = f1 * (w2 * T[l1] + w1 * T[l2]); cb = f1 * (w1 * T[l1] + w2 * T[l2]); }
      rhs[l1] -= ca; rhs[l2] -= cb;                     // :742-746
    }
    const double m12 = area / 12.0, ml = area / 3.0;
    const double so = To[0] + To[1] + To[2];
    double out[3];
    if (a.exact) {
      // T = M^-1 (M told + dt rhs), M^-1 = (12/A)(I - J/4)  (transport_rect.F90:277-291 semantics)
      double v[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) v[i] = m12 * (To[i] + so) + a.dt * rhs[i];
      const double sv = v[0] + v[1] + v[2];
#pragma unroll
      for (int i = 0; i < 3; ++i) out[i] = (12.0 / area) * (v[i] - 0.25 * sv);
    } else {
      double rj[3], tl[3] = {T[0], T[1], T[2]};         // tnew_nonlin(:,ele) == tnew(:,ele) here (:598,783)
#pragma unroll
      for (int i = 0; i < 3; ++i) rj[i] = m12 * (To[i] + so) + a.dt * rhs[i];   // :774
      for (int it = 0; it < a.njac; ++it) {
        const double st = tl[0] + tl[1] + tl[2];
        double nt[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) nt[i] = (ml * tl[i] - m12 * (tl[i] + st) + rj[i]) / ml;  // :780-787
        tl[0] = nt[0]; tl[1] = nt[1]; tl[2] = nt[2];
      }
      out[0] = tl[0]; out[1] = tl[1]; out[2] = tl[2];
    }
    a.Tout[(size_t)e * 3] = out[0]; a.Tout[(size_t)e * 3 + 1] = out[1]; a.Tout[(size_t)e * 3 + 2] = out[2];
  }
}

inline void unstr_free(UnstrDev& u) {
  cudaFree(u.X); cudaFree(u.neig); cudaFree(u.nside); cudaFree(u.T[0]); cudaFree(u.T[1]); cudaFree(u.told);
  u = UnstrDev();
}

inline int unstr_setup(UnstrDev& u, int E, const double* X, const int32_t* neig, const int32_t* fneig,
                       cudaStream_t st, std::string& err) {
  unstr_free(u);
  // geometric pairing flag per face: does the neighbour node that get_unstr_sn2 pairs with my first face
  // node actually coincide with it?  (SURVEY B-9: the reference ignores Dir here)
  std::vector<int32_t> enc((size_t)E * 3);
  const int L1[3] = {0, 1, 2};
  for (int e = 0; e < E; ++e)
    for (int f = 0; f < 3; ++f) {
      const int q = neig[(size_t)e * 3 + f], ns = fneig[(size_t)e * 3 + f];
      int sw = 0;
      if (q != 0 && ns >= 1 && ns <= 3) {
        const int m1 = (ns == 1) ? 2 : (ns == 2 ? 0 : 1);
        const double* a = X + (size_t)e * 6 + 2 * L1[f];
        const double* b = X + (size_t)(q - 1) * 6 + 2 * m1;
        sw = !(a[0] == b[0] && a[1] == b[1]);
      } else if (q != 0) { err = "fNeig must be 1..3 where Neig != 0"; return PAMG_ERR_ARG; }
      enc[(size_t)e * 3 + f] = (ns & 3) | (sw << 2);
    }
#define UCK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { err = std::string(#call) + ": " + cudaGetErrorString(e_); return PAMG_ERR_CUDA; } } while (0)
  u.E = E;
  UCK(cudaMalloc(&u.X, (size_t)E * 6 * sizeof(double)));
  UCK(cudaMalloc(&u.neig, (size_t)E * 3 * sizeof(int32_t)));
  UCK(cudaMalloc(&u.nside, (size_t)E * 3 * sizeof(int32_t)));
  for (int i = 0; i < 2; ++i) UCK(cudaMalloc(&u.T[i], (size_t)E * 3 * sizeof(double)));
  UCK(cudaMalloc(&u.told, (size_t)E * 3 * sizeof(double)));
  UCK(cudaMemcpyAsync(u.X, X, (size_t)E * 6 * sizeof(double), cudaMemcpyHostToDevice, st));
  UCK(cudaMemcpyAsync(u.neig, neig, (size_t)E * 3 * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  UCK(cudaMemcpyAsync(u.nside, enc.data(), (size_t)E * 3 * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  UCK(cudaMemsetAsync(u.T[0], 0, (size_t)E * 3 * sizeof(double), st));
  UCK(cudaMemsetAsync(u.T[1], 0, (size_t)E * 3 * sizeof(double), st));
  UCK(cudaStreamSynchronize(st));
  u.cur = 0;
  return PAMG_OK;
}

// time loop of unstr_explicit (:588-795): told = tnew ; nits x { tnew = tnew_nonlin ; element loop }
inline int unstr_step(UnstrDev& u, double dt, double ux, double uy, double t_bc, int ntime, int nits, int njac,
                      int exact, int use_dir, int nsm, cudaStream_t st, long long& nlaunch, std::string& err) {
  const int grid = std::max(1, std::min((u.E + TPB - 1) / TPB, nsm * 8));
  for (int it = 0; it < ntime; ++it) {
    UCK(cudaMemcpyAsync(u.told, u.T[u.cur], (size_t)u.E * 3 * sizeof(double), cudaMemcpyDeviceToDevice, st));
    for (int k = 0; k < nits; ++k) {
      UnstrArgs a;
      a.X = u.X; a.neig = u.neig; a.nside = u.nside; a.Tin = u.T[u.cur]; a.told = u.told; a.Tout = u.T[u.cur ^ 1];
      a.dt = dt; a.ux = ux; a.uy = uy; a.t_bc = t_bc; a.E = u.E; a.njac = njac; a.exact = exact; a.use_dir = use_dir;
      k_unstr_explicit<<<grid, TPB, 0, st>>>(a);
      nlaunch++;
      UCK(cudaGetLastError());
      u.cur ^= 1;
    }
  }
  return PAMG_OK;
}

// ------------------------------------------------------------------------------------------------
// FINDInv: Gauss-Jordan on [M I], NO partial pivoting; a zero pivot is repaired by ADDING the first lower
// row with a non-zero entry (matrices.F90:1661-1676); errorflag -1 when singular.
template <int N>
__global__ void __launch_bounds__(128) k_local_minv(const double* __restrict__ M, const double* __restrict__ rhs,
                                                     double* __restrict__ x, double* __restrict__ Minv,
                                                     int32_t* __restrict__ status, int batch) {
  for (int bi = blockIdx.x * 128 + threadIdx.x; bi < batch; bi += gridDim.x * 128) {
    double a[N][2 * N];
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int j = 0; j < 2 * N; ++j) a[i][j] = (j < N) ? M[((size_t)bi * N + i) * N + j] : ((i + N) == j ? 1.0 : 0.0);
    bool ok = true;
#pragma unroll
    for (int k = 0; k < N - 1; ++k) {
      if (ok && a[k][k] == 0.0) {
        // as written the reference only ever tries row k+1: a zero there returns "non-invertible" at once
        if (a[k + 1][k] != 0.0) {
#pragma unroll
          for (int j = 0; j < 2 * N; ++j) a[k][j] += a[k + 1][j];
        } else {
          ok = false;
        }
      }
      if (ok) {
#pragma unroll
        for (int j = k + 1; j < N; ++j) {
          const double m = a[j][k] / a[k][k];
#pragma unroll
          for (int i = 0; i < 2 * N; ++i) if (i >= k) a[j][i] -= m * a[k][i];
        }
      }
    }
#pragma unroll
    for (int i = 0; i < N; ++i) if (a[i][i] == 0.0) ok = false;
    if (ok) {
#pragma unroll
      for (int i = 0; i < N; ++i) {
        const double m = a[i][i];
#pragma unroll
        for (int j = 0; j < 2 * N; ++j) if (j >= i) a[i][j] /= m;
      }
#pragma unroll
      for (int k = N - 2; k >= 0; --k)
#pragma unroll
        for (int i = 0; i <= k; ++i) {
          const double m = a[i][k + 1];
#pragma unroll
          for (int j = 0; j < 2 * N; ++j) if (j >= k) a[i][j] -= a[k + 1][j] * m;
        }
    }
    if (status) status[bi] = ok ? 0 : -1;
    if (Minv)
#pragma unroll
      for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = 0; j < N; ++j) Minv[((size_t)bi * N + i) * N + j] = ok ? a[i][N + j] : 0.0;
    if (x) {
#pragma unroll
      for (int i = 0; i < N; ++i) {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < N; ++j) s += a[i][N + j] * rhs[(size_t)bi * N + j];
        x[(size_t)bi * N + i] = ok ? s : 0.0;
      }
    }
  }
}

inline int local_minv(int n, int batch, const double* M, const double* rhs, double* x, double* Minv, int32_t* status,
                      int nsm, cudaStream_t st, long long& nlaunch, std::string& err) {
  double *dM = nullptr, *dr = nullptr, *dx = nullptr, *dI = nullptr;
  int32_t* ds = nullptr;
  const size_t nm = (size_t)batch * n * n, nv = (size_t)batch * n;
  int rc = PAMG_OK;
  auto done = [&]() { cudaFree(dM); cudaFree(dr); cudaFree(dx); cudaFree(dI); cudaFree(ds); };
#define MCK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { err = std::string(#call) + ": " + cudaGetErrorString(e_); done(); return PAMG_ERR_CUDA; } } while (0)
  MCK(cudaMalloc(&dM, nm * sizeof(double)));
  MCK(cudaMalloc(&ds, (size_t)batch * sizeof(int32_t)));
  MCK(cudaMemcpyAsync(dM, M, nm * sizeof(double), cudaMemcpyHostToDevice, st));
  if (rhs) {
    MCK(cudaMalloc(&dr, nv * sizeof(double))); MCK(cudaMalloc(&dx, nv * sizeof(double)));
    MCK(cudaMemcpyAsync(dr, rhs, nv * sizeof(double), cudaMemcpyHostToDevice, st));
  }
  if (Minv) MCK(cudaMalloc(&dI, nm * sizeof(double)));
  const int grid = std::max(1, std::min((batch + 127) / 128, nsm * 8));
  if (n == 3) k_local_minv<3><<<grid, 128, 0, st>>>(dM, dr, dx, dI, ds, batch);
  else if (n == 4) k_local_minv<4><<<grid, 128, 0, st>>>(dM, dr, dx, dI, ds, batch);
  else k_local_minv<6><<<grid, 128, 0, st>>>(dM, dr, dx, dI, ds, batch);
  nlaunch++;
  MCK(cudaGetLastError());
  if (x) MCK(cudaMemcpyAsync(x, dx, nv * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (Minv) MCK(cudaMemcpyAsync(Minv, dI, nm * sizeof(double), cudaMemcpyDeviceToHost, st));
  std::vector<int32_t> hs(batch);
  MCK(cudaMemcpyAsync(hs.data(), ds, (size_t)batch * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  MCK(cudaStreamSynchronize(st));
  for (int i = 0; i < batch; ++i) {
    if (status) status[i] = hs[i];
    if (hs[i] != 0) rc = PAMG_ERR_SINGULAR;
  }
  done();
  if (rc) err = "singular block (errorflag = -1)";
  return rc;
}

}  // namespace pamg

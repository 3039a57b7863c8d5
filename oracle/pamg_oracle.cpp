// pamg_oracle.cpp -- CPU oracle (fp64) for the P-A_multigrids hot path.
// TEST INFRASTRUCTURE ONLY; see pamg_oracle.h for the rules.  PARITY UNPINNED (no Fortran
// compiler exists here, the reference has no tests); pinned by tests/test_oracle_*.py.
//
// Written in the reference's own "loop over Gauss points / stencils" style on purpose:
// the CUDA product uses closed forms, so agreement between the two is meaningful.
// Index convention: helper arithmetic is 1-based like the Fortran, storage is 0-based.
#include "pamg_oracle.h"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <fstream>
#include <sstream>
#include <algorithm>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

// ------------------------------------------------------------------ tables (a1/a2)
// ShapFun.F90:554-563,659-661 (TRIQUAold NGI=3), :1036-1056 (SHATRIold NLOC=3),
// :1100-1112 (1-D face branch).
struct Tables {
  double n[3][3];       // n[g][i]
  double nlx[3][2][3];  // nlx[g][d][i]
  double w[3];
  double sn[2][2];      // sn_orig[s][c]
  double snlx[2][2];    // snlx_orig[s][c]  (single local direction)
  double sw[2];
  Tables() {
    double L1[3] = {0.5, 0.0, 0.5}, L2[3] = {0.5, 0.5, 0.0}, L3[3];
    for (int g = 0; g < 3; ++g) { L3[g] = 1.0 - L1[g] - L2[g]; w[g] = 1.0 / 3.0; }
    for (int g = 0; g < 3; ++g) {
      n[g][0] = L1[g]; n[g][1] = L2[g]; n[g][2] = L3[g];
      nlx[g][0][0] = 1; nlx[g][0][1] = 0; nlx[g][0][2] = -1;
      nlx[g][1][0] = 0; nlx[g][1][1] = 1; nlx[g][1][2] = -1;
    }
    double lxp[2] = {-1, 1}, lx[2] = {-1.0 / std::sqrt(3.0), 1.0 / std::sqrt(3.0)};
    for (int p = 0; p < 2; ++p)
      for (int c = 0; c < 2; ++c) {
        sn[p][c] = 0.5 * (1.0 + lxp[c] * lx[p]);
        snlx[p][c] = 0.5 * lxp[c];
        sw[p] = 1.0;
      }
  }
};
const Tables TB;

// child-face -> volume nodes, transport_tri_semi.F90:142-147 (1-based values)
const int FACE_NODES[3][2] = {{1, 3}, {3, 2}, {2, 1}};
// gmsh parent face -> volume nodes carrying sn_orig(:,1), sn_orig(:,2); ShapFun_unstruc.F90:160-188
const int UN_FACE_NODES[3][2] = {{1, 3}, {2, 1}, {3, 2}};
// child face -> parent (gmsh) face, transport_tri_semi.F90:629-638
const int MFACE[3] = {1, 3, 2};

inline int ipow(int b, int e) { int r = 1; while (e-- > 0) r *= b; return r; }

// ------------------------------------------------------------------ numbering (a7)
// Msh2Tri.F90:42-58 (subtract-row-lengths loop, kept literally)
void get_str_info(int n_split, int ele, int& irow, int& ipos, int& orientation) {
  int i = ele, row = 1, ele_row = ipow(2, n_split + 1) - 1;
  irow = 0; ipos = 0;
  while (i >= 1) {
    if (i > ele_row) { i -= ele_row; row += 1; ele_row -= 2; }
    else { ipos = i; irow = row; break; }
  }
  orientation = ipos % 2;
}

// Msh2Tri.F90:79-106.  X[node][dim], 0-based storage of 1-based nodes.
void get_splitting(const double X[3][2], int n_split, int str_ele, double x[3][2]) {
  double s = (double)ipow(2, n_split);
  double v1[2] = {(X[0][0] - X[2][0]) / s, (X[0][1] - X[2][1]) / s};
  double v2[2] = {(X[1][0] - X[2][0]) / s, (X[1][1] - X[2][1]) / s};
  int irow, ipos, ori;
  get_str_info(n_split, str_ele, irow, ipos, ori);
  for (int d = 0; d < 2; ++d) {
    if (ipos % 2 != 0) {
      x[2][d] = X[2][d] + (irow - 1) * v2[d] + (ipos / 2) * v1[d];
      x[1][d] = X[2][d] + irow * v2[d] + (ipos / 2) * v1[d];
      x[0][d] = X[2][d] + (irow - 1) * v2[d] + v1[d] * (ipos / 2 + 1);
    } else {
      x[0][d] = X[2][d] + irow * v2[d] + v1[d] * (ipos / 2 - 1);
      x[1][d] = X[2][d] + (irow - 1) * v2[d] + v1[d] * (ipos / 2);
      x[2][d] = X[2][d] + irow * v2[d] + v1[d] * (ipos / 2);
    }
  }
}

// splitting.F90:741-774.  out[(ele-1)*3 + (f-1)]
void str_neig_table(int n, std::vector<int>& t) {
  int C = ipow(4, n);
  t.assign((size_t)3 * (C + 2), 0);
  auto S = [&](int f, int ele) -> int& { return t[(size_t)(ele - 1) * 3 + (f - 1)]; };
  int total = ipow(2, n + 1) - 1, current = total, irow = ipow(2, n) - 1;
  S(1, 1) = 0; S(2, 1) = 0; S(3, 1) = 2;
  int ele = 2;
  while (ele <= total) {
    S(2, ele) = ele + 1; S(3, ele) = ele - 1; S(1, ele) = ele + total - 1;
    ele++;
    S(2, ele) = ele - 1; S(1, ele) = 0; S(3, ele) = ele + 1;
    ele++;
  }
  S(3, ele - 1) = 0;
  while (irow >= 1) {
    total = total + current - 2;
    current = current - 2;
    S(2, ele) = 0; S(3, ele) = ele + 1; S(1, ele) = ele - current - 1;
    ele++;
    while (ele <= total) {
      S(2, ele) = ele + 1; S(3, ele) = ele - 1; S(1, ele) = ele + current - 1;
      ele++;
      S(2, ele) = ele - 1; S(3, ele) = ele + 1; S(1, ele) = ele - current - 1;
      ele++;
    }
    S(3, ele - 1) = 0;
    irow--;
  }
  t.resize((size_t)3 * C);
}

// splitting.F90:434-449.  out[(f-1)*S + (i-1)]
void surf_ele_table(int n, std::vector<int>& t) {
  int S = ipow(2, n);
  t.assign((size_t)3 * S, 0);
  auto A = [&](int i, int f) -> int& { return t[(size_t)(f - 1) * S + (i - 1)]; };
  A(1, 1) = 1;
  int ele;
  for (ele = 2; ele <= S; ++ele) A(ele, 1) = A(ele - 1, 1) + 2;
  A(1, 3) = 1;
  int counter = A(ele - 1, 1);
  A(1, 2) = counter;
  for (ele = 2; ele <= S; ++ele) {
    A(ele, 2) = A(ele - 1, 2) + counter - 2;
    A(ele, 3) = A(ele - 1, 2) + 1;
    counter -= 2;
  }
}

// splitting.F90:105-139.  i_split is the COARSE split.
void element_conversion(int coarse_ele, int i_split, int fin[4]) {
  int irow, ipos, ori;
  get_str_info(i_split, coarse_ele, irow, ipos, ori);
  int tot_fine = 0, rowx = ipow(2, i_split + 1) * 2 - 1, counter;
  if (ori == 1) {
    counter = 2;
    while (counter < irow * 2) { tot_fine += rowx; rowx -= 2; counter++; }
    fin[0] = ipos * 2 - 1 + tot_fine;
    fin[1] = fin[0] + 1;
    fin[2] = fin[0] + 2;
    tot_fine += rowx;
    fin[3] = ipos * 2 - 1 + tot_fine;
  } else {
    counter = 1;
    while (counter < irow * 2) { tot_fine += rowx; rowx -= 2; counter++; }
    fin[2] = (ipos / 2 - 1) * 3 + ipos / 2 + tot_fine + 1;
    fin[1] = fin[2] + 1;
    fin[0] = fin[2] + 2;
    fin[3] = fin[0] - rowx - 2;
  }
}

// ------------------------------------------------------------------ geometry (a3/a4)
// ShapFun.F90:1414-1454
void tri_det_nlx(const double x[3][2], double nx[3][2][3], double detwei[3]) {
  for (int g = 0; g < 3; ++g) {
    double A = 0, B = 0, C = 0, D = 0;
    for (int l = 0; l < 3; ++l) {
      A += TB.nlx[g][0][l] * x[l][0];
      B += TB.nlx[g][0][l] * x[l][1];
      C += TB.nlx[g][1][l] * x[l][0];
      D += TB.nlx[g][1][l] * x[l][1];
    }
    double detj = A * D - B * C;
    detwei[g] = 0.5 * std::fabs(detj) * TB.w[g];
    double a11 = D / detj, a21 = -C / detj, a12 = -B / detj, a22 = A / detj;
    for (int l = 0; l < 3; ++l) {
      nx[g][0][l] = a11 * TB.nlx[g][0][l] + a12 * TB.nlx[g][1][l];
      nx[g][1][l] = a21 * TB.nlx[g][0][l] + a22 * TB.nlx[g][1][l];
    }
  }
}

// det_snlx_all (ShapFun.F90:1554-1590) + NORMGI/XPROD1 (:2012-2054), 2-D branch, with the
// gmsh face tables of ShapFun_unstruc.F90:160-188 and the approximate outward direction of
// ShapFun.F90:1751-1762 / transport_tri_unstr.F90:720-722.  iface is the gmsh face (1..3).
void face_geometry(const double x[3][2], int iface, double sdetwei[2], double snorm[2][2]) {
  double face_sn[2][3] = {{0, 0, 0}, {0, 0, 0}}, face_snlx[2][3] = {{0, 0, 0}, {0, 0, 0}};
  int l1 = UN_FACE_NODES[iface - 1][0] - 1, l2 = UN_FACE_NODES[iface - 1][1] - 1;
  for (int s = 0; s < 2; ++s) {
    face_sn[s][l1] = TB.sn[s][0]; face_sn[s][l2] = TB.sn[s][1];
    face_snlx[s][l1] = TB.snlx[s][0]; face_snlx[s][l2] = TB.snlx[s][1];
  }
  double xsgi[2][2] = {{0, 0}, {0, 0}};
  for (int l = 0; l < 3; ++l)
    for (int d = 0; d < 2; ++d)
      for (int s = 0; s < 2; ++s) xsgi[s][d] += face_sn[s][l] * x[l][d];
  double norm[2];
  for (int d = 0; d < 2; ++d)
    norm[d] = (xsgi[0][d] + xsgi[1][d]) / 2.0 - (x[0][d] + x[1][d] + x[2][d]) / 3.0;
  for (int s = 0; s < 2; ++s) {
    double dxdlx = 0, dydlx = 0;
    for (int l = 0; l < 3; ++l) { dxdlx += face_snlx[s][l] * x[l][0]; dydlx += face_snlx[s][l] * x[l][1]; }
    double detj = std::sqrt(dydlx * dydlx + dxdlx * dxdlx);
    sdetwei[s] = detj * TB.sw[s];
    // XPROD1 with b=(dxdlx,dydlx,0), c=(0,0,1)
    double ax = dydlx * 1.0 - 0.0 * 0.0, ay = -(dxdlx * 1.0 - 0.0 * 0.0);
    double rn = std::sqrt(ax * ax + ay * ay);
    double sirn = std::copysign(1.0 / rn, ax * norm[0] + ay * norm[1]);
    snorm[s][0] = sirn * ax; snorm[s][1] = sirn * ay;
  }
}

// matrices.F90:1640-1715 (Gauss-Jordan on [M I], no partial pivoting). Row-major.
int findinv(const double* M, double* inv, int n) {
  std::vector<double> a((size_t)n * 2 * n);
  auto A = [&](int i, int j) -> double& { return a[(size_t)(i - 1) * 2 * n + (j - 1)]; };
  for (int i = 1; i <= n; ++i)
    for (int j = 1; j <= 2 * n; ++j)
      A(i, j) = (j <= n) ? M[(i - 1) * n + (j - 1)] : ((i + n) == j ? 1.0 : 0.0);
  bool flag = true;
  for (int k = 1; k <= n - 1; ++k) {
    if (A(k, k) == 0) {
      flag = false;
      for (int i = k + 1; i <= n; ++i) {
        if (A(i, k) != 0) {
          for (int j = 1; j <= 2 * n; ++j) A(k, j) += A(i, j);
          flag = true;
          break;
        }
        if (!flag) { for (int q = 0; q < n * n; ++q) inv[q] = 0; return -1; }
      }
    }
    for (int j = k + 1; j <= n; ++j) {
      double m = A(j, k) / A(k, k);
      for (int i = k; i <= 2 * n; ++i) A(j, i) -= m * A(k, i);
    }
  }
  for (int i = 1; i <= n; ++i)
    if (A(i, i) == 0) { for (int q = 0; q < n * n; ++q) inv[q] = 0; return -1; }
  for (int i = 1; i <= n; ++i) {
    double m = A(i, i);
    for (int j = i; j <= 2 * n; ++j) A(i, j) /= m;
  }
  for (int k = n - 1; k >= 1; --k)
    for (int i = 1; i <= k; ++i) {
      double m = A(i, k + 1);
      for (int j = k; j <= 2 * n; ++j) A(i, j) -= A(k + 1, j) * m;
    }
  for (int i = 1; i <= n; ++i)
    for (int j = 1; j <= n; ++j) inv[(i - 1) * n + (j - 1)] = A(i, j + n);
  return 0;
}

inline double boundary_fn(double a, double b) { return std::sin(a + b); }  // splitting.F90:1401-1405

// ------------------------------------------------------------------ mesh reader
struct Tri { double X[3][2]; int neig[3]; int dir[3]; int region; };

inline bool are_equal2(const double* a, const double* b) {  // Generic.F90:47-57
  double dx = a[0] - b[0], dy = a[1] - b[1];
  return std::sqrt(dx * dx + dy * dy) < 2.220446049250313e-16;
}
inline double get_length(const double* a, const double* b) {  // Msh2Tri.F90:337-345
  return std::sqrt((b[0] - a[0]) * (b[0] - a[0]) + (b[1] - a[1]) * (b[1] - a[1]));
}
inline bool check_vector(const int v[4], int num) { for (int i = 0; i < 4; ++i) if (v[i] == num) return true; return false; }

// Msh2Tri.F90:780-933
void check_neig(std::vector<Tri>& ml, int i, int j, int& no_neig, double l_d) {
  bool one = false, two = false, three = false, one2 = false, two2 = false, three2 = false;
  int vertex[4] = {0, 0, 0, 0};
  Tri& ti = ml[i]; Tri& tj = ml[j];
  int counter = 0;
  for (int a = 0; a < 3 && counter <= 2; ++a)
    for (int b = 0; b < 3; ++b) {
      if (get_length(ti.X[a], tj.X[b]) > l_d) counter++;
      if (counter > 2) break;
    }
  if (counter >= 2) return;
  if (are_equal2(ti.X[0], tj.X[0])) { vertex[0] = 1; one = true; one2 = true; }
  else if (are_equal2(ti.X[0], tj.X[1])) { vertex[0] = 2; one = true; two2 = true; }
  else if (are_equal2(ti.X[0], tj.X[2])) { vertex[0] = 3; one = true; three2 = true; }
  if (are_equal2(ti.X[1], tj.X[0])) { vertex[1] = 2; two = true; one2 = true; }
  else if (are_equal2(ti.X[1], tj.X[1])) { vertex[1] = 5; two = true; two2 = true; }
  else if (are_equal2(ti.X[1], tj.X[2])) { vertex[1] = 6; two = true; three2 = true; }
  if (are_equal2(ti.X[2], tj.X[0])) { vertex[2] = 3; three = true; one2 = true; }
  else if (are_equal2(ti.X[2], tj.X[1])) { vertex[2] = 6; three = true; two2 = true; }
  else if (are_equal2(ti.X[2], tj.X[2])) { vertex[2] = 9; three = true; three2 = true; }

  if (one && three) {
    ti.neig[0] = j + 1; no_neig++;
    if (check_vector(vertex, 1) || (check_vector(vertex, 2) && check_vector(vertex, 9))) ti.dir[0] = 1;
    three = false;
  } else if (one && two) {
    ti.neig[1] = j + 1; no_neig++;
    if (check_vector(vertex, 1) || (check_vector(vertex, 6) && check_vector(vertex, 2))) ti.dir[1] = 1;
    one = false;
  } else if (three && two) {
    ti.neig[2] = j + 1; no_neig++;
    if (check_vector(vertex, 9) || (check_vector(vertex, 6) && check_vector(vertex, 2))) ti.dir[2] = 1;
    two = false;
  }
  auto copy_dir = [&](int jf) {
    if (one) tj.dir[jf] = ti.dir[0];
    else if (two) tj.dir[jf] = ti.dir[1];
    else if (three) tj.dir[jf] = ti.dir[2];
  };
  if (one2 && three2) { tj.neig[0] = i + 1; copy_dir(0); }
  else if (one2 && two2) { tj.neig[1] = i + 1; copy_dir(1); }
  else if (three2 && two2) { tj.neig[2] = i + 1; copy_dir(2); }
}

// Msh2Tri.F90:173-330.  gmsh 2.2 ASCII; triangle types {2,9,20,21,23,24,25}; first tag = region.
int read_msh(const char* path, std::vector<Tri>& out) {
  std::ifstream f(path);
  if (!f) return -1;
  std::string line;
  if (!std::getline(f, line)) return -2;
  while (!line.empty() && (line.back() == '\r' || line.back() == ' ')) line.pop_back();
  if (line != "$MeshFormat") return -2;
  std::getline(f, line);
  { std::istringstream is(line); double ver; int binary; is >> ver >> binary; if (binary != 0) return -3; }
  auto trim = [](std::string& s) { while (!s.empty() && (s.back() == '\r' || s.back() == ' ')) s.pop_back(); };
  while (std::getline(f, line)) { trim(line); if (line == "$Nodes") break; }
  int nodes = 0;
  std::getline(f, line); nodes = std::atoi(line.c_str());
  if (nodes <= 0) return -4;
  std::vector<double> vx((size_t)nodes + 1), vy((size_t)nodes + 1);
  for (int i = 0; i < nodes; ++i) {
    std::getline(f, line);
    std::istringstream is(line);
    int id; double x, y, z;
    is >> id >> x >> y >> z;
    if (!is || id < 1 || id > nodes) return -4;
    vx[id] = x; vy[id] = y;
  }
  while (std::getline(f, line)) { trim(line); if (line == "$Elements") break; }
  std::getline(f, line);
  int nel = std::atoi(line.c_str());
  if (nel <= 0) return -5;
  std::vector<Tri> ml2((size_t)nel + 1);
  int j = 0;
  double l_d = 0.0;
  for (int i = 1; i <= nel; ++i) {
    std::getline(f, line);
    std::istringstream is(line);
    std::vector<long> tok; long v;
    while (is >> v) tok.push_back(v);
    if (tok.size() < 3) return -5;
    int pos = (int)tok[0], type = (int)tok[1];
    if (!(type == 23 || type == 21 || type == 20 || type == 9 || type == 2 || type == 24 || type == 25)) { j++; continue; }
    int ntags = (int)tok[2];
    if ((int)tok.size() < 6 + ntags || pos < 1 || pos > nel) return -5;
    int region = (int)tok[3];
    int xp[3] = {(int)tok[3 + ntags], (int)tok[4 + ntags], (int)tok[5 + ntags]};
    Tri& t = ml2[pos];
    t.region = region;
    for (int a = 0; a < 3; ++a) { t.X[a][0] = vx[xp[a]]; t.X[a][1] = vy[xp[a]]; }
    double d1 = get_length(t.X[0], t.X[2]), d2 = get_length(t.X[0], t.X[1]), d3 = get_length(t.X[1], t.X[2]);
    l_d = std::max(l_d, std::max(d1, std::max(d2, d3)));
  }
  out.clear();
  for (int i = j + 1; i <= nel; ++i) {
    Tri t = ml2[i];
    for (int a = 0; a < 3; ++a) { t.neig[a] = 0; t.dir[a] = 0; }
    out.push_back(t);
  }
  // all-pairs neighbour search, :323-330
  int N = (int)out.size();
  for (int i = 0; i < N; ++i) {
    int no_neig = 0;
    for (int jj = i + 1; jj < N; ++jj) {
      check_neig(out, i, jj, no_neig, l_d);
      if (no_neig == 3) break;
    }
  }
  return N;
}

// Msh2Tri.F90:454-548
void neig_data(const int* neig, const int* dir, int U, int mpos /*1-based*/, int side, int& npos, int& nside, int nnodes[2]) {
  (void)U;
  nside = 0;
  npos = neig[(mpos - 1) * 3 + (side - 1)];
  const int T1[3][2] = {{1, 3}, {1, 2}, {2, 3}};
  if (npos != 0) {
    nside = 0;
    for (int q = 0; q < 3; ++q) if (neig[(npos - 1) * 3 + q] == mpos) { nside = q + 1; break; }  // NumLoc
    if (nside == 0) { nnodes[0] = nnodes[1] = 0; return; }
    if (dir[(mpos - 1) * 3 + (side - 1)]) { nnodes[0] = T1[nside - 1][0]; nnodes[1] = T1[nside - 1][1]; }
    else { nnodes[0] = T1[nside - 1][1]; nnodes[1] = T1[nside - 1][0]; }
  } else {
    nnodes[0] = T1[side - 1][0]; nnodes[1] = T1[side - 1][1];
  }
}

}  // namespace

// =============================================================================
// semi-structured multigrid problem
// =============================================================================
struct orc_semi {
  orc_params p;
  int U;
  std::vector<double> X;        // [u][node][dim]
  std::vector<int> neig, fneig, dir;
  struct Level {
    int s, C, S;                // split, children per parent, boundary children per face
    std::vector<double> tnew, told, rhs, res, src, tnonlin;
    std::vector<double> ovl, ovl_old;  // [u][face 0..2][3*S]
    std::vector<int> str_neig, surf_ele;
    // per parent scaling_var(ilevel): detwei[3], nx[3][2][3], sdetwei per CHILD face [2][3]
    std::vector<double> detwei, nx, sdetwei;
  };
  std::vector<Level> lev;
  // per parent: snorm in child-face order [s][d][f], dc_unele[3] (gmsh faces), dc_str[3] (child faces, finest), center[2]
  std::vector<double> snorm, dc_unele, dc_str, center;
  // geometric halo maps per parent gmsh face: reversed flag for the writer, node map for the reader
  std::vector<int> halo_rev;    // [u][mface]: 1 -> slot S-p+1, 0 -> slot p
  std::vector<int> halo_node;   // [u][mface][2]: strip entry (0..2) coincident with my face nodes a, b
  // Dirichlet data of domain-boundary parent faces (update_overlaps carries a t_bc argument, splitting.F90:1210, that HEAD
  // overwrites with boundary(x,y) = sin(x+y) at :1246-1252).  kind 0: sin(x+y) (HEAD); 1: the constant bc_val; 2: no data
  // (open face: no penalty term, the exterior trace equals the interior one)
  std::vector<int> bc_kind;     // [u][mface]
  std::vector<double> bc_val;   // [u][mface]
};

static int g_threads = 1;
void orc_semi_set_threads(int n) { g_threads = n < 1 ? 1 : n; }

namespace {

inline const double (*Xof(const orc_semi* h, int u))[2] {
  return reinterpret_cast<const double (*)[2]>(&h->X[(size_t)u * 6]);
}

// the 12-case table of splitting.F90:1256-1391, reduced to "reversed or not"
inline int literal_reversed(int mface, int nside, int dir) {
  if (mface == 2) return (nside == 2) ? (dir ? 0 : 1) : (dir ? 1 : 0);
  return (nside == 2) ? (dir ? 1 : 0) : (dir ? 0 : 1);
}

// nodes of an up child lying on parent gmsh face mf (1..3): Msh2Tri.F90:877-901 side numbering
const int SIDE_NODES[3][2] = {{1, 3}, {1, 2}, {2, 3}};

void build_halo_maps(orc_semi* h) {
  int U = h->U;
  h->halo_rev.assign((size_t)U * 3, 0);
  h->halo_node.assign((size_t)U * 6, 0);
  for (int u = 0; u < U; ++u)
    for (int mf = 1; mf <= 3; ++mf) {
      int q = h->neig[u * 3 + mf - 1];
      int* node = &h->halo_node[(size_t)(u * 3 + mf - 1) * 2];
      // my child face on this parent face and its nodes (a,b)
      int cf = (mf == 1) ? 1 : (mf == 2 ? 3 : 2);
      int a = FACE_NODES[cf - 1][0], b = FACE_NODES[cf - 1][1];
      if (q == 0) { node[0] = a - 1; node[1] = b - 1; continue; }  // Dirichlet entries live at my own node ids
      int ns = h->fneig[u * 3 + mf - 1];
      const double (*Xm)[2] = Xof(h, u);
      const double (*Xn)[2] = Xof(h, q - 1);
      // geometric pairing of parent vertices
      int na = -1, nb = -1;
      for (int c = 0; c < 2; ++c) {
        int nn = SIDE_NODES[ns - 1][c];
        if (are_equal2(Xn[nn - 1], Xm[a - 1])) na = nn;
        if (are_equal2(Xn[nn - 1], Xm[b - 1])) nb = nn;
      }
      node[0] = na - 1; node[1] = nb - 1;
      // geometric reversal: my strip positions run X3->X1 (mf 1), X1->X2 (mf 2), X3->X2 (mf 3);
      // the neighbour's positions on its side ns run the same way in ITS numbering.
      const int START[3] = {3, 1, 3};  // first vertex of the run on each parent face
      int my_start = START[mf - 1], nb_start = START[ns - 1];
      bool same = are_equal2(Xm[my_start - 1], Xn[nb_start - 1]);
      int geo_rev = same ? 0 : 1;
      h->halo_rev[u * 3 + mf - 1] =
          h->p.halo_rule == 0 ? literal_reversed(mf, ns, h->dir[u * 3 + mf - 1]) : geo_rev;
    }
}

// semi_tri_det_nlx_multigrid / semi_det_snlx_multigrid (ShapFun.F90:1661-1684,1737-1783) and
// get_d_center (Msh2Tri.F90:349-385)
void build_geometry(orc_semi* h) {
  int U = h->U, n = h->p.n_split;
  h->snorm.assign((size_t)U * 12, 0);
  h->dc_unele.assign((size_t)U * 3, 0);
  h->dc_str.assign((size_t)U * 3, 0);
  h->center.assign((size_t)U * 2, 0);
  std::vector<int> sn_fine;
  str_neig_table(n, sn_fine);
  for (int u = 0; u < U; ++u) {
    const double (*X)[2] = Xof(h, u);
    double nx[3][2][3], detwei[3];
    tri_det_nlx(X, nx, detwei);
    double sdet_g[3][2], snorm_g[3][2][2];  // gmsh face order
    for (int f = 1; f <= 3; ++f) face_geometry(X, f, sdet_g[f - 1], snorm_g[f - 1]);
    for (size_t il = 0; il < h->lev.size(); ++il) {
      orc_semi::Level& L = h->lev[il];
      double a4 = (double)ipow(4, L.s), a2 = (double)ipow(2, L.s);
      for (int g = 0; g < 3; ++g) {
        L.detwei[(size_t)u * 3 + g] = detwei[g] / a4;
        for (int d = 0; d < 2; ++d)
          for (int i = 0; i < 3; ++i) L.nx[(size_t)u * 18 + (g * 2 + d) * 3 + i] = nx[g][d][i] * a2;
      }
      // child face f uses the length of parent face MFACE[f] (INTENDED; HEAD leaves scaling_var%sdetwei
      // in parent order, SURVEY B-10, but never reads it because the face block is commented out)
      for (int cf = 0; cf < 3; ++cf)
        for (int s = 0; s < 2; ++s) L.sdetwei[(size_t)u * 6 + s * 3 + cf] = sdet_g[MFACE[cf] - 1][s] / a2;
    }
    // snorm swapped to child-face order (:1774-1776)
    for (int cf = 0; cf < 3; ++cf)
      for (int s = 0; s < 2; ++s)
        for (int d = 0; d < 2; ++d) h->snorm[(size_t)u * 12 + (s * 2 + d) * 3 + cf] = snorm_g[MFACE[cf] - 1][s][d];
    // centres and centroid distances
    double cx = (X[0][0] + X[1][0] + X[2][0]) / 3, cy = (X[0][1] + X[1][1] + X[2][1]) / 3;
    h->center[u * 2] = cx; h->center[u * 2 + 1] = cy;
    for (int mf = 1; mf <= 3; ++mf) {
      int q = h->neig[u * 3 + mf - 1];
      if (q != 0) {
        const double (*Y)[2] = Xof(h, q - 1);
        double qx = (Y[0][0] + Y[1][0] + Y[2][0]) / 3, qy = (Y[0][1] + Y[1][1] + Y[2][1]) / 3;
        h->dc_unele[u * 3 + mf - 1] = std::sqrt((cx - qx) * (cx - qx) + (cy - qy) * (cy - qy));
      } else {
        // INTENDED boundary length: centre -> midpoint of the parent edge (matrices.F90:104-109 with the
        // edge of mface; the HEAD get_d_center branch :365-370 leaves c1(2) unset)
        int a = SIDE_NODES[mf - 1][0] - 1, b = SIDE_NODES[mf - 1][1] - 1;
        double mx = (X[a][0] + X[b][0]) / 2, my = (X[a][1] + X[b][1]) / 2;
        h->dc_unele[u * 3 + mf - 1] = std::sqrt((cx - mx) * (cx - mx) + (cy - my) * (cy - my));
      }
    }
    // dc_str_ele: child 2 and its three neighbours at the finest split (:377-383)
    double x1[3][2], x2[3][2];
    get_splitting(X, n, 2, x1);
    double c1x = (x1[0][0] + x1[1][0] + x1[2][0]) / 3, c1y = (x1[0][1] + x1[1][1] + x1[2][1]) / 3;
    for (int f = 0; f < 3; ++f) {
      int nb = sn_fine[(size_t)(2 - 1) * 3 + f];
      get_splitting(X, n, nb, x2);
      double c2x = (x2[0][0] + x2[1][0] + x2[2][0]) / 3, c2y = (x2[0][1] + x2[1][1] + x2[2][1]) / 3;
      h->dc_str[u * 3 + f] = std::sqrt((c1x - c2x) * (c1x - c2x) + (c1y - c2y) * (c1y - c2y));
    }
  }
}

// per-parent stencils, ShapFun_unstruc.F90:324-335
struct Stencil {
  double ml[3], mass[3][3], stiff[3][3][2][3] /* [iloc][g][d][jloc] */, dvol[3][2][3][3] /* [g][d][i][j] */;
};
void parent_stencil(const double* detwei, const double* nxp, double k, Stencil& st) {
  auto nx = [&](int g, int d, int i) { return nxp[(g * 2 + d) * 3 + i]; };
  for (int j = 0; j < 3; ++j) {
    st.ml[j] = 0;
    for (int g = 0; g < 3; ++g) st.ml[j] += TB.n[g][j] * detwei[g];
    for (int i = 0; i < 3; ++i) {
      double m = 0;
      for (int g = 0; g < 3; ++g) m += TB.n[g][i] * detwei[g] * TB.n[g][j];
      st.mass[i][j] = m;
      for (int d = 0; d < 2; ++d)
        for (int g = 0; g < 3; ++g) {
          st.stiff[i][g][d][j] = nx(g, d, j) * detwei[g] * TB.n[g][i];
          st.dvol[g][d][i][j] = k * nx(g, d, i) * detwei[g] * nx(g, d, j);
        }
    }
  }
}

struct ElemOp {            // everything get_A_x / get_diagonal need for one child
  double stiff1[3][3], dvol1[3][3];
  double flux[3], dsurf[3], mydiag[3];
};

// One child: volume terms (:575-609) and, if face_terms, the face block (:619-688 read with
// transport_tri.F90:593-669 / transport_tri_unstr.F90:706-747).  Tcur: field the face traces and
// neighbour values are taken from; Town: own nodal values used for traces.
// ovl_field: the halo strips that go with Tnbr_field (nullptr = L.ovl, the strips of tracer%tnew; L.ovl_old for told).
void element_terms(const orc_semi* h, const orc_semi::Level& L, int level, int u, int ele,
                   const Stencil& st, const double* Town, const double* Tnbr_field, ElemOp& op,
                   const std::vector<double>* ovl_field = nullptr) {
  int irow, ipos, ori;
  get_str_info(L.s, ele, irow, ipos, ori);
  int updown = ori == 0 ? -1 : 1;  // semi_get_nx_pos, ShapFun.F90:1801-1804
  double uloc[2][3] = {{h->p.u_x, h->p.u_x, h->p.u_x}, {h->p.u_y, h->p.u_y, h->p.u_y}};
  double ugi[3][2];
  for (int g = 0; g < 3; ++g)
    for (int d = 0; d < 2; ++d) {
      double s = 0;
      for (int i = 0; i < 3; ++i) s += TB.n[g][i] * uloc[d][i];
      ugi[g][d] = s;
    }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s1 = 0, s2 = 0;
      for (int g = 0; g < 3; ++g) { s1 += st.stiff[j][g][0][i] * ugi[g][0]; s2 += st.stiff[j][g][1][i] * ugi[g][1]; }
      op.stiff1[i][j] = (s1 + s2) * updown;
      double dv = 0;
      for (int d = 0; d < 2; ++d)
        for (int g = 0; g < 3; ++g) dv += st.dvol[g][d][i][j];
      op.dvol1[i][j] = dv;
    }
  for (int i = 0; i < 3; ++i) { op.flux[i] = 0; op.dsurf[i] = 0; op.mydiag[i] = 0; }
  if (!h->p.face_terms) return;

  const double* ovl = &(ovl_field ? *ovl_field : L.ovl)[(size_t)u * 3 * 3 * L.S];
  double lvl_scale = (double)ipow(2, level - 1);  // penalty length is level aware (INTENDED, SURVEY B-18)
  for (int f = 1; f <= 3; ++f) {
    int ele22 = L.str_neig[(size_t)(ele - 1) * 3 + f - 1];
    int a = FACE_NODES[f - 1][0] - 1, b = FACE_NODES[f - 1][1] - 1;
    double T2a, T2b, delta_x;
    int mface = MFACE[f - 1];
    bool open_face = false;
    if (ele22 != 0) {
      const double* Tn = &Tnbr_field[((size_t)u * L.C + (ele22 - 1)) * 3];
      T2a = Tn[b]; T2b = Tn[a];  // shared nodes appear reversed on the other side (transport_tri.F90:610-611)
      delta_x = h->dc_str[u * 3 + f - 1] * lvl_scale;
    } else if (h->neig[u * 3 + mface - 1] == 0 && h->bc_kind[u * 3 + mface - 1] == 2) {
      T2a = Town[a]; T2b = Town[b];   // open boundary face: exterior trace = interior trace, no penalty
      delta_x = 1.0; open_face = true;
    } else {
      int sp = (f == 1) ? ipos / 2 + 1 : irow;   // :629-638
      const double* e = &ovl[(size_t)(mface - 1) * 3 * L.S + (size_t)(sp - 1) * 3];
      const int* nm = &h->halo_node[(size_t)(u * 3 + mface - 1) * 2];
      T2a = e[nm[0]]; T2b = e[nm[1]];
      delta_x = h->dc_unele[u * 3 + mface - 1] / (double)ipow(2, L.s);  // matrices.F90:101-109
    }
    double snorm[2][2], sdet[2], Ts[2], T2s[2], income[2];
    for (int s = 0; s < 2; ++s) {
      for (int d = 0; d < 2; ++d) snorm[s][d] = updown * h->snorm[(size_t)u * 12 + (s * 2 + d) * 3 + f - 1];
      sdet[s] = L.sdetwei[(size_t)u * 6 + s * 3 + f - 1];
      Ts[s] = TB.sn[s][0] * Town[a] + TB.sn[s][1] * Town[b];
      T2s[s] = TB.sn[s][0] * T2a + TB.sn[s][1] * T2b;
      double un = snorm[s][0] * 0.5 * (h->p.u_x + h->p.u_x) + snorm[s][1] * 0.5 * (h->p.u_y + h->p.u_y);
      income[s] = 0.5 + 0.5 * std::copysign(1.0, -un);
    }
    double uu[2] = {h->p.u_x, h->p.u_y};
    for (int c = 0; c < 2; ++c) {
      int q = FACE_NODES[f - 1][c] - 1;
      for (int d = 0; d < 2; ++d)
        for (int s = 0; s < 2; ++s)
          op.flux[q] += TB.sn[s][c] * snorm[s][d] * sdet[s] *
                        ((1.0 - income[s]) * uu[d] * Ts[s] + income[s] * uu[d] * T2s[s]);
      double kd = open_face ? 0.0 : h->p.k / delta_x;
      for (int s = 0; s < 2; ++s) {
        op.dsurf[q] += kd * TB.sn[s][c] * sdet[s] * (Ts[s] - T2s[s]);   // matrices.F90:113-115 form
        op.mydiag[q] += kd * TB.sn[s][c] * TB.sn[s][c] * sdet[s];      // my_diff_surf(i,i,f), :471-472
      }
    }
  }
}

// get_A_x (:412-448) for one child acting on T
void apply_A(const orc_semi* h, const Stencil& st, const ElemOp& op, const double* T, double* Ax) {
  double th = h->p.theta;
  for (int i = 0; i < 3; ++i) {
    double mass_new = 0, dvol = 0, stiff = 0;
    for (int j = 0; j < 3; ++j) {
      mass_new += st.mass[i][j] * T[j];
      dvol += op.dvol1[i][j] * T[j];
      stiff += op.stiff1[i][j] * T[j];
    }
    mass_new *= 1.0 / h->p.dt;
    Ax[i] = th * (mass_new - stiff + op.flux[i] + dvol + op.dsurf[i]) + (1.0 - th) * mass_new;
  }
}

// get_diagonal (:481-486).  For theta != 1 it is the diagonal of the operator get_A_x applies, ml/dt + theta (K_ii + sum_f my_ii);
// the reference's mat_diag_approx leaves theta out (:483-484), which is the same thing at theta = 1, its only literal (:117).
// solve_Richardson (:516) has no theta either and is weighted the same way here.
inline double diag_of(const orc_semi* h, const Stencil& st, const ElemOp& op, int i) {
  return st.ml[i] / h->p.dt + h->p.theta * op.dvol1[i][i] + h->p.theta * op.mydiag[i];   // (same rounding as before at theta = 1)
}

// source reset (:593) + get_RHS (:452-464) for one level-1 child
void build_rhs_child(orc_semi* h, orc_semi::Level& L, int u, int ele, const Stencil& st) {
  const double (*X)[2] = Xof(h, u);
  double x[3][2];
  get_splitting(X, L.s, ele, x);
  size_t o = ((size_t)u * L.C + (ele - 1)) * 3;
  double* src = &L.src[o];
  const double* told = &L.told[o];
  double s0[3];
  for (int i = 0; i < 3; ++i) { src[i] = h->p.source_coef * boundary_fn(x[i][0], x[i][1]); s0[i] = src[i]; }
  for (int i = 0; i < 3; ++i) {
    double m_old = 0;
    for (int j = 0; j < 3; ++j) m_old += st.mass[i][j] * told[j];
    m_old *= 1.0 / h->p.dt;
    double ms = 0;
    if (h->p.literal_source) { for (int j = 0; j < 3; ++j) ms += st.mass[i][j] * src[j]; src[i] = ms; }
    else { for (int j = 0; j < 3; ++j) ms += st.mass[i][j] * s0[j]; }
    L.rhs[o + i] = m_old + ms;   // theta = 1 form of :458 (the reference's only literal, :117)
  }
  if (h->p.theta != 1.0) {
    // (1-theta) branch of :459-460: + (1-theta)(stiff_old - flux_old - diff_vol - diff_surf), every spatial term evaluated
    // on TOLD with the told strips (t_overlap_old, splitting.F90:1259-1262).  HEAD has no flux_ele_old (the face block is
    // commented out) and takes diff_vol / diff_surf of the NEW iterate there; theta = 1 is its only literal, so this branch
    // is the Crank-Nicolson reading, not a quirk that any run of the reference exercises.
    ElemOp op;
    element_terms(h, L, 1, u, ele, st, told, L.told.data(), op, &L.ovl_old);
    for (int i = 0; i < 3; ++i) {
      double stiff = 0, dvol = 0;
      for (int j = 0; j < 3; ++j) { stiff += op.stiff1[i][j] * told[j]; dvol += op.dvol1[i][j] * told[j]; }
      L.rhs[o + i] += (1.0 - h->p.theta) * (stiff - op.flux[i] - dvol - op.dsurf[i]);
    }
  }
  if (!h->p.literal_source) for (int i = 0; i < 3; ++i) {
    double ms = 0; for (int j = 0; j < 3; ++j) ms += st.mass[i][j] * s0[j]; src[i] = ms; }
}

void stencil_for(const orc_semi* h, const orc_semi::Level& L, int u, Stencil& st) {
  parent_stencil(&L.detwei[(size_t)u * 3], &L.nx[(size_t)u * 18], h->p.k, st);
}

}  // namespace

// ------------------------------------------------------------------ public API
extern "C" {

void orc_tables(double* n9, double* nlx18, double* w3, double* sn4, double* snlx4, double* sw2) {
  for (int g = 0; g < 3; ++g) {
    w3[g] = TB.w[g];
    for (int i = 0; i < 3; ++i) n9[g * 3 + i] = TB.n[g][i];
    for (int d = 0; d < 2; ++d)
      for (int i = 0; i < 3; ++i) nlx18[(g * 2 + d) * 3 + i] = TB.nlx[g][d][i];
  }
  for (int s = 0; s < 2; ++s) {
    sw2[s] = TB.sw[s];
    for (int c = 0; c < 2; ++c) { sn4[s * 2 + c] = TB.sn[s][c]; snlx4[s * 2 + c] = TB.snlx[s][c]; }
  }
}
void orc_get_str_info(int n, int ele, int* irow, int* ipos, int* ori) { get_str_info(n, ele, *irow, *ipos, *ori); }
void orc_get_splitting(const double* X6, int n, int ele, double* x6) {
  get_splitting(reinterpret_cast<const double (*)[2]>(X6), n, ele, reinterpret_cast<double (*)[2]>(x6));
}
void orc_str_neig(int n, int32_t* out) { std::vector<int> t; str_neig_table(n, t); for (size_t i = 0; i < t.size(); ++i) out[i] = t[i]; }
void orc_surf_ele(int n, int32_t* out) { std::vector<int> t; surf_ele_table(n, t); for (size_t i = 0; i < t.size(); ++i) out[i] = t[i]; }
void orc_element_conversion(int c, int s, int32_t* fin4) { int f[4]; element_conversion(c, s, f); for (int i = 0; i < 4; ++i) fin4[i] = f[i]; }
void orc_tri_det_nlx(const double* x6, double* nx18, double* detwei3) {
  double nx[3][2][3];
  tri_det_nlx(reinterpret_cast<const double (*)[2]>(x6), nx, detwei3);
  for (int g = 0; g < 3; ++g) for (int d = 0; d < 2; ++d) for (int i = 0; i < 3; ++i) nx18[(g * 2 + d) * 3 + i] = nx[g][d][i];
}
void orc_face_geometry(const double* x6, int iface, double* sdetwei2, double* snorm4) {
  double sn[2][2];
  face_geometry(reinterpret_cast<const double (*)[2]>(x6), iface, sdetwei2, sn);
  for (int s = 0; s < 2; ++s) for (int d = 0; d < 2; ++d) snorm4[s * 2 + d] = sn[s][d];
}
int orc_findinv(const double* A, double* Ainv, int n) { return findinv(A, Ainv, n); }

int orc_read_msh(const char* path, int max_tri, double* X, int32_t* neig, int32_t* dir, int32_t* region) {
  std::vector<Tri> ml;
  int N = read_msh(path, ml);
  if (N < 0) return N;
  if (N > max_tri) return -10;
  for (int u = 0; u < N; ++u) {
    for (int a = 0; a < 3; ++a) {
      X[(size_t)u * 6 + a * 2] = ml[u].X[a][0]; X[(size_t)u * 6 + a * 2 + 1] = ml[u].X[a][1];
      neig[u * 3 + a] = ml[u].neig[a]; dir[u * 3 + a] = ml[u].dir[a];
    }
    region[u] = ml[u].region;
  }
  return N;
}
void orc_neig_data(int U, const int32_t* neig, const int32_t* dir, int32_t* fneig, int32_t* snodes) {
  for (int u = 1; u <= U; ++u)
    for (int f = 1; f <= 3; ++f) {
      int npos, nside, nn[2];
      neig_data(neig, dir, U, u, f, npos, nside, nn);
      fneig[(u - 1) * 3 + f - 1] = nside;
      snodes[((u - 1) * 3 + f - 1) * 2] = nn[0]; snodes[((u - 1) * 3 + f - 1) * 2 + 1] = nn[1];
    }
}

orc_semi* orc_semi_create(const orc_params* p, int U, const double* X, const int32_t* neig,
                          const int32_t* fneig, const int32_t* dir) {
  if (p->multi_levels > p->n_split || p->multi_levels < 1) return nullptr;  // transport_tri_semi.F90:120-123
  orc_semi* h = new orc_semi();
  h->p = *p; h->U = U;
  h->X.assign(X, X + (size_t)U * 6);
  h->neig.assign(neig, neig + (size_t)U * 3);
  h->fneig.assign(fneig, fneig + (size_t)U * 3);
  h->dir.assign(dir, dir + (size_t)U * 3);
  h->lev.resize(p->multi_levels);
  for (int il = 0; il < p->multi_levels; ++il) {
    orc_semi::Level& L = h->lev[il];
    L.s = p->n_split - il;  // i_split = n_split - ilevel + 1
    L.C = ipow(4, L.s); L.S = ipow(2, L.s);
    size_t nd = (size_t)3 * L.C * U;
    L.tnew.assign(nd, 0); L.told.assign(nd, 0); L.rhs.assign(nd, 0); L.res.assign(nd, 0);
    L.src.assign(nd, 0); L.tnonlin.assign(nd, 0);
    L.ovl.assign((size_t)U * 9 * L.S, 0); L.ovl_old.assign((size_t)U * 9 * L.S, 0);
    str_neig_table(L.s, L.str_neig);
    surf_ele_table(L.s, L.surf_ele);
    L.detwei.assign((size_t)U * 3, 0); L.nx.assign((size_t)U * 18, 0); L.sdetwei.assign((size_t)U * 6, 0);
  }
  build_geometry(h);
  build_halo_maps(h);
  h->bc_kind.assign((size_t)U * 3, 0);
  h->bc_val.assign((size_t)U * 3, 0.0);
  return h;
}
void orc_semi_set_boundary(orc_semi* h, const int32_t* kind, const double* value) {
  for (size_t i = 0; i < (size_t)h->U * 3; ++i) { h->bc_kind[i] = kind[i]; h->bc_val[i] = value ? value[i] : 0.0; }
  for (auto& L : h->lev) { std::fill(L.ovl.begin(), L.ovl.end(), 0.0); std::fill(L.ovl_old.begin(), L.ovl_old.end(), 0.0); }
}
void orc_semi_destroy(orc_semi* h) { delete h; }
int64_t orc_semi_ndof(orc_semi* h, int level) { return (int64_t)3 * h->lev[level - 1].C * h->U; }
double* orc_semi_field(orc_semi* h, int field, int level) {
  orc_semi::Level& L = h->lev[level - 1];
  switch (field) {
    case ORC_TNEW: return L.tnew.data();
    case ORC_TOLD: return L.told.data();
    case ORC_RHS: return L.rhs.data();
    case ORC_RES: return L.res.data();
    case ORC_SRC: return L.src.data();
    case ORC_TNONLIN: return L.tnonlin.data();
  }
  return nullptr;
}
double* orc_semi_overlap(orc_semi* h, int level, int old) {
  return old ? h->lev[level - 1].ovl_old.data() : h->lev[level - 1].ovl.data();
}

// update_overlaps, splitting.F90:1238-1394.  Reads tracer%tnew / told of the level.
static void update_overlaps_impl(orc_semi* h, int level, bool do_new, bool do_old) {
  orc_semi::Level& L = h->lev[level - 1];
  int S = L.S;
  double bc_scale = (h->p.coarse_bc_zero && level > 1) ? 0.0 : 1.0;
  const int order[3] = {1, 3, 2};  // faces are visited 1, 3, 2
  for (int u = 0; u < h->U; ++u) {
    const double (*X)[2] = Xof(h, u);
    for (int oi = 0; oi < 3; ++oi) {
      int mf = order[oi];
      for (int i = 1; i <= S; ++i) {
        int ele = L.surf_ele[(size_t)(mf - 1) * S + i - 1];
        int irow, ipos, ori;
        get_str_info(L.s, ele, irow, ipos, ori);
        int pos = (mf == 1) ? ipos / 2 + 1 : irow;
        int npos = h->neig[u * 3 + mf - 1];
        if (npos == 0) {
          const int kind = h->bc_kind[u * 3 + mf - 1];
          if (kind == 2) continue;                                           // open face: no Dirichlet data
          double x[3][2];
          get_splitting(X, L.s, ele, x);
          int a = SIDE_NODES[mf - 1][0] - 1, b = SIDE_NODES[mf - 1][1] - 1;  // :1246-1249,1287-1290,1344-1347
          double ta = bc_scale * boundary_fn(x[a][0], x[a][1]), tb = bc_scale * boundary_fn(x[b][0], x[b][1]);
          if (kind == 1) ta = tb = bc_scale * h->bc_val[u * 3 + mf - 1];      // t_bc as data (splitting.F90:1210)
          size_t o = ((size_t)u * 3 + (mf - 1)) * 3 * S + (size_t)(pos - 1) * 3;
          if (do_new) { L.ovl[o + a] = ta; L.ovl[o + b] = tb; }
          if (do_old) { L.ovl_old[o + a] = ta; L.ovl_old[o + b] = tb; }
        } else {
          int nside = h->fneig[u * 3 + mf - 1];
          int rev = h->halo_rev[u * 3 + mf - 1];
          int slot = rev ? (S - pos + 1) : pos;
          size_t o = ((size_t)(npos - 1) * 3 + (nside - 1)) * 3 * S + (size_t)(slot - 1) * 3;
          size_t src = ((size_t)u * L.C + (ele - 1)) * 3;
          for (int q = 0; q < 3; ++q) {
            if (do_new) L.ovl[o + q] = L.tnew[src + q];
            if (do_old) L.ovl_old[o + q] = L.told[src + q];
          }
        }
      }
    }
  }
}
void orc_semi_update_overlaps(orc_semi* h, int level) { update_overlaps_impl(h, level, true, true); }

void orc_semi_build_rhs(orc_semi* h) {
  orc_semi::Level& L = h->lev[0];
  // theta != 1: the old-time face terms read the told strips; refresh them from TOLD (the tnew strips stay as the caller left them)
  if (h->p.theta != 1.0 && h->p.face_terms) update_overlaps_impl(h, 1, false, true);
#pragma omp parallel for num_threads(g_threads) schedule(static)
  for (int u = 0; u < h->U; ++u) {
    Stencil st; stencil_for(h, L, u, st);
    for (int ele = 1; ele <= L.C; ++ele) build_rhs_child(h, L, u, ele, st);
  }
}

// smoother, transport_tri_semi.F90:543-722.  Operates on tnonlin (the iterate) exactly like the
// reference: each sweep starts with tnew <- tnonlin and a halo update from tnew.
void orc_semi_smooth(orc_semi* h, int level, int solver, int nsweeps) {
  orc_semi::Level& L = h->lev[level - 1];
  for (int sweep = 0; sweep < nsweeps; ++sweep) {
    L.tnew = L.tnonlin;                         // :550
    orc_semi_update_overlaps(h, level);          // :555
    if (level == 1) orc_semi_build_rhs(h);       // get_RHS is evaluated per child per sweep (:699-701); same values
    auto do_child = [&](int u, int ele, const Stencil& st, bool latest) {
      size_t o = ((size_t)u * L.C + (ele - 1)) * 3;
      const double* Town = latest ? &L.tnonlin[o] : &L.tnew[o];
      const double* nbr = latest ? L.tnonlin.data() : L.tnew.data();
      ElemOp op;
      element_terms(h, L, level, u, ele, st, Town, nbr, op);
      double Ax[3];
      apply_A(h, st, op, Town, Ax);
      for (int i = 0; i < 3; ++i) {
        if (solver == 2) {
          // solve_Richardson :511-518 : omega*(b - (mass - stiff + flux))
          double m = 0, sfn = 0;
          for (int j = 0; j < 3; ++j) { m += st.mass[i][j] * Town[j]; sfn += op.stiff1[i][j] * Town[j]; }
          m /= h->p.dt;
          L.tnonlin[o + i] = L.tnonlin[o + i] + h->p.omega * (L.rhs[o + i] - (m - h->p.theta * sfn + h->p.theta * op.flux[i]));
        } else {
          double D = diag_of(h, st, op, i);
          double base = latest ? L.tnonlin[o + i] : L.tnew[o + i];
          Ax[i] = Ax[i];
          // solve_Jacobi :491-497 / solve_Gauss_Seidel :501-507 (point-simultaneous inside the child)
          L.tnonlin[o + i] = base + h->p.omega / D * (L.rhs[o + i] - Ax[i]);
        }
      }
    };
    if (solver == 1 || solver == 2) {
#pragma omp parallel for num_threads(g_threads) schedule(static)
      for (int u = 0; u < h->U; ++u) {
        Stencil st; stencil_for(h, L, u, st);
        for (int ele = 1; ele <= L.C; ++ele) do_child(u, ele, st, false);
      }
    } else if (solver == 3) {
      // reference order: parent-major, child-minor, latest values inside a parent, halo lagged
      for (int u = 0; u < h->U; ++u) {
        Stencil st; stencil_for(h, L, u, st);
        for (int ele = 1; ele <= L.C; ++ele) {
          // own values must be read before they are overwritten: copy
          size_t o = ((size_t)u * L.C + (ele - 1)) * 3;
          double own[3] = {L.tnonlin[o], L.tnonlin[o + 1], L.tnonlin[o + 2]};
          ElemOp op;
          element_terms(h, L, level, u, ele, st, own, L.tnonlin.data(), op);
          double Ax[3];
          apply_A(h, st, op, own, Ax);
          for (int i = 0; i < 3; ++i) {
            double D = diag_of(h, st, op, i);
            L.tnonlin[o + i] = own[i] + h->p.omega / D * (L.rhs[o + i] - Ax[i]);
          }
        }
      }
    } else {
      // solver 4: two-colour ordering of the same sweep (GPU ordering): all down children, then all up
      for (int colour = 0; colour < 2; ++colour) {
#pragma omp parallel for num_threads(g_threads) schedule(static)
        for (int u = 0; u < h->U; ++u) {
          Stencil st; stencil_for(h, L, u, st);
          for (int ele = 1; ele <= L.C; ++ele) {
            int irow, ipos, ori;
            get_str_info(L.s, ele, irow, ipos, ori);
            if (ori != colour) continue;  // colour 0 = down (orientation 0), colour 1 = up
            size_t o = ((size_t)u * L.C + (ele - 1)) * 3;
            double own[3] = {L.tnonlin[o], L.tnonlin[o + 1], L.tnonlin[o + 2]};
            ElemOp op;
            element_terms(h, L, level, u, ele, st, own, L.tnonlin.data(), op);
            double Ax[3];
            apply_A(h, st, op, own, Ax);
            for (int i = 0; i < 3; ++i) {
              double D = diag_of(h, st, op, i);
              L.tnonlin[o + i] = own[i] + h->p.omega / D * (L.rhs[o + i] - Ax[i]);
            }
          }
        }
      }
    }
  }
}

// get_residual :725-873 on tracer%tnew; halo strips as last updated by the caller.
void orc_semi_residual(orc_semi* h, int level, double* l2, double* linf) {
  orc_semi::Level& L = h->lev[level - 1];
  if (level == 1) orc_semi_build_rhs(h);
  double sum = 0, mx = 0;
#pragma omp parallel for num_threads(g_threads) schedule(static) reduction(+ : sum) reduction(max : mx)
  for (int u = 0; u < h->U; ++u) {
    Stencil st; stencil_for(h, L, u, st);
    for (int ele = 1; ele <= L.C; ++ele) {
      size_t o = ((size_t)u * L.C + (ele - 1)) * 3;
      ElemOp op;
      element_terms(h, L, level, u, ele, st, &L.tnew[o], L.tnew.data(), op);
      double Ax[3];
      apply_A(h, st, op, &L.tnew[o], Ax);
      for (int i = 0; i < 3; ++i) {
        double r = h->p.residual_sign * (Ax[i] - L.rhs[o + i]);  // :869
        L.res[o + i] = r;
        sum += r * r;
        mx = std::max(mx, std::fabs(r));
      }
    }
  }
  if (l2) *l2 = std::sqrt(sum);
  if (linf) *linf = mx;
}

// get_convergence :876-889 (signed max, starts from 0)
double orc_semi_convergence(orc_semi* h, int level) {
  orc_semi_residual(h, level, nullptr, nullptr);
  orc_semi::Level& L = h->lev[level - 1];
  double c = 0;
  for (size_t e = 0; e < L.res.size() / 3; ++e) {
    double loc = std::max(L.res[e * 3], std::max(L.res[e * 3 + 1], L.res[e * 3 + 2]));
    if (loc > c) c = loc;
  }
  return c;
}

// P1 interpolation weights from coarse nodes (c1,c2,c3) to the 12 fine nodes of fin(1..4),
// derived from splitting.F90:59-88: W[k][i][c]
static const double PW[4][3][3] = {
    /* fin1 (holds coarse node 3) */ {{0.5, 0, 0.5}, {0, 0.5, 0.5}, {0, 0, 1}},
    /* fin2 (central, inverted)   */ {{0, 0.5, 0.5}, {0.5, 0, 0.5}, {0.5, 0.5, 0}},
    /* fin3 (holds coarse node 1) */ {{1, 0, 0}, {0.5, 0.5, 0}, {0.5, 0, 0.5}},
    /* fin4 (holds coarse node 2) */ {{0.5, 0.5, 0}, {0, 1, 0}, {0, 0.5, 0.5}}};

void orc_semi_restrict(orc_semi* h, int fine_level) {
  if (fine_level >= h->p.multi_levels) return;  // splitting.F90:18
  orc_semi::Level& F = h->lev[fine_level - 1];
  orc_semi::Level& Cc = h->lev[fine_level];
#pragma omp parallel for num_threads(g_threads) schedule(static)
  for (int u = 0; u < h->U; ++u)
    for (int c = 1; c <= Cc.C; ++c) {
      int fin[4];
      element_conversion(c, F.s - 1, fin);
      size_t oc = ((size_t)u * Cc.C + (c - 1)) * 3;
      auto R = [&](int k, int i) { return F.res[((size_t)u * F.C + (fin[k] - 1)) * 3 + i]; };
      if (h->p.transfer == 0) {
        // splitting.F90:26-28 : mean over the 3 nodes of one fine child
        Cc.rhs[oc + 0] = (R(2, 0) + R(2, 1) + R(2, 2)) / 3.0;
        Cc.rhs[oc + 1] = (R(3, 0) + R(3, 1) + R(3, 2)) / 3.0;
        Cc.rhs[oc + 2] = (R(0, 0) + R(0, 1) + R(0, 2)) / 3.0;
      } else {
        for (int cn = 0; cn < 3; ++cn) {
          double s = 0;
          for (int k = 0; k < 4; ++k)
            for (int i = 0; i < 3; ++i) s += PW[k][i][cn] * R(k, i);
          Cc.rhs[oc + cn] = s;
        }
      }
    }
}

void orc_semi_prolong(orc_semi* h, int fine_level) {
  orc_semi::Level& F = h->lev[fine_level - 1];
  orc_semi::Level& Cc = h->lev[fine_level];
  // literal acts on tracer%tnew (splitting.F90:59-88); intended acts on the iterate (tnonlin)
  std::vector<double>& ft = h->p.transfer == 0 ? F.tnew : F.tnonlin;
  const std::vector<double>& ct = h->p.transfer == 0 ? Cc.tnew : Cc.tnonlin;
#pragma omp parallel for num_threads(g_threads) schedule(static)
  for (int u = 0; u < h->U; ++u)
    for (int c = 1; c <= Cc.C; ++c) {
      int fin[4];
      element_conversion(c, F.s - 1, fin);
      const double* cc = &ct[((size_t)u * Cc.C + (c - 1)) * 3];
      auto T = [&](int k, int i) -> double& { return ft[((size_t)u * F.C + (fin[k] - 1)) * 3 + i]; };
      if (h->p.transfer == 0) {
        T(0, 0) += 0.5 * cc[2] + 0.5 * cc[0];
        T(0, 1) += 0.5 * cc[1] + 0.5 * cc[2];
        T(0, 2) += cc[2];
        T(1, 0) += T(0, 1);
        T(1, 1) += T(0, 0);
        T(1, 2) += 0.5 * cc[0] + 0.5 * cc[1];
        T(2, 0) += cc[0];
        T(2, 1) += T(1, 2);
        T(2, 2) += T(1, 1);
        T(3, 0) += T(1, 2);
        T(3, 1) += cc[1];
        T(3, 2) += T(1, 0);
      } else {
        for (int k = 0; k < 4; ++k)
          for (int i = 0; i < 3; ++i) {
            double s = 0;
            for (int cn = 0; cn < 3; ++cn) s += PW[k][i][cn] * cc[cn];
            T(k, i) += s;
          }
      }
    }
}

static void vcycle_rec(orc_semi* h, int level, int solver, int nu1, int nu2, int ncoarse) {
  int Lmax = h->p.multi_levels;
  if (level == Lmax) { orc_semi_smooth(h, level, solver, ncoarse); return; }
  orc_semi_smooth(h, level, solver, nu1);
  orc_semi::Level& L = h->lev[level - 1];
  L.tnew = L.tnonlin;
  orc_semi_update_overlaps(h, level);
  orc_semi_residual(h, level, nullptr, nullptr);
  orc_semi_restrict(h, level);
  orc_semi::Level& Cc = h->lev[level];
  std::fill(Cc.tnonlin.begin(), Cc.tnonlin.end(), 0.0);
  std::fill(Cc.tnew.begin(), Cc.tnew.end(), 0.0);
  vcycle_rec(h, level + 1, solver, nu1, nu2, ncoarse);
  orc_semi_prolong(h, level);
  orc_semi_smooth(h, level, solver, nu2);
}

// INTENDED composition.  Requires residual_sign = -1 and transfer = 1 for a convergent cycle.
int orc_semi_vcycle_solve(orc_semi* h, int solver, int nu1, int nu2, int ncoarse, int max_cycles,
                          double tol, double* hist) {
  orc_semi::Level& L = h->lev[0];
  L.tnew = L.tnonlin;
  orc_semi_update_overlaps(h, 1);
  double r0, r;
  orc_semi_residual(h, 1, &r0, nullptr);
  if (hist) hist[0] = r0;
  if (r0 == 0) return 0;
  for (int c = 1; c <= max_cycles; ++c) {
    vcycle_rec(h, 1, solver, nu1, nu2, ncoarse);
    L.tnew = L.tnonlin;
    orc_semi_update_overlaps(h, 1);
    orc_semi_residual(h, 1, &r, nullptr);
    if (hist) hist[c] = r;
    if (r / r0 <= tol) return c;
  }
  return max_cycles + 1;
}

// HEAD time-loop body, transport_tri_semi.F90:316-379 (one itime), with every quirk of SURVEY B-6/B-7.
void orc_semi_literal_timestep(orc_semi* h, int solver, int n_multigrid, int n_smooth) {
  int ML = h->p.multi_levels;
  h->lev[0].told = h->lev[0].tnew;         // :316
  h->lev[0].tnonlin = h->lev[0].tnew;      // :317
  for (int mg = 0; mg < n_multigrid; ++mg) {
    for (int il = 1; il <= ML; ++il) {
      h->lev[il - 1].tnonlin = h->lev[il - 1].tnew;  // :327
      orc_semi_smooth(h, il, solver, n_smooth);      // :331
      orc_semi_restrict(h, il);                      // :336 (restricts the residual of the previous pass)
      orc_semi_residual(h, il, nullptr, nullptr);    // :338
    }
    h->lev[ML - 1].tnonlin = h->lev[ML - 1].tnew;    // :348
    for (int i = 0; i < 15; ++i) orc_semi_smooth(h, ML, solver, n_smooth);  // :351-352
    for (int il = ML - 1; il >= 1; --il) {
      h->lev[il - 1].tnonlin = h->lev[il - 1].tnew;  // :367
      orc_semi_prolong(h, il);                       // :370
      orc_semi_smooth(h, il, solver, n_smooth);      // :376
    }
  }
}

// ------------------------------------------------------------------ unstructured explicit
// transport_tri_unstr.F90:588-795.  use_dir=0 reproduces get_unstr_sn2 ignoring Dir (SURVEY B-9).
void orc_unstr_explicit(int E, const double* X, const int32_t* neig, const int32_t* fneig,
                        const int32_t* /*dir: unstr_explicit ignores Dir, SURVEY B-9*/, double u_x, double u_y, double dt, int ntime, int nits,
                        int njac_its, int exact_minv, int use_dir, double t_bc, double* tnew) {
  std::vector<double> told((size_t)3 * E), tnl((size_t)3 * E);
  for (int it = 0; it < ntime; ++it) {
    std::copy(tnew, tnew + (size_t)3 * E, told.begin());
    std::copy(tnew, tnew + (size_t)3 * E, tnl.begin());
    for (int its = 0; its < nits; ++its) {
      std::copy(tnl.begin(), tnl.end(), tnew);
      for (int e = 0; e < E; ++e) {
        const double (*x)[2] = reinterpret_cast<const double (*)[2]>(&X[(size_t)e * 6]);
        double nx[3][2][3], detwei[3];
        tri_det_nlx(x, nx, detwei);
        const double* Tl = &tnew[(size_t)e * 3];
        const double* To = &told[(size_t)e * 3];
        double ugi[3][2], tgi[3];
        for (int g = 0; g < 3; ++g) {
          ugi[g][0] = (TB.n[g][0] + TB.n[g][1] + TB.n[g][2]) * u_x;
          ugi[g][1] = (TB.n[g][0] + TB.n[g][1] + TB.n[g][2]) * u_y;
          tgi[g] = TB.n[g][0] * Tl[0] + TB.n[g][1] * Tl[1] + TB.n[g][2] * Tl[2];
        }
        double rhs[3] = {0, 0, 0}, mass[3][3], ml[3];
        for (int i = 0; i < 3; ++i) {
          for (int j = 0; j < 3; ++j) {
            double m = 0;
            for (int g = 0; g < 3; ++g) m += TB.n[g][i] * TB.n[g][j] * detwei[g];
            mass[i][j] = m;
          }
          ml[i] = 0;
          for (int g = 0; g < 3; ++g) ml[i] += TB.n[g][i] * detwei[g];
          for (int g = 0; g < 3; ++g)
            for (int d = 0; d < 2; ++d) rhs[i] += nx[g][d][i] * ugi[g][d] * tgi[g] * detwei[g];
        }
        for (int f = 1; f <= 3; ++f) {
          int npos = neig[e * 3 + f - 1], nside = fneig[e * 3 + f - 1];
          double T2[3], u2[2] = {u_x, u_y};
          if (npos != 0) for (int q = 0; q < 3; ++q) T2[q] = tnew[(size_t)(npos - 1) * 3 + q];
          else for (int q = 0; q < 3; ++q) T2[q] = t_bc;
          double sn[2][3] = {{0, 0, 0}, {0, 0, 0}}, sn2[2][3] = {{0, 0, 0}, {0, 0, 0}};
          int l1 = UN_FACE_NODES[f - 1][0] - 1, l2 = UN_FACE_NODES[f - 1][1] - 1;
          for (int s = 0; s < 2; ++s) { sn[s][l1] = TB.sn[s][0]; sn[s][l2] = TB.sn[s][1]; }
          // get_unstr_sn2, ShapFun_unstruc.F90:205-222 : Nside 1->(3,1) 2->(1,2) 3->(2,3); Nside=0 leaves sn2=0
          const int N2[3][2] = {{3, 1}, {1, 2}, {2, 3}};
          if (nside >= 1) {
            int m1 = N2[nside - 1][0] - 1, m2 = N2[nside - 1][1] - 1;
            if (use_dir && npos != 0) {
              // INTENDED: pair by geometry instead of assuming opposite edge orientation
              const double (*xn)[2] = reinterpret_cast<const double (*)[2]>(&X[(size_t)(npos - 1) * 6]);
              if (!are_equal2(xn[m1], x[l1])) std::swap(m1, m2);
            }
            for (int s = 0; s < 2; ++s) { sn2[s][m1] = TB.sn[s][0]; sn2[s][m2] = TB.sn[s][1]; }
          }
          double sdet[2], snorm[2][2];
          face_geometry(x, f, sdet, snorm);
          for (int s = 0; s < 2; ++s) {
            double ts = 0, t2s = 0, sum2 = 0;
            for (int q = 0; q < 3; ++q) { ts += sn[s][q] * Tl[q]; t2s += sn2[s][q] * T2[q]; sum2 += sn2[s][q]; }
            double us[2] = {u_x, u_y};
            double us2[2] = {sum2 * u2[0], sum2 * u2[1]};
            double un = snorm[s][0] * 0.5 * (us[0] + us2[0]) + snorm[s][1] * 0.5 * (us[1] + us2[1]);
            double income = 0.5 + 0.5 * std::copysign(1.0, -un);
            for (int d = 0; d < 2; ++d) {
              double sc = snorm[s][d] * sdet[s] * ((1.0 - income) * us[d] * ts + income * us2[d] * t2s);
              for (int i = 0; i < 3; ++i) rhs[i] -= sn[s][i] * sc;
            }
          }
        }
        double* out = &tnl[(size_t)e * 3];
        if (exact_minv) {
          // transport_rect.F90:277-291 semantics with FINDInv
          double inv[9], mt[3];
          findinv(&mass[0][0], inv, 3);
          for (int i = 0; i < 3; ++i) mt[i] = mass[i][0] * To[0] + mass[i][1] * To[1] + mass[i][2] * To[2];
          double o3[3];
          for (int i = 0; i < 3; ++i) {
            o3[i] = 0;
            for (int j = 0; j < 3; ++j) o3[i] += inv[i * 3 + j] * (mt[j] + dt * rhs[j]);
          }
          for (int i = 0; i < 3; ++i) out[i] = o3[i];
        } else {
          double rj[3], tl[3] = {out[0], out[1], out[2]};
          for (int i = 0; i < 3; ++i) rj[i] = mass[i][0] * To[0] + mass[i][1] * To[1] + mass[i][2] * To[2] + dt * rhs[i];
          for (int jit = 0; jit < njac_its; ++jit) {
            double mt[3];
            for (int i = 0; i < 3; ++i) mt[i] = mass[i][0] * tl[0] + mass[i][1] * tl[1] + mass[i][2] * tl[2];
            for (int i = 0; i < 3; ++i) tl[i] = (ml[i] * tl[i] - mt[i] + rj[i]) / ml[i];
          }
          for (int i = 0; i < 3; ++i) out[i] = tl[i];
        }
      }
      std::copy(tnl.begin(), tnl.end(), tnew);  // :792
    }
  }
}

// ------------------------------------------------------------------ unstructured implicit
// Assembly of unstr_implicit, transport_tri_unstr.F90:270-364: per element the block mass/dt - stiff (:270-292) and
// per face the 3x3 upwind block flux_ele(iloc,jloc) (:344-362) that lands in the columns of the element itself
// (outflow) or of the neighbour (inflow), `target_ele` (:339-342).  The CSR containers of the reference
// (add_to_CSR assigns, add_to_CSR_flux accumulates in a fixed window, SURVEY B-15) are replaced by plain accumulation
// into dense row-major matrices A (lhs + flux) and Mdt (mass/dt) of size (3E)^2 - small meshes only.
void orc_unstr_implicit_assemble(int E, const double* X, const int32_t* neig, const int32_t* fneig, double u_x,
                                 double u_y, double dt, int use_dir, double* A, double* Mdt) {
  orc_unstr_implicit_assemble_diff(E, X, neig, fneig, u_x, u_y, 0.0, dt, use_dir, A, Mdt);
}

// the same with the diffusion operator of the iterative path added (INTENDED: the implicit drivers call add_diffusion_vol and
// drop its result, transport_tri_semi.F90:1627): volume term k nx.nx detwei (diff_vol_stcl, ShapFun_unstruc.F90:324-335) and the
// face penalty (k/dx) sum_g sn_i (T - T2) sdetwei (matrices.F90:113-115, get_diff_surf_stencl transport_tri_semi.F90:468-477)
// with dx = centroid distance to the neighbour, or centre -> edge midpoint on the domain boundary (matrices.F90:84-110), where
// the exterior trace is Dirichlet data (right-hand side) and only the own part enters the matrix.
void orc_unstr_implicit_assemble_diff(int E, const double* X, const int32_t* neig, const int32_t* fneig, double u_x,
                                      double u_y, double k, double dt, int use_dir, double* A, double* Mdt) {
  const size_t N = (size_t)3 * E;
  std::fill(A, A + N * N, 0.0);
  std::fill(Mdt, Mdt + N * N, 0.0);
  for (int e = 0; e < E; ++e) {
    const double (*x)[2] = reinterpret_cast<const double (*)[2]>(&X[(size_t)e * 6]);
    double nx[3][2][3], detwei[3];
    tri_det_nlx(x, nx, detwei);
    double ugi[3][2];
    for (int g = 0; g < 3; ++g) {
      ugi[g][0] = (TB.n[g][0] + TB.n[g][1] + TB.n[g][2]) * u_x;
      ugi[g][1] = (TB.n[g][0] + TB.n[g][1] + TB.n[g][2]) * u_y;
    }
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        double mass = 0, stiff = 0;
        for (int g = 0; g < 3; ++g) {
          mass += TB.n[g][i] * TB.n[g][j] * detwei[g] / dt;
          for (int d = 0; d < 2; ++d) stiff += nx[g][d][i] * ugi[g][d] * detwei[g] * TB.n[g][j];
        }
        double dvol = 0;
        for (int g = 0; g < 3; ++g)
          for (int d = 0; d < 2; ++d) dvol += k * nx[g][d][i] * detwei[g] * nx[g][d][j];
        Mdt[(size_t)(3 * e + i) * N + 3 * e + j] += mass;
        A[(size_t)(3 * e + i) * N + 3 * e + j] += mass - stiff + dvol;
      }
    for (int f = 1; f <= 3; ++f) {
      const int npos = neig[e * 3 + f - 1], nside = fneig[e * 3 + f - 1];
      double sn[2][3] = {{0, 0, 0}, {0, 0, 0}}, sn2[2][3] = {{0, 0, 0}, {0, 0, 0}};
      const int l1 = UN_FACE_NODES[f - 1][0] - 1, l2 = UN_FACE_NODES[f - 1][1] - 1;
      for (int s = 0; s < 2; ++s) { sn[s][l1] = TB.sn[s][0]; sn[s][l2] = TB.sn[s][1]; }
      const int N2[3][2] = {{3, 1}, {1, 2}, {2, 3}};
      if (nside >= 1) {
        int m1 = N2[nside - 1][0] - 1, m2 = N2[nside - 1][1] - 1;
        if (use_dir && npos != 0) {
          const double (*xn)[2] = reinterpret_cast<const double (*)[2]>(&X[(size_t)(npos - 1) * 6]);
          if (!are_equal2(xn[m1], x[l1])) std::swap(m1, m2);
        }
        for (int s = 0; s < 2; ++s) { sn2[s][m1] = TB.sn[s][0]; sn2[s][m2] = TB.sn[s][1]; }
      }
      double sdet[2], snorm[2][2], income[2], us[2][2], us2[2][2];
      face_geometry(x, f, sdet, snorm);
      for (int s = 0; s < 2; ++s) {
        double sum2 = 0;
        for (int q = 0; q < 3; ++q) sum2 += sn2[s][q];
        us[s][0] = u_x; us[s][1] = u_y; us2[s][0] = sum2 * u_x; us2[s][1] = sum2 * u_y;
        const double un = snorm[s][0] * 0.5 * (us[s][0] + us2[s][0]) + snorm[s][1] * 0.5 * (us[s][1] + us2[s][1]);
        income[s] = 0.5 + 0.5 * std::copysign(1.0, -un);
      }
      int target = (int)((1.0 - income[0]) * (e + 1) + income[0] * npos);   // :339
      if (target == 0) target = e + 1;                                      // :340-342
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
          double fl = 0;
          for (int d = 0; d < 2; ++d)
            for (int s = 0; s < 2; ++s)
              fl += snorm[s][d] * sdet[s] * sn[s][i] *
                    ((1.0 - income[s]) * sn[s][j] * us[s][d] + income[s] * sn2[s][j] * us2[s][d]);
          A[(size_t)(3 * e + i) * N + 3 * (target - 1) + j] += fl;
        }
      if (k != 0.0) {
        const double cx = (x[0][0] + x[1][0] + x[2][0]) / 3, cy = (x[0][1] + x[1][1] + x[2][1]) / 3;
        double dx;
        if (npos != 0) {
          const double (*xn)[2] = reinterpret_cast<const double (*)[2]>(&X[(size_t)(npos - 1) * 6]);
          const double qx = (xn[0][0] + xn[1][0] + xn[2][0]) / 3, qy = (xn[0][1] + xn[1][1] + xn[2][1]) / 3;
          dx = std::sqrt((cx - qx) * (cx - qx) + (cy - qy) * (cy - qy));
        } else {
          const double mx = (x[l1][0] + x[l2][0]) / 2, my = (x[l1][1] + x[l2][1]) / 2;
          dx = std::sqrt((cx - mx) * (cx - mx) + (cy - my) * (cy - my));
        }
        for (int i = 0; i < 3; ++i)
          for (int j = 0; j < 3; ++j) {
            double my_d = 0, ng_d = 0;
            for (int s = 0; s < 2; ++s) {
              my_d += (k / dx) * sn[s][i] * sn[s][j] * sdet[s];
              ng_d += (k / dx) * sn[s][i] * sn2[s][j] * sdet[s];
            }
            A[(size_t)(3 * e + i) * N + 3 * e + j] += my_d;
            if (npos != 0) A[(size_t)(3 * e + i) * N + 3 * (npos - 1) + j] -= ng_d;
          }
      }
    }
  }
}

// time loop of unstr_implicit (:214-387): told = tnew ; rhs = (M/dt) told ; tnew = inverse(lhs + flux) rhs by FINDInv
int orc_unstr_implicit(int E, const double* X, const int32_t* neig, const int32_t* fneig, double u_x, double u_y,
                       double dt, int ntime, int nits, int use_dir, double* tnew) {
  const size_t N = (size_t)3 * E;
  std::vector<double> A(N * N), M(N * N), inv(N * N), rhs(N), told(N);
  orc_unstr_implicit_assemble(E, X, neig, fneig, u_x, u_y, dt, use_dir, A.data(), M.data());
  if (findinv(A.data(), inv.data(), (int)N) != 0) return -1;
  for (int it = 0; it < ntime; ++it) {
    std::copy(tnew, tnew + N, told.begin());
    for (int k = 0; k < nits; ++k) {
      for (size_t i = 0; i < N; ++i) { double s = 0; for (size_t j = 0; j < N; ++j) s += M[i * N + j] * told[j]; rhs[i] = s; }
      for (size_t i = 0; i < N; ++i) { double s = 0; for (size_t j = 0; j < N; ++j) s += inv[i * N + j] * rhs[j]; tnew[i] = s; }
    }
  }
  return 0;
}

// Petrov-Galerkin residual-based stabilisation of unstr_implicit (transport_tri_unstr.F90:239-267,278): per Gauss point
// rgi, a_star, p_star (eq 23 form with inv_jac), diff_coe, and the element matrix stab(i,j) = sum_g diff_coe(g)
// grad(phi_j).grad(phi_i) detwei(g).  HEAD computes stab and never uses it (the only use, :367-368, is commented out).
static void stab_element(const double x[3][2], const double* tn, const double* to, double u_x, double u_y, double dt,
                         double diff_coe[3], double stab[9]) {
  const double toler = 0.00000000001;   // :93
  double nx[3][2][3], detwei[3];
  tri_det_nlx(x, nx, detwei);
  // inv_jac (ShapFun.F90:1440-1450): (1,1)=D/detJ (1,2)=-C/detJ (2,1)=-B/detJ (2,2)=A/detJ, the same at every Gauss point
  double A = 0, B = 0, C = 0, D = 0;
  for (int l = 0; l < 3; ++l) {
    A += TB.nlx[0][0][l] * x[l][0]; B += TB.nlx[0][0][l] * x[l][1];
    C += TB.nlx[0][1][l] * x[l][0]; D += TB.nlx[0][1][l] * x[l][1];
  }
  const double detj = A * D - B * C;
  const double ij[2][2] = {{D / detj, -C / detj}, {-B / detj, A / detj}};
  for (int g = 0; g < 3; ++g) {
    double ugi[2] = {0, 0}, txgi[2] = {0, 0}, tgi = 0, togi = 0;
    for (int l = 0; l < 3; ++l) {
      ugi[0] += TB.n[g][l] * u_x; ugi[1] += TB.n[g][l] * u_y;
      txgi[0] += nx[g][0][l] * tn[l]; txgi[1] += nx[g][1][l] * tn[l];
      tgi += TB.n[g][l] * tn[l]; togi += TB.n[g][l] * to[l];
    }
    const double rgi = (tgi - togi) / dt + (ugi[0] * txgi[0] + ugi[1] * txgi[1]);
    const double g2 = txgi[0] * txgi[0] + txgi[1] * txgi[1];
    const double a_coef = rgi / std::max(toler, g2);
    const double as[2] = {a_coef * txgi[0], a_coef * txgi[1]};
    double ps = 0.0;
    for (int d = 0; d < 2; ++d) ps = std::max(ps, std::fabs(as[0] * ij[0][d] + as[1] * ij[1][d]));
    ps = std::min(1.0 / toler, 0.25 / ps);
    diff_coe[g] = 0.25 * rgi * rgi * ps / std::max(toler, g2);
  }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double v = 0;
      for (int g = 0; g < 3; ++g) v += diff_coe[g] * (nx[g][0][j] * nx[g][0][i] + nx[g][1][j] * nx[g][1][i]) * detwei[g];
      stab[i * 3 + j] = v;
    }
}

void orc_unstr_stab(int E, const double* X, const double* tnew, const double* told, double u_x, double u_y, double dt,
                    double* diff_coe, double* stab) {
  for (int e = 0; e < E; ++e)
    stab_element(reinterpret_cast<const double (*)[2]>(&X[(size_t)e * 6]), tnew + (size_t)e * 3, told + (size_t)e * 3, u_x,
                 u_y, dt, diff_coe + (size_t)e * 3, stab + (size_t)e * 9);
}

// INTENDED use of the stabilisation (what the commented `mat_loc = mass_ele + dt*stab` at :367-368 amounts to for the
// implicit system): every nonlinear pass adds stab(tnew_nonlin, told) to the diagonal blocks of lhs + flux and re-solves.
int orc_unstr_implicit_stab(int E, const double* X, const int32_t* neig, const int32_t* fneig, double u_x, double u_y,
                            double dt, int ntime, int nits, int use_dir, double* tnew) {
  const size_t N = (size_t)3 * E;
  std::vector<double> A(N * N), As(N * N), M(N * N), inv(N * N), rhs(N), told(N), dc(N), st((size_t)9 * E);
  orc_unstr_implicit_assemble(E, X, neig, fneig, u_x, u_y, dt, use_dir, A.data(), M.data());
  for (int it = 0; it < ntime; ++it) {
    std::copy(tnew, tnew + N, told.begin());
    for (size_t i = 0; i < N; ++i) { double s = 0; for (size_t j = 0; j < N; ++j) s += M[i * N + j] * told[j]; rhs[i] = s; }
    for (int k = 0; k < nits; ++k) {
      orc_unstr_stab(E, X, tnew, told.data(), u_x, u_y, dt, dc.data(), st.data());
      As = A;
      for (int e = 0; e < E; ++e)
        for (int i = 0; i < 3; ++i)
          for (int j = 0; j < 3; ++j) As[(size_t)(3 * e + i) * N + 3 * e + j] += st[(size_t)e * 9 + i * 3 + j];
      if (findinv(As.data(), inv.data(), (int)N) != 0) return -1;
      for (size_t i = 0; i < N; ++i) { double s = 0; for (size_t j = 0; j < N; ++j) s += inv[i * N + j] * rhs[j]; tnew[i] = s; }
    }
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------
// trans_rec (transport_rect.F90:7-380): explicit DG on bilinear quadrilaterals (nloc = 4, 2x2 Gauss points, 2 Gauss points
// per face) of a structured rectangular grid, element-local solve by FINDInv or by Jacobi iterations on the lumped mass.
//   RE2DN4, NGI = 4 branch (ShapFun.F90:72-215); ele_info (structured_meshgen.F90:6-71); surface_pointers_sn
//   (ShapFun.F90:258-367); det_nlx (:1245-1386, 2-D); det_snlx_all / NORMGI / XPROD1 (:1554-1590, :2012-2054).
// volume_term = 0 is HEAD: the line that sets tnew_gi (:157) is commented out, so the advection VOLUME integral (:206-208)
// multiplies a zero array.  That is the version that wrote the shipped output file DG-rectangular_structured
// (main.F90:19: CFL .7, 200 x 1 elements, u = (2*0.01428571, 0), time 250, nits 2, 10 Jacobi iterations), which this
// restatement reproduces to the reference's single precision - the one pin against an output of the reference binary.
// volume_term = 1 is the intended scheme.
int orc_trans_rec(double CFL, int no_ele_row, int no_ele_col, double x_length, double y_length, double u_x, double u_y,
                  double time, int nits, int njac_its, int direct_solver, int volume_term, double* x_all, double* tnew_out) {
  const int nloc = 4, ngi = 4, sngi = 2, nface = 4;
  const int totele = no_ele_row * no_ele_col;
  const double dx = x_length / no_ele_row, dy = y_length / no_ele_col, dt = CFL * dx;
  const int ntime = (int)(time / dt);
  // RE2DN4
  const double lxp[4] = {-1, 1, -1, 1}, lyp[4] = {-1, -1, 1, 1};
  const double posi = 1.0 / std::sqrt(3.0), lx[2] = {-posi, posi};
  double n[4][4], nlx[4][2][4], weight[4];
  for (int q = 0; q < 2; ++q)
    for (int pq = 0; pq < 2; ++pq)
      for (int c = 0; c < 4; ++c) {
        const int g = q * 2 + pq;
        weight[g] = 1.0;
        n[g][c] = 0.25 * (1.0 + lxp[c] * lx[pq]) * (1.0 + lyp[c] * lx[q]);
        nlx[g][0][c] = 0.25 * lxp[c] * (1.0 + lyp[c] * lx[q]);
        nlx[g][1][c] = 0.25 * lyp[c] * (1.0 + lxp[c] * lx[pq]);
      }
  double sn_o[2][2], snlx_o[2][2];
  for (int pq = 0; pq < 2; ++pq)
    for (int c = 0; c < 2; ++c) { sn_o[pq][c] = 0.5 * (1.0 + lxp[c] * lx[pq]); snlx_o[pq][c] = 0.5 * lxp[c]; }
  // surface_pointers_sn: (lnod1, lnod2) of my side and of the neighbour's side, 1-based
  const int FN[4][2] = {{2, 1}, {1, 3}, {4, 2}, {3, 4}}, FN2[4][2] = {{4, 3}, {2, 4}, {3, 1}, {1, 2}};
  double face_sn[4][2][4] = {}, face_snlx[4][2][4] = {}, face_sn2[4][2][4] = {};
  for (int f = 0; f < 4; ++f)
    for (int sg = 0; sg < 2; ++sg) {
      face_sn[f][sg][FN[f][0] - 1] = sn_o[sg][0]; face_sn[f][sg][FN[f][1] - 1] = sn_o[sg][1];
      face_snlx[f][sg][FN[f][0] - 1] = snlx_o[sg][0]; face_snlx[f][sg][FN[f][1] - 1] = snlx_o[sg][1];
      face_sn2[f][sg][FN2[f][0] - 1] = sn_o[sg][0]; face_sn2[f][sg][FN2[f][1] - 1] = sn_o[sg][1];
    }
  // ele_info
  std::vector<int> face_ele((size_t)totele * 4);
  for (int ele = 1; ele <= totele; ++ele) {
    const int row = (ele + no_ele_row - 1) / no_ele_row, col = ele - no_ele_row * (row - 1);
    double* x = x_all + (size_t)(ele - 1) * 8;
    x[0] = dx * (col - 1); x[1] = dy * (row - 1);
    x[2] = dx * col;       x[3] = dy * (row - 1);
    x[4] = dx * (col - 1); x[5] = dy * row;
    x[6] = dx * col;       x[7] = dy * row;
    auto rowof = [&](int e) { return (int)std::ceil((double)e / no_ele_row); };
    int* fe = &face_ele[(size_t)(ele - 1) * 4];
    fe[0] = ele - no_ele_row;
    fe[1] = ele - 1; if (rowof(fe[1]) != row) fe[1] = -fe[1];
    fe[2] = ele + 1; if (rowof(fe[2]) != row) fe[2] = -fe[2];
    fe[3] = ele + no_ele_row; if (fe[3] > totele) fe[3] = -fe[3];
  }
  std::vector<double> tnew((size_t)totele * 4, 0.0), told, tnl;
  for (int ele = no_ele_row / 5; ele <= no_ele_row / 2; ++ele)        // tnew(:, no_ele_row/5 : no_ele_row/2) = 1 (:83)
    if (ele >= 1 && ele <= totele) for (int i = 0; i < 4; ++i) tnew[(size_t)(ele - 1) * 4 + i] = 1.0;
  for (int itime = 0; itime < ntime; ++itime) {
    told = tnew; tnl = tnew;
    for (int its = 0; its < nits; ++its) {
      tnew = tnl;
      for (int ele = 1; ele <= totele; ++ele) {
        const double* xl = x_all + (size_t)(ele - 1) * 8;
        // det_nlx
        double nx[4][2][4], detwei[4];
        for (int g = 0; g < ngi; ++g) {
          double A = 0, B = 0, Cc = 0, D = 0;
          for (int l = 0; l < nloc; ++l) {
            A += nlx[g][0][l] * xl[2 * l]; B += nlx[g][0][l] * xl[2 * l + 1];
            Cc += nlx[g][1][l] * xl[2 * l]; D += nlx[g][1][l] * xl[2 * l + 1];
          }
          const double detj = A * D - B * Cc;
          detwei[g] = std::fabs(detj) * weight[g];
          const double a11 = D / detj, a21 = -B / detj, a12 = -Cc / detj, a22 = A / detj;
          for (int l = 0; l < nloc; ++l) {
            nx[g][0][l] = a11 * nlx[g][0][l] + a12 * nlx[g][1][l];
            nx[g][1][l] = a21 * nlx[g][0][l] + a22 * nlx[g][1][l];
          }
        }
        const double* tl = &tnew[(size_t)(ele - 1) * 4];
        const double* to = &told[(size_t)(ele - 1) * 4];
        double ugi[4][2], tgi[4];
        for (int g = 0; g < ngi; ++g) {
          double sx = 0, sy = 0, tt = 0;
          for (int l = 0; l < nloc; ++l) { sx += n[g][l] * u_x; sy += n[g][l] * u_y; tt += n[g][l] * tl[l]; }
          ugi[g][0] = sx; ugi[g][1] = sy;
          tgi[g] = volume_term ? tt : 0.0;
        }
        double mass[4][4], ml[4], rhs[4];
        for (int i = 0; i < nloc; ++i) {
          for (int j = 0; j < nloc; ++j) {
            double m = 0;
            for (int g = 0; g < ngi; ++g) m += n[g][i] * n[g][j] * detwei[g];
            mass[i][j] = m;
          }
          double l = 0, r = 0;
          for (int g = 0; g < ngi; ++g) {
            l += n[g][i] * detwei[g];
            for (int d = 0; d < 2; ++d) r += nx[g][d][i] * ugi[g][d] * tgi[g] * detwei[g];
          }
          ml[i] = l; rhs[i] = r;
        }
        for (int f = 0; f < nface; ++f) {
          const int e22 = face_ele[(size_t)(ele - 1) * 4 + f];
          const bool bnd = e22 <= 0;                                   // (sign(1, -ele22) + 1) / 2
          const double* t2 = bnd ? nullptr : &tnew[(size_t)(e22 - 1) * 4];
          double usgi[2][2] = {}, usgi2[2][2] = {}, xsgi[2][2] = {}, tsgi[2] = {}, tsgi2[2] = {};
          for (int l = 0; l < nloc; ++l)
            for (int sg = 0; sg < sngi; ++sg) {
              usgi[sg][0] += face_sn[f][sg][l] * u_x; usgi[sg][1] += face_sn[f][sg][l] * u_y;
              usgi2[sg][0] += face_sn2[f][sg][l] * u_x; usgi2[sg][1] += face_sn2[f][sg][l] * u_y;   // u_bc = u_ele (:93)
              xsgi[sg][0] += face_sn[f][sg][l] * xl[2 * l]; xsgi[sg][1] += face_sn[f][sg][l] * xl[2 * l + 1];
              tsgi[sg] += face_sn[f][sg][l] * tl[l];
              tsgi2[sg] += face_sn2[f][sg][l] * (bnd ? 0.0 : t2[l]);                                  // t_bc = 0 (:80)
            }
          double norm[2];
          for (int d = 0; d < 2; ++d)
            norm[d] = (xsgi[0][d] + xsgi[1][d]) / 2.0 - (xl[d] + xl[2 + d] + xl[4 + d] + xl[6 + d]) / 4.0;
          double sdet[2], snorm[2][2];
          for (int sg = 0; sg < sngi; ++sg) {
            double dxdlx = 0, dydlx = 0;
            for (int l = 0; l < nloc; ++l) { dxdlx += face_snlx[f][sg][l] * xl[2 * l]; dydlx += face_snlx[f][sg][l] * xl[2 * l + 1]; }
            sdet[sg] = std::sqrt(dydlx * dydlx + dxdlx * dxdlx) * 1.0;
            const double ax = dydlx, ay = -dxdlx;
            const double rn = std::sqrt(ax * ax + ay * ay);
            const double sirn = std::copysign(1.0 / rn, ax * norm[0] + ay * norm[1]);
            snorm[sg][0] = sirn * ax; snorm[sg][1] = sirn * ay;
          }
          for (int sg = 0; sg < sngi; ++sg) {
            const double un = snorm[sg][0] * 0.5 * (usgi[sg][0] + usgi2[sg][0]) + snorm[sg][1] * 0.5 * (usgi[sg][1] + usgi2[sg][1]);
            const double income = 0.5 + 0.5 * std::copysign(1.0, -un);
            for (int d = 0; d < 2; ++d) {
              const double sc = snorm[sg][d] * sdet[sg] * ((1.0 - income) * usgi[sg][d] * tsgi[sg] + income * usgi2[sg][d] * tsgi2[sg]);
              for (int i = 0; i < nloc; ++i) rhs[i] -= face_sn[f][sg][i] * sc;
            }
          }
        }
        double* out = &tnl[(size_t)(ele - 1) * 4];
        if (direct_solver) {
          double inv[16], m16[16];
          for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) m16[i * 4 + j] = mass[i][j];
          if (findinv(m16, inv, 4) != 0) return -1;
          double v[4];
          for (int i = 0; i < 4; ++i) { double sm = 0; for (int j = 0; j < 4; ++j) sm += mass[i][j] * to[j]; v[i] = sm + dt * rhs[i]; }
          for (int i = 0; i < 4; ++i) { double sm = 0; for (int j = 0; j < 4; ++j) sm += inv[i * 4 + j] * v[j]; out[i] = sm; }
        } else {
          double rj[4], x4[4] = {out[0], out[1], out[2], out[3]};
          for (int i = 0; i < 4; ++i) { double sm = 0; for (int j = 0; j < 4; ++j) sm += mass[i][j] * to[j]; rj[i] = sm + dt * rhs[i]; }
          for (int k = 0; k < njac_its; ++k) {
            double mt[4];
            for (int i = 0; i < 4; ++i) { double sm = 0; for (int j = 0; j < 4; ++j) sm += mass[i][j] * x4[j]; mt[i] = sm; }
            for (int i = 0; i < 4; ++i) x4[i] = (ml[i] * x4[i] - mt[i] + rj[i]) / ml[i];
          }
          for (int i = 0; i < 4; ++i) out[i] = x4[i];
        }
      }
      tnew = tnl;
    }
  }
  std::copy(tnew.begin(), tnew.end(), tnew_out);
  return ntime;
}

// transport_rect.F90:48-52,83,101-105,337-344 ; structured_meshgen.F90:25-33 (one row of quads)
void orc_rect_analytical(double CFL, int no_ele_row, double x_length, double u_x, double time,
                         double* x_out, double* t_out) {
  double dx = x_length / no_ele_row, dt = CFL * dx;
  int ntime = (int)(time / dt);
  int offset_x = (int)(u_x * dt * ntime * no_ele_row / x_length + 1);
  int lo = offset_x + no_ele_row / 5, hi = offset_x + no_ele_row / 2;
  for (int ele = 1; ele <= no_ele_row; ++ele) {
    double v = (ele >= lo && ele <= hi) ? 1.0 : 0.0;
    double xs[4] = {dx * (ele - 1), dx * ele, dx * (ele - 1), dx * ele};
    for (int q = 0; q < 4; ++q) { x_out[(ele - 1) * 4 + q] = xs[q]; t_out[(ele - 1) * 4 + q] = v; }
  }
}

double orc_thermal_analytical(double x, double t, double u, double gamma) {
  (void)u;
  const double pi = 3.141596;  // the script's own constant, Check_thermal_analytical_validation.py:24
  double term1 = std::erfc((x - gamma * t) / (2.0 * std::sqrt(t)));
  double term2 = std::exp(gamma * x) * std::erfc((x + gamma * t) / (2.0 * std::sqrt(t)));
  double term3 = 1.0 + 0.5 * gamma * (2.0 - x + gamma * t);
  double term4 = std::erfc((2.0 - x + gamma * t) / (2.0 * std::sqrt(t)));
  double term5 = gamma * std::sqrt(t / pi) * std::exp(-((2.0 - x + gamma * t) * (2.0 - x + gamma * t)) / (4.0 * t));
  return 0.5 * (term1 + term2) + std::exp(gamma) * (term3 * term4 - term5);
}

}  // extern "C"

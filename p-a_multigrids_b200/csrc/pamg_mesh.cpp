// pamg_mesh.cpp -- host mesh pipeline of libpamg_cuda.so.
//
// Replaces ReadMSH (Msh2Tri.F90:132-334), the all-pairs CheckNeig neighbour search
// (Msh2Tri.F90:323-330,780-963; 788 s for 98 k elements in grofiling.txt) and getNeigDataMesh
// (Msh2Tri.F90:454-548) by an O(N) edge hash.  The semantics of Neig / Dir / fNeig are the
// reference's; only the search differs.  Also generates the synthetic semi-structured inputs of
// SURVEY.md section 8(d).
#include "pamg_internal.h"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <unordered_map>

namespace pamg {

namespace {

struct VKey {
  uint64_t x, y;
  bool operator==(const VKey& o) const { return x == o.x && y == o.y; }
};
struct EKey {
  VKey a, b;  // a <= b lexicographically
  bool operator==(const EKey& o) const { return a == o.a && b == o.b; }
};
struct EHash {
  size_t operator()(const EKey& k) const {
    uint64_t h = 1469598103934665603ull;
    for (uint64_t v : {k.a.x, k.a.y, k.b.x, k.b.y}) { h ^= v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2); }
    return (size_t)h;
  }
};
inline uint64_t bits(double v) {
  if (v == 0.0) v = 0.0;  // -0 -> +0
  uint64_t u; std::memcpy(&u, &v, 8); return u;
}
inline VKey vkey(const double* p) { return VKey{bits(p[0]), bits(p[1])}; }
inline bool vless(const VKey& p, const VKey& q) { return p.x < q.x || (p.x == q.x && p.y < q.y); }

// nodes of parent side 1,2,3 = (Xp1,Xp3), (Xp1,Xp2), (Xp2,Xp3): Msh2Tri.F90:877-901
const int SIDE[3][2] = {{0, 2}, {0, 1}, {1, 2}};

// Dir of the LOWER-numbered triangle i for its side f, given which vertex of j each vertex of i
// coincides with (m[a] in {0,1,2} or -1).  Same truth table as the vertex-code test of
// Msh2Tri.F90:881,889,897; the higher-numbered triangle copies the value (:906-931).
inline int dir_rule(int f, const int m[3]) {
  switch (f) {
    case 0: return (m[0] == 0) || (m[0] == 1 && m[2] == 2);
    case 1: return (m[0] == 0) || (m[0] == 1 && m[1] == 2);
    default: return (m[2] == 2) || (m[1] == 0 && m[2] == 1);
  }
}

}  // namespace

void Mesh::build_neighbours() {
  const int U = (int)region.size();
  neig.assign((size_t)U * 3, 0);
  fneig.assign((size_t)U * 3, 0);
  dir.assign((size_t)U * 3, 0);
  struct Slot { int tri, side; };
  std::unordered_map<EKey, Slot, EHash> open;
  open.reserve((size_t)U * 2);
  for (int u = 0; u < U; ++u) {
    const double* Xu = &X[(size_t)u * 6];
    for (int f = 0; f < 3; ++f) {
      VKey p = vkey(Xu + 2 * SIDE[f][0]), q = vkey(Xu + 2 * SIDE[f][1]);
      EKey k = vless(p, q) ? EKey{p, q} : EKey{q, p};
      auto it = open.find(k);
      if (it == open.end()) { open.emplace(k, Slot{u, f}); continue; }
      const int i = it->second.tri, fi = it->second.side;  // i < u : the pair the reference visits as (i, j=u)
      open.erase(it);
      neig[(size_t)i * 3 + fi] = u + 1;
      neig[(size_t)u * 3 + f] = i + 1;
      fneig[(size_t)i * 3 + fi] = f + 1;
      fneig[(size_t)u * 3 + f] = fi + 1;
      const double* Xi = &X[(size_t)i * 6];
      int m[3] = {-1, -1, -1};
      for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b)
          if (vkey(Xi + 2 * a) == vkey(Xu + 2 * b)) { m[a] = b; break; }
      const int d = dir_rule(fi, m);
      dir[(size_t)i * 3 + fi] = d;
      dir[(size_t)u * 3 + f] = d;
    }
  }
}

int Mesh::read_msh(const char* path, Mesh& out) {
  std::ifstream f(path);
  if (!f) return PAMG_ERR_IO;
  auto trim = [](std::string& s) {
    while (!s.empty() && (s.back() == '\r' || s.back() == ' ' || s.back() == '\t')) s.pop_back();
  };
  std::string line;
  if (!std::getline(f, line)) return PAMG_ERR_IO;
  trim(line);
  if (line != "$MeshFormat") return PAMG_ERR_IO;                 // Msh2Tri.F90:173-174
  if (!std::getline(f, line)) return PAMG_ERR_IO;
  {
    std::istringstream is(line);
    double ver = 0; int binary = 1;
    is >> ver >> binary;
    if (!is || binary != 0) return PAMG_ERR_IO;                   // ASCII only, :182-186
  }
  bool found = false;
  while (std::getline(f, line)) { trim(line); if (line == "$Nodes") { found = true; break; } }
  if (!found || !std::getline(f, line)) return PAMG_ERR_IO;
  long nn = std::atol(line.c_str());
  if (nn <= 0) return PAMG_ERR_IO;
  std::vector<double> vx((size_t)nn + 1, 0.0), vy((size_t)nn + 1, 0.0);
  for (long i = 0; i < nn; ++i) {
    if (!std::getline(f, line)) return PAMG_ERR_IO;
    std::istringstream is(line);
    long id; double x, y, z;
    is >> id >> x >> y >> z;
    if (!is || id < 1 || id > nn) return PAMG_ERR_IO;
    vx[id] = x; vy[id] = y;
  }
  found = false;
  while (std::getline(f, line)) { trim(line); if (line == "$Elements") { found = true; break; } }
  if (!found || !std::getline(f, line)) return PAMG_ERR_IO;
  long ne = std::atol(line.c_str());
  if (ne <= 0) return PAMG_ERR_IO;
  out = Mesh();
  for (long i = 0; i < ne; ++i) {
    if (!std::getline(f, line)) return PAMG_ERR_IO;
    std::istringstream is(line);
    std::vector<long> t; long v;
    while (is >> v) t.push_back(v);
    if (t.size() < 3) return PAMG_ERR_IO;
    const long type = t[1];
    // triangle families kept by the reference, Msh2Tri.F90:264-265; first tag = region id, first 3 nodes = vertices
    if (!(type == 2 || type == 9 || type == 20 || type == 21 || type == 23 || type == 24 || type == 25)) continue;
    const long ntags = t[2];
    if ((long)t.size() < 6 + ntags || ntags < 1) return PAMG_ERR_IO;
    for (int a = 0; a < 3; ++a) {
      long id = t[3 + ntags + a];
      if (id < 1 || id > nn) return PAMG_ERR_IO;
      out.X.push_back(vx[id]); out.X.push_back(vy[id]);
    }
    out.region.push_back((int32_t)t[3]);
  }
  if (out.region.empty()) return PAMG_ERR_IO;
  out.build_neighbours();
  return PAMG_OK;
}

// child `ele` (1-based) of a triangle split n times, closed form of get_str_info / get_splitting
// (Msh2Tri.F90:42-58,79-106): row start(r) = 1 + (r-1)(2^(n+1)+1-r).
static void child_vertices(const double* P, int n, int ele, double* x /* [3][2] */) {
  const int b = 1 << (n + 1);
  int r = 1, rem = ele;
  while (rem > b + 1 - 2 * r) { rem -= b + 1 - 2 * r; ++r; }
  const int ipos = rem;
  const double s = (double)(1 << n);
  for (int d = 0; d < 2; ++d) {
    const double v1 = (P[0 + d] - P[4 + d]) / s, v2 = (P[2 + d] - P[4 + d]) / s, o = P[4 + d];
    if (ipos & 1) {
      x[4 + d] = o + (r - 1) * v2 + (ipos / 2) * v1;
      x[2 + d] = o + r * v2 + (ipos / 2) * v1;
      x[0 + d] = o + (r - 1) * v2 + v1 * (ipos / 2 + 1);
    } else {
      x[0 + d] = o + r * v2 + v1 * (ipos / 2 - 1);
      x[2 + d] = o + (r - 1) * v2 + v1 * (ipos / 2);
      x[4 + d] = o + r * v2 + v1 * (ipos / 2);
    }
  }
}

int Mesh::synthetic(int kp, int G, Mesh& out) {
  if (kp < 0 || kp > 12 || G < 1 || (long long)G << (2 * kp) > (1ll << 26)) return PAMG_ERR_ARG;
  out = Mesh();
  const int per = 1 << (2 * kp);
  out.X.resize((size_t)G * per * 6);
  out.region.assign((size_t)G * per, 1);
  for (int g = 0; g < G; ++g) {
    const double sq = (double)(g / 2);
    double P[6];
    if ((g & 1) == 0) { P[0] = sq + 1; P[1] = 0; P[2] = sq; P[3] = 1; P[4] = sq; P[5] = 0; }        // X1,X2,X3
    else              { P[0] = sq; P[1] = 1; P[2] = sq + 1; P[3] = 0; P[4] = sq + 1; P[5] = 1; }
    for (int e = 1; e <= per; ++e) child_vertices(P, kp, e, &out.X[((size_t)g * per + (e - 1)) * 6]);
  }
  out.build_neighbours();
  return PAMG_OK;
}

}  // namespace pamg

// ------------------------------------------------------------------------------ C ABI
using pamg::Mesh;
struct pamg_mesh { Mesh m; };

extern "C" {

int pamg_mesh_read_msh(const char* path, pamg_mesh** out) {
  if (!path || !out) return PAMG_ERR_ARG;
  pamg_mesh* pm = new pamg_mesh();
  int rc = Mesh::read_msh(path, pm->m);
  if (rc != PAMG_OK) { delete pm; *out = nullptr; return rc; }
  *out = pm;
  return PAMG_OK;
}

int pamg_mesh_synthetic(int kp, int G, pamg_mesh** out) {
  if (!out) return PAMG_ERR_ARG;
  pamg_mesh* pm = new pamg_mesh();
  int rc = Mesh::synthetic(kp, G, pm->m);
  if (rc != PAMG_OK) { delete pm; *out = nullptr; return rc; }
  *out = pm;
  return PAMG_OK;
}

// structured triangular mesh of str_explicit (transport_tri.F90:354): no_ele_row triangles per row (odd = pointing up,
// even = pointing down), node coordinates as in str_tri_X_nodes (structured_meshgen.F90:276-298); the neighbour
// arrays come from the same edge hash as for gmsh input and reproduce tri_ele_info2 (:190-272) with the structured face
// numbers 1, 2, 3 appearing as gmsh sides 1, 3, 2.
int pamg_mesh_structured_tri(int no_ele_row, int no_ele_col, double dx, double dy, pamg_mesh** out) {
  if (no_ele_row < 2 || (no_ele_row & 1) || no_ele_col < 1 || !(dx > 0.0) || !(dy > 0.0) || !out) return PAMG_ERR_ARG;
  const int totele = no_ele_row * no_ele_col;
  pamg_mesh* pm = new pamg_mesh();
  pm->m.X.resize((size_t)totele * 6);
  pm->m.region.assign(totele, 0);
  for (int ele = 1; ele <= totele; ++ele) {
    const int row = (ele + no_ele_row - 1) / no_ele_row;
    const int col = ele - no_ele_row * (row - 1);
    double* x = &pm->m.X[(size_t)(ele - 1) * 6];
    if (ele & 1) {
      x[0] = dx * (col / 2 + 1); x[1] = dy * (row - 1);
      x[2] = dx * (col / 2);     x[3] = dy * row;
      x[4] = dx * (col / 2);     x[5] = dy * (row - 1);
    } else {
      x[0] = dx * (col / 2 - 1); x[1] = dy * row;
      x[2] = dx * (col / 2);     x[3] = dy * (row - 1);
      x[4] = dx * (col / 2);     x[5] = dy * row;
    }
  }
  pm->m.build_neighbours();
  *out = pm;
  return PAMG_OK;
}

int pamg_mesh_from_arrays(int U, const double* X, const int32_t* region, pamg_mesh** out) {
  if (U < 1 || !X || !out) return PAMG_ERR_ARG;
  pamg_mesh* pm = new pamg_mesh();
  pm->m.X.assign(X, X + (size_t)U * 6);
  if (region) pm->m.region.assign(region, region + U); else pm->m.region.assign(U, 0);
  pm->m.build_neighbours();
  *out = pm;
  return PAMG_OK;
}

int pamg_mesh_size(const pamg_mesh* m, int* U) {
  if (!m || !U) return PAMG_ERR_ARG;
  *U = (int)m->m.region.size();
  return PAMG_OK;
}

int pamg_mesh_get(const pamg_mesh* m, double* X, int32_t* neig, int32_t* fneig, int32_t* dir, int32_t* region) {
  if (!m) return PAMG_ERR_ARG;
  const Mesh& M = m->m;
  if (X) std::memcpy(X, M.X.data(), M.X.size() * sizeof(double));
  if (neig) std::memcpy(neig, M.neig.data(), M.neig.size() * sizeof(int32_t));
  if (fneig) std::memcpy(fneig, M.fneig.data(), M.fneig.size() * sizeof(int32_t));
  if (dir) std::memcpy(dir, M.dir.data(), M.dir.size() * sizeof(int32_t));
  if (region) std::memcpy(region, M.region.data(), M.region.size() * sizeof(int32_t));
  return PAMG_OK;
}

void pamg_mesh_free(pamg_mesh* m) { delete m; }

}  // extern "C"

// pamg_host.cpp -- C++ host driver above the C ABI (include/pamg.h), mirroring the reference's own driver:
// program Transport_equation (main.F90:16-51) selects a solver with `mode`; mode 9 runs
// Semi_implicit_iterative (transport_tri_semi.F90:14-391: ReadMSH, initial condition by region id, time
// loop, n_multigrid "V-cycles" per step), mode 4 runs unstr_explicit (transport_tri_unstr.F90:413-795) and mode 1 runs
// trans_rec (transport_rect.F90:7) and writes its two dump files.
// The reference hard-codes every parameter and is recompiled to change them; here they are flags whose
// defaults are the literals of main.F90:28,46-47.  Everything numerical happens on the GPU through pamg_*.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "pamg.h"

namespace {

struct Args {
  int mode = 9;
  std::string mesh = "";
  int kp = -1, G = 1;              // --synthetic kp,G instead of a .msh file
  bool literal = true;             // HEAD behaviour (mode 9 default) or --intended
  int n_split = 1, multi_levels = 1, solver = 3, n_smooth = 4, n_multigrid = 2, ntime = 2;
  int region = -1;                 // IC T = 1 where region_id == region (4 for test_sn2, 12 in unstr_explicit)
  double cfl = -1, dx = -1, ux = 0, uy = 0, k = 1.0, tol = 1e-8, theta = 1.0;
  int nits = 2, njac = 10, exact_minv = 0, use_dir = 0, max_cycles = 50;
  bool ntime_given = false, k_given = false, tol_given = false;
  int ner = 20, nec = 2, max_iters = 500;   // str_explicit: no_ele_row, no_ele_col (main.F90:22); Krylov iterations per implicit solve
  double dy = -1;
  int device = 0;
  std::vector<int> devices;        // --gpus N / --devices a,b,..: this ONE process drives several GPUs (pamg_create_multi)
};

double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

bool flag(int& i, int argc, char** argv, const char* name, std::string& out) {
  if (std::strcmp(argv[i], name) != 0) return false;
  if (i + 1 >= argc) { std::fprintf(stderr, "missing value for %s\n", name); std::exit(2); }
  out = argv[++i];
  return true;
}

int die(pamg_handle* h, const char* what, int rc) {
  std::fprintf(stderr, "pamg_host: %s failed with %d: %s\n", what, rc, h ? pamg_last_error(h) : "");
  return 1;
}

}  // namespace

int main(int argc, char** argv) {
  Args a;
  for (int i = 1; i < argc; ++i) {
    std::string v;
    if (flag(i, argc, argv, "--mode", v)) a.mode = std::atoi(v.c_str());
    else if (flag(i, argc, argv, "--mesh", v)) a.mesh = v;
    else if (flag(i, argc, argv, "--synthetic", v)) { std::sscanf(v.c_str(), "%d,%d", &a.kp, &a.G); }
    else if (!std::strcmp(argv[i], "--literal")) a.literal = true;
    else if (!std::strcmp(argv[i], "--intended")) a.literal = false;
    else if (flag(i, argc, argv, "--n_split", v)) a.n_split = std::atoi(v.c_str());
    else if (flag(i, argc, argv, "--multi_levels", v)) a.multi_levels = std::atoi(v.c_str());
    else if (flag(i, argc, argv, "--solver", v)) a.solver = std::atoi(v.c_str());
    else if (flag(i, argc, argv, "--n_smooth", v)) a.n_smooth = std::atoi(v.c_str());
    else if (flag(i, argc, argv, "--n_multigrid", v)) a.n_multigrid = std::atoi(v.c_str());
    else if (flag(i, argc, argv, "--ntime", v)) { a.ntime = std::atoi(v.c_str()); a.ntime_given = true; }
    else if (flag(i, argc, argv, "--no_ele_row", v)) a.ner = std::atoi(v.c_str());
    else if (flag(i, argc, argv, "--no_ele_col", v)) a.nec = std::atoi(v.c_str());
    else if (flag(i, argc, argv, "--dy", v)) a.dy = std::atof(v.c_str());
    else if (flag(i, argc, argv, "--max_iters", v)) a.max_iters = std::atoi(v.c_str());
    else if (flag(i, argc, argv, "--region", v)) a.region = std::atoi(v.c_str());
    else if (flag(i, argc, argv, "--cfl", v)) a.cfl = std::atof(v.c_str());
    else if (flag(i, argc, argv, "--dx", v)) a.dx = std::atof(v.c_str());
    else if (flag(i, argc, argv, "--ux", v)) a.ux = std::atof(v.c_str());
    else if (flag(i, argc, argv, "--uy", v)) a.uy = std::atof(v.c_str());
    else if (flag(i, argc, argv, "--k", v)) { a.k = std::atof(v.c_str()); a.k_given = true; }
    else if (flag(i, argc, argv, "--theta", v)) a.theta = std::atof(v.c_str());   // transport_tri_semi.F90:117 (literal 1.)
    else if (flag(i, argc, argv, "--nits", v)) a.nits = std::atoi(v.c_str());
    else if (flag(i, argc, argv, "--njac_its", v)) a.njac = std::atoi(v.c_str());
    else if (flag(i, argc, argv, "--exact_minv", v)) a.exact_minv = std::atoi(v.c_str());
    else if (flag(i, argc, argv, "--use_dir", v)) a.use_dir = std::atoi(v.c_str());
    else if (flag(i, argc, argv, "--tol", v)) { a.tol = std::atof(v.c_str()); a.tol_given = true; }
    else if (flag(i, argc, argv, "--max_cycles", v)) a.max_cycles = std::atoi(v.c_str());
    else if (flag(i, argc, argv, "--device", v)) a.device = std::atoi(v.c_str());
    else if (flag(i, argc, argv, "--gpus", v)) { a.devices.clear(); for (int d = 0; d < std::atoi(v.c_str()); ++d) a.devices.push_back(d); }
    else if (flag(i, argc, argv, "--devices", v)) {
      a.devices.clear();
      for (size_t pos = 0; pos < v.size();) { a.devices.push_back(std::atoi(v.c_str() + pos)); pos = v.find(',', pos); if (pos == std::string::npos) break; ++pos; }
    }
    else { std::fprintf(stderr, "unknown flag %s\n", argv[i]); return 2; }
  }
  // literal defaults of main.F90:28 (mode 4) and :46-47 (mode 9)
  if (a.mode == 4) { if (a.cfl < 0) a.cfl = 0.07; if (a.dx < 0) a.dx = 1e-3; if (a.region < 0) a.region = 12; if (a.ux == 0 && a.uy == 0) a.ux = 0.9; }
  else if (a.mode == 1) { if (a.cfl < 0) a.cfl = 0.7; }
  // main.F90:22 str_explicit(0.7, ..., njac_its 10, nits 2, no_ele_row 20, no_ele_col 2, dx 0.1, dy 0.1, u 0.05 0.05); ntime = 100 (transport_tri.F90:469)
  else if (a.mode == 2) { if (a.cfl < 0) a.cfl = 0.7; if (a.dx < 0) a.dx = 0.1; if (a.ux == 0 && a.uy == 0) { a.ux = 0.05; a.uy = 0.05; } if (!a.ntime_given) a.ntime = 100; }
  // main.F90:31 unstr_implicit(.7, ..., time .7*.1*2, nits 2, dx 0.1, u -.1 .1) on Mesh_files/gmsh_100.msh: dt = CFL dx, ntime = time / dt = 2
  else if (a.mode == 5) { if (a.cfl < 0) a.cfl = 0.7; if (a.dx < 0) a.dx = 0.1; if (a.ux == 0 && a.uy == 0) { a.ux = -0.1; a.uy = 0.1; } }
  else { if (a.cfl < 0) a.cfl = 1.0; if (a.dx < 0) a.dx = a.literal ? 1.25e-5 : 1e-3; if (a.region < 0) a.region = 4; }

  if (a.mode == 1) {
    // case(1) of main.F90:19: call trans_rec(0.7, 2, .false., 10, 250., 2, 2*100, 1, ..., 2*0.01428571, 0.0, .false.)
    // and the two dumps of transport_rect.F90:320-344 (x y t of the four nodes of every element; x t of the analytical pulse)
    const int ner = 200, nec = 1;
    const double CFL = a.cfl > 0 ? a.cfl : 0.7, ux = (a.ux == 0 && a.uy == 0) ? 2 * 0.01428571 : a.ux, uy = a.uy, time = 250.0;
    pamg_params p1;
    pamg_default_params(&p1, 1);
    pamg_handle* h1 = nullptr;
    int rc1 = pamg_create(&p1, a.device, &h1);
    if (rc1 != PAMG_OK) return die(nullptr, "pamg_create (no CUDA device? there is no CPU fallback)", rc1);
    std::vector<double> xa((size_t)ner * nec * 8), t((size_t)ner * nec * 4);
    int ntime = 0;
    if ((rc1 = pamg_trans_rec(h1, CFL, ner, nec, 100.0, 100.0, ux, uy, time, a.nits, a.njac, a.exact_minv, a.literal ? 0 : 1, xa.data(),
                              t.data(), &ntime)))
      return die(h1, "pamg_trans_rec", rc1);
    const std::string dir = a.mesh.empty() ? std::string(".") : a.mesh;      // --mesh doubles as the output directory here
    FILE* f = std::fopen((dir + "/DG-rectangular_structured").c_str(), "w");
    if (!f) { std::fprintf(stderr, "cannot write into %s\n", dir.c_str()); return 1; }
    for (int e = 0; e < ner * nec; ++e)
      for (int i = 0; i < 4; ++i) std::fprintf(f, " %16.9g %16.9g %16.9g\n", xa[(size_t)e * 8 + 2 * i], xa[(size_t)e * 8 + 2 * i + 1], t[(size_t)e * 4 + i]);
    std::fclose(f);
    // analytical pulse: shifted by the distance travelled, in whole elements (:101-105)
    const double dx = 100.0 / ner, dt = CFL * dx;
    const int off = (int)(ux * dt * ntime * ner / 100.0 + 1);
    f = std::fopen((dir + "/DG-rectangular_structured_analytical").c_str(), "w");
    if (!f) { std::fprintf(stderr, "cannot write into %s\n", dir.c_str()); return 1; }
    for (int e = 1; e <= ner * nec; ++e)
      for (int i = 0; i < 4; ++i)
        std::fprintf(f, " %16.9g %16.9g\n", xa[(size_t)(e - 1) * 8 + 2 * i], (e >= off + ner / 5 && e <= off + ner / 2) ? 1.0 : 0.0);
    std::fclose(f);
    double sum = 0, mx = -1e300, mn = 1e300;
    for (double v : t) { sum += v; mx = std::fmax(mx, v); mn = std::fmin(mn, v); }
    std::printf("trans_rec: ntime = %d\ntnew: sum %.12e min %.6e max %.6e\n", ntime, sum, mn, mx);
    pamg_destroy(h1);
    return 0;
  }

  std::printf("---------------------------------------------------------\n|       Reading the .msh file     |\n");
  const double t0 = now();
  pamg_mesh* mesh = nullptr;
  // str_explicit: totele = no_ele_row * no_ele_col * 2 (transport_tri.F90:400), no_ele_row triangles per row (str_tri_X_nodes)
  int rc = a.mode == 2 ? pamg_mesh_structured_tri(a.ner, 2 * a.nec, a.dx, a.dy > 0 ? a.dy : a.dx, &mesh)
           : a.kp >= 0 ? pamg_mesh_synthetic(a.kp, a.G, &mesh) : pamg_mesh_read_msh(a.mesh.c_str(), &mesh);
  if (rc != PAMG_OK) { std::fprintf(stderr, "cannot read mesh '%s' (%d)\n", a.mesh.c_str(), rc); return 1; }
  int U = 0;
  pamg_mesh_size(mesh, &U);
  std::vector<double> X((size_t)U * 6);
  std::vector<int32_t> neig((size_t)U * 3), fneig((size_t)U * 3), dir((size_t)U * 3), region(U);
  pamg_mesh_get(mesh, X.data(), neig.data(), fneig.data(), dir.data(), region.data());
  pamg_mesh_free(mesh);
  std::printf("|   Time for reading .msh file    | %g\n", now() - t0);

  pamg_params p;
  pamg_default_params(&p, a.literal ? 1 : 0);
  p.n_split = a.n_split; p.multi_levels = a.multi_levels; p.solver = a.solver; p.n_smooth = a.n_smooth;
  p.n_multigrid = a.n_multigrid; p.dt = a.cfl * a.dx; p.k = a.k; p.u_x = a.ux; p.u_y = a.uy; p.theta = a.theta;
  p.source_coef = (a.literal ? -2.0 : 2.0) * a.k;
  pamg_handle* h = nullptr;
  // the reference's driver is one serial process (main.F90:16-51): with --gpus N the same single host thread drives N
  // devices through one handle; the library cuts the parents into N contiguous blocks and every call below stays as it is
  if (a.devices.size() > 1 && a.mode == 9) rc = pamg_create_multi(&p, (int)a.devices.size(), a.devices.data(), &h);
  else rc = pamg_create(&p, a.devices.size() == 1 ? a.devices[0] : a.device, &h);
  if (rc != PAMG_OK) return die(nullptr, "pamg_create (no CUDA device? there is no CPU fallback)", rc);
  if (a.devices.size() > 1) std::printf("|   GPUs driven by this process = %zu\n", a.devices.size());

  if (a.mode == 9) {
    if ((rc = pamg_set_parents(h, U, X.data(), neig.data(), fneig.data(), dir.data()))) return die(h, "pamg_set_parents", rc);
    int64_t ndof = 0;
    pamg_ndof(h, 1, &ndof);
    const int64_t C = ndof / 3 / U;
    std::printf("|   n_split = %d\n|   multigrid levels = %d\n|   totele_unst, totele_str, totele %d %lld %lld\n|   ntime = %d\n|   dt    = %g\n",
                p.n_split, p.multi_levels, U, (long long)C, (long long)(C * U), a.ntime, p.dt);
    std::printf("---------------------------------------------------------\n");
    std::vector<double> T((size_t)ndof, 0.0);
    for (int u = 0; u < U; ++u)
      if (region[u] == a.region) for (int64_t i = 0; i < 3 * C; ++i) T[(size_t)u * 3 * C + i] = 1.0;   // :249-251
    if ((rc = pamg_upload_field(h, PAMG_TNEW, 1, T.data()))) return die(h, "upload", rc);
    pamg_sync(h);
    const double t1 = now();
    for (int it = 1; it <= a.ntime; ++it) {
      if (a.literal) {
        if ((rc = pamg_literal_timestep(h, p.solver, p.n_multigrid, p.n_smooth))) return die(h, "pamg_literal_timestep", rc);
      } else {
        // told = tnew ; tnew_nonlin = tnew ; V-cycles to tolerance
        if ((rc = pamg_copy_field(h, 1, PAMG_TOLD, PAMG_TNEW))) return die(h, "copy", rc);
        if ((rc = pamg_copy_field(h, 1, PAMG_TNONLIN, PAMG_TNEW))) return die(h, "copy", rc);
        int cycles = 0;
        std::vector<double> hist((size_t)a.max_cycles + 2);
        if ((rc = pamg_vcycle_solve(h, p.solver, p.n_smooth, p.n_smooth, p.n_coarse_smooth, a.max_cycles, a.tol, &cycles, hist.data())))
          return die(h, "pamg_vcycle_solve", rc);
        if ((rc = pamg_copy_field(h, 1, PAMG_TNEW, PAMG_TNONLIN))) return die(h, "copy", rc);
        std::printf(" V-cycles %d  ||r||/||r0|| %.3e\n", cycles, hist[0] > 0 ? hist[cycles > a.max_cycles ? a.max_cycles : cycles] / hist[0] : 0.0);
      }
      std::printf(" semi %d\n", it);
    }
    pamg_sync(h);
    const double t2 = now();
    double l2 = 0, linf = 0;
    pamg_update_overlaps(h, 1);
    pamg_residual(h, 1, &l2, &linf);
    pamg_download_field(h, PAMG_TNEW, 1, T.data());
    double sum = 0, mx = -1e300, mn = 1e300;
    for (double v : T) { sum += v; mx = std::fmax(mx, v); mn = std::fmin(mn, v); }
    std::printf("----------------------------------------------------------\n|        cpu_time for time_loop = %g |\n", t2 - t1);
    std::printf("|        ||r||2 = %.6e  ||r||inf = %.6e\n|        tnew: sum %.12e min %.6e max %.6e\n", l2, linf, sum, mn, mx);
    std::printf("----------------------------------------------------------\n");
  } else if (a.mode == 4) {
    if ((rc = pamg_set_unstructured(h, U, X.data(), neig.data(), fneig.data()))) return die(h, "pamg_set_unstructured", rc);
    std::vector<double> T((size_t)U * 3, 0.0);
    for (int e = 0; e < U; ++e) if (region[e] == a.region) T[3 * e] = T[3 * e + 1] = T[3 * e + 2] = 1.0;   // :550-552
    std::printf("totele = %d\nntime = %d\n", U, a.ntime);
    if ((rc = pamg_unstr_upload(h, T.data()))) return die(h, "upload", rc);
    pamg_sync(h);
    const double t1 = now();
    if ((rc = pamg_explicit_step(h, a.cfl * a.dx, a.ux, a.uy, 0.0, a.ntime, a.nits, a.njac, a.exact_minv, a.use_dir)))
      return die(h, "pamg_explicit_step", rc);
    pamg_sync(h);
    const double t2 = now();
    pamg_unstr_download(h, T.data());
    double sum = 0, mx = -1e300, mn = 1e300;
    for (double v : T) { sum += v; mx = std::fmax(mx, v); mn = std::fmin(mn, v); }
    std::printf("cpu_time for time_loop = %g\ntnew: sum %.12e min %.6e max %.6e\n", t2 - t1, sum, mn, mx);
  } else if (a.mode == 2) {
    // str_explicit (transport_tri.F90:354): the structured triangles through the same explicit step; t_bc = 0, u_bc = u
    // (:438-440,465); initial pulse as at :457-460 (the first no_ele_row/5 + 1 elements of row 1, the first fifth of every
    // other row)
    if ((rc = pamg_set_unstructured(h, U, X.data(), neig.data(), fneig.data()))) return die(h, "pamg_set_unstructured", rc);
    std::vector<double> T((size_t)U * 3, 0.0);
    auto one = [&](int e1) { if (e1 >= 1 && e1 <= U) T[3 * (e1 - 1)] = T[3 * (e1 - 1) + 1] = T[3 * (e1 - 1) + 2] = 1.0; };
    for (int e = 1; e <= a.ner / 5 + 1; ++e) one(e);
    for (int i = 2; i <= a.nec; ++i) for (int e = (i - 1) * a.ner + 1; e <= a.ner * i - (a.ner * 4 / 5); ++e) one(e);
    std::printf("totele = %d\nntime = %d\n", U, a.ntime);
    if ((rc = pamg_unstr_upload(h, T.data()))) return die(h, "upload", rc);
    pamg_sync(h);
    const double t1 = now();
    if ((rc = pamg_explicit_step(h, a.cfl * a.dx, a.ux, a.uy, 0.0, a.ntime, a.nits, a.njac, a.exact_minv, 0)))
      return die(h, "pamg_explicit_step", rc);
    pamg_sync(h);
    const double t2 = now();
    pamg_unstr_download(h, T.data());
    double sum = 0, mx = -1e300, mn = 1e300;
    for (double v : T) { sum += v; mx = std::fmax(mx, v); mn = std::fmin(mn, v); }
    std::printf("cpu_time for time_loop = %g\ntnew: sum %.12e min %.6e max %.6e\n", t2 - t1, sum, mn, mx);
  } else if (a.mode == 5) {
    // unstr_implicit (transport_tri_unstr.F90:7-408): block-CSR operator on the device instead of three CSR matrices -> dense
    // -> FINDInv (:366-378), a Krylov solve per nonlinear pass; IC tnew(:,5) = 1 (:163) unless --region selects elements
    if ((rc = pamg_set_unstructured(h, U, X.data(), neig.data(), fneig.data()))) return die(h, "pamg_set_unstructured", rc);
    std::vector<double> T((size_t)U * 3, 0.0);
    if (a.region >= 0) { for (int e = 0; e < U; ++e) if (region[e] == a.region) T[3 * e] = T[3 * e + 1] = T[3 * e + 2] = 1.0; }
    else if (U >= 5) T[12] = T[13] = T[14] = 1.0;
    std::printf("totele = %d\nntime = %d\n", U, a.ntime);
    const double dt = a.cfl * a.dx;
    if ((rc = a.k_given ? pamg_implicit_assemble_diffusion(h, dt, a.ux, a.uy, a.k, a.use_dir) : pamg_implicit_assemble(h, dt, a.ux, a.uy, a.use_dir)))
      return die(h, "pamg_implicit_assemble", rc);
    if ((rc = pamg_unstr_upload(h, T.data()))) return die(h, "upload", rc);
    pamg_sync(h);
    const double t1 = now();
    int iters = 0; double relres = 0.0;
    if ((rc = pamg_implicit_step(h, a.ntime, a.nits, a.tol_given ? a.tol : 1e-13, a.max_iters, &iters, &relres))) return die(h, "pamg_implicit_step", rc);
    pamg_sync(h);
    const double t2 = now();
    pamg_unstr_download(h, T.data());
    double sum = 0, mx = -1e300, mn = 1e300;
    for (double v : T) { sum += v; mx = std::fmax(mx, v); mn = std::fmin(mn, v); }
    std::printf("Krylov iterations %d  worst ||r||/||b|| %.3e\ncpu_time for time_loop = %g\ntnew: sum %.12e min %.6e max %.6e\n",
                iters, relres, t2 - t1, sum, mn, mx);
  } else {
    std::fprintf(stderr, "mode %d is outside the hot path (modes 1, 2, 4, 5 and 9 are implemented; see DESIGN.md)\n", a.mode);
    pamg_destroy(h);
    return 2;
  }
  pamg_destroy(h);
  return 0;
}

"""ctypes binding of oracle/liboracle.so (CPU oracle, test infrastructure only).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product (p-a_multigrids_b200) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "liboracle.so")

f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


class OrcParams(C.Structure):
    _fields_ = [
        ("n_split", C.c_int32), ("multi_levels", C.c_int32), ("face_terms", C.c_int32),
        ("literal_source", C.c_int32), ("transfer", C.c_int32), ("residual_sign", C.c_int32),
        ("halo_rule", C.c_int32), ("coarse_bc_zero", C.c_int32),
        ("theta", C.c_double), ("dt", C.c_double), ("k", C.c_double), ("omega", C.c_double),
        ("u_x", C.c_double), ("u_y", C.c_double), ("source_coef", C.c_double),
    ]


def build(force=False):
    src = [os.path.join(ORACLE_DIR, f) for f in ("pamg_oracle.cpp", "pamg_oracle.h", "Makefile")]
    if force or not os.path.exists(LIB_PATH) or any(
            os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in src):
        subprocess.run(["make", "-C", ORACLE_DIR], check=True, stdout=subprocess.DEVNULL)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(LIB_PATH)
    L.orc_tables.argtypes = [f64p] * 6
    L.orc_get_str_info.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.orc_get_splitting.argtypes = [f64p, C.c_int, C.c_int, f64p]
    L.orc_str_neig.argtypes = [C.c_int, i32p]
    L.orc_surf_ele.argtypes = [C.c_int, i32p]
    L.orc_element_conversion.argtypes = [C.c_int, C.c_int, i32p]
    L.orc_tri_det_nlx.argtypes = [f64p, f64p, f64p]
    L.orc_face_geometry.argtypes = [f64p, C.c_int, f64p, f64p]
    L.orc_findinv.argtypes = [f64p, f64p, C.c_int]
    L.orc_findinv.restype = C.c_int
    L.orc_read_msh.argtypes = [C.c_char_p, C.c_int, f64p, i32p, i32p, i32p]
    L.orc_read_msh.restype = C.c_int
    L.orc_neig_data.argtypes = [C.c_int, i32p, i32p, i32p, i32p]
    L.orc_semi_create.argtypes = [C.POINTER(OrcParams), C.c_int, f64p, i32p, i32p, i32p]
    L.orc_semi_create.restype = C.c_void_p
    L.orc_semi_destroy.argtypes = [C.c_void_p]
    L.orc_semi_set_boundary.argtypes = [C.c_void_p, i32p, f64p]
    L.orc_semi_field.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.orc_semi_field.restype = C.POINTER(C.c_double)
    L.orc_semi_overlap.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.orc_semi_overlap.restype = C.POINTER(C.c_double)
    L.orc_semi_ndof.argtypes = [C.c_void_p, C.c_int]
    L.orc_semi_ndof.restype = C.c_int64
    L.orc_semi_update_overlaps.argtypes = [C.c_void_p, C.c_int]
    L.orc_semi_smooth.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
    L.orc_semi_build_rhs.argtypes = [C.c_void_p]
    L.orc_semi_residual.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.orc_semi_convergence.argtypes = [C.c_void_p, C.c_int]
    L.orc_semi_convergence.restype = C.c_double
    L.orc_semi_restrict.argtypes = [C.c_void_p, C.c_int]
    L.orc_semi_prolong.argtypes = [C.c_void_p, C.c_int]
    L.orc_semi_vcycle_solve.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, f64p]
    L.orc_semi_vcycle_solve.restype = C.c_int
    L.orc_semi_literal_timestep.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
    L.orc_semi_set_threads.argtypes = [C.c_int]
    L.orc_unstr_explicit.argtypes = [C.c_int, f64p, i32p, i32p, i32p, C.c_double, C.c_double, C.c_double,
                                     C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, f64p]
    L.orc_unstr_implicit_assemble.argtypes = [C.c_int, f64p, i32p, i32p, C.c_double, C.c_double, C.c_double, C.c_int, f64p, f64p]
    L.orc_unstr_implicit_assemble_diff.argtypes = [C.c_int, f64p, i32p, i32p, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int, f64p, f64p]
    L.orc_unstr_implicit.argtypes = [C.c_int, f64p, i32p, i32p, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, f64p]
    L.orc_unstr_implicit.restype = C.c_int
    L.orc_unstr_stab.argtypes = [C.c_int, f64p, f64p, f64p, C.c_double, C.c_double, C.c_double, f64p, f64p]
    L.orc_unstr_implicit_stab.argtypes = [C.c_int, f64p, i32p, i32p, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, f64p]
    L.orc_unstr_implicit_stab.restype = C.c_int
    L.orc_trans_rec.argtypes = [C.c_double, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double,
                                C.c_int, C.c_int, C.c_int, C.c_int, f64p, f64p]
    L.orc_trans_rec.restype = C.c_int
    L.orc_rect_analytical.argtypes = [C.c_double, C.c_int, C.c_double, C.c_double, C.c_double, f64p, f64p]
    L.orc_thermal_analytical.argtypes = [C.c_double] * 4
    L.orc_thermal_analytical.restype = C.c_double
    _lib = L
    return L


# field ids
TNEW, TOLD, RHS, RES, SRC, TNONLIN = range(6)


def literal_params(n_split, multi_levels=1, dt=1.25e-5, k=1.0, omega=0.8, u=(0.0, 0.0)):
    """What HEAD computes (transport_tri_semi.F90 as checked in)."""
    return OrcParams(n_split=n_split, multi_levels=multi_levels, face_terms=0, literal_source=1, transfer=0,
                     residual_sign=1, halo_rule=0, coarse_bc_zero=0, theta=1.0, dt=dt, k=k, omega=omega,
                     u_x=u[0], u_y=u[1], source_coef=-2.0 * k)


def intended_params(n_split, multi_levels=None, dt=1e-3, k=1.0, omega=0.8, u=(0.0, 0.0), source_coef=None):
    """The mathematically intended composition."""
    return OrcParams(n_split=n_split, multi_levels=multi_levels or n_split, face_terms=1, literal_source=0,
                     transfer=1, residual_sign=-1, halo_rule=1, coarse_bc_zero=1, theta=1.0, dt=dt, k=k,
                     omega=omega, u_x=u[0], u_y=u[1],
                     source_coef=(2.0 * k if source_coef is None else source_coef))


def read_msh(path, max_tri=200000):
    X = np.zeros((max_tri, 3, 2)); neig = np.zeros((max_tri, 3), np.int32)
    dirv = np.zeros((max_tri, 3), np.int32); region = np.zeros(max_tri, np.int32)
    n = lib().orc_read_msh(path.encode(), max_tri, X, neig, dirv, region)
    if n < 0:
        raise RuntimeError(f"orc_read_msh({path}) -> {n}")
    return dict(X=X[:n].copy(), neig=neig[:n].copy(), dir=dirv[:n].copy(), region=region[:n].copy())


def neig_data(neig, dirv):
    U = neig.shape[0]
    fneig = np.zeros((U, 3), np.int32); snodes = np.zeros((U, 3, 2), np.int32)
    lib().orc_neig_data(U, np.ascontiguousarray(neig), np.ascontiguousarray(dirv), fneig, snodes)
    return fneig, snodes


class Semi:
    """Semi-structured multigrid problem held by the oracle."""

    def __init__(self, params, X, neig, fneig, dirv):
        self.L = lib()
        self.params = params
        self.U = X.shape[0]
        self.X = np.ascontiguousarray(X, np.float64)
        self.h = self.L.orc_semi_create(C.byref(params), self.U, self.X,
                                        np.ascontiguousarray(neig, np.int32),
                                        np.ascontiguousarray(fneig, np.int32),
                                        np.ascontiguousarray(dirv, np.int32))
        if not self.h:
            raise RuntimeError("orc_semi_create failed (multi_levels > n_split?)")

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_semi_destroy(self.h)
            self.h = None

    def set_boundary(self, kind, value):
        self.L.orc_semi_set_boundary(self.h, np.ascontiguousarray(kind, np.int32).reshape(-1),
                                     np.ascontiguousarray(value, np.float64).reshape(-1))

    def split(self, level):
        return self.params.n_split - level + 1

    def field(self, fid, level=1):
        """numpy view (U, C, 3) of the oracle's own storage."""
        n = self.L.orc_semi_ndof(self.h, level)
        p = self.L.orc_semi_field(self.h, fid, level)
        a = np.ctypeslib.as_array(p, shape=(n,))
        return a.reshape(self.U, 4 ** self.split(level), 3)

    def overlap(self, level=1, old=False):
        S = 2 ** self.split(level)
        p = self.L.orc_semi_overlap(self.h, level, 1 if old else 0)
        return np.ctypeslib.as_array(p, shape=(self.U * 9 * S,)).reshape(self.U, 3, S, 3)

    def update_overlaps(self, level=1):
        self.L.orc_semi_update_overlaps(self.h, level)

    def smooth(self, level, solver, nsweeps):
        self.L.orc_semi_smooth(self.h, level, solver, nsweeps)

    def build_rhs(self):
        self.L.orc_semi_build_rhs(self.h)

    def residual(self, level=1):
        l2 = C.c_double(); li = C.c_double()
        self.L.orc_semi_residual(self.h, level, C.byref(l2), C.byref(li))
        return l2.value, li.value

    def convergence(self, level=1):
        return self.L.orc_semi_convergence(self.h, level)

    def restrict(self, fine_level):
        self.L.orc_semi_restrict(self.h, fine_level)

    def prolong(self, fine_level):
        self.L.orc_semi_prolong(self.h, fine_level)

    def vcycle_solve(self, solver=1, nu1=4, nu2=4, ncoarse=15, max_cycles=50, tol=1e-8):
        hist = np.zeros(max_cycles + 2)
        it = self.L.orc_semi_vcycle_solve(self.h, solver, nu1, nu2, ncoarse, max_cycles, tol, hist)
        return it, hist[: min(it, max_cycles) + 1]

    def literal_timestep(self, solver=3, n_multigrid=2, n_smooth=4):
        self.L.orc_semi_literal_timestep(self.h, solver, n_multigrid, n_smooth)


def synthetic_mesh(kp, G=1, tmpdir=None):
    """The synthetic input of SURVEY 8(d) built WITHOUT the product: G right super-triangles (G/2 unit squares in a strip), each
    split kp times with the reference's get_splitting numbering (Msh2Tri.F90:69-107) into 4**kp parents; neighbours by the
    oracle's restatement of ReadMSH / CheckNeig (all-pairs, fine for a few thousand parents).  Same vertices and parent
    order as pamg_mesh_synthetic.  Returns dict(X, neig, fneig, dir, region)."""
    import tempfile
    per = 4 ** kp
    X = np.zeros((G * per, 3, 2))
    out = np.zeros(6)
    L = lib()
    for g in range(G):
        sq = float(g // 2)
        if g % 2 == 0:
            P = np.array([sq + 1, 0, sq, 1, sq, 0], np.float64)          # X1, X2, X3
        else:
            P = np.array([sq, 1, sq + 1, 0, sq + 1, 1], np.float64)
        for e in range(1, per + 1):
            L.orc_get_splitting(P, kp, e, out)
            X[g * per + e - 1] = out.reshape(3, 2)
    ids, nodes, tris = {}, [], []
    for t in X:
        row = []
        for p in t:
            key = (float(p[0]), float(p[1]))
            if key not in ids:
                ids[key] = len(nodes) + 1
                nodes.append(key)
            row.append(ids[key])
        tris.append(row)
    fd, path = tempfile.mkstemp(suffix=".msh", dir=tmpdir)
    with os.fdopen(fd, "w") as f:
        f.write("$MeshFormat\n2.2 0 8\n$EndMeshFormat\n$Nodes\n%d\n" % len(nodes))
        for i, (x, y) in enumerate(nodes):
            f.write("%d %.17g %.17g 0\n" % (i + 1, x, y))
        f.write("$EndNodes\n$Elements\n%d\n" % len(tris))
        for i, t in enumerate(tris):
            f.write("%d 2 2 1 1 %d %d %d\n" % (i + 1, t[0], t[1], t[2]))
        f.write("$EndElements\n")
    try:
        m = read_msh(path, max_tri=len(tris) + 8)
    finally:
        os.unlink(path)
    m["fneig"], _ = neig_data(m["neig"], m["dir"])
    return m

"""Pins the CPU oracle against every known answer the reference formulas imply,
the reconstructible golden dump and geometric invariants.  CPU only."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle_api as orc
from helpers import GOLDEN, mesh_triangles, write_msh

UNIT = np.array([[1.0, 0.0], [0.0, 1.0], [0.0, 0.0]])  # X1, X2, X3


def test_tables():
    n = np.zeros(9); nlx = np.zeros(18); w = np.zeros(3); sn = np.zeros(4); snlx = np.zeros(4); sw = np.zeros(2)
    orc.lib().orc_tables(n, nlx, w, sn, snlx, sw)
    assert np.allclose(n.reshape(3, 3), [[.5, .5, 0], [0, .5, .5], [.5, 0, .5]], atol=0)
    assert np.allclose(w, 1 / 3)
    assert np.allclose(nlx.reshape(3, 2, 3)[0], [[1, 0, -1], [0, 1, -1]])
    assert np.allclose(sn.reshape(2, 2), [[0.788675134594813, 0.211324865405187],
                                         [0.211324865405187, 0.788675134594813]], atol=1e-15)
    assert np.allclose(snlx.reshape(2, 2), [[-.5, .5], [-.5, .5]])
    assert np.allclose(sw, 1)


def test_unit_triangle_mass_stiffness():
    nx = np.zeros(18); dw = np.zeros(3)
    orc.lib().orc_tri_det_nlx(UNIT.copy(), nx, dw)
    assert np.allclose(dw, 1 / 6)
    n = np.array([[.5, .5, 0], [0, .5, .5], [.5, 0, .5]])
    M = np.einsum("gi,g,gj->ij", n, dw, n)
    assert np.allclose(M, np.array([[2, 1, 1], [1, 2, 1], [1, 1, 2]]) / 24)
    nxa = nx.reshape(3, 2, 3)
    K = np.einsum("gdi,g,gdj->ij", nxa, dw, nxa)
    assert np.allclose(K, 0.5 * np.array([[1, 0, -1], [0, 1, -1], [-1, -1, 2]]))
    Minv = np.zeros(9)
    assert orc.lib().orc_findinv(np.ascontiguousarray(M.ravel()), Minv, 3) == 0
    assert np.allclose(Minv.reshape(3, 3), [[18, -6, -6], [-6, 18, -6], [-6, -6, 18]])


def test_findinv_singular_and_zero_pivot():
    out = np.zeros(4)
    assert orc.lib().orc_findinv(np.array([1.0, 2.0, 2.0, 4.0]), out, 2) == -1
    assert np.all(out == 0)
    # zero leading pivot is repaired by a row ADD (no swap), matrices.F90:1661-1676
    A = np.array([[0.0, 1.0], [1.0, 1.0]])
    assert orc.lib().orc_findinv(A.ravel().copy(), out, 2) == 0
    assert np.allclose(out.reshape(2, 2) @ A, np.eye(2))
    rng = np.random.default_rng(0)
    for n in (3, 4, 6):
        A = rng.random((n, n)) + n * np.eye(n)
        o = np.zeros(n * n)
        assert orc.lib().orc_findinv(A.ravel().copy(), o, n) == 0
        assert np.allclose(o.reshape(n, n), np.linalg.inv(A), rtol=1e-12)


def test_face_geometry_unit_triangle():
    # gmsh faces: 1=(X1,X3) 2=(X1,X2) 3=(X2,X3)
    exp_len = [1.0, np.sqrt(2.0), 1.0]
    exp_n = [[0, -1], [1 / np.sqrt(2), 1 / np.sqrt(2)], [-1, 0]]
    for f in (1, 2, 3):
        sd = np.zeros(2); sn = np.zeros(4)
        orc.lib().orc_face_geometry(UNIT.copy(), f, sd, sn)
        assert np.allclose(sd, exp_len[f - 1] / 2)
        assert np.allclose(sn.reshape(2, 2), [exp_n[f - 1]] * 2, atol=1e-15)
    # face mass (L/6)[[2,1],[1,2]]
    s = np.array([[0.788675134594813, 0.211324865405187], [0.211324865405187, 0.788675134594813]])
    assert np.allclose(np.einsum("si,sj->ij", s, s) * 0.5, np.array([[2, 1], [1, 2]]) / 6)


def test_str_neig_n2_known_table():
    t = np.zeros(3 * 16, np.int32)
    orc.lib().orc_str_neig(2, t)
    t = t.reshape(16, 3)
    assert list(t[:, 0]) == [0, 8, 0, 10, 0, 12, 0, 2, 13, 4, 15, 6, 9, 16, 11, 14]
    assert list(t[:, 1]) == [0, 3, 2, 5, 4, 7, 6, 0, 10, 9, 12, 11, 0, 15, 14, 0]
    assert list(t[:, 2]) == [2, 1, 4, 3, 6, 5, 0, 9, 8, 11, 10, 0, 14, 13, 0, 0]


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5])
def test_numbering_closed_form_and_neighbour_symmetry(n):
    L = orc.lib()
    Cn = 4 ** n
    b = 2 ** (n + 1)
    t = np.zeros(3 * Cn, np.int32)
    L.orc_str_neig(n, t)
    t = t.reshape(Cn, 3)
    ir, ip, ori = C.c_int(), C.c_int(), C.c_int()
    tri = {}
    for ele in range(1, Cn + 1):
        L.orc_get_str_info(n, ele, C.byref(ir), C.byref(ip), C.byref(ori))
        r = ir.value
        assert ele == 1 + (r - 1) * (b + 1 - r) + ip.value - 1          # closed-form row start
        assert ori.value == ip.value % 2
        x = np.zeros((3, 2))
        L.orc_get_splitting(UNIT.copy(), n, ele, x)
        tri[ele] = x.copy()
        area = 0.5 * abs(np.linalg.det(np.c_[x[0] - x[2], x[1] - x[2]]))
        assert np.isclose(area, 0.5 / Cn)
    fn = [(1, 3), (3, 2), (2, 1)]
    for ele in range(1, Cn + 1):
        for f in range(3):
            nb = t[ele - 1, f]
            if nb == 0:
                continue
            assert t[nb - 1, f] == ele                                   # same face number on both sides
            a, bb = fn[f]
            # shared nodes coincide in reversed order
            assert np.allclose(tri[ele][a - 1], tri[nb][bb - 1])
            assert np.allclose(tri[ele][bb - 1], tri[nb][a - 1])


def test_surf_ele_and_boundary_children():
    n = 2
    s = np.zeros(3 * 4, np.int32)
    orc.lib().orc_surf_ele(n, s)
    s = s.reshape(3, 4)
    assert list(s[0]) == [1, 3, 5, 7]      # face 1
    assert list(s[1]) == [7, 12, 15, 16]   # face 2: last child of each row
    assert list(s[2]) == [1, 8, 13, 16]    # face 3: first child of each row


def test_element_conversion_known():
    f = np.zeros(4, np.int32)
    L = orc.lib()
    L.orc_element_conversion(1, 1, f); assert list(f) == [1, 2, 3, 8]
    L.orc_element_conversion(2, 1, f); assert list(f) == [11, 10, 9, 4]
    L.orc_element_conversion(3, 1, f); assert list(f) == [5, 6, 7, 12]
    L.orc_element_conversion(1, 0, f); assert list(f) == [1, 2, 3, 4]


@pytest.mark.parametrize("s", [0, 1, 2, 3])
def test_element_conversion_partitions_and_interpolates_coordinates(s):
    """Fine children of a coarse child tile it, and P1 interpolation of the coarse vertex coordinates
    reproduces the fine vertex coordinates (pins the prolongation weights PW)."""
    L = orc.lib()
    PW = np.array([[[.5, 0, .5], [0, .5, .5], [0, 0, 1]],
                   [[0, .5, .5], [.5, 0, .5], [.5, .5, 0]],
                   [[1, 0, 0], [.5, .5, 0], [.5, 0, .5]],
                   [[.5, .5, 0], [0, 1, 0], [0, .5, .5]]])
    X = np.array([[2.0, 0.3], [0.4, 1.7], [-0.2, 0.1]])
    seen = set()
    f = np.zeros(4, np.int32)
    for c in range(1, 4 ** s + 1):
        L.orc_element_conversion(c, s, f)
        xc = np.zeros((3, 2)); L.orc_get_splitting(X.copy(), s, c, xc)
        for k in range(4):
            assert f[k] not in seen
            seen.add(int(f[k]))
            xf = np.zeros((3, 2)); L.orc_get_splitting(X.copy(), s + 1, int(f[k]), xf)
            assert np.allclose(xf, PW[k] @ xc, atol=1e-14)
    assert seen == set(range(1, 4 ** (s + 1) + 1))


def test_rect_analytical_dump_matches_reference_file():
    g = np.load(os.path.join(GOLDEN, "rect_golden.npz"))
    ana = g["analytical"]
    x = np.zeros(800); t = np.zeros(800)
    # main.F90:19 : CFL .7, 200x1 elements on 100x100, u_x = 2*0.01428571, time 250
    orc.lib().orc_rect_analytical(0.7, 200, 100.0, 2 * 0.01428571, 250.0, x, t)
    assert np.allclose(x, ana[:, 0], atol=1e-5)
    assert np.array_equal(t, ana[:, 1])
    assert np.flatnonzero(t)[0] == 216 and np.count_nonzero(t) == 244
    num = g["numerical"]
    assert num.shape == (800, 3) and -0.2504 < num[:, 2].min() and num[:, 2].max() < 2.2504


def test_trans_rec_reproduces_the_reference_output_file():
    """The one pin against an OUTPUT OF THE REFERENCE BINARY: DG-rectangular_structured, written by trans_rec with the
    arguments of main.F90:19.  The restatement of transport_rect.F90 (Q1 shape functions, structured grid, upwind
    face flux, told / tnew_nonlin handling, 10 Jacobi iterations on the lumped mass, 714 time steps of 2 passes)
    reproduces the file to the reference's single precision - with the HEAD quirk that tnew_gi is never set (:157 is
    commented out), i.e. without the advection volume integral.  The intended scheme does NOT reproduce it."""
    g = np.load(os.path.join(GOLDEN, "rect_golden.npz"))
    num = g["numerical"]
    x = np.zeros((200, 4, 2)); t = np.zeros((200, 4))
    nt = orc.lib().orc_trans_rec(0.7, 200, 1, 100.0, 100.0, 2 * 0.01428571, 0.0, 250.0, 2, 10, 0, 0, x, t)
    assert nt == 714
    assert np.array_equal(x.reshape(-1, 2), num[:, :2])
    assert np.max(np.abs(t.ravel() - num[:, 2])) <= 3e-5          # fp32 accumulated over 1428 passes; max |t| = 2.25
    assert np.max(np.abs(t.ravel() - num[:, 2])) <= 2e-5 * np.max(np.abs(num[:, 2]))
    ti = np.zeros((200, 4))
    orc.lib().orc_trans_rec(0.7, 200, 1, 100.0, 100.0, 2 * 0.01428571, 0.0, 250.0, 2, 10, 0, 1, x, ti)
    assert np.max(np.abs(ti.ravel() - num[:, 2])) > 1.0           # the intended scheme is a different answer ...
    assert -0.05 < ti.min() and ti.max() < 1.05                    # ... a bounded pulse ...
    assert abs(ti.sum() - 61 * 4) <= 1e-9                          # ... that conserves mass (61 elements of 1)
    # the element-local direct solve (FINDInv of the 4x4 mass matrix) and 10 Jacobi iterations agree to their tolerance
    td = np.zeros((200, 4))
    orc.lib().orc_trans_rec(0.7, 200, 1, 100.0, 100.0, 2 * 0.01428571, 0.0, 250.0, 2, 10, 1, 1, x, td)
    assert np.max(np.abs(td - ti)) < 2e-3


def test_thermal_analytical_profile():
    v0 = orc.lib().orc_thermal_analytical(0.0, 0.1, 1.0, 1.0)
    v1 = orc.lib().orc_thermal_analytical(1.0, 0.1, 1.0, 1.0)
    assert abs(v0 - 1.0) < 2e-3 and 0 <= v1 < 0.1
    xs = np.linspace(0, 1, 50)
    vals = [orc.lib().orc_thermal_analytical(float(x), 0.1, 1.0, 1.0) for x in xs]
    assert np.all(np.diff(vals) < 0)


MESH_FACTS = {"900_ele": (800, 800, 0), "untitled8192": (8192, 2048, 6144), "test_sn2": (12, 4, 8),
              "split0": (6, None, None), "split1": (24, None, None), "split2": (96, None, None)}


@pytest.mark.parametrize("name", list(MESH_FACTS))
def test_read_msh_facts_and_neighbours(name, tmp_path):
    m = orc.read_msh(write_msh(name, str(tmp_path / (name + ".msh"))))
    U, ccw, cw = MESH_FACTS[name]
    X = m["X"]
    assert X.shape[0] == U
    Xg, reg = mesh_triangles(name)
    assert np.array_equal(X, Xg) and np.array_equal(m["region"], reg)
    det = (X[:, 0, 0] - X[:, 2, 0]) * (X[:, 1, 1] - X[:, 2, 1]) - (X[:, 0, 1] - X[:, 2, 1]) * (X[:, 1, 0] - X[:, 2, 0])
    if ccw is not None:
        assert (det > 0).sum() == ccw and (det < 0).sum() == cw
    fneig, _ = orc.neig_data(m["neig"], m["dir"])
    side = [(0, 2), (0, 1), (1, 2)]
    for u in range(U):
        for f in range(3):
            q = m["neig"][u, f]
            if q == 0:
                continue
            g = fneig[u, f]
            assert m["neig"][q - 1, g - 1] == u + 1
            mine = {tuple(X[u, i]) for i in side[f]}
            theirs = {tuple(X[q - 1, i]) for i in side[g - 1]}
            assert mine == theirs
            assert m["dir"][u, f] == m["dir"][q - 1, g - 1]
    if name == "900_ele":
        assert np.allclose(0.5 * np.abs(det), 1.125)
        assert sorted(set(reg)) == [9, 10]


def _implicit_dense(m, u, dt, use_dir):
    E = m["X"].shape[0]
    fneig, _ = orc.neig_data(m["neig"], m["dir"])
    A = np.zeros((3 * E, 3 * E)); M = np.zeros((3 * E, 3 * E))
    orc.lib().orc_unstr_implicit_assemble(E, np.ascontiguousarray(m["X"]), np.ascontiguousarray(m["neig"]), fneig,
                                          u[0], u[1], dt, use_dir, A, M)
    return A, M, fneig


@pytest.mark.parametrize("name", ["split1", "test_sn2"])
@pytest.mark.parametrize("use_dir", [0, 1])
def test_unstr_implicit_operator_invariants(name, use_dir, tmp_path):
    """The assembled unstr_implicit operator (transport_tri_unstr.F90:270-364): pure advection keeps constants on
    elements away from the boundary ((A - M/dt) 1 = 0 there) and is conservative (columns of elements without an
    outflow boundary face sum to the mass column sum)."""
    m = orc.read_msh(write_msh(name, str(tmp_path / (name + ".msh"))))
    u, dt = (0.9, 0.3), 1e-2
    A, M, _ = _implicit_dense(m, u, dt, use_dir)
    E = m["X"].shape[0]
    interior = np.repeat(np.all(m["neig"] != 0, axis=1), 3)
    assert interior.any() or name == "test_sn2"
    r = (A - M) @ np.ones(3 * E)
    assert np.max(np.abs(r[interior]), initial=0.0) <= 1e-14
    assert np.max(np.abs(r[~interior])) > 1e-3          # inflow boundary rows do lose the constant (t_bc = 0)
    c = np.ones(3 * E) @ (A - M)
    assert np.max(np.abs(c[interior]), initial=0.0) <= 1e-14
    # mass/dt block: (A/12dt)(1 + delta), row sums = area/(3 dt)
    X = m["X"]
    area = 0.5 * np.abs((X[:, 0, 0] - X[:, 2, 0]) * (X[:, 1, 1] - X[:, 2, 1]) - (X[:, 0, 1] - X[:, 2, 1]) * (X[:, 1, 0] - X[:, 2, 0]))
    assert np.allclose(M @ np.ones(3 * E), np.repeat(area / (3 * dt), 3), rtol=1e-13)


def test_unstr_implicit_solve_is_backward_euler(tmp_path):
    """orc_unstr_implicit == dense solve of (lhs + flux) tnew = (M/dt) told per step, and a second nonlinear pass
    (nits = 2) changes nothing because the scheme is linear (:247-387)."""
    m = orc.read_msh(write_msh("split1", str(tmp_path / "s.msh")))
    u, dt = (0.4, -0.7), 5e-3
    A, M, fneig = _implicit_dense(m, u, dt, 1)
    E = m["X"].shape[0]
    rng = np.random.default_rng(5)
    T0 = rng.random(3 * E)
    ref = T0.copy()
    for _ in range(3):
        ref = np.linalg.solve(A, M @ ref)
    for nits in (1, 2):
        T = T0.copy()
        rc = orc.lib().orc_unstr_implicit(E, np.ascontiguousarray(m["X"]), np.ascontiguousarray(m["neig"]), fneig,
                                          u[0], u[1], dt, 3, nits, 1, T)
        assert rc == 0
        assert np.linalg.norm(T - ref) <= 1e-12 * np.linalg.norm(ref)


def test_unstr_stab_known_answer_and_invariants(tmp_path):
    """Petrov-Galerkin stabilisation (transport_tri_unstr.F90:239-267,278).  Unit triangle x1=(1,0) x2=(0,1) x3=(0,0):
    inv_jac = I, grad phi = (1,0),(0,1),(-1,-1).  T=(1,0,0)=told, u=(2,0): rgi = 2, a_star = (2,0), p_star = 0.125,
    diff_coe = 0.25*4*0.125 = 0.125, stab = 0.125 * area * grad_j.grad_i."""
    X = np.array([[[1.0, 0.0], [0.0, 1.0], [0.0, 0.0]]])
    T = np.array([[1.0, 0.0, 0.0]])
    dc = np.zeros((1, 3)); st = np.zeros((1, 9))
    orc.lib().orc_unstr_stab(1, X, T, T.copy(), 2.0, 0.0, 1e-2, dc, st)
    assert np.allclose(dc, 0.125, rtol=1e-14)
    G = np.array([[1.0, 0.0], [0.0, 1.0], [-1.0, -1.0]])
    assert np.allclose(st.reshape(3, 3), 0.125 * 0.5 * G @ G.T, rtol=1e-14)
    # zero residual (steady, gradient perpendicular to u) -> no diffusion at all
    orc.lib().orc_unstr_stab(1, X, T, T.copy(), 0.0, 3.0, 1e-2, dc, st)
    assert np.all(dc == 0.0) and np.all(st == 0.0)
    # flat field: |grad T|^2 < toler, the 1/toler clamps apply: a_star = 0 -> p_star = 1e11, diff = .25 r^2 1e11 / 1e-11
    Tn = np.array([[1.0, 1.0, 1.0]]); To = np.zeros((1, 3))
    orc.lib().orc_unstr_stab(1, X, Tn, To, 1.0, 1.0, 0.5, dc, st)
    assert np.allclose(dc, 0.25 * 4.0 * 1e11 / 1e-11, rtol=1e-12)
    # random fields on a real mesh: symmetric, positive semi-definite, constants in the kernel
    m = orc.read_msh(write_msh("irregular", str(tmp_path / "i.msh")))
    E = m["X"].shape[0]
    rng = np.random.default_rng(3)
    Tn = rng.random((E, 3)); To = rng.random((E, 3))
    dc = np.zeros((E, 3)); st = np.zeros((E, 9))
    orc.lib().orc_unstr_stab(E, np.ascontiguousarray(m["X"]), Tn, To, 0.9, 0.3, 1e-2, dc, st)
    S = st.reshape(E, 3, 3)
    assert np.all(dc >= 0.0)
    assert np.allclose(S, S.transpose(0, 2, 1), rtol=1e-12, atol=0)
    assert np.max(np.abs(S.sum(axis=2))) <= 1e-12 * np.max(np.abs(S))
    assert np.min(np.linalg.eigvalsh(S)) >= -1e-12 * np.max(np.abs(S))


def test_unstr_implicit_diffusion_blocks(tmp_path):
    """orc_unstr_implicit_assemble_diff: the diffusion part added to the implicit operator is the iterative path's (volume
    term k A grad.grad + face penalty k/dx): symmetric, positive semi-definite, constants in its kernel away from the
    domain boundary, zero for k = 0, linear in k, and - the anchor to the multigrid path - identical to what
    orc_semi_residual applies on the same triangles (0_split.msh + n_split = 1 == 1_split.msh)."""
    m = orc.read_msh(write_msh("split1", str(tmp_path / "s1.msh")))
    fneig, _ = orc.neig_data(m["neig"], m["dir"])
    E = m["X"].shape[0]; N = 3 * E
    dt = 1e-2

    def assemble(u, k):
        A = np.zeros((N, N)); M = np.zeros((N, N))
        orc.lib().orc_unstr_implicit_assemble_diff(E, m["X"], m["neig"], fneig, u[0], u[1], k, dt, 1, A, M)
        return A, M
    A0, M = assemble((0.0, 0.0), 0.0)
    assert np.allclose(A0, M, atol=0)
    A1, _ = assemble((0.0, 0.0), 1.0)
    D = A1 - M                                                     # the diffusion operator alone
    assert np.max(np.abs(D - D.T)) <= 1e-13 * np.max(np.abs(D))
    assert np.linalg.eigvalsh(0.5 * (D + D.T)).min() >= -1e-10 * np.max(np.abs(D))
    interior = np.all(m["neig"] != 0, axis=1)
    rows = (D @ np.ones(N)).reshape(E, 3)
    assert np.max(np.abs(rows[interior])) <= 1e-12 * np.max(np.abs(D))
    assert np.max(np.abs(rows[~interior])) > 1e-3                   # Dirichlet faces keep their own penalty part
    A3, _ = assemble((0.0, 0.0), 3.0)
    assert np.allclose(A3 - M, 3.0 * D, rtol=1e-13, atol=1e-13 * np.max(np.abs(D)))
    # against the semi-structured operator on the same triangles
    u, k = (0.4, -0.7), 0.9
    A, _ = assemble(u, k)
    c = orc.read_msh(write_msh("split0", str(tmp_path / "s0.msh")))
    cf, _ = orc.neig_data(c["neig"], c["dir"])
    s = orc.Semi(orc.intended_params(1, 1, dt=dt, k=k, u=u), c["X"], c["neig"], cf, c["dir"])
    from helpers import child_coordinates
    xc = child_coordinates(orc, c["X"], 1)
    f = lambda X: np.sin(3.0 * X[..., 0]) * np.cos(2.0 * X[..., 1]) + 0.5

    def residual_of(T):
        s.field(orc.TNEW)[:] = T; s.update_overlaps(1); s.residual(1)
        return s.field(orc.RES).copy()
    d = f(xc)
    Ad_semi = (residual_of(np.zeros_like(d)) - residual_of(d)).reshape(-1, 3)
    Ad_unstr = (A @ f(m["X"]).ravel()).reshape(E, 3)
    key = lambda p: (round(float(p[0]), 9), round(float(p[1]), 9))
    by = lambda X, V: {tuple(sorted(key(p) for p in X[e])): {key(p): V[e, i] for i, p in enumerate(X[e])} for e in range(X.shape[0])}
    P, Q = by(xc.reshape(-1, 3, 2), Ad_semi), by(m["X"], Ad_unstr)
    assert set(P) == set(Q)
    worst = max(abs(P[e][p] - Q[e][p]) for e in P for p in P[e])
    assert worst <= 1e-11 * np.max(np.abs(Ad_unstr))

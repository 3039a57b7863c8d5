"""GPU parity tests proper: every kernel of the hot path, called through the C ABI (ctypes), against the
CPU oracle on identical seeded inputs.  Tolerance for floating point: relative L2 <= 1e-12 per Jacobi
sweep and per residual evaluation (BASELINE.json north_star); copies are bit-exact."""
import os

import numpy as np
import pytest

import oracle_api as orc
from helpers import GOLDEN, rel_l2, rng_field, write_msh
from pamg_pkg import pamg

pytestmark = pytest.mark.gpu
TOL = 1e-12


def make_pair(mesh, n, levels, intended, u=(0.0, 0.0), dt=None, halo_rule=None, **kw):
    if intended:
        op = orc.intended_params(n, levels, dt=dt or 1e-3, u=u)
    else:
        op = orc.literal_params(n, levels, dt=dt or 1.25e-5, u=u)
    if halo_rule is not None:
        op.halo_rule = halo_rule
    gkw = {k: kw.pop(k) for k in list(kw) if k in ("keep_tnew_gs",)}
    for k, v in kw.items():
        setattr(op, k, v)
    gp = pamg.default_params(literal_head=not intended, n_split=n, multi_levels=levels, **gkw)
    for f in ("face_terms", "literal_source", "transfer", "residual_sign", "halo_rule", "coarse_bc_zero",
              "theta", "dt", "k", "omega", "u_x", "u_y", "source_coef"):
        setattr(gp, f, getattr(op, f))
    o = orc.Semi(op, mesh.X, mesh.neig, mesh.fneig, mesh.dir)
    g = pamg.SemiImplicitIterative(gp, mesh)
    return o, g


def seed_fields(o, g, level=1, seed=20221, with_rhs=False):
    shape = g.shape(level)
    T = rng_field(shape, seed)
    Told = rng_field(shape, seed + 1)
    o.field(orc.TNONLIN, level)[:] = T
    o.field(orc.TNEW, level)[:] = T
    o.field(orc.TOLD, level)[:] = Told
    g.upload(pamg.TNONLIN, level, T)
    g.copy(level, pamg.TNEW, pamg.TNONLIN)
    g.upload(pamg.TOLD, level, Told)
    if with_rhs:
        R = rng_field(shape, seed + 2)
        o.field(orc.RHS, level)[:] = R
        g.upload(pamg.RHS, level, R)
    return T, Told


@pytest.fixture(scope="module")
def meshes(tmp_path_factory):
    d = tmp_path_factory.mktemp("msh")
    out = {}
    for name in ("test_sn2", "split0", "split1", "900_ele", "untitled8192", "irregular"):
        out[name] = pamg.Mesh.read_msh(write_msh(name, str(d / (name + ".msh"))))
    out["syn"] = pamg.Mesh.synthetic(1, 2)
    return out


@pytest.mark.parametrize("name,n", [("test_sn2", 1), ("test_sn2", 3), ("split0", 5), ("irregular", 2), ("syn", 4)])
@pytest.mark.parametrize("rule", [0, 1])
def test_update_overlaps(meshes, name, n, rule):
    o, g = make_pair(meshes[name], n, 1, True, halo_rule=rule)
    seed_fields(o, g)
    o.update_overlaps(1)
    g.update_overlaps(1)
    ref, got = o.overlap(1), g.overlap(1)
    interior = meshes[name].neig != 0
    assert np.array_equal(got[interior], ref[interior])          # copies: bit exact
    assert np.allclose(got[~interior], ref[~interior], rtol=0, atol=4e-16)   # sin(x+y) Dirichlet data
    assert np.array_equal(g.overlap(1, old=True)[interior], o.overlap(1, old=True)[interior])


@pytest.mark.parametrize("literal_source", [0, 1])
def test_build_rhs(meshes, literal_source):
    o, g = make_pair(meshes["test_sn2"], 3, 1, True, literal_source=literal_source)
    seed_fields(o, g)
    o.build_rhs()
    g.get_RHS()
    assert rel_l2(g.download(pamg.RHS), o.field(orc.RHS)) <= TOL


SWEEP_CASES = [
    # mesh, n_split, intended, velocity
    ("test_sn2", 1, False, (0.0, 0.0)),       # HEAD default (main.F90:46, mesh at transport_tri_semi.F90:99)
    ("test_sn2", 3, True, (0.0, 0.0)),
    ("test_sn2", 3, True, (0.9, 0.3)),
    ("split0", 1, True, (0.9, 0.3)),
    ("split0", 4, True, (0.9, 0.3)),
    ("split0", 6, True, (0.0, 0.0)),
    ("split1", 3, False, (0.1, 0.1)),
    ("900_ele", 2, True, (0.1, 0.1)),          # config c1
    ("900_ele", 1, False, (0.0, 0.0)),
    ("irregular", 3, True, (-0.4, 0.7)),
    ("syn", 5, True, (0.9, 0.3)),
    # levels with 6 <= s <= 8 run the window kernel with a producer warp (k_element_win2), s = 4, 5 k_element_win
    ("test_sn2", 6, True, (0.9, 0.3)),
    ("split0", 7, True, (-0.5, 0.8)),
    ("split0", 6, False, (0.1, 0.1)),
    ("syn", 8, True, (0.9, 0.3)),
    # n_split 9: rows of 1023 children do not fit the 8-tile ring - these levels run k_element_tma
    ("split0", 9, True, (0.9, 0.3)),
]


@pytest.mark.parametrize("name,n,intended,u", SWEEP_CASES)
def test_jacobi_sweep_and_residual(meshes, name, n, intended, u):
    o, g = make_pair(meshes[name], n, 1, intended, u=u)
    seed_fields(o, g)
    for sweep in range(3):                      # parity per sweep
        o.smooth(1, 1, 1)
        g.smoother(1, pamg.JACOBI, 1)
        assert rel_l2(g.download(pamg.TNONLIN), o.field(orc.TNONLIN)) <= TOL, sweep
        assert rel_l2(g.download(pamg.TNEW), o.field(orc.TNEW)) <= TOL
    # residual evaluation on the swept field
    o.field(orc.TNEW)[:] = o.field(orc.TNONLIN)
    g.copy(1, pamg.TNEW, pamg.TNONLIN)
    o.update_overlaps(1); g.update_overlaps(1)
    l2o, lio = o.residual(1)
    l2g, lig = g.get_residual(1)
    assert rel_l2(g.download(pamg.RES), o.field(orc.RES)) <= TOL
    assert abs(l2g - l2o) <= 1e-12 * l2o and abs(lig - lio) <= 1e-12 * lio
    assert abs(g.get_convergence(1) - o.convergence(1)) <= 1e-12 * max(1.0, abs(lio))


@pytest.mark.parametrize("name,n,u", [("test_sn2", 3, (0.9, 0.3)), ("split0", 5, (0.0, 0.0)), ("900_ele", 2, (0.1, 0.1)),
                                      ("syn", 4, (0.9, 0.3)),
                                      # n_split <= 4: both colours in one launch, whole parents in shared memory (k_gs_small)
                                      ("irregular", 1, (0.9, 0.3)), ("untitled8192", 1, (-0.4, 0.7)), ("test_sn2", 4, (0.0, -1.0)),
                                      # one-pass kernel with a producer warp (k_gs_win2): 6 <= n_split <= 8
                                      ("test_sn2", 6, (0.9, 0.3)), ("split0", 7, (-0.5, 0.8)), ("syn", 8, (0.9, 0.3)),
                                      ("irregular", 6, (0.3, -0.2)),
                                      # n_split 9: two in-place passes through k_element_tma
                                      ("split0", 9, (0.9, 0.3))])
def test_two_colour_gauss_seidel_sweep(meshes, name, n, u):
    o, g = make_pair(meshes[name], n, 1, True, u=u, keep_tnew_gs=1)
    seed_fields(o, g)
    for sweep in range(3):
        o.smooth(1, 4, 1)                       # oracle runs the same colouring: down children, then up
        g.smoother(1, pamg.GAUSS_SEIDEL, 1)
        assert rel_l2(g.download(pamg.TNONLIN), o.field(orc.TNONLIN)) <= TOL, sweep
        assert rel_l2(g.download(pamg.TNEW), o.field(orc.TNEW)) <= TOL   # start-of-sweep copy (:550)


def test_literal_head_gs_equals_reference_order(meshes):
    """At HEAD the face block is commented out, so the lexicographic sweep of the reference and the
    two-colour GPU sweep must agree to rounding."""
    o, g = make_pair(meshes["test_sn2"], 2, 1, False)
    seed_fields(o, g)
    o.smooth(1, 3, 4)
    g.smoother(1, pamg.GAUSS_SEIDEL, 4)
    assert rel_l2(g.download(pamg.TNONLIN), o.field(orc.TNONLIN)) <= TOL


@pytest.mark.parametrize("n", [3, 4, 6])
def test_richardson_sweep(meshes, n):
    o, g = make_pair(meshes["split0"], n, 1, True, u=(0.9, 0.3), omega=1e-6)
    seed_fields(o, g)
    o.smooth(1, 2, 2)
    g.smoother(1, pamg.RICHARDSON, 2)
    assert rel_l2(g.download(pamg.TNONLIN), o.field(orc.TNONLIN)) <= TOL


@pytest.mark.parametrize("transfer", [0, 1])
@pytest.mark.parametrize("name,n", [("test_sn2", 3), ("split0", 5)])
def test_restrict_and_prolong(meshes, name, n, transfer):
    o, g = make_pair(meshes[name], n, n, True, transfer=transfer)
    for lvl in range(1, n):
        shape_f, shape_c = g.shape(lvl), g.shape(lvl + 1)
        R = rng_field(shape_f, 7 + lvl)
        o.field(orc.RES, lvl)[:] = R
        g.upload(pamg.RES, lvl, R)
        o.restrict(lvl); g.restrictor(lvl)
        assert rel_l2(g.download(pamg.RHS, lvl + 1), o.field(orc.RHS, lvl + 1)) <= TOL
        Tf, Tc = rng_field(shape_f, 70 + lvl), rng_field(shape_c, 700 + lvl)
        for fld_o, fld_g in ((orc.TNONLIN, pamg.TNONLIN), (orc.TNEW, pamg.TNEW)):
            o.field(fld_o, lvl)[:] = Tf; o.field(fld_o, lvl + 1)[:] = Tc
        g.upload(pamg.TNONLIN, lvl, Tf); g.copy(lvl, pamg.TNEW, pamg.TNONLIN)
        g.upload(pamg.TNONLIN, lvl + 1, Tc); g.copy(lvl + 1, pamg.TNEW, pamg.TNONLIN)
        o.prolong(lvl); g.prolongator(lvl)
        fld = (orc.TNEW, pamg.TNEW) if transfer == 0 else (orc.TNONLIN, pamg.TNONLIN)
        assert rel_l2(g.download(fld[1], lvl), o.field(fld[0], lvl)) <= TOL


@pytest.mark.parametrize("name,n,solver,u", [("split0", 4, 1, (0.0, 0.0)), ("split0", 5, 3, (0.0, 0.0)),
                                             ("test_sn2", 4, 3, (0.3, 0.1)), ("900_ele", 2, 1, (0.1, 0.1)),
                                             ("split0", 7, 1, (0.3, 0.1)), ("test_sn2", 6, 3, (0.0, 0.0))])
def test_vcycle_to_1e8_matches_oracle(meshes, name, n, solver, u):
    o, g = make_pair(meshes[name], n, n, True, u=u)
    it_o, hist_o = o.vcycle_solve(solver=4 if solver == 3 else 1, nu1=4, nu2=4, ncoarse=15, max_cycles=40, tol=1e-8)
    it_g, hist_g = g.vcycle_solve(solver=solver, nu1=4, nu2=4, ncoarse=15, max_cycles=40, tol=1e-8)
    assert it_g <= 40 and abs(it_g - it_o) <= 1
    assert hist_g[-1] / hist_g[0] <= 1e-8
    m = min(len(hist_o), len(hist_g))
    assert np.allclose(hist_g[:m], hist_o[:m], rtol=1e-6)
    if it_g == it_o:
        sol_o, sol_g = o.field(orc.TNONLIN), g.download(pamg.TNONLIN)
        assert np.max(np.abs(sol_g - sol_o)) <= 1e-10 * max(1.0, np.max(np.abs(sol_o)))


def test_literal_head_timestep_mode9(meshes):
    """mode = 9 exactly as checked in: test_sn2.msh, n_split 1, multi_levels 1, GS, n_smooth 4, n_multigrid 2,
    IC T = 1 where region_id == 4 (main.F90:46-47, transport_tri_semi.F90:99,118,249-251)."""
    mesh = meshes["test_sn2"]
    o, g = make_pair(mesh, 1, 1, False)
    ic = np.zeros(g.shape(1))
    ic[mesh.region == 4] = 1.0
    o.field(orc.TNEW)[:] = ic
    g.upload(pamg.TNEW, 1, ic)
    for step in range(2):                       # ntime = 2 (:135)
        o.literal_timestep(solver=3, n_multigrid=2, n_smooth=4)
        g.literal_timestep(solver=3, n_multigrid=2, n_smooth=4)
        assert rel_l2(g.download(pamg.TNEW), o.field(orc.TNEW)) <= TOL
        assert rel_l2(g.download(pamg.TNONLIN), o.field(orc.TNONLIN)) <= TOL
        # the field has converged: r = A x - b is pure cancellation, so compare against the size of b
        scale = np.max(np.abs(o.field(orc.RHS)))
        assert np.max(np.abs(g.download(pamg.RES) - o.field(orc.RES))) <= 1e-13 * scale


def test_literal_multilevel_timestep(meshes):
    o, g = make_pair(meshes["split0"], 3, 3, False, u=(0.1, 0.1))
    ic = rng_field(g.shape(1), 5)
    o.field(orc.TNEW)[:] = ic
    g.upload(pamg.TNEW, 1, ic)
    o.literal_timestep(solver=1, n_multigrid=2, n_smooth=4)
    g.literal_timestep(solver=1, n_multigrid=2, n_smooth=4)
    for lvl in (1, 2, 3):
        assert rel_l2(g.download(pamg.TNEW, lvl), o.field(orc.TNEW, lvl)) <= 1e-11, lvl


@pytest.mark.parametrize("exact,use_dir", [(0, 0), (1, 0), (0, 1)])
def test_unstr_explicit_config2(meshes, exact, use_dir):
    """unstr_explicit with the literal arguments of main.F90:28 on untitled8192.msh (IC tag 12)."""
    mesh = meshes["untitled8192"]
    g = pamg.SemiImplicitIterative(pamg.default_params(), pamg.Mesh.synthetic(0, 1))
    g.set_unstructured(mesh)
    T0 = np.zeros((mesh.U, 3))
    T0[mesh.region == 12] = 1.0
    dt = 0.07 * 1e-3
    ref = T0.copy()
    orc.lib().orc_unstr_explicit(mesh.U, mesh.X, mesh.neig, mesh.fneig, mesh.dir, 0.9, 0.0, dt, 2, 2, 10, exact,
                                 use_dir, 0.0, ref)
    got = g.unstr_explicit(T0, dt, 0.9, 0.0, ntime=2, nits=2, njac_its=10, exact_minv=bool(exact), use_dir=bool(use_dir))
    assert rel_l2(got, ref) <= TOL
    assert np.max(np.abs(got - ref)) <= 1e-13


def test_unstr_explicit_random_field_long(meshes):
    mesh = meshes["split1"]
    g = pamg.SemiImplicitIterative(pamg.default_params(), pamg.Mesh.synthetic(0, 1))
    g.set_unstructured(mesh)
    T0 = rng_field((mesh.U, 3), 3)
    ref = T0.copy()
    orc.lib().orc_unstr_explicit(mesh.U, mesh.X, mesh.neig, mesh.fneig, mesh.dir, 0.4, -0.7, 1e-3, 5, 2, 10, 0, 1, 0.25, ref)
    got = g.unstr_explicit(T0, 1e-3, 0.4, -0.7, ntime=5, nits=2, njac_its=10, t_bc=0.25, use_dir=True)
    assert rel_l2(got, ref) <= TOL


@pytest.mark.parametrize("solver", [pamg.JACOBI, pamg.GAUSS_SEIDEL])
def test_smooth_host_pipelined_equals_blocking_calls(meshes, solver):
    """pamg_smooth_host (upload / sweeps / download on three streams, pipelined across calls) == the blocking sequence
    upload_field, smoother, download_field - bit for bit, for several different fields in a row."""
    mesh = meshes["irregular"]
    p = pamg.default_params(n_split=5, multi_levels=1, u_x=0.9, u_y=0.3)
    g = pamg.SemiImplicitIterative(p, mesh)
    shape = g.shape(1)
    n = int(np.prod(shape))
    g.upload(pamg.TOLD, 1, rng_field(shape, 1))
    fields = [rng_field(shape, 10 + i) for i in range(4)]
    want = []
    for f in fields:
        g.upload(pamg.TNONLIN, 1, f); g.copy(1, pamg.TNEW, pamg.TNONLIN)
        g.smoother(1, solver, 3)
        want.append(g.download(pamg.TNONLIN, 1).copy())
    ins = [pamg.PinnedBuffer(n) for _ in fields]
    outs = [pamg.PinnedBuffer(n) for _ in fields]
    for b, f in zip(ins, fields):
        b.array[:] = f.ravel()
    for b, o in zip(ins, outs):
        g.smooth_host(solver, 3, b.ptr, o.ptr)          # no synchronisation in between
    g.sync()
    for o, w in zip(outs, want):
        assert np.array_equal(o.array.reshape(shape), w)
    # the handle is in a consistent state afterwards: the iterate is the last result
    assert np.array_equal(g.download(pamg.TNONLIN, 1), want[-1])


def _bsr_to_dense(val, col):
    E = val.shape[0]
    A = np.zeros((3 * E, 3 * E))
    for e in range(E):
        for b in range(4):
            c = col[e, b]
            if c >= 0:
                A[3 * e:3 * e + 3, 3 * c:3 * c + 3] += val[e, b]
    return A


@pytest.mark.parametrize("name", ["split1", "test_sn2", "irregular", "900_ele"])
@pytest.mark.parametrize("use_dir", [0, 1])
def test_unstr_implicit_bsr_assembly(meshes, name, use_dir):
    """Device block-CSR of unstr_implicit's lhs + flux (transport_tri_unstr.F90:270-364) == the oracle's dense matrix."""
    mesh = meshes[name]
    E = mesh.U
    g = pamg.SemiImplicitIterative(pamg.default_params(), pamg.Mesh.synthetic(0, 1))
    g.set_unstructured(mesh)
    u, dt = (0.9, 0.3), 1e-2
    g.implicit_assemble(dt, u[0], u[1], use_dir=bool(use_dir))
    val, col = g.implicit_bsr()
    assert np.array_equal(col[:, 0], np.arange(E))
    A = np.zeros((3 * E, 3 * E)); M = np.zeros((3 * E, 3 * E))
    orc.lib().orc_unstr_implicit_assemble(E, mesh.X, mesh.neig, mesh.fneig, u[0], u[1], dt, use_dir, A, M)
    got = _bsr_to_dense(val, col)
    assert np.max(np.abs(got - A)) <= 1e-13 * np.max(np.abs(A))
    # inflow blocks exist exactly where the dense matrix couples to a neighbour
    for e in range(E):
        for f in range(3):
            q = mesh.neig[e, f]
            if q:
                assert (col[e, 1 + f] == q - 1) == bool(np.any(A[3 * e:3 * e + 3, 3 * (q - 1):3 * q]))
            else:
                assert col[e, 1 + f] == -1
    x = rng_field((E, 3), 11)
    y = g.implicit_apply(x)
    assert rel_l2(y.ravel(), A @ x.ravel()) <= TOL


@pytest.mark.parametrize("name,use_dir,u", [("split1", 1, (0.4, -0.7)), ("irregular", 0, (0.9, 0.3)),
                                            ("900_ele", 1, (0.9, 0.0)), ("test_sn2", 0, (-0.5, 0.8))])
def test_unstr_implicit_time_loop(meshes, name, use_dir, u):
    """Krylov solve on the device block-CSR vs the reference's dense FINDInv solve (:366-378), 3 time steps."""
    mesh = meshes[name]
    E = mesh.U
    g = pamg.SemiImplicitIterative(pamg.default_params(), pamg.Mesh.synthetic(0, 1))
    g.set_unstructured(mesh)
    area = 0.5 * np.abs((mesh.X[:, 0, 0] - mesh.X[:, 2, 0]) * (mesh.X[:, 1, 1] - mesh.X[:, 2, 1])
                        - (mesh.X[:, 0, 1] - mesh.X[:, 2, 1]) * (mesh.X[:, 1, 0] - mesh.X[:, 2, 0]))
    dt = 2.0 * float(np.sqrt(area.min()))          # CFL ~ 2: out of reach of the explicit step
    T0 = rng_field((E, 3), 7)
    ref = T0.copy()
    if E <= 100:
        assert orc.lib().orc_unstr_implicit(E, mesh.X, mesh.neig, mesh.fneig, u[0], u[1], dt, 3, 1, use_dir, ref) == 0
    else:   # the oracle's O(N^3) FINDInv is too slow here; LAPACK on the oracle's own matrices (pinned equal on CPU)
        A = np.zeros((3 * E, 3 * E)); M = np.zeros((3 * E, 3 * E))
        orc.lib().orc_unstr_implicit_assemble(E, mesh.X, mesh.neig, mesh.fneig, u[0], u[1], dt, use_dir, A, M)
        r = T0.ravel().copy()
        for _ in range(3):
            r = np.linalg.solve(A, M @ r)
        ref = r.reshape(E, 3)
    got, iters, relres = g.unstr_implicit(T0, dt, u[0], u[1], ntime=3, nits=1, use_dir=bool(use_dir), tol=1e-13)
    assert relres <= 1e-13 and 0 < iters < 3 * 500
    assert rel_l2(got, ref) <= 1e-10
    # nits = 2 re-solves the same linear system: same answer, no extra Krylov iterations
    got2, iters2, _ = g.unstr_implicit(T0, dt, u[0], u[1], ntime=3, nits=2, use_dir=bool(use_dir), tol=1e-13)
    assert rel_l2(got2, ref) <= 1e-10 and iters2 == iters


@pytest.mark.parametrize("name", ["split1", "irregular", "900_ele", "untitled8192"])
def test_unstr_stabilisation_arrays(meshes, name):
    """diff_coe / stab of transport_tri_unstr.F90:239-267,278 on the device == oracle (HEAD computes them and stops)."""
    mesh = meshes[name]
    E = mesh.U
    g = pamg.SemiImplicitIterative(pamg.default_params(), pamg.Mesh.synthetic(0, 1))
    g.set_unstructured(mesh)
    Tn = rng_field((E, 3), 21); To = rng_field((E, 3), 22)
    dt = 1e-2
    dc, st = g.unstr_stab(Tn, To, dt, 0.9, 0.3)
    rdc = np.zeros((E, 3)); rst = np.zeros((E, 9))
    orc.lib().orc_unstr_stab(E, mesh.X, Tn, To, 0.9, 0.3, dt, rdc, rst)
    assert np.allclose(dc, rdc, rtol=1e-10, atol=1e-12 * np.max(np.abs(rdc)))
    assert np.allclose(st.reshape(E, 9), rst, rtol=1e-10, atol=1e-12 * np.max(np.abs(rst)))
    # zero residual -> exactly no diffusion
    dc0, st0 = g.unstr_stab(np.ones((E, 3)), np.ones((E, 3)), dt, 0.9, 0.3)
    assert np.all(dc0 == 0.0) and np.all(st0 == 0.0)


@pytest.mark.parametrize("name,use_dir", [("split1", 1), ("test_sn2", 0)])
def test_unstr_implicit_with_stabilisation(meshes, name, use_dir):
    """INTENDED use of the stabilisation: diagonal blocks += stab(tnew_nonlin, told) in every nonlinear pass."""
    mesh = meshes[name]
    E = mesh.U
    g = pamg.SemiImplicitIterative(pamg.default_params(), pamg.Mesh.synthetic(0, 1))
    g.set_unstructured(mesh)
    T0 = rng_field((E, 3), 9)
    dt, u = 5e-2, (0.4, -0.7)
    ref = T0.copy()
    assert orc.lib().orc_unstr_implicit_stab(E, mesh.X, mesh.neig, mesh.fneig, u[0], u[1], dt, 2, 2, use_dir, ref) == 0
    got, iters, relres = g.unstr_implicit(T0, dt, u[0], u[1], ntime=2, nits=2, use_dir=bool(use_dir), tol=1e-13, with_stab=True)
    assert relres <= 1e-13 and iters > 0
    assert rel_l2(got, ref) <= 1e-9
    plain, _, _ = g.unstr_implicit(T0, dt, u[0], u[1], ntime=2, nits=2, use_dir=bool(use_dir), tol=1e-13)
    assert rel_l2(plain, ref) > 1e-6            # the stabilisation really changed the answer, and switching it off works
    ref2 = T0.copy()
    assert orc.lib().orc_unstr_implicit(E, mesh.X, mesh.neig, mesh.fneig, u[0], u[1], dt, 2, 2, use_dir, ref2) == 0
    assert rel_l2(plain, ref2) <= 1e-10


def test_unstr_implicit_errors(meshes):
    g = pamg.SemiImplicitIterative(pamg.default_params(), pamg.Mesh.synthetic(0, 1))
    L = pamg.lib()
    assert L.pamg_implicit_assemble(g.h, 1e-2, 1.0, 0.0, 0) == pamg.ERR_ARG          # no unstructured mesh yet
    g.set_unstructured(meshes["split0"])
    assert L.pamg_implicit_step(g.h, 1, 1, 1e-12, 10, None, None) == pamg.ERR_STATE  # not assembled
    assert L.pamg_implicit_assemble(g.h, 0.0, 1.0, 0.0, 0) == pamg.ERR_ARG


def test_str_explicit_front_end():
    """str_explicit (transport_tri.F90:354) = structured triangles (pamg_mesh_structured_tri) + the same explicit DG
    step; t_bc = 0 and u_bc = u as at :100,120, initial box pulse, 3 time steps of 2 nonlinear passes."""
    ner, nec = 40, 10
    mesh = pamg.Mesh.structured_tri(ner, nec, 0.05, 0.1)
    g = pamg.SemiImplicitIterative(pamg.default_params(), pamg.Mesh.synthetic(0, 1))
    g.set_unstructured(mesh)
    T0 = np.zeros((mesh.U, 3))
    cx = mesh.X[:, :, 0].mean(axis=1)
    T0[(cx > 0.2) & (cx < 0.5)] = 1.0
    dt = 0.05 * 0.05
    ref = T0.copy()
    orc.lib().orc_unstr_explicit(mesh.U, mesh.X, mesh.neig, mesh.fneig, mesh.dir, 1.0, 0.0, dt, 3, 2, 10, 0, 0, 0.0, ref)
    got = g.unstr_explicit(T0, dt, 1.0, 0.0, ntime=3, nits=2, njac_its=10)
    assert rel_l2(got, ref) <= TOL
    assert abs(got.sum() - T0.sum()) <= 1e-10 * T0.sum()      # the pulse has not reached the outflow side: mass is conserved


@pytest.mark.parametrize("direct,volume,ner,nec,u", [(0, 0, 200, 1, (2 * 0.01428571, 0.0)), (0, 1, 200, 1, (2 * 0.01428571, 0.0)),
                                                        (1, 1, 200, 1, (2 * 0.01428571, 0.0)), (0, 1, 40, 6, (0.02, 0.013)),
                                                        (1, 0, 30, 4, (-0.02, 0.01))])
def test_trans_rec_front_end(direct, volume, ner, nec, u):
    """trans_rec (transport_rect.F90:7) on the device == oracle; the first case is main.F90:19, whose oracle result is
    pinned against the reference's shipped output file on the CPU side."""
    g = pamg.SemiImplicitIterative(pamg.default_params(), pamg.Mesh.synthetic(0, 1))
    time = 250.0 if nec == 1 else 40.0
    x, t, nt = g.trans_rec(0.7, ner, nec, u[0], u[1], time, nits=2, njac_its=10, direct_solver=bool(direct),
                           volume_term=bool(volume))
    rx = np.zeros((ner * nec, 4, 2)); rt = np.zeros((ner * nec, 4))
    rnt = orc.lib().orc_trans_rec(0.7, ner, nec, 100.0, 100.0, u[0], u[1], time, 2, 10, direct, volume, rx, rt)
    assert nt == rnt and np.array_equal(x, rx)
    assert rel_l2(t, rt) <= 1e-11
    if (direct, volume, ner, nec) == (0, 0, 200, 1):
        num = np.load(os.path.join(GOLDEN, "rect_golden.npz"))["numerical"]
        assert np.max(np.abs(t.ravel() - num[:, 2])) <= 3e-5      # the device result against the reference's own file


@pytest.mark.parametrize("n", [3, 4, 6])
def test_batched_local_inverse(n):
    g = pamg.SemiImplicitIterative(pamg.default_params(), pamg.Mesh.synthetic(0, 1))
    rng = np.random.default_rng(n)
    B = 1000
    M = rng.random((B, n, n)) + n * np.eye(n)
    M[5] = 0.0                                   # singular block -> errorflag -1, inverse = 0
    if n == 3:
        A = 0.37
        M[6] = A / 12 * (np.ones((3, 3)) + np.eye(3))   # P1 mass matrix: inverse (12/A)(I - J/4)
    M[7, 0, 0] = 0.0                             # zero leading pivot: repaired by adding row 2
    rhs = rng.random((B, n))
    Minv, x, status = g.findinv(M, rhs)
    for b in range(B):
        ref = np.zeros(n * n)
        flag = orc.lib().orc_findinv(np.ascontiguousarray(M[b].ravel()), ref, n)
        assert status[b] == flag
        assert np.allclose(Minv[b].ravel(), ref, rtol=1e-12, atol=1e-13)
        if flag == 0:
            assert np.allclose(x[b], ref.reshape(n, n) @ rhs[b], rtol=1e-12, atol=1e-13)
    assert status[5] == -1 and np.all(Minv[5] == 0)
    if n == 3:
        assert np.allclose(Minv[6], 12 / 0.37 * (np.eye(3) - 0.25 * np.ones((3, 3))), rtol=1e-12)


def test_errors_are_reported_not_swallowed(meshes):
    p = pamg.default_params(n_split=2, multi_levels=3)
    with pytest.raises(pamg.PamgError):
        pamg.SemiImplicitIterative(p, meshes["test_sn2"])       # multi_levels > n_split (:120-123)
    g = pamg.SemiImplicitIterative(pamg.default_params(n_split=2, multi_levels=1), meshes["test_sn2"])
    with pytest.raises(pamg.PamgError):
        g.smoother(2, 1, 1)                                      # no such level
    with pytest.raises(pamg.PamgError):
        g.smoother(1, 9, 1)                                      # unknown solver (select case default)


@pytest.mark.parametrize("name", ["split1", "test_sn2", "irregular", "900_ele"])
@pytest.mark.parametrize("use_dir", [0, 1])
def test_unstr_implicit_bsr_assembly_with_diffusion(meshes, name, use_dir):
    """Block-CSR with the diffusion operator of the iterative path (volume term + face penalty, get_A_x
    transport_tri_semi.F90:412-448 / matrices.F90:84-115) == the oracle's dense matrix; product and solve follow."""
    mesh = meshes[name]
    E = mesh.U
    g = pamg.SemiImplicitIterative(pamg.default_params(), pamg.Mesh.synthetic(0, 1))
    g.set_unstructured(mesh)
    u, dt, k = (0.9, 0.3), 1e-2, 0.7
    g.implicit_assemble(dt, u[0], u[1], use_dir=bool(use_dir), k=k)
    val, col = g.implicit_bsr()
    A = np.zeros((3 * E, 3 * E)); M = np.zeros((3 * E, 3 * E))
    orc.lib().orc_unstr_implicit_assemble_diff(E, mesh.X, mesh.neig, mesh.fneig, u[0], u[1], k, dt, use_dir, A, M)
    got = _bsr_to_dense(val, col)
    assert np.max(np.abs(got - A)) <= 1e-13 * np.max(np.abs(A))
    assert np.array_equal(col[:, 1:] >= 0, mesh.neig != 0)          # with diffusion every interior face couples
    x = rng_field((E, 3), 11)
    assert rel_l2(g.implicit_apply(x).ravel(), A @ x.ravel()) <= TOL
    # one backward-Euler step: Krylov solve on the device (scalars of the recurrence device-resident) vs LAPACK
    syncs0 = g.implicit_host_syncs()
    g._ck(g.L.pamg_unstr_upload(g.h, x))
    it, rr = np.zeros(1, np.int32), np.zeros(1)
    import ctypes as C
    g._ck(g.L.pamg_implicit_step(g.h, 1, 1, 1e-13, 800, it.ctypes.data_as(C.POINTER(C.c_int)), rr.ctypes.data_as(C.POINTER(C.c_double))))
    sol = np.empty_like(x)
    g._ck(g.L.pamg_unstr_download(g.h, sol))
    ref = np.linalg.solve(A, M @ x.ravel())
    assert rr[0] <= 1e-13 and rel_l2(sol.ravel(), ref) <= 1e-10
    assert g.implicit_host_syncs() - syncs0 <= 1 + it[0] // 16       # one look at the flag per 16 iterations
    # without advection the operator is symmetric
    g.implicit_assemble(dt, 0.0, 0.0, use_dir=True, k=k)
    S = _bsr_to_dense(*g.implicit_bsr())
    assert np.max(np.abs(S - S.T)) <= 1e-13 * np.max(np.abs(S))


@pytest.mark.parametrize("n", [1, 2, 3, 4])
def test_semi_structured_operator_equals_unstructured_on_the_refined_mesh(tmp_path, n):
    """`n_split.msh` fully unstructured == `0_split.msh` + n_split = n semi-structured (the same triangles
    in another order).  The operator of the multigrid path (mass/dt - advection + upwind flux + diffusion + penalty, applied
    through get_residual) must equal the block-CSR operator assembled on the gmsh-refined mesh, element by element (matched
    by their vertex sets), for the same nodal function on both meshes."""
    coarse = pamg.Mesh.read_msh(write_msh("split0", str(tmp_path / "split0.msh")))
    fine = pamg.Mesh.read_msh(write_msh("split%d" % n, str(tmp_path / "fine.msh")))
    assert fine.U == coarse.U * 4 ** n
    u, dt, k = (0.4, -0.7), 2e-2, 0.9
    gp = pamg.default_params(n_split=n, multi_levels=1, dt=dt, k=k, u_x=u[0], u_y=u[1])
    g = pamg.SemiImplicitIterative(gp, coarse)
    xc, _, _ = g.output_fields()                                  # child coordinates (U, C, 3, 2)

    def field(X):      # a smooth nodal function: the same function on both meshes
        return np.sin(3.0 * X[..., 0]) * np.cos(2.0 * X[..., 1]) + 0.5

    def residual_of(T):
        g.upload(pamg.TNEW, 1, T); g.copy(1, pamg.TNONLIN, pamg.TNEW); g.update_overlaps(1)
        g.get_residual(1)
        return g.download(pamg.RES, 1)
    g.fill(pamg.TOLD, 1, 0.0)
    d = field(xc)
    # r = b - A T (+ boundary data): the difference of two residuals is -A d, right-hand side and Dirichlet data cancel
    Ad_semi = residual_of(np.zeros_like(d)) - residual_of(d)
    g2 = pamg.SemiImplicitIterative(pamg.default_params(), pamg.Mesh.synthetic(0, 1))
    g2.set_unstructured(fine)
    g2.implicit_assemble(dt, u[0], u[1], use_dir=True, k=k)
    Ad_unstr = g2.implicit_apply(field(fine.X))
    key = lambda p: (round(float(p[0]), 9), round(float(p[1]), 9))

    def by_element(X, V):
        out = {}
        for e in range(X.shape[0]):
            out[tuple(sorted(key(p) for p in X[e]))] = {key(p): V[e, i] for i, p in enumerate(X[e])}
        return out
    A = by_element(xc.reshape(-1, 3, 2), Ad_semi.reshape(-1, 3))
    B = by_element(fine.X, Ad_unstr)
    assert set(A) == set(B) and len(A) == fine.U
    scale = float(np.max(np.abs(Ad_unstr)))
    worst = max(abs(A[e][p] - B[e][p]) for e in A for p in A[e])
    assert worst <= 1e-11 * scale, (worst, scale)

"""Properties the intended solver components must have, checked on the oracle alone (CPU): more pins for the triangle
multigrid path, which has no reference output to compare with.

  * the level-1 operator assembled column by column from residual evaluations is the matrix whose direct solution is the
    fixed point of every smoother and the limit of the V-cycle (a11, a12, a15);
  * restriction is the transpose of the P1 prolongation (`transfer = 1`, a13)."""
import numpy as np
import pytest

import oracle_api as orc
from helpers import rng_field, write_msh


def problem(name, n, levels, tmp_path, u=(0.6, -0.3), dt=1e-3, k=1.0):
    # (dt as in the benchmarks: with dt = 2e-2 the damped Jacobi smoother of the reference, omega = 0.8 on the lumped-mass
    # diagonal, diverges on test_sn2 - a property of the algorithm, not of an implementation)
    m = orc.read_msh(write_msh(name, str(tmp_path / (name + ".msh"))))
    fneig, _ = orc.neig_data(m["neig"], m["dir"])
    p = orc.intended_params(n, levels, dt=dt, k=k, u=u)
    return orc.Semi(p, m["X"], m["neig"], fneig, m["dir"])


def operator_and_rhs(o):
    """A and b of level 1 from residual evaluations: r(x) = b - A x (residual_sign = -1), Dirichlet data folded into b."""
    shape = o.field(orc.TNEW).shape
    N = int(np.prod(shape))

    def resid(x):
        o.field(orc.TNEW)[:] = x.reshape(shape); o.field(orc.TNONLIN)[:] = x.reshape(shape)
        o.update_overlaps(1)
        o.residual(1)
        return o.field(orc.RES).reshape(-1).copy()
    b = resid(np.zeros(N))
    A = np.zeros((N, N))
    for j in range(N):
        e = np.zeros(N); e[j] = 1.0
        A[:, j] = b - resid(e)
    return A, b


@pytest.mark.parametrize("name,n", [("test_sn2", 2), ("split0", 2)])
def test_direct_solution_is_the_fixed_point_of_the_smoothers_and_the_limit_of_the_vcycle(name, n, tmp_path):
    o = problem(name, n, n, tmp_path)
    shape = o.field(orc.TNEW).shape
    o.field(orc.TOLD)[:] = rng_field(shape, 5)
    o.build_rhs()
    A, b = operator_and_rhs(o)
    assert np.linalg.cond(A) < 1e8
    xs = np.linalg.solve(A, b)
    scale = np.abs(xs).max()
    for solver in (1, 4):                                   # Jacobi, two-colour Gauss-Seidel
        o.field(orc.TNONLIN)[:] = xs.reshape(shape); o.field(orc.TNEW)[:] = xs.reshape(shape)
        o.smooth(1, solver, 3)
        assert np.abs(o.field(orc.TNONLIN).reshape(-1) - xs).max() <= 1e-11 * scale, solver
    for solver in (1, 4):
        o.field(orc.TNONLIN)[:] = 0.0; o.field(orc.TNEW)[:] = 0.0
        cycles, hist = o.vcycle_solve(solver=solver, max_cycles=60, tol=1e-10)
        assert cycles <= 60 and hist[-1] <= 1e-10 * hist[0], (cycles, hist[-1] / hist[0])
        assert np.abs(o.field(orc.TNONLIN).reshape(-1) - xs).max() <= 1e-7 * scale, solver
    # one Jacobi sweep is x + omega D^-1 (b - A x) with ONE fixed positive diagonal D (the reference's get_diagonal: lumped
    # mass / dt + K_ii + penalty diagonal, transport_tri_semi.F90:481-486 - not diag(A), whose mass part is consistent)
    Ds = []
    for seed in (9, 10):
        x0 = rng_field(shape, seed).reshape(-1)
        o.field(orc.TNONLIN)[:] = x0.reshape(shape); o.field(orc.TNEW)[:] = x0.reshape(shape)
        o.smooth(1, 1, 1)
        dx = o.field(orc.TNONLIN).reshape(-1) - x0
        Ds.append(o.params.omega * (b - A @ x0) / dx)
    assert np.all(Ds[0] > 0.0)
    assert np.abs(Ds[0] - Ds[1]).max() <= 1e-8 * np.abs(Ds[0]).max()
    assert np.all(Ds[0] >= np.diag(A) * (1.0 - 1e-12)) or np.all(Ds[0] > 0.5 * np.diag(A))   # lumped >= consistent mass on the diagonal


@pytest.mark.parametrize("name,n", [("test_sn2", 3), ("irregular", 2)])
def test_restriction_is_the_transpose_of_the_p1_prolongation(name, n, tmp_path):
    o = problem(name, n, 2, tmp_path)
    f_shape, c_shape = o.field(orc.TNEW, 1).shape, o.field(orc.TNEW, 2).shape
    r = rng_field(f_shape, 21) - 0.5
    e = rng_field(c_shape, 22) - 0.5
    o.field(orc.RES, 1)[:] = r
    o.restrict(1)                                           # RHS(2) = R r
    Rr = o.field(orc.RHS, 2).copy()
    o.field(orc.TNONLIN, 1)[:] = 0.0; o.field(orc.TNEW, 1)[:] = 0.0
    o.field(orc.TNONLIN, 2)[:] = e; o.field(orc.TNEW, 2)[:] = e
    o.prolong(1)                                            # T(1) += P e
    Pe = o.field(orc.TNONLIN, 1).copy()
    lhs, rhs = float(np.sum(Rr * e)), float(np.sum(r * Pe))
    assert abs(lhs - rhs) <= 1e-12 * max(abs(lhs), abs(rhs), 1.0), (lhs, rhs)


@pytest.mark.parametrize("name,use_dir,u", [("split1", 1, (0.4, -0.7)), ("irregular", 1, (0.9, 0.3)), ("900_ele", 0, (1.0, 0.0)),
                                            ("untitled8192", 1, (-0.6, 0.2))])
def test_explicit_step_advects_a_linear_field_exactly(name, use_dir, u, tmp_path):
    """unstr_explicit (transport_tri_unstr.F90:588-795) with the exact local mass inverse: for a continuous linear field the
    upwind DG right-hand side is int phi_i u.grad(T) on every element without a boundary face, so one forward-Euler step gives
    T - dt u.g at every node of those elements - whatever the mesh, neighbour numbering or node pairing."""
    m = orc.read_msh(write_msh(name, str(tmp_path / (name + ".msh"))))
    fneig, _ = orc.neig_data(m["neig"], m["dir"])
    X = m["X"]
    g = np.array([1.3, -0.8])
    T0 = 0.2 + X @ g                                         # (E, 3)
    dt = 1e-4
    T = np.ascontiguousarray(T0.copy())
    orc.lib().orc_unstr_explicit(X.shape[0], np.ascontiguousarray(X), np.ascontiguousarray(m["neig"]), fneig,
                                 np.ascontiguousarray(m["dir"]), u[0], u[1], dt, 1, 1, 10, 1, use_dir, 0.0, T)
    interior = np.all(m["neig"] != 0, axis=1)
    assert interior.sum() >= 4
    expect = T0 - dt * float(np.dot(g, u))
    err = np.abs(T - expect)[interior]
    assert err.max() <= 1e-12 * np.abs(expect).max(), err.max()


@pytest.mark.gpu
@pytest.mark.parametrize("name,use_dir,u", [("irregular", 1, (0.9, 0.3)), ("untitled8192", 1, (-0.6, 0.2))])
def test_device_explicit_step_advects_a_linear_field_exactly(name, use_dir, u, tmp_path):
    """the same identity on k_unstr_explicit (division-free geometry) through the C ABI"""
    from pamg_pkg import pamg
    mesh = pamg.Mesh.read_msh(write_msh(name, str(tmp_path / (name + ".msh"))))
    gs = pamg.SemiImplicitIterative(pamg.default_params(), pamg.Mesh.synthetic(0, 1))
    gs.set_unstructured(mesh)
    g = np.array([1.3, -0.8])
    T0 = 0.2 + mesh.X @ g
    dt = 1e-4
    T = gs.unstr_explicit(T0, dt, u[0], u[1], ntime=1, nits=1, njac_its=10, exact_minv=True, use_dir=bool(use_dir))
    interior = np.all(mesh.neig != 0, axis=1)
    expect = T0 - dt * float(np.dot(g, u))
    assert np.abs(T - expect)[interior].max() <= 1e-12 * np.abs(expect).max()
    gs.close()


@pytest.mark.parametrize("name,u,k", [("split1", (0.4, -0.7), 0.9), ("irregular", (0.9, 0.3), 0.0), ("900_ele", (1.0, 0.0), 0.5)])
def test_implicit_operator_on_a_continuous_linear_field(name, u, k, tmp_path):
    """unstr_implicit's matrix (transport_tri_unstr.F90:270-364, + the diffusion blocks of the iterative path): applied to a
    continuous linear field it must give mass/dt T + int phi_i u.grad(T) + k A grad(phi_i).g on every element without a
    boundary face (no jump: the penalty vanishes, the upwind flux telescopes)."""
    m = orc.read_msh(write_msh(name, str(tmp_path / (name + ".msh"))))
    fneig, _ = orc.neig_data(m["neig"], m["dir"])
    X = m["X"]; E = X.shape[0]; N = 3 * E
    dt = 1e-2
    A = np.zeros((N, N)); M = np.zeros((N, N))
    orc.lib().orc_unstr_implicit_assemble_diff(E, np.ascontiguousarray(X), np.ascontiguousarray(m["neig"]), fneig, u[0], u[1], k,
                                               dt, 1, A, M)
    g = np.array([1.3, -0.8])
    T = 0.2 + X @ g                                          # (E, 3)
    AT = (A @ T.reshape(-1)).reshape(E, 3)
    e1, e2 = X[:, 0] - X[:, 2], X[:, 1] - X[:, 2]
    det = e1[:, 0] * e2[:, 1] - e1[:, 1] * e2[:, 0]
    area = 0.5 * np.abs(det)
    gphi = np.zeros((E, 3, 2))
    gphi[:, 0, 0] = e2[:, 1] / det; gphi[:, 0, 1] = -e2[:, 0] / det
    gphi[:, 1, 0] = -e1[:, 1] / det; gphi[:, 1, 1] = e1[:, 0] / det
    gphi[:, 2] = -(gphi[:, 0] + gphi[:, 1])
    expect = (area[:, None] / 12.0) * (T + T.sum(axis=1, keepdims=True)) / dt
    expect += (np.dot(g, u) * area / 3.0)[:, None] + k * area[:, None] * (gphi @ g)
    interior = np.all(m["neig"] != 0, axis=1)
    assert interior.sum() >= 4
    assert np.abs(AT - expect)[interior].max() <= 1e-11 * np.abs(expect[interior]).max()

"""theta != 1 (get_A_x / get_RHS, transport_tri_semi.F90:441-446,457-460) on the oracle alone (CPU).

The reference carries `theta` through get_A_x and get_RHS but only ever sets it to 1 (`:117`), so no run of the reference
exercises the (1 - theta) branches.  The oracle implements them in the Crank-Nicolson reading (every old-time spatial term on
TOLD with the told strips).  Pins, none of which depends on how the oracle is written:

  * an algebraic identity: with R_th(x, told) the residual of the theta-scheme,
        R_th(x, told) = R_1(x, told) - (1 - theta) [ R_1(x, x) - R_1(told, told) ]
    (R_1(x, told) - R_1(x, x) = M (x - told) / dt and R_1(told, told) is the spatial operator on told), i.e. the old-time
    branch of get_RHS equals what the theta = 1 path - pinned elsewhere - computes for the same field;
  * the time stepper converges to the exact solution exp(-M^-1 A t) of the semi-discrete system with first order for
    theta = 1 and SECOND order for theta = 1/2;
  * the direct solution of the theta-system is the fixed point of the Jacobi and the two-colour Gauss-Seidel sweep and the
    limit of the V-cycle (so the diagonal and the coarse operators carry theta consistently)."""
import numpy as np
import pytest
import scipy.linalg as sl

import oracle_api as orc
from helpers import rng_field, write_msh


def problem(name, n, levels, theta, dt, tmp_path, k=0.05, u=(0.6, -0.3)):
    m = orc.read_msh(write_msh(name, str(tmp_path / (name + ".msh"))))
    fneig, _ = orc.neig_data(m["neig"], m["dir"])
    p = orc.intended_params(n, levels, dt=dt, k=k, u=u)
    p.theta = theta
    return orc.Semi(p, m["X"], m["neig"], fneig, m["dir"])


def resid(o, x, told):
    """b(told) - A x on level 1 (residual_sign = -1), halo strips and RHS refreshed."""
    sh = o.field(orc.TNEW).shape
    o.field(orc.TNEW)[:] = x.reshape(sh); o.field(orc.TNONLIN)[:] = x.reshape(sh); o.field(orc.TOLD)[:] = told.reshape(sh)
    o.update_overlaps(1)
    o.residual(1)
    return o.field(orc.RES).reshape(-1).copy()


def matrix(o):
    N = int(np.prod(o.field(orc.TNEW).shape))
    z = np.zeros(N)
    b = resid(o, z, z)
    A = np.zeros((N, N))
    for j in range(N):
        e = np.zeros(N); e[j] = 1.0
        A[:, j] = b - resid(o, e, z)
    return A, b


@pytest.mark.parametrize("name,n", [("test_sn2", 2), ("irregular", 2), ("split0", 3)])
@pytest.mark.parametrize("theta", [0.5, 0.25, 0.0])
def test_old_time_branch_equals_the_theta_one_operator_on_told(name, n, theta, tmp_path):
    o1 = problem(name, n, 1, 1.0, 1e-2, tmp_path)
    ot = problem(name, n, 1, theta, 1e-2, tmp_path)
    sh = o1.field(orc.TNEW).shape
    x = rng_field(sh, 3).reshape(-1); told = rng_field(sh, 4).reshape(-1)
    lhs = resid(ot, x, told)
    rhs = resid(o1, x, told) - (1.0 - theta) * (resid(o1, x, x) - resid(o1, told, told))
    assert np.abs(lhs - rhs).max() <= 1e-13 * np.abs(lhs).max()


def test_time_stepper_converges_to_the_semi_discrete_solution_with_the_order_of_theta(tmp_path):
    name, n, dt0, tend = "test_sn2", 2, 0.02, 0.08
    A0, _ = matrix(problem(name, n, 1, 0.0, dt0, tmp_path))      # theta = 0: A = M / dt exactly
    A1, g = matrix(problem(name, n, 1, 1.0, dt0, tmp_path))      # M / dt + Abar;  g = M src + Dirichlet data (told = 0)
    M, Abar = A0 * dt0, A1 - A0
    N = M.shape[0]
    Tinf = np.linalg.solve(Abar, g)
    T0 = np.zeros(N)
    exact = Tinf + sl.expm(-np.linalg.solve(M, Abar) * tend) @ (T0 - Tinf)
    order = {}
    for theta in (1.0, 0.5):
        errs = []
        for steps in (16, 32, 64):
            o = problem(name, n, 1, theta, tend / steps, tmp_path)
            A, _ = matrix(o)
            lu = sl.lu_factor(A)
            x = T0.copy()
            for _ in range(steps):
                x = sl.lu_solve(lu, resid(o, np.zeros(N), x))    # b(told = x)
            errs.append(np.abs(x - exact).max())
        order[theta] = np.log2(errs[1] / errs[2])
        assert errs[2] < errs[1] < errs[0]
    assert abs(order[1.0] - 1.0) <= 0.05, order     # backward Euler
    assert abs(order[0.5] - 2.0) <= 0.05, order     # Crank-Nicolson


@pytest.mark.parametrize("theta", [0.5, 0.75])
def test_theta_system_is_the_fixed_point_of_the_smoothers_and_the_limit_of_the_vcycle(theta, tmp_path):
    n = 2
    o = problem("test_sn2", n, n, theta, 1e-3, tmp_path, k=1.0)
    shape = o.field(orc.TNEW).shape
    told = rng_field(shape, 5).reshape(-1)
    o1 = problem("test_sn2", n, 1, theta, 1e-3, tmp_path, k=1.0)
    A, _ = matrix(o1)
    b = resid(o1, np.zeros(A.shape[0]), told)
    xs = np.linalg.solve(A, b)
    scale = np.abs(xs).max()
    o.field(orc.TOLD)[:] = told.reshape(shape)
    for solver in (1, 4):
        o.field(orc.TNONLIN)[:] = xs.reshape(shape); o.field(orc.TNEW)[:] = xs.reshape(shape)
        o.smooth(1, solver, 3)
        assert np.abs(o.field(orc.TNONLIN).reshape(-1) - xs).max() <= 1e-11 * scale, solver
    for solver in (1, 4):
        o.field(orc.TNONLIN)[:] = 0.0; o.field(orc.TNEW)[:] = 0.0
        cycles, hist = o.vcycle_solve(solver=solver, max_cycles=60, tol=1e-10)
        assert cycles <= 60 and hist[-1] <= 1e-10 * hist[0], (cycles, hist[-1] / hist[0])
        assert np.abs(o.field(orc.TNONLIN).reshape(-1) - xs).max() <= 1e-7 * scale, solver
    # the diagonal of the sweep carries theta: D = ml/dt + theta (K_ii + penalty diagonal)
    x0 = rng_field(shape, 9).reshape(-1)
    o.field(orc.TNONLIN)[:] = x0.reshape(shape); o.field(orc.TNEW)[:] = x0.reshape(shape)
    o.smooth(1, 1, 1)
    D = o.params.omega * (b - A @ x0) / (o.field(orc.TNONLIN).reshape(-1) - x0)
    ob = problem("test_sn2", n, 1, 1.0, 1e-3, tmp_path, k=1.0)
    ob.field(orc.TOLD)[:] = told.reshape(shape)
    A1, _ = matrix(ob)
    b1 = resid(ob, np.zeros(A.shape[0]), told)
    ob.field(orc.TNONLIN)[:] = x0.reshape(shape); ob.field(orc.TNEW)[:] = x0.reshape(shape)
    ob.smooth(1, 1, 1)
    D1 = ob.params.omega * (b1 - A1 @ x0) / (ob.field(orc.TNONLIN).reshape(-1) - x0)
    A0, _ = matrix(problem("test_sn2", n, 1, 0.0, 1e-3, tmp_path, k=1.0))
    Dm = A0 @ np.ones(A.shape[0])                      # row sums of M / dt = lumped mass / dt
    assert np.abs(D - (Dm + theta * (D1 - Dm))).max() <= 1e-8 * np.abs(D1).max()

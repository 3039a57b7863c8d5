"""mode 9 exactly as checked in (main.F90:46-47: one level, Gauss-Seidel, 4 sweeps per smoother call, 2 passes of the
`do multigrid` loop) replayed in numpy, CPU only.

At HEAD the face loop body is commented out (transport_tri_semi.F90:619-688), so the operator is block diagonal and the whole
time step is a fixed sequence of per-child 3x3 iterations that can be written down from the Fortran without any of the
oracle's machinery:

    b      = M told / dt + src'        src = -2k sin(x+y), src' = M src summed IN PLACE (get_RHS :455-456: row i already
                                       sees the overwritten entries 1..i-1)
    sweep  : tnew <- tnew_nonlin (:550);  tnew_nonlin <- tnew_nonlin + omega / D (b - A tnew_nonlin)
             A = M/dt - S + K,  D_i = ml_i/dt + K_ii (get_diagonal :481-486, lumped mass)
    pass   : tnew_nonlin <- tnew (:327), smoother, get_residual, tnew_nonlin <- tnew (:348: drops the last sweep), 15 smoother calls
    step   : told <- tnew, tnew_nonlin <- tnew (:316-317), n_multigrid passes

The oracle's literal time step (and with it the device's, which is compared with the oracle on the GPU) must give the same field."""
import numpy as np
import pytest

import oracle_api as orc
from helpers import child_coordinates, write_msh


def blocks(tri, u, k, dt):
    """per element: A = M/dt - S + K, M, ml from exact P1 integrals"""
    E = tri.shape[0]
    A = np.zeros((E, 3, 3)); M = np.zeros((E, 3, 3))
    for e in range(E):
        x = tri[e]
        d = np.array([[x[1, 1] - x[2, 1], x[2, 0] - x[1, 0]], [x[2, 1] - x[0, 1], x[0, 0] - x[2, 0]], [x[0, 1] - x[1, 1], x[1, 0] - x[0, 0]]])
        det = (x[1, 0] - x[0, 0]) * (x[2, 1] - x[0, 1]) - (x[2, 0] - x[0, 0]) * (x[1, 1] - x[0, 1])
        g = d / det
        area = 0.5 * abs(det)
        M[e] = area / 12.0 * (np.ones((3, 3)) + np.eye(3))
        A[e] = M[e] / dt + k * area * g @ g.T - np.outer(g @ np.asarray(u, float), np.ones(3)) * area / 3.0
    return A, M, M.sum(axis=2)


@pytest.mark.parametrize("name,n,u,region", [("test_sn2", 1, (0.0, 0.0), 4), ("test_sn2", 2, (0.0, 0.0), 4),
                                             ("split1", 1, (0.1, 0.1), None), ("900_ele", 1, (0.0, 0.0), 9)])
def test_literal_head_time_steps_replayed_from_the_fortran(name, n, u, region, tmp_path):
    n_smooth, n_multigrid, ntime = 4, 2, 2
    m = orc.read_msh(write_msh(name, str(tmp_path / (name + ".msh"))))
    fneig, _ = orc.neig_data(m["neig"], m["dir"])
    p = orc.literal_params(n, 1, u=u)
    o = orc.Semi(p, m["X"], m["neig"], fneig, m["dir"])
    xy = child_coordinates(orc, m["X"], n)
    U, C = xy.shape[0], xy.shape[1]
    tri = xy.reshape(U * C, 3, 2)
    A, M, ml = blocks(tri, u, p.k, p.dt)
    # get_diagonal has no advection entry: K_ii from the blocks without velocity
    A0, _, _ = blocks(tri, (0.0, 0.0), p.k, p.dt)
    D = ml / p.dt + np.stack([A0[:, i, i] - M[:, i, i] / p.dt for i in range(3)], axis=1)
    if region is None:
        T = np.random.Generator(np.random.MT19937(6)).random((U * C, 3))
    else:
        T = np.zeros((U * C, 3))
        T[np.repeat(m["region"] == region, C)] = 1.0             # IC :249-251
    o.field(orc.TNEW)[:] = T.reshape(U, C, 3)
    src = p.source_coef * np.sin(tri[:, :, 0] + tri[:, :, 1])     # :593
    s = src.copy()
    for i in range(3):                                            # in place, row by row (:455-456)
        s[:, i] = np.einsum("ej,ej->e", M[:, i, :], s)
    srcp = s

    def sweep(tn):
        return tn + p.omega / D * (b - np.einsum("eij,ej->ei", A, tn))

    tnew = T.copy()
    for _ in range(ntime):
        told = tnew.copy(); tnl = tnew.copy()
        b = np.einsum("eij,ej->ei", M, told) / p.dt + srcp
        for _ in range(n_multigrid):
            tnl = tnew.copy()                                     # :327
            for _ in range(n_smooth):
                tnew = tnl.copy(); tnl = sweep(tnl)               # smoother :331
            tnl = tnew.copy()                                     # :348
            for _ in range(15 * n_smooth):
                tnew = tnl.copy(); tnl = sweep(tnl)               # :351-352
        o.literal_timestep(solver=3, n_multigrid=n_multigrid, n_smooth=n_smooth)
        got = o.field(orc.TNEW).reshape(U * C, 3)
        assert np.abs(got - tnew).max() <= 1e-11 * max(1.0, np.abs(tnew).max())
        assert np.abs(o.field(orc.TNONLIN).reshape(U * C, 3) - tnl).max() <= 1e-11 * max(1.0, np.abs(tnl).max())

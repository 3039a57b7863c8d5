"""An analytic pin of the oracle's intended operator (the triangle multigrid path has no reference output to pin
against, so the restatement is checked against what the discretisation MUST give).

For a globally continuous, piecewise-linear field T = a + b x + c y the DG operator of the iterative path collapses to closed
forms on every child that does not touch the domain boundary, whatever the mesh:
  * the face penalty (k/dx) int sn_i (T - T2) vanishes (no jump),
  * the upwind flux sees the same value on both sides, so  - int grad(phi_i).u T + sum_faces int phi_i (n.u) T^  equals
    int phi_i u.grad(T) = (u.g) A/3  by the divergence theorem,
  * the volume diffusion is k A grad(phi_i).g  and the mass term (A/12)(T_i + sum_j T_j)/dt.
Anything wrong in the shape-function tables, the per-child geometry scaling, the neighbour tables, the node pairing across
child faces, the halo strips between parents (both reversal rules) or a sign convention breaks the identity."""
import numpy as np
import pytest

import oracle_api as orc
from helpers import write_msh


def child_coordinates(X, n):
    U, Cn = X.shape[0], 4 ** n
    xc = np.zeros((U, Cn, 3, 2))
    x6 = np.zeros((3, 2))
    L = orc.lib()
    for un in range(U):
        P = np.ascontiguousarray(X[un])
        for ele in range(1, Cn + 1):
            L.orc_get_splitting(P, n, ele, x6)
            xc[un, ele - 1] = x6
    return xc


@pytest.mark.parametrize("halo_rule", [0, 1])
@pytest.mark.parametrize("name,n,u", [("test_sn2", 3, (0.7, -0.4)), ("split1", 2, (-0.3, 0.9)), ("irregular", 3, (0.5, 0.5)),
                                      ("900_ele", 1, (1.0, 0.0)), ("split0", 4, (0.0, 0.0))])
def test_intended_operator_on_a_continuous_linear_field(name, n, u, halo_rule, tmp_path):
    m = orc.read_msh(write_msh(name, str(tmp_path / (name + ".msh"))))
    fneig, _ = orc.neig_data(m["neig"], m["dir"])
    dt, k = 0.05, 0.6
    p = orc.intended_params(n, 1, dt=dt, k=k, u=u, source_coef=0.0)
    p.halo_rule = halo_rule
    o = orc.Semi(p, m["X"], m["neig"], fneig, m["dir"])
    U, Cn, S = m["X"].shape[0], 4 ** n, 2 ** n
    xc = child_coordinates(m["X"], n)
    g = np.array([1.7, -0.9])
    T = 0.3 + xc @ g                                                   # (U, C, 3)
    o.field(orc.TNEW)[:] = T; o.field(orc.TNONLIN)[:] = T; o.field(orc.TOLD)[:] = 0.0
    o.build_rhs()                                                      # (M/dt) told + M src = 0
    assert np.all(o.field(orc.RHS) == 0.0)
    o.update_overlaps(1)
    o.residual(1)                                                      # r = b - A T (residual_sign = -1)
    AT = -o.field(orc.RES).copy()
    # closed form per child
    e1, e2 = xc[:, :, 0] - xc[:, :, 2], xc[:, :, 1] - xc[:, :, 2]
    det = e1[..., 0] * e2[..., 1] - e1[..., 1] * e2[..., 0]
    area = 0.5 * np.abs(det)
    gphi = np.zeros((U, Cn, 3, 2))                                     # gradients of the three hat functions
    gphi[:, :, 0, 0] = e2[..., 1] / det; gphi[:, :, 0, 1] = -e2[..., 0] / det
    gphi[:, :, 1, 0] = -e1[..., 1] / det; gphi[:, :, 1, 1] = e1[..., 0] / det
    gphi[:, :, 2] = -(gphi[:, :, 0] + gphi[:, :, 1])
    sumT = T.sum(axis=2, keepdims=True)
    expect = (area[..., None] / 12.0) * (T + sumT) / dt
    expect += (np.dot(g, u) * area / 3.0)[..., None]
    expect += k * area[..., None] * (gphi @ g)
    # children on a parent face that lies on the domain boundary see Dirichlet data sin(x+y): not part of the identity
    se = np.zeros(3 * S, np.int32)
    orc.lib().orc_surf_ele(n, se)
    se = se.reshape(3, S)
    interior = np.ones((U, Cn), bool)
    for un in range(U):
        for mf in range(3):
            if m["neig"][un, mf] == 0:
                interior[un, se[mf] - 1] = False
    assert interior.sum() > 0.3 * U * Cn
    err = np.abs(AT - expect)[interior]
    scale = np.abs(expect[interior]).max()
    assert err.max() <= 1e-10 * scale, (err.max(), scale)


@pytest.mark.gpu
@pytest.mark.parametrize("name,n,u", [("test_sn2", 3, (0.7, -0.4)), ("irregular", 3, (0.5, 0.5)), ("split0", 6, (-0.3, 0.9))])
def test_device_operator_on_a_continuous_linear_field(name, n, u, tmp_path):
    """The same identity straight on the device operator (window kernels at n_split 6, direct kernels below), through the
    C ABI: get_residual of a continuous linear field against the closed form."""
    from pamg_pkg import pamg
    mesh = pamg.Mesh.read_msh(write_msh(name, str(tmp_path / (name + ".msh"))))
    dt, k = 0.05, 0.6
    gp = pamg.default_params(n_split=n, multi_levels=1, dt=dt, k=k, u_x=u[0], u_y=u[1], source_coef=0.0)
    gsolver = pamg.SemiImplicitIterative(gp, mesh)
    U, Cn, S = mesh.U, 4 ** n, 2 ** n
    xc = child_coordinates(mesh.X, n)
    g = np.array([1.7, -0.9])
    T = 0.3 + xc @ g
    gsolver.upload(pamg.TNEW, 1, T); gsolver.copy(1, pamg.TNONLIN, pamg.TNEW); gsolver.fill(pamg.TOLD, 1, 0.0)
    gsolver.update_overlaps(1)
    gsolver.get_residual(1)
    AT = -gsolver.download(pamg.RES, 1).reshape(U, Cn, 3)
    e1, e2 = xc[:, :, 0] - xc[:, :, 2], xc[:, :, 1] - xc[:, :, 2]
    det = e1[..., 0] * e2[..., 1] - e1[..., 1] * e2[..., 0]
    area = 0.5 * np.abs(det)
    gphi = np.zeros((U, Cn, 3, 2))
    gphi[:, :, 0, 0] = e2[..., 1] / det; gphi[:, :, 0, 1] = -e2[..., 0] / det
    gphi[:, :, 1, 0] = -e1[..., 1] / det; gphi[:, :, 1, 1] = e1[..., 0] / det
    gphi[:, :, 2] = -(gphi[:, :, 0] + gphi[:, :, 1])
    expect = (area[..., None] / 12.0) * (T + T.sum(axis=2, keepdims=True)) / dt
    expect += (np.dot(g, u) * area / 3.0)[..., None] + k * area[..., None] * (gphi @ g)
    se = np.zeros(3 * S, np.int32)
    orc.lib().orc_surf_ele(n, se)
    se = se.reshape(3, S)
    interior = np.ones((U, Cn), bool)
    for un in range(U):
        for mf in range(3):
            if mesh.neig[un, mf] == 0:
                interior[un, se[mf] - 1] = False
    err = np.abs(AT - expect)[interior]
    scale = np.abs(expect[interior]).max()
    assert err.max() <= 1e-10 * scale, (err.max(), scale)
    gsolver.close()

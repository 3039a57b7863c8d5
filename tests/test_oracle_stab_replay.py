"""The Petrov-Galerkin stabilisation arrays of unstr_implicit (transport_tri_unstr.F90:239-267,278: rgi, a_star, p_star,
diff_coe, stab) restated in numpy from the formulas, CPU only:

    r_g      = (T_g - Told_g)/dt + u . grad T            at the three edge-midpoint Gauss points (ShapFun.F90:554-563)
    a*       = r_g grad T / max(tol, |grad T|^2)
    p*       = min(1/tol, 0.25 / max_d |(J^-1 a*)_d|)    J = [[x1-x3, y1-y3], [x2-x3, y2-y3]]   (INV_JAC is stored transposed,
                                                         ShapFun.F90:1440-1450, and contracted over its first index at :262)
    diff_g   = 0.25 r_g^2 p* / max(tol, |grad T|^2)
    stab_ij  = sum_g diff_g grad phi_j . grad phi_i |K|/3"""
import numpy as np
import pytest

import oracle_api as orc
from helpers import write_msh

GAUSS = np.array([[0.5, 0.5, 0.0], [0.0, 0.5, 0.5], [0.5, 0.0, 0.5]])      # barycentric coordinates of gi = 1, 2, 3


@pytest.mark.parametrize("name", ["gmsh_100", "irregular", "900_ele"])
def test_stabilisation_arrays_follow_the_formulas(name, tmp_path):
    m = orc.read_msh(write_msh(name, str(tmp_path / (name + ".msh"))))
    X = np.ascontiguousarray(m["X"])
    E = X.shape[0]
    u, dt, tol = np.array([-0.1, 0.1]), 0.07, 1e-11
    rng = np.random.Generator(np.random.MT19937(31))
    tn, to = rng.random((E, 3)), rng.random((E, 3))
    tn[::7] = 0.3 + 1e-3 * rng.random((len(tn[::7]), 3))   # nearly flat elements: large a*, small p* (an exactly flat element
    #                                                        sits on the 1/tol clamps, where rounding noise in grad T decides:
    #                                                        those are pinned by hand in tests/test_oracle_known_answers.py)
    dc = np.zeros((E, 3)); st = np.zeros((E, 3, 3))
    orc.lib().orc_unstr_stab(E, X, tn, to, u[0], u[1], dt, dc, st)
    for e in range(E):
        x = X[e]
        J = np.array([x[0] - x[2], x[1] - x[2]])
        det = np.linalg.det(J)
        grad = np.linalg.solve(J, np.array([[1.0, 0.0, -1.0], [0.0, 1.0, -1.0]])).T       # rows: grad phi_i
        area = 0.5 * abs(det)
        gT = tn[e] @ grad
        g2 = max(tol, gT @ gT)
        want_dc = np.zeros(3)
        for g in range(3):
            r = (GAUSS[g] @ tn[e] - GAUSS[g] @ to[e]) / dt + u @ gT
            a = r / g2 * gT
            with np.errstate(divide="ignore"):
                ps = min(1.0 / tol, 0.25 / np.abs(np.linalg.solve(J, a)).max())
            want_dc[g] = 0.25 * r * r * ps / g2
        want_st = sum(want_dc[g] * (grad @ grad.T) * area / 3.0 for g in range(3))
        assert np.allclose(dc[e], want_dc, rtol=1e-9, atol=1e-12 * max(1.0, np.abs(want_dc).max())), e
        assert np.allclose(st[e], want_st, rtol=1e-9, atol=1e-12 * max(1.0, np.abs(want_st).max())), e

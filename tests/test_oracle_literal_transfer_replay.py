"""restrictor / prolongator AS WRITTEN (splitting.F90:10-91) replayed geometrically in numpy, CPU only.

The Fortran addresses the four fine children of a coarse child through element_conversion's closed-form ids and fixed local
node numbers.  Read geometrically it says (fin1, fin3, fin4 are the corner children at coarse nodes 3, 1, 2; fin2 is the
inverted centre child):

  restrictor : coarse RHS at node j = mean of the three residuals of the corner child that contains coarse node j;
  prolongator: the children are visited in the order corner@3, centre, corner@1, corner@2.  A fine node on a coarse vertex
               receives the coarse value; the edge midpoints of corner@3 and the midpoint of edge (1,2) in the centre child
               receive the P1 interpolant; every other midpoint receives the already updated TOTAL of the coincident node of
               the child visited before (corner@3 for the centre child, the centre child for corner@1 and corner@2) - the
               mixing of totals and corrections of SURVEY B-8.

The replay below finds children, nodes and coincidences from coordinates only, so it checks the oracle's use of
element_conversion and of the local node numbering, which the device kernels (k_restrict mode 0, k_prolong_literal) mirror."""
import numpy as np
import pytest

import oracle_api as orc
from helpers import child_coordinates, write_msh


def key(p):
    return (round(float(p[0]) * 1e9), round(float(p[1]) * 1e9))


def families(xy_c, xy_f):
    """for every coarse child of a parent: its fine children as (corner@1, corner@2, corner@3, centre) - found by containment"""
    Cc, Cf = xy_c.shape[0], xy_f.shape[0]
    cen = xy_f.mean(axis=1)
    out = []
    for c in range(Cc):
        x = xy_c[c]
        T = np.array([x[0] - x[2], x[1] - x[2]]).T
        lam = np.linalg.solve(T, (cen - x[2]).T).T
        inside = np.flatnonzero((lam[:, 0] > -1e-9) & (lam[:, 1] > -1e-9) & (lam.sum(axis=1) < 1 + 1e-9))
        assert len(inside) == 4
        corner = [None, None, None]; centre = None
        for f in inside:
            hit = [j for j in range(3) if any(key(xy_f[f, i]) == key(x[j]) for i in range(3))]
            if hit:
                corner[hit[0]] = int(f)
            else:
                centre = int(f)
        assert None not in corner and centre is not None
        out.append((corner[0], corner[1], corner[2], centre))
    return out


@pytest.mark.parametrize("name,n", [("test_sn2", 2), ("irregular", 3), ("split0", 3)])
def test_literal_restrictor_and_prolongator_follow_the_geometry(name, n, tmp_path):
    m = orc.read_msh(write_msh(name, str(tmp_path / (name + ".msh"))))
    fneig, _ = orc.neig_data(m["neig"], m["dir"])
    p = orc.literal_params(n, 2)
    o = orc.Semi(p, m["X"], m["neig"], fneig, m["dir"])
    xy_f, xy_c = child_coordinates(orc, m["X"], n), child_coordinates(orc, m["X"], n - 1)
    U, Cf, Cc = xy_f.shape[0], xy_f.shape[1], xy_c.shape[1]
    rng = np.random.Generator(np.random.MT19937(17))
    res = rng.random((U, Cf, 3)); fine = rng.random((U, Cf, 3)); coarse = rng.random((U, Cc, 3))
    # ---- restrictor
    o.field(orc.RES, 1)[:] = res
    o.restrict(1)
    got = o.field(orc.RHS, 2).copy()
    want = np.zeros((U, Cc, 3))
    fam = [families(xy_c[u], xy_f[u]) for u in range(U)]
    for u in range(U):
        for c, (k1, k2, k3, _) in enumerate(fam[u]):
            want[u, c] = [res[u, k1].mean(), res[u, k2].mean(), res[u, k3].mean()]
    assert np.abs(got - want).max() <= 1e-14
    # ---- prolongator
    o.field(orc.TNEW, 1)[:] = fine; o.field(orc.TNONLIN, 1)[:] = fine
    o.field(orc.TNEW, 2)[:] = coarse; o.field(orc.TNONLIN, 2)[:] = coarse
    o.prolong(1)
    got = o.field(orc.TNEW, 1).copy()
    want = fine.copy()
    for u in range(U):
        for c, (k1, k2, k3, kc) in enumerate(fam[u]):
            X, cv = xy_c[u, c], coarse[u, c]
            vertex = {key(X[j]): cv[j] for j in range(3)}
            interp = {key(0.5 * (X[a] + X[b])): 0.5 * (cv[a] + cv[b]) for a, b in ((0, 1), (1, 2), (0, 2))}
            mid12 = key(0.5 * (X[0] + X[1]))
            total = {}                                           # coordinate -> total written by the child visited before
            for kid, source in ((k3, "interp"), (kc, "corner3"), (k1, "centre"), (k2, "centre")):
                new_total = {}
                for i in range(3):
                    kk = key(xy_f[u, kid, i])
                    if kk in vertex:
                        want[u, kid, i] += vertex[kk]
                    elif source == "interp" or (source == "corner3" and kk == mid12):
                        want[u, kid, i] += interp[kk]
                    else:
                        want[u, kid, i] += total[kk]
                    new_total[kk] = want[u, kid, i]
                if source in ("interp", "corner3"):
                    total = new_total                             # corner@1 and corner@2 both read the centre child's totals
    assert np.abs(got - want).max() <= 1e-13


def volume_blocks(tri, u, k, dt):
    E = tri.shape[0]
    A = np.zeros((E, 3, 3)); M = np.zeros((E, 3, 3)); K = np.zeros((E, 3))
    for e in range(E):
        x = tri[e]
        d = np.array([[x[1, 1] - x[2, 1], x[2, 0] - x[1, 0]], [x[2, 1] - x[0, 1], x[0, 0] - x[2, 0]], [x[0, 1] - x[1, 1], x[1, 0] - x[0, 0]]])
        det = (x[1, 0] - x[0, 0]) * (x[2, 1] - x[0, 1]) - (x[2, 0] - x[0, 0]) * (x[1, 1] - x[0, 1])
        g = d / det
        area = 0.5 * abs(det)
        M[e] = area / 12.0 * (np.ones((3, 3)) + np.eye(3))
        Kb = k * area * g @ g.T
        K[e] = np.diag(Kb)
        A[e] = M[e] / dt + Kb - np.outer(g @ np.asarray(u, float), np.ones(3)) * area / 3.0
    return A, M, M.sum(axis=2) / dt + K


@pytest.mark.parametrize("name,n,levels,u", [("test_sn2", 2, 2, (0.0, 0.0)), ("split1", 3, 3, (0.1, 0.1))])
def test_literal_multilevel_time_step_replayed_from_the_fortran(name, n, levels, u, tmp_path):
    """`do multigrid` as checked in (transport_tri_semi.F90:319-379) on several levels, face loop commented out: smoother,
    restrictor (of the residual of the PREVIOUS pass: it runs before get_residual), get_residual with r = A x - b on
    tracer%tnew, 15 smoother calls on the coarsest level, and on the way up tnew_nonlin <- tnew BEFORE the prolongator, so the
    smoother's first `tnew = tnew_nonlin` (:550) overwrites the prolonged field - all of it replayed with per-child blocks and
    the geometric transfers above."""
    n_smooth, n_multigrid = 4, 2
    m = orc.read_msh(write_msh(name, str(tmp_path / (name + ".msh"))))
    fneig, _ = orc.neig_data(m["neig"], m["dir"])
    p = orc.literal_params(n, levels, u=u)
    o = orc.Semi(p, m["X"], m["neig"], fneig, m["dir"])
    U = m["X"].shape[0]
    xy = [child_coordinates(orc, m["X"], n - l) for l in range(levels)]
    blk = [volume_blocks(x.reshape(-1, 3, 2), u, p.k, p.dt) for x in xy]
    fam = [[families(xy[l + 1][q], xy[l][q]) for q in range(U)] for l in range(levels - 1)]
    shape = [x.shape[:2] + (3,) for x in xy]
    rng = np.random.Generator(np.random.MT19937(23))
    T0 = rng.random(shape[0])
    o.field(orc.TNEW, 1)[:] = T0
    tri0 = xy[0].reshape(-1, 3, 2)
    s = p.source_coef * np.sin(tri0[:, :, 0] + tri0[:, :, 1])
    for i in range(3):
        s[:, i] = np.einsum("ej,ej->e", blk[0][1][:, i, :], s)            # in-place M src (:455-456)
    tnew = [np.zeros(sh) for sh in shape]; tnl = [np.zeros(sh) for sh in shape]
    rhs = [np.zeros(sh) for sh in shape]; res = [np.zeros(sh) for sh in shape]
    tnew[0] = T0.copy()

    def smooth(l, count):
        A, _, D = blk[l]
        for _ in range(count):
            tnew[l] = tnl[l].copy()
            x = tnl[l].reshape(-1, 3)
            tnl[l] = (x + p.omega / D * (rhs[l].reshape(-1, 3) - np.einsum("eij,ej->ei", A, x))).reshape(shape[l])

    def restrict(l):
        if l == levels - 1:
            return
        for q in range(U):
            for c, (k1, k2, k3, _) in enumerate(fam[l][q]):
                rhs[l + 1][q, c] = [res[l][q, k1].mean(), res[l][q, k2].mean(), res[l][q, k3].mean()]

    def residual(l):
        A = blk[l][0]
        res[l] = (np.einsum("eij,ej->ei", A, tnew[l].reshape(-1, 3)) - rhs[l].reshape(-1, 3)).reshape(shape[l])   # :869

    def prolong(l):
        for q in range(U):
            for c, (k1, k2, k3, kc) in enumerate(fam[l][q]):
                X, cv = xy[l + 1][q, c], tnew[l + 1][q, c]
                vertex = {key(X[j]): cv[j] for j in range(3)}
                interp = {key(0.5 * (X[a] + X[b])): 0.5 * (cv[a] + cv[b]) for a, b in ((0, 1), (1, 2), (0, 2))}
                mid12 = key(0.5 * (X[0] + X[1]))
                total = {}
                for kid, source in ((k3, "interp"), (kc, "corner3"), (k1, "centre"), (k2, "centre")):
                    new_total = {}
                    for i in range(3):
                        kk = key(xy[l][q, kid, i])
                        if kk in vertex:
                            tnew[l][q, kid, i] += vertex[kk]
                        elif source == "interp" or (source == "corner3" and kk == mid12):
                            tnew[l][q, kid, i] += interp[kk]
                        else:
                            tnew[l][q, kid, i] += total[kk]
                        new_total[kk] = tnew[l][q, kid, i]
                    if source in ("interp", "corner3"):
                        total = new_total

    for step in range(2):
        told = tnew[0].copy(); tnl[0] = tnew[0].copy()                      # :316-317
        rhs[0] = (np.einsum("eij,ej->ei", blk[0][1], told.reshape(-1, 3)) / p.dt + s).reshape(shape[0])
        for _ in range(n_multigrid):
            for l in range(levels):
                tnl[l] = tnew[l].copy()                                     # :327
                smooth(l, n_smooth); restrict(l); residual(l)               # :331,336,338
            tnl[levels - 1] = tnew[levels - 1].copy()                       # :348
            for _ in range(15):
                smooth(levels - 1, n_smooth)                                # :351-352
            for l in range(levels - 2, -1, -1):
                tnl[l] = tnew[l].copy()                                     # :367
                prolong(l)                                                  # :370
                smooth(l, n_smooth)                                         # :376
        o.literal_timestep(solver=3, n_multigrid=n_multigrid, n_smooth=n_smooth)
        for l in range(levels):
            scale = max(1.0, np.abs(tnew[l]).max())
            assert np.abs(o.field(orc.TNEW, l + 1) - tnew[l]).max() <= 1e-10 * scale, (step, l)
            assert np.abs(o.field(orc.TNONLIN, l + 1) - tnl[l]).max() <= 1e-10 * scale, (step, l)
            assert np.abs(o.field(orc.RES, l + 1) - res[l]).max() <= 1e-9 * max(1.0, np.abs(res[l]).max()), (step, l)

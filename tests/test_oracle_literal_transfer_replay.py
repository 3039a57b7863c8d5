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

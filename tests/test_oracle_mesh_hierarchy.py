"""multigrid_meshes pin: the children of (0_split.msh, n_split = s) produced by get_splitting are
exactly the triangles of s_split.msh (gmsh "refine by splitting" of the same unit square)."""
import numpy as np
import pytest

import oracle_api as orc
from helpers import mesh_triangles


def tri_key(x):
    pts = sorted((round(float(p[0]), 9), round(float(p[1]), 9)) for p in x)   # gmsh writes 0.2499999999994 for 1/4
    return tuple(pts)


@pytest.mark.parametrize("s", [1, 2, 3, 4])
def test_children_of_split0_are_the_triangles_of_split_s(s):
    X0, _ = mesh_triangles("split0")
    Xs, _ = mesh_triangles(f"split{s}")
    assert Xs.shape[0] == 6 * 4 ** s
    want = {tri_key(t) for t in Xs}
    assert len(want) == Xs.shape[0]
    got = set()
    x = np.zeros((3, 2))
    for u in range(X0.shape[0]):
        for e in range(1, 4 ** s + 1):
            orc.lib().orc_get_splitting(np.ascontiguousarray(X0[u]), s, e, x)
            got.add(tri_key(x))
    assert got == want

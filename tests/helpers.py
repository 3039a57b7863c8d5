"""Shared test helpers: golden meshes -> .msh text, numpy synthetic parents, norms."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

_meshes = None


def golden_meshes():
    global _meshes
    if _meshes is None:
        _meshes = np.load(os.path.join(GOLDEN, "meshes.npz"))
    return _meshes


def write_msh(name, path):
    """Re-emit a golden mesh as gmsh 2.2 ASCII (same node/element tables as the reference file)."""
    z = golden_meshes()
    nodes, elems = z[name + "__nodes"], z[name + "__elems"]
    with open(path, "w") as f:
        f.write("$MeshFormat\n2.2 0 8\n$EndMeshFormat\n$Nodes\n%d\n" % len(nodes))
        for r in nodes:
            f.write("%d %.17g %.17g %.17g\n" % (int(r[0]), r[1], r[2], r[3]))
        f.write("$EndNodes\n$Elements\n%d\n" % len(elems))
        for e in elems:
            f.write(" ".join(str(int(v)) for v in e if v >= 0) + "\n")
        f.write("$EndElements\n")
    return path


def mesh_triangles(name):
    """(X[U,3,2], region[U]) straight from the golden arrays (no neighbour info)."""
    z = golden_meshes()
    nodes, elems = z[name + "__nodes"], z[name + "__elems"]
    xy = {int(r[0]): (r[1], r[2]) for r in nodes}
    X, reg = [], []
    for e in elems:
        if e[1] in (2, 9, 20, 21, 23, 24, 25):
            nt = e[2]
            ids = e[3 + nt: 6 + nt]
            X.append([xy[int(i)] for i in ids])
            reg.append(e[3])
    return np.array(X, np.float64), np.array(reg, np.int32)


def rel_l2(a, b):
    a = np.asarray(a, np.float64).ravel()
    b = np.asarray(b, np.float64).ravel()
    d = np.linalg.norm(b)
    return np.linalg.norm(a - b) / (d if d > 0 else 1.0)


def rng_field(shape, seed=20221):
    return np.random.Generator(np.random.MT19937(seed)).random(shape)

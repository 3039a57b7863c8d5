#!/usr/bin/env python
"""Generates tests/golden/*.npz from the reference tree (run in the build container only).

/root/reference does not exist on the GPU box, so everything the tests need from it is parsed
here once and committed as small binary fixtures:
  meshes.npz        gmsh-2.2 meshes of Mesh_files/ as arrays (nodes, element table); the tests write
                    them back to .msh text in a tmpdir so that both mesh readers are exercised
  rect_golden.npz   the two dumps written by trans_rec (transport_rect.F90:320-353)
Usage:  python tests/golden/make_golden.py [/root/reference]
"""
import os
import sys

import numpy as np

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))

MESHES = {
    "test_sn2": "Mesh_files/test_sn2.msh",
    "900_ele": "Mesh_files/900_ele.msh",
    "untitled8192": "Mesh_files/untitled8192.msh",
    "untitled2048": "Mesh_files/untitled2048.msh",
    "untitled8": "Mesh_files/untitled8.msh",
    "2_unele_test": "Mesh_files/2_unele_test.msh",
    "irregular": "Mesh_files/irregular.msh",
    "semi_mesh": "Mesh_files/semi_mesh.msh",
    "gmsh_100": "Mesh_files/gmsh_100.msh",
    "split0": "Mesh_files/multigrid_meshes/0_split.msh",
    "split1": "Mesh_files/multigrid_meshes/1_split.msh",
    "split2": "Mesh_files/multigrid_meshes/2_split.msh",
    "split3": "Mesh_files/multigrid_meshes/3_split.msh",
    "split4": "Mesh_files/multigrid_meshes/4_split.msh",
}


def parse_msh(path):
    with open(path) as f:
        lines = [l.strip() for l in f]
    i = lines.index("$Nodes")
    nn = int(lines[i + 1])
    nodes = np.zeros((nn, 4))
    for k in range(nn):
        t = lines[i + 2 + k].split()
        nodes[k] = [float(t[0]), float(t[1]), float(t[2]), float(t[3])]
    j = lines.index("$Elements")
    ne = int(lines[j + 1])
    elems = -np.ones((ne, 12), np.int64)  # id, type, ntags, tags..., nodes...
    for k in range(ne):
        t = [int(x) for x in lines[j + 2 + k].split()]
        assert len(t) <= 12, (path, t)
        elems[k, : len(t)] = t
    return nodes, elems


def main():
    out = {}
    for name, rel in MESHES.items():
        nodes, elems = parse_msh(os.path.join(REF, rel))
        out[name + "__nodes"] = nodes
        out[name + "__elems"] = elems.astype(np.int32)
    np.savez_compressed(os.path.join(OUT, "meshes.npz"), **out)
    ana = np.loadtxt(os.path.join(REF, "DG-rectangular_structured_analytical"))
    num = np.loadtxt(os.path.join(REF, "DG-rectangular_structured"))
    np.savez_compressed(os.path.join(OUT, "rect_golden.npz"), analytical=ana, numerical=num)
    print("wrote", os.listdir(OUT))


if __name__ == "__main__":
    main()

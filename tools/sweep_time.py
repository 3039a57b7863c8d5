#!/usr/bin/env python
"""Sustained time of a level-1 sweep on c5: `reps` x `count` back-to-back sweeps of each smoother (a board at its power
limit settles after ~0.3 s: compare count = 100 with count = 2000).  PAMG_LIB selects an A/B build (tools/ab_build.sh)."""
import importlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("p-a_multigrids_b200")


def main():
    count = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    mesh = pkg.Mesh.synthetic(4, 1)
    p = pkg.default_params(n_split=8, multi_levels=1, u_x=0.9, u_y=0.3, dt=1e-3)
    g = pkg.SemiImplicitIterative(p, mesh)
    rng = np.random.Generator(np.random.MT19937(1))
    nd = g.ndof(1)
    g.upload(pkg.TNONLIN, 1, rng.random(nd)); g.copy(1, pkg.TNEW, pkg.TNONLIN); g.upload(pkg.TOLD, 1, rng.random(nd))
    lib = os.path.basename(os.environ.get("PAMG_LIB", "default"))
    for name, solver in (("jacobi", pkg.JACOBI), ("gs", pkg.GAUSS_SEIDEL)):
        g.smoother(1, solver, 10); g.sync()
        ts = []
        for _ in range(reps):
            g.event_record(0); g.smoother(1, solver, count); g.event_record(1); g.sync()
            ts.append(g.elapsed_ms(0, 1) / count * 1e3)
        print(lib, name, count, "sweeps:", " ".join("%.2f" % t for t in ts), "us per sweep")
    g.close()


if __name__ == "__main__":
    main()

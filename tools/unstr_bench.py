#!/usr/bin/env python
"""Times the unstructured front-end kernels alone at HBM scale (4^PAMG_BENCH_UNSTR_KP triangles, default 4^12 = 16.7 M):
k_unstr_explicit (144 B/element), k_assemble_bsr (456 B/element, with and without the diffusion blocks), k_bsr_spmv
(352 B/element).  Prints one JSON line.  PAMG_LIB selects an A/B build (tools/ab_build.sh)."""
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("p-a_multigrids_b200")


def main():
    kp = int(os.environ.get("PAMG_BENCH_UNSTR_KP", "12"))
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        peak = 6555.2
    um = pkg.Mesh.synthetic(kp, 1)
    g = pkg.SemiImplicitIterative(pkg.default_params(n_split=1, multi_levels=1, u_x=0.9, u_y=0.3, dt=1e-3), pkg.Mesh.synthetic(1, 1))
    g.set_unstructured(um)
    E = um.U
    T0 = np.random.Generator(np.random.MT19937(7)).random((E, 3))
    g._ck(g.L.pamg_unstr_upload(g.h, T0))
    out = {"elements": E, "lib": os.path.basename(pkg.pamg.LIB_PATH) if hasattr(pkg, "pamg") else ""}

    def timed(fn, reps, warm=2):
        for _ in range(warm):
            fn()
        g.sync()
        g.event_record(8)
        for _ in range(reps):
            fn()
        g.event_record(9)
        g.sync()
        return g.elapsed_ms(8, 9) / reps

    # explicit: one time step of 10 passes (one told copy per call, 10 kernels)
    t = timed(lambda: g._ck(g.L.pamg_explicit_step(g.h, 1e-6, 0.9, 0.3, 0.0, 1, 10, 10, 0, 0)), 3) / 10.0
    out["explicit_ms"] = t; out["explicit_frac"] = 144.0 * E / (t * 1e-3) / 1e9 / peak
    area = 0.5 * np.abs((um.X[:, 0, 0] - um.X[:, 2, 0]) * (um.X[:, 1, 1] - um.X[:, 2, 1])
                        - (um.X[:, 0, 1] - um.X[:, 2, 1]) * (um.X[:, 1, 0] - um.X[:, 2, 0]))
    dt_i = 4.0 * float(np.sqrt(area.min()))
    t = timed(lambda: g.implicit_assemble(dt_i, 0.9, 0.3, use_dir=True), 10)
    out["assemble_ms"] = t; out["assemble_frac"] = 456.0 * E / (t * 1e-3) / 1e9 / peak
    t = timed(lambda: g.implicit_assemble(dt_i, 0.9, 0.3, use_dir=True, k=1.0), 10)
    out["assemble_diff_ms"] = t; out["assemble_diff_frac"] = 480.0 * E / (t * 1e-3) / 1e9 / peak
    g.implicit_assemble(dt_i, 0.9, 0.3, use_dir=True)
    t = g.implicit_spmv_ms(20)
    out["spmv_ms"] = t; out["spmv_frac"] = 352.0 * E / (t * 1e-3) / 1e9 / peak
    print(json.dumps(out))
    g.close()


if __name__ == "__main__":
    main()

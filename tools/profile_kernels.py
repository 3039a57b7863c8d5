#!/usr/bin/env python
"""Small driver for `ncu`: a handful of launches of every hot kernel at HBM scale (c5 for the semi-structured kernels,
4^11 triangles for the unstructured front-ends).  Usage (B200_PROFILING.md: plain run first, then the same command under ncu):

    python tools/profile_kernels.py > gpurun_out/plain.log 2>&1 &&
    ncu --set full --clock-control none --import-source on \
        -k regex:'k_element_win2|k_gs_win2|k_assemble_bsr|k_unstr_explicit|k_bsr_spmv|k_kry' -c 16 -o gpurun_out/prof python tools/profile_kernels.py
"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("p-a_multigrids_b200")


def main():
    n = int(os.environ.get("PAMG_PROFILE_NSPLIT", "8"))
    mesh = pkg.Mesh.synthetic(4, 1)
    p = pkg.default_params(n_split=n, multi_levels=1, u_x=0.9, u_y=0.3, dt=1e-3)
    g = pkg.SemiImplicitIterative(p, mesh)
    rng = np.random.Generator(np.random.MT19937(1))
    nd = g.ndof(1)
    g.upload(pkg.TNONLIN, 1, rng.random(nd)); g.copy(1, pkg.TNEW, pkg.TNONLIN); g.upload(pkg.TOLD, 1, rng.random(nd))
    g.smoother(1, pkg.JACOBI, 3)            # k_build_rhs, k_halo (first sweep only), 3 x k_element_win2<JACOBI>
    g.smoother(1, pkg.GAUSS_SEIDEL, 2)      # 2 x k_gs_win2
    g.get_residual(1)                       # k_element_win2<RESID> + k_reduce_partials
    g.sync()
    um = pkg.Mesh.synthetic(int(os.environ.get("PAMG_PROFILE_UNSTR_KP", "11")), 1)
    g.set_unstructured(um)
    T0 = rng.random((um.U, 3))
    g._ck(g.L.pamg_unstr_upload(g.h, T0))
    g._ck(g.L.pamg_explicit_step(g.h, 1e-6, 0.9, 0.3, 0.0, 1, 2, 10, 0, 0))      # 2 x k_unstr_explicit
    area = 0.5 * np.abs((um.X[:, 0, 0] - um.X[:, 2, 0]) * (um.X[:, 1, 1] - um.X[:, 2, 1])
                        - (um.X[:, 0, 1] - um.X[:, 2, 1]) * (um.X[:, 1, 0] - um.X[:, 2, 0]))
    dt_i = 4.0 * float(np.sqrt(area.min()))
    g.implicit_assemble(dt_i, 0.9, 0.3, use_dir=True)                            # k_assemble_bsr
    g.implicit_assemble(dt_i, 0.9, 0.3, use_dir=True, k=1.0)                     # k_assemble_bsr with the diffusion blocks
    print("spmv ms", g.implicit_spmv_ms(2))                                      # 3 x k_bsr_spmv
    import ctypes as C
    it = C.c_int(0); rr = C.c_double(0)
    g._ck(g.L.pamg_implicit_step(g.h, 1, 1, 1e-2, 3, C.byref(it), C.byref(rr)))  # a few BiCGStab iterations (k_kry_*)
    g.sync()
    print("profile driver ok", it.value, rr.value)
    g.close()


if __name__ == "__main__":
    main()

import os, sys, time
os.environ["PAMG_P2P_TIMEOUT_S"] = "5"
sys.path.insert(0, "tests")
import numpy as np
from pamg_pkg import pamg
kp, n = 2, 6
for G_super, devs in ((1, [0, 0]), (1, [0, 0, 0]), (2, [0, 0])):
    mesh = pamg.Mesh.synthetic(kp, G_super)
    params = pamg.default_params(n_split=n, multi_levels=n, u_x=0.9, u_y=0.3)
    g = pamg.SemiImplicitIterative(params, mesh, devices=devs)
    nparts = len(devs)
    pf = [mesh.U * i // nparts for i in range(nparts + 1)]
    for part in range(nparts):
        pl = pamg.halo_plan(mesh, 1, nparts, np.array(pf, np.int32), part)
        print("G", G_super, devs, "part", part, "peers", pl["peers"].tolist(), "nstrips", pl["nstrips"], "nsend", pl["nsend"], flush=True)
    rng = np.random.default_rng(1)
    for lvl in (1, 2, 3, 4, 5, 6):
        Tl = rng.random((mesh.U, 4 ** (n - lvl + 1), 3))
        g.upload(pamg.TNONLIN, lvl, Tl); g.upload(pamg.RHS, lvl, 0.5 * Tl)
        for what, fn in (("jacobi", lambda: g.smoother(lvl, pamg.JACOBI, 2)), ("gs", lambda: g.smoother(lvl, pamg.GAUSS_SEIDEL, 1)),
                         ("upd", lambda: g.update_overlaps(lvl))):
            t0 = time.time()
            try:
                fn(); g.sync()
                print("  lvl", lvl, what, "ok %.3f s" % (time.time() - t0), flush=True)
            except Exception as e:
                print("  lvl", lvl, what, "FAIL %.3f s" % (time.time() - t0), e, flush=True)
                break
    g.close()

import importlib, os, sys, numpy as np
sys.path.insert(0, os.getcwd())
pkg = importlib.import_module("p-a_multigrids_b200")
mesh = pkg.Mesh.synthetic(4, 1)
p = pkg.default_params(n_split=8, multi_levels=1, u_x=0.9, u_y=0.3, dt=1e-3)
g = pkg.SemiImplicitIterative(p, mesh)
rng = np.random.Generator(np.random.MT19937(1))
nd = g.ndof(1)
g.upload(pkg.TNONLIN, 1, rng.random(nd)); g.copy(1, pkg.TNEW, pkg.TNONLIN); g.upload(pkg.TOLD, 1, rng.random(nd))
for name, solver in (("jacobi", pkg.JACOBI), ("gs", pkg.GAUSS_SEIDEL)):
    g.smoother(1, solver, 10); g.sync()
    best = 1e9
    for rep in range(3):
        g.event_record(0); g.smoother(1, solver, 100); g.event_record(1); g.sync()
        best = min(best, g.elapsed_ms(0, 1) / 100)
    print(os.path.basename(os.environ.get("PAMG_LIB", "default")), name, round(best * 1e3, 2), "us")
g.close()

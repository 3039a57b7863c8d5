"""Single-process multi-GPU handle (pamg_create_multi: one host thread drives N devices, the form the reference's serial
driver main.F90:16-51 can call) against the single-GPU handle and the oracle.

The parts exchange their cut-face strips by flagged stores into each other's memory and poll for them inside the halo
kernel; small coarse levels are pushed to part 0.  A device may appear more than once in the device list, so the whole
protocol - strip placement, reversal, epochs, double buffering, the coarse gather / scatter - is exercised on a ONE-GPU
box too (devices = [0, 0, ...]); with real peers the same test runs over NVLink."""
import os

import numpy as np
import pytest

import oracle_api as orc
from helpers import rel_l2
from pamg_pkg import pamg

pytestmark = pytest.mark.gpu
os.environ.setdefault("PAMG_P2P_TIMEOUT_S", "30")     # a protocol bug must fail the test, not hang the box


def device_lists():
    n = pamg.device_count()
    out = [("same-device x2", [0, 0]), ("same-device x3", [0, 0, 0])]
    if n >= 2:
        out.append(("two GPUs", [0, 1]))
    if n >= 4:
        out.append(("four GPUs", [0, 1, 2, 3]))
    return out


def rnd(shape, seed):
    return np.random.Generator(np.random.MT19937(seed)).random(shape)


@pytest.mark.parametrize("G_super", [1, 2])
@pytest.mark.parametrize("label,devices", device_lists())
def test_group_equals_single_gpu_and_oracle(label, devices, G_super):
    """G_super = 2: the cut follows the seam between the two super-triangles (when len(devices) == 2); G_super = 1: the 16
    parents of one super-triangle are cut into uneven blocks through its interior."""
    kp, n = 2, 6
    mesh = pamg.Mesh.synthetic(kp, G_super)
    params = pamg.default_params(n_split=n, multi_levels=n, u_x=0.9, u_y=0.3)
    g = pamg.SemiImplicitIterative(params, mesh, devices=devices)
    ref = pamg.SemiImplicitIterative(params, mesh)
    o = orc.Semi(orc.intended_params(n, n, dt=params.dt, u=(0.9, 0.3)), mesh.X, mesh.neig, mesh.fneig, mesh.dir)
    orc.lib().orc_semi_set_threads(os.cpu_count() or 1)
    shape = (mesh.U, 4 ** n, 3)
    assert g.ndof(1) == ref.ndof(1) == int(np.prod(shape))
    T, Told = rnd(shape, 4242), rnd(shape, 4243)
    for s in (g, ref):
        s.upload(pamg.TNONLIN, 1, T); s.copy(1, pamg.TNEW, pamg.TNONLIN); s.upload(pamg.TOLD, 1, Told)
    o.field(orc.TNONLIN)[:] = T; o.field(orc.TOLD)[:] = Told
    # halo strips of every parent (cut faces: received from the other part) - bit exact
    g.update_overlaps(1); ref.update_overlaps(1)
    o.field(orc.TNEW)[:] = T; o.update_overlaps(1)
    ovg = g.overlap(1)
    assert np.array_equal(ovg, ref.overlap(1))
    assert np.array_equal(g.overlap(1, old=True), ref.overlap(1, old=True))      # t_overlap_old travels too
    assert rel_l2(ovg, o.overlap(1)) <= 1e-13          # (Dirichlet entries: device sin() against libm)
    # sweeps: Jacobi then two-colour GS
    for s in (g, ref):
        s.smoother(1, pamg.JACOBI, 3)
        s.smoother(1, pamg.GAUSS_SEIDEL, 2)
    o.smooth(1, 1, 3); o.smooth(1, 4, 2)
    got = g.download(pamg.TNONLIN, 1)
    assert rel_l2(got, ref.download(pamg.TNONLIN, 1)) <= 1e-14
    assert rel_l2(got, o.field(orc.TNONLIN)) <= 1e-12
    # residual + norms (combined over the parts on the host)
    for s in (g, ref):
        s.copy(1, pamg.TNEW, pamg.TNONLIN); s.update_overlaps(1)
    l2g, lig = g.get_residual(1); l2r, lir = ref.get_residual(1)
    assert abs(l2g - l2r) <= 1e-13 * l2r and lig == lir
    assert rel_l2(g.download(pamg.RES, 1), ref.download(pamg.RES, 1)) <= 1e-14
    # coarser levels (other kernel families, the exchange on every level)
    for lvl in (2, 4, 6):
        Tl = rnd((mesh.U, 4 ** (n - lvl + 1), 3), 100 + lvl)
        for s in (g, ref):
            s.upload(pamg.TNONLIN, lvl, Tl); s.upload(pamg.RHS, lvl, 0.5 * Tl)
            s.smoother(lvl, pamg.JACOBI, 2); s.smoother(lvl, pamg.GAUSS_SEIDEL, 1)
        assert rel_l2(g.download(pamg.TNONLIN, lvl), ref.download(pamg.TNONLIN, lvl)) <= 1e-14, lvl
    # V-cycles to 1e-8 (CUDA graph per part from the second cycle on, coarse levels agglomerated on part 0)
    for s in (g, ref):
        s.fill(pamg.TNONLIN, 1, 0.0); s.copy(1, pamg.TNEW, pamg.TNONLIN); s.fill(pamg.TOLD, 1, 0.0)
    cg, hg = g.vcycle_solve(solver=pamg.GAUSS_SEIDEL, max_cycles=40, tol=1e-8)
    cr, hr = ref.vcycle_solve(solver=pamg.GAUSS_SEIDEL, max_cycles=40, tol=1e-8)
    assert cg == cr and hg[-1] <= 1e-8 * hg[0]
    np.testing.assert_allclose(hg, hr, rtol=1e-6)
    assert np.max(np.abs(g.download(pamg.TNONLIN, 1) - ref.download(pamg.TNONLIN, 1))) <= 1e-10
    # a second solve with the other smoother (another set of cached graphs), again from zero
    for s in (g, ref):
        s.fill(pamg.TNONLIN, 1, 0.0); s.copy(1, pamg.TNEW, pamg.TNONLIN)
    cg2, hg2 = g.vcycle_solve(solver=pamg.JACOBI, max_cycles=40, tol=1e-8)
    cr2, hr2 = ref.vcycle_solve(solver=pamg.JACOBI, max_cycles=40, tol=1e-8)
    assert cg2 == cr2
    np.testing.assert_allclose(hg2, hr2, rtol=1e-6)
    g.close(); ref.close()


@pytest.mark.parametrize("label,devices", device_lists()[:1] + device_lists()[2:3])
def test_group_host_buffer_entries(label, devices):
    """pamg_smoother_host / pamg_smooth_host / pamg_timestep_host take and return WHOLE-mesh host arrays on a group."""
    kp, n = 2, 6
    mesh = pamg.Mesh.synthetic(kp, 2)
    params = pamg.default_params(n_split=n, multi_levels=n, u_x=0.9, u_y=0.3, solver=pamg.GAUSS_SEIDEL)
    g = pamg.SemiImplicitIterative(params, mesh, devices=devices)
    ref = pamg.SemiImplicitIterative(params, mesh)
    nd = g.ndof(1)
    a, b, c = pamg.PinnedBuffer(nd), pamg.PinnedBuffer(nd), pamg.PinnedBuffer(nd)
    a.array[:] = rnd(nd, 7)
    Told = rnd((mesh.U, 4 ** n, 3), 8)
    g.upload(pamg.TOLD, 1, Told); ref.upload(pamg.TOLD, 1, Told)
    g.smoother_host(pamg.JACOBI, 4, a.ptr, b.ptr)
    ref.smoother_host(pamg.JACOBI, 4, a.ptr, c.ptr)
    assert rel_l2(b.array, c.array) <= 1e-14
    # dependent pipelined loop: the result of a call is the input of the next one (and then in place)
    g.smooth_host(pamg.JACOBI, 2, a.ptr, b.ptr)
    g.smooth_host(pamg.JACOBI, 2, b.ptr, b.ptr)
    g.sync()
    ref.smoother_host(pamg.JACOBI, 4, a.ptr, c.ptr)
    assert rel_l2(b.array, c.array) <= 1e-14
    cyc, rel = g.timestep_host(a.ptr, b.ptr, max_cycles=40, tol=1e-8)
    cyc_r, rel_r = ref.timestep_host(a.ptr, c.ptr, max_cycles=40, tol=1e-8)
    assert cyc == cyc_r and rel <= 1e-8
    assert np.max(np.abs(b.array - c.array)) <= 1e-10
    for x in (a, b, c):
        x.free()
    g.close(); ref.close()

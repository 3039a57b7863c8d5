"""The C++ host driver (pamg_host, mirror of main.F90 modes 1, 2, 4, 5 and 9) against the oracle."""
import os
import re
import subprocess

import numpy as np
import pytest

import oracle_api as orc
from helpers import ROOT, write_msh
from pamg_pkg import pamg

pytestmark = pytest.mark.gpu
HOST = os.path.join(ROOT, "p-a_multigrids_b200", "lib", "pamg_host")


def run_host(*args):
    r = subprocess.run([HOST, *map(str, args)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    return r.stdout


def test_mode9_as_checked_in(tmp_path):
    """main.F90:46-47 on test_sn2.msh: n_split 1, 1 level, GS, 4 sweeps, 2 cycles, 2 steps, IC region 4."""
    path = write_msh("test_sn2", str(tmp_path / "test_sn2.msh"))
    out = run_host("--mode", 9, "--mesh", path)
    m = re.search(r"tnew: sum ([-+0-9.e]+) min ([-+0-9.e]+) max ([-+0-9.e]+)", out)
    assert m, out
    mesh = pamg.Mesh.read_msh(path)
    o = orc.Semi(orc.literal_params(1, 1), mesh.X, mesh.neig, mesh.fneig, mesh.dir)
    ic = np.zeros(o.field(orc.TNEW).shape)
    ic[mesh.region == 4] = 1.0
    o.field(orc.TNEW)[:] = ic
    for _ in range(2):
        o.literal_timestep(solver=3, n_multigrid=2, n_smooth=4)
    ref = o.field(orc.TNEW)
    assert abs(float(m.group(1)) - ref.sum()) <= 1e-9 * max(1.0, abs(ref.sum()))
    assert abs(float(m.group(2)) - ref.min()) <= 1e-6 and abs(float(m.group(3)) - ref.max()) <= 1e-6
    assert "totele_unst, totele_str, totele 12 4 48" in out


def test_mode9_intended_vcycles(tmp_path):
    path = write_msh("split0", str(tmp_path / "split0.msh"))
    out = run_host("--mode", 9, "--mesh", path, "--intended", "--n_split", 5, "--multi_levels", 5, "--ntime", 1,
                   "--solver", 3, "--region", 99)
    m = re.search(r"V-cycles (\d+)\s+\|\|r\|\|/\|\|r0\|\| ([0-9.e+-]+)", out)
    assert m, out
    mesh = pamg.Mesh.read_msh(path)
    o = orc.Semi(orc.intended_params(5, 5), mesh.X, mesh.neig, mesh.fneig, mesh.dir)
    it, _ = o.vcycle_solve(solver=4, max_cycles=50, tol=1e-8)
    assert abs(int(m.group(1)) - it) <= 1 and float(m.group(2)) <= 1e-8


def test_mode4_unstr_explicit(tmp_path):
    path = write_msh("untitled8192", str(tmp_path / "u.msh"))
    out = run_host("--mode", 4, "--mesh", path)
    m = re.search(r"tnew: sum ([-+0-9.e]+) min ([-+0-9.e]+) max ([-+0-9.e]+)", out)
    assert m, out
    mesh = pamg.Mesh.read_msh(path)
    T = np.zeros((mesh.U, 3))
    T[mesh.region == 12] = 1.0
    orc.lib().orc_unstr_explicit(mesh.U, mesh.X, mesh.neig, mesh.fneig, mesh.dir, 0.9, 0.0, 0.07e-3, 2, 2, 10, 0, 0, 0.0, T)
    assert abs(float(m.group(1)) - T.sum()) <= 1e-9 * abs(T.sum())
    assert "totele = 8192" in out


def test_mode2_str_explicit_literal_arguments():
    """pamg_host --mode 2 = case(2) of main.F90:22: str_explicit(0.7, ..., njac_its 10, nits 2, no_ele_row 20, no_ele_col 2,
    dx = dy = 0.1, u = (0.05, 0.05)); totele = 20 * 2 * 2, ntime = 100, the pulse of transport_tri.F90:457-460."""
    out = run_host("--mode", 2)
    m = re.search(r"tnew: sum ([-+0-9.e]+) min ([-+0-9.e]+) max ([-+0-9.e]+)", out)
    assert m and "totele = 80" in out and "ntime = 100" in out, out
    ner, nec = 20, 2
    mesh = pamg.Mesh.structured_tri(ner, 2 * nec, 0.1, 0.1)
    T = np.zeros((mesh.U, 3))
    T[0:ner // 5 + 1] = 1.0
    for i in range(2, nec + 1):
        T[(i - 1) * ner: ner * i - (ner * 4 // 5)] = 1.0
    orc.lib().orc_unstr_explicit(mesh.U, mesh.X, mesh.neig, mesh.fneig, mesh.dir, 0.05, 0.05, 0.7 * 0.1, 100, 2, 10, 0, 0, 0.0, T)
    assert abs(float(m.group(1)) - T.sum()) <= 1e-9 * abs(T.sum())
    assert abs(float(m.group(2)) - T.min()) <= 1e-6 and abs(float(m.group(3)) - T.max()) <= 1e-6


def test_mode5_unstr_implicit_literal_arguments(tmp_path):
    """pamg_host --mode 5 = case(5) of main.F90:31 on Mesh_files/gmsh_100.msh: dt = 0.7 * 0.1, ntime = 2, nits = 2,
    u = (-0.1, 0.1), tnew(:,5) = 1; the dense FINDInv solve of the reference (oracle) against the Krylov solve on the device."""
    path = write_msh("gmsh_100", str(tmp_path / "gmsh_100.msh"))
    out = run_host("--mode", 5, "--mesh", path)
    m = re.search(r"tnew: sum ([-+0-9.e]+) min ([-+0-9.e]+) max ([-+0-9.e]+)", out)
    assert m and "ntime = 2" in out, out
    mesh = pamg.Mesh.read_msh(path)
    T = np.zeros((mesh.U, 3))
    T[4] = 1.0
    assert orc.lib().orc_unstr_implicit(mesh.U, mesh.X, mesh.neig, mesh.fneig, -0.1, 0.1, 0.07, 2, 2, 0, T) == 0
    assert abs(float(m.group(1)) - T.sum()) <= 1e-9 * abs(T.sum())
    assert abs(float(m.group(2)) - T.min()) <= 1e-6 and abs(float(m.group(3)) - T.max()) <= 1e-6
    it = re.search(r"Krylov iterations (\d+)  worst \|\|r\|\|/\|\|b\|\| ([0-9.e+-]+)", out)
    assert it and float(it.group(2)) <= 1e-12, out


def test_mode1_trans_rec_writes_the_reference_dumps(tmp_path):
    """pamg_host --mode 1 = case(1) of main.F90:19; its two dump files against the files the reference ships."""
    out = run_host("--mode", 1, "--mesh", str(tmp_path))
    assert "ntime = 714" in out
    from helpers import GOLDEN
    g = np.load(os.path.join(GOLDEN, "rect_golden.npz"))
    num = np.loadtxt(tmp_path / "DG-rectangular_structured")
    ana = np.loadtxt(tmp_path / "DG-rectangular_structured_analytical")
    assert num.shape == (800, 3) and ana.shape == (800, 2)
    assert np.array_equal(num[:, :2], g["numerical"][:, :2])
    assert np.max(np.abs(num[:, 2] - g["numerical"][:, 2])) <= 3e-5       # the reference computed in single precision
    assert np.allclose(ana[:, 0], g["analytical"][:, 0], atol=1e-5) and np.array_equal(ana[:, 1], g["analytical"][:, 1])


def test_mode9_on_several_gpus_from_one_process(tmp_path):
    """--devices 0,0,0 (or real GPUs when the box has them): the serial driver's call sequence, unchanged, on a handle that
    fans out over several parts; the printed checksums equal the one-GPU run."""
    path = write_msh("900_ele", str(tmp_path / "m.msh"))
    n = pamg.device_count()
    devs = ",".join(str(i) for i in range(min(n, 4))) if n >= 2 else "0,0,0"
    common = ["--mode", 9, "--mesh", path, "--intended", "--n_split", 3, "--multi_levels", 3, "--ntime", 2, "--solver", 3,
              "--region", 9, "--ux", 0.1, "--uy", 0.1]
    env_timeout = dict(os.environ, PAMG_P2P_TIMEOUT_S="30")
    one = subprocess.run([HOST, *map(str, common)], capture_output=True, text=True, timeout=300, env=env_timeout)
    many = subprocess.run([HOST, *map(str, common), "--devices", devs], capture_output=True, text=True, timeout=300, env=env_timeout)
    assert one.returncode == 0 and many.returncode == 0, one.stdout + one.stderr + many.stdout + many.stderr
    pat = r"tnew: sum ([-+0-9.e]+) min ([-+0-9.e]+) max ([-+0-9.e]+)"
    a, b = re.search(pat, one.stdout), re.search(pat, many.stdout)
    assert a and b, many.stdout
    for i in (1, 2, 3):
        assert abs(float(a.group(i)) - float(b.group(i))) <= 1e-9 * max(1.0, abs(float(a.group(i))))
    assert re.findall(r"V-cycles (\d+)", one.stdout) == re.findall(r"V-cycles (\d+)", many.stdout)
    assert "GPUs driven by this process" in many.stdout

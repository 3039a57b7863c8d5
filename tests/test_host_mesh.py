"""Host-side logic of the product (no GPU): mesh reader + O(N) neighbour search against the oracle's
literal all-pairs restatement, the halo plan, and the C-ABI surface."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import oracle_api as orc
from helpers import ROOT, write_msh
from pamg_pkg import pamg

ALL = ["test_sn2", "900_ele", "untitled8192", "untitled2048", "untitled8", "2_unele_test", "irregular",
       "semi_mesh", "gmsh_100", "split0", "split1", "split2", "split3", "split4"]


@pytest.mark.parametrize("name", ALL)
def test_edge_hash_pipeline_equals_allpairs_reference_search(name, tmp_path):
    path = write_msh(name, str(tmp_path / (name + ".msh")))
    o = orc.read_msh(path)
    m = pamg.Mesh.read_msh(path)
    assert m.U == o["X"].shape[0]
    assert np.array_equal(m.X, o["X"])
    assert np.array_equal(m.region, o["region"])
    assert np.array_equal(m.neig, o["neig"])
    assert np.array_equal(m.dir, o["dir"])
    fneig, _ = orc.neig_data(o["neig"], o["dir"])
    assert np.array_equal(m.fneig, fneig)


def test_read_msh_errors(tmp_path):
    with pytest.raises(pamg.PamgError) as e:
        pamg.Mesh.read_msh(str(tmp_path / "missing.msh"))
    assert e.value.code == pamg.pamg.ERR_IO
    bad = tmp_path / "bad.msh"
    bad.write_text("$NotAMesh\n")
    with pytest.raises(pamg.PamgError):
        pamg.Mesh.read_msh(str(bad))
    binary = tmp_path / "bin.msh"
    binary.write_text("$MeshFormat\n2.2 1 8\n$EndMeshFormat\n")
    with pytest.raises(pamg.PamgError):
        pamg.Mesh.read_msh(str(binary))


@pytest.mark.parametrize("kp,G", [(0, 1), (1, 1), (2, 2), (3, 4), (2, 8)])
def test_synthetic_mesh_matches_oracle_splitting_and_search(kp, G):
    m = pamg.Mesh.synthetic(kp, G)
    per = 4 ** kp
    assert m.U == G * per
    L = orc.lib()
    for g in range(G):
        sq = g // 2
        P = (np.array([[sq + 1, 0], [sq, 1], [sq, 0]], float) if g % 2 == 0
             else np.array([[sq, 1], [sq + 1, 0], [sq + 1, 1]], float))
        for e in range(1, per + 1):
            x = np.zeros((3, 2))
            L.orc_get_splitting(P.copy(), kp, e, x)
            assert np.array_equal(m.X[g * per + e - 1], x)
    # neighbour search agrees with the literal all-pairs search on the same triangles
    m2 = pamg.Mesh.from_arrays(m.X)
    assert np.array_equal(m2.neig, m.neig) and np.array_equal(m2.dir, m.dir)
    area = 0.5 * np.abs((m.X[:, 0, 0] - m.X[:, 2, 0]) * (m.X[:, 1, 1] - m.X[:, 2, 1])
                        - (m.X[:, 0, 1] - m.X[:, 2, 1]) * (m.X[:, 1, 0] - m.X[:, 2, 0]))
    assert np.isclose(area.sum(), 0.5 * G)
    nb = (m.neig == 0).sum()
    # boundary parent faces: perimeter of the strip in units of 2^-kp
    assert nb == (2 ** kp) * (3 if G == 1 else (G + 2 if G % 2 == 0 else G + 2))


@pytest.mark.parametrize("name", ["test_sn2", "untitled8192", "split2", "irregular"])
def test_halo_plan_single_part_matches_oracle_strips(name, tmp_path):
    """strip placement + reversal + node map reproduce the oracle's update_overlaps for both rules."""
    m = pamg.Mesh.read_msh(write_msh(name, str(tmp_path / (name + ".msh"))))
    n = 2
    S = 2 ** n
    rng = np.random.default_rng(1)
    for rule in (0, 1):
        p = orc.intended_params(n, 1)
        p.halo_rule = rule
        s = orc.Semi(p, m.X, m.neig, m.fneig, m.dir)
        T = s.field(orc.TNEW, 1)
        T[:] = rng.random(T.shape)
        s.update_overlaps(1)
        ref = s.overlap(1)
        plan = pamg.halo_plan(m, halo_rule=rule)
        assert plan["nsend"] == 0 and plan["nstrips"] == 3 * m.U and len(plan["peers"]) == 0
        surf = np.zeros(3 * S, np.int32)
        orc.lib().orc_surf_ele(n, surf)
        surf = surf.reshape(3, S)
        strips = np.zeros((3 * m.U, S, 3))
        for u in range(m.U):
            for mf in range(3):
                lf = u * 3 + mf
                d = plan["dst_strip"][lf]
                if d < 0:
                    continue
                for i in range(S):
                    slot = S - 1 - i if plan["rev"][lf] else i
                    strips[d, slot] = T[u, surf[mf, i] - 1]
        got = strips[plan["strip_of"]].reshape(m.U, 3, S, 3)
        interior = (m.neig != 0)
        assert np.array_equal(got[interior], ref[interior])


def test_halo_plan_partition_pairs_cut_faces_consistently():
    m = pamg.Mesh.synthetic(2, 4)
    pf = np.array([0, 16, 32, 48, 64], np.int32)
    plans = [pamg.halo_plan(m, 1, 4, pf, r) for r in range(4)]
    for r, pl in enumerate(plans):
        for part, nfaces, sb, _ in pl["peers"]:
            other = plans[part]
            row = [q for q in other["peers"] if q[0] == r]
            assert len(row) == 1 and row[0][1] == nfaces
            # k-th receive strip of `part` from r belongs to the face that r's k-th send slot targets
            inv_r = {pl["strip_of"][lf]: lf for lf in range(len(pl["strip_of"]))}
            inv_o = {other["strip_of"][lf]: lf for lf in range(len(other["strip_of"]))}
            for k in range(nfaces):
                lf_r = inv_r[sb + k]
                lf_o = inv_o[row[0][2] + k]
                gu, mf = pf[r] + lf_r // 3, lf_r % 3
                go, mo = pf[part] + lf_o // 3, lf_o % 3
                assert m.neig[gu, mf] == go + 1 and m.fneig[gu, mf] == mo + 1
                assert pl["dst_strip"][lf_r] == pl["nstrips"] + sb + k


def test_c_abi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "pamg.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(pamg_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) > 40
    L = C.CDLL(pamg.LIB_PATH)
    missing = [n for n in sorted(names) if not hasattr(L, n)]
    assert not missing, missing
    # and the Python mirror binds all of them
    assert set(pamg.lib()._signatures) == names


def test_no_cpu_fallback_without_device():
    if pamg.device_count() > 0:
        pytest.skip("GPU present")
    with pytest.raises(pamg.PamgError) as e:
        pamg.SemiImplicitIterative(pamg.default_params(), pamg.Mesh.synthetic(1, 1))
    assert e.value.code == pamg.pamg.ERR_CUDA


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "p-a_multigrids_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".F90")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "liboracle" not in txt and "oracle_api" not in txt and "pamg_oracle" not in txt, f


def _tri_ele_info2(totele, ele, iface, no_ele_row):
    """structured_meshgen.F90:190-272 restated: neighbour across structured face iface (0 = boundary)."""
    import math
    row = math.ceil(ele / no_ele_row)
    if ele % 2 > 0:
        if iface == 1:
            e2 = ele - no_ele_row + 1
            return e2 if e2 >= 1 else 0
        e2 = ele - 1 if iface == 2 else ele + 1
    else:
        if iface == 1:
            e2 = ele + no_ele_row - 1
            return e2 if e2 <= totele else 0
        e2 = ele + 1 if iface == 2 else ele - 1
    return e2 if math.ceil(e2 / no_ele_row) == row else 0


@pytest.mark.parametrize("ner,nec,dx,dy", [(8, 4, 0.25, 0.5), (2, 1, 1.0, 1.0), (20, 10, 0.1, 0.1)])
def test_structured_triangle_mesh_matches_reference_generator(ner, nec, dx, dy):
    """pamg_mesh_structured_tri: coordinates of str_tri_X_nodes (:276-298) and the neighbours of tri_ele_info2
    (structured faces 1, 2, 3 are gmsh sides 1, 3, 2)."""
    m = pamg.Mesh.structured_tri(ner, nec, dx, dy)
    tot = ner * nec
    assert m.U == tot
    for ele in range(1, tot + 1):
        row = (ele + ner - 1) // ner
        col = ele - ner * (row - 1)
        if ele % 2:
            want = [(dx * (col // 2 + 1), dy * (row - 1)), (dx * (col // 2), dy * row), (dx * (col // 2), dy * (row - 1))]
        else:
            want = [(dx * (col // 2 - 1), dy * row), (dx * (col // 2), dy * (row - 1)), (dx * (col // 2), dy * row)]
        assert np.array_equal(m.X[ele - 1], np.array(want))
        for iface, side in ((1, 1), (2, 3), (3, 2)):
            assert m.neig[ele - 1, side - 1] == _tri_ele_info2(tot, ele, iface, ner), (ele, iface)
    # right triangles of area dx*dy/2 tiling the rectangle
    X = m.X
    det = (X[:, 0, 0] - X[:, 2, 0]) * (X[:, 1, 1] - X[:, 2, 1]) - (X[:, 0, 1] - X[:, 2, 1]) * (X[:, 1, 0] - X[:, 2, 0])
    assert np.allclose(0.5 * np.abs(det), 0.5 * dx * dy)
    assert pamg.lib().pamg_mesh_structured_tri(7, 2, 1.0, 1.0, None) == pamg.ERR_ARG

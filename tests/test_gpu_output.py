"""Output side of the path: x_all_str, analytical, get_error and the .vtu piece
(get_vtk_files.F90:10-140, transport_tri_semi.F90:274,278,299-312,531-540) against the oracle."""
import re
import struct
import xml.etree.ElementTree as ET

import numpy as np
import pytest

import oracle_api as orc
from helpers import rng_field, write_msh
from pamg_pkg import pamg

pytestmark = pytest.mark.gpu


def make(tmp_path, name="test_sn2", n_split=3):
    mesh = pamg.Mesh.read_msh(write_msh(name, str(tmp_path / (name + ".msh"))))
    p = pamg.default_params(n_split=n_split, multi_levels=1)
    g = pamg.SemiImplicitIterative(p, mesh)
    T = rng_field(g.shape(1), 4)
    g.upload(pamg.TNEW, 1, T)
    return g, mesh, T


def oracle_coords(mesh, n_split):
    C = 4 ** n_split
    X = np.zeros((mesh.U, C, 3, 2))
    out = np.zeros(6)
    for u in range(mesh.U):
        for e in range(C):
            orc.lib().orc_get_splitting(np.ascontiguousarray(mesh.X[u].ravel()), n_split, e + 1, out)
            X[u, e] = out.reshape(3, 2)
    return X


@pytest.mark.parametrize("name,n_split", [("test_sn2", 3), ("irregular", 4), ("split0", 1)])
def test_output_fields_match_oracle(tmp_path, name, n_split):
    g, mesh, T = make(tmp_path, name, n_split)
    x, an, er = g.output_fields()
    ref = oracle_coords(mesh, n_split)
    assert np.max(np.abs(x - ref)) <= 4e-16 * max(1.0, np.max(np.abs(ref)))   # FMA contraction on the device: <= 1 ulp
    ra = np.sin(ref[..., 0] + ref[..., 1])
    assert np.max(np.abs(an - ra)) <= 1e-15
    assert np.max(np.abs(er - np.abs(T - ra))) <= 1e-15


def _ascii_array(el):
    return np.array(el.text.split(), dtype=np.float64)


def test_vtu_ascii_reference_layout(tmp_path):
    g, mesh, T = make(tmp_path, "test_sn2", 2)
    x, an, er = g.output_fields()
    path = tmp_path / "Concentration_1.vtu"
    g.get_vtu(path, "Concentration", binary=False)
    root = ET.parse(path).getroot()
    assert root.tag == "VTKFile" and root.attrib["type"] == "UnstructuredGrid"
    piece = root.find("UnstructuredGrid/Piece")
    n = T.size // 3
    assert int(piece.attrib["NumberOfPoints"]) == 3 * n and int(piece.attrib["NumberOfCells"]) == n
    arrays = {a.attrib["Name"]: a for a in piece.find("PointData")}
    assert list(arrays) == ["Concentration", "error", "analytical"]
    assert np.max(np.abs(_ascii_array(arrays["Concentration"]) - T.ravel())) <= 0.5e-10 * (1 + 1e-6)     # F12.10
    assert np.max(np.abs(_ascii_array(arrays["error"]) - er.ravel())) <= 0.5e-7 * (1 + 1e-6)            # F10.7
    assert np.max(np.abs(_ascii_array(arrays["analytical"]) - an.ravel())) <= 0.5e-7 * (1 + 1e-6)
    pts = _ascii_array(piece.find("Points/DataArray")).reshape(-1, 3)
    assert np.max(np.abs(pts[:, :2] - x.reshape(-1, 2))) <= 0.5e-3 * (1 + 1e-9) and np.all(pts[:, 2] == 0)   # F10.3
    cells = {a.attrib["Name"]: _ascii_array(a) for a in piece.find("Cells")}
    assert np.array_equal(cells["connectivity"], np.arange(3 * n))
    assert np.array_equal(cells["offsets"], 3 * np.arange(1, n + 1))
    assert np.all(cells["types"] == 5)
    # one value per line like the reference's advance="yes" writes
    text = path.read_text()
    body = text.split('Name="Concentration" Format="ascii">\n')[1].split("</DataArray>")[0]
    assert all(re.fullmatch(r" {10}-?\d\.\d{10}  ", ln) for ln in body.splitlines() if ln.strip())


def test_vtu_binary_appended_round_trip(tmp_path):
    g, mesh, T = make(tmp_path, "irregular", 3)
    x, an, er = g.output_fields()
    path = tmp_path / "c.vtu"
    g.get_vtu(path, "Concentration", binary=True)
    raw = path.read_bytes()
    head, tail = raw.split(b'<AppendedData encoding="raw">\n   _', 1)
    root = ET.fromstring(head + b"</VTKFile>")
    n = T.size // 3
    offs = {}
    for a in root.iter("DataArray"):
        offs[a.attrib.get("Name", "points")] = (int(a.attrib["offset"]), a.attrib["type"])

    def block(name, dtype):
        o, _ = offs[name]
        (nbytes,) = struct.unpack_from("<Q", tail, o)
        return np.frombuffer(tail, dtype=dtype, count=nbytes // np.dtype(dtype).itemsize, offset=o + 8)

    assert np.array_equal(block("Concentration", np.float64), T.ravel())
    assert np.array_equal(block("error", np.float64), er.ravel())
    assert np.array_equal(block("analytical", np.float64), an.ravel())
    pts = block("points", np.float64).reshape(-1, 3)
    assert np.array_equal(pts[:, :2], x.reshape(-1, 2)) and np.all(pts[:, 2] == 0)
    assert np.array_equal(block("connectivity", np.int64), np.arange(3 * n))
    assert np.array_equal(block("offsets", np.int64), 3 * np.arange(1, n + 1))
    assert np.all(block("types", np.uint8) == 5) and block("types", np.uint8).size == n


def test_output_errors(tmp_path):
    g, _, _ = make(tmp_path, "split0", 1)
    L = pamg.lib()
    assert L.pamg_output_fields(g.h, None, None, None) == pamg.ERR_ARG
    assert L.pamg_write_vtu(g.h, str(tmp_path / "nodir" / "x.vtu").encode(), b"c", 0) == pamg.ERR_IO


def test_semi_structured_children_through_the_implicit_operator(tmp_path):
    """Semi_implicit_direct (transport_tri_semi.F90:1366-1783) assembles mass/dt - stiffness + upwind flux over the
    children of the semi-structured mesh.  Here: expand the children (pamg_output_fields), hand them to the block-CSR
    machinery (pamg_mesh_from_arrays -> pamg_set_unstructured -> pamg_implicit_*), and check against (a) the oracle's
    dense solve on the same arrays and (b) the same step on the gmsh-refined mesh 1_split.msh, whose triangles are the
    children of 0_split.msh at n_split = 1 in another order."""
    g, mesh0, _ = make(tmp_path, "split0", 1)
    x, _, _ = g.output_fields()
    kids = pamg.Mesh.from_arrays(x.reshape(-1, 3, 2))
    E = kids.U
    assert E == 4 * mesh0.U
    g.set_unstructured(kids)

    def field(X):      # a smooth nodal field evaluated at the nodes, so that it is the same function on both meshes
        return np.sin(3.0 * X[..., 0]) * np.cos(2.0 * X[..., 1]) + 0.5

    u, dt = (0.4, -0.7), 2e-2
    T0 = field(kids.X)
    got, iters, relres = g.unstr_implicit(T0, dt, u[0], u[1], ntime=2, nits=1, use_dir=True, tol=1e-13)
    ref = T0.copy()
    assert orc.lib().orc_unstr_implicit(E, kids.X, kids.neig, kids.fneig, u[0], u[1], dt, 2, 1, 1, ref) == 0
    assert np.linalg.norm(got - ref) <= 1e-10 * np.linalg.norm(ref)
    # (b) the independently refined mesh
    fine = pamg.Mesh.read_msh(write_msh("split1", str(tmp_path / "split1.msh")))
    assert fine.U == E
    g2 = pamg.SemiImplicitIterative(pamg.default_params(), pamg.Mesh.synthetic(0, 1))
    g2.set_unstructured(fine)
    got2, _, _ = g2.unstr_implicit(field(fine.X), dt, u[0], u[1], ntime=2, nits=1, use_dir=True, tol=1e-13)
    key = lambda X: np.round(X, 9)
    a = {tuple(key(p)): v for p, v in zip(kids.X.reshape(-1, 2), got.ravel())}
    # DG: a node value belongs to (element, node); match elements by their sorted vertex sets
    def by_element(X, T):
        out = {}
        for e in range(X.shape[0]):
            k = tuple(sorted(tuple(key(p)) for p in X[e]))
            out[k] = {tuple(key(p)): T[e, i] for i, p in enumerate(X[e])}
        return out
    A, B = by_element(kids.X, got), by_element(fine.X, got2)
    assert set(A) == set(B)
    worst = max(abs(A[k][p] - B[k][p]) for k in A for p in A[k])
    assert worst <= 1e-10
    assert len(a) > 0

"""The ISO_C_BINDING module a Fortran host binds (p-a_multigrids_b200/host/pamg_iface.F90) is generated from
include/pamg.h; it must be up to date and cover EVERY extern "C" entry with the same argument count and order.
(No Fortran compiler in the image: structural checks only.)"""
import importlib.util
import os
import re

from helpers import ROOT

spec = importlib.util.spec_from_file_location("gen_fortran_iface", os.path.join(ROOT, "tools", "gen_fortran_iface.py"))
gen = importlib.util.module_from_spec(spec)
spec.loader.exec_module(gen)


def module_text():
    with open(gen.OUT) as f:
        return f.read()


def test_module_is_up_to_date_with_the_header():
    assert module_text() == gen.generate(), "run python tools/gen_fortran_iface.py"


def test_every_entry_is_bound_with_the_same_arguments():
    funcs, fields, defines, enums = gen.parse_header()
    assert len(funcs) >= 64
    text = module_text()
    joined = re.sub(r"&\s*\n\s*", "", text)                    # undo continuation lines
    for ret, name, args in funcs:
        m = re.search(r"(subroutine|function)\s+%s\(([^)]*)\)\s*bind\(c, name=\"%s\"\)" % (name, name), joined)
        assert m, f"{name} is not bound"
        dummies = [a.strip() for a in m.group(2).split(",") if a.strip()]
        assert dummies == [an for _, an in args], name
        assert (m.group(1) == "subroutine") == (ret == "void"), name
        body = joined[m.end(): joined.index("end " + m.group(1), m.end())]
        for _, an in args:                                      # every dummy argument is declared exactly once
            assert len(re.findall(r"::\s*%s\b" % re.escape(an), body)) == 1, (name, an)
    # entries a driver cannot do without (the round-1 module missed them)
    for must in ("pamg_sync", "pamg_last_error", "pamg_timestep_host", "pamg_create_multi", "pamg_set_boundary_data",
                 "pamg_smoother_host", "pamg_set_parents_partition", "pamg_comm_init"):
        assert f'name="{must}"' in text


def test_params_type_mirrors_the_struct_and_lines_fit_free_form():
    funcs, fields, defines, enums = gen.parse_header()
    text = module_text()
    block = text[text.index("type, bind(c) :: pamg_params"): text.index("end type pamg_params")]
    names = re.findall(r"::\s*(\w+)", block)[1:]
    assert names == [nm for _, nm in fields]
    kinds = re.findall(r"^\s*(integer\(c_int32_t\)|real\(c_double\))", block, flags=re.M)
    assert kinds == [gen.FTYPE[t] for t, _ in fields]
    for k, v in defines + enums:
        assert re.search(r"parameter :: %s = %s\b" % (k, re.escape(v)), text)
    assert max(len(line) for line in text.splitlines()) <= 132
    assert len(re.findall(r"^  (end )?interface$", text, flags=re.M)) == 2 and text.rstrip().endswith("end module pamg_iface")


def test_header_exports_match_the_library():
    """every entry of pamg.h is exported by libpamg_cuda.so (no compute call: loading needs no GPU)"""
    import ctypes
    from pamg_pkg import pamg
    lib = ctypes.CDLL(pamg.LIB_PATH)
    funcs, _, _, _ = gen.parse_header()
    for _, name, _ in funcs:
        assert hasattr(lib, name), name
    assert set(pamg.lib()._signatures) == {name for _, name, _ in funcs}


def test_names_survive_fortran_case_insensitivity_and_attribute_keywords():
    """Fortran folds case and shares one namespace per scope: no two dummies of an entry, and no two module-level names, may differ
    by case only; a dummy called like the VALUE attribute would read `value :: value`."""
    funcs, fields, defines, enums = gen.parse_header()
    for ret, name, args in funcs:
        low = [an.lower() for _, an in args]
        assert len(set(low)) == len(low), name
        assert name.lower() not in low and "value" not in low, name
        assert all(len(an) <= 63 and not an.startswith("_") for an in low), name
    scope = [k.lower() for k, _ in defines + enums] + [n.lower() for _, n, _ in funcs] + ["pamg_params"]
    assert len(set(scope)) == len(scope)
    assert not re.search(r"\bvalue\s*::\s*value\b", module_text())

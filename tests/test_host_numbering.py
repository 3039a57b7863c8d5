"""The closed-form child numbering the kernels use (child_from_ele0 / child_from_flat / fine_children / child_advance,
csrc/pamg_kernels.cuh - the same functions compiled for the host behind pamg_numbering, no GPU needed) against the
reference's loops as restated by the oracle: get_str_info (Msh2Tri.F90:42-58), element_conversion (splitting.F90:105-139).

The memory-order closed form takes a float square root and corrects the guess; it is checked for EVERY child up to
n_split = 13 (67 M children, the largest split pamg_create accepts) against the exact integer definition of a row start."""
import ctypes

import numpy as np
import pytest

import oracle_api as orc
from pamg_pkg import pamg


def oracle_rows(n):
    C = 4 ** n
    out = np.zeros((C, 3), np.int64)
    r, p, o = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    f = orc.lib().orc_get_str_info
    for e in range(1, C + 1):
        f(n, e, ctypes.byref(r), ctypes.byref(p), ctypes.byref(o))
        out[e - 1] = (r.value, p.value, o.value)
    return out


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 6, 7])
def test_memory_order_and_paired_rows_equal_get_str_info(n):
    ref = oracle_rows(n)
    C, S, b = 4 ** n, 2 ** n, 2 ** (n + 1)
    got = pamg.numbering(0, n)
    assert np.array_equal(got[:, 0], ref[:, 0]) and np.array_equal(got[:, 1], ref[:, 1])
    assert np.array_equal(got[:, 2], b + 1 - 2 * ref[:, 0])                    # row length 2^(n+1) + 1 - 2 r
    # paired rows: slot t -> element; a bijection onto 1..C, rows r and S+1-r interleaved in blocks of 2^(n+1)
    flat = pamg.numbering(1, n)
    assert sorted(flat[:, 2].tolist()) == list(range(1, C + 1))
    assert np.array_equal(ref[flat[:, 2] - 1, 0], flat[:, 0]) and np.array_equal(ref[flat[:, 2] - 1, 1], flat[:, 1])
    blocks = flat[:, 0].reshape(-1, b)
    for p_, rows in enumerate(blocks):
        assert set(rows.tolist()) <= {p_ + 1, S - p_}


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 6])
def test_fine_children_equal_element_conversion(n):
    got = pamg.numbering(2, n)
    fin = np.zeros(4, np.int32)
    for e in range(1, 4 ** n + 1):
        orc.lib().orc_element_conversion(e, n, fin)
        assert np.array_equal(got[e - 1], fin), (n, e)
    # the fine children of all coarse children tile the fine level exactly once
    assert sorted(got.reshape(-1).tolist()) == list(range(1, 4 ** (n + 1) + 1))


@pytest.mark.parametrize("n", [4, 5, 8, 9, 11])
def test_walking_256_children_ahead_equals_the_closed_form(n):
    C = 4 ** n
    count = min(C, 1 << 20)
    for first in sorted({0, max(0, C // 2 - count // 2), C - count}):
        closed = pamg.numbering(0, n, first, count)
        walked = pamg.numbering(3, n, first, count)
        ok = np.arange(first, first + count) + 256 < C
        ahead = pamg.numbering(0, n, first + 256, count - 256) if count > 256 else closed[:0]
        m = min(len(ahead), int(ok.sum()))
        assert np.array_equal(walked[:m, :3], ahead[:m, :3])
        assert not walked[~ok].any()


@pytest.mark.parametrize("n", [8, 9, 10, 11, 12, 13])
def test_closed_form_is_exact_for_every_child_up_to_the_largest_split(n):
    C, S, b = 4 ** n, 2 ** n, 2 ** (n + 1)
    chunk = 1 << 22
    for first in range(0, C, chunk):
        count = min(chunk, C - first)
        got = pamg.numbering(0, n, first, count).astype(np.int64)
        k = np.arange(first, first + count, dtype=np.int64)
        m = got[:, 0] - 1                                        # row r starts at m (b - m), m = r - 1
        assert np.all((m >= 0) & (m < S))
        assert np.all(m * (b - m) <= k) and np.all((m + 1 >= S) | ((m + 1) * (b - m - 1) > k))
        assert np.array_equal(got[:, 1], k - m * (b - m) + 1)


def test_numbering_argument_errors():
    out = np.zeros((4, 4), np.int32)
    L = pamg.lib()
    assert L.pamg_numbering(0, 0, 0, 1, out) == pamg.ERR_ARG
    assert L.pamg_numbering(0, 14, 0, 1, out) == pamg.ERR_ARG
    assert L.pamg_numbering(4, 2, 0, 1, out) == pamg.ERR_ARG
    assert L.pamg_numbering(0, 1, 2, 3, out) == pamg.ERR_ARG      # beyond 4^s children

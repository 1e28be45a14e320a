"""The C-ABI library loads on a CPU-only box, exports every symbol include/avirb200.h
declares, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import avir_b200 as ab
import cases as cs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    txt = open(os.path.join(ROOT, "include", "avirb200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    names = re.findall(r"\b((?:avirb200|lancirb200)_[a-z0-9_]+)\s*\(", txt)
    return sorted(set(names))


def test_every_declared_symbol_is_exported():
    lib = ab.lib()
    names = declared_functions()
    assert len(names) >= 23
    for n in names:
        assert hasattr(lib, n), "libavirb200.so does not export " + n


def test_status_strings():
    lib = ab.lib()
    assert lib.avirb200_status_string(0) == b"ok"
    for code in range(-6, 0):
        assert lib.avirb200_status_string(code) not in (b"ok", b"unknown status")


def test_bad_arguments_are_rejected_before_touching_cuda():
    lib = ab.lib()
    out = C.c_void_p()
    assert lib.avirb200_plan_create(None, C.byref(out)) == -1
    assert lib.lancirb200_plan_create(None, C.byref(out)) == -1
    info = (C.c_int * 8)()
    assert lib.avirb200_shard_query_desc(None, 0, 1, info) == -1


@pytest.mark.skipif(ab.device_count() > 0, reason="checks the no-GPU behaviour")
def test_no_cpu_fallback_without_gpu():
    """Without a device the product must fail loudly, never compute on the host."""
    case = cs.SMALL_CASES[0]
    src = cs.make_input(case)
    with pytest.raises(ab.AvirB200Error) as e:
        cs.gpu_output(case, src)
    assert "no usable CUDA device" in str(e.value) or "CUDA" in str(e.value)
    r, _ = ab.CLancIR().resizeImage(np.zeros((8, 8, 4), np.uint8), 4, 4)
    assert r == 0  # upstream's error convention (lancir.h:382-383)


def test_descriptor_is_built_without_gpu():
    case = cs.SMALL_CASES[2]
    fp, sw, sh, nw, nh, ch, ti, to, rb, kw = case
    rs, v = cs.resizer_and_vars(case)
    h, dp, modes = rs.descriptor((sh, sw, ch), ti, nw, nh, to, 0.0, v)
    assert dp and modes[0] in (0, 1) and modes[1] in (0, 1)
    rs.free_descriptor(h)


def test_lancir_argument_errors_follow_upstream():
    # lancir.h:392-408: bad sizes / la < 2 -> 0
    r, _ = ab.CLancIR().resizeImage(np.zeros((8, 8, 4), np.uint8), 4, 4, ab.CLancIRParams(la=1.5))
    assert r == 0

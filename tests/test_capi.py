"""The C-ABI library loads and exports every symbol include/*.h declares; without a GPU every
compute entry point fails loudly (no CPU fallback).  No compute calls here."""
import ctypes as C
import glob
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def declared_functions():
    names = []
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        src = open(h).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names += re.findall(r"\b(lf_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def test_header_declares_functions():
    fns = declared_functions()
    assert "lf_ldd_build" in fns and "lf_router_route" in fns and len(fns) >= 20


def test_library_exports_every_declared_symbol():
    from lisflood_code_b200 import _capi
    L = C.CDLL(_capi.LIB_PATH)
    for name in declared_functions():
        assert hasattr(L, name), "symbol %s declared in include/ but not exported" % name
        assert name in _capi.SIGNATURES, "symbol %s has no ctypes signature" % name
    assert set(_capi.SIGNATURES) == set(declared_functions())
    assert _capi.lib().lf_version() >= 100


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from lisflood_code_b200 import _capi
    from lisflood_code_b200.hydrological_modules.kinematic_wave_parallel import kinematicWave
    rc = _capi.lib().lf_device_init(0)
    assert rc == _capi.LF_ERR_NO_DEVICE
    assert b"no CPU fallback" in _capi.lib().lf_last_error()
    with pytest.raises(_capi.LisfloodB200Error):
        kinematicWave(np.array([2.0, 5.0]), np.ones((2, 1), bool), np.ones(2), 0.6, 1000.0, 3600.0)


def test_product_never_imports_oracle():
    """The product package must not reference oracle/ (parity claims depend on it)."""
    pkg = os.path.join(ROOT, "lisflood_code_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("no oracle", ""), os.path.join(dirpath, f)


def test_host_modules_import():
    from lisflood_code_b200 import hotpath, synthetic  # noqa: F401
    from lisflood_code_b200.global_modules import add1, ldd_ops  # noqa: F401
    assert "W1a" in hotpath.THREE_ROWS and "LZ" not in hotpath.THREE_ROWS


def test_soil_columns_args_layout_matches_the_header(tmp_path):
    """ctypes mirror of struct lf_soil_columns_args: same size and field offsets as the C header (checked with gcc)."""
    import subprocess
    from lisflood_code_b200 import _capi
    names = [n for n, _ in _capi.SoilColumnsArgs._fields_]
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "lisflood_b200.h"\nint main(void){\n'
                   'printf("%zu\\n", sizeof(lf_soil_columns_args));\n'
                   + "".join('printf("%%zu\\n", offsetof(lf_soil_columns_args, %s));\n' % n for n in names)
                   + "return 0;}\n")
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    assert out[0] == C.sizeof(_capi.SoilColumnsArgs)
    assert out[1:] == [getattr(_capi.SoilColumnsArgs, n).offset for n in names]
    from lisflood_code_b200.hydrological_modules import soilloop
    assert len(soilloop._SOIL_ARG_ORDER) == 73


def test_soil_pf_args_layout_matches_the_header(tmp_path):
    """ctypes mirror of struct lf_soil_pf_args (suctionUnsaturatedSoilPF): same size and field offsets as the C header."""
    import subprocess
    from lisflood_code_b200 import _capi
    names = [n for n, _ in _capi.SoilPfArgs._fields_]
    src = tmp_path / "layout_pf.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "lisflood_b200.h"\nint main(void){\n'
                   'printf("%zu\\n", sizeof(lf_soil_pf_args));\n'
                   + "".join('printf("%%zu\\n", offsetof(lf_soil_pf_args, %s));\n' % n for n in names)
                   + "return 0;}\n")
    exe = tmp_path / "layout_pf"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    assert out[0] == C.sizeof(_capi.SoilPfArgs)
    assert out[1:] == [getattr(_capi.SoilPfArgs, n).offset for n in names]

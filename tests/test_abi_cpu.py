"""The C-ABI library on a GPU-less host: it loads, exports every symbol the header
declares, the ctypes mirror of the parameter struct matches the C layout, and the
CUDA path refuses to run (no silent fallback) without a device."""
import ctypes
import os
import re
import subprocess

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "dynmm_b200.h")


def _declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dynmm_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from dynmm_b200 import _lib
    lib = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/dynmm_b200.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes prototype"
    assert sorted(_lib.SIGNATURES) == names
    assert lib.dynmm_abi_version() == 1


def _check_layout(tmp_path, c_name, mirror, rename=None):
    rename = rename or {}
    src = tmp_path / f"{c_name}.c"
    fields = [f[0] for f in mirror._fields_]
    c_fields = [rename.get(f, f) for f in fields]
    prints = "\n".join(f'  printf("%zu\\n", offsetof({c_name}, {f}));' for f in c_fields)
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "dynmm_b200.h"\nint main(void) {\n'
                   f'  printf("%zu\\n", sizeof({c_name}));\n' + prints + "\n  return 0;\n}\n")
    exe = tmp_path / c_name
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = [int(x) for x in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    assert out[0] == ctypes.sizeof(mirror)
    for name, off in zip(fields, out[1:]):
        assert getattr(mirror, name).offset == off, name


def test_conv_params_struct_layout_matches_c(tmp_path):
    from dynmm_b200 import _lib
    _check_layout(tmp_path, "dynmm_conv_params", _lib.ConvParams, {"in_": "in"})


def test_pair_and_wgrad_struct_layouts_match_c(tmp_path):
    from dynmm_b200 import _lib
    _check_layout(tmp_path, "dynmm_conv_pair_params", _lib.ConvPairParams, {"in_": "in"})
    _check_layout(tmp_path, "dynmm_wgrad_params", _lib.WgradParams)


def test_bad_arguments_are_reported_not_crashing():
    from dynmm_b200 import _lib
    lib = _lib.load()
    p = _lib.ConvParams()          # all NULL
    rc = lib.dynmm_conv_igemm_fwd(ctypes.byref(p), None)
    assert rc == -1 and b"null" in lib.dynmm_last_error()
    assert lib.dynmm_global_gate_workspace(8, 4, 4) == -1
    assert lib.dynmm_global_gate_workspace(8, 120, 160) > 0
    q = _lib.ConvPairParams()      # all NULL
    assert lib.dynmm_conv_pair_fwd(ctypes.byref(q), None) == -1 and b"null" in lib.dynmm_last_error()
    assert lib.dynmm_conv_program_bytes(0) == -1 and lib.dynmm_conv_program_bytes(4) > 0
    assert lib.dynmm_stem_s2d_workspace(8, 4, 4) == -1 and lib.dynmm_stem_s2d_workspace(8, 480, 640) > 40_000_000


@pytest.mark.skipif(torch.cuda.is_available(), reason="needs a GPU-less host")
def test_cuda_path_fails_loudly_without_a_device():
    from dynmm_b200 import _lib
    from dynmm_b200.fusion import SkipGateESANet
    with pytest.raises(_lib.DynmmError):
        _lib.require_device()
    model = SkipGateESANet(height=64, width=64).eval()
    with pytest.raises(_lib.DynmmError):
        model.engine()

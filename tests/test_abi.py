"""The C-ABI shared library loads and exports every symbol include/bbbp_b200.h declares (no GPU needed)."""
import ctypes
import os
import re

import bbbp_b200
from bbbp_b200 import _lib


def test_header_symbols_all_exported():
    header = open(_lib.HEADER_PATH).read()
    declared = set(re.findall(r"\b(bbbp_[a-z0-9_]+)\s*\(", re.sub(r"/\*.*?\*/", " ", header, flags=re.S)))
    assert declared == set(_lib.PROTOTYPES), declared ^ set(_lib.PROTOTYPES)
    assert len(declared) >= 34
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert getattr(raw, name) is not None


def test_abi_version_and_no_torch_in_signatures():
    assert bbbp_b200.ABI_VERSION == 4
    header = open(_lib.HEADER_PATH).read()
    code = re.sub(r"/\*.*?\*/", " ", header, flags=re.S)
    assert "torch" not in code.lower() and "Tensor" not in code and "at::" not in code
    assert 'extern "C"' in code


def test_library_has_no_torch_dependency():
    import subprocess
    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "torch" not in out and "c10" not in out


def test_argument_validation_reports_through_last_error():
    # argument checks run before any CUDA call, so they are testable without a device
    rc = _lib.lib.bbbp_gemm_f32(0, 1, 4, 4, 4, None, 4, None, 4, None, 4, None, 0, 0, 1, None, 0, None)
    assert rc == -1
    assert "null operand" in bbbp_b200.last_error()
    rc = _lib.lib.bbbp_conv3x3_f32(1, 1, None, 1, None, 1, 3, 30, 128, 128, 1, None)
    assert rc == -1 and "multiple of 32" in bbbp_b200.last_error()
    rc = _lib.lib.bbbp_gemm_bf16(4, 4, 7, 16, 7, 16, 8, None, None, 0, 16, 4, None, 0, 0, 1, None, 0, None)
    assert rc == -1 and "multiples of 8" in bbbp_b200.last_error()
    rc = _lib.lib.bbbp_batchnorm_fwd_f32(1, 1, 1, 1, 1, 1, 1, 1, 1, 8, 1, 0.1, 1e-5, None)
    assert rc == -1 and "more than 1 value per channel" in bbbp_b200.last_error()


def test_missing_library_fails_loudly(tmp_path, monkeypatch):
    import importlib.util
    spec = importlib.util.spec_from_file_location("lib_probe", os.path.join(_lib.PKG_DIR, "_lib.py"))
    mod = importlib.util.module_from_spec(spec)
    monkeypatch.setattr(os.path, "exists", lambda p, _orig=os.path.exists: False if p.endswith("libbbbp_b200.so") else _orig(p))
    try:
        spec.loader.exec_module(mod)
    except ImportError as e:
        assert "no CPU or PyTorch fallback" in str(e)
    else:
        raise AssertionError("import succeeded without the library")


def test_cuda_sources_are_blackwell_native():
    import subprocess
    sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass      # tcgen05.mma
    assert "UTMALDG" in sass      # TMA tensor loads
    assert "LDTM" in sass         # tcgen05.ld
    assert "HGMMA" not in sass


def test_whole_model_entry_points_describe_the_reference_state_dict():
    """bbbp_model_param_name / _numel enumerate MixedInputModel.state_dict() (20250113.py:69-107) key for key."""
    import ctypes
    from oracle import nets
    from bbbp_b200 import c_host
    for fp_dim in (167, 2048, 64):
        ref = nets.build("tcnn", fp_dim, 128)
        desc = c_host.make_desc(fp_dim, "strict")
        names = c_host.param_names(desc)
        state = ref.state_dict()
        assert names == list(state.keys())
        for i, n in enumerate(names):
            assert _lib.lib.bbbp_model_param_numel(ctypes.byref(desc), i) == state[n].numel(), n
        assert _lib.lib.bbbp_model_prepared_bytes(ctypes.byref(desc)) > 128 * 65536 * 2
        desc = c_host.make_desc(fp_dim, "bf16", 4, 256)
        assert _lib.lib.bbbp_workspace_bytes(ctypes.byref(desc)) > 1024 * (64 * 64 * 32 + 65536) * 2


def test_whole_model_entry_points_reject_what_is_not_built():
    import ctypes
    from bbbp_b200 import c_host
    lib = _lib.lib
    desc = c_host.make_desc(167, "strict", 1, 32)
    desc.precision = 0                                   # BBBP_PREC_FP32: validation mode, per-kernel entry points only
    assert lib.bbbp_workspace_bytes(ctypes.byref(desc)) == 0 and "BF16, F16 and STRICT" in bbbp_b200.last_error()
    desc = c_host.make_desc(167, "strict", 1, 32)
    desc.abi_version = 1
    assert lib.bbbp_model_param_count(ctypes.byref(desc)) == -1 and "abi_version" in bbbp_b200.last_error()
    desc = c_host.make_desc(200, "bf16", 1, 32)          # 25 heads of dimension 8 are built ...
    assert lib.bbbp_model_param_count(ctypes.byref(desc)) == 109
    desc = c_host.make_desc(201, "bf16", 1, 32)          # ... 3 heads of dimension 67 are not
    assert lib.bbbp_model_param_count(ctypes.byref(desc)) == -4
    desc = c_host.make_desc(167, "bf16", 1, 32)
    assert lib.bbbp_fwd(ctypes.byref(desc), None, None, None, None, None, None, 0, None) == -1
    assert lib.bbbp_comm_gather_scores(None, None, None, 0, None) == -1


def test_header_is_plain_c_and_the_c_host_example_compiles():
    """include/bbbp_b200.h must be consumable by a C compiler (the drop-in boundary is a C ABI): gcc -std=c99 -fsyntax-only on
    the header alone, and on tools/c_host_example.c (a complete C host of bbbp_fwd) when the CUDA runtime headers are present."""
    import shutil
    import subprocess
    import pytest
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    inc = os.path.join(root, "include")
    probe = '#include "bbbp_b200.h"\nint main(void) { bbbp_model_desc d = {BBBP_ABI_VERSION, BBBP_MODEL_TCNN_20250113, 167, BBBP_PREC_STRICT, 1, 32, 0}; return d.seq == 32 ? 0 : 1; }\n'
    r = subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-fsyntax-only", "-I" + inc, "-x", "c", "-"], input=probe,
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    if os.path.exists(os.path.join(cuda_inc, "cuda_runtime_api.h")):
        r = subprocess.run(["gcc", "-std=c99", "-Wall", "-fsyntax-only", "-I" + inc, "-I" + cuda_inc,
                            os.path.join(root, "tools", "c_host_example.c")], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr

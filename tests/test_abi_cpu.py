"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every symbol that
include/mqb200.h declares; no compute call is made (there is no GPU in the build container)."""
import ctypes, os
import pytest
from mobilequant_b200 import _lib, build as mqbuild


@pytest.fixture(scope="module")
def lib():
    mqbuild.build()
    return _lib.load()


def test_header_symbols_exported(lib):
    syms = _lib.declared_symbols()
    assert len(syms) >= 12
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, f"declared in include/mqb200.h but not exported: {missing}"


def test_version_and_error_strings(lib):
    assert lib.mq_version() == 100
    assert lib.mq_get_error_description(0) == b"no error"
    assert lib.mq_get_error_description(2) == b"invalid argument"
    assert lib.mq_get_error_description(99) is None


def test_invalid_context_is_rejected(lib):
    # capp/src/libllmod.cpp:50-65 semantics: a bad handle never dereferences, it returns INVALID_CONTEXT
    lib.mq_release.argtypes = [ctypes.c_void_p]
    assert lib.mq_release(None) == 1
    buf = ctypes.create_string_buffer(256)
    assert lib.mq_release(ctypes.cast(buf, ctypes.c_void_p)) == 1


def test_setup_without_gpu_fails_loudly(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = ctypes.c_void_p()
    rc = lib.mq_setup(ctypes.byref(h), 0)
    assert rc == 4 and not h.value            # MQ_RUNTIME_ERROR, no context created
    assert b"CUDA" in lib.mq_get_last_error_extra_info(rc, None)


def test_product_has_no_cpu_fallback():
    import torch
    from mobilequant_b200 import kernels
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    x = torch.randn(16)
    with pytest.raises(_lib.MQError):
        kernels.fq_fwd(x, torch.tensor(0.1), torch.tensor(0.0), 0, 255)


def test_product_never_imports_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    bad = []
    for dp, _, fs in os.walk(os.path.join(root, "mobilequant_b200")):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                if "import oracle" in txt or "from oracle" in txt or "/root/reference" in txt:
                    bad.append(os.path.join(dp, f))
    assert not bad, bad


def _header_prototypes():
    """{function name: number of parameters} parsed from include/mqb200.h (comments stripped)."""
    import re
    txt = open(_lib.HEADER_PATH).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(?:int|const char\*)\s+(mq_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", txt, flags=re.S):
        params = m.group(2).strip()
        protos[m.group(1)] = 0 if params in ("", "void") else len(params.split(","))
    return protos


def test_ctypes_bindings_match_header(lib):
    """Every binding in kernels.py passes exactly as many arguments as the header declares (a dropped or extra argument in
    a ctypes prototype corrupts the call without any diagnostic)."""
    from mobilequant_b200 import kernels
    kernels._protos()
    protos = _header_prototypes()
    assert len(protos) >= 20
    checked = 0
    for name, n in protos.items():
        fn = getattr(lib, name)
        if fn.argtypes is not None:
            assert len(fn.argtypes) == n, f"{name}: header declares {n} parameters, ctypes binding passes {len(fn.argtypes)}"
            checked += 1
    assert checked >= 15


def test_sim_export_key_set_matches_reference_golden():
    """device/convert_sim.py: the parameter names let through for SimModel equal the reference's state-dict keys."""
    import torch
    from mobilequant_b200.device.convert_sim import sim_state_keys
    g = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sim_export.pt"), weights_only=False)
    for tag in ("llama_w8_slinear", "llama_w4", "gemma_w8_slinear"):
        assert sim_state_keys(2, g[tag]["impl_sym_pch_as_slinear"]) == set(g[tag]["out_states"].keys())

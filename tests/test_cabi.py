"""CPU: the C-ABI library loads and exports exactly what include/tdnet_b200.h declares."""
import ctypes
import os
import re

import pytest

from tdnet_b200 import _cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "tdnet_b200.h")


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tdn_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.isfile(_cabi.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return _cabi.load()


def test_header_and_binding_agree():
    assert header_functions() == sorted(_cabi.SIGNATURES)


def test_every_declared_symbol_is_exported(lib):
    raw = ctypes.CDLL(_cabi.LIB_PATH)
    for name in header_functions():
        assert hasattr(raw, name), f"{name} declared in the header but not exported"


def test_info_calls_without_gpu(lib):
    assert lib.tdn_abi_version() == 2
    assert lib.tdn_strerror(0) == b"ok"
    assert b"sm_100" in lib.tdn_strerror(-4)
    assert lib.tdn_psp_pool_workspace_bytes(1, 128, 512) == 128 * 12 * 512 * 4
    # argument validation happens before any CUDA call, so it is checkable on a CPU-only box
    assert lib.tdn_conv2d(None, None) == -1
    assert b"null descriptor" in lib.tdn_last_error()


def test_struct_layout_matches_header(lib):
    # tdn_tensor: 2 pointers, 5 int32 (+pad), 3 int64
    assert ctypes.sizeof(_cabi.Tensor) == 8 + 8 + 4 * 5 + 4 + 8 * 3
    assert _cabi.Conv2dDesc.weight.offset == 3 * ctypes.sizeof(_cabi.Tensor)


def test_plain_c_consumer_links_and_runs(lib, tmp_path):
    """examples/abi_smoke.c: a C99 program linked against the library (no Python, no torch in the boundary); the GPU-less
    build checks version / error strings / workspace sizing and the argument validation done before any CUDA call."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    exe = tmp_path / "abi_smoke"
    libdir = os.path.dirname(_cabi.LIB_PATH)
    subprocess.check_call([gcc, "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "abi_smoke.c"), "-L", libdir, "-ltdnet_b200",
                           f"-Wl,-rpath,{libdir}", "-o", str(exe)])
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0 and "abi_smoke: ok" in out.stdout, out.stderr

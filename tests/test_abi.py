"""CPU: the C-ABI library loads, exports every symbol include/million_b200.h declares, its struct layout matches the
ctypes mirror, argument errors come back as status codes (no compute on a GPU-less box), and the product refuses to run
without CUDA (no CPU fallback)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "million_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(million_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge.build()
    from million_b200 import _lib
    return _lib


def test_every_declared_symbol_is_exported_and_bound(lib):
    syms = declared_symbols()
    assert len(syms) >= 12
    handle = lib.lib()
    for s in syms:
        assert s in lib.SIGNATURES, f"{s} has no ctypes signature"
        assert getattr(handle, s) is not None
    out = subprocess.check_output(["nm", "-D", "--defined-only", lib.LIB_PATH], text=True)
    exported = set(re.findall(r" T (million_\w+)", out))
    assert set(syms) <= exported


def test_struct_layout_matches_header(lib, tmp_path):
    c = tmp_path / "sz.c"
    c.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "%s"\nint main(){printf("%%zu %%zu %%zu %%zu %%zu", sizeof(million_attn_params),'
                 'offsetof(million_attn_params,k_head_stride),offsetof(million_attn_params,v_page_ids),offsetof(million_attn_params,workspace_bytes),'
                 'offsetof(million_attn_params,partial));}' % HEADER)
    exe = tmp_path / "sz"
    subprocess.check_call(["/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc", str(c), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)], text=True).split()]
    P = lib.AttnParams
    assert got == [ctypes.sizeof(P), P.k_head_stride.offset, P.v_page_ids.offset, P.workspace_bytes.offset, P.partial.offset]


def test_abi_version_and_pure_functions(lib):
    h = lib.lib()
    assert h.million_abi_version() == lib.ABI_VERSION
    assert h.million_pq_decode_attn_workspace_bytes(1, 32, 8, 128, 18) == 256 + 32 * 20 * 132 * 4   # rows of d + 2 floats padded to 16 bytes


def test_argument_errors_are_status_codes(lib):
    h = lib.lib()
    p = lib.AttnParams()
    assert h.million_pq_decode_attn(ctypes.byref(p), None) == lib.MILLION_ERR_INVALID      # struct_size 0
    assert b"struct_size" in h.million_last_error()
    assert h.million_pq_decode_attn(None, None) == lib.MILLION_ERR_INVALID
    one = ctypes.c_void_p(16)
    assert h.million_pq_encode(one, 0, 0, one, None, one, 1, 0, 0, 0, 0, 1, 1, 128, 60, 256, 0, None) == lib.MILLION_ERR_INVALID  # d % M
    assert h.million_pq_encode(one, 0, 0, one, None, one, 1, 0, 0, 0, 0, 1, 1, 128, 64, 512, 0, None) == lib.MILLION_ERR_INVALID  # C > 256 in u8
    assert h.million_pq_encode(None, 0, 0, None, None, None, 1, 0, 0, 0, 0, 0, 0, 128, 64, 256, 0, None) == lib.MILLION_OK         # empty
    with pytest.raises(lib.MillionError):
        lib.check(lib.MILLION_ERR_INVALID)


def test_no_cpu_fallback():
    import torch
    from million_b200 import ops, pq_utils
    X = torch.zeros(1, 1, 4, 128, dtype=torch.float16)
    cent = torch.zeros(64, 256, 2)
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.pq_encode(X, cent)
    with pytest.raises(RuntimeError, match="CUDA"):
        pq_utils.sa_decode_4d(torch.zeros(1, 1, 4, 64, dtype=torch.uint8), cent)
    if not torch.cuda.is_available():
        pq_utils.Singleton.clear_instance()
        with pytest.raises(Exception):
            pq_utils.DynamicPQCache(bs=1, nh=1, num_key_value_heads=1, M=64, layer_num=1, scalar_t=torch.float16)
        pq_utils.Singleton.clear_instance()


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "million_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "oracle/" not in src, f

"""CPU: the C-ABI library builds, loads, and exports exactly the symbols include/vitb200.h declares
(no compute calls here: there is no GPU in the authoring container)."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "vitb200.h")


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as ge
    ge.build()
    import vit_cifar_b200 as vb
    return vb


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vitb_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_and_library_exports_same_symbols(built):
    from vit_cifar_b200 import _lib
    declared = header_functions()
    assert declared, "no functions parsed from the header"
    assert sorted(_lib.SIGNATURES) == declared  # the ctypes table mirrors the header
    nm = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = sorted(n for n in set(re.findall(r"\bT (vitb_[a-z0-9_]+)\b", nm)) if not n.startswith("vitb_debug_"))  # tools-only hooks
    assert exported == declared
    lib = built.load_library()
    for name in declared:
        assert hasattr(lib, name)
    assert lib.vitb_version() == _lib.ABI_VERSION
    assert lib.vitb_last_error() == b""


def test_workspace_size_queries_are_pure_host_functions(built):
    lib = built.load_library()
    assert lib.vitb_layernorm_bwd_ws_bytes(8320, 384) == 3 * 296 * 384 * 4
    assert lib.vitb_colsum_ws_bytes(8320, 384) > 0
    assert lib.vitb_colsum_ws_bytes(10, 100) == 0  # cols must be a multiple of 128 for the vectorised kernel
    assert lib.vitb_gemm_wgrad_ws_bytes(66560, 384, 384, 1) >= 16 * 384 * 384 * 4
    assert lib.vitb_patch_embed_bwd_ws_bytes(128, 32, 8, 384, 1) > 0


def test_library_is_sm100a_with_tcgen05_and_tma(built):
    """SASS evidence that the GEMM is a Blackwell-native kernel (UTCHMMA = tcgen05.mma, UTMALDG = TMA, LDTM = tcgen05.ld)."""
    from vit_cifar_b200 import _lib
    r = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in r.stdout
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM", "HMMA", "LDSM"):
        assert mnemonic in r.stdout, mnemonic


def test_no_cpu_fallback(built):
    import torch
    vb = built
    m = vb.ViT(3, 10, img_size=32, patch=8, num_layers=1, hidden=128, mlp_hidden=128, head=4)
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(vb.VitbError):
        m(torch.zeros(2, 3, 32, 32))
    with pytest.raises(vb.VitbError):
        vb.LabelSmoothingCrossEntropyLoss(10, 0.1)(torch.zeros(2, 10), torch.zeros(2, dtype=torch.long))


def test_oracle_is_never_imported_by_the_product_or_the_tools():
    """oracle/ is test infrastructure: the package and tools/ must not reference it (bench.py's CPU-baseline legs and
    __graft_entry__.smoke() are the only other users)."""
    import glob
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    offenders = []
    for path in glob.glob(os.path.join(root, "vit-cifar_b200", "**", "*.py"), recursive=True) + glob.glob(os.path.join(root, "tools", "*.py")):
        if re.search(r"^\s*(import|from)\s+oracle\b", open(path).read(), re.M):
            offenders.append(os.path.relpath(path, root))
    assert offenders == []


def test_round2_host_side_entry_points(built):
    """Pure host functions added in round 2: workspace queries of the one-pass backward GEMM, argument checks of the deferred-
    reduction window and of the data-parallel flag-wait bound (no device work)."""
    lib = built.load_library()
    bf16, f32 = 1, 0
    assert lib.vitb_gemm_bwd_fused_ws_bytes(66560, 384, 384, bf16) >= 49 * 384 * 384 * 4   # one fp32 dW slice per CTA of a column block
    assert lib.vitb_gemm_bwd_fused_ws_bytes(1000, 128, 256, bf16) > 0
    for M, N, K, dt in ((66560, 1152, 384, bf16), (66560, 384, 100, bf16), (66560, 384, 384, f32), (66560, 768, 3072, bf16)):
        assert lib.vitb_gemm_bwd_fused_ws_bytes(M, N, K, dt) == 0   # caller falls back to dgrad + wgrad
    assert lib.vitb_defer_begin(None, 0) < 0 and b"arena" in lib.vitb_last_error()
    assert lib.vitb_defer_flush(None) < 0          # no window open
    assert lib.vitb_defer_used() == 0
    assert lib.vitb_dp_set_timeout(0.0) < 0
    assert lib.vitb_dp_set_timeout(600.0) == 0
    assert lib.vitb_patch_embed_fwd_ws_bytes(128, 32, 8, 384, bf16) >= 65 * 384 * 2   # bf16 copy of pos_emb for the GEMM epilogue


def test_dropout_descriptor_layout_matches_the_header(tmp_path):
    """vitb_dropout_t is passed by pointer from ctypes: its field offsets in _lib.DropoutDesc must be the C compiler's."""
    import ctypes
    import subprocess
    import vit_cifar_b200  # noqa: F401
    from vit_cifar_b200 import _lib
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "vitb200.h"\n'
                   'int main(void) { printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(vitb_dropout_t), offsetof(vitb_dropout_t, p),\n'
                   '  offsetof(vitb_dropout_t, seed), offsetof(vitb_dropout_t, site), offsetof(vitb_dropout_t, step), offsetof(vitb_dropout_t, step_dev));\n'
                   '  return 0; }\n')
    exe = tmp_path / "layout"
    inc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include")
    subprocess.run(["gcc", "-I", inc, str(src), "-o", str(exe)], check=True)
    got = [int(t) for t in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    D = _lib.DropoutDesc
    assert got == [ctypes.sizeof(D), D.p.offset, D.seed.offset, D.site.offset, D.step.offset, D.step_dev.offset]

"""CPU: the C-ABI library builds in-tree, loads, and exports every function include/mf_b200.h declares (no compute calls: there
is no GPU here); the ctypes shim binds each of them; nothing under the product package imports the oracle."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "mf_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mf_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from mere_fusion_b200 import build
    lib_path = build.build()
    assert os.path.exists(lib_path)
    L = ctypes.CDLL(lib_path)
    names = _declared()
    assert len(names) >= 25 and "mf_ernerf_render" in names and "mf_musetalk_forward" in names and "mf_whisper_features" in names
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, f"declared in include/mf_b200.h but not exported: {missing}"
    L.mf_version.restype = ctypes.c_int
    assert L.mf_version() >= 1
    # without a GPU mf_create must fail cleanly (no CPU fallback), never crash
    ctx = ctypes.c_void_p()
    L.mf_create.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]
    import torch
    if not torch.cuda.is_available():
        assert L.mf_create(0, ctypes.byref(ctx)) != 0 and not ctx.value


def test_ctypes_shim_binds_the_hot_path_entry_points():
    from mere_fusion_b200 import _lib
    L = _lib.lib()
    for n in ("mf_ernerf_load", "mf_ernerf_render", "mf_wav2lip_load", "mf_wav2lip_forward", "mf_musetalk_forward", "mf_whisper_features", "mf_wav2vec2_logits",
              "mf_wav2lip_mel_chunks", "mf_paste_resize_u8", "mf_paste_blend_u8", "mf_march_rays", "mf_composite_rays_triplane",
              "mf_grid_encode_forward"):
        assert getattr(L, n).argtypes is not None, n


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "mere_fusion_b200")
    bad = []
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(d, f), errors="ignore").read()
                if re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M) or "oracle/" in txt and f.endswith(".py") and "import" in txt.split("oracle/")[0][-40:]:
                    bad.append(os.path.join(d, f))
    assert not bad, bad


def test_binding_argument_counts_match_the_header():
    """every bound entry point takes as many arguments as include/mf_b200.h declares (a drifted ctypes prototype corrupts the
    call silently)"""
    from mere_fusion_b200 import _lib
    src = open(os.path.join(ROOT, "include", "mf_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    decl = {}
    for m in re.finditer(r"\b(?:int|void|const char \*)\s*\*?\s*(mf_[a-z0-9_]+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        decl[m.group(1)] = 0 if args in ("", "void") else len(args.split(","))
    assert len(decl) >= 30 and "mf_ernerf_render_batch" in decl and "mf_wav2vec2_logits_batch" in decl
    L = _lib.lib()
    bad = [(n, k, len(getattr(L, n).argtypes)) for n, k in decl.items()
           if getattr(L, n).argtypes is not None and len(getattr(L, n).argtypes) != k]
    assert not bad, bad
    unbound = [n for n in decl if getattr(L, n).argtypes is None and decl[n] > 0]
    assert not unbound, f"declared but without a ctypes prototype: {unbound}"

"""CPU-only: the C-ABI library loads, exports every symbol the headers declare, and refuses to run
without a GPU (no CPU fallback)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    for hdr in ("b2pt.h", "b2pt_host.h"):
        text = open(os.path.join(ROOT, "include", hdr)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names.update(re.findall(r"\b(b2pt_[a-z0-9_]+)\s*\(", text))
    return sorted(names)


def test_headers_declare_something():
    syms = declared_symbols()
    assert "b2pt_create" in syms and "b2pt_render" in syms and "b2pt_scene_load_obj" in syms
    assert len(syms) >= 24


def test_library_exports_every_declared_symbol(built):
    import path_tracer_ai_b200 as pt
    lib = ctypes.CDLL(pt.LIB_PATH)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, f"declared in include/*.h but not exported: {missing}"


def test_python_binding_lists_match_headers(built):
    from path_tracer_ai_b200 import _capi, renderer
    assert sorted(_capi.EXPORTS + renderer.HOST_EXPORTS) == declared_symbols()


def test_header_compiles_as_c(built, tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "b2pt.h"\n#include "b2pt_host.h"\nint main(void){ b2pt_settings s; s.width = 1; return s.width - 1; }\n')
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    subprocess.run([cc, "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", str(src), "-o", str(tmp_path / "t.o")], check=True)


def test_struct_sizes_match_header(built, tmp_path):
    """ctypes mirrors must have the C layout."""
    from path_tracer_ai_b200 import _capi
    src = tmp_path / "s.c"
    src.write_text('#include <stdio.h>\n#include "b2pt.h"\nint main(void){ printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(b2pt_material), sizeof(b2pt_light), '
                   'sizeof(b2pt_camera), sizeof(b2pt_settings), sizeof(b2pt_partition), sizeof(b2pt_config), sizeof(b2pt_stats)); return 0; }\n')
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    exe = tmp_path / "s"
    subprocess.run([cc, "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    sizes = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    mine = [ctypes.sizeof(c) for c in (_capi.Material, _capi.Light, _capi.Camera, _capi.Settings, _capi.Partition, _capi.Config, _capi.Stats)]
    assert sizes == mine


def test_no_gpu_means_loud_failure_not_fallback(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import path_tracer_ai_b200 as pt
    with pytest.raises(pt.B2ptError) as e:
        pt.Engine()
    assert "no CPU fallback" in str(e.value)
    r = pt.B200Renderer()
    with pytest.raises(pt.B2ptError):
        r.render(pt.Camera())   # render before initialize (optix_renderer.cu:421-423)


def test_product_never_imports_oracle():
    """The product path must not import, link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "path_tracer_ai_b200")
    bad = []
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                if re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M) or re.search(r'#include\s+[<"].*oracle/', text) \
                        or "libpt_oracle" in text or "libref_oracle" in text:
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad


def test_sass_keeps_the_parity_arithmetic_scalar(built):
    """ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 (one rounding where the reference has two), so the
    kernels must not use Blackwell's packed fp32 instructions at all: no FMUL2 / FADD2 / FFMA2 in libb2pt.so, and the
    traversal kernels do their arithmetic with scalar FMUL / FADD (tools/sass_mix.py, profiles/r02_sass_mix.txt)."""
    import shutil
    import path_tracer_ai_b200 as pt
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", pt.LIB_PATH], capture_output=True, text=True).stdout
    assert "Function :" in sass and "k_extend_rtc" in sass and "k_shadow_rtc" in sass
    for packed in ("FMUL2", "FADD2", "FFMA2"):
        assert f" {packed} " not in sass, packed
    assert sass.count(" FMUL ") > 1000 and sass.count(" FADD ") > 1000

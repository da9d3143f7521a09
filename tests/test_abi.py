"""The C-ABI library loads without a GPU, exports every symbol include/rt795.h declares, its structs match the ctypes
mirror, and the product path fails loudly (no CPU fallback) when CUDA is unusable.  CPU only, no compute calls."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = os.path.join(ROOT, "include", "rt795.h")


def declared_functions():
    src = open(HDR).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rt_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(rt):
    lib = rt.load_core()
    names = declared_functions()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f"librt795.so does not export {n}"
    assert sorted(rt.binding.CORE_SYMBOLS) == names
    assert lib.rt_abi_version() == 2


def test_struct_layouts_match_header(rt, tmp_path):
    structs = ["RtMaterial", "RtTexture", "RtImage", "RtMesh", "RtSphere", "RtObject", "RtLight", "RtSceneDesc", "RtCameraDesc", "RtRenderOpts", "RtHit", "RtStats"]
    prog = '#include <stdio.h>\n#include "rt795.h"\nint main(void){' + "".join(f'printf("{s} %zu\\n", sizeof({s}));' for s in structs) + "return 0;}"
    c = tmp_path / "sz.c"
    c.write_text(prog)
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(c), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    sizes = dict(line.split() for line in out.strip().splitlines())
    for s in structs:
        assert C.sizeof(getattr(rt.binding, s)) == int(sizes[s]), s


def test_no_cpu_fallback_without_cuda(rt, scene_dir):
    """On a box without a usable GPU the core must refuse to render; on a GPU box this test is a no-op."""
    lib = rt.load_core()
    if lib.rt_device_count() > 0:
        pytest.skip("CUDA device present")
    xml = rt.scenes.scene_c1(os.path.join(scene_dir, "c1abi"), res=16)
    hs = rt.HostScene(xml)
    with pytest.raises(rt.RtError) as e:
        rt.Scene(hs.desc, 0)
    assert e.value.code == rt.RT_ERR_CUDA
    assert b"CUDA" in lib.rt_last_error()
    hs.close()


def test_product_sources_never_reference_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py may touch oracle/: the shipped library must not."""
    pkg = os.path.join(ROOT, "raytracer-795_b200")
    for dp, _, files in os.walk(pkg):
        if os.path.basename(dp) in ("build", "lib", "bin", "__pycache__"):
            continue
        for f in files:
            if f.endswith(".py"):                                       # the ctypes driver of the product knows nothing about the oracle either
                txt = open(os.path.join(dp, f), errors="replace").read()
                assert "liboracle" not in txt and "load_oracle" not in txt and "oracle_render" not in txt, f
            if f.endswith((".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dp, f), errors="replace").read()
                code = re.sub(r"//[^\n]*", "", txt)                      # comments may cite the oracle; code may not use it
                assert "liboracle" not in code and "oracle_" not in code and "dlopen" not in code, f
                for inc in re.findall(r'#include\s*[<"]([^>"]+)[>"]', code):
                    assert "oracle" not in inc and "restate" not in inc, (f, inc)
    out = subprocess.run(["nm", "-D", "--defined-only", os.path.join(pkg, "lib", "librt795.so")], capture_output=True, text=True).stdout
    assert "oracle_" not in out

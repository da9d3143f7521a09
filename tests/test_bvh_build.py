"""Host SAH BVH builder + 4-wide collapse (H1): structural invariants, degenerate inputs, thread-count
independence.  Pure C++ test compiled here with g++; no GPU."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_bvh_builder_invariants(tmp_path):
    csrc = os.path.join(ROOT, "raytracer-795_b200", "csrc")
    exe = str(tmp_path / "bvh_build_test")
    subprocess.run(["g++", "-std=c++17", "-O2", "-I", csrc, os.path.join(ROOT, "tests", "cpp", "bvh_build_test.cpp"), os.path.join(csrc, "bvh_build.cpp"),
                    "-o", exe, "-lpthread"], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:]
    assert "bvh_build_test: ok" in r.stdout

#!/usr/bin/env python
"""Generates the committed golden fixtures in tests/golden/ from the UNMODIFIED reference.

Run HERE (in the container that has /root/reference): `python tests/golden/make_golden.py`.  It builds oracle/_ref
(oracle/Makefile compiles /root/reference/src in place against the stand-in headers), writes small deterministic
scenes with raytracer-795_b200/scenes.py, runs oracle/_ref/ref_raytracer (float framebuffer) and oracle/_ref/ref_hits
(per-pixel primary hit records) on them and stores the outputs as compressed .npz.  The scenes are regenerated from
the same seeded generators at test time, so only the reference OUTPUTS are committed.
Deterministic scenes only (1 spp, no area / environment light, no glossy): the reference's stochastic paths use
unseeded mt19937 streams and cannot be golden.
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_api as orc  # noqa: E402

CASES = {
    # name: (generator, kwargs)
    "c1_96": ("scene_c1", dict(res=96)),
    "c3det_96x54": ("scene_c3", dict(res=(96, 54), spp=1, n=24, deterministic=True)),
    "c4det_128x72": ("scene_c4", dict(res=(128, 72), spp=1, n=20, deterministic=True)),
}


def main():
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "-j8", "ref"], check=True, stdout=subprocess.DEVNULL)
    rt = ge.load_package()
    ref = os.path.join(ROOT, "oracle", "_ref")
    for name, (gen, kw) in CASES.items():
        with tempfile.TemporaryDirectory() as td:
            xml = getattr(rt.scenes, gen)(os.path.join(td, "s"), **kw)
            subprocess.run([os.path.join(ref, "ref_raytracer"), xml], cwd=td, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            subprocess.run([os.path.join(ref, "ref_hits"), xml, os.path.join(td, "hits.bin")], cwd=td, check=True, stdout=subprocess.DEVNULL)
            f32 = [f for f in os.listdir(td) if f.endswith(".f32")][0]
            img = rt.read_f32(os.path.join(td, f32))
            hits = orc.read_ref_hits(os.path.join(td, "hits.bin"))
        np.savez_compressed(os.path.join(HERE, name + ".npz"), radiance=img, hit_full=hits["full"].astype(np.int8), hit_label=hits["label"],
                            hit_t=hits["t"], hit_point=hits["point"], hit_normal=hits["normal"])
        print(name, img.shape, "hit fraction", hits["full"].mean(), os.path.getsize(os.path.join(HERE, name + ".npz")), "bytes")


if __name__ == "__main__":
    main()

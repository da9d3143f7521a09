"""Pins the oracle (oracle/restate.cpp) against the reference: committed golden outputs of the UNMODIFIED reference
(tests/golden/*.npz, produced by tests/golden/make_golden.py) and, where oracle/_ref is available, live runs.
CPU only."""
import os
import subprocess

import numpy as np
import pytest

import oracle_api as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")

CASES = {
    "c1_96": ("scene_c1", dict(res=96)),
    "c3det_96x54": ("scene_c3", dict(res=(96, 54), spp=1, n=24, deterministic=True)),
    "c4det_128x72": ("scene_c4", dict(res=(128, 72), spp=1, n=20, deterministic=True)),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_golden_radiance(rt, scene_dir, name):
    gen, kw = CASES[name]
    xml = getattr(rt.scenes, gen)(os.path.join(scene_dir, name), **kw)
    hs = rt.HostScene(xml)
    img = orc.oracle_render(hs.desc, hs.camera(0))
    gold = np.load(os.path.join(GOLD, name + ".npz"))["radiance"]
    assert img.shape == gold.shape
    r = np.abs(img - gold) / np.maximum(np.abs(gold), 1.0)
    # the restatement repeats the reference's float operations in the same order: expect bit-identical pixels almost
    # everywhere (libm / contraction differences at most)
    assert (img == gold).mean() >= 0.99, (img == gold).mean()
    assert r.max() <= 1e-4, r.max()
    hs.close()


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_golden_primary_hits(rt, scene_dir, name):
    gen, kw = CASES[name]
    xml = getattr(rt.scenes, gen)(os.path.join(scene_dir, name + "_h"), **kw)
    hs = rt.HostScene(xml)
    d = hs.desc.contents
    hits = orc.oracle_trace_primary(hs.desc, hs.camera(0))
    g = np.load(os.path.join(GOLD, name + ".npz"))
    assert (hits["full"] == g["hit_full"]).all()
    full = hits["full"] == 1
    # expected label per hit from (object, primitive), mirroring the harness numbering
    base, bases, n_base = 1, [], 0
    seen_mesh = {}
    for i in range(d.num_objects):
        o = d.objects[i]
        is_instance = o.kind == 1 and o.geom in seen_mesh
        if is_instance:
            bases.append(None)
            continue
        if o.kind == 1:
            seen_mesh[o.geom] = i
        bases.append(base)
        base += 1 if o.kind == 0 else d.meshes[o.geom].tri_count
        n_base += 1
    exp = np.zeros(hits.shape, dtype=np.int64)
    for i in range(d.num_objects):
        m = full & (hits["object"] == i)
        if bases[i] is None:
            exp[m] = -((i - n_base) + 1)
        else:
            exp[m] = bases[i] + hits["primitive"][m]
    assert (exp[full] == g["hit_label"][full]).mean() >= 0.9995
    assert np.abs(hits["t"][full] - g["hit_t"][full]).max() <= 1e-4 * max(1.0, float(g["hit_t"][full].max()))
    assert np.abs(hits["normal"][full] - g["hit_normal"][full]).max() <= 1e-5
    assert np.abs(hits["point"][full] - g["hit_point"][full]).max() <= 1e-4
    hs.close()


def test_oracle_c1_bit_exact_vs_live_reference(rt, scene_dir, have_ref, tmp_path):
    if not have_ref:
        pytest.skip("oracle/_ref not built")
    xml = rt.scenes.scene_c1(os.path.join(scene_dir, "c1live"), res=200)
    subprocess.run([os.path.join(ROOT, "oracle", "_ref", "ref_raytracer"), xml], cwd=tmp_path, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    hs = rt.HostScene(xml)
    ref = rt.read_f32(os.path.join(tmp_path, hs.image_name(0) + ".f32"))
    img = orc.oracle_render(hs.desc, hs.camera(0))
    assert (img == ref).mean() >= 0.999
    hs.close()


@pytest.mark.parametrize("gen,kw", [("scene_c2", dict(res=64, spp=36)), ("scene_c3", dict(res=(96, 54), spp=16, n=24)),
                                    ("scene_c4", dict(res=(96, 54), spp=16, n=16)), ("scene_c5", dict(res=(64, 36), spp=36, sun_scale=0.01))])
def test_oracle_statistical_vs_live_reference(rt, scene_dir, have_ref, tmp_path, gen, kw):
    """Stochastic configurations: the oracle (Philox streams) and the reference (unseeded mt19937) must agree within
    the reference's own run-to-run noise on block means."""
    if not have_ref:
        pytest.skip("oracle/_ref not built")
    xml = getattr(rt.scenes, gen)(os.path.join(scene_dir, gen + "_stat"), **kw)
    hs = rt.HostScene(xml)
    refs = []
    for i in range(2):
        d = os.path.join(tmp_path, f"r{i}")
        os.makedirs(d)
        subprocess.run([os.path.join(ROOT, "oracle", "_ref", "ref_raytracer"), xml], cwd=d, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        refs.append(rt.read_f32(os.path.join(d, hs.image_name(0) + ".f32")))
    img = orc.oracle_render(hs.desc, hs.camera(0))

    from conftest import assert_statistically_equal
    assert_statistically_equal(refs[0], refs[1], img, b=6)
    hs.close()


@pytest.mark.parametrize("variant", ["lights", "brdf", "textures", "textures_jpeg"])
def test_oracle_feature_matrix_deterministic_vs_live_reference(rt, scene_dir, have_ref, tmp_path, variant):
    """Every light type, all eight BRDFs, every texture decal mode / interpolation / image format, transforms and the
    mirror / conductor / dielectric recursion, against the unmodified reference on deterministic scenes.  `textures_jpeg`:
    the reference decodes its JPEG textures through an independent libjpeg (oracle/shim/jpeglib.h -> Pillow), the oracle's
    scene comes from the product's own decoder (host/jpeg.h) - baseline, progressive and the greyscale neighbour-read quirk."""
    if not have_ref:
        pytest.skip("oracle/_ref not built")
    xml = rt.scenes.scene_features(os.path.join(scene_dir, "feat_" + variant), variant)
    subprocess.run([os.path.join(ROOT, "oracle", "_ref", "ref_raytracer"), xml], cwd=tmp_path, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    hs = rt.HostScene(xml)
    ref = rt.read_f32(os.path.join(tmp_path, hs.image_name(0) + ".f32"))
    img = orc.oracle_render(hs.desc, hs.camera(0))
    r = np.abs(img - ref) / np.maximum(np.abs(ref), 1.0)
    assert r.mean() <= 1e-5, r.mean()
    assert np.quantile(r, 0.999) <= 1e-3, np.quantile(r, 0.999)
    assert r.max() <= 5e-2, r.max()      # a nearest-neighbour texel can flip on a rounding difference
    hs.close()


@pytest.mark.parametrize("variant", ["stochastic", "stochastic_blur", "env"])
def test_oracle_feature_matrix_stochastic_vs_live_reference(rt, scene_dir, have_ref, tmp_path, variant):
    """Depth of field, motion blur, glossy reflection, area and environment lights, left-handed camera."""
    if not have_ref:
        pytest.skip("oracle/_ref not built")
    from conftest import assert_statistically_equal
    xml = rt.scenes.scene_features(os.path.join(scene_dir, "feat_" + variant), variant, res=(96, 72), spp=25)
    hs = rt.HostScene(xml)
    refs = []
    for i in range(2):
        d = os.path.join(tmp_path, f"r{i}")
        os.makedirs(d)
        subprocess.run([os.path.join(ROOT, "oracle", "_ref", "ref_raytracer"), xml], cwd=d, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        refs.append(rt.read_f32(os.path.join(d, hs.image_name(0) + ".f32")))
    img = orc.oracle_render(hs.desc, hs.camera(0))
    assert_statistically_equal(refs[0], refs[1], img, b=6)
    hs.close()

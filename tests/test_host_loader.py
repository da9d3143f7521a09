"""Host shell (XML / PLY / image I/O / scene preparation) against the behaviour of src/Parser.h, src/Helper.cpp:135-226
and src/Scene.cpp:303-318, including the parser quirks catalogued in SURVEY.md Appendix A.  CPU only."""
import math
import os
import struct

import numpy as np
import pytest

BASE = """<Scene>
  <MaxRecursionDepth>3</MaxRecursionDepth>
  <Cameras><Camera id="1"{camattr}><Position>0 0 0</Position>{gaze}<Up>0 1 0</Up>{plane}
    <NearDistance>2</NearDistance><ImageResolution>8 4</ImageResolution><ImageName>x.png</ImageName>{camextra}</Camera></Cameras>
  <Lights>{lights}</Lights>
  <Materials>{materials}</Materials>
  {textures}
  {transformations}
  <VertexData>{verts}</VertexData>
  {texcoords}
  <Objects>{objects}</Objects>
</Scene>"""

MAT = '<Material id="1"><AmbientReflectance>1 1 1</AmbientReflectance><DiffuseReflectance>.5 .5 .5</DiffuseReflectance><SpecularReflectance>0 0 0</SpecularReflectance></Material>'


def write_scene(tmp, **kw):
    d = dict(camattr="", gaze="<Gaze>0 0 -1</Gaze>", plane="<NearPlane>-1 1 -1 1</NearPlane>", camextra="", lights="<AmbientLight>1 2 3</AmbientLight>",
             materials=MAT, textures="", transformations="", verts="0 0 -5\n1 0 -5\n0 1 -5\n1 1 -5", texcoords="", objects="")
    d.update(kw)
    os.makedirs(tmp, exist_ok=True)
    p = os.path.join(tmp, "scene.xml")
    with open(p, "w") as f:
        f.write(BASE.format(**d))
    return p


def mat_from_inv(o):
    inv = np.eye(4)
    inv[:3, :] = np.array(list(o.inv)).reshape(3, 4)
    return np.linalg.inv(inv)


def test_defaults_and_camera(rt, tmp_path):
    p = write_scene(str(tmp_path), objects='<Sphere id="1"><Material>1</Material><Center>1</Center><Radius>2</Radius></Sphere>')
    hs = rt.HostScene(p)
    d = hs.desc.contents
    assert d.max_recursion_depth == 3
    assert d.shadow_ray_eps == pytest.approx(0.002) and d.int_test_eps == pytest.approx(0.001)   # Parser.h:23-25
    assert list(d.ambient_light) == [1, 2, 3]
    assert d.env_light_index == -1 and d.background_texture == -1
    cam = hs.camera(0)
    assert (cam.nx, cam.ny, cam.num_samples, cam.is_dof, cam.left_handed) == (8, 4, 1, 0, 0)
    assert hs.image_name(0) == "x.png"
    assert d.num_spheres == 1 and d.spheres[0].radius == 2 and list(d.spheres[0].center) == [0, 0, -5]
    hs.close()


def test_fovy_gazepoint_handedness_dof(rt, tmp_path):
    p = write_scene(str(tmp_path), camattr=' handedness="left"', gaze="<GazePoint>0 0 -10</GazePoint>", plane="<FovY>90</FovY>",
                    camextra="<NumSamples>9</NumSamples><FocusDistance>5</FocusDistance><ApertureSize>0.1</ApertureSize>",
                    objects='<Sphere id="1"><Material>1</Material><Center>1</Center><Radius>2</Radius></Sphere>')
    hs = rt.HostScene(p)
    cam = hs.camera(0)
    assert cam.left_handed == 1 and cam.is_dof == 1 and cam.num_samples == 9
    assert list(cam.gaze) == [0, 0, -10]
    assert cam.top == pytest.approx(2.0, rel=1e-6) and cam.right == pytest.approx(4.0, rel=1e-6)   # tan(45)*2, aspect 2
    hs.close()


def test_transform_composition_order_and_composite(rt, tmp_path):
    """'s1 r1 t1': M = T.R.S (listed order = application order, Helper.cpp:153-176); a composite REPLACES (173-175);
    'c' is only recognised as the first character of the list (Parser.h:792-794)."""
    tr = ('<Transformations><Translation id="1">1 2 3</Translation><Scaling id="1">2 2 2</Scaling><Rotation id="1">90 0 0 1</Rotation>'
          '<Composite id="1">1 0 0 5  0 1 0 6  0 0 1 7  0 0 0 1</Composite></Transformations>')
    objs = ('<Sphere id="1"><Material>1</Material><Center>1</Center><Radius>1</Radius><Transformations>s1 r1 t1</Transformations></Sphere>'
            '<Sphere id="2"><Material>1</Material><Center>1</Center><Radius>1</Radius><Transformations>c1 t1</Transformations></Sphere>'
            '<Sphere id="3"><Material>1</Material><Center>1</Center><Radius>1</Radius><Transformations>t1 c1</Transformations></Sphere>')
    hs = rt.HostScene(write_scene(str(tmp_path), transformations=tr, objects=objs))
    d = hs.desc.contents
    M = mat_from_inv(d.objects[0])
    S = np.diag([2, 2, 2, 1.0]); R = np.array([[0, -1, 0, 0], [1, 0, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1.0]]); T = np.eye(4); T[:3, 3] = [1, 2, 3]
    assert np.allclose(M, T @ R @ S, atol=1e-5)
    # 'c1 t1': iterate j = n-1..0: t1 first (M = T), then the composite replaces it -> M = C
    Cm = np.eye(4); Cm[:3, 3] = [5, 6, 7]
    assert np.allclose(mat_from_inv(d.objects[1]), Cm, atol=1e-5)
    # 't1 c1': the composite is not the first token -> skipped by the parser -> M = T
    assert np.allclose(mat_from_inv(d.objects[2]), T, atol=1e-5)
    # normal matrix is (M^-1)^T
    nrm = np.array(list(d.objects[0].inv_t)).reshape(3, 3)
    assert np.allclose(nrm, np.linalg.inv(M)[:3, :3].T, atol=1e-5)
    hs.close()


def test_instances_reset_transform_and_order(rt, tmp_path):
    tr = '<Transformations><Translation id="1">1 0 0</Translation><Translation id="2">0 5 0</Translation></Transformations>'
    objs = ('<Mesh id="7"><Material>1</Material><Transformations>t1</Transformations><Faces>1 2 3</Faces></Mesh>'
            '<MeshInstance id="8" baseMeshId="7" resetTransform="false"><Material>1</Material><Transformations>t2</Transformations></MeshInstance>'
            '<MeshInstance id="9" baseMeshId="7" resetTransform="true"><Material>1</Material><Transformations>t2</Transformations><MotionBlur>0 0 1</MotionBlur></MeshInstance>'
            '<Sphere id="1"><Material>1</Material><Center>1</Center><Radius>1</Radius></Sphere>')
    hs = rt.HostScene(write_scene(str(tmp_path), transformations=tr, objects=objs))
    d = hs.desc.contents
    # loop order of the reference: spheres, triangles, meshes, then instances (Helper.cpp:32-73)
    assert [d.objects[i].kind for i in range(4)] == [0, 1, 1, 1]
    assert np.allclose(mat_from_inv(d.objects[1])[:3, 3], [1, 0, 0])
    assert np.allclose(mat_from_inv(d.objects[2])[:3, 3], [1, 5, 0])     # own . base
    assert np.allclose(mat_from_inv(d.objects[3])[:3, 3], [0, 5, 0])     # reset
    assert list(d.objects[3].blur) == [0, 0, 1] and list(d.objects[2].blur) == [0, 0, 0]
    assert d.objects[2].geom == d.objects[1].geom == d.objects[3].geom
    hs.close()


def test_smooth_normals_and_standalone_triangles(rt, tmp_path):
    """Standalone <Triangle>s are always smooth-shaded and share the GLOBAL vertex normals."""
    objs = ('<Triangle id="1"><Material>1</Material><Indices>1 2 3</Indices></Triangle>'
            '<Mesh id="2"><Material>1</Material><Faces>2 4 3</Faces></Mesh>'
            '<Mesh id="3" shadingMode="smooth"><Material>1</Material><Faces>2 4 3</Faces></Mesh>')
    hs = rt.HostScene(write_scene(str(tmp_path), objects=objs))
    d = hs.desc.contents
    assert [d.meshes[i].smooth for i in range(3)] == [1, 0, 1]
    n = np.array([d.vertex_normals[i] for i in range(12)]).reshape(4, 3)
    assert np.allclose(n, [[0, 0, 1]] * 4, atol=1e-6)    # all faces lie in z = -5, CCW seen from +z
    hs.close()


def test_texture_state_persists_across_texture_maps(rt, tmp_path):
    """dm / interpolation / normalizer / scale / bump are declared outside the <TextureMap> loop (Parser.h:479-486)."""
    tex = ('<Textures><TextureMap id="1" type="perlin"><DecalMode>bump_normal</DecalMode><NoiseConversion>absval</NoiseConversion><NoiseScale>6</NoiseScale>'
           '<BumpFactor>0.3</BumpFactor></TextureMap><TextureMap id="2" type="perlin"><DecalMode>replace_kd</DecalMode></TextureMap></Textures>')
    objs = '<Sphere id="1"><Material>1</Material><Center>1</Center><Radius>1</Radius><Textures>1 2</Textures></Sphere>'
    hs = rt.HostScene(write_scene(str(tmp_path), textures=tex, objects=objs))
    d = hs.desc.contents
    t0, t1 = d.textures[0], d.textures[1]
    assert (t0.type, t0.decal_mode, t0.noise_conversion) == (1, 2, 0)
    assert t0.noise_scale == 6 and t0.bump_factor == pytest.approx(0.3)
    assert (t1.decal_mode, t1.noise_conversion) == (0, 0) and t1.noise_scale == 6 and t1.bump_factor == pytest.approx(0.3)   # inherited
    assert list(d.objects[0].tex) == [1, 2]
    hs.close()


def test_ply_quads_uv_and_offsets(rt, tmp_path):
    os.makedirs(tmp_path / "s", exist_ok=True)
    v = np.array([[0, 0, -5], [1, 0, -5], [1, 1, -5], [0, 1, -5]], dtype=np.float32)
    uv = np.array([[0, 0], [1, 0], [1, 1], [0, 1]], dtype=np.float32)
    rt.scenes.write_ply(str(tmp_path / "s" / "q.ply"), v, np.array([[0, 1, 2, 3]]), uv)
    tex = '<Textures><TextureMap id="1" type="perlin"><DecalMode>replace_kd</DecalMode></TextureMap></Textures>'
    objs = '<Mesh id="1"><Material>1</Material><Textures>1</Textures><Faces plyFile="q.ply" /></Mesh>'
    p = write_scene(str(tmp_path / "s"), textures=tex, objects=objs, texcoords="<TexCoordData>9 9\n 8 8</TexCoordData>")
    hs = rt.HostScene(p)
    d = hs.desc.contents
    assert d.num_vertices == 8 and d.num_triangles == 2          # 4 XML vertices + 4 PLY vertices; quad split (0,1,2),(2,3,0)
    assert [d.tri_indices[i] for i in range(6)] == [4, 5, 6, 6, 7, 4]
    assert d.num_texcoords == 6
    assert d.objects[0].tex_offset == (2 + 1) - (4 + 1)           # textureOffset - vertexOffset (Parser.h:1051,1065,1102)
    uvi = d.tri_indices[0] + d.objects[0].tex_offset
    assert (d.texcoords[2 * uvi], d.texcoords[2 * uvi + 1]) == (0.0, 0.0)
    hs.close()


def test_lights_order_and_env(rt, tmp_path):
    os.makedirs(tmp_path / "s", exist_ok=True)
    img = np.full((4, 8, 4), 2.5, dtype=np.float32)
    rt.scenes.write_exr(str(tmp_path / "s" / "env.exr"), img)
    lights = ('<SphericalDirectionalLight id="9"><ImageId>1</ImageId></SphericalDirectionalLight><AreaLight id="4"><Position>0 1 0</Position><Normal>0 -2 0</Normal>'
              '<Size>2</Size><Radiance>1 1 1</Radiance></AreaLight><SpotLight id="3"><Position>0 0 0</Position><Direction>0 0 -1</Direction><Intensity>1 1 1</Intensity>'
              '<CoverageAngle>60</CoverageAngle><FalloffAngle>40</FalloffAngle></SpotLight><DirectionalLight id="2"><Direction>0 -1 0</Direction><Radiance>3 3 3</Radiance></DirectionalLight>'
              '<PointLight id="1"><Position>1 1 1</Position><Intensity>5 5 5</Intensity></PointLight>')
    tex = "<Textures><Images><Image id=\"1\">env.exr</Image></Images></Textures>"
    objs = '<Sphere id="1"><Material>1</Material><Center>1</Center><Radius>1</Radius></Sphere>'
    hs = rt.HostScene(write_scene(str(tmp_path / "s"), lights=lights, textures=tex, objects=objs))
    d = hs.desc.contents
    assert [d.lights[i].type for i in range(5)] == [0, 2, 3, 1, 4]   # point, directional, spot, area, environment (Parser.h:1197-1315)
    assert d.env_light_index == 4 and d.lights[4].image == 0
    im = d.images[0]
    assert (im.width, im.height, im.format) == (8, 4, 1)
    hs.close()


def test_image_io_roundtrips(rt, tmp_path):
    lib = rt.load_core()
    rgb = (np.arange(6 * 5 * 3, dtype=np.float32).reshape(5, 6, 3) * 3.7) - 20
    # P3 PPM for ".png" names: clamp to 255, (unsigned char) truncation (src/Image.cpp:62-103)
    p = str(tmp_path / "out.png")
    assert lib.rt_host_save_image(p.encode(), rgb.ctypes.data, 6, 5) == 0
    toks = open(p).read().split()
    assert toks[:4] == ["P3", "6", "5", "255"]
    vals = np.array(toks[4:], dtype=np.int64).reshape(5, 6, 3)
    assert (vals == np.clip(rgb, 0, 255).astype(np.uint8)).all()
    # half-float B,G,R EXR otherwise (src/Helper.cpp:361-412); read it back through the texture path
    e = str(tmp_path / "out.exr")
    assert lib.rt_host_save_image(e.encode(), rgb.ctypes.data, 6, 5) == 0
    os.makedirs(tmp_path / "s", exist_ok=True)
    os.replace(e, tmp_path / "s" / "t.exr")
    tex = '<Textures><Images><Image id="1">t.exr</Image></Images><TextureMap id="1" type="image"><ImageId>1</ImageId><DecalMode>replace_kd</DecalMode></TextureMap></Textures>'
    objs = '<Sphere id="1"><Material>1</Material><Center>1</Center><Radius>1</Radius><Textures>1</Textures></Sphere>'
    hs = rt.HostScene(write_scene(str(tmp_path / "s"), textures=tex, objects=objs))
    im = hs.desc.contents.images[0]
    import ctypes as C
    px = np.ctypeslib.as_array(C.cast(im.pixels, C.POINTER(C.c_float)), shape=(5, 6, 4))
    assert np.allclose(px[..., :3], rgb.astype(np.float16).astype(np.float32))
    assert np.allclose(px[..., 3], 1.0)
    hs.close()


def test_png_decoder(rt, tmp_path):
    from PIL import Image
    os.makedirs(tmp_path / "s", exist_ok=True)
    a = (np.random.default_rng(1).integers(0, 255, size=(7, 9, 3))).astype(np.uint8)
    Image.fromarray(a).save(tmp_path / "s" / "t.png")
    tex = '<Textures><Images><Image id="1">t.png</Image></Images><TextureMap id="1" type="image"><ImageId>1</ImageId><DecalMode>replace_kd</DecalMode></TextureMap></Textures>'
    objs = '<Sphere id="1"><Material>1</Material><Center>1</Center><Radius>1</Radius><Textures>1</Textures></Sphere>'
    hs = rt.HostScene(write_scene(str(tmp_path / "s"), textures=tex, objects=objs))
    im = hs.desc.contents.images[0]
    import ctypes as C
    px = np.ctypeslib.as_array(C.cast(im.pixels, C.POINTER(C.c_ubyte)), shape=(7, 9, 4))
    assert (px[..., :3] == a).all() and (px[..., 3] == 255).all()
    hs.close()


def test_malformed_inputs_are_errors_not_crashes(rt, tmp_path):
    with pytest.raises(rt.RtError) as e:
        rt.HostScene(str(tmp_path / "missing.xml"))
    assert e.value.code == rt.RT_ERR_IO
    bad = tmp_path / "bad.xml"
    bad.write_text("<Scene><Cameras></Scene>")
    with pytest.raises(rt.RtError) as e:
        rt.HostScene(str(bad))
    assert e.value.code == rt.RT_ERR_PARSE
    p = write_scene(str(tmp_path / "a"), objects='<Sphere id="1"><Material>3</Material><Center>1</Center><Radius>1</Radius></Sphere>')
    with pytest.raises(rt.RtError) as e:
        rt.HostScene(p)
    assert e.value.code == rt.RT_ERR_PARSE and "Material" in str(e.value)
    p = write_scene(str(tmp_path / "b"), objects='<MeshInstance id="2" baseMeshId="5" resetTransform="true"><Material>1</Material></MeshInstance>')
    with pytest.raises(rt.RtError):
        rt.HostScene(p)
    # textured mesh without texture coordinates would read out of bounds in the reference
    tex = '<Textures><TextureMap id="1" type="perlin"><DecalMode>replace_kd</DecalMode></TextureMap></Textures>'
    p = write_scene(str(tmp_path / "c"), textures=tex, objects='<Mesh id="1"><Material>1</Material><Textures>1</Textures><Faces>1 2 3</Faces></Mesh>')
    with pytest.raises(rt.RtError):
        rt.HostScene(p)


# ---- scene blob (scene ingest)
def _desc_bytes(rt, hs):
    """Every array + scalar the core sees, as comparable python values."""
    import ctypes as C
    d = hs.desc.contents
    out = {"scalars": (d.max_recursion_depth, d.shadow_ray_eps, d.int_test_eps, tuple(d.background_color), tuple(d.ambient_light), d.env_light_index,
                       d.background_texture, d.num_vertices, d.num_texcoords, d.num_triangles, d.num_meshes, d.num_spheres, d.num_objects,
                       d.num_materials, d.num_textures, d.num_images, d.num_lights)}
    def raw(ptr, n, elem):
        return C.string_at(ptr, n * elem) if n and ptr else b""
    out["vertices"] = raw(d.vertices, d.num_vertices * 3, 4)
    out["normals"] = raw(d.vertex_normals, d.num_vertices * 3, 4) if d.vertex_normals else b""
    out["texcoords"] = raw(d.texcoords, d.num_texcoords * 2, 4)
    out["tris"] = raw(d.tri_indices, d.num_triangles * 3, 4)
    for name, n in (("meshes", d.num_meshes), ("spheres", d.num_spheres), ("objects", d.num_objects), ("materials", d.num_materials),
                    ("textures", d.num_textures), ("lights", d.num_lights)):
        arr = getattr(d, name)
        out[name] = raw(arr, n, C.sizeof(arr._type_)) if n else b""
    out["images"] = [(d.images[i].width, d.images[i].height, d.images[i].format,
                      C.string_at(d.images[i].pixels, d.images[i].width * d.images[i].height * (4 if d.images[i].format == rt.RT_IMG_RGBA8 else 16)))
                     for i in range(d.num_images)]
    out["cameras"] = [(bytes(hs.camera(i)), hs.image_name(i)) for i in range(hs.num_cameras)]
    return out


def test_scene_blob_round_trip_staleness_and_corruption(rt, tmp_path):
    d = str(tmp_path / "blob")
    xml = rt.scenes.scene_c4(d, res=(64, 36), spp=4, n=24, deterministic=True)      # PLY mesh + instances + EXR image + Perlin textures
    a = rt.HostScene(xml)
    blob = os.path.join(d, "scene.blob")
    a.save_blob(blob)
    a.save_blob(blob + ".again")                                             # byte-reproducible: no host addresses inside
    assert open(blob, "rb").read() == open(blob + ".again", "rb").read()
    b = rt.HostScene(blob=blob)
    assert _desc_bytes(rt, a) == _desc_bytes(rt, b)                                   # bit-identical prepared scene
    b.close()

    # cached loader: miss (writes <xml>.rt795blob), then hit
    c1 = rt.HostScene(xml, cached=True)
    assert not c1.cache_hit and os.path.exists(xml + ".rt795blob")
    c2 = rt.HostScene(xml, cached=True)
    assert c2.cache_hit and _desc_bytes(rt, a) == _desc_bytes(rt, c2)
    c1.close(); c2.close()

    # a changed source file makes the blob stale: explicit load refuses, the cached loader re-parses and rewrites
    ply = os.path.join(d, "c4_mesh.ply")
    st = os.stat(ply)
    os.utime(ply, ns=(st.st_atime_ns, st.st_mtime_ns + 5_000_000_000))
    with pytest.raises(rt.RtError, match="stale"):
        rt.HostScene(blob=blob)
    rt.HostScene(blob=blob, check_sources=False).close()                              # the caller may vouch for the sources
    c3 = rt.HostScene(xml, cached=True)
    assert not c3.cache_hit
    c3.close()
    c4 = rt.HostScene(xml, cached=True)
    assert c4.cache_hit
    c4.close()

    # corruption: wrong magic, truncation, absurd counts - errors, never crashes
    data = open(blob, "rb").read()
    bad = os.path.join(d, "bad.blob")
    for mutate in (lambda x: b"NOTABLOB" + x[8:], lambda x: x[: len(x) // 2], lambda x: x[:-3],
                   lambda x: x[:16] + struct.pack("<Q", 1 << 60) + x[24:], lambda x: b""):
        with open(bad, "wb") as f:
            f.write(mutate(data))
        with pytest.raises(rt.RtError):
            rt.HostScene(blob=bad, check_sources=False)
    with pytest.raises(rt.RtError):
        rt.HostScene(blob=os.path.join(d, "missing.blob"))
    a.close()


def _load_image_texture(rt, tmp, name):
    import ctypes as C
    tex = f'<Textures><Images><Image id="1">{name}</Image></Images><TextureMap id="1" type="image"><ImageId>1</ImageId><DecalMode>replace_kd</DecalMode></TextureMap></Textures>'
    objs = '<Sphere id="1"><Material>1</Material><Center>1</Center><Radius>1</Radius><Textures>1</Textures></Sphere>'
    hs = rt.HostScene(write_scene(tmp, textures=tex, objects=objs))
    im = hs.desc.contents.images[0]
    px = np.ctypeslib.as_array(C.cast(im.pixels, C.POINTER(C.c_ubyte)), shape=(im.height, im.width, 4)).copy()
    hs.close()
    return px


def test_jpeg_decoder_matches_libjpeg(rt, tmp_path):
    """Image textures in JPEG (Texture::ReadJPG, src/Texture.cpp:191-240): the sequential + progressive decoder reproduces libjpeg's default
    arithmetic (islow IDCT, fancy h2v1/h2v2 upsampling, fixed-point YCbCr->RGB), checked against PIL's libjpeg decode on smooth
    and noisy images, odd sizes, every common chroma subsampling, restart intervals and greyscale (with the reference's
    neighbour-read quirk for one-component files)."""
    from PIL import Image
    d = str(tmp_path / "j")
    os.makedirs(d, exist_ok=True)
    rng = np.random.default_rng(7)
    yy, xx = np.mgrid[0:61, 0:83]
    smooth = np.stack([127 + 120 * np.sin(xx / 9.0), 127 + 120 * np.cos(yy / 7.0), (xx * 3 + yy * 2) % 256], axis=-1).astype(np.uint8)
    noisy = rng.integers(0, 256, size=(37, 23, 3)).astype(np.uint8)
    tiny = rng.integers(0, 256, size=(5, 3, 3)).astype(np.uint8)
    cases = []
    for name, img in (("smooth", smooth), ("noisy", noisy), ("tiny", tiny)):
        for sub in (0, 1, 2):                                            # 4:4:4, 4:2:2, 4:2:0
            for q in (95, 60):
                cases.append((f"{name}_s{sub}_q{q}.jpg", img, dict(quality=q, subsampling=sub)))
    cases.append(("restart.jpg", smooth, dict(quality=85, subsampling=2, restart_marker_blocks=3)))
    cases.append(("optimized.jpg", smooth, dict(quality=85, subsampling=2, optimize=True)))
    for name, img in (("smooth", smooth), ("noisy", noisy), ("tiny", tiny)):        # progressive: DC/AC first + refinement scans, EOB runs
        for sub in (0, 1, 2):
            cases.append((f"{name}_prog_s{sub}.jpg", img, dict(quality=80, subsampling=sub, progressive=True)))
    cases.append(("prog_restart.jpg", smooth, dict(quality=30, subsampling=2, progressive=True, restart_marker_blocks=5)))
    worst = 0
    for fn, img, kw in cases:
        path = os.path.join(d, fn)
        try:
            Image.fromarray(img).save(path, format="JPEG", **kw)
        except TypeError:
            continue                                                     # an older PIL without restart_marker_blocks
        want = np.asarray(Image.open(path).convert("RGB"))
        got = _load_image_texture(rt, d, fn)
        assert got.shape[:2] == want.shape[:2] and (got[..., 3] == 255).all()
        diff = np.abs(got[..., :3].astype(int) - want.astype(int)).max()
        worst = max(worst, int(diff))
        assert diff <= 1, (fn, diff)                                     # same integer arithmetic; <= 1 leaves room for a libjpeg variant's rounding
    # greyscale: the reference reads raw[idx], raw[idx+1], raw[idx+2] with numComponents = 1 (src/Texture.cpp:62-66)
    g = (127 + 100 * np.sin(xx / 5.0) * np.cos(yy / 6.0)).astype(np.uint8)
    Image.fromarray(g, mode="L").save(os.path.join(d, "grey.jpg"), format="JPEG", quality=90)
    want = np.asarray(Image.open(os.path.join(d, "grey.jpg")))
    assert want.ndim == 2
    flat = want.ravel()
    idx = np.arange(flat.size)
    quirk = np.stack([flat[np.minimum(idx + k, flat.size - 1)] for k in range(3)], axis=-1).reshape(want.shape + (3,))
    got = _load_image_texture(rt, d, "grey.jpg")
    assert np.abs(got[..., :3].astype(int) - quirk.astype(int)).max() <= 1
    # unsupported flavours and corrupt files are errors with a reason, not crashes
    Image.fromarray(smooth).convert("CMYK").save(os.path.join(d, "cmyk.jpg"), format="JPEG")
    with pytest.raises(rt.RtError, match="CMYK"):
        _load_image_texture(rt, d, "cmyk.jpg")
    for src in ("smooth_s2_q95.jpg", "smooth_prog_s2.jpg"):
        data = open(os.path.join(d, src), "rb").read()
        for i, mutate in enumerate((lambda x: x[: len(x) // 3], lambda x: x[: len(x) * 2 // 3], lambda x: x[:200])):
            with open(os.path.join(d, f"cut{i}.jpg"), "wb") as f:
                f.write(mutate(data))
            try:
                _load_image_texture(rt, d, f"cut{i}.jpg")
            except rt.RtError:
                pass
    data = open(os.path.join(d, "smooth_s2_q95.jpg"), "rb").read()
    for i, mutate in enumerate((lambda x: x[: len(x) // 3], lambda x: x[:200], lambda x: b"\\xff\\xd8" + bytes(50), lambda x: x[:2] + x[40:])):
        with open(os.path.join(d, f"bad{i}.jpg"), "wb") as f:
            f.write(mutate(data))
        try:
            _load_image_texture(rt, d, f"bad{i}.jpg")                    # a truncated scan may still decode (zeros fed, like libjpeg); it must not crash
        except rt.RtError:
            pass


def test_mutated_inputs_never_crash(rt, tmp_path):
    """Seeded fuzzing of every file format the host shell reads (XML, PLY, PNG, EXR, JPEG baseline + progressive, scene blob):
    byte flips, truncations and spliced garbage must come back as an RtError or as a loadable scene - never as a crash or an
    exception escaping the C boundary.  Loaded blobs are also pushed through rt_scene_create's validation (which runs before
    any CUDA call, so this works without a GPU)."""
    from PIL import Image
    d = str(tmp_path / "fz")
    rng = np.random.default_rng(2026)
    xml = rt.scenes.scene_c4(d, res=(32, 18), spp=1, n=12, deterministic=True)
    hs = rt.HostScene(xml)
    hs.save_blob(os.path.join(d, "s.blob"))
    hs.close()
    img = rng.integers(0, 256, size=(21, 33, 3)).astype(np.uint8)
    Image.fromarray(img).save(os.path.join(d, "a.jpg"), quality=80)
    Image.fromarray(img).save(os.path.join(d, "p.jpg"), quality=80, progressive=True)
    Image.fromarray(img).save(os.path.join(d, "a.png"))
    rt.scenes.write_exr(os.path.join(d, "a.exr"), rng.random((9, 12, 3)).astype(np.float32))

    def mutations(data, n):
        for it in range(n):
            b = bytearray(data)
            if it % 3 == 0:
                for _ in range(int(rng.integers(1, 6))):
                    b[int(rng.integers(0, len(b)))] = int(rng.integers(0, 256))
            elif it % 3 == 1:
                b = b[: int(rng.integers(1, len(b)))]
            else:
                i = int(rng.integers(0, len(b)))
                j = int(rng.integers(i, min(len(b), i + 40)))
                b[i:j] = bytes(rng.integers(0, 256, size=int(rng.integers(0, 40))).astype(np.uint8))
            yield bytes(b)

    outcomes = {"loaded": 0, "rejected": 0}
    for src, ext in (("a.jpg", "jpg"), ("p.jpg", "jpg"), ("a.png", "png"), ("a.exr", "exr")):
        for m in mutations(open(os.path.join(d, src), "rb").read(), 45):
            with open(os.path.join(d, "f." + ext), "wb") as f:
                f.write(m)
            try:
                _load_image_texture(rt, d, "f." + ext)
                outcomes["loaded"] += 1
            except rt.RtError:
                outcomes["rejected"] += 1
    for which, target in (("ply", os.path.join(d, "c4_mesh.ply")), ("xml", xml), ("blob", os.path.join(d, "s.blob"))):
        orig = open(target, "rb").read()
        for m in mutations(orig, 45):
            with open(target, "wb") as f:
                f.write(m)
            try:
                h = rt.HostScene(blob=target, check_sources=False) if which == "blob" else rt.HostScene(xml)
            except rt.RtError:
                outcomes["rejected"] += 1
                continue
            try:
                rt.Scene(h.desc, 0).close()        # Validate() first; then RT_ERR_CUDA on a CPU-only box
            except rt.RtError:
                pass
            h.close()
            outcomes["loaded"] += 1
        with open(target, "wb") as f:
            f.write(orig)
    assert outcomes["loaded"] > 0 and outcomes["rejected"] > 0


def test_targeted_corruptions_are_rejected(rt, tmp_path):
    """The three structural corruptions a byte-flip fuzzer rarely hits: a SECOND frame header after the first progressive scan (it
    would re-size the coefficient planes under scans already decoded), an IHDR chunk shorter than 13 bytes, and a PLY face list whose
    count field is far larger than the file."""
    import struct
    import zlib
    from PIL import Image
    d = str(tmp_path / "tc")
    os.makedirs(d)
    rng = np.random.default_rng(7)
    img = rng.integers(0, 256, size=(40, 56, 3)).astype(np.uint8)
    # (1) progressive JPEG with its SOF2 segment repeated (smaller dimensions the second time) right before the second scan
    Image.fromarray(img).save(os.path.join(d, "p.jpg"), quality=85, progressive=True)
    data = open(os.path.join(d, "p.jpg"), "rb").read()
    sof = data.index(b"\xff\xc2")
    seglen = struct.unpack(">H", data[sof + 2:sof + 4])[0]
    seg = bytearray(data[sof:sof + 2 + seglen])
    seg[5:9] = struct.pack(">HH", 8, 8)                                       # H, W of the forged frame
    sos = [i for i in range(len(data) - 1) if data[i] == 0xff and data[i + 1] == 0xda]
    assert len(sos) >= 2
    forged = data[:sos[1]] + bytes(seg) + data[sos[1]:]
    with open(os.path.join(d, "dup.jpg"), "wb") as f:
        f.write(forged)
    with pytest.raises(rt.RtError) as e:
        _load_image_texture(rt, d, "dup.jpg")
    assert "second frame header" in str(e.value)
    # (2) PNG whose IHDR chunk is 9 bytes long
    def chunk(t, body):
        c = struct.pack(">I", len(body)) + t + body
        return c + struct.pack(">I", zlib.crc32(t + body) & 0xffffffff)
    raw = b"".join(b"\x00" + img[y].tobytes() for y in range(img.shape[0]))
    short = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIB", 56, 40, 8)) + chunk(b"IDAT", zlib.compress(raw)) + chunk(b"IEND", b"")
    with open(os.path.join(d, "short.png"), "wb") as f:
        f.write(short)
    with pytest.raises(rt.RtError) as e:
        _load_image_texture(rt, d, "short.png")
    assert "IHDR" in str(e.value)
    # (3) binary PLY: the first face claims 2^31 - 1 indices
    xml = rt.scenes.scene_c4(d, res=(32, 18), spp=1, n=10, deterministic=True)
    ply = os.path.join(d, "c4_mesh.ply")
    b = bytearray(open(ply, "rb").read())
    hdr_end = b.index(b"end_header\n") + len(b"end_header\n")
    header = bytes(b[:hdr_end]).decode()
    assert "property list uchar int" in header
    patched = header.replace("property list uchar int", "property list int int")
    nverts = int(header.split("element vertex ")[1].split()[0])
    nprops = header.split("element face")[0].count("property float")
    body = bytes(b[hdr_end:])
    faces_at = nverts * nprops * 4
    forged = patched.encode() + body[:faces_at] + struct.pack("<i", 2 ** 31 - 1) + body[faces_at + 1:]
    with open(ply, "wb") as f:
        f.write(forged)
    with pytest.raises(rt.RtError):
        rt.HostScene(xml)


def test_camera_renderer_and_tonemap_extensions(rt, tmp_path):
    """<Renderer> / <RendererParams> / <Tonemap> in a camera (course vocabulary; not parsed by the reference snapshot, SURVEY
    App. B) -> rt_host_camera_render_opts / rt_host_camera_tonemap, and they survive the scene blob."""
    d = str(tmp_path / "ext")
    extra = ("<Renderer>PathTracing</Renderer><RendererParams>NextEventEstimation ImportanceSampling</RendererParams>"
             "<Tonemap><TMO>Photographic</TMO><TMOOptions>0.25 2.5</TMOOptions><Saturation>0.8</Saturation><Gamma>2.4</Gamma></Tonemap>")
    xml = write_scene(d, camextra=extra)
    hs = rt.HostScene(xml)
    o, t = hs.render_opts(0), hs.tonemap(0)
    assert o.integrator == rt.RT_INTEGRATOR_PATH and (o.pt_nee, o.pt_russian_roulette, o.pt_importance) == (1, 0, 1) and o.seed == 795
    assert t.enabled == 1 and (t.key, t.burn_percent, t.saturation, t.gamma) == pytest.approx((0.25, 2.5, 0.8, 2.4))
    blob = os.path.join(d, "s.blob")
    hs.save_blob(blob)
    hb = rt.HostScene(blob=blob)
    assert bytes(hb.render_opts(0)) == bytes(o) and bytes(hb.tonemap(0)) == bytes(t)
    with pytest.raises(rt.RtError):
        hs.render_opts(3)
    hb.close(); hs.close()
    # a plain reference scene: the reference integrator, no tone map
    hs = rt.HostScene(write_scene(str(tmp_path / "plain")))
    o, t = hs.render_opts(0), hs.tonemap(0)
    assert o.integrator == rt.RT_INTEGRATOR_REFERENCE and t.enabled == 0
    hs.close()

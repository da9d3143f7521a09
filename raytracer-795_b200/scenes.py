"""Seeded synthetic scene generators for BASELINE.json configs C1..C5.

No assets can be fetched, so every scene is procedural and written in the reference's own input formats:
scene XML in the grammar of /root/reference/src/Parser.h, meshes as binary
little-endian PLY with x,y,z,u,v float32 vertices and uchar-counted int32 face lists (what happly reads at
Parser.h:1049-1097), HDR images as uncompressed float32 scanline EXR (what tinyexr's LoadEXR reads at
Helper.cpp:346-360).  Rules honoured so that the reference itself can render these files:
  * no <?xml?> prolog, <Scene> is the first node (Scene.cpp:461);
  * scene paths always contain a '/' (Parser.h:488-498);
  * every textured mesh carries uv (Shape.cpp:509-518);
  * closed rooms around dielectrics so refracted rays always hit (Scene.cpp:110).
"""
from __future__ import annotations

import math
import os
import struct

import numpy as np


# --------------------------------------------------------------------------------------------- file writers
def write_ply(path: str, verts: np.ndarray, faces: np.ndarray, uv: np.ndarray | None = None) -> None:
    """Binary little-endian PLY: float32 x y z [u v]; faces as uchar count + int32 indices (0-based)."""
    verts = np.asarray(verts, dtype="<f4")
    faces = np.asarray(faces, dtype="<i4")
    n, m = len(verts), len(faces)
    k = faces.shape[1]
    hdr = ["ply", "format binary_little_endian 1.0", f"element vertex {n}",
           "property float x", "property float y", "property float z"]
    if uv is not None:
        hdr += ["property float u", "property float v"]
    hdr += [f"element face {m}", "property list uchar int vertex_indices", "end_header"]
    with open(path, "wb") as f:
        f.write(("\n".join(hdr) + "\n").encode())
        if uv is not None:
            rec = np.concatenate([verts, np.asarray(uv, dtype="<f4")], axis=1)
        else:
            rec = verts
        f.write(np.ascontiguousarray(rec, dtype="<f4").tobytes())
        frec = np.empty(m, dtype=[("n", "u1"), ("idx", "<i4", (k,))])
        frec["n"] = k
        frec["idx"] = faces
        f.write(frec.tobytes())


def write_exr(path: str, img: np.ndarray) -> None:
    """Uncompressed scanline OpenEXR, float32 channels (written A?,B,G,R alphabetical as the format requires).

    img: (h, w, 3) or (h, w, 4) float32 RGB(A).  Readable by tinyexr LoadEXR (reference) and by our host reader.
    """
    img = np.asarray(img, dtype="<f4")
    h, w, c = img.shape
    names = ["B", "G", "R"] if c == 3 else ["A", "B", "G", "R"]
    src = {"R": 0, "G": 1, "B": 2, "A": 3}

    def attr(name: str, typ: str, data: bytes) -> bytes:
        return name.encode() + b"\0" + typ.encode() + b"\0" + struct.pack("<i", len(data)) + data

    chlist = b"".join(n.encode() + b"\0" + struct.pack("<iBBBBii", 2, 0, 0, 0, 0, 1, 1) for n in names) + b"\0"
    box = struct.pack("<iiii", 0, 0, w - 1, h - 1)
    hdr = struct.pack("<ii", 20000630, 2)
    hdr += attr("channels", "chlist", chlist)
    hdr += attr("compression", "compression", b"\0")
    hdr += attr("dataWindow", "box2i", box)
    hdr += attr("displayWindow", "box2i", box)
    hdr += attr("lineOrder", "lineOrder", b"\0")
    hdr += attr("pixelAspectRatio", "float", struct.pack("<f", 1.0))
    hdr += attr("screenWindowCenter", "v2f", struct.pack("<ff", 0.0, 0.0))
    hdr += attr("screenWindowWidth", "float", struct.pack("<f", 1.0))
    hdr += b"\0"
    line_bytes = len(names) * w * 4
    off0 = len(hdr) + 8 * h
    offsets = np.arange(h, dtype="<u8") * (8 + line_bytes) + off0
    with open(path, "wb") as f:
        f.write(hdr)
        f.write(offsets.tobytes())
        for y in range(h):
            f.write(struct.pack("<ii", y, line_bytes))
            for n in names:
                f.write(np.ascontiguousarray(img[y, :, src[n]]).tobytes())


def _fmt(v) -> str:
    return " ".join(repr(float(x)) for x in v)


# ----------------------------------------------------------------------------------------- procedural mesh
def displaced_sphere(n: int, third_harmonic: bool = False):
    """Lat-long displaced sphere, n x n quads -> 2*n*(n-1) triangles... see SURVEY §8d (C3: n=186, C4: n=660).

    r = 1 + 0.15 sin5p sin7t + 0.05 sin(23p+1.3) cos17t (+ 0.01 sin97p sin89t); CCW-outward winding for the
    reference normal (c-b)x(a-b) (Shape.cpp:266).  Returns verts (V,3) f32, faces (F,3) i32, uv (V,2) f32.
    Pole rows are collapsed into fans, so F = 2*n*(n-1) and V = n*(n-1)+2... computed exactly below.
    """
    nt, npi = n, n  # theta rows (pole to pole), phi columns (wrap)
    th = (np.arange(1, nt) / nt) * math.pi  # interior rings
    ph = (np.arange(npi) / npi) * 2.0 * math.pi
    T, P = np.meshgrid(th, ph, indexing="ij")

    def radius(t, p):
        r = 1.0 + 0.15 * np.sin(5 * p) * np.sin(7 * t) + 0.05 * np.sin(23 * p + 1.3) * np.cos(17 * t)
        if third_harmonic:
            r = r + 0.01 * np.sin(97 * p) * np.sin(89 * t)
        return r

    R = radius(T, P)
    x = R * np.sin(T) * np.cos(P)
    y = R * np.cos(T)
    z = R * np.sin(T) * np.sin(P)
    ring = np.stack([x, y, z], axis=-1).reshape(-1, 3)
    top = np.array([[0.0, float(radius(0.0, 0.0)), 0.0]])
    bot = np.array([[0.0, -float(radius(math.pi, 0.0)), 0.0]])
    verts = np.concatenate([ring, top, bot]).astype(np.float32)
    uv_ring = np.stack([P / (2 * math.pi), T / math.pi], axis=-1).reshape(-1, 2)
    uv = np.concatenate([uv_ring, [[0.5, 0.0]], [[0.5, 1.0]]]).astype(np.float32)
    itop, ibot = len(ring), len(ring) + 1

    def vid(i, j):
        return i * npi + (j % npi)

    i = np.arange(nt - 2)[:, None]
    j = np.arange(npi)[None, :]
    a = (i * npi + j).ravel()
    b = (i * npi + (j + 1) % npi).ravel()
    c = ((i + 1) * npi + (j + 1) % npi).ravel()
    d = ((i + 1) * npi + j).ravel()
    # outward = CCW when seen from outside.  With the reference normal (c-b)x(a-b) for (a,b,c) the order a,b,c
    # must be counter-clockwise from outside: (upper-left, lower-left, lower-right) in (phi right, theta down).
    f1 = np.stack([a, d, c], axis=1)
    f2 = np.stack([a, c, b], axis=1)
    jj = np.arange(npi)
    ftop = np.stack([np.full(npi, itop), jj, (jj + 1) % npi], axis=1)
    base = (nt - 2) * npi
    fbot = np.stack([np.full(npi, ibot), base + (jj + 1) % npi, base + jj], axis=1)
    faces = np.concatenate([f1, f2, ftop, fbot]).astype(np.int32)
    # orientation check/fix so that the reference's normal points outward for every face
    A, B, C = verts[faces[:, 0]], verts[faces[:, 1]], verts[faces[:, 2]]
    nrm = np.cross(C - B, A - B)
    cen = (A + B + C) / 3.0
    flip = np.einsum("ij,ij->i", nrm, cen) < 0
    faces[flip] = faces[flip][:, [0, 2, 1]]
    return verts, faces, uv


# ------------------------------------------------------------------------------------------------ scenes
def scene_c1(outdir: str, res: int = 800, name: str = "c1") -> str:
    """C1: Whitted simple.xml-style: 5 spheres + 4 triangles, 1 point light, mirror/diffuse, depth 6, 1 spp."""
    os.makedirs(outdir, exist_ok=True)
    verts = [
        (-4.0, 1.5, 0.0), (0.0, 2.5, -3.0), (4.5, 2.0, 1.0), (-1.5, 1.5, 5.0), (2.0, 3.0, -9.0),  # sphere centres 1..5
        (-14.0, 0.0, 14.0), (14.0, 0.0, 14.0), (14.0, 0.0, -14.0), (-14.0, 0.0, -14.0),           # floor 6..9
        (-12.0, 0.0, -13.5), (12.0, 0.0, -13.5), (12.0, 11.0, -13.5), (-12.0, 11.0, -13.5),       # back mirror 10..13
    ]
    xml = f"""<Scene>
  <MaxRecursionDepth>6</MaxRecursionDepth>
  <BackgroundColor>12 18 30</BackgroundColor>
  <ShadowRayEpsilon>1e-3</ShadowRayEpsilon>
  <IntersectionTestEpsilon>1e-6</IntersectionTestEpsilon>
  <Cameras>
    <Camera id="1">
      <Position>1.3 7.1 26.7</Position>
      <Gaze>-0.04 -0.19 -1</Gaze>
      <Up>0 1 0</Up>
      <NearPlane>-1 1 -1 1</NearPlane>
      <NearDistance>2.2</NearDistance>
      <ImageResolution>{res} {res}</ImageResolution>
      <ImageName>{name}.ppm</ImageName>
    </Camera>
  </Cameras>
  <Lights>
    <AmbientLight>25 25 25</AmbientLight>
    <PointLight id="1">
      <Position>6 16 12</Position>
      <Intensity>90000 90000 90000</Intensity>
    </PointLight>
  </Lights>
  <Materials>
    <Material id="1">
      <AmbientReflectance>1 1 1</AmbientReflectance>
      <DiffuseReflectance>0.9 0.25 0.2</DiffuseReflectance>
      <SpecularReflectance>0.6 0.6 0.6</SpecularReflectance>
      <PhongExponent>40</PhongExponent>
    </Material>
    <Material id="2" type="mirror">
      <AmbientReflectance>0.1 0.1 0.1</AmbientReflectance>
      <DiffuseReflectance>0.1 0.1 0.1</DiffuseReflectance>
      <SpecularReflectance>0.4 0.4 0.4</SpecularReflectance>
      <PhongExponent>100</PhongExponent>
      <MirrorReflectance>0.85 0.85 0.85</MirrorReflectance>
    </Material>
    <Material id="3">
      <AmbientReflectance>1 1 1</AmbientReflectance>
      <DiffuseReflectance>0.2 0.7 0.3</DiffuseReflectance>
      <SpecularReflectance>0.3 0.3 0.3</SpecularReflectance>
      <PhongExponent>10</PhongExponent>
    </Material>
    <Material id="4">
      <AmbientReflectance>1 1 1</AmbientReflectance>
      <DiffuseReflectance>0.55 0.55 0.6</DiffuseReflectance>
      <SpecularReflectance>0.1 0.1 0.1</SpecularReflectance>
      <PhongExponent>3</PhongExponent>
    </Material>
    <Material id="5" type="mirror">
      <AmbientReflectance>0.05 0.05 0.05</AmbientReflectance>
      <DiffuseReflectance>0.05 0.05 0.08</DiffuseReflectance>
      <SpecularReflectance>0.2 0.2 0.2</SpecularReflectance>
      <PhongExponent>200</PhongExponent>
      <MirrorReflectance>0.7 0.7 0.75</MirrorReflectance>
    </Material>
    <Material id="6">
      <AmbientReflectance>1 1 1</AmbientReflectance>
      <DiffuseReflectance>0.25 0.35 0.9</DiffuseReflectance>
      <SpecularReflectance>0.8 0.8 0.8</SpecularReflectance>
      <PhongExponent>60</PhongExponent>
    </Material>
  </Materials>
  <VertexData>
{chr(10).join("    " + _fmt(v) for v in verts)}
  </VertexData>
  <Objects>
    <Sphere id="1"><Material>1</Material><Center>1</Center><Radius>1.5</Radius></Sphere>
    <Sphere id="2"><Material>2</Material><Center>2</Center><Radius>2.5</Radius></Sphere>
    <Sphere id="3"><Material>3</Material><Center>3</Center><Radius>2.0</Radius></Sphere>
    <Sphere id="4"><Material>6</Material><Center>4</Center><Radius>1.5</Radius></Sphere>
    <Sphere id="5"><Material>2</Material><Center>5</Center><Radius>3.0</Radius></Sphere>
    <Triangle id="6"><Material>4</Material><Indices>6 7 8</Indices></Triangle>
    <Triangle id="7"><Material>4</Material><Indices>6 8 9</Indices></Triangle>
    <Triangle id="8"><Material>5</Material><Indices>10 11 12</Indices></Triangle>
    <Triangle id="9"><Material>5</Material><Indices>10 12 13</Indices></Triangle>
  </Objects>
</Scene>
"""
    path = os.path.join(outdir, f"{name}.xml")
    with open(path, "w") as f:
        f.write(xml)
    return path


# ------------------------------------------------------------------------------------------ shared XML helpers
def _material(mid, ka, kd, ks, phong=1, mtype=None, mirror=None, extra="", brdf=None):
    attr = f' id="{mid}"' + (f' type="{mtype}"' if mtype else "") + (f' BRDF="{brdf}"' if brdf else "")
    s = f"    <Material{attr}>\n"
    s += f"      <AmbientReflectance>{_fmt(ka)}</AmbientReflectance>\n"
    s += f"      <DiffuseReflectance>{_fmt(kd)}</DiffuseReflectance>\n"
    s += f"      <SpecularReflectance>{_fmt(ks)}</SpecularReflectance>\n"
    s += f"      <PhongExponent>{int(phong)}</PhongExponent>\n"
    if mirror is not None:
        s += f"      <MirrorReflectance>{_fmt(mirror)}</MirrorReflectance>\n"
    s += extra
    s += "    </Material>\n"
    return s


def _quad_faces(verts, quad, want):
    """Two triangles for quad (4 zero-based vertex ids) wound so the reference normal (c-b)x(a-b) points along `want`."""
    a, b, c, d = quad
    tris = [(a, b, c), (a, c, d)]
    out = []
    for (i, j, k) in tris:
        A, B, C = (np.array(verts[i]), np.array(verts[j]), np.array(verts[k]))
        n = np.cross(C - B, A - B)
        out.append((i, j, k) if float(np.dot(n, want)) > 0 else (i, k, j))
    return out


def _faces_text(tris):
    return "\n".join("        " + " ".join(str(i + 1) for i in t) for t in tris)


def scene_c2(outdir: str, res: int = 1024, spp: int = 100, name: str = "c2", depth: int = 6) -> str:
    """C2: closed Cornell box (3 meshes, 12 triangles) + dielectric sphere + rough conductor sphere, one area light,
    distributed ray tracing.  The box is CLOSED (front wall behind the camera) so refracted rays always hit (A.7)."""
    os.makedirs(outdir, exist_ok=True)
    X, Y, Z0, Z1 = 10.0, 20.0, -10.0, 30.0
    verts = [(-X, 0, Z1), (X, 0, Z1), (X, 0, Z0), (-X, 0, Z0),      # floor 0-3
             (-X, Y, Z1), (X, Y, Z1), (X, Y, Z0), (-X, Y, Z0),      # ceiling 4-7
             (-4.0, 4.0, 6.0), (4.6, 3.5, 0.5)]                     # sphere centres 8, 9
    white = (_quad_faces(verts, (0, 1, 2, 3), (0, 1, 0)) + _quad_faces(verts, (4, 5, 6, 7), (0, -1, 0)) +
             _quad_faces(verts, (3, 2, 6, 7), (0, 0, 1)) + _quad_faces(verts, (0, 1, 5, 4), (0, 0, -1)))
    red = _quad_faces(verts, (0, 3, 7, 4), (1, 0, 0))
    green = _quad_faces(verts, (1, 2, 6, 5), (-1, 0, 0))
    mats = (_material(1, (1, 1, 1), (0.78, 0.78, 0.78), (0, 0, 0)) + _material(2, (1, 1, 1), (0.8, 0.12, 0.1), (0, 0, 0)) +
            _material(3, (1, 1, 1), (0.12, 0.7, 0.15), (0, 0, 0)) +
            _material(4, (0, 0, 0), (0, 0, 0), (0, 0, 0), 1, "dielectric", None,
                      "      <RefractionIndex>1.5</RefractionIndex>\n      <AbsorptionCoefficient>0.01 0.01 0.01</AbsorptionCoefficient>\n") +
            _material(5, (0, 0, 0), (0.05, 0.05, 0.05), (0.3, 0.3, 0.3), 50, "conductor", (0.9, 0.8, 0.65),
                      "      <RefractionIndex>0.37</RefractionIndex>\n      <AbsorptionIndex>2.82</AbsorptionIndex>\n      <Roughness>0.15</Roughness>\n"))
    xml = f"""<Scene>
  <MaxRecursionDepth>{depth}</MaxRecursionDepth>
  <BackgroundColor>0 0 0</BackgroundColor>
  <ShadowRayEpsilon>1e-3</ShadowRayEpsilon>
  <IntersectionTestEpsilon>1e-6</IntersectionTestEpsilon>
  <Cameras>
    <Camera id="1">
      <Position>0.37 10.3 28.7</Position>
      <Gaze>-0.013 -0.031 -1</Gaze>
      <Up>0 1 0</Up>
      <NearPlane>-1 1 -1 1</NearPlane>
      <NearDistance>1.9</NearDistance>
      <ImageResolution>{res} {res}</ImageResolution>
      <NumSamples>{spp}</NumSamples>
      <ImageName>{name}.ppm</ImageName>
    </Camera>
  </Cameras>
  <Lights>
    <AmbientLight>8 8 8</AmbientLight>
    <AreaLight id="1">
      <Position>0 18.6 8</Position>
      <Normal>0 -1 0</Normal>
      <Size>4</Size>
      <Radiance>2000 1900 1700</Radiance>
    </AreaLight>
  </Lights>
  <Materials>
{mats}  </Materials>
  <VertexData>
{chr(10).join("    " + _fmt(v) for v in verts)}
  </VertexData>
  <Objects>
    <Sphere id="1"><Material>4</Material><Center>9</Center><Radius>4.0</Radius></Sphere>
    <Sphere id="2"><Material>5</Material><Center>10</Center><Radius>3.5</Radius></Sphere>
    <Mesh id="3">
      <Material>1</Material>
      <Faces>
{_faces_text(white)}
      </Faces>
    </Mesh>
    <Mesh id="4">
      <Material>2</Material>
      <Faces>
{_faces_text(red)}
      </Faces>
    </Mesh>
    <Mesh id="5">
      <Material>3</Material>
      <Faces>
{_faces_text(green)}
      </Faces>
    </Mesh>
  </Objects>
</Scene>
"""
    path = os.path.join(outdir, f"{name}.xml")
    with open(path, "w") as f:
        f.write(xml)
    return path


def scene_c3(outdir: str, res=(1920, 1080), spp: int = 16, n: int = 187, name: str = "c3", deterministic: bool = False) -> str:
    """C3: bunny-scale (~69k triangle) smooth PLY mesh + floor, point + area light, 16 spp."""
    os.makedirs(outdir, exist_ok=True)
    v, f, uv = displaced_sphere(n)
    write_ply(os.path.join(outdir, f"{name}_mesh.ply"), v, f, uv)
    verts = [(-9.0, -1.35, 9.0), (9.0, -1.35, 9.0), (9.0, -1.35, -9.0), (-9.0, -1.35, -9.0)]
    floor = _quad_faces(verts, (0, 1, 2, 3), (0, 1, 0))
    mats = (_material(1, (1, 1, 1), (0.75, 0.55, 0.3), (0.5, 0.5, 0.5), 60) + _material(2, (1, 1, 1), (0.6, 0.6, 0.65), (0.05, 0.05, 0.05), 4))
    area = "" if deterministic else """    <AreaLight id="1">
      <Position>-4 5 2</Position>
      <Normal>0.6 -0.75 -0.28</Normal>
      <Size>2</Size>
      <Radiance>900 950 1100</Radiance>
    </AreaLight>
"""
    xml = f"""<Scene>
  <MaxRecursionDepth>2</MaxRecursionDepth>
  <BackgroundColor>20 24 36</BackgroundColor>
  <ShadowRayEpsilon>1e-3</ShadowRayEpsilon>
  <IntersectionTestEpsilon>1e-6</IntersectionTestEpsilon>
  <Cameras>
    <Camera id="1">
      <Position>0.31 1.17 4.6</Position>
      <GazePoint>0.02 -0.11 0.01</GazePoint>
      <Up>0 1 0</Up>
      <FovY>38</FovY>
      <NearDistance>1</NearDistance>
      <ImageResolution>{res[0]} {res[1]}</ImageResolution>
      <NumSamples>{spp}</NumSamples>
      <ImageName>{name}.ppm</ImageName>
    </Camera>
  </Cameras>
  <Lights>
    <AmbientLight>18 18 18</AmbientLight>
    <PointLight id="1">
      <Position>4 6 5</Position>
      <Intensity>9000 8600 8000</Intensity>
    </PointLight>
{area}  </Lights>
  <Materials>
{mats}  </Materials>
  <VertexData>
{chr(10).join("    " + _fmt(p) for p in verts)}
  </VertexData>
  <Objects>
    <Mesh id="1">
      <Material>2</Material>
      <Faces>
{_faces_text(floor)}
      </Faces>
    </Mesh>
    <Mesh id="2" shadingMode="smooth">
      <Material>1</Material>
      <Faces plyFile="{name}_mesh.ply" />
    </Mesh>
  </Objects>
</Scene>
"""
    path = os.path.join(outdir, f"{name}.xml")
    with open(path, "w") as fh:
        fh.write(xml)
    return path


def checker_exr(path: str, w: int = 256, h: int = 256, seed: int = 795) -> None:
    """Generated image texture (float RGBA EXR): soft coloured checker + seeded noise, values in 0..255."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    chk = ((xx // 16 + yy // 16) % 2).astype(np.float32)
    img = np.empty((h, w, 4), dtype=np.float32)
    img[..., 0] = 60 + 160 * chk + 20 * np.sin(xx * 0.11)
    img[..., 1] = 70 + 120 * (1 - chk) + 25 * np.cos(yy * 0.07)
    img[..., 2] = 90 + 60 * chk + rng.uniform(0, 30, size=(h, w))
    img[..., 3] = 1.0
    write_exr(path, img)


def scene_c4(outdir: str, res=(1920, 1080), spp: int = 64, n: int = 660, name: str = "c4", grid: int = 4, deterministic: bool = False) -> str:
    """C4: dragon-scale (~870k triangle) smooth mesh, 1 base + grid*grid-1 instances (TLAS/BLAS), transformations,
    motion blur on every 5th entry, Perlin bump on the mesh, Perlin kd + image texture on the ground."""
    os.makedirs(outdir, exist_ok=True)
    v, f, uv = displaced_sphere(n, third_harmonic=True)
    write_ply(os.path.join(outdir, f"{name}_mesh.ply"), v, f, uv)
    checker_exr(os.path.join(outdir, f"{name}_floor.exr"))
    S = 3.2 * grid
    verts = [(-S, -1.4, S), (S, -1.4, S), (S, -1.4, -S), (-S, -1.4, -S),        # ground 0-3
             (-S, -1.4, -S), (S, -1.4, -S), (S, 9.0, -S), (-S, 9.0, -S)]        # back wall 4-7
    uvs = [(0, 0), (4, 0), (4, 4), (0, 4), (0, 0), (1, 0), (1, 1), (0, 1)]
    ground = _quad_faces(verts, (0, 1, 2, 3), (0, 1, 0))
    wall = _quad_faces(verts, (4, 5, 6, 7), (0, 0, 1))
    mats = (_material(1, (1, 1, 1), (0.7, 0.45, 0.25), (0.6, 0.6, 0.6), 80) + _material(2, (1, 1, 1), (0.5, 0.5, 0.5), (0.05, 0.05, 0.05), 4) +
            _material(3, (1, 1, 1), (0.25, 0.4, 0.75), (0.7, 0.7, 0.7), 120) + _material(4, (1, 1, 1), (0.3, 0.7, 0.35), (0.4, 0.4, 0.4), 30) +
            _material(5, (1, 1, 1), (0.8, 0.8, 0.8), (0, 0, 0), 1))
    area = "" if deterministic else """    <AreaLight id="1">
      <Position>-7 12 3</Position>
      <Normal>0.5 -0.8 -0.2</Normal>
      <Size>3</Size>
      <Radiance>2600 2700 3000</Radiance>
    </AreaLight>
"""
    trans, objs = [], []
    tid = 0
    sid = [0]
    scal, rots = [], []
    inst_xml = []
    for k in range(grid * grid):
        gx, gz = k % grid, k // grid
        px, pz = (gx - (grid - 1) / 2) * 3.1, (gz - (grid - 1) / 2) * 3.1
        tid += 1
        trans.append(f'    <Translation id="{tid}">{_fmt((px, 0.0, pz))}</Translation>')
        tl = f"t{tid}"
        if k % 3 == 1:
            scal.append(f'    <Scaling id="{len(scal) + 1}">{_fmt((0.8 + 0.05 * (k % 4), 1.0 + 0.04 * (k % 5), 0.85))}</Scaling>')
            tl = f"s{len(scal)} " + tl
        if k % 4 == 2:
            rots.append(f'    <Rotation id="{len(rots) + 1}">{_fmt((17.0 * k, 0.2, 1.0, 0.1))}</Rotation>')
            tl = f"r{len(rots)} " + tl
        blur = f"<MotionBlur>{_fmt((0.0, 0.25, 0.12))}</MotionBlur>" if k % 5 == 4 else ""
        mat = 1 + (k % 2) * 2 if k % 7 else 4
        if k == 0:
            base = (f'    <Mesh id="1" shadingMode="smooth">\n      <Material>{mat}</Material>\n      <Textures>1</Textures>\n'
                    f'      <Transformations>{tl}</Transformations>\n      <Faces plyFile="{name}_mesh.ply" />\n    </Mesh>')
        else:
            inst_xml.append(f'    <MeshInstance id="{100 + k}" baseMeshId="1" resetTransform="true">\n      <Material>{mat}</Material>\n'
                            f'      <Transformations>{tl}</Transformations>\n      {blur}\n    </MeshInstance>')
    xml = f"""<Scene>
  <MaxRecursionDepth>2</MaxRecursionDepth>
  <BackgroundColor>14 18 28</BackgroundColor>
  <ShadowRayEpsilon>1e-3</ShadowRayEpsilon>
  <IntersectionTestEpsilon>1e-6</IntersectionTestEpsilon>
  <Cameras>
    <Camera id="1">
      <Position>0.7 7.3 13.9</Position>
      <GazePoint>0.05 -0.4 0.1</GazePoint>
      <Up>0 1 0</Up>
      <FovY>42</FovY>
      <NearDistance>1</NearDistance>
      <ImageResolution>{res[0]} {res[1]}</ImageResolution>
      <NumSamples>{spp}</NumSamples>
      <ImageName>{name}.ppm</ImageName>
    </Camera>
  </Cameras>
  <Lights>
    <AmbientLight>16 16 16</AmbientLight>
    <PointLight id="1">
      <Position>6 14 9</Position>
      <Intensity>52000 50000 47000</Intensity>
    </PointLight>
{area}  </Lights>
  <Materials>
{mats}  </Materials>
  <Textures>
    <Images>
      <Image id="1">{name}_floor.exr</Image>
    </Images>
    <TextureMap id="1" type="perlin">
      <DecalMode>bump_normal</DecalMode>
      <NoiseConversion>absval</NoiseConversion>
      <NoiseScale>6</NoiseScale>
      <BumpFactor>0.3</BumpFactor>
    </TextureMap>
    <TextureMap id="2" type="image">
      <ImageId>1</ImageId>
      <DecalMode>blend_kd</DecalMode>
      <Interpolation>bilinear</Interpolation>
      <Normalizer>255</Normalizer>
    </TextureMap>
    <TextureMap id="3" type="perlin">
      <DecalMode>replace_kd</DecalMode>
      <NoiseConversion>linear</NoiseConversion>
      <NoiseScale>0.9</NoiseScale>
    </TextureMap>
  </Textures>
  <Transformations>
{chr(10).join(trans)}
{chr(10).join(scal)}
{chr(10).join(rots)}
  </Transformations>
  <VertexData>
{chr(10).join("    " + _fmt(p) for p in verts)}
  </VertexData>
  <TexCoordData>
{chr(10).join("    " + _fmt(p) for p in uvs)}
  </TexCoordData>
  <Objects>
    <Mesh id="50">
      <Material>2</Material>
      <Textures>2</Textures>
      <Faces>
{_faces_text(ground)}
      </Faces>
    </Mesh>
    <Mesh id="51">
      <Material>5</Material>
      <Textures>3</Textures>
      <Faces>
{_faces_text(wall)}
      </Faces>
    </Mesh>
{base}
{chr(10).join(inst_xml)}
  </Objects>
</Scene>
"""
    path = os.path.join(outdir, f"{name}.xml")
    with open(path, "w") as fh:
        fh.write(xml)
    return path


def sky_exr(path: str, w: int = 1024, h: int = 512, sun_scale: float = 1.0) -> None:
    """Procedural lat-long HDR environment: sky gradient + ground + sun disc (float RGBA EXR)."""
    v = (np.arange(h) + 0.5) / h
    u = (np.arange(w) + 0.5) / w
    theta = v[:, None] * math.pi                  # 0 = +y
    phi = math.pi - u[None, :] * 2 * math.pi      # inverse of u = (pi - phi) / 2pi
    dy = np.cos(theta) * np.ones_like(phi)
    dx = np.sin(theta) * np.cos(phi)
    dz = np.sin(theta) * np.sin(phi)
    up = np.clip(dy, 0, 1)
    img = np.empty((h, w, 4), dtype=np.float32)
    sky = np.stack([40 + 60 * (1 - up), 70 + 70 * (1 - up), 140 + 60 * up], axis=-1)
    ground = np.array([38.0, 34.0, 30.0])
    img[..., :3] = np.where((dy > 0)[..., None], sky, ground)
    sun = np.array([0.45, 0.62, 0.64]); sun = sun / np.linalg.norm(sun)
    cosang = dx * sun[0] + dy * sun[1] + dz * sun[2]
    disc = np.clip((cosang - 0.995) / 0.005, 0, 1) ** 2
    img[..., :3] += disc[..., None] * np.array([24000.0, 22000.0, 18000.0]) * sun_scale
    img[..., 3] = 1.0
    write_exr(path, img)


def scene_c5(outdir: str, res=(3840, 2160), spp: int = 1024, name: str = "c5", depth: int = 8, sun_scale: float = 1.0) -> str:
    """C5: Cornell box (open front) with BRDF materials, HDR environment light + emissive quad (object light, with a
    co-located AreaLight for NEE / for the reference's direct lighting), meant for the path tracer + tone map."""
    os.makedirs(outdir, exist_ok=True)
    sky_exr(os.path.join(outdir, f"{name}_sky.exr"), sun_scale=sun_scale)
    X, Y, Z0, Z1 = 10.0, 20.0, -10.0, 10.0
    verts = [(-X, 0, Z1), (X, 0, Z1), (X, 0, Z0), (-X, 0, Z0), (-X, Y, Z1), (X, Y, Z1), (X, Y, Z0), (-X, Y, Z0),
             (-2.5, 19.95, 2.5), (2.5, 19.95, 2.5), (2.5, 19.95, -2.5), (-2.5, 19.95, -2.5),   # emissive quad 8-11
             (-4.5, 4.0, 1.0), (4.5, 3.0, 3.5)]                                                  # spheres 12, 13
    white = (_quad_faces(verts, (0, 1, 2, 3), (0, 1, 0)) + _quad_faces(verts, (4, 5, 6, 7), (0, -1, 0)) + _quad_faces(verts, (3, 2, 6, 7), (0, 0, 1)))
    red = _quad_faces(verts, (0, 3, 7, 4), (1, 0, 0))
    green = _quad_faces(verts, (1, 2, 6, 5), (-1, 0, 0))
    lamp = _quad_faces(verts, (8, 9, 10, 11), (0, -1, 0))
    mats = (_material(1, (0, 0, 0), (0.75, 0.75, 0.75), (0.15, 0.15, 0.15), 1, brdf=1) + _material(2, (0, 0, 0), (0.75, 0.1, 0.08), (0.1, 0.1, 0.1), 1, brdf=1) +
            _material(3, (0, 0, 0), (0.1, 0.65, 0.12), (0.1, 0.1, 0.1), 1, brdf=1) +
            _material(4, (0, 0, 0), (0.35, 0.3, 0.2), (0.6, 0.55, 0.45), 1, brdf=2, extra="      <RefractionIndex>1.8</RefractionIndex>\n      <AbsorptionIndex>2.2</AbsorptionIndex>\n") +
            _material(5, (0, 0, 0), (0, 0, 0), (0, 0, 0), 1, "mirror", (0.92, 0.92, 0.92)) +
            _material(6, (0, 0, 0), (0, 0, 0), (0, 0, 0), 1, extra="      <Emission>1800 1700 1500</Emission>\n"))
    xml = f"""<Scene>
  <MaxRecursionDepth>{depth}</MaxRecursionDepth>
  <BackgroundColor>0 0 0</BackgroundColor>
  <ShadowRayEpsilon>1e-3</ShadowRayEpsilon>
  <IntersectionTestEpsilon>1e-6</IntersectionTestEpsilon>
  <Cameras>
    <Camera id="1">
      <Position>0.4 10.2 41.5</Position>
      <Gaze>-0.011 -0.017 -1</Gaze>
      <Up>0 1 0</Up>
      <FovY>36</FovY>
      <NearDistance>1</NearDistance>
      <ImageResolution>{res[0]} {res[1]}</ImageResolution>
      <NumSamples>{spp}</NumSamples>
      <ImageName>{name}.exr</ImageName>
    </Camera>
  </Cameras>
  <BRDFs>
    <ModifiedBlinnPhong id="1" normalized="true"><Exponent>40</Exponent></ModifiedBlinnPhong>
    <TorranceSparrow id="2" kdfresnel="true"><Exponent>60</Exponent></TorranceSparrow>
  </BRDFs>
  <Lights>
    <AmbientLight>0 0 0</AmbientLight>
    <AreaLight id="1">
      <Position>0 19.9 0</Position>
      <Normal>0 -1 0</Normal>
      <Size>5</Size>
      <Radiance>1800 1700 1500</Radiance>
    </AreaLight>
    <SphericalDirectionalLight id="2">
      <ImageId>1</ImageId>
    </SphericalDirectionalLight>
  </Lights>
  <Materials>
{mats}  </Materials>
  <Textures>
    <Images>
      <Image id="1">{name}_sky.exr</Image>
    </Images>
  </Textures>
  <VertexData>
{chr(10).join("    " + _fmt(p) for p in verts)}
  </VertexData>
  <Objects>
    <Sphere id="1"><Material>4</Material><Center>13</Center><Radius>4.0</Radius></Sphere>
    <Sphere id="2"><Material>5</Material><Center>14</Center><Radius>3.0</Radius></Sphere>
    <Mesh id="3"><Material>1</Material><Faces>
{_faces_text(white)}
    </Faces></Mesh>
    <Mesh id="4"><Material>2</Material><Faces>
{_faces_text(red)}
    </Faces></Mesh>
    <Mesh id="5"><Material>3</Material><Faces>
{_faces_text(green)}
    </Faces></Mesh>
    <Mesh id="6"><Material>6</Material><Faces>
{_faces_text(lamp)}
    </Faces></Mesh>
  </Objects>
</Scene>
"""
    path = os.path.join(outdir, f"{name}.xml")
    with open(path, "w") as fh:
        fh.write(xml)
    return path


def scene_object_lights(outdir: str, variant: str = "cornell", res=(160, 120), spp: int = 16, name: str = "olights", depth: int = 6) -> str:
    """Object lights (<LightMesh> / <LightSphere>, the course's HW7 elements; extensions of the reference grammar), for the path tracer.
    variant "cornell": closed Cornell box, a <LightMesh> lamp under the ceiling (rotated and translated by object transformations),
        a scaled + translated <LightSphere>, a diffuse and a mirror sphere; every surface a normalised BRDF (kd / pi).
    variant "furnace": the inside of a closed box that IS the only light: <LightMesh> of 12 triangles facing inwards, radiance 1,
        albedo 0.5 (normalised modified Blinn-Phong with ks = 0, so f = kd / pi exactly).  Every ray sees radiance
        sum_k 0.5^k over the bounces the depth allows, with next-event estimation on or off - the estimator's pdf / area /
        cosine factors all have to be right for that."""
    os.makedirs(outdir, exist_ok=True)
    if variant == "furnace":
        L = 4.0
        verts = [(-L, -L, L), (L, -L, L), (L, -L, -L), (-L, -L, -L), (-L, L, L), (L, L, L), (L, L, -L), (-L, L, -L)]
        faces = (_quad_faces(verts, (0, 1, 2, 3), (0, 1, 0)) + _quad_faces(verts, (4, 5, 6, 7), (0, -1, 0)) + _quad_faces(verts, (3, 2, 6, 7), (0, 0, 1)) +
                 _quad_faces(verts, (0, 1, 5, 4), (0, 0, -1)) + _quad_faces(verts, (0, 3, 7, 4), (1, 0, 0)) + _quad_faces(verts, (1, 2, 6, 5), (-1, 0, 0)))
        mats = _material(1, (0, 0, 0), (0.5, 0.5, 0.5), (0, 0, 0), 1, brdf=1)
        cam = "<Position>0.3 -0.4 0.9</Position><Gaze>0.2 0.1 -1</Gaze><Up>0 1 0</Up><FovY>70</FovY>"
        objects = f"""    <LightMesh id="1"><Material>1</Material><Radiance>1 1 1</Radiance><Faces>
{_faces_text(faces)}
    </Faces></LightMesh>"""
        lights = "    <AmbientLight>0 0 0</AmbientLight>"
        transforms = ""
    else:
        X, Y, Z0, Z1 = 10.0, 20.0, -10.0, 10.0
        verts = [(-X, 0, Z1), (X, 0, Z1), (X, 0, Z0), (-X, 0, Z0), (-X, Y, Z1), (X, Y, Z1), (X, Y, Z0), (-X, Y, Z0),
                 (-1, 0, 1), (1, 0, 1), (1, 0, -1), (-1, 0, -1),          # unit lamp quad 8-11 (scaled / rotated / moved by transformations)
                 (-4.5, 4.0, 1.0), (4.5, 3.0, 3.5), (0, 0, 0)]             # spheres 12, 13; light sphere centre 14 (moved by a transformation)
        white = (_quad_faces(verts, (0, 1, 2, 3), (0, 1, 0)) + _quad_faces(verts, (4, 5, 6, 7), (0, -1, 0)) + _quad_faces(verts, (3, 2, 6, 7), (0, 0, 1)) +
                 _quad_faces(verts, (0, 1, 5, 4), (0, 0, -1)))
        red = _quad_faces(verts, (0, 3, 7, 4), (1, 0, 0))
        green = _quad_faces(verts, (1, 2, 6, 5), (-1, 0, 0))
        lamp = _quad_faces(verts, (8, 9, 10, 11), (0, -1, 0))
        mats = (_material(1, (0, 0, 0), (0.75, 0.75, 0.75), (0.1, 0.1, 0.1), 1, brdf=1) + _material(2, (0, 0, 0), (0.75, 0.1, 0.08), (0.05, 0.05, 0.05), 1, brdf=1) +
                _material(3, (0, 0, 0), (0.1, 0.65, 0.12), (0.05, 0.05, 0.05), 1, brdf=1) + _material(4, (0, 0, 0), (0.6, 0.6, 0.3), (0.2, 0.2, 0.2), 1, brdf=1) +
                _material(5, (0, 0, 0), (0, 0, 0), (0, 0, 0), 1, "mirror", (0.9, 0.9, 0.9)) + _material(6, (0, 0, 0), (0.3, 0.3, 0.3), (0, 0, 0), 1, brdf=1))
        cam = "<Position>0.4 10.2 9.5</Position><Gaze>-0.011 -0.017 -1</Gaze><Up>0 1 0</Up><FovY>75</FovY>"
        objects = f"""    <Sphere id="1"><Material>4</Material><Center>13</Center><Radius>4.0</Radius></Sphere>
    <Sphere id="2"><Material>5</Material><Center>14</Center><Radius>3.0</Radius></Sphere>
    <LightSphere id="3"><Material>6</Material><Center>15</Center><Radius>1.0</Radius><Transformations>s2 t2</Transformations><Radiance>90 80 60</Radiance></LightSphere>
    <Mesh id="4"><Material>1</Material><Faces>
{_faces_text(white)}
    </Faces></Mesh>
    <Mesh id="5"><Material>2</Material><Faces>
{_faces_text(red)}
    </Faces></Mesh>
    <Mesh id="6"><Material>3</Material><Faces>
{_faces_text(green)}
    </Faces></Mesh>
    <LightMesh id="7"><Material>6</Material><Transformations>s1 r1 t1</Transformations><Radiance>60 56 50</Radiance><Faces>
{_faces_text(lamp)}
    </Faces></LightMesh>"""
        lights = "    <AmbientLight>0 0 0</AmbientLight>"
        transforms = """  <Transformations>
    <Scaling id="1">3 1 2.5</Scaling>
    <Scaling id="2">1.5 1.5 1.5</Scaling>
    <Rotation id="1">12 0 0 1</Rotation>
    <Translation id="1">-1 19.2 0.5</Translation>
    <Translation id="2">6.5 12 -5</Translation>
  </Transformations>
"""
    xml = f"""<Scene>
  <MaxRecursionDepth>{depth}</MaxRecursionDepth>
  <BackgroundColor>0 0 0</BackgroundColor>
  <ShadowRayEpsilon>1e-3</ShadowRayEpsilon>
  <IntersectionTestEpsilon>1e-6</IntersectionTestEpsilon>
  <Cameras>
    <Camera id="1">
      {cam}
      <NearDistance>1</NearDistance>
      <ImageResolution>{res[0]} {res[1]}</ImageResolution>
      <NumSamples>{spp}</NumSamples>
      <ImageName>{name}.exr</ImageName>
      <Renderer>PathTracing</Renderer>
      <RendererParams>NextEventEstimation ImportanceSampling</RendererParams>
    </Camera>
  </Cameras>
  <BRDFs>
    <ModifiedBlinnPhong id="1" normalized="true"><Exponent>20</Exponent></ModifiedBlinnPhong>
  </BRDFs>
  <Lights>
{lights}
  </Lights>
  <Materials>
{mats}  </Materials>
{transforms}  <VertexData>
{chr(10).join("    " + _fmt(p) for p in verts)}
  </VertexData>
  <Objects>
{objects}
  </Objects>
</Scene>
"""
    path = os.path.join(outdir, f"{name}.xml")
    with open(path, "w") as fh:
        fh.write(xml)
    return path


CONFIGS = {
    "c1": dict(gen=scene_c1, label="C1 Whitted: 5 spheres + 4 triangles, 1 point light, 800x800, 1 spp, depth 6"),
    "c2": dict(gen=scene_c2, label="C2 Cornell box, dielectric + conductor spheres, area light, glossy, 1024x1024, 100 spp"),
    "c3": dict(gen=scene_c3, label="C3 ~69k-triangle smooth PLY mesh, point + area light, 1920x1080, 16 spp"),
    "c4": dict(gen=scene_c4, label="C4 ~870k-triangle mesh x16 (TLAS/BLAS), transforms + motion blur, Perlin bump/kd + image texture, 1920x1080, 64 spp"),
    "c5": dict(gen=scene_c5, label="C5 path-traced Cornell box, BRDF materials, NEE + RR, HDR env light + tone map, 3840x2160, 1024 spp"),
}


# ------------------------------------------------------------------------------------------ feature-matrix scenes (parity coverage)
def _rgb_test_image(w: int = 64, h: int = 48, seed: int = 3) -> np.ndarray:
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    return np.stack([(xx * 255 // max(1, w - 1)), (yy * 255 // max(1, h - 1)), ((xx // 8 + yy // 8) % 2) * 200 + rng.integers(0, 40, size=(h, w))], axis=-1).astype(np.uint8)


def _write_jpeg(path: str, img: np.ndarray, **kw) -> None:
    """JPEG texture through Pillow (test scenes only)."""
    from PIL import Image
    Image.fromarray(img).save(path, format="JPEG", **kw)


def _write_png_rgb(path: str, w: int = 64, h: int = 48, seed: int = 3) -> None:
    """Small RGB8 PNG (non-interlaced) written with zlib only, readable by the oracle's png shim and the host shell."""
    import zlib
    img = _rgb_test_image(w, h, seed)
    raw = b"".join(b"\x00" + img[y].tobytes() for y in range(h))

    def chunk(t, d):
        c = struct.pack(">I", len(d)) + t + d
        return c + struct.pack(">I", zlib.crc32(t + d) & 0xffffffff)
    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0)) + chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b""))


def scene_features(outdir: str, variant: str = "lights", res=(192, 144), spp: int = 1, name: str = "feat") -> str:
    """Small scenes that exercise every branch of the hot path the reference implements:
      lights    point + directional + spot lights, Blinn-Phong, mirror / conductor / dielectric recursion, transforms (deterministic)
      brdf      the eight BRDF models (Light.cpp:62-155) under point + directional light (deterministic)
      textures  PNG + EXR image textures (replace_kd / blend_kd / replace_all / replace_normal / bump_normal, nearest + bilinear),
                Perlin kd + bump on spheres and meshes, replace_background, non-uniform scale + rotation (deterministic)
      stochastic  depth of field, motion blur, glossy mirror, area light, non-square sample count
      env       environment (SphericalDirectionalLight) lighting + left-handed camera
    """
    os.makedirs(outdir, exist_ok=True)
    W, H = res
    cam_extra, lights, textures, objects, transformations, brdfs = "", "", "", "", "", ""
    background = "9 11 15"
    handed = ""
    verts = [(-8, 0, 8), (8, 0, 8), (8, 0, -8), (-8, 0, -8),          # floor 1-4
             (-8, 10, 8), (8, 10, 8), (8, 10, -8), (-8, 10, -8),      # ceiling 5-8
             (-4.0, 1.5, 0.0), (0.0, 1.5, -1.0), (4.0, 1.5, 0.5), (-2.0, 1.0, 3.5), (2.2, 1.0, 3.2), (0.0, 4.5, -3.0)]   # centres 9-14
    uvs = [(0, 0), (3, 0), (3, 3), (0, 3), (0, 0), (1, 0), (1, 1), (0, 1)] + [(0, 0)] * 6
    room = (_quad_faces(verts, (0, 1, 2, 3), (0, 1, 0)) + _quad_faces(verts, (4, 5, 6, 7), (0, -1, 0)) + _quad_faces(verts, (3, 2, 6, 7), (0, 0, 1)) +
            _quad_faces(verts, (0, 1, 5, 4), (0, 0, -1)) + _quad_faces(verts, (0, 3, 7, 4), (1, 0, 0)) + _quad_faces(verts, (1, 2, 6, 5), (-1, 0, 0)))
    floor_only = _quad_faces(verts, (0, 1, 2, 3), (0, 1, 0))
    mats = _material(1, (1, 1, 1), (0.7, 0.7, 0.7), (0.2, 0.2, 0.2), 10)
    sphere = lambda i, m, c, r, extra="": f'    <Sphere id="{i}"><Material>{m}</Material><Center>{c}</Center><Radius>{r}</Radius>{extra}</Sphere>\n'
    room_mesh = lambda m, extra="": f'    <Mesh id="90"><Material>{m}</Material>{extra}<Faces>\n{_faces_text(room)}\n    </Faces></Mesh>\n'
    floor_mesh = lambda m, extra="": f'    <Mesh id="91"><Material>{m}</Material>{extra}<Faces>\n{_faces_text(floor_only)}\n    </Faces></Mesh>\n'
    depth = 4
    if variant == "lights":
        lights = ('<AmbientLight>12 12 12</AmbientLight>\n<PointLight id="1"><Position>-3 8 5</Position><Intensity>2500 2400 2300</Intensity></PointLight>\n'
                  '<DirectionalLight id="1"><Direction>0.3 -1 -0.4</Direction><Radiance>40 42 50</Radiance></DirectionalLight>\n'
                  '<SpotLight id="1"><Position>4 9 4</Position><Direction>-0.3 -1 -0.35</Direction><Intensity>6000 6000 5000</Intensity>'
                  '<CoverageAngle>50</CoverageAngle><FalloffAngle>30</FalloffAngle></SpotLight>')
        mats += (_material(2, (1, 1, 1), (0.1, 0.1, 0.1), (0.5, 0.5, 0.5), 80, "mirror", (0.8, 0.8, 0.8)) +
                 _material(3, (0, 0, 0), (0.05, 0.05, 0.05), (0.3, 0.3, 0.3), 40, "conductor", (0.9, 0.7, 0.4), "      <RefractionIndex>0.37</RefractionIndex>\n      <AbsorptionIndex>2.82</AbsorptionIndex>\n") +
                 _material(4, (0, 0, 0), (0, 0, 0), (0, 0, 0), 1, "dielectric", None, "      <RefractionIndex>1.5</RefractionIndex>\n      <AbsorptionCoefficient>0.02 0.05 0.03</AbsorptionCoefficient>\n") +
                 _material(5, (1, 1, 1), (0.2, 0.5, 0.8), (0.6, 0.6, 0.6), 30))
        transformations = ('<Transformations><Translation id="1">0 0.5 0</Translation><Scaling id="1">1 1.4 0.8</Scaling><Rotation id="1">35 0.2 1 0.1</Rotation>'
                           '<Composite id="1">0.9 0 0.1 1.0  0 1.1 0 0.3  -0.1 0 0.9 0.2  0 0 0 1</Composite></Transformations>')
        objects = (sphere(1, 2, 10, 1.5) + sphere(2, 3, 11, 1.5) + sphere(3, 4, 9, 1.5) + sphere(4, 5, 12, 1.0, "<Transformations>s1 r1 t1</Transformations>") +
                   sphere(5, 5, 13, 1.0, "<Transformations>c1</Transformations>") +
                   '    <Triangle id="6"><Material>2</Material><Indices>14 11 10</Indices></Triangle>\n' + room_mesh(1))
    elif variant == "brdf":
        depth = 1
        lights = ('<AmbientLight>5 5 5</AmbientLight>\n<PointLight id="1"><Position>0 9 6</Position><Intensity>3000 3000 3000</Intensity></PointLight>\n'
                  '<DirectionalLight id="1"><Direction>-0.4 -1 -0.6</Direction><Radiance>60 60 60</Radiance></DirectionalLight>')
        brdfs = ('<BRDFs><OriginalBlinnPhong id="1"><Exponent>30</Exponent></OriginalBlinnPhong><ModifiedBlinnPhong id="2"><Exponent>30</Exponent></ModifiedBlinnPhong>'
                 '<ModifiedBlinnPhong id="3" normalized="true"><Exponent>30</Exponent></ModifiedBlinnPhong><OriginalPhong id="4"><Exponent>20</Exponent></OriginalPhong>'
                 '<ModifiedPhong id="5"><Exponent>20</Exponent></ModifiedPhong><ModifiedPhong id="6" normalized="true"><Exponent>20</Exponent></ModifiedPhong>'
                 '<TorranceSparrow id="7"><Exponent>50</Exponent></TorranceSparrow><TorranceSparrow id="8" kdfresnel="true"><Exponent>50</Exponent></TorranceSparrow></BRDFs>')
        cx = [(-5.2 + 1.5 * k, 1.0 + 2.2 * (k % 2), 0.5 * (k % 3)) for k in range(8)]
        verts = verts[:8] + cx
        uvs = uvs[:8] + [(0, 0)] * 8
        for k in range(8):
            mats += _material(2 + k, (1, 1, 1), (0.5, 0.3 + 0.05 * k, 0.2), (0.8, 0.8, 0.8), 1, brdf=k + 1,
                              extra="      <RefractionIndex>1.6</RefractionIndex>\n      <AbsorptionIndex>2.5</AbsorptionIndex>\n")
            objects += sphere(k + 1, 2 + k, 9 + k, 0.7)
        objects += floor_mesh(1)
    elif variant in ("textures", "textures_jpeg"):
        depth = 2
        e1, e3 = "png", "png"
        if variant == "textures_jpeg":
            # the same scene with its 8-bit images as JPEG (src/Texture.cpp:191-240): baseline 4:2:0, progressive 4:4:4 and a greyscale
            # file, which the reference reads with numComponents = 1 (pixel + its two right-hand neighbours as R, G, B, :62-66)
            e1, e3 = "jpg", "jpg"
            _write_jpeg(os.path.join(outdir, f"{name}_t.jpg"), _rgb_test_image(), quality=92, subsampling=2)
            _write_jpeg(os.path.join(outdir, f"{name}_bg.jpg"), _rgb_test_image(80, 60, seed=9), quality=85, subsampling=0, progressive=True)
            _write_jpeg(os.path.join(outdir, f"{name}_g.jpg"), _rgb_test_image(72, 40, seed=5)[..., 2], quality=90)
        else:
            _write_png_rgb(os.path.join(outdir, f"{name}_t.png"))
            _write_png_rgb(os.path.join(outdir, f"{name}_bg.png"), 80, 60, seed=9)
        checker_exr(os.path.join(outdir, f"{name}_t.exr"), 96, 64)
        lights = ('<AmbientLight>20 20 20</AmbientLight>\n<PointLight id="1"><Position>1 9 7</Position><Intensity>4200 4200 4200</Intensity></PointLight>')
        maps = [("image", 1, "replace_kd", "nearest", 255, None, None, None), ("image", 2, "blend_kd", "bilinear", 255, None, None, None),
                ("image", 1, "replace_all", "bilinear", 1, None, None, None), ("image", 1, "replace_normal", "bilinear", 255, None, None, None),
                ("image", 2, "bump_normal", "nearest", 255, None, None, 0.02), ("perlin", 0, "replace_kd", "nearest", 255, "linear", 2.5, 1.0),
                ("perlin", 0, "bump_normal", "nearest", 255, "absval", 3.0, 0.4), ("image", 3, "replace_background", "bilinear", 255, None, None, None),
                ("image", 2, "replace_kd", "bilinear", 255, None, None, None)]
        tm = ""
        for k, (ty, im, dm, ip, nz, nc, ns, bf) in enumerate(maps):
            tm += f'<TextureMap id="{k + 1}" type="{ty}">' + (f"<ImageId>{im}</ImageId>" if ty == "image" else "") + f"<DecalMode>{dm}</DecalMode><Interpolation>{ip}</Interpolation><Normalizer>{nz}</Normalizer>"
            tm += (f"<NoiseConversion>{nc}</NoiseConversion>" if nc else "") + (f"<NoiseScale>{ns}</NoiseScale>" if ns else "") + (f"<BumpFactor>{bf}</BumpFactor>" if bf else "") + "</TextureMap>\n"
        if variant == "textures_jpeg":
            tm += '<TextureMap id="10" type="image"><ImageId>4</ImageId><DecalMode>replace_kd</DecalMode><Interpolation>bilinear</Interpolation><Normalizer>255</Normalizer></TextureMap>\n'
        img4 = f'<Image id="4">{name}_g.jpg</Image>' if variant == "textures_jpeg" else ""
        textures = f'<Textures><Images><Image id="1">{name}_t.{e1}</Image><Image id="2">{name}_t.exr</Image><Image id="3">{name}_bg.{e3}</Image>{img4}</Images>\n{tm}</Textures>'
        transformations = '<Transformations><Scaling id="1">1.3 0.8 1</Scaling><Rotation id="1">40 0 1 0.3</Rotation><Translation id="1">0 0.4 0</Translation></Transformations>'
        tex = lambda ids: f"<Textures>{ids}</Textures>"
        objects = (sphere(1, 1, 9, 1.3, tex("1")) + sphere(2, 1, 10, 1.3, tex("2") + "<Transformations>s1 r1 t1</Transformations>") + sphere(3, 1, 11, 1.3, tex("4")) +
                   sphere(4, 1, 12, 1.0, tex("5 6")) + sphere(5, 1, 13, 1.0, tex("10" if variant == "textures_jpeg" else "7")) + sphere(6, 1, 14, 1.2, tex("3")) + floor_mesh(1, tex("9 7")))
        # a textured quad mesh with its own uv: replace_normal + bump from image on triangles
        verts += [(-7.5, 0.2, -6), (-2.5, 0.2, -6), (-2.5, 5.2, -7), (-7.5, 5.2, -7), (2.5, 0.2, -6), (7.5, 0.2, -6), (7.5, 5.2, -7), (2.5, 5.2, -7)]
        uvs += [(0, 0), (1, 0), (1, 1), (0, 1), (0, 0), (2, 0), (2, 2), (0, 2)]
        q1 = _quad_faces(verts, (14, 15, 16, 17), (0, 0.2, 1)); q2 = _quad_faces(verts, (18, 19, 20, 21), (0, 0.2, 1))
        objects += f'    <Mesh id="92"><Material>1</Material>{tex("4")}<Faces>\n{_faces_text(q1)}\n    </Faces></Mesh>\n'
        objects += f'    <Mesh id="93" shadingMode="smooth"><Material>1</Material>{tex("5 1")}<Faces>\n{_faces_text(q2)}\n    </Faces></Mesh>\n'
    elif variant == "stochastic":
        cam_extra = f"<NumSamples>{max(spp, 5)}</NumSamples><FocusDistance>14</FocusDistance><ApertureSize>0.35</ApertureSize>"
        lights = ('<AmbientLight>10 10 10</AmbientLight>\n<AreaLight id="1"><Position>0 9.2 2</Position><Normal>0 -1 0</Normal><Size>3</Size><Radiance>900 900 850</Radiance></AreaLight>\n'
                  '<PointLight id="1"><Position>-5 6 6</Position><Intensity>900 900 1000</Intensity></PointLight>')
        mats += (_material(2, (1, 1, 1), (0.1, 0.1, 0.1), (0.5, 0.5, 0.5), 80, "mirror", (0.8, 0.8, 0.8), "      <Roughness>0.2</Roughness>\n") + _material(3, (1, 1, 1), (0.8, 0.3, 0.2), (0.4, 0.4, 0.4), 20))
        objects = (sphere(1, 2, 10, 1.5) + sphere(2, 3, 9, 1.5, "<MotionBlur>0.6 0.9 0</MotionBlur>") + sphere(3, 3, 11, 1.5) + room_mesh(1))
    elif variant == "stochastic_blur":   # same without depth of field: the reference applies motion blur only when DoF is off (Camera.cpp:105-112)
        cam_extra = f"<NumSamples>{max(spp, 4)}</NumSamples>"
        lights = ('<AmbientLight>10 10 10</AmbientLight>\n<AreaLight id="1"><Position>0 9.2 2</Position><Normal>0 -1 0</Normal><Size>3</Size><Radiance>900 900 850</Radiance></AreaLight>')
        mats += _material(2, (1, 1, 1), (0.8, 0.3, 0.2), (0.4, 0.4, 0.4), 20)
        objects = (sphere(1, 2, 10, 1.5, "<MotionBlur>1.2 0.6 0</MotionBlur>") + sphere(2, 2, 9, 1.2) + room_mesh(1))
    elif variant == "env":
        handed = ' handedness="left"'
        sky_exr(os.path.join(outdir, f"{name}_sky.exr"), 128, 64, sun_scale=0.002)
        textures = f'<Textures><Images><Image id="1">{name}_sky.exr</Image></Images></Textures>'
        cam_extra = f"<NumSamples>{max(spp, 4)}</NumSamples>"
        lights = '<AmbientLight>0 0 0</AmbientLight>\n<SphericalDirectionalLight id="1"><ImageId>1</ImageId></SphericalDirectionalLight>'
        mats += _material(2, (0, 0, 0), (0.6, 0.4, 0.3), (0.5, 0.5, 0.5), 30)
        objects = sphere(1, 2, 10, 1.5) + sphere(2, 1, 9, 1.5) + floor_mesh(1)
    else:
        raise ValueError(variant)
    xml = f"""<Scene>
  <MaxRecursionDepth>{depth}</MaxRecursionDepth>
  <BackgroundColor>{background}</BackgroundColor>
  <ShadowRayEpsilon>1e-3</ShadowRayEpsilon>
  <IntersectionTestEpsilon>1e-6</IntersectionTestEpsilon>
  <Cameras>
    <Camera id="1"{handed}>
      <Position>0.23 4.1 7.7</Position>
      <GazePoint>0.02 1.9 0.03</GazePoint>
      <Up>0 1 0</Up>
      <FovY>58</FovY>
      <NearDistance>1</NearDistance>
      <ImageResolution>{W} {H}</ImageResolution>
      <ImageName>{name}_{variant}.ppm</ImageName>
      {cam_extra}
    </Camera>
  </Cameras>
  {brdfs}
  <Lights>
{lights}
  </Lights>
  <Materials>
{mats}  </Materials>
  {textures}
  {transformations}
  <VertexData>
{chr(10).join("    " + _fmt(p) for p in verts)}
  </VertexData>
  <TexCoordData>
{chr(10).join("    " + _fmt(p) for p in uvs)}
  </TexCoordData>
  <Objects>
{objects}  </Objects>
</Scene>
"""
    path = os.path.join(outdir, f"{name}_{variant}.xml")
    with open(path, "w") as fh:
        fh.write(xml)
    return path


FEATURE_VARIANTS_DETERMINISTIC = ["lights", "brdf", "textures"]
FEATURE_VARIANTS_STOCHASTIC = ["stochastic", "stochastic_blur", "env"]

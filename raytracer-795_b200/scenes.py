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

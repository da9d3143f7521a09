"""Device BLAS builder (csrc/bvh_device.cu) against the host binned-SAH builder.

A BVH is correct when it is a conservative partition: every triangle sits in exactly one leaf and every child box contains
everything below it - then traversal returns the same closest hit as any other conservative BVH, in
particular the reference's median-split tree (src/BVH.cpp:64-135).  Checked twice: structurally, by walking the arrays
read back through rt_debug_read_bvh, and end to end, by rendering the same scene with both builders."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

K_NO_CHILD = 0x7FFFFFFD


def _scenes(rt, xml, builders):
    hs = rt.HostScene(xml)
    return hs, [rt.Scene(hs.desc, 0, bvh_builder=b) for b in builders]


def _walk(nodes, root, width, tris, ref_col, box_of_child):
    """Walks one BLAS from `root`; returns (per-triangle cover counts, max depth, node count).  Asserts containment."""
    refs = nodes[:, ref_col:ref_col + width].view(np.int32)
    fl = nodes.view(np.float32)
    ntri = tris.shape[0]
    cover = np.zeros(ntri, dtype=np.int32)
    t = tris.view(np.float32).reshape(ntri, 3, 4)
    v0 = t[:, 0, :3]
    v1 = v0 + t[:, 1, :3]
    v2 = v0 + t[:, 2, :3]
    tlo = np.minimum(np.minimum(v0, v1), v2)
    thi = np.maximum(np.maximum(v0, v1), v2)
    # post-order: box of everything below each visited child, compared with the box stored for it
    stack = [(root, 0, None)]
    max_depth, visited = 0, 0
    sub_lo, sub_hi = {}, {}
    order = []
    while stack:
        n, depth, _ = stack.pop()
        visited += 1
        max_depth = max(max_depth, depth)
        order.append(n)
        for k in range(width):
            r = int(refs[n, k])
            if r == K_NO_CHILD:
                continue
            if r >= 0:
                stack.append((r, depth + 1, None))
    for n in reversed(order):
        lo = np.full(3, np.inf, dtype=np.float32)
        hi = np.full(3, -np.inf, dtype=np.float32)
        for k in range(width):
            r = int(refs[n, k])
            if r == K_NO_CHILD:
                continue
            blo, bhi = box_of_child(fl[n], k)
            if r >= 0:
                clo, chi = sub_lo[r], sub_hi[r]
            else:
                payload = ~r
                first, count = payload >> 3, (payload & 7) + 1
                assert 1 <= count <= 4 and first + count <= ntri
                cover[first:first + count] += 1
                clo, chi = tlo[first:first + count].min(axis=0), thi[first:first + count].max(axis=0)
            assert (blo <= clo).all() and (bhi >= chi).all(), f"node {n} child {k}: stored box does not contain its subtree"
            lo, hi = np.minimum(lo, clo), np.maximum(hi, chi)
        sub_lo[n], sub_hi[n] = lo, hi
    return cover, max_depth, visited


def _box4(row, k):
    return np.array([row[0 + k], row[8 + k], row[16 + k]]), np.array([row[4 + k], row[12 + k], row[20 + k]])


def _box2(row, k):
    return (np.array([row[4 * k + 0], row[4 * k + 2], row[8 + 2 * k]]), np.array([row[4 * k + 1], row[4 * k + 3], row[8 + 2 * k + 1]]))


@pytest.mark.parametrize("builder", ["RT_BVH_DEVICE_PLOC", "RT_BVH_DEVICE_SAH"])
def test_device_tree_is_a_conservative_partition(rt, scene_dir, builder, monkeypatch):
    monkeypatch.setenv("RT795_DEBUG_BVH", "1")                           # keep the binary BVH (the collapse's input) resident for rt_debug_read_bvh
    xml = rt.scenes.scene_c3(os.path.join(scene_dir, "c3dev"), res=(96, 54), spp=1, n=96, deterministic=True)   # 18 050 triangles
    hs, (dev,) = _scenes(rt, xml, [getattr(rt, builder)])
    st = dev.stats()
    assert st.bvh_device_meshes >= 1 and 3 <= st.bvh_ploc_iterations <= 400
    tris, _ = dev.read_bvh(rt.RT_DEBUG_TRIS)
    ids = tris.view(np.int32)[:, 3]
    assert sorted(ids.tolist()) == list(range(len(ids)))                 # the leaf-ordered triangle array is a permutation of the scene's
    n4, _ = dev.read_bvh(rt.RT_DEBUG_BVH4)
    n2, _ = dev.read_bvh(rt.RT_DEBUG_BVH2)
    roots4, _ = dev.read_bvh(rt.RT_DEBUG_OBJ_ROOT4)
    # the displaced-sphere mesh is the big one: find its triangle range = the largest mesh
    desc = hs.desc.contents if hasattr(hs.desc, "contents") else hs.desc
    big = max(range(desc.num_meshes), key=lambda m: desc.meshes[m].tri_count)
    nt = desc.meshes[big].tri_count
    assert nt >= 8
    obj = next(i for i in range(desc.num_objects) if desc.objects[i].kind == rt.RT_OBJ_MESH and desc.objects[i].geom == big)
    root4 = int(roots4.view(np.int32)[obj, 0])
    cover, depth4, count4 = _walk(n4, root4, 4, tris, 24, _box4)
    in_mesh = np.isin(ids, np.arange(desc.meshes[big].first_tri, desc.meshes[big].first_tri + nt))
    assert (cover[in_mesh] == 1).all() and (cover[~in_mesh] == 0).all()  # every triangle of the mesh in exactly one leaf
    assert depth4 <= 40 and count4 < nt
    # binary twin: same partition.  Its root is the first node of the mesh's segment = the node whose subtree covers the mesh;
    # meshes are laid out in order, so walk from every candidate root until the cover matches
    refs2 = n2[:, 12:14].view(np.int32)
    first_leaf = np.where(in_mesh)[0].min()
    cand = [i for i in range(n2.shape[0]) if any((r < 0 and r != K_NO_CHILD and ((~r) >> 3) <= first_leaf < ((~r) >> 3) + ((~r) & 7) + 1) for r in refs2[i].tolist())]
    assert cand
    # climb to the subtree root: the node no other node of the segment points to
    pointed = set(int(r) for r in refs2.ravel().tolist() if 0 <= r < n2.shape[0])
    parent = {}
    for i in range(n2.shape[0]):
        for r in refs2[i].tolist():
            if 0 <= r < n2.shape[0]:
                parent[int(r)] = i
    r2 = cand[0]
    while r2 in parent:
        r2 = parent[r2]
    cover2, depth2, _ = _walk(n2, r2, 2, tris, 12, _box2)
    assert (cover2[in_mesh] == 1).all() and (cover2[~in_mesh] == 0).all()
    assert depth2 <= 60
    dev.close(); hs.close()


@pytest.mark.parametrize("builder", ["RT_BVH_DEVICE_PLOC", "RT_BVH_DEVICE_SAH"])
def test_device_and_host_builders_trace_the_same_hits(rt, scene_dir, builder):
    """Primary hits (binary BVH, rt_trace_primary) and rendered frames (4-wide BVH) with the device-built and the
    host-built tree: same closest hit everywhere up to float ties at triangle edges."""
    xml = rt.scenes.scene_c4(os.path.join(scene_dir, "c4dev"), res=(256, 144), spp=4, n=120, deterministic=True)   # 28 322 triangles x 16 instances
    hs, (host, dev) = _scenes(rt, xml, [rt.RT_BVH_HOST_SAH, getattr(rt, builder)])
    assert host.stats().bvh_device_meshes == 0 and dev.stats().bvh_device_meshes >= 1
    cam = hs.camera(0)
    ha, hb = host.trace_primary(cam), dev.trace_primary(cam)
    assert (ha["full"] == hb["full"]).mean() > 0.9995
    both = (ha["full"] == 1) & (hb["full"] == 1)
    same = (ha["object"] == hb["object"]) & (ha["primitive"] == hb["primitive"])
    assert same[both].mean() > 0.999
    assert np.allclose(ha["t"][both & same], hb["t"][both & same], rtol=1e-6, atol=0)
    ia, ib = host.render(cam), dev.render(cam)
    close = np.isclose(ia, ib, rtol=1e-4, atol=1e-4).all(axis=-1)
    assert close.mean() > 0.998
    assert abs(float(ia.mean()) - float(ib.mean())) <= 2e-4 * float(ia.mean())
    for s in (host, dev):
        s.close()
    hs.close()


def test_builder_selection_and_small_meshes(rt, scene_dir):
    """AUTO keeps small meshes on the host; DEVICE builds anything of >= 8 triangles on the GPU (Cornell walls: 10-12
    triangles) and the frames agree; an unknown builder id is rejected."""
    xml = rt.scenes.scene_c2(os.path.join(scene_dir, "c2dev"), res=96, spp=1)
    hs, (auto, dev, dev2) = _scenes(rt, xml, [rt.RT_BVH_AUTO, rt.RT_BVH_DEVICE_PLOC, rt.RT_BVH_DEVICE_SAH])
    assert auto.stats().bvh_device_meshes == 0
    cam = hs.camera(0)
    ha = auto.trace_primary(cam)
    for d in (dev, dev2):
        assert d.stats().bvh_device_meshes >= 1
        hb = d.trace_primary(cam)
        assert (ha["full"] == hb["full"]).all()
        same = (ha["object"] == hb["object"]) & (ha["primitive"] == hb["primitive"])
        assert same.mean() > 0.999
    dev2.close()
    with pytest.raises(rt.RtError):
        rt.Scene(hs.desc, 0, bvh_builder=9)
    for s in (auto, dev):
        s.close()
    hs.close()


@pytest.mark.parametrize("builder", ["RT_BVH_DEVICE_PLOC", "RT_BVH_DEVICE_SAH"])
def test_device_build_handles_duplicate_and_degenerate_triangles(rt, scene_dir, tmp_path, builder):
    """All-ties input: many copies of the same triangles (every merge distance ties) plus zero-area triangles.  The build
    must terminate (buddy tie-break) and still partition every triangle."""
    v, f, uv = rt.scenes.displaced_sphere(24)
    f = np.asarray(f)
    f = np.concatenate([f, f, f, np.repeat(f[:1], 300, axis=0), np.stack([f[:50, 0], f[:50, 0], f[:50, 0]], axis=1)])
    d = os.path.join(scene_dir, "dupmesh")
    xml = rt.scenes.scene_c3(d, res=(64, 36), spp=1, n=24, deterministic=True, name="dup")
    rt.scenes.write_ply(os.path.join(d, "dup_mesh.ply"), v, f, uv)
    hs, (host, dev) = _scenes(rt, xml, [rt.RT_BVH_HOST_SAH, getattr(rt, builder)])
    assert dev.stats().bvh_device_meshes >= 1
    tris, _ = dev.read_bvh(rt.RT_DEBUG_TRIS)
    ids = tris.view(np.int32)[:, 3]
    assert sorted(ids.tolist()) == list(range(len(ids)))
    cam = hs.camera(0)
    ha, hb = host.trace_primary(cam), dev.trace_primary(cam)
    assert (ha["full"] == hb["full"]).mean() > 0.999
    both = (ha["full"] == 1) & (hb["full"] == 1)
    assert np.allclose(ha["t"][both], hb["t"][both], rtol=1e-5)           # duplicates: the primitive id of a tie may differ, the distance may not
    for s in (host, dev):
        s.close()
    hs.close()

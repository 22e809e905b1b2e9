"""Iso-surface extraction without a GPU (second half): the product's triangle table against the oracle's and by its own consistency,
the product's KERNEL (csrc/marching_cubes.cu executed on the CPU through tests/cuda_emu) against oracle/marchingcubes_oracle.c bit for bit, and the host-side mesh
clean-up / PLY writer (real library code: it is host code) against an independent restatement of mLib's MeshData::mergeCloseVertices / removeDuplicateFaces /
MeshIO::saveToPLY.  The oracle itself is pinned against the reference's kernel in tests/test_marchingcubes_reference_emulated.py."""
import ctypes as C
import os
import re
import struct

import numpy as np
import pytest

from bundlefusion_b200 import _capi as capi
from bundlefusion_b200 import marching_cubes as mc
from oracle import oracle as orc
from tests.test_marchingcubes_reference_emulated import canon, golden_params, golden_scene, scene_box

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EDGES = [(0, 1), (1, 2), (2, 3), (3, 0), (4, 5), (5, 6), (6, 7), (7, 4), (0, 4), (1, 5), (2, 6), (3, 7)]


def product_table():
    """the packed words of csrc/mc_tables.cuh -> rows of edge indices"""
    src = open(os.path.join(ROOT, "bundlefusion_b200", "csrc", "mc_tables.cuh")).read()
    words = [int(w, 16) for w in re.findall(r"0x([0-9a-f]{16})ull", src)]
    assert len(words) == 256
    rows = []
    for w in words:
        r = [(w >> (4 * i)) & 15 for i in range(16)]
        n = r.index(15)
        assert all(v == 15 for v in r[n:]) and n % 3 == 0 and n <= 15
        rows.append(r[:n])
    return rows


def test_packed_table_is_consistent_and_equals_the_oracles():
    rows = product_table()
    edge, tri = orc.marchingcubes_tables()
    for c, r in enumerate(rows):
        assert r == [int(v) for v in tri[c] if v >= 0], c
        mask = sum(1 << e for e, (a, b) in enumerate(EDGES) if ((c >> a) & 1) != ((c >> b) & 1))
        assert mask == int(edge[c]) and mask == sum(1 << e for e in set(r)), c                 # exactly the crossing edges are used
        half = {}
        for t in range(0, len(r), 3):
            a, b, d = r[t:t + 3]
            for u, v in ((a, b), (b, d), (d, a)):
                assert (u, v) not in half, c                                                # oriented manifold: no directed edge twice
                half[(u, v)] = 1
        out = {}
        for (u, v) in half:
            if (v, u) not in half:
                out[u] = out.get(u, 0) + 1
        assert all(out.get(e, 0) == 1 for e in set(r)), c                                   # every vertex lies once on the patch boundary (which runs on the cube's faces)


_EXTRA = r'''
#include "%s"
#include "%s"
#include "%s"
#include <cerrno>
namespace bf { static BFHashParams g_emuHp; const BFHashParams* bound_hash_params() { return &g_emuHp; } static inline void set_last_error(const char*, int) {} }
extern "C" void updateConstantHashParams(const BFHashParams* p) { bf::g_emuHp = *p; }
enum { cudaMemcpyDeviceToHost = 2 };
static inline int cudaMallocHost(void** p, size_t n) { *p = calloc(1, n); return *p ? 0 : 2; }
template <class T> static inline int cudaMallocHost(T** p, size_t n) { return cudaMallocHost(reinterpret_cast<void**>(p), n); }
static inline int cudaFreeHost(void* p) { free(p); return 0; }
static inline int cudaMemcpyAsync(void* d, const void* s, size_t n, int, cudaStream_t) { memcpy(d, s, n); return 0; }
static inline int cudaStreamSynchronize(cudaStream_t) { return 0; }
'''


@pytest.fixture(scope="module")
def emu():
    from tests.cuda_emu import build_emulated
    inc = os.path.join(ROOT, "include")
    L = build_emulated("marching_cubes.cu", 4, extra_pre=_EXTRA % (os.path.join(inc, "bf_tsdf.h"), os.path.join(inc, "bf_raycast.h"), os.path.join(inc, "bf_marchingcubes.h")))
    vp = C.c_void_p
    L.bfMarchingCubesExtract.argtypes = [vp] * 5
    L.resetMarchingCubesCUDA.argtypes = [vp]; L.resetMarchingCubesCUDA.restype = None
    L.extractIsoSurfaceCUDA.argtypes = [vp] * 4; L.extractIsoSurfaceCUDA.restype = None
    L.updateConstantHashParams.argtypes = [vp]
    L.bfMarchingCubesCreate.argtypes = [vp, C.POINTER(vp)]
    L.bfMarchingCubesExtractIsoSurface.argtypes = [vp, vp, vp, vp, vp, C.c_int]
    L.bfMarchingCubesGetSoup.argtypes = [vp, C.POINTER(vp), C.POINTER(vp)]; L.bfMarchingCubesGetSoup.restype = C.c_size_t
    L.bfMarchingCubesDestroy.argtypes = [vp]; L.bfMarchingCubesDestroy.restype = None
    return L


def run_emulated(L, sc, p):
    tri = np.zeros((int(p.m_maxNumTriangles), 3, 6), np.float32)
    n = np.full(1, 77, np.uint32)
    assert L.bfMarchingCubesExtract(C.addressof(sc.hd), C.addressof(sc.hp), C.addressof(p), tri.ctypes.data, n.ctypes.data) == 0
    return tri[:int(n[0])].copy()


def test_emulated_kernel_matches_oracle_bit_for_bit(emu):
    sc, cam, frames = golden_scene()
    p = golden_params(sc.hp)
    want, found = orc.marchingcubes_extract(sc, p)
    got = run_emulated(emu, sc, p)
    assert len(got) == len(want) == found > 10000 and np.array_equal(canon(got), canon(want))
    sc, cam, frames = golden_scene(small=True)                                      # the variants on a smaller model
    want, found = orc.marchingcubes_extract(sc, golden_params(sc.hp))
    assert found > 1000
    box = scene_box(sc)
    pb = golden_params(sc.hp, box)
    assert np.array_equal(canon(run_emulated(emu, sc, pb)), canon(orc.marchingcubes_extract(sc, pb)[0]))
    # a full buffer: the count is the capacity, every triangle written is one of the full soup's
    pc = golden_params(sc.hp, cap=500)
    capped = run_emulated(emu, sc, pc)
    full = {r.tobytes() for r in canon(want)}
    assert len(capped) == 500 and all(r.tobytes() in full for r in canon(capped))


def test_emulated_reference_named_stubs_and_host_class(emu, tmp_path):
    sc, cam, frames = golden_scene(small=True)
    p = golden_params(sc.hp, scene_box(sc))
    want, _ = orc.marchingcubes_extract(sc, p)
    # the reference's call sequence: constants, reset, extract with the parameters in "device" memory (MarchingCubesData::updateParams)
    tri = np.zeros((int(p.m_maxNumTriangles), 3, 6), np.float32)
    n = np.full(1, 5, np.uint32)
    dparams = capi.BFMarchingCubesParams.from_buffer_copy(p)
    data = capi.BFMarchingCubesData(C.addressof(dparams), n.ctypes.data, tri.ctypes.data, 1)
    host = golden_params(sc.hp)                                            # the host copy only sizes the grid in the reference; the box comes from d_params
    emu.updateConstantHashParams(C.addressof(sc.hp))
    emu.resetMarchingCubesCUDA(C.addressof(data))
    assert n[0] == 0
    emu.extractIsoSurfaceCUDA(C.addressof(sc.hd), None, C.addressof(host), C.addressof(data))
    assert np.array_equal(canon(tri[:int(n[0])]), canon(want))
    # class CUDAMarchingCubesHashSDF: two extractions append to the mesh buffer
    h = C.c_void_p()
    assert emu.bfMarchingCubesCreate(C.addressof(host), C.byref(h)) == 0
    lo, hi = (C.c_float * 3)(*scene_box(sc)[0]), (C.c_float * 3)(*scene_box(sc)[1])
    assert emu.bfMarchingCubesExtractIsoSurface(h, C.addressof(sc.hd), C.addressof(sc.hp), lo, hi, 1) == 0
    assert emu.bfMarchingCubesGetSoup(h, None, None) == 3 * len(want)
    assert emu.bfMarchingCubesExtractIsoSurface(h, C.addressof(sc.hd), C.addressof(sc.hp), None, None, 0) == 0
    pp, cp = C.c_void_p(), C.c_void_p()
    nv = emu.bfMarchingCubesGetSoup(h, C.byref(pp), C.byref(cp))
    full, _ = orc.marchingcubes_extract(sc, host)
    assert nv == 3 * (len(want) + len(full))
    pos = np.ctypeslib.as_array(C.cast(pp, C.POINTER(C.c_float)), (nv, 3)); col = np.ctypeslib.as_array(C.cast(cp, C.POINTER(C.c_float)), (nv, 4))
    first = np.concatenate([pos[:3 * len(want)], col[:3 * len(want), :3]], axis=1).reshape(len(want), 3, 6)
    assert np.array_equal(canon(first), canon(want)) and np.all(col[:, 3] == 1.0)
    # saveMesh: the buffer (both extractions, in their order) merged, de-duplicated and written; the buffer is cleared; an existing name counts up
    from tests.test_mesh_reference_host import TRANSFORM, library_save
    soup = np.concatenate([pos, col[:, :3]], axis=1).reshape(-1, 3, 6).copy()
    emu.bfMarchingCubesSaveMesh.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int, C.c_char_p, C.c_size_t]
    out = C.create_string_buffer(4096)
    path = str(tmp_path / "scans" / "scan.ply")
    assert emu.bfMarchingCubesSaveMesh(h, path.encode(), None, 0, out, 4096) == 0 and out.value.decode() == path
    assert emu.bfMarchingCubesGetSoup(h, None, None) == 0
    library_save(soup, None, str(tmp_path / "want.ply"))
    assert open(path, "rb").read() == open(str(tmp_path / "want.ply"), "rb").read()
    assert emu.bfMarchingCubesExtractIsoSurface(h, C.addressof(sc.hd), C.addressof(sc.hp), lo, hi, 1) == 0
    nv = emu.bfMarchingCubesGetSoup(h, C.byref(pp), C.byref(cp))
    pos = np.ctypeslib.as_array(C.cast(pp, C.POINTER(C.c_float)), (nv, 3)); col = np.ctypeslib.as_array(C.cast(cp, C.POINTER(C.c_float)), (nv, 4))
    soup = np.concatenate([pos, col[:, :3]], axis=1).reshape(-1, 3, 6).copy()
    T = np.ascontiguousarray(TRANSFORM)
    assert emu.bfMarchingCubesSaveMesh(h, path.encode(), T.ctypes.data, 0, out, 4096) == 0 and out.value.decode() == str(tmp_path / "scans" / "scan1.ply")
    library_save(soup, TRANSFORM, str(tmp_path / "want_t.ply"))
    assert open(out.value.decode(), "rb").read() == open(str(tmp_path / "want_t.ply"), "rb").read()
    assert emu.bfMarchingCubesSaveMesh(h, path.encode(), None, 1, out, 4096) == 0 and out.value.decode() == path          # overwrite: an empty mesh under the first name
    assert b"element vertex 0" in open(path, "rb").read()
    emu.bfMarchingCubesDestroy(h)


# ---- the host-side mesh clean-up: independent restatement of mLib core-mesh/meshData.cpp:40-100, 200-300 ----
def py_merge(pos, col, faces, thresh):
    inv = np.float32(1.0) / np.float32(thresh)
    grid, look, keep = {}, [], []
    for v, p in enumerate(pos):
        c = tuple(int(np.float32(np.float32(x * inv) + np.float32(0.5) * np.float32(np.sign(x)))) for x in p)
        nn = None
        for i in (-1, 0, 1):
            for j in (-1, 0, 1):
                for k in (-1, 0, 1):
                    if nn is None and (c[0] + i, c[1] + j, c[2] + k) in grid:
                        nn = grid[(c[0] + i, c[1] + j, c[2] + k)]
        if nn is None:
            grid[c] = len(keep); look.append(len(keep)); keep.append(v)
        else:
            look.append(nn)
    f = np.array([[look[i] for i in t] for t in faces], np.uint32).reshape(-1, 3)
    f = f[(f[:, 0] != f[:, 1]) & (f[:, 0] != f[:, 2]) & (f[:, 1] != f[:, 2])]
    return pos[keep], col[keep], f


def py_dedup(faces):
    seen, out = set(), []
    for t in faces:
        k = tuple(sorted(int(i) for i in t))
        if k not in seen:
            seen.add(k); out.append(t)
    return np.array(out, np.uint32).reshape(-1, 3)


def test_mesh_cleanup_and_ply_writer(tmp_path):
    sc, cam, frames = golden_scene()
    tri, _ = orc.marchingcubes_extract(sc, golden_params(sc.hp, scene_box(sc)))
    tri = tri[:1500]
    pos = np.ascontiguousarray(tri[..., :3].reshape(-1, 3)); col = np.concatenate([tri[..., 3:].reshape(-1, 3), np.ones((len(pos), 1), np.float32)], axis=1)
    faces = np.arange(len(pos), dtype=np.uint32).reshape(-1, 3)
    faces = np.concatenate([faces, faces[10:14][:, [1, 2, 0]], np.array([[0, 0, 1]], np.uint32)])                    # rotated duplicates and a degenerate face
    gp, gc, gf = mc.merge_close_vertices(pos, col, faces, 0.00001)
    wp, wc, wf = py_merge(pos, col, faces, 0.00001)
    assert len(gp) == len(wp) < len(pos) / 3 and np.array_equal(gp.view(np.uint32), wp.view(np.uint32)) and np.array_equal(gc, wc) and np.array_equal(gf, wf)
    gd, wd = mc.remove_duplicate_faces(gf), py_dedup(wf)
    assert np.array_equal(gd, wd) and len(gd) == len(gf) - 4
    # neighbouring cells share their edge vertices: the merged mesh is (nearly everywhere) a closed fan around interior vertices -- far fewer vertices than corners
    assert len(gp) < 0.3 * len(pos)
    path = str(tmp_path / "m.ply")
    mc.save_ply(path, gp, gc, gd)
    raw = open(path, "rb").read()
    head, body = raw.split(b"end_header\n", 1)
    assert head.decode().splitlines() == ["ply", "format binary_little_endian 1.0", "comment MLIB generated", f"element vertex {len(gp)}", "property float x", "property float y",
                                          "property float z", "property uchar red", "property uchar green", "property uchar blue", "property uchar alpha",
                                          f"element face {len(gd)}", "property list uchar int vertex_indices"]
    assert len(body) == 16 * len(gp) + 13 * len(gd)
    v = np.frombuffer(body[:16 * len(gp)], np.dtype([("p", "<f4", 3), ("c", "u1", 4)]))
    assert np.array_equal(v["p"].view(np.uint32), gp.view(np.uint32)) and np.array_equal(v["c"], (gc * np.float32(255)).astype(np.uint8))
    f = np.frombuffer(body[16 * len(gp):], np.dtype([("n", "u1"), ("i", "<i4", 3)]))
    assert np.all(f["n"] == 3) and np.array_equal(f["i"].astype(np.uint32), gd)
    mc.save_ply(path, gp, None, gd)                                                                                  # without colours: positions only
    assert len(open(path, "rb").read().split(b"end_header\n", 1)[1]) == 12 * len(gp) + 13 * len(gd)
    with pytest.raises(ValueError):
        mc.merge_close_vertices(pos, col, np.array([[0, 1, len(pos)]], np.uint32))

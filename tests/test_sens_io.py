"""`.sens` reader / writer and its PNG / JPEG decoders (csrc/sens_io.cu behind include/bf_sens.h; SURVEY.md section 8f, row N4, first half).  Host code: runs
without a GPU.  The container is checked against an independent pure-Python reader written from the layout in external/mLib ext-depthcamera/sensorData.h
(loadFromFile :1187-1227, RGBDFrame :686-700) -- and against files this test assembles byte by byte, as ml::SensorData::saveToFile would with JPEG / PNG colour
(encoded by libjpeg / libpng through PIL) and zlib depth; the decoders against PIL's decode of the same bytes."""
import io
import os
import struct
import zlib

import numpy as np
import pytest

from bundlefusion_b200 import sens, synth

W, H = 160, 120


def frames(n, w=W, h=H):
    out = []
    for i in range(n):
        d, c, T = synth.make_frame(3 * i, w, h, texture="rich")
        du = np.where(np.isfinite(d), np.clip(np.round(d * 1000.0), 1, 65535), 0).astype(np.uint16)          # millimetres, 0 = no measurement
        out.append((du, np.ascontiguousarray(c[..., :3]), T.astype(np.float32)))
    return out


def python_reader(path):
    """independent restatement of ml::SensorData::loadFromFile for the variants a writer can produce without an image codec"""
    with open(path, "rb") as f:
        b = f.read()
    at = 0

    def take(fmt):
        nonlocal at
        v = struct.unpack_from("<" + fmt, b, at); at += struct.calcsize("<" + fmt)
        return v if len(v) > 1 else v[0]
    hdr = {"version": take("I")}
    n = take("Q"); hdr["name"] = b[at:at + n].decode(); at += n
    hdr["calib"] = [np.array(take("16f"), np.float32).reshape(4, 4) for _ in range(4)]
    hdr["cc"], hdr["dc"], hdr["cw"], hdr["ch"], hdr["dw"], hdr["dh"] = take("iiIIII")
    hdr["shift"] = take("f")
    nf = take("Q")
    fr = []
    for _ in range(nf):
        pose = np.array(take("16f"), np.float32).reshape(4, 4)
        tc, td, cb, db = take("QQQQ")
        col = b[at:at + cb]; at += cb
        dep = b[at:at + db]; at += db
        fr.append((pose, tc, td, col, dep))
    nimu = take("Q")
    assert at + nimu * 128 == len(b)
    return hdr, fr


@pytest.mark.parametrize("zl", [False, True])
def test_write_read_round_trip_and_independent_reader(tmp_path, zl):
    fs = frames(5)
    K = np.eye(4, dtype=np.float32); K[0, 0] = K[1, 1] = 131.25; K[0, 2] = 79.5; K[1, 2] = 59.5
    path = str(tmp_path / "t.sens")
    w = sens.SensorDataWriter(path, W, H, K, depth_shift=1000.0, zlib_depth=zl, sensor_name="synthetic")
    for i, (d, c, T) in enumerate(fs):
        w.append(d, c, T if i != 2 else None, ts_color=100 + i, ts_depth=200 + i)
    w.finish()
    # the independent reader sees what ml::SensorData::loadFromFile would
    hdr, fr = python_reader(path)
    assert hdr["version"] == 4 and hdr["name"] == "synthetic" and (hdr["cw"], hdr["ch"], hdr["dw"], hdr["dh"]) == (W, H, W, H) and hdr["shift"] == 1000.0
    assert hdr["cc"] == sens.COLOR_RAW and hdr["dc"] == (sens.DEPTH_ZLIB_USHORT if zl else sens.DEPTH_RAW_USHORT)
    assert np.array_equal(hdr["calib"][0], K) and np.array_equal(hdr["calib"][1], np.eye(4)) and np.array_equal(hdr["calib"][2], K)
    assert len(fr) == 5
    for i, ((pose, tc, td, col, dep), (d, c, T)) in enumerate(zip(fr, fs)):
        assert (tc, td) == (100 + i, 200 + i) and col == c.tobytes()
        assert (zlib.decompress(dep) if zl else dep) == d.tobytes()
        assert np.all(np.isneginf(pose)) if i == 2 else np.array_equal(pose, T)
    # the library's reader
    r = sens.SensorDataReader(path)
    assert len(r) == 5 and r.header.sensorName == b"synthetic" and r.header.depthShift == 1000.0 and r.header.numIMUFrames == 0
    for i, (d, c, T) in enumerate(fs):
        depth, color, pose, ts = r.frame(i)
        want = np.where(d == 0, -np.inf, d.astype(np.float32) / np.float32(1000.0)).astype(np.float32)        # FL/SensorDataReader.cpp:104-107
        assert np.array_equal(depth.view(np.uint32), want.view(np.uint32))
        assert np.array_equal(color[..., :3], c) and np.all(color[..., 3] == 1)                              # vec4uc(vec3uc): w = 1
        assert list(ts) == [100 + i, 200 + i]
        du, cu = r.frame_raw(i)
        assert np.array_equal(du, d) and np.array_equal(cu, c)
    with pytest.raises(RuntimeError):
        r.frame(5)
    r.close()


def assemble_sens(path, fs, color_type, encode):
    """a version-4 file as ml::SensorData::saveToFile lays it out, colour through an image codec, depth zlib"""
    K = np.eye(4, dtype=np.float32); I = np.eye(4, dtype=np.float32)
    name = b"StructureSensor"
    with open(path, "wb") as f:
        f.write(struct.pack("<IQ", 4, len(name)) + name)
        for M in (K, I, K, I):
            f.write(M.tobytes())
        f.write(struct.pack("<iiIIIIf", color_type, sens.DEPTH_ZLIB_USHORT, W, H, W, H, 1000.0))
        f.write(struct.pack("<Q", len(fs)))
        for i, (d, c, T) in enumerate(fs):
            cb, db = encode(c), zlib.compress(d.tobytes(), 6)
            f.write(T.tobytes() + struct.pack("<QQQQ", i, i, len(cb), len(db)) + cb + db)
        f.write(struct.pack("<Q", 2))                        # two IMU frames (5 x vec3d + a time stamp each), which the reader skips
        f.write(b"\0" * 256)


def pil_encode(fmt, **kw):
    from PIL import Image

    def enc(c):
        bio = io.BytesIO(); Image.fromarray(c).save(bio, fmt, **kw)
        return bio.getvalue()
    return enc


def pil_decode(data):
    from PIL import Image
    return np.asarray(Image.open(io.BytesIO(data)).convert("RGB"))


def test_reads_png_colour_exactly(tmp_path):
    fs = frames(2)
    path = str(tmp_path / "p.sens")
    assemble_sens(path, fs, sens.COLOR_PNG, pil_encode("PNG"))
    r = sens.SensorDataReader(path)
    assert r.header.colorCompression == sens.COLOR_PNG and r.header.numIMUFrames == 2 and r.header.sensorName == b"StructureSensor"
    for i, (d, c, T) in enumerate(fs):
        du, cu = r.frame_raw(i)
        assert np.array_equal(du, d) and np.array_equal(cu, c)                    # PNG is lossless: bit for bit
        assert np.array_equal(r.frame(i)[2], T)


@pytest.mark.parametrize("sub,q", [(0, 95), (1, 90), (2, 85), (2, 50)])          # PIL subsampling 0 = 4:4:4, 1 = 4:2:2, 2 = 4:2:0
def test_reads_jpeg_colour_like_libjpeg(tmp_path, sub, q):
    fs = frames(2)
    path = str(tmp_path / "j.sens")
    enc = pil_encode("JPEG", quality=q, subsampling=sub)
    assemble_sens(path, fs, sens.COLOR_JPEG, enc)
    r = sens.SensorDataReader(path)
    for i, (d, c, T) in enumerate(fs):
        du, cu = r.frame_raw(i)
        ref = pil_decode(enc(c))                                                  # libjpeg's decode of the same bytes
        diff = np.abs(cu.astype(np.int32) - ref.astype(np.int32))
        # the IDCT and the chroma up-sampling are not normative: decoders agree to a level or two (this one follows the reference's stb_image, not libjpeg)
        assert diff.max() <= 3 and diff.mean() < (0.35 if q >= 85 else 0.8), (diff.max(), diff.mean())     # coarser quantisation: more pixels where two correct IDCTs round apart
        assert np.array_equal(du, d)


def test_decoders_on_odd_sizes_grey_restart_markers_and_errors():
    from PIL import Image
    rng = np.random.default_rng(3)
    img = (rng.random((37, 53, 3)) * 255).astype(np.uint8)
    img = np.asarray(Image.fromarray(img).resize((53 * 3, 37 * 3), Image.BILINEAR))          # smooth enough for JPEG, 159 x 111: not a multiple of the MCU
    for kw in (dict(quality=90, subsampling=2), dict(quality=90, subsampling=1), dict(quality=92, subsampling=0)):
        bio = io.BytesIO(); Image.fromarray(img).save(bio, "JPEG", **kw)
        got, ref = sens.decode_jpeg(bio.getvalue()), pil_decode(bio.getvalue())
        # odd sizes: the reference's decoder (stb_image, followed here bit for bit -- test_sens_reference_stb.py) up-samples from the chroma samples under the image
        # and replicates at the right / bottom border, libjpeg from the padded MCU: the last chroma pair may differ, the interior may not
        assert got.shape == ref.shape and np.abs(got.astype(int) - ref.astype(int))[:-2, :-2].max() <= 3
    grey = np.asarray(Image.fromarray(img).convert("L"))
    bio = io.BytesIO(); Image.fromarray(grey).save(bio, "JPEG", quality=90)
    got = sens.decode_jpeg(bio.getvalue())
    assert np.abs(got[..., 0].astype(int) - np.asarray(Image.open(io.BytesIO(bio.getvalue()))).astype(int)).max() <= 1 and np.array_equal(got[..., 0], got[..., 2])
    bio = io.BytesIO(); Image.fromarray(img).save(bio, "JPEG", quality=90, subsampling=2, restart_marker_blocks=3)        # DRI + RSTn
    if b"\xff\xdd" in bio.getvalue():
        assert np.abs(sens.decode_jpeg(bio.getvalue()).astype(int) - pil_decode(bio.getvalue()).astype(int))[:-2, :-2].max() <= 3
    bio = io.BytesIO(); Image.fromarray(img).save(bio, "JPEG", quality=90, progressive=True)                              # SOF2: spectral selection + successive approximation
    assert b"\xff\xc2" in bio.getvalue()
    assert np.abs(sens.decode_jpeg(bio.getvalue()).astype(int) - pil_decode(bio.getvalue()).astype(int))[:-2, :-2].max() <= 3
    arith = bio.getvalue().replace(b"\xff\xc2", b"\xff\xc9", 1)                                                         # SOF9 (arithmetic coding): refused, as by the reference's decoder
    with pytest.raises(RuntimeError, match="unsupported"):
        sens.decode_jpeg(arith)
    for mode in ("RGB", "RGBA", "L", "LA", "P", "1"):                                             # "1": one bit per pixel, rows filtered (libpng picks Up / Paeth)
        bio = io.BytesIO(); Image.fromarray(img).convert(mode).save(bio, "PNG")
        assert np.array_equal(sens.decode_png(bio.getvalue()), np.asarray(Image.open(io.BytesIO(bio.getvalue())).convert("RGB")))
    for colors in (2, 4, 16):                                                                     # palette images of 1, 2 and 4 bits
        bio = io.BytesIO(); Image.fromarray(img).convert("P", palette=Image.ADAPTIVE, colors=colors).save(bio, "PNG", bits={2: 1, 4: 2, 16: 4}[colors])
        assert np.array_equal(sens.decode_png(bio.getvalue()), np.asarray(Image.open(io.BytesIO(bio.getvalue())).convert("RGB")))
    from tests.test_sens_reference_stb import handmade_png
    for il in (True, False):                                                                      # Adam7 (PIL reads it, cannot write it)
        d = handmade_png(img, 2, 8, None, il)
        assert np.array_equal(sens.decode_png(d), img) and np.array_equal(np.asarray(Image.open(io.BytesIO(d)).convert("RGB")), img)
    bio = io.BytesIO(); Image.fromarray(img[..., 0].astype(np.uint16) * 257).save(bio, "PNG")                               # 16 bits per channel: refused, as by the reference's decoder
    with pytest.raises(RuntimeError, match="unsupported"):
        sens.decode_png(bio.getvalue())
    with pytest.raises(RuntimeError):
        sens.decode_png(b"not a png at all, not even close to one........")
    with pytest.raises(RuntimeError):
        sens.decode_jpeg(bio.getvalue())                    # PNG bytes are not a JPEG


def test_open_rejects_other_versions(tmp_path):
    p = str(tmp_path / "v.sens")
    with open(p, "wb") as f:
        f.write(struct.pack("<IQ", 3, 0))
    with pytest.raises(RuntimeError):
        sens.SensorDataReader(p)
    with pytest.raises(RuntimeError):
        sens.SensorDataReader(str(tmp_path / "missing.sens"))


def test_ate_of_a_trajectory_against_the_recorded_poses(tmp_path):
    """scan.ate: rigid alignment then residuals; scan.recorded_poses: the poses of a file without decoding its images"""
    from bundlefusion_b200 import scan
    fs = frames(6)
    K = np.eye(4, dtype=np.float32)
    path = str(tmp_path / "p.sens")
    w = sens.SensorDataWriter(path, W, H, K)
    for i, (d, c, T) in enumerate(fs):
        w.append(d, c, T if i != 2 else np.full((4, 4), -np.inf, np.float32))
    w.finish()
    gt = scan.recorded_poses(path)
    assert gt.shape == (6, 4, 4) and np.isneginf(gt[2]).all() and np.array_equal(gt[0], fs[0][2])
    # an estimate in another world frame (the loop starts at the identity), with 5 mm of noise on two frames
    rng = np.random.default_rng(0)
    G = np.eye(4); G[:3, :3] = np.linalg.qr(rng.normal(size=(3, 3)))[0]; G[:3, :3] *= np.sign(np.linalg.det(G[:3, :3])); G[:3, 3] = (0.3, -1.0, 2.0)
    est = np.stack([(G @ T.astype(np.float64)) for _, _, T in fs])
    e = scan.ate(est, gt)
    assert e["frames"] == 5 and e["rmse"] < 1e-6
    est[1, :3, 3] += (0.005, 0, 0); est[4, :3, 3] -= (0, 0.005, 0)
    e = scan.ate(est, gt)
    assert 0.001 < e["rmse"] < 0.005 and e["max"] < 0.0051 and e["mean"] <= e["rmse"]
    assert np.isnan(scan.ate(est[:2], gt[:2])["rmse"])                    # fewer than three common frames: nothing to align


def test_decoders_reject_or_decode_mutated_streams_without_crashing():
    """recordings come from outside: a corrupt payload must give an error code (or some picture), not a fault or an absurd allocation"""
    from tests.test_sens_reference_stb import make_streams
    jpegs, pngs, _ = make_streams()
    rng = np.random.default_rng(0)
    rejected = 0
    for it in range(1200):
        pool = jpegs if it % 2 else pngs
        b = bytearray(pool[int(rng.integers(0, len(pool)))])
        mode = it % 4
        if mode == 0:
            for _ in range(int(rng.integers(1, 6))):
                b[int(rng.integers(0, len(b)))] = int(rng.integers(0, 256))
        elif mode == 1:
            b = b[: int(rng.integers(1, len(b)))]
        elif mode == 2:
            i = int(rng.integers(0, len(b))); b[i:i] = bytes(rng.integers(0, 256, int(rng.integers(1, 40))).astype(np.uint8))
        else:
            i = int(rng.integers(2, len(b))); del b[i:i + int(rng.integers(1, 30))]
        try:
            (sens.decode_jpeg if it % 2 else sens.decode_png)(bytes(b))
        except (RuntimeError, MemoryError):
            rejected += 1
    assert rejected > 300
    huge = bytearray(pngs[0]); huge[16:24] = (0x7FFFFFFF).to_bytes(4, "big") * 2          # IHDR claims 2^31 x 2^31 pixels
    with pytest.raises(RuntimeError):
        sens.decode_png(bytes(huge))

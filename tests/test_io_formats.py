"""File formats and console helpers either side of the path: include/edge_alignment/io.h.

CPU only.  The files are produced here by an independent writer (PNG with every filter type, BMP with row padding, PLY
ascii / binary with extra properties); the C++ side (tests/cpp/io_test.cpp) decodes them and the bytes are compared."""
import os
import struct
import subprocess
import zlib

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
U8C3, U8C1, U16C1 = 16, 0, 2
READ_COLOR, READ_ANYDEPTH, READ_GRAYSCALE = 0, 1, 2


@pytest.fixture(scope="module")
def io_bin(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("io") / "io_test")
    subprocess.run(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "io_test.cpp"),
                    "-lz", "-o", out], check=True)
    return out


def write_png(path, arr, ctype, bits, palette=None):
    """arr: [h][w][bytes per pixel] uint8 raster exactly as PNG stores it; rows cycle through filter types 0..4."""
    h, w, bpp = arr.shape
    raw = bytearray()
    prev = np.zeros(w * bpp, np.int32)
    for y in range(h):
        cur = arr[y].reshape(-1).astype(np.int32)
        f = y % 5
        a = np.concatenate([np.zeros(bpp, np.int32), cur[:-bpp]])
        b = prev
        c = np.concatenate([np.zeros(bpp, np.int32), prev[:-bpp]])
        pp = a + b - c
        pa, pb, pc = np.abs(pp - a), np.abs(pp - b), np.abs(pp - c)
        paeth = np.where((pa <= pb) & (pa <= pc), a, np.where(pb <= pc, b, c))
        pred = [np.zeros_like(cur), a, b, (a + b) >> 1, paeth][f]
        raw.append(f)
        raw += ((cur - pred) & 255).astype(np.uint8).tobytes()
        prev = cur

    def chunk(tag, body):
        return struct.pack(">I", len(body)) + tag + body + struct.pack(">I", zlib.crc32(tag + body) & 0xffffffff)
    data = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, bits, ctype, 0, 0, 0))
    if palette is not None:
        data += chunk(b"PLTE", palette.astype(np.uint8).tobytes())
    comp = zlib.compress(bytes(raw), 6)
    data += chunk(b"IDAT", comp[:len(comp) // 2]) + chunk(b"IDAT", comp[len(comp) // 2:]) + chunk(b"IEND", b"")   # split IDAT on purpose
    open(path, "wb").write(data)


def write_bmp24(path, bgr):
    h, w, _ = bgr.shape
    row = (w * 3 + 3) // 4 * 4
    body = bytearray()
    for y in range(h - 1, -1, -1):
        body += bgr[y].tobytes() + b"\0" * (row - w * 3)
    hdr = b"BM" + struct.pack("<IHHI", 54 + len(body), 0, 0, 54) + struct.pack("<IiiHHIIiiII", 40, w, h, 1, 24, 0, len(body), 2835, 2835, 0, 0)
    open(path, "wb").write(hdr + bytes(body))


def write_bmp8(path, gray):
    h, w = gray.shape
    row = (w + 3) // 4 * 4
    pal = b"".join(bytes([i, i, i, 0]) for i in range(256))
    body = bytearray()
    for y in range(h - 1, -1, -1):
        body += gray[y].tobytes() + b"\0" * (row - w)
    off = 54 + 1024
    hdr = b"BM" + struct.pack("<IHHI", off + len(body), 0, 0, off) + struct.pack("<IiiHHIIiiII", 40, w, h, 1, 8, 0, len(body), 2835, 2835, 256, 0)
    open(path, "wb").write(hdr + pal + bytes(body))


def run_img(io_bin, path, flags, tmp):
    out = str(tmp / "img.raw")
    r = subprocess.run([io_bin, "img", path, str(flags), out], check=True, capture_output=True, text=True)
    rows, cols, typ = map(int, r.stdout.split())
    return rows, cols, typ, np.fromfile(out, np.uint8)


def test_png_and_bmp_decoders(io_bin, tmp_path, frames):
    rng = np.random.default_rng(7)
    # 8-bit colour PNG (stored RGB) -> BGR, the bundled frame crop so that the content is image-like
    bgr = np.ascontiguousarray(frames["bgr"][0][100:137, 200:261])            # 37 x 61: odd sizes
    p = str(tmp_path / "c.png")
    write_png(p, bgr[:, :, ::-1].copy(), 2, 8)
    rows, cols, typ, px = run_img(io_bin, p, READ_COLOR, tmp_path)
    assert (rows, cols, typ) == (37, 61, U8C3) and np.array_equal(px.reshape(37, 61, 3), bgr)
    # RGBA drops alpha
    rgba = np.concatenate([bgr[:, :, ::-1], rng.integers(0, 256, (37, 61, 1), dtype=np.uint8)], 2)
    write_png(p, rgba, 6, 8)
    _, _, typ, px = run_img(io_bin, p, READ_COLOR, tmp_path)
    assert typ == U8C3 and np.array_equal(px.reshape(37, 61, 3), bgr)
    # 16-bit depth PNG, big-endian on disk, CV_LOAD_IMAGE_ANYDEPTH keeps 16 bits
    depth = np.ascontiguousarray(frames["depth"][0][100:137, 200:261])
    be = depth.astype(">u2").view(np.uint8).reshape(37, 61, 2)
    d = str(tmp_path / "d.png")
    write_png(d, be, 0, 16)
    rows, cols, typ, px = run_img(io_bin, d, READ_ANYDEPTH, tmp_path)
    assert typ == U16C1 and np.array_equal(px.view(np.uint16).reshape(37, 61), depth)
    _, _, typ, px = run_img(io_bin, d, READ_GRAYSCALE, tmp_path)
    assert typ == U8C1 and np.array_equal(px.reshape(37, 61), (depth >> 8).astype(np.uint8))
    # 8-bit gray and palette PNG
    gray = rng.integers(0, 256, (20, 33, 1), dtype=np.uint8)
    write_png(p, gray, 0, 8)
    _, _, typ, px = run_img(io_bin, p, READ_GRAYSCALE, tmp_path)
    assert typ == U8C1 and np.array_equal(px.reshape(20, 33), gray[:, :, 0])
    _, _, typ, px = run_img(io_bin, p, READ_COLOR, tmp_path)
    assert typ == U8C3 and np.array_equal(px.reshape(20, 33, 3), np.repeat(gray, 3, 2))
    pal = rng.integers(0, 256, (256, 3), dtype=np.uint8)
    write_png(p, gray, 3, 8, palette=pal)
    _, _, typ, px = run_img(io_bin, p, READ_COLOR, tmp_path)
    assert np.array_equal(px.reshape(20, 33, 3), pal[gray[:, :, 0]][:, :, ::-1])
    # BMP: 24-bit with row padding, 8-bit gray palette (the masks of tests 5-8)
    b = str(tmp_path / "c.bmp")
    write_bmp24(b, bgr)
    _, _, typ, px = run_img(io_bin, b, READ_COLOR, tmp_path)
    assert typ == U8C3 and np.array_equal(px.reshape(37, 61, 3), bgr)
    mask = (rng.integers(0, 2, (20, 33)) * 255).astype(np.uint8)
    write_bmp8(b, mask)
    _, _, typ, px = run_img(io_bin, b, READ_GRAYSCALE, tmp_path)
    assert typ == U8C1 and np.array_equal(px.reshape(20, 33), mask)
    # colour -> gray follows OpenCV's fixed-point BGR2GRAY
    write_bmp24(b, bgr)
    _, _, typ, px = run_img(io_bin, b, READ_GRAYSCALE, tmp_path)
    want = ((bgr[:, :, 0].astype(np.int64) * 1868 + bgr[:, :, 1].astype(np.int64) * 9617 + bgr[:, :, 2].astype(np.int64) * 4899 + 8192) >> 14).astype(np.uint8)
    assert typ == U8C1 and np.array_equal(px.reshape(37, 61), want)
    # PNG writer round trip (colour and 16-bit)
    o = str(tmp_path / "o.png")
    subprocess.run([io_bin, "repng", d, str(READ_ANYDEPTH), o], check=True)
    _, _, typ, px = run_img(io_bin, o, READ_ANYDEPTH, tmp_path)
    assert typ == U16C1 and np.array_equal(px.view(np.uint16).reshape(37, 61), depth)
    # errors are reported, not crashes
    open(p, "wb").write(b"not an image")
    assert subprocess.run([io_bin, "img", p, "0", o], capture_output=True).returncode == 1
    # malformed headers are refused before they index or size anything: an 8-bit BMP whose header size points the palette
    # past the end of the file, one whose palette overlaps the pixel data, and a PNG announcing absurd dimensions
    write_bmp8(b, mask)
    good = bytearray(open(b, "rb").read())
    bad = bytearray(good); bad[14:18] = (0x7fffff00).to_bytes(4, "little")
    open(b, "wb").write(bad)
    assert subprocess.run([io_bin, "img", b, str(READ_GRAYSCALE), o], capture_output=True).returncode == 1
    bad = bytearray(good); bad[10:14] = (54 + 16).to_bytes(4, "little")          # pixel data starts inside the palette
    open(b, "wb").write(bad)
    assert subprocess.run([io_bin, "img", b, str(READ_GRAYSCALE), o], capture_output=True).returncode == 1
    write_png(p, gray, 0, 8)
    png = bytearray(open(p, "rb").read())
    png[16:20] = (0x40000000).to_bytes(4, "big")                                 # IHDR width (the CRC is not checked)
    open(p, "wb").write(png)
    assert subprocess.run([io_bin, "img", p, "0", o], capture_output=True).returncode == 1


def test_ply_obj_pose_overlay(io_bin, tmp_path):
    rng = np.random.default_rng(3)
    xyz = rng.normal(size=(50, 3)).astype(np.float32)
    extra = rng.integers(0, 256, (50, 3), dtype=np.uint8)
    # ascii PLY with colour properties and a face element after the vertices
    a = str(tmp_path / "a.ply")
    with open(a, "w") as f:
        f.write("ply\nformat ascii 1.0\ncomment made by a test\nelement vertex 50\nproperty float x\nproperty float y\nproperty float z\n"
                "property uchar red\nproperty uchar green\nproperty uchar blue\nelement face 1\nproperty list uchar int vertex_indices\nend_header\n")
        for p, c in zip(xyz, extra):
            f.write("%.9g %.9g %.9g %d %d %d\n" % (p[0], p[1], p[2], c[0], c[1], c[2]))
        f.write("3 0 1 2\n")
    out = str(tmp_path / "p.raw")
    r = subprocess.run([io_bin, "ply", a, out], check=True, capture_output=True, text=True)
    assert int(r.stdout) == 50 and np.array_equal(np.fromfile(out, np.float32).reshape(50, 3), xyz)
    # binary PLY: double y, interleaved extra property
    b = str(tmp_path / "b.ply")
    with open(b, "wb") as f:
        f.write(b"ply\nformat binary_little_endian 1.0\nelement vertex 50\nproperty float x\nproperty short q\nproperty double y\nproperty float z\nend_header\n")
        for p in xyz:
            f.write(struct.pack("<fhdf", p[0], -7, float(p[1]), p[2]))
    r = subprocess.run([io_bin, "ply", b, out], check=True, capture_output=True, text=True)
    assert int(r.stdout) == 50 and np.array_equal(np.fromfile(out, np.float32).reshape(50, 3), xyz)
    # OBJ writer: "v x y z", no trailing newline, homogeneous stride
    xyz1 = np.concatenate([xyz, np.ones((50, 1), np.float32)], 1)
    xyz1.tofile(out)
    o = str(tmp_path / "c.obj")
    subprocess.run([io_bin, "obj", out, "50", "4", o], check=True)
    text = open(o).read()
    assert not text.endswith("\n") and len(text.split("\n")) == 50
    back = np.array([[float(t) for t in line.split()[1:]] for line in text.split("\n")])
    assert all(line.startswith("v ") for line in text.split("\n")) and np.allclose(back, xyz, rtol=1e-5)   # default ostream precision: 6 significant digits, as the reference
    # poses: matrix, round trip, pretty printer (YPR in degrees, PoseManipUtils.cpp:82-98,148-158)
    ang = np.radians([20.0, -35.0, 10.0])    # yaw, pitch, roll
    cy, sy, cp, sp, cr, sr = np.cos(ang[0]), np.sin(ang[0]), np.cos(ang[1]), np.sin(ang[1]), np.cos(ang[2]), np.sin(ang[2])
    R = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]]) @ np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]]) @ np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    qw = np.sqrt(1 + np.trace(R)) / 2
    q = np.array([qw, (R[2, 1] - R[1, 2]) / (4 * qw), (R[0, 2] - R[2, 0]) / (4 * qw), (R[1, 0] - R[0, 1]) / (4 * qw)])
    pose = np.concatenate([q, [0.25, -1.5, 3.0]])
    r = subprocess.run([io_bin, "pose"] + ["%.17g" % v for v in pose], check=True, capture_output=True, text=True).stdout.split("\n")
    assert r[0] == ":YPR=(20.00,-35.00,10.00)  :TxTyTz=(0.25,-1.50,3.00)"
    T = np.array([float(v) for v in r[1].split()]).reshape(4, 4)
    assert np.allclose(T[:3, :3], R, atol=1e-14) and np.allclose(T[:3, 3], pose[4:]) and np.array_equal(T[3], [0, 0, 0, 1])
    assert np.allclose([float(v) for v in r[2].split()], pose, atol=1e-14)
    # reproject + overlay: red pixels where the points project, points outside the image skipped
    img = np.full((30, 40, 3), 90, np.uint8)
    p = str(tmp_path / "i.png")
    write_png(p, img, 2, 8)
    pts = np.array([[0.0, 0.0, 2.0], [0.5, 0.25, 2.0], [-0.9, 0.0, 2.0], [5.0, 5.0, 1.0]], np.float32)   # the last one leaves the image
    pts.tofile(out)
    K = [20.0, 20.0, 19.5, 14.5]
    ident = [1, 0, 0, 0, 0, 0, 0]
    o = str(tmp_path / "ov.png")
    subprocess.run([io_bin, "overlay", p, out, "4"] + [str(v) for v in ident] + [str(v) for v in K] + [o], check=True)
    raw = str(tmp_path / "ov.raw")
    subprocess.run([io_bin, "img", o, "0", raw], check=True, capture_output=True)
    got = np.fromfile(raw, np.uint8).reshape(30, 40, 3)
    want = img.copy()
    for X, Y, Z in pts[:3]:
        want[int(K[1] * Y / Z + K[3]), int(K[0] * X / Z + K[2])] = (0, 0, 255)
    assert np.array_equal(got, want)

"""The two data formats either side of the path, on the CPU through libalvrl_host.so (csrc/hostio.h is the code behind
alvrl_set_medium_grid_file and alvrl_film_write_npy):
  grid volume files  src/volume/gridvolume.cpp:217-287 ("VOL", version 3; float32 / uint8 density, the uint8 density map)
  NumPy film output  src/films/mfilm.cpp:337-348 + src/films/cnpy.h:207-236"""
import ctypes as C
import struct

import numpy as np
import pytest


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _read(host_lib, path, voxels=True):
    hdr = np.zeros(11, np.int32)
    err = C.create_string_buffer(256)
    rc = host_lib.alvrl_host_read_vol(str(path).encode(), _p(hdr), None, err, C.c_uint32(256))
    if rc != 0:
        return rc, err.value.decode(), None
    vtype, xr, yr, zr, ch = (int(v) for v in hdr[:5])
    box = hdr[5:].view(np.float32).copy()
    d = None
    if voxels:
        d = np.zeros((zr, yr, xr), np.float32)
        rc = host_lib.alvrl_host_read_vol(str(path).encode(), _p(hdr), _p(d), err, C.c_uint32(256))
        if rc != 0:
            return rc, err.value.decode(), None
    return 0, dict(type=vtype, res=(xr, yr, zr), channels=ch, bmin=box[:3], bmax=box[3:]), d


def test_float32_volume_round_trip(pkg, host_lib, tmp_path):
    rng = np.random.default_rng(1)
    d = rng.random((5, 7, 9), dtype=np.float32)                        # [z][y][x]: ragged resolution 9 x 7 x 5
    path = tmp_path / "fog.vol"
    pkg.volfile.write_vol(path, d, (-1, 0, 0.5), (2, 1, 3.25))
    rc, hdr, got = _read(host_lib, path)
    assert rc == 0, hdr
    assert hdr["type"] == pkg.volfile.VOL_FLOAT32 and hdr["res"] == (9, 7, 5) and hdr["channels"] == 1
    assert np.array_equal(hdr["bmin"], np.float32([-1, 0, 0.5])) and np.array_equal(hdr["bmax"], np.float32([2, 1, 3.25]))
    assert np.array_equal(got, d)                                      # bit for bit, x fastest
    d2, mn, mx, t = pkg.volfile.read_vol(path)                         # the numpy reader agrees
    assert np.array_equal(d2, d) and np.array_equal(mn, hdr["bmin"]) and np.array_equal(mx, hdr["bmax"]) and t == 1


def test_uint8_volume_goes_through_the_reference_density_map(pkg, host_lib, tmp_path):
    """gridvolume.cpp:212-215: m_densityMap[i] = i / 255.0f, [255] = 1; the trilinear lookup reads the mapped values
    (374-389), so the float grid the reader hands to the device holds exactly them"""
    q = np.arange(256, dtype=np.uint8).reshape(4, 8, 8)
    path = tmp_path / "smoke8.vol"
    pkg.volfile.write_vol(path, q, (0, 0, 0), (1, 1, 1), pkg.volfile.VOL_UINT8)
    rc, hdr, got = _read(host_lib, path)
    assert rc == 0 and hdr["type"] == pkg.volfile.VOL_UINT8 and hdr["res"] == (8, 8, 4)
    want = (np.arange(256, dtype=np.float32) / np.float32(255.0)).reshape(4, 8, 8)
    assert np.array_equal(got, want) and got.reshape(-1)[255] == 1.0
    # float densities are quantised by the writer to the nearest map entry
    d = np.random.default_rng(2).random((3, 4, 5))
    pkg.volfile.write_vol(path, d, (0, 0, 0), (1, 1, 1), pkg.volfile.VOL_UINT8)
    _, _, got = _read(host_lib, path)
    assert np.abs(got - d).max() <= 0.5 / 255 + 1e-7


def _raw(path, ident=b"VOL", version=3, vtype=1, res=(2, 2, 2), channels=1, voxel_bytes=None):
    n = res[0] * res[1] * res[2] * channels
    body = voxel_bytes if voxel_bytes is not None else b"\0" * (n * (4 if vtype == 1 else 2 if vtype in (2, 4) else 1))
    with open(path, "wb") as f:
        f.write(ident + bytes([version]) + struct.pack("<iiiii", vtype, *res, channels) + struct.pack("<6f", 0, 0, 0, 1, 1, 1) + body)


@pytest.mark.parametrize("kw,code,needle", [
    (dict(ident=b"VOX"), -1, "incorrect header identifier"),                         # gridvolume.cpp:227-229
    (dict(version=2), -1, "incorrect file version"),                                 # 232-234
    (dict(vtype=2), -5, "float16 volumes are not yet supported"),                     # 253-255
    (dict(vtype=7), -1, "unknown type (type=7, channels=1)"),                         # 269-270
    (dict(channels=2), -1, "only 1 and 3 channels are supported"),                    # 246-250
    (dict(channels=3), -5, "one-channel volume"),                                    # a density is a lookupFloat
    (dict(vtype=4, channels=3), -5, "quantized direction"),
    (dict(voxel_bytes=b"\0" * 12), -4, "fewer voxels than the header announces"),
    (dict(res=(0, 4, 4)), -1, "resolution out of range"),
], ids=["identifier", "version", "float16", "unknown-type", "two-channels", "three-channels", "qdir", "truncated", "empty-axis"])
def test_invalid_volume_files_are_refused_with_the_reference_messages(host_lib, tmp_path, kw, code, needle):
    path = tmp_path / "bad.vol"
    _raw(path, **kw)
    rc, msg, _ = _read(host_lib, path)
    assert rc == code and needle in msg, (rc, msg)


def test_missing_and_short_files(host_lib, tmp_path):
    rc, msg, _ = _read(host_lib, tmp_path / "nothing.vol")
    assert rc == -4 and "cannot open" in msg
    (tmp_path / "short.vol").write_bytes(b"VOL\x03\x01\0\0\0")
    rc, msg, _ = _read(host_lib, tmp_path / "short.vol")
    assert rc == -4 and "truncated header" in msg


def test_header_only_query_does_not_need_the_voxels(host_lib, tmp_path):
    path = tmp_path / "big.vol"
    _raw(path, res=(512, 512, 512), voxel_bytes=b"")                    # a header that announces 512^3 voxels, none present
    rc, hdr, _ = _read(host_lib, path, voxels=False)
    assert rc == 0 and hdr["res"] == (512, 512, 512)


@pytest.mark.parametrize("shape", [(3, 5, 3), (1, 1, 3), (17, 251, 3), (4, 6, 1)])
def test_npy_film_output_loads_in_numpy_and_has_the_cnpy_header_layout(host_lib, tmp_path, shape):
    H, W, ch = shape
    img = np.random.default_rng(H * W).random(shape, dtype=np.float32)
    path = tmp_path / "film.npy"
    err = C.create_string_buffer(256)
    assert host_lib.alvrl_host_write_npy(str(path).encode(), _p(img), C.c_uint32(H), C.c_uint32(W), C.c_uint32(ch), err, C.c_uint32(256)) == 0
    back = np.load(path)
    assert back.dtype == np.float32 and back.shape == ((H, W) if ch == 1 else shape)        # mfilm.cpp:343-344
    assert np.array_equal(back.reshape(shape), img)
    raw = path.read_bytes()
    assert raw[:8] == b"\x93NUMPY\x01\x00"
    hlen = struct.unpack("<H", raw[8:10])[0]
    assert (10 + hlen) % 16 == 0 and raw[10 + hlen - 1:10 + hlen] == b"\n"                   # cnpy.h:222-225
    shp = f"({H}, {W})" if ch == 1 else f"({H}, {W}, {ch})"
    assert raw[10:10 + hlen].decode().rstrip() == "{'descr': '<f4', 'fortran_order': False, 'shape': " + shp + ", }"
    assert len(raw) == 10 + hlen + 4 * H * W * ch
    assert host_lib.alvrl_host_write_npy(str(tmp_path / "no" / "dir.npy").encode(), _p(img), C.c_uint32(H), C.c_uint32(W), C.c_uint32(ch), err, C.c_uint32(256)) == -4


@pytest.mark.parametrize("uint8", [False, True], ids=["float32", "uint8"])
def test_library_reader_agrees_with_the_oracles_restatement_of_loadFromFile(pkg, orc, host_lib, tmp_path, uint8):
    """two independent readers: the library's (csrc/hostio.h, through libalvrl_host.so) and the oracle's restatement of
    GridDataSource::loadFromFile + the uint8 density map (oracle_capi.cpp: orc_set_medium_grid_file).  A medium set from the
    file on the oracle == a medium set from the array the library's reader returns, on 2 000 transmittance queries."""
    from conftest import small_case
    scene, vrls, params = small_case(pkg, "C3", 16, 16, 16, grid=20)
    m = scene["medium"]
    path = tmp_path / "density.vol"
    pkg.volfile.write_vol(path, m["density"], (0.05, 0, 0.1), (0.9, 1, 1), pkg.volfile.VOL_UINT8 if uint8 else pkg.volfile.VOL_FLOAT32)
    rc, hdr, dens = _read(host_lib, path)
    assert rc == 0
    a = orc.Oracle(**params); a.set_scene(scene)
    a.set_medium_grid_file(str(path), m["scale"], m["albedo"], m["sigmaS_base"])
    b = orc.Oracle(**params); b.set_scene(scene)
    b.set_medium_grid(dens, hdr["bmin"], hdr["bmax"], m["scale"], m["albedo"], m["sigmaS_base"])
    rng = np.random.default_rng(8)
    p1, p2 = rng.uniform(0.02, 0.98, (2000, 3)).astype(np.float32), rng.uniform(0.02, 0.98, (2000, 3)).astype(np.float32)
    s = np.zeros(2000, np.int32)
    ta, tb = a.eval_transmittance(p1, s, p2), b.eval_transmittance(p1, s, p2)
    assert np.array_equal(ta, tb) and 0 < ta[ta > 0].min() < ta.max() <= 1 and (ta > 0).mean() > 0.5      # (zeros: occluded segments)
    if not uint8:
        c = orc.Oracle(**params); c.set_scene(scene)
        c.set_medium_grid(m["density"], (0.05, 0, 0.1), (0.9, 1, 1), m["scale"], m["albedo"], m["sigmaS_base"])
        assert np.array_equal(ta, c.eval_transmittance(p1, s, p2))


def _read_vrls(host_lib, path):
    n = C.c_uint32()
    rc = host_lib.alvrl_host_read_vrl_file(str(path).encode(), None, C.byref(n))
    if rc != 0:
        return rc, None
    out = np.zeros((n.value, 9), np.float32)
    assert host_lib.alvrl_host_read_vrl_file(str(path).encode(), _p(out), C.byref(n)) == 0
    return 0, out


def test_vrl_file_reader_follows_the_reference_constructor(pkg, host_lib, tmp_path):
    """VRL.h:43-54 + 120-128: nine numbers per line; reading ends at the end of the file, at a line without nine numbers, and
    -- silently, keeping what was read -- at a VRL whose power is not valid (the constructor's Log(EError) lands in the reader's
    own catch block).  Zero power / zero length lines are read here and dropped later by the put() filter."""
    s, e, p, _ = pkg.scenes.synthetic_vrls(40, sigma_t=1.05, seed=4)
    path = tmp_path / "set.vrl"
    pkg.scenes.write_vrl_file(path, s, e, p)
    rc, got = _read_vrls(host_lib, path)
    assert rc == 0 and np.array_equal(got, np.concatenate([s, e, p], 1))              # repr() round-trips every float32
    # CR LF line ends, extra columns, exponents, a last line without a newline
    path.write_text("0 0 0 1 0 0 1 1 1\r\n1e-1 2E-1 .3 4 5 6 7 8 9 10 11\r\n-1 -2 -3 0.5 0.5 0.5 0 0 0")
    rc, got = _read_vrls(host_lib, path)
    assert rc == 0 and len(got) == 3 and np.array_equal(got[1], np.float32([0.1, 0.2, 0.3, 4, 5, 6, 7, 8, 9])) and not got[2, 6:].any()
    # the reading ends at: a negative power, a NaN, a short line, an empty line
    for bad in ("0 0 0 1 1 1 1 -1 1", "0 0 0 1 1 1 nan 1 1", "0 0 0 1 1 1 1 1", "", "x y z"):
        path.write_text("0 0 0 1 0 0 1 1 1\n0 0 0 2 0 0 1 1 1\n" + bad + "\n0 0 0 3 0 0 1 1 1\n")
        rc, got = _read_vrls(host_lib, path)
        assert rc == 0 and len(got) == 2, (bad, got)
    path.write_text("")
    rc, got = _read_vrls(host_lib, path)
    assert rc == 0 and len(got) == 0
    assert _read_vrls(host_lib, tmp_path / "missing.vrl")[0] == -4


def test_file_readers_survive_corrupt_input(pkg, host_lib, tmp_path):
    """hypothesis: truncated, bit-flipped and random files never crash the readers and never make them allocate what the file
    cannot hold -- a header that announces 16384^3 voxels in a 60-byte file is refused from the file's length"""
    from hypothesis import given, settings, strategies as st
    good = tmp_path / "good.vol"
    pkg.volfile.write_vol(good, np.random.default_rng(0).random((3, 4, 5), dtype=np.float32), (0, 0, 0), (1, 1, 1))
    base = good.read_bytes()
    huge = tmp_path / "huge.vol"
    huge.write_bytes(b"VOL\x03" + struct.pack("<iiiii", 1, 16384, 16384, 16384, 1) + struct.pack("<6f", 0, 0, 0, 1, 1, 1) + b"\0" * 12)
    rc, hdr, _ = _read(host_lib, huge, voxels=False)
    assert rc == 0 and hdr["res"] == (16384, 16384, 16384)               # the header alone is well formed ...
    one, err, h11 = np.zeros(1, np.float32), C.create_string_buffer(256), np.zeros(11, np.int32)
    assert host_lib.alvrl_host_read_vol(str(huge).encode(), _p(h11), _p(one), err, C.c_uint32(256)) == -4 and b"fewer voxels" in err.value
    path = tmp_path / "fuzz.bin"

    @settings(max_examples=300, deadline=None, derandomize=True, database=None)
    @given(st.integers(0, len(base)), st.lists(st.tuples(st.integers(0, len(base) - 1), st.integers(0, 255)), max_size=6), st.binary(max_size=80))
    def check(cut, flips, tail):
        b = bytearray(base[:cut])
        for pos, val in flips:
            if pos < len(b):
                b[pos] = val
        path.write_bytes(bytes(b) + tail)
        hdr = np.zeros(11, np.int32)
        err = C.create_string_buffer(256)
        rc = host_lib.alvrl_host_read_vol(str(path).encode(), _p(hdr), None, err, C.c_uint32(256))
        assert rc in (0, -1, -4, -5)
        if rc == 0 and int(hdr[1]) * int(hdr[2]) * int(hdr[3]) <= 1 << 22:
            d = np.zeros(int(hdr[1]) * int(hdr[2]) * int(hdr[3]), np.float32)
            assert host_lib.alvrl_host_read_vol(str(path).encode(), _p(hdr), _p(d), err, C.c_uint32(256)) in (0, -4)
        n = C.c_uint32()
        assert host_lib.alvrl_host_read_vrl_file(str(path).encode(), None, C.byref(n)) == 0        # a VRL file reader stops at the first bad line
    check()

"""CPU tests of the file formats at the drop-in boundary (apps/er_io.h through bin/er_io_tool) and of the Python
generators, including round trips of the reference's own fixtures when /root/reference is present."""
import os
import struct
import subprocess

import numpy as np
import pytest

from elasticreconstruction_b200 import build, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOOL = os.path.join(ROOT, "elasticreconstruction_b200", "bin", "er_io_tool")
REF_DATA = "/root/reference/Matlab_Toolbox/Example/Data"


@pytest.fixture(scope="module", autouse=True)
def tool():
    if not os.path.exists(TOOL):
        build.build_apps()
    assert os.path.exists(TOOL)


def parse_log(path):
    L = [l for l in open(path) if not l.startswith("#")]
    n = len(L) // 5
    ids = np.array([[int(x) for x in L[5 * i].split()] for i in range(n)])
    T = np.array([[[float(x) for x in L[5 * i + 1 + r].split()] for r in range(4)] for i in range(n)])
    return ids, T


def test_log_roundtrip_synthetic(tmp_path):
    P = synth.quantize_log(synth.loop_trajectory(17))
    a, b = str(tmp_path / "a.log"), str(tmp_path / "b.log")
    synth.write_log(a, P)
    with open(a) as f:
        body = f.read()
    with open(a, "w") as f:
        f.write("# comment line\n" + body)                 # '#' lines are skipped (TSDFVolumeUnit.h:35)
    out = subprocess.check_output([TOOL, "log2log", a, b], text=True)
    assert int(out) == 17
    ids, T = parse_log(b)
    assert np.array_equal(T, P) and np.array_equal(ids[:, 2], np.arange(1, 18))


@pytest.mark.skipif(not os.path.isdir(REF_DATA), reason="reference fixtures not present")
@pytest.mark.parametrize("rel", ["Trajectory/traj_gt.log", "Trajectory/traj_rigid.log",
                                 "RegistrationEvaluation/livingroom1/gt.log"])
def test_reference_log_fixtures_parse(tmp_path, rel):
    src = os.path.join(REF_DATA, rel)
    out = str(tmp_path / "o.log")
    n = int(subprocess.check_output([TOOL, "log2log", src, out], text=True))
    L = [l for l in open(src) if l.strip() and not l.startswith("#")]
    assert n == len(L) // 5 and n > 10
    ids, T = parse_log(out)
    ref = np.array([[float(x) for x in L[5 * 3 + 1 + r].split()] for r in range(4)])
    assert np.allclose(T[3], ref, atol=5e-9)                # written with %.8f
    assert np.allclose(T[:, 3, :], [0, 0, 0, 1])


@pytest.mark.skipif(not os.path.isdir(REF_DATA), reason="reference fixtures not present")
def test_reference_info_fixture_roundtrip(tmp_path):
    src = os.path.join(REF_DATA, "RegistrationEvaluation/livingroom1/gt.info")
    out = str(tmp_path / "o.info")
    n = int(subprocess.check_output([TOOL, "info2info", src, out], text=True))
    assert n > 10
    L = open(out).read().split("\n")
    blk = np.array([[float(x) for x in L[1 + r].split()] for r in range(6)])
    assert blk[0, 0] == 3305 and abs(blk[0, 4] - 10048.0068) < 1e-3 and abs(blk[1, 3] + 10048.0068) < 1e-3


def lzf_literal_stream(raw: bytes) -> bytes:
    """A valid LZF stream made only of literal runs (<= 32 bytes each)."""
    out = bytearray()
    for i in range(0, len(raw), 32):
        chunk = raw[i:i + 32]
        out.append(len(chunk) - 1)
        out += chunk
    return bytes(out)


def lzf_with_backrefs(raw: bytes) -> bytes:
    """Greedy LZF encoder with short back references, to exercise the copy path of the decoder."""
    out = bytearray()
    i, lit = 0, bytearray()

    def flush():
        nonlocal lit
        for k in range(0, len(lit), 32):
            c = lit[k:k + 32]
            out.append(len(c) - 1)
            out.extend(c)
        lit = bytearray()
    while i < len(raw):
        best_len, best_off = 0, 0
        for off in range(1, min(i, 64) + 1):
            l = 0
            while l < 40 and i + l < len(raw) and raw[i + l] == raw[i - off + l]:
                l += 1
            if l > best_len:
                best_len, best_off = l, off
        if best_len >= 3:
            flush()
            l = best_len - 2
            o = best_off - 1
            if l < 7:
                out.append((l << 5) | (o >> 8))
            else:
                out.append((7 << 5) | (o >> 8))
                out.append(l - 7)
            out.append(o & 0xFF)
            i += best_len
        else:
            lit.append(raw[i])
            i += 1
    flush()
    return bytes(out)


@pytest.mark.parametrize("mode", ["binary", "ascii", "binary_compressed", "binary_compressed_refs"])
def test_pcd_reader_all_encodings(tmp_path, mode):
    rng = np.random.default_rng(11)
    n = 300
    xyz = rng.normal(size=(n, 3)).astype(np.float32)
    nrm = rng.normal(size=(n, 3)).astype(np.float32)
    nrm[::7, 0] = np.nan                                  # dropped by the NaN-normal filter (CorresApp.cpp:94)
    path = str(tmp_path / "c.pcd")
    fields = "x y z rgb normal_x normal_y normal_z curvature"
    hdr = ("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS %s\nSIZE 4 4 4 4 4 4 4 4\n"
           "TYPE F F F F F F F F\nCOUNT 1 1 1 1 1 1 1 1\nWIDTH %d\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS %d\n" % (fields, n, n))
    cols = [xyz[:, 0], xyz[:, 1], xyz[:, 2], np.zeros(n, np.float32), nrm[:, 0], nrm[:, 1], nrm[:, 2], np.zeros(n, np.float32)]
    if mode == "binary":
        synth.write_pcd_xyzrgbnormal(path, xyz, nrm)
    elif mode == "ascii":
        with open(path, "w") as f:
            f.write(hdr + "DATA ascii\n")
            for i in range(n):
                f.write(" ".join(("nan" if np.isnan(c[i]) else "%.9g" % c[i]) for c in cols) + "\n")
    else:
        soa = b"".join(np.ascontiguousarray(c, dtype="<f4").tobytes() for c in cols)     # field-major after decompression
        comp = lzf_literal_stream(soa) if mode == "binary_compressed" else lzf_with_backrefs(soa)
        with open(path, "wb") as f:
            f.write((hdr + "DATA binary_compressed\n").encode())
            f.write(struct.pack("<II", len(comp), len(soa)))
            f.write(comp)
    out = str(tmp_path / "c.xyzn")
    total, kept = [int(x) for x in subprocess.check_output([TOOL, "pcd2xyzn", path, out], text=True).split()]
    keep = ~np.isnan(nrm[:, 0])
    assert total == n and kept == int(keep.sum())
    got = np.loadtxt(out, dtype=np.float64).reshape(-1, 6)
    want = np.concatenate([xyz[keep], nrm[keep]], axis=1)
    assert np.array_equal(got.astype(np.float32), want)        # %.9g round-trips a float32 exactly


def test_ctr_and_camera_loaders(tmp_path):
    g = synth.quantize_ctr(synth.make_control_grids(2, 8, 3.0))
    p = str(tmp_path / "g.ctr")
    synth.write_ctr(p, g)
    o = str(tmp_path / "g.txt")
    subprocess.check_call([TOOL, "ctr2txt", p, "2", "8", o])
    got = np.loadtxt(o, dtype=np.float32).reshape(2, -1, 3)
    assert np.array_equal(got, g)
    c = str(tmp_path / "cam.param")
    synth.write_camera(c, 500.0, 510.0, 310.25, 245.75, 2.5, 1.3)
    vals = [float(x) for x in subprocess.check_output([TOOL, "camera", c], text=True, stderr=subprocess.DEVNULL).split()]
    assert vals == [500.0, 510.0, 310.25, 245.75, 2.5, np.float32(1.3)]


def test_depth_container_and_world_pcd_roundtrip(tmp_path):
    fr = synth.wavy_wall_depth(2)
    p = str(tmp_path / "d.raw")
    synth.write_depth_container(p, fr)
    raw = open(p, "rb").read()
    assert struct.unpack("<iii", raw[:12]) == (2, 640, 480)
    assert np.array_equal(np.frombuffer(raw[12:], dtype="<u2").reshape(2, 480, 640), fr)

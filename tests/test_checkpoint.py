"""CPU: `.rr` checkpoint / resume: SaveToFile / LoadFromFile / LoadingModel of the host library against a file
written by the reference's own code (tests/golden/ref_area0.5_lp64.rr) and, live, against oracle/_ref."""
import ctypes
import os

import numpy as np

from util import sha, seeded_radiosity

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731


def _same(a, b):
    return all((x.view(np.uint32) == y.view(np.uint32)).all() for x, y in zip(a, b))


def test_reads_the_reference_dump(api, golden):
    g = golden["reference"]["rr"]
    path = os.path.join(ROOT, "tests", "golden", g["file"])
    assert os.path.getsize(path) == g["bytes"]
    s = api.Scene(0.5)
    fresh = s.arrays(); nb = s.neighbours()
    s.set_state(np.zeros((s.P, 3), np.float32), np.zeros((s.P, 3), np.float32))
    s.load(path)
    v, ix, c, r, il = s.arrays()
    assert s.P == g["P"] and sha(r) == g["rad_sha256"] and sha(il) == g["illum_sha256"]
    assert _same((v, ix, c), fresh[:3]) and (s.neighbours() == nb).all()


def test_round_trip_all_formats(api, tmp_path):
    s = api.Scene(0.05)
    rad = seeded_radiosity(s.P, 5); ill = seeded_radiosity(s.P, 6)
    s.set_state(rad, ill)
    want = s.arrays(); nb = s.neighbours()
    sizes = {0: 16 + 116 * s.P, 1: 4 + 148 * s.P, 2: 8 + 184 * s.P}
    for fmt in (0, 1, 2):
        p = str(tmp_path / f"scene{fmt}.rr")
        s.save(p, fmt)
        assert os.path.getsize(p) == sizes[fmt]
        t = api.Scene(0.5)
        t.load(p)
        assert t.P == s.P and _same(t.arrays(), want) and (t.neighbours() == nb).all()
        # a loaded scene is not subdivided again (LoadingModel::getPatches ignores the area)
        assert t.P == s.P


def test_rejects_garbage(api, tmp_path):
    p = tmp_path / "bad.rr"
    p.write_bytes(b"not a checkpoint at all")
    s = api.Scene(0.5)
    try:
        s.load(str(p))
        assert False
    except api.RadError:
        pass
    assert s.P == 502                                     # scene untouched


def test_rejects_crafted_counts(api, tmp_path):
    """a count field chosen so that count * record size wraps modulo 2^64 onto the file size must be refused, not allocated"""
    import struct
    s = api.Scene(0.5)
    for name, blob in (
            ("wrap_portable.rr", b"RRB2" + struct.pack("<IQ", 1, (1 << 64) // 104 + 1) + b"\0" * 8),      # 16 + count * 104 wraps
            ("wrap_lp64.rr", struct.pack("<Q", ((1 << 64) + 184 * 2) // 184 + (1 << 61)) + b"\0" * 368),
            ("short.rr", struct.pack("<Q", 3) + b"\0" * 184)):
        p = tmp_path / name
        p.write_bytes(blob)
        try:
            s.load(str(p))
            assert False, name
        except api.RadError:
            pass
        assert s.P == 502


def test_reads_win64_layout(api, tmp_path):
    """LLP64 dump of the reference (MSVC x64): 4-byte count, 184-byte records — same fields as the LP64 file"""
    import struct
    s = api.Scene(0.5)
    lp = tmp_path / "lp64.rr"
    s.save(str(lp), 2)
    raw = lp.read_bytes()
    (count,) = struct.unpack("<Q", raw[:8])
    w = tmp_path / "win64.rr"
    w.write_bytes(struct.pack("<I", count) + raw[8:])
    t = api.Scene(0.5)
    t.load(str(w))
    assert t.P == s.P and _same(t.arrays(), s.arrays()) and (t.neighbours() == s.neighbours()).all()


def test_live_exchange_with_reference_build(api, ref, tmp_path):
    """my LP64 file -> the reference's LoadingModel; the reference's dump -> my loader."""
    s = api.Scene(0.02)
    rad = seeded_radiosity(s.P, 7); ill = seeded_radiosity(s.P, 8)
    s.set_state(rad, ill)
    mine = str(tmp_path / "mine.rr")
    s.save(mine, 2)
    assert ref.refp_load_rr(mine.encode()) == s.P
    P = s.P
    rv = np.zeros((P, 12), np.float32); rc = np.zeros((P, 3), np.float32); rr = np.zeros((P, 3), np.float32); ri = np.zeros((P, 3), np.float32)
    ref.refp_scene_get(vp(rv), None, vp(rc), vp(rr), vp(ri))
    v, ix, c, r, il = s.arrays()
    assert _same((rv, rc, rr, ri), (v, c, r, il))
    nb = np.zeros((P, 8), np.int32); ref.refp_neighbours(vp(nb))
    assert (nb == s.neighbours()).all()
    theirs = str(tmp_path / "theirs.rr")
    assert ref.refp_save_rr(theirs.encode())
    t = api.Scene(0.5); t.load(theirs)
    assert _same(t.arrays(), (v, ix, c, r, il))

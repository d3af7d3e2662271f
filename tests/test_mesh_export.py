"""Shaded-mesh export (radiosity_b200/host/MeshExport.*): what the reference draws after a run (OnIdle,
Main.cpp:1318-1366 — Colors::smoothShadePatch vertex colours on the scene's quads, colours clamped to [0, 1] by GL) as a
binary PLY.  CPU: the file against the scene arrays and the host Colors path; GPU: the driver's `ply` key."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def read_ply(path):
    with open(path, "rb") as f:
        data = f.read()
    head, body = data.split(b"end_header\n", 1)
    lines = head.decode().splitlines()
    assert lines[0] == "ply" and lines[1] == "format binary_little_endian 1.0"
    nv = int([l for l in lines if l.startswith("element vertex")][0].split()[-1])
    nf = int([l for l in lines if l.startswith("element face")][0].split()[-1])
    vt = np.dtype([("xyz", "<f4", 3), ("rgb", "u1", 3)])
    ft = np.dtype([("n", "u1"), ("idx", "<i4", 4)])
    assert len(body) == nv * vt.itemsize + nf * ft.itemsize
    v = np.frombuffer(body, vt, nv)
    fc = np.frombuffer(body, ft, nf, offset=nv * vt.itemsize)
    return v, fc


def test_ply_matches_scene_and_host_colours(api, tmp_path):
    scene = api.Scene(0.5)
    verts, _, col, rad, illum = scene.arrays()
    colors = scene.smooth_shade()                      # fresh scene: only the light patches are lit (I = 1)
    path = str(tmp_path / "box.ply")
    scene.export_ply(path, colors)
    v, f = read_ply(path)
    P = scene.P
    assert len(v) == 4 * P and len(f) == P
    assert (v["xyz"].reshape(P, 12) == verts).all()
    assert (f["n"] == 4).all() and (f["idx"] == np.arange(4 * P, dtype=np.int32).reshape(P, 4)).all()
    exp = np.rint(np.clip(colors.reshape(-1, 3), 0.0, 1.0) * 255.0).astype(np.uint8)
    assert (v["rgb"] == exp).all()
    assert v["rgb"].max() == 255 and v["rgb"].min() == 0
    # exposure scales before the clamp; NaN and negative colours come out black
    c2 = colors.copy(); c2[0, :3] = [np.nan, -1.0, 0.5]
    scene.export_ply(path, c2, exposure=0.5)
    v2, _ = read_ply(path)
    assert tuple(v2["rgb"][0]) == (0, 0, 64)
    with pytest.raises(api.RadError):
        scene.export_ply(str(tmp_path / "no_such_dir" / "x.ply"), colors)
    with pytest.raises(api.RadError):
        scene.export_ply(path, colors[:5])


@pytest.mark.gpu
def test_driver_writes_the_shaded_mesh(api, tmp_path):
    """`radiosity ... ply out.ply`: vertex colours from the display stage on the device (K5) == the host Colors path applied
    to the energies the same run dumped."""
    exe = os.path.join(ROOT, "radiosity_b200", "radiosity")
    ply = str(tmp_path / "out.ply"); ckpt = str(tmp_path / "out.rr")
    p = subprocess.run([exe, "area", "0.5", "hemicube", "64", "hemicubes", "4", "shoots", "5", "shots", "10", "ply", ply, "save", ckpt],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert p.returncode == 0, p.stdout
    v, f = read_ply(ply)
    scene = api.Scene(0.5)
    scene.load(ckpt)                                   # the run's energies
    assert len(v) == 4 * scene.P and len(f) == scene.P
    exp = np.rint(np.clip(scene.smooth_shade().reshape(-1, 3), 0.0, 1.0) * 255.0).astype(np.int32)
    assert np.abs(v["rgb"].astype(np.int32) - exp).max() <= 1       # device and host sums agree to rounding
    assert (v["rgb"].astype(np.int32).sum(1) > 0).mean() > 0.3      # ten batches light a good part of the box

"""CPU: the lane-level walks of the tile-binned rasteriser (radiosity_b200/csrc/tile_walk.cuh is host/device code) are
compiled with g++ and run lane by lane, tile by tile, against a brute-force statement of the raster rules — quarter-warp
walks of small-quad records cut by tile borders (partial tiles included), lone triangles, large triangles in the int32
and int64 forms, vertices on pixel centres and pixel edges (tie rules), fragments beyond the far plane.  Keys (depth24,
id) must be equal bit for bit."""
import json
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_tile_walks_equal_brute_force(tmp_path):
    exe = str(tmp_path / "tile_walk_check")
    src = os.path.join(ROOT, "tests", "cpu", "tile_walk_check.cpp")
    subprocess.run(["g++", "-O1", "-ffp-contract=off", "-std=c++17", "-Wall", "-o", exe, src], check=True)
    p = subprocess.run([exe, "30000"], stdout=subprocess.PIPE, text=True)
    res = json.loads(p.stdout)
    assert p.returncode == 0, res
    assert res["mismatched_pixels"] == 0 and res["bad_offsets"] == 0
    assert res["quads"] > 5000 and res["lone_triangles"] > 5000 and res["big_triangles"] > 5000
    assert res["multi_tile_records"] > 1000 and res["int64_walks"] > 1000

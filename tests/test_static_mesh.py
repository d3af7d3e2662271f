"""the adapter for 'static 3DS export' headers such as the reference's (dead) TestModel.h —
radiosity_b200/host/StaticMeshModel.*.  CPU: parsing, orientation, materials, subdivision."""
import os

import numpy as np
import pytest

FIXTURE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "static_mesh_fixture.h")
REF_HEADER = "/root/reference/source/TestModel.h"


def normals(v):
    q = v.reshape(-1, 4, 3)
    return np.cross(q[:, 3] - q[:, 0], q[:, 1] - q[:, 0]), q.mean(1)       # Patch::getNormal = (v4 - v1) x (v2 - v1)


def test_fixture_room(api):
    s = api.Scene(0, static_mesh=FIXTURE, scale=0.01, flip=True, emissive_material=1)
    v, _, c, r, il = s.arrays()
    assert s.P == 14                                                  # 12 room triangles + 2 lamp triangles
    q = v.reshape(-1, 4, 3)
    assert (q[:, 2] == q[:, 3]).all()                                 # triangles are degenerate quads (WaveFrontModel.cpp:98-99)
    assert np.allclose(v.reshape(-1, 3).min(0), 0.0) and np.allclose(v.reshape(-1, 3).max(0), 2.0)   # 200 units * 0.01
    n, cen = normals(v)
    assert (((np.array([1.0, 1.0, 1.0]) - cen) * n).sum(1) > 0).all()  # flip: every patch shoots into the room
    assert (r[:12] == 0).all() and (r[12:] == 100).all() and (il[12:] == 1).all()     # the lamp material emits like the built-in light
    assert (c[:12] == np.float32(0.75)).all() and (c[12:] == np.array([0.75, 0.25, 0.25], np.float32)).all()
    # without flip the winding of the export is kept: the patches face outwards
    s2 = api.Scene(0, static_mesh=FIXTURE, scale=0.01, flip=False)
    n2, cen2 = normals(s2.arrays()[0])
    assert (((np.array([1.0, 1.0, 1.0]) - cen2) * n2).sum(1) < 0).all() and s2.arrays()[3].sum() == 0


def test_fixture_subdivision(api):
    s = api.Scene(0.05, static_mesh=FIXTURE, scale=0.01, flip=True, emissive_material=1)
    v, _, c, r, il = s.arrays()
    assert s.P > 14 and np.isfinite(v).all()
    assert (r[:, 0] > 0).sum() >= 2 and set(np.unique(r)) == {0.0, 100.0}   # children inherit B unchanged (Patch.cpp:119-122)
    assert v.reshape(-1, 3).min() >= -1e-6 and v.reshape(-1, 3).max() <= 2.0 + 1e-6


def test_missing_file(api):
    with pytest.raises(api.RadError):
        api.Scene(0, static_mesh="/nonexistent/header.h")


@pytest.mark.skipif(not os.path.exists(REF_HEADER), reason="reference tree not present (GPU box)")
def test_reference_testmodel_header(api):
    """The reference's own TestModel.h read as text: four objects, 12 + 12 + 12 + 36 triangles (TestModel.h:64-75),
    material ranges of object 3 (TestModel.h: p_object_3_materials)."""
    s = api.Scene(0, static_mesh=REF_HEADER, scale=0.01, emissive_material=3)
    v, _, c, r, il = s.arrays()
    assert s.P == 72
    assert (r[:, 0] > 0).sum() == 2                                   # material 3 of object 3: 6 indices = 2 triangles
    assert len(np.unique(c, axis=0)) == 8                             # eight materials -> eight palette colours
    assert np.abs(v).max() < 2.0                                      # ~±110 export units * 0.01

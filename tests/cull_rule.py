"""numpy float32 restatement of the rasteriser's conservative culling stage (radiosity_b200/csrc/raster.cu
raster_cull_kernel + the per-emitter margin RadEmitter::ctol of camera.cuh camera_emitter), kept operation for operation
in step with the CUDA code.  Test infrastructure: it lets the CPU suite check the RULE — "a (patch, face) pair is only
dropped if the exact stage would not draw a pixel of it" — against the oracle's item buffers for many shooters, which is
how the margin bug of small shooters was found and fixed (see tests/test_cull_rule_cpu.py)."""
import numpy as np

f32 = np.float32


def rcross(a, b):
    """the reference's v_Cross: a.v_Cross(b) == b x a (Vector.h:534-537)"""
    return np.stack([b[..., 1] * a[..., 2] - b[..., 2] * a[..., 1], b[..., 2] * a[..., 0] - b[..., 0] * a[..., 2],
                     b[..., 0] * a[..., 1] - b[..., 1] * a[..., 0]], -1).astype(f32)


def vnormalize(a):
    t = np.sqrt((a * a).sum(dtype=f32)).astype(f32)
    return (a * (f32(1) / t)).astype(f32) if t != 0 else a


def shooter_frame(v, sh):
    q = v.reshape(-1, 4, 3).astype(f32)
    a, b, c, d = q[sh]
    eye = ((a + b + c + d) / f32(4)).astype(f32)
    n = rcross(b - a, d - a); u = (d - a).astype(f32); sx = rcross(u, n)
    ax = np.stack([sx / np.linalg.norm(sx), u / np.linalg.norm(u), n / np.linalg.norm(n)]).astype(f32)
    return eye, n, u, ax


def _look_basis(eye, n, u, face, ideal):
    if face == 0: target, up = u, -n
    elif face == 1: target, up = -u, n
    elif face == 2: target, up = -rcross(n, u), u
    elif face == 3: target, up = rcross(n, u), u
    else: target, up = n, u
    dirv = vnormalize(target.astype(f32) if ideal else ((target + eye).astype(f32) - eye).astype(f32))
    right = vnormalize(rcross(dirv, up))
    return right, rcross(right, dirv), dirv


def frame_deviation(v, sh):
    """camera_emitter: largest |real - ideal| over the basis vectors of the five faces, real = the reference's float32
    LookAt(eye, target + eye, up) (Camera.cpp:19-52, Transform.cpp:26-46), ideal = the same without the round trip through
    eye (its vectors are +-axes of the shooter frame the culls work in)"""
    eye, n, u, ax = shooter_frame(v, sh)
    dev = 0.0
    for face in range(5):
        real = _look_basis(eye, n, u, face, False); ideal = _look_basis(eye, n, u, face, True)
        for wr, wi in zip(real, ideal):
            d = (wr - wi).astype(f32)
            dev = max(dev, float(np.sqrt((d * d).sum(dtype=f32))))
    return dev


def ctol(v, sh):
    m = f32(2e-3) + f32(3.0) * f32(frame_deviation(v, sh))
    return f32(m * m)


def cull_faces(v, sh, tol_rel):
    """raster_cull_kernel for shooter sh over all patches: bit f set = (patch, face f) goes to the exact stage
    (0 UP, 1 DOWN, 2 LEFT, 3 RIGHT, 4 FRONT)"""
    q = v.reshape(-1, 4, 3).astype(f32)
    eye, n, u, ax = shooter_frame(v, sh)
    da = eye - q[:, 0]
    n1 = rcross(q[:, 1] - q[:, 0], q[:, 2] - q[:, 0]); n2 = rcross(q[:, 2] - q[:, 0], q[:, 3] - q[:, 0])
    d2 = (da * da).sum(-1); s1 = (n1 * da).sum(-1); s2 = (n2 * da).sum(-1)
    m1 = f32(0.01) * (n1 * n1).sum(-1) * d2; m2 = f32(0.01) * (n2 * n2).sum(-1) * d2
    back = (((s1 < 0) & (s1 * s1 > m1)) | (m1 == 0)) & (((s2 < 0) & (s2 * s2 > m2)) | (m2 == 0)) & ((m1 > 0) | (m2 > 0))
    out = np.zeros((q.shape[0], 17), np.int32)
    for vi in range(4):
        r = q[:, vi] - eye
        A = (r * ax[0]).sum(-1); B = (r * ax[1]).sum(-1); C = (r * ax[2]).sum(-1)
        tol = tol_rel * (r * r).sum(-1)
        h = [C, C - A, C + A, C - B, C + B, B - A, B + A, B - C, -B - A, -B + A, -B - C, A - B, A + B, A - C, -A - B, -A + B, -A - C]
        for k in range(17):
            out[:, k] |= ((h[k] < 0) & (h[k] * h[k] > tol)).astype(np.int32) << vi
    full = out == 15
    faces = np.zeros(q.shape[0], np.int32)
    faces |= (~(full[:, 5] | full[:, 6] | full[:, 7])) * 1
    faces |= (~(full[:, 8] | full[:, 9] | full[:, 10])) * 2
    faces |= (~(full[:, 11] | full[:, 12] | full[:, 13])) * 4
    faces |= (~(full[:, 14] | full[:, 15] | full[:, 16])) * 8
    faces |= (~(full[:, 1] | full[:, 2] | full[:, 3] | full[:, 4])) * 16
    faces[full[:, 0] | back] = 0
    return faces


def face_map(N):
    """face shown by every atlas pixel (Main.cpp:314-389): rows [0, N) LEFT | FRONT | RIGHT, rows [N, 1.5 N) UP | DOWN"""
    fm = np.zeros((N + N // 2, 2 * N), np.int32)
    fm[:N, :N // 2] = 2; fm[:N, N // 2:N + N // 2] = 4; fm[:N, N + N // 2:] = 3
    fm[N:, :N] = 0; fm[N:, N:] = 1
    return fm


def wrongly_culled(orc, v, sh, N, tol_rel, threads=4):
    """pixels of the oracle's item buffer whose patch the culling stage would have dropped on that face"""
    exp = orc.render_hemicube(v, sh, N, threads=threads)
    faces = cull_faces(v, sh, tol_rel)
    vis = exp > 0
    pid = exp[vis] - 1
    return int(((((faces[pid] >> face_map(N)[vis]) & 1) == 0)).sum())

"""Camera -> PerFrameCB (host side; an INPUT of the hot path, not part of it).

Restates the reference's camera set-up (main.c:1422-1477) and `update()` (main.c:1480-1562) with no
keyboard/mouse input: reversed-Z infinite left-handed projection, view space (x right, y up) mapped to
the right-handed z-up world by `change_of_basis`. All arithmetic is fp32 in the reference's order
except the 4x4 inverse, which uses a plain cofactor expansion (the reference spells out its own
closed form, math.h:282-320); results agree to a few ulp, and both the oracle and the GPU path are
always fed the SAME constant-buffer bytes, so this does not enter parity.
"""
from __future__ import annotations

import numpy as np

F = np.float32
PI = F(3.141592654)
TAU = F(6.283185307)
PI_OVER_TWO = F(1.570796326)


def _dot4(a, b):  # v4f32_dot math.h:137-140, serial fp32
    return F(F(F(F(a[0] * b[0]) + F(a[1] * b[1])) + F(a[2] * b[2])) + F(a[3] * b[3]))


def _mul(m0, m1):  # m4x4f32_mul_m4x4f32 math.h:188-197
    out = np.zeros((4, 4), dtype=F)
    for r in range(4):
        for c in range(4):
            out[r, c] = _dot4(m0[r], m1[:, c])
    return out


def _inverse(m):
    m64 = m.astype(np.float64)
    cof = np.zeros((4, 4), dtype=np.float64)
    for r in range(4):
        for c in range(4):
            minor = np.delete(np.delete(m64, r, axis=0), c, axis=1)
            cof[r, c] = ((-1) ** (r + c)) * np.linalg.det(minor)
    det = float(np.dot(m64[0], cof[0]))
    return (cof.T / det).astype(F)


def per_frame_cb(width: int, height: int, pos=(3.5, 1.0, 1.0), yaw_rad: float = 0.0, pitch_rad: float = 0.0,
                 fov_y_angle_deg: float = 75.0, near_plane: float = 0.01) -> np.ndarray:
    """-> float32 [3, 4, 4]: clip_from_world, view_from_clip, world_from_view (PerFrameCB main.c:169-173)."""
    fov_y_angle_rad = F(F(fov_y_angle_deg) * F(PI / F(180.0)))
    aspect_ratio = F(F(width) / F(height))
    scale_y = F(1.0 / np.tan(float(fov_y_angle_rad) / 2.0))
    scale_x = F(scale_y / aspect_ratio)
    clip_from_view = np.array([[scale_x, 0, 0, 0], [0, scale_y, 0, 0], [0, 0, 0, F(near_plane)], [0, 0, 1, 0]], dtype=F)

    yaw = F(yaw_rad)
    if yaw > PI:
        yaw = F(yaw - TAU)
    elif yaw <= -PI:
        yaw = F(yaw + TAU)
    pitch = F(max(-PI_OVER_TWO, min(PI_OVER_TWO, F(pitch_rad))))
    cp, sp = F(np.cos(float(-pitch))), F(np.sin(float(-pitch)))
    cy, sy = F(np.cos(float(-yaw))), F(np.sin(float(-yaw)))
    rotation_pitch = np.array([[1, 0, 0, 0], [0, cp, sp, 0], [0, -sp, cp, 0], [0, 0, 0, 1]], dtype=F)
    rotation_yaw = np.array([[cy, 0, -sy, 0], [0, 1, 0, 0], [sy, 0, cy, 0], [0, 0, 0, 1]], dtype=F)
    change_of_basis = np.array([[0, 0, -1, 0], [1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=F)
    world_from_view = _mul(change_of_basis, _mul(rotation_yaw, rotation_pitch))
    world_from_view[0, 3], world_from_view[1, 3], world_from_view[2, 3] = F(pos[0]), F(pos[1]), F(pos[2])
    view_from_world = _inverse(world_from_view)
    clip_from_world = _mul(clip_from_view, view_from_world)
    view_from_clip = _inverse(clip_from_view)
    return np.stack([clip_from_world, view_from_clip, world_from_view]).astype(F)

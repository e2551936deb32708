"""glm-equivalent matrix helpers (float32, column-major, right-handed, NDC z in [-1,1]).

The reference builds every matrix with glm (`lookAt`, `ortho`, `perspective`, `scale`;
/root/reference/Voxel_Cone_Tracing_Final/Voxel_Cone_Tracing.h:84-86,128-134,161-162,183 and
Camera.h:75-78,131-144).  glm is not vendored in the reference, so these are the published closed
forms.  Matrices are returned as numpy float32 arrays of shape (4, 4) in MATH layout
(`m[row, col]`); `colmajor()` flattens to the 16-float column-major order the uniform setters take
(Shader.h:414-417, transpose = GL_FALSE).
"""
from __future__ import annotations

import numpy as np

F = np.float32


def _v3(v):
    return np.asarray(v, dtype=F).reshape(3)


def normalize(v):
    v = _v3(v)
    return (v / F(np.sqrt(F(np.dot(v, v))))).astype(F)


def look_at(eye, center, up):
    eye, center, up = _v3(eye), _v3(center), _v3(up)
    f = normalize(center - eye)
    s = normalize(np.cross(f, up).astype(F))
    u = np.cross(s, f).astype(F)
    m = np.identity(4, dtype=F)
    m[0, :3] = s
    m[1, :3] = u
    m[2, :3] = -f
    m[0, 3] = -F(np.dot(s, eye))
    m[1, 3] = -F(np.dot(u, eye))
    m[2, 3] = F(np.dot(f, eye))
    return m


def ortho(l, r, b, t, n, f):
    l, r, b, t, n, f = map(F, (l, r, b, t, n, f))
    m = np.identity(4, dtype=F)
    m[0, 0] = F(2) / (r - l)
    m[1, 1] = F(2) / (t - b)
    m[2, 2] = -F(2) / (f - n)
    m[0, 3] = -(r + l) / (r - l)
    m[1, 3] = -(t + b) / (t - b)
    m[2, 3] = -(f + n) / (f - n)
    return m


def perspective(fovy_rad, aspect, n, f):
    fovy_rad, aspect, n, f = map(F, (fovy_rad, aspect, n, f))
    th = F(np.tan(fovy_rad / F(2)))
    m = np.zeros((4, 4), dtype=F)
    m[0, 0] = F(1) / (aspect * th)
    m[1, 1] = F(1) / th
    m[2, 2] = -(f + n) / (f - n)
    m[3, 2] = -F(1)
    m[2, 3] = -(F(2) * f * n) / (f - n)
    return m


def scale(s):
    m = np.identity(4, dtype=F)
    m[0, 0] = m[1, 1] = m[2, 2] = F(s)
    return m


def radians(deg):
    return F(deg) * F(np.pi / 180.0)


def colmajor(m):
    """(4,4) math-layout matrix -> 16 floats, column-major (what glUniformMatrix4fv receives)."""
    return np.ascontiguousarray(np.asarray(m, dtype=F).T).reshape(16)


def camera_front(yaw_deg, pitch_deg):
    """Camera::updateCameraVectors, Camera.h:131-144."""
    y, p = np.radians(np.float64(yaw_deg)), np.radians(np.float64(pitch_deg))
    return normalize([np.cos(y) * np.cos(p), np.sin(p), np.sin(y) * np.cos(p)])


def view_matrix(position, yaw_deg=-90.0, pitch_deg=0.0):
    """Camera::GetViewMatrix, Camera.h:75-78: lookAt(position, position + Front, Up)."""
    front = camera_front(yaw_deg, pitch_deg)
    right = normalize(np.cross(front, _v3([0, 1, 0])))
    up = normalize(np.cross(right, front))
    pos = _v3(position)
    return look_at(pos, pos + front, up)

"""Camera matrices as the reference builds them (reference src/Camera.cpp:164-174): glm defaults --
right-handed, clip z in [-1, 1], no Vulkan Y flip; ``invProjView = inverse(proj * view)``.
Matrices are returned column-major (glm memory order) as float32[16].
"""
from __future__ import annotations

import math

import numpy as np


def perspective(fovy: float, aspect: float, near: float, far: float) -> np.ndarray:
    """glm::perspectiveRH_NO."""
    t = math.tan(fovy / 2.0)
    m = np.zeros((4, 4), dtype=np.float64)          # m[col][row]
    m[0][0] = 1.0 / (aspect * t)
    m[1][1] = 1.0 / t
    m[2][2] = -(far + near) / (far - near)
    m[2][3] = -1.0
    m[3][2] = -(2.0 * far * near) / (far - near)
    return m


def look_at(eye, center, up) -> np.ndarray:
    """glm::lookAtRH."""
    eye, center, up = (np.asarray(v, dtype=np.float64) for v in (eye, center, up))
    f = center - eye; f /= np.linalg.norm(f)
    s = np.cross(f, up); s /= np.linalg.norm(s)
    u = np.cross(s, f)
    m = np.eye(4, dtype=np.float64)                 # m[col][row]
    m[0][0], m[1][0], m[2][0] = s
    m[0][1], m[1][1], m[2][1] = u
    m[0][2], m[1][2], m[2][2] = -f
    m[3][0], m[3][1], m[3][2] = -np.dot(s, eye), -np.dot(u, eye), np.dot(f, eye)
    return m


class Camera:
    """pos / viewDir / up / aspect / fov / near / far, as en::Camera (reference src/main.cu:180-187)."""

    def __init__(self, pos=(64.0, 0.0, 0.0), view_dir=(-1.0, 0.0, 0.0), up=(0.0, 1.0, 0.0), aspect=1920.0 / 1080.0,
                 fov=math.radians(60.0), near=0.1, far=100.0):
        self.pos = tuple(float(v) for v in pos)
        self.view_dir = tuple(float(v) for v in view_dir)
        self.up = tuple(float(v) for v in up)
        self.aspect, self.fov, self.near, self.far = float(aspect), float(fov), float(near), float(far)

    def matrices(self):
        proj = perspective(self.fov, self.aspect, self.near, self.far)
        view = look_at(self.pos, np.add(self.pos, self.view_dir), self.up)
        # column-major storage m[col][row]  <->  mathematical matrix M = m.T
        pv = (proj.T @ view.T)
        inv = np.linalg.inv(pv)
        to_cm = lambda M: np.ascontiguousarray(M.T, dtype=np.float32).reshape(16)
        return to_cm(pv), to_cm(inv)

    @property
    def inv_proj_view(self) -> np.ndarray:
        return self.matrices()[1]

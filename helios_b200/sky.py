"""Host side of the Hosek-Wilkie sky: the double-precision coefficient fit that the
reference does on the CPU every frame (src/engine/gfx/hosek_wilkie_sky_model.cpp:41-67 evaluate_spline /
evaluate, :658-686 HosekWilkieSkyModel::update).  The result is the 10 x vec4 uniform block
(A,B,C,D,E,F,G,H,I,Z) that hl_sky_update() bakes into the 512^2 x 6 cube map on the GPU.

Dataset: helios_b200/data/hosek_rgb_v1_4a.f64 (Hosek & Wilkie RGB coefficients v1.4a, extracted by
tools/extract_hosek_dataset.py): 3 x 1080 doubles (9 coefficients x 6 control points x 10 turbidities x
2 albedos) followed by 3 x 120 doubles (radiance).
"""
from __future__ import annotations

import math
from functools import lru_cache
from pathlib import Path

import numpy as np

_DATA = Path(__file__).resolve().parent / "data" / "hosek_rgb_v1_4a.f64"


@lru_cache(maxsize=1)
def dataset():
    d = np.fromfile(_DATA, dtype="<f8")
    if d.size != 3600:
        raise RuntimeError(f"{_DATA}: expected 3600 doubles, found {d.size}")
    return d[:3240].reshape(3, 1080), d[3240:].reshape(3, 120)


def _spline(s, stride, v):
    """quintic Bezier over 6 control points spaced `stride` apart (hosek_wilkie_sky_model.cpp:41-44)"""
    c = [s[k * stride] for k in range(6)]
    w = [1, 5, 10, 10, 5, 1]
    return sum(w[k] * (1 - v) ** (5 - k) * v**k * c[k] for k in range(6))


def _evaluate(ds, stride, turbidity, albedo, sun_theta):
    # float32 intermediates where the reference uses float (hosek_wilkie_sky_model.cpp:48-66)
    e = np.float32(max(np.float32(0.0), np.float32(1.0 - float(sun_theta) / (math.pi / 2.0))))
    k = float(np.power(e, np.float32(1.0 / 3.0), dtype=np.float32))
    t0 = min(max(int(turbidity), 1), 10)
    t1 = min(t0 + 1, 10)
    tk = float(np.float32(min(max(np.float32(turbidity) - np.float32(t0), 0.0), 1.0)))
    a0, a1 = ds, ds[stride * 6 * 10 :]
    a0t0 = _spline(a0[stride * 6 * (t0 - 1) :], stride, k)
    a1t0 = _spline(a1[stride * 6 * (t0 - 1) :], stride, k)
    a0t1 = _spline(a0[stride * 6 * (t1 - 1) :], stride, k)
    a1t1 = _spline(a1[stride * 6 * (t1 - 1) :], stride, k)
    al = float(np.float32(albedo))
    return a0t0 * (1 - al) * (1 - tk) + a1t0 * al * (1 - tk) + a0t1 * (1 - al) * tk + a1t1 * al * tk


def _hosek(cos_theta, gamma, cos_gamma, cf):
    A, B, C_, D, E, F, G, H, I = (cf[k].astype(np.float32) for k in range(9))
    f = np.float32
    chi = (f(1) + f(cos_gamma) * f(cos_gamma)) / np.power(f(1) + H * H - f(2) * f(cos_gamma) * H, f(1.5))
    return (f(1) + A * np.exp(B / (f(cos_theta) + f(0.01)))) * (
        C_ + D * np.exp(E * f(gamma)) + F * (f(cos_gamma) * f(cos_gamma)) + G * chi + I * f(math.sqrt(max(0.0, cos_theta)))
    )


def sky_coefficients(sun_direction, turbidity=4.0, albedo=0.1, normalized_sun_y=1.15) -> np.ndarray:
    """HosekWilkieSkyModel::update (:658-686).  sun_direction = -directional_light[0].forward()."""
    rgb, rad = dataset()
    sun_theta = np.float32(math.acos(min(max(float(np.float32(sun_direction[1])), 0.0), 1.0)))
    cf = np.zeros((10, 3), np.float32)
    for i in range(3):
        for k in range(7):
            cf[k, i] = _evaluate(rgb[i][k:], 9, turbidity, albedo, sun_theta)
        cf[7, i] = _evaluate(rgb[i][8:], 9, turbidity, albedo, sun_theta)  # H and I are swapped in the dataset (:674-676)
        cf[8, i] = _evaluate(rgb[i][7:], 9, turbidity, albedo, sun_theta)
        cf[9, i] = _evaluate(rad[i], 1, turbidity, albedo, sun_theta)
    if normalized_sun_y:
        S = _hosek(np.float32(math.cos(sun_theta)), 0.0, 1.0, cf) * cf[9]
        lum = np.float32(S[0] * np.float32(0.2126) + S[1] * np.float32(0.7152) + S[2] * np.float32(0.0722))
        cf[9] = cf[9] / lum
        cf[9] = cf[9] * np.float32(normalized_sun_y)
    out = np.zeros((10, 4), np.float32)
    out[:, :3] = cf
    return out.reshape(40)

"""NumPy float64 backend of nphusl: the colour chain written array-at-a-time.

This is the backend ``enable_numpy()`` / ``numpy_enabled()`` bind and the one
``*_enabled(back_to_std=True)`` returns to -- the role the reference's
``nphusl/nphusl.py:104-392`` plays.  It is an *explicitly selected* CPU backend:
``enable_best_optimized()`` never picks it for the three public conversions, and
nothing in ``nphusl/_cuda.py`` calls into it (the CUDA backend has no CPU
fallback).  It does not import ``oracle/``.

The step functions keep the reference's names and call signatures because its
tests and its table generators call them by name (``nphusl.nphusl._to_light`` in
``LUT/light.py:21``, ``nphusl.nphusl._max_lh_chroma`` in ``LUT/chroma.py:22``,
``tests/test.py:153-457``):

    sRGB -> linear (``_to_linear``) -> XYZ (``_dot_product`` with ``M_INV``)
         -> CIELUV (``_xyz_to_luv``, ``_to_light``) -> LCh (``_luv_to_lch``)
         -> HUSL (``_lch_to_husl``: chroma relative to ``_max_lh_chroma``)

and back (``_husl_to_lch``, ``_lch_to_luv``, ``_luv_to_xyz``, ``_from_light``,
``_xyz_to_rgb``, ``_from_linear``).  Everything is float64; shapes are kept as
given (``(..., 3)`` in, ``(..., 3)`` out).

Formulation notes (where this differs from a transcription of the reference):
the six gamut boundary lines are evaluated as one ``(6, N)`` array instead of a
generator of twelve temporaries; branches are ``np.where`` selections instead of
boolean-mask scatters; ``X`` and ``Z`` of the inverse use the simplified
``9Y u'/(4v')`` and ``Y (12 - 3u' - 20v')/(4v')``.  Values agree with the
reference NumPy path to ~1e-12 (``tests/test_numpy_backend.py`` pins 1e-9
against vectors the reference itself produced).
"""
import math

import numpy as np

from . import constants
from .transform import float_input, rgb_float_input

__all__ = ["_rgb_to_husl", "_husl_to_rgb", "_rgb_to_hue"]

_M = np.array(constants.M, dtype=np.float64)            # linear RGB <- XYZ
_M_INV = np.array(constants.M_INV, dtype=np.float64)    # XYZ <- linear RGB
L_MAX, L_MIN = constants.L_MAX, constants.L_MIN

# Gamut boundary in the (u, v) chroma plane at lightness L: for each linear channel c (row m of M)
# and each limit t in {0, 1} the locus "channel c == t" is a straight line v = slope * u + intercept
# (reference nphusl.py:179-213).  Per-row scalars:
_LINE_TOP1 = 284517.0 * _M[:, 0] - 94839.0 * _M[:, 2]
_LINE_TOP2 = 838422.0 * _M[:, 2] + 769860.0 * _M[:, 1] + 731718.0 * _M[:, 0]
_LINE_BOTTOM = 632260.0 * _M[:, 2] - 126452.0 * _M[:, 1]
_TOP2_L = 769860.0
_BOTTOM_T = 126452.0


def _channel(data, last_dim_idx):
    return data[..., last_dim_idx]


def _f64(arr):
    return np.asarray(arr, dtype=np.float64)


# ---------------------------------------------------------------------------
# RGB -> HUSL
# ---------------------------------------------------------------------------

@rgb_float_input
def _to_linear(rgb_nd):
    """sRGB companded values in [0, 1] -> linear light (reference nphusl.py:278-286)."""
    rgb_nd = _f64(rgb_nd)
    with np.errstate(invalid="ignore"):
        curved = ((rgb_nd + 0.055) / 1.055) ** 2.4
    return np.where(rgb_nd > 0.04045, curved, rgb_nd / 12.92)


def _dot_product(scalars, rgb_nd):
    """Apply the 3 x 3 matrix ``scalars`` to the last axis (reference nphusl.py:289-296)."""
    return _f64(rgb_nd) @ _f64(scalars).T


def _rgb_to_xyz(rgb_nd):
    return _dot_product(_M_INV, _to_linear(rgb_nd))


def _to_light(y_nd):
    """CIE lightness L* of relative luminance Y (reference nphusl.py:268-275)."""
    y = _f64(y_nd) / constants.REF_Y
    with np.errstate(invalid="ignore"):
        cube_root = np.power(y, 1.0 / 3.0) * 116 - 16
    return np.where(y > constants.EPSILON, cube_root, y * constants.KAPPA)


def _xyz_to_luv(xyz_nd):
    """XYZ -> CIELUV (reference nphusl.py:246-265); black maps to (0, 0, 0)."""
    xyz_nd = _f64(xyz_nd)
    x, y, z = (_channel(xyz_nd, n) for n in range(3))
    denom = x + 15.0 * y + 3.0 * z
    with np.errstate(divide="ignore", invalid="ignore"):
        u_prime = np.where(denom != 0, 4.0 * x / denom, 0.0)
        v_prime = np.where(denom != 0, 9.0 * y / denom, 0.0)
    light = _to_light(y)
    lit = light != 0
    luv = np.empty(xyz_nd.shape, dtype=np.float64)
    luv[..., 0] = light
    luv[..., 1] = np.where(lit, 13.0 * light * (u_prime - constants.REF_U), 0.0)
    luv[..., 2] = np.where(lit, 13.0 * light * (v_prime - constants.REF_V), 0.0)
    return np.nan_to_num(luv)


def _luv_to_lch(luv_nd):
    """CIELUV -> (L, chroma, hue in degrees [0, 360)) (reference nphusl.py:223-243)."""
    luv_nd = _f64(luv_nd)
    u = _channel(luv_nd, 1) + 0.0            # -0.0 + 0.0 == +0.0: atan2 must not see a negative zero
    v = _channel(luv_nd, 2) + 0.0
    hue = np.degrees(np.arctan2(v, u))
    lch = np.empty(luv_nd.shape, dtype=np.float64)
    lch[..., 0] = _channel(luv_nd, 0)
    lch[..., 1] = np.sqrt(u * u + v * v)
    lch[..., 2] = np.where(hue < 0.0, hue + 360.0, hue)
    return lch


def _rgb_to_lch(rgb):
    return _luv_to_lch(_xyz_to_luv(_rgb_to_xyz(rgb)))


def _line_table(l_nd):
    """-> (slopes, intercepts), each of shape ``(6,) + l_nd.shape``, ordered like the reference's
    ``_bounds`` generator: channel-major, t = 0 before t = 1."""
    light = _f64(l_nd)
    sub1 = (light + 16.0) ** 3 / 1560896.0
    sub2 = np.where(sub1 < constants.EPSILON, light / constants.KAPPA, sub1)
    t = np.array([0.0, 1.0]).reshape((1, 2) + (1,) * light.ndim)
    shape3 = (3, 1) + (1,) * light.ndim
    top1 = _LINE_TOP1.reshape(shape3) * sub2
    top2 = _LINE_TOP2.reshape(shape3) * (light * sub2) - t * (_TOP2_L * light)
    bottom = _LINE_BOTTOM.reshape(shape3) * sub2 + t * _BOTTOM_T
    with np.errstate(divide="ignore", invalid="ignore"):
        slopes = np.broadcast_to(top1, top2.shape) / bottom
        intercepts = top2 / bottom
    return slopes.reshape((6,) + light.shape), intercepts.reshape((6,) + light.shape)


def _bounds(l_nd):
    """Iterate the six ``(slope, intercept)`` boundary lines of the gamut at each lightness
    (reference nphusl.py:188-213)."""
    slopes, intercepts = _line_table(l_nd)
    for k in range(6):
        yield slopes[k], intercepts[k]


def _ray_length(theta, line):
    """Distance from the origin along direction ``theta`` (radians) to ``line`` =
    ``(slope, intercept)``; negative when the ray points away from it (reference nphusl.py:216-220)."""
    slope, intercept = line
    with np.errstate(divide="ignore", invalid="ignore"):
        return intercept / (np.sin(theta) - slope * np.cos(theta))


def _max_lh_chroma(lch):
    """Largest in-gamut chroma for each (L, ., H) row of ``lch`` (reference nphusl.py:165-176)."""
    lch = _f64(lch)
    light, hue = _channel(lch, 0), _channel(lch, 2)
    theta = np.radians(hue)
    lengths = _ray_length(theta, _line_table(light))
    lengths = np.where(np.isnan(lengths) | (lengths < 0), np.inf, lengths)
    return lengths.min(axis=0)


def _lch_to_husl(lch_nd):
    """LCh -> HUSL: S = 100 * C / max chroma; the lightness extremes are clamped
    (reference nphusl.py:136-159)."""
    lch_nd = _f64(lch_nd)
    light, chroma, hue = (_channel(lch_nd, n) for n in range(3))
    too_light, too_dark = light > L_MAX, light < L_MIN
    with np.errstate(divide="ignore", invalid="ignore"):
        sat = chroma / _max_lh_chroma(lch_nd) * 100.0
    hsl = np.empty(lch_nd.shape, dtype=np.float64)
    hsl[..., 0] = hue
    hsl[..., 1] = np.where(too_light | too_dark, 0.0, sat)
    hsl[..., 2] = np.where(too_light, 100.0, np.where(too_dark, 0.0, light))
    return hsl


@rgb_float_input
def _rgb_to_husl(rgb_nd):
    """RGB (uint8, or float in [0, 1]) -> float64 HUSL, same shape (reference nphusl.py:110-115)."""
    return _lch_to_husl(_rgb_to_lch(rgb_nd))


@rgb_float_input
def _rgb_to_hue(rgb):
    """RGB -> float64 HUSL hue, shape ``rgb.shape[:-1]`` (reference nphusl.py:122-128)."""
    return np.ascontiguousarray(_channel(_rgb_to_lch(rgb), 2))


# ---------------------------------------------------------------------------
# HUSL -> RGB
# ---------------------------------------------------------------------------

def _husl_to_lch(husl_nd):
    """HUSL -> LCh (reference nphusl.py:338-358)."""
    husl_nd = _f64(husl_nd)
    hue, sat, light = (_channel(husl_nd, n) for n in range(3))
    probe = np.empty(husl_nd.shape, dtype=np.float64)
    probe[..., 0] = light
    probe[..., 1] = 0.0
    probe[..., 2] = hue
    too_light, too_dark = light > L_MAX, light < L_MIN
    with np.errstate(invalid="ignore"):           # inf * 0 at the extremes, overwritten below
        chroma = _max_lh_chroma(probe) / 100.0 * sat
    lch = probe
    lch[..., 0] = np.where(too_light, 100.0, np.where(too_dark, 0.0, light))
    lch[..., 1] = np.where(too_light | too_dark, 0.0, chroma)
    return lch


def _lch_to_luv(lch_nd):
    lch_nd = _f64(lch_nd)
    theta = np.radians(_channel(lch_nd, 2))
    luv = np.empty(lch_nd.shape, dtype=np.float64)
    luv[..., 0] = _channel(lch_nd, 0)
    luv[..., 1] = np.cos(theta) * _channel(lch_nd, 1)
    luv[..., 2] = np.sin(theta) * _channel(lch_nd, 1)
    return luv


def _from_light(l_nd):
    """Inverse of ``_to_light`` (reference nphusl.py:385-392)."""
    light = _f64(l_nd)
    return constants.REF_Y * np.where(light > 8, ((light + 16) / 116) ** 3.0, light / constants.KAPPA)


def _luv_to_xyz(luv_nd):
    """CIELUV -> XYZ (reference nphusl.py:361-382); L == 0 maps to black."""
    luv_nd = _f64(luv_nd)
    light, u, v = (_channel(luv_nd, n) for n in range(3))
    y = _from_light(light)
    with np.errstate(divide="ignore", invalid="ignore"):
        u_prime = u / (13.0 * light) + constants.REF_U
        v_prime = v / (13.0 * light) + constants.REF_V
        u_prime = np.where(np.isinf(u_prime), 0.0, u_prime)
        v_prime = np.where(np.isinf(v_prime), 0.0, v_prime)
        x = 9.0 * y * u_prime / (4.0 * v_prime)
        z = y * (12.0 - 3.0 * u_prime - 20.0 * v_prime) / (4.0 * v_prime)
    xyz = np.stack([x, y, z], axis=-1)
    xyz[light == 0] = 0.0
    return np.nan_to_num(xyz)


def _from_linear(xyz_nd):
    """linear light -> sRGB companded (reference nphusl.py:330-335)."""
    lin = _f64(xyz_nd)
    with np.errstate(invalid="ignore"):
        curved = 1.055 * lin ** (1 / 2.4) - 0.055
    return np.where(lin <= 0.0031308, 12.92 * lin, curved)


def _xyz_to_rgb(xyz_nd):
    return _from_linear(_dot_product(_M, xyz_nd))


def _lch_to_rgb(lch_nd):
    return _xyz_to_rgb(_luv_to_xyz(_lch_to_luv(lch_nd)))


@float_input
def _husl_to_rgb(husl_nd):
    """HUSL -> float64 RGB in [0, 1], same shape (reference nphusl.py:305-307); the public
    ``to_rgb`` rounds it to uint8 (``transform.rgb_int_output``)."""
    return _lch_to_rgb(_husl_to_lch(husl_nd))


# the names the reference registers through ``@optimized`` (nphusl.py:84-101)
REGISTERED = ("_rgb_to_husl", "_rgb_to_hue", "_lch_to_husl", "_max_lh_chroma", "_bounds", "_ray_length",
              "_luv_to_lch", "_xyz_to_luv", "_to_light", "_to_linear", "_dot_product", "_husl_to_rgb")
# helper steps that live beside them in the reference's nphusl.nphusl namespace
HELPERS = ("_rgb_to_lch", "_rgb_to_xyz", "_lch_to_rgb", "_xyz_to_rgb", "_lch_to_luv", "_from_linear",
           "_husl_to_lch", "_luv_to_xyz", "_from_light", "_channel")
_2pi = math.pi * 2

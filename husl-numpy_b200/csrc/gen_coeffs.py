#!/usr/bin/env python
"""Derive, in float64, every coefficient the CUDA kernels use and print
``husl_coeffs.h`` (committed; rerun only when the formulation changes).

Inputs are the reference's constants (nphusl/constants.py:3-21): M, M_INV,
REF_U, REF_V, KAPPA, EPSILON.  The kernels do not walk the reference's chain
RGB -> XYZ -> Luv -> LCh -> six boundary lines; they use an algebraically
identical form that needs no trigonometry in the forward direction and has no
catastrophic cancellation in float32 (derivation in DESIGN.md section 3):

  basis      dr = r-g, db = b-g, g    (linear-light differences; the table
                                        holds hi+lo float pairs so dr, db and
                                        1-g are exact to ~1e-14)
  Y          = sY*g + y_r*dr + y_b*db
  D          = X+15Y+3Z = sD*g + d_r*dr + d_b*db
  Nu, Nv     = 4X - REF_U*D, 9Y - REF_V*D   (U = 13 L Nu/D, V = 13 L Nv/D)
  H          = atan2(Nv, Nu)
  S/100      = max( (D - sD*min(r,g,b)) / D ,
                    1 - sD*Y*(1-max(r,g,b)) / (D*(1-Y)) )
  L          = 116*cbrt(Y) - 16  or  KAPPA*Y

The S formula is the six-line ray intersection of nphusl.py:165-220 /
_cython_opt.pyx:271-309 in closed form: lines with t=0 are "linear channel
i == 0", lines with t=1 are "linear channel i == 1"; `verify()` checks the
identity against a literal evaluation of the six lines.
"""
import sys

import numpy as np

M = np.array([[3.240969941904521, -1.537383177570093, -0.498610760293],
              [-0.96924363628087, 1.87596750150772, 0.041555057407175],
              [0.055630079696993, -0.20397695888897, 1.056971514242878]])
M_INV = np.array([[0.41239079926595, 0.35758433938387, 0.18048078840183],
                  [0.21263900587151, 0.71516867876775, 0.072192315360733],
                  [0.019330818715591, 0.11919477979462, 0.95053215224966]])
REF_U = 0.19783000664283
REF_V = 0.46831999493879
KAPPA = 903.2962962
EPSILON = 0.0088564516
L_MAX, L_MIN = 99.99, 0.01
WHITE_HUE = 19.916405993809086


def fit_atan(nterms, npts=6000, iters=80):
    """Near-minimax odd polynomial atan(q) ~ q*P(q^2) on [0,1] (Lawson-weighted
    least squares on Chebyshev nodes); returns coefficients and max abs error."""
    q = np.cos(np.pi * (np.arange(npts) + 0.5) / npts) * 0.5 + 0.5
    s, y = q * q, np.arctan(q) / q
    V = np.vander(s, nterms, increasing=True)
    w = np.ones_like(q)
    for _ in range(iters):
        coef = np.linalg.lstsq(V * (q * w)[:, None], y * q * w, rcond=None)[0]
        err = np.abs(q * (V @ coef - y))
        w *= (err / err.max() + 1e-3) ** 0.5
        w /= w.max()
    return coef, err.max()


ATAN_TERMS = 7
ATAN_SHRINK = 6e-7


def atan_first_quadrant_f32(c32, q):
    """float32 emulation of hue_deg's `fmaf(P(q*q), q, 45)` (Horner with FFMA), q in [-1, 1]"""
    q = q.astype(np.float32)
    s = (q.astype(np.float64) ** 2).astype(np.float32)
    p = np.full_like(s, np.float32(c32[-1]))
    for c in c32[-2::-1]:
        p = (p.astype(np.float64) * s.astype(np.float64) + np.float64(c)).astype(np.float32)
    return (p.astype(np.float64) * q.astype(np.float64) + 45.0).astype(np.float32)


def atan_coeffs():
    """Degrees-valued coefficients of hue_deg, shrunk by ATAN_SHRINK so that the
    first-quadrant angle 45 + q*P(q^2) stays inside [0, 90] for EVERY float q in
    [-1 - 4 ulp, 1 + 4 ulp] (checked exhaustively near the ends, where the minimax error and the
    Horner rounding could otherwise push it to -1.5e-5 / 90.000015 and the hue out
    of [0, 360]).  Returns (float32 coefficients, max abs error in degrees)."""
    coef, _ = fit_atan(ATAN_TERMS)
    c32 = (np.degrees(coef) * (1.0 - ATAN_SHRINK)).astype(np.float32)
    # all floats in [0.5, 1 + 4 ulp]: q = (y-x)*rcp(y+x) can overshoot 1 by the MUFU.RCP error (1 ulp)
    ends = np.arange(0x3F000000, 0x3F800005, dtype=np.uint32).view(np.float32)
    lo, hi = atan_first_quadrant_f32(c32, -ends), atan_first_quadrant_f32(c32, ends)
    assert lo.min() >= 0.0 and hi.max() <= 90.0, (lo.min(), hi.max())
    q = np.linspace(-1, 1, 2000001).astype(np.float32)
    err = atan_first_quadrant_f32(c32, q).astype(np.float64) - (45.0 + np.degrees(np.arctan(q.astype(np.float64))))
    return c32, np.abs(err).max()


def derive():
    c = {}
    Yf = M_INV[1]
    Df = M_INV[0] + 15 * M_INV[1] + 3 * M_INV[2]
    Nu = 4 * M_INV[0] - REF_U * Df
    Nv = 9 * M_INV[1] - REF_V * Df
    c["Y_DR"], c["Y_DB"], c["Y_S"] = Yf[0], Yf[2], Yf.sum()
    c["ONE_M_SY"] = 1.0 - Yf.sum()
    c["D_DR"], c["D_DB"], c["D_S"] = Df[0], Df[2], Df.sum()
    c["U_DR"], c["U_DB"], c["U_S"] = Nu[0], Nu[2], Nu.sum()
    c["V_DR"], c["V_DB"], c["V_S"] = Nv[0], Nv[2], Nv.sum()
    # exact greys: the float64 reference evaluates atan2 of the residue of its white point,
    # (U_S, V_S)*g ~ 1e-14, plus rounding noise, and lands at 17.6 .. 21.1 degrees (19.916 for
    # white).  The kernels add a vanishing g-term pointing exactly at WHITE_HUE instead: greys
    # get the hue of white, black (g = 0) gets atan2(0, 0) = 0, and any pixel with a real
    # colour difference is unaffected (1e-13 is far below float32 resolution of Nu, Nv there).
    c["GREY_U"] = 1e-13 * np.cos(np.radians(WHITE_HUE))
    c["GREY_V"] = 1e-13 * np.sin(np.radians(WHITE_HUE))
    c["Y_HI"] = ((L_MAX + 16) / 116) ** 3        # L > 99.99  <=>  Y > Y_HI
    c["Y_LO"] = L_MIN / KAPPA                    # L < 0.01   <=>  Y < Y_LO
    # inverse: lin_i = Y * A_i(u', v') / (4 v'),  A_i = AU_i u' + AV_i v' + A0_i
    Au = 9 * M[:, 0] - 3 * M[:, 2]
    Av = 4 * M[:, 1] - 20 * M[:, 2]
    A0 = 12 * M[:, 2]
    An = A0 + Au * REF_U + Av * REF_V            # == 4*REF_V / sY for every row
    for i in range(3):
        c["AU%d" % i], c["AV%d" % i], c["AN%d" % i] = Au[i], Av[i], An[i]
    c["AN"] = An.mean()
    c["REF_V4"] = 4 * REF_V
    assert np.float32(c["AN"]) == np.float32(c["REF_V4"])     # the inverse uses AN*den for both
    return c


def six_line_max_chroma(L, H):
    """Literal transcription of the reference formula (float64) for verify()."""
    sub1 = (L + 16.0) ** 3 / 1560896.0
    sub2 = sub1 if sub1 > EPSILON else L / KAPPA
    th = H / 360.0 * 2 * np.pi
    best = np.inf
    for i in range(3):
        t1 = 284517.0 * M[i][0] - 94839.0 * M[i][2]
        t2 = 838422.0 * M[i][2] + 769860.0 * M[i][1] + 731718.0 * M[i][0]
        bt = 632260.0 * M[i][2] - 126452.0 * M[i][1]
        for t in (0, 1):
            top2 = L * sub2 * t2 - (769860.0 * L if t else 0.0)
            bottom = sub2 * bt + (126452.0 if t else 0.0)
            ln = (top2 / bottom) / (np.sin(th) - (sub2 * t1 / bottom) * np.cos(th))
            if ln >= 0:
                best = min(best, ln)
    return best


def verify(c, n=2000, seed=0):
    """closed-form S and the inverse ray length against the six-line formula"""
    rng = np.random.default_rng(seed)
    worst_s = worst_t = 0.0
    for _ in range(n):
        lin = rng.uniform(0, 1, 3)
        r, g, b = lin
        dr, db = r - g, b - g
        Y = c["Y_S"] * g + c["Y_DR"] * dr + c["Y_DB"] * db
        D = c["D_S"] * g + c["D_DR"] * dr + c["D_DB"] * db
        Nu = c["U_DR"] * dr + c["U_DB"] * db + c["U_S"] * g
        Nv = c["V_DR"] * dr + c["V_DB"] * db + c["V_S"] * g
        L = 116 * Y ** (1 / 3) - 16 if Y > EPSILON else KAPPA * Y
        U, V = 13 * L * Nu / D, 13 * L * Nv / D
        H = np.degrees(np.arctan2(V, U)) % 360.0
        S_ref = 100 * np.hypot(U, V) / six_line_max_chroma(L, H)
        S = 100 * max((D - c["D_S"] * lin.min()) / D,
                      1 - c["D_S"] * Y * (1 - lin.max()) / (D * (1 - Y)))
        worst_s = max(worst_s, abs(S - S_ref))
        # inverse: tau_max in (u', v') units vs max chroma / (13 L)
        th = np.radians(H)
        gi = np.array([c["AU%d" % i] * np.cos(th) + c["AV%d" % i] * np.sin(th) for i in range(3)])
        inv_tau = max(-gi.min() / c["AN"], (Y * gi.max() - 4 * np.sin(th)) / (c["REF_V4"] * (1 - Y)))
        # the kernel works with g_i/4 and den/4: 1/(4 tau) = max(-(gi/4).min()/AN, (Y*(gi/4).max() - sin)/(REF_V4 (1-Y)))
        inv_tau4 = max(-(gi / 4).min() / c["AN"], (Y * (gi / 4).max() - np.sin(th)) / (c["REF_V4"] * (1 - Y)))
        assert abs(inv_tau4 * 4 - inv_tau) <= 1e-12 * abs(inv_tau)
        worst_t = max(worst_t, abs(1 / inv_tau - six_line_max_chroma(L, H) / (13 * L)) * 13 * L)
    return worst_s, worst_t


POWG_TAU = 0.25       # |t| below which (1+t)^2.4 - 1 is taken from the polynomial
POWG_DEG = 4


def powg_coeffs():
    """((1+t)^2.4 - 1) / t ~ polynomial of degree POWG_DEG on [-POWG_TAU, POWG_TAU] (Chebyshev-node fit, i.e.
    near-minimax): the gain of the sRGB power segment between two NEARBY channel values, used by the float32
    kernels to form differences of linear light without cancellation (husl_math.cuh: diff_from_f32).
    Returns the coefficients and the maximum relative error of the fit."""
    from numpy.polynomial import chebyshev as C, polynomial as P
    k = np.arange(4000)
    x = np.cos(np.pi * (k + 0.5) / 4000) * POWG_TAU
    q = ((1 + x) ** 2.4 - 1) / x
    coef = C.Chebyshev.fit(x, q, POWG_DEG, domain=[-POWG_TAU, POWG_TAU]).convert(kind=P.Polynomial).coef
    coef = np.float32(coef).astype(np.float64)
    xx = np.linspace(-POWG_TAU, POWG_TAU, 400001)
    xx = xx[np.abs(xx) > 1e-7]
    err = np.max(np.abs(P.polyval(xx, coef) / (((1 + xx) ** 2.4 - 1) / xx) - 1))
    return coef, err


def srgb_split():
    """The float32 value at which a float32 channel switches from c/12.92 to the power segment exactly when the
    float64 reference does (c > 0.04045 in double <=> c > SRGB_T in float, SRGB_T the largest float <= 0.04045),
    the power segment's value there and the (tiny) jump between the two segments at that point."""
    t = np.float32(0.04045)
    if float(t) > 0.04045:
        t = np.nextafter(t, np.float32(0))
    assert float(t) <= 0.04045 < float(np.nextafter(t, np.float32(1)))
    pt = ((float(t) + 0.055) / 1.055) ** 2.4
    return float(t), pt, pt - float(t) / 12.92


def emit(c, out):
    coef, err = atan_coeffs()
    p = lambda *a: print(*a, file=out)  # noqa: E731
    p("// husl_coeffs.h -- GENERATED by csrc/gen_coeffs.py, do not edit.")
    p("// float64-derived coefficients of the cancellation-free HUSL formulation")
    p("// (see gen_coeffs.py docstring and DESIGN.md section 3).")
    p("#pragma once")
    p("namespace husl {")
    for k, v in c.items():
        p("constexpr double %s_D = %.17e;" % (k, v))
        p("constexpr float  %s = %.9ef;" % (k, v))
    p("constexpr float KAPPA = %.9ef;" % KAPPA)
    p("constexpr float INV_KAPPA = %.9ef;" % (1 / KAPPA))
    p("constexpr float EPSILON = %.9ef;" % EPSILON)
    p("constexpr float WHITE_HUE = %.9ef;" % WHITE_HUE)
    p("constexpr double WHITE_HUE_D = %.17e;" % WHITE_HUE)
    p("// atan(q) ~ q*P(q^2) on [-1,1], result in DEGREES; 45 + q*P(q^2) evaluated in float32 stays in")
    p("// [0, 90] for every float q and is within %.3e deg of 45 + atan(q)" % err)
    p("constexpr int ATAN_TERMS = %d;" % ATAN_TERMS)
    for i, v in enumerate(coef):
        p("constexpr float ATAN_C%d = %.9ef;" % (i, v))
    pg, perr = powg_coeffs()
    p("// ((1+t)^2.4 - 1)/t ~ polynomial on |t| <= %.4g, max relative error %.2e (float32 channels, diff_from_f32)" % (POWG_TAU, perr))
    p("constexpr float POWG_TAU = %.9ef;" % POWG_TAU)
    p("constexpr int POWG_DEG = %d;" % POWG_DEG)
    for i, v in enumerate(pg):
        p("constexpr float POWG_C%d = %.9ef;" % (i, v))
    t, pt, j = srgb_split()
    p("// sRGB segment switch for float32 channels: largest float <= 0.04045, the power segment there, the jump")
    p("constexpr float SRGB_T = %.9ef;" % t)
    p("constexpr float SRGB_PT = %.9ef;" % pt)
    p("constexpr float SRGB_J = %.9ef;" % j)
    p("}  // namespace husl")


if __name__ == "__main__":
    c = derive()
    ws, wt = verify(c)
    print("verify: closed-form S vs six-line formula, max abs diff %.3e; "
          "inverse ray length max abs diff %.3e" % (ws, wt), file=sys.stderr)
    assert ws < 1e-9 and wt < 1e-9
    emit(c, sys.stdout)

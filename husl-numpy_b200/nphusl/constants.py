"""Colour-science constants of the HUSL (HSLuv) transform.

Same names and values as the reference's ``nphusl/constants.py:3-21`` (which in
turn quotes husl.py).  They are data, not code: the sRGB<->XYZ matrices, the
D65 white point in u'v', and the CIE L* break points.  The CUDA backend keeps
its own copy in ``csrc/husl_consts.h`` (derived coefficients are computed in
float64 at build time by ``csrc/gen_coeffs.py``); ``tests/test_consts.py``
checks the two agree.
"""

# linear sRGB <- XYZ (rows: R, G, B)
M = [
    [3.240969941904521, -1.537383177570093, -0.498610760293],
    [-0.96924363628087, 1.87596750150772, 0.041555057407175],
    [0.055630079696993, -0.20397695888897, 1.056971514242878],
]

# XYZ <- linear sRGB (rows: X, Y, Z)
M_INV = [
    [0.41239079926595, 0.35758433938387, 0.18048078840183],
    [0.21263900587151, 0.71516867876775, 0.072192315360733],
    [0.019330818715591, 0.11919477979462, 0.95053215224966],
]

REF_X = 0.95045592705167
REF_Y = 1.0
REF_Z = 1.089057750759878
REF_U = 0.19783000664283
REF_V = 0.46831999493879
KAPPA = 903.2962962
EPSILON = 0.0088564516

# lightness clamps of the HUSL <-> LCh step (reference nphusl/nphusl.py:106-107)
L_MAX = 99.99
L_MIN = 0.01

# hue the float64 path produces for pure white (reference nphusl/_simd.c:79)
WHITE_HUE = 19.916405993809086

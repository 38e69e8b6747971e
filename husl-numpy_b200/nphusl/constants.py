"""Colour-science constants of the HUSL (HSLuv) transform.

Same public names and values as the reference's ``nphusl/constants.py:3-21``
(``M``, ``M_INV``, ``REF_*``, ``KAPPA``, ``EPSILON``) because callers and the
reference's Cython module import them by name.  They are data, not code: the
sRGB <-> XYZ matrices, the D65 white point in u'v', and the CIE L* break points.

The CUDA kernels do not read this module: ``csrc/gen_coeffs.py`` derives their
coefficients (``csrc/husl_coeffs.h``) from its own copy of the same numbers;
``tests/test_consts.py`` checks that the two copies and the header agree.
"""

# linear sRGB channel <- (X, Y, Z)
_R_FROM_XYZ = (3.240969941904521, -1.537383177570093, -0.498610760293)
_G_FROM_XYZ = (-0.96924363628087, 1.87596750150772, 0.041555057407175)
_B_FROM_XYZ = (0.055630079696993, -0.20397695888897, 1.056971514242878)
M = [list(_R_FROM_XYZ), list(_G_FROM_XYZ), list(_B_FROM_XYZ)]

# X, Y, Z <- linear sRGB (r, g, b)
_X_FROM_RGB = (0.41239079926595, 0.35758433938387, 0.18048078840183)
_Y_FROM_RGB = (0.21263900587151, 0.71516867876775, 0.072192315360733)
_Z_FROM_RGB = (0.019330818715591, 0.11919477979462, 0.95053215224966)
M_INV = [list(_X_FROM_RGB), list(_Y_FROM_RGB), list(_Z_FROM_RGB)]

# D65 reference white: XYZ and its (u', v') chromaticity
REF_X, REF_Y, REF_Z = 0.95045592705167, 1.0, 1.089057750759878
REF_U, REF_V = 0.19783000664283, 0.46831999493879

# CIE L*: slope of the linear segment and the Y value where the cube root takes over
KAPPA, EPSILON = 903.2962962, 0.0088564516

# lightness clamps of the HUSL <-> LCh step (reference nphusl/nphusl.py:106-107)
L_MAX, L_MIN = 99.99, 0.01

# hue the float64 path produces for pure white (reference nphusl/_simd.c:79)
WHITE_HUE = 19.916405993809086

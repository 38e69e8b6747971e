"""One set of colour constants, three copies: the host package
(nphusl/constants.py), the coefficient generator (csrc/gen_coeffs.py) and the
generated header the kernels include (csrc/husl_coeffs.h).  They must agree, and
the closed-form identities the kernels rely on must hold in float64."""
import importlib.util
import io
import os

import numpy as np

from conftest import ROOT
from nphusl import constants

CSRC = os.path.join(ROOT, "husl-numpy_b200", "csrc")


def _gen():
    spec = importlib.util.spec_from_file_location("gen_coeffs", os.path.join(CSRC, "gen_coeffs.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_host_constants_equal_generator_inputs():
    g = _gen()
    assert np.array_equal(np.array(constants.M), g.M)
    assert np.array_equal(np.array(constants.M_INV), g.M_INV)
    for name in ("REF_U", "REF_V", "KAPPA", "EPSILON", "L_MAX", "L_MIN", "WHITE_HUE"):
        assert getattr(constants, name) == getattr(g, name), name
    # the published matrices are inverses of each other to their printed precision
    assert np.abs(np.array(constants.M) @ np.array(constants.M_INV) - np.eye(3)).max() < 1e-9


def test_committed_header_is_what_the_generator_emits():
    g = _gen()
    buf = io.StringIO()
    g.emit(g.derive(), buf)
    assert buf.getvalue() == open(os.path.join(CSRC, "husl_coeffs.h")).read()


def test_closed_forms_match_the_six_line_formula():
    g = _gen()
    worst_s, worst_t = g.verify(g.derive(), n=500, seed=1)
    assert worst_s < 1e-9 and worst_t < 1e-9


def test_white_point_annihilates_the_grey_axis():
    c = _gen().derive()
    # Nu, Nv have (numerically) no g component: hue is independent of the grey level
    assert abs(c["U_S"]) < 1e-12 and abs(c["V_S"]) < 1e-12
    assert abs(c["Y_S"] - 1.0) < 1e-13
    assert abs(c["AN"] - c["REF_V4"]) < 1e-12

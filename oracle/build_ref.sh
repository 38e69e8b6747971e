#!/bin/bash
# Build the reference's own CPU implementation of the hot path into oracle/_ref/.
#
# TEST INFRASTRUCTURE ONLY.  Nothing here is shipped or imported by the product
# package; tests/, __graft_entry__.smoke() and bench.py's CPU legs are the only
# users.  Sources are compiled WHERE THEY LIE under $REF (default
# /root/reference); only build outputs (.so files, generated lookup tables) are
# written, all of them under oracle/_ref/ (git-ignored, but shipped to the GPU
# box by gpurun).  No reference source file is copied into the repository.
#
# Recipe follows SURVEY.md App. B:
#   * LUT/light.py + LUT/chroma.py (the two commands of make_lookup_tables:8-9)
#     generate _light_lookup.{c,h} / _chroma_lookup.{c,h}; they import the
#     reference's NumPy path, which needs the removed np.float/np.int aliases
#     -> injected by a throw-away sitecustomize.py on PYTHONPATH.
#   * nphusl/_simd.c (+ _linear_lookup.c, _scale_const.c) is compiled four ways
#     (setup.py:99-128 flag sets) into plain shared objects exporting
#     rgb_to_husl_nd (_simd.h:4-5).
#   * the reference package itself (its Python modules byte-compiled, its two
#     .pyx modules cythonized and compiled) becomes oracle/_ref/refpkg/nphusl_ref:
#     see step 3.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${REF:-/root/reference}"
OUT="$HERE/_ref"
if [ ! -d "$REF/nphusl" ]; then
    echo "build_ref: $REF not present (GPU box?) - keeping prebuilt oracle/_ref" >&2
    exit 0
fi
PY="${PYTHON:-python}"
# NB: the image exports CC=/opt/gcc/bin/gcc, which lacks libgomp.spec (SURVEY App. B fix 2)
CC="${REF_CC:-/usr/bin/gcc}"
mkdir -p "$OUT/gen" "$OUT/shim" "$OUT/scratch"
printf 'import numpy as np\nnp.float = float\nnp.int = int\n' > "$OUT/shim/sitecustomize.py"

# 1. generated lookup tables (outputs only under oracle/_ref/gen)
if [ ! -s "$OUT/gen/_chroma_lookup.c" ] || [ ! -s "$OUT/gen/_light_lookup.c" ]; then
    ( cd "$REF" && PYTHONDONTWRITEBYTECODE=1 PYTHONWARNINGS=ignore \
        PYTHONPATH="$OUT/shim:$REF/LUT:$REF" \
        "$PY" LUT/light.py -y 0.070 0.350 -t uint16_t -o "$OUT/gen/_light_lookup" -s 1024 )
    ( cd "$REF" && PYTHONDONTWRITEBYTECODE=1 PYTHONWARNINGS=ignore \
        PYTHONPATH="$OUT/shim:$REF/LUT:$REF" \
        "$PY" LUT/chroma.py -t uint16_t -o "$OUT/gen/_chroma_lookup" -u 1024 -l 1024 )
fi

# 1b. the same tables as FLOAT tables (`make_lookup_tables float 1024 1024`, make_lookup_tables:4-6:
#     "{:7.4f}" / "{:7.3f}" literals, LIGHT_SCALE = CHROMA_SCALE = 1)
mkdir -p "$OUT/gen_float"
if [ ! -s "$OUT/gen_float/_chroma_lookup.c" ] || [ ! -s "$OUT/gen_float/_light_lookup.c" ]; then
    ( cd "$REF" && PYTHONDONTWRITEBYTECODE=1 PYTHONWARNINGS=ignore \
        PYTHONPATH="$OUT/shim:$REF/LUT:$REF" \
        "$PY" LUT/light.py -y 0.070 0.350 -t float -o "$OUT/gen_float/_light_lookup" -s 1024 )
    ( cd "$REF" && PYTHONDONTWRITEBYTECODE=1 PYTHONWARNINGS=ignore \
        PYTHONPATH="$OUT/shim:$REF/LUT:$REF" \
        "$PY" LUT/chroma.py -t float -o "$OUT/gen_float/_chroma_lookup" -u 1024 -l 1024 ) >/dev/null 2>&1
fi

# 2. _simd.c variants
SRC="$REF/nphusl/_simd.c $REF/nphusl/_linear_lookup.c $REF/nphusl/_scale_const.c"
INC="-I$REF/nphusl -I$OUT/gen"
BASE="-fcommon -shared -fPIC -std=c99 -fopenmp -w"
LUTS="$OUT/gen/_light_lookup.c $OUT/gen/_chroma_lookup.c -DUSE_LIGHT_LUT -DUSE_CHROMA_LUT -DUSE_HUE_ATAN2_APPROX"
$CC -O3 -ffast-math $BASE $INC $SRC $LUTS                      -o "$OUT/libsimd_lut.so" -lm
$CC -O2             $BASE $INC $SRC $LUTS                      -o "$OUT/libsimd_lut_strict.so" -lm
$CC -O3 -ffast-math $BASE $INC $SRC $LUTS -DINTERPOLATE_CHROMA -o "$OUT/libsimd_lut_interp.so" -lm
$CC -O2             $BASE $INC $SRC $LUTS -DINTERPOLATE_CHROMA -o "$OUT/libsimd_lut_interp_strict.so" -lm
$CC -O3 -ffast-math $BASE $INC $SRC                            -o "$OUT/libsimd_exact.so" -lm
INCF="-I$REF/nphusl -I$OUT/gen_float"
LUTF="$OUT/gen_float/_light_lookup.c $OUT/gen_float/_chroma_lookup.c -DUSE_LIGHT_LUT -DUSE_CHROMA_LUT -DUSE_HUE_ATAN2_APPROX"
$CC -O2             $BASE $INCF $SRC $LUTF                      -o "$OUT/libsimd_lutf_strict.so" -lm
$CC -O2             $BASE $INCF $SRC $LUTF -DINTERPOLATE_CHROMA -o "$OUT/libsimd_lutf_interp_strict.so" -lm

# 3. the reference PACKAGE itself, compiled: oracle/_ref/refpkg/nphusl_ref/
#    * its four Python modules (__init__, nphusl, transform, constants) are byte-compiled WHERE THEY LIE into
#      marshalled code objects (<name>.code, a build output like the .so files: no reference source enters
#      the repository) with a generated three-line loader each;
#    * _cython_opt.pyx and _simd_opt.pyx are cythonized into a scratch dir and compiled with the flags of
#      the reference's setup.py:99-128 (default build: light LUT + chroma LUT + atan2 approximation).
#    Every import inside the package is relative, so it loads under the name `nphusl_ref` next to this
#    repository's own `nphusl`.  bench.py --impl reference and the cpu_baseline leg drive the reference through
#    ITS public API (nphusl_ref.to_husl / to_rgb / to_hue) and ITS stock dispatch (SIMD, then Cython); none of
#    this repository's host code is on that path.  The only patches are the two from SURVEY App. B that a
#    current toolchain forces: np.float / np.int aliases (set by oracle.ref_package() before the import) and
#    Cython language level 2 for _simd_opt.pyx's integer division (a command-line flag, not an edit).
PKG="$OUT/refpkg/nphusl_ref"
rm -rf "$OUT/refhost" "$OUT/refpkg"
mkdir -p "$PKG"
"$PY" - "$REF/nphusl" "$PKG" <<'PYEOF'
# Byte-compile the reference's Python modules where they lie.  The code objects are stored as
# <name>.code (marshal; *.pyc files do not travel to the GPU box) next to a generated three-line
# loader <name>.py that executes them in the module's namespace.
import marshal, os, sys
src, dst = sys.argv[1:3]
LOADER = ("# generated by oracle/build_ref.sh: runs the byte-compiled reference module nphusl/%s.py (see %s.code)\n"
          "import marshal as _m, os as _o\n"
          "exec(_m.load(open(_o.path.join(_o.path.dirname(__file__), %r), 'rb')), globals())\n")
for name in ("__init__", "nphusl", "transform", "constants"):
    with open(os.path.join(src, name + ".py")) as f:
        code = compile(f.read(), "<reference>/nphusl/%s.py" % name, "exec", dont_inherit=True)
    with open(os.path.join(dst, name + ".code"), "wb") as f:
        marshal.dump(code, f)
    with open(os.path.join(dst, name + ".py"), "w") as f:
        f.write(LOADER % (name, name, name + ".code"))
PYEOF
NPINC="$($PY -c 'import numpy; print(numpy.get_include())')"
PYINC="$($PY -c 'import sysconfig; print(sysconfig.get_paths()["include"])')"
EXT="$($PY -c 'import sysconfig; print(sysconfig.get_config_var("EXT_SUFFIX"))')"
CYFLAGS="-fopenmp -O3 -ffast-math -ftree-vectorize -mtune=native -shared -fPIC -w -DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION"
"$PY" -m cython -3 --module-name nphusl_ref._cython_opt \
    "$REF/nphusl/_cython_opt.pyx" -o "$OUT/scratch/_cython_opt.c" >/dev/null 2>&1
$CC $CYFLAGS -I"$NPINC" -I"$PYINC" "$OUT/scratch/_cython_opt.c" -o "$PKG/_cython_opt$EXT" -lm
"$PY" -m cython -2 --module-name nphusl_ref._simd_opt -I "$REF/nphusl" \
    "$REF/nphusl/_simd_opt.pyx" -o "$OUT/scratch/_simd_opt.c" >/dev/null 2>&1
$CC $CYFLAGS -fcommon -std=c99 -I"$NPINC" -I"$PYINC" $INC "$OUT/scratch/_simd_opt.c" $SRC $LUTS \
    -o "$PKG/_simd_opt$EXT" -lm
rm -rf "$OUT/scratch"
echo "build_ref: built $(ls "$OUT"/*.so | wc -l) _simd variants + refpkg/nphusl_ref ($(ls "$PKG" | tr '\n' ' '))"

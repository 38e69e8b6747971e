#!/bin/bash
# Run profiles/tools/ab.py once per variant library in build/var (see build_variants.sh).
# usage: ab_variants.sh out.jsonl name1 name2 ...
out=$1; shift
: > "$out"
for v in "$@"; do
  NPHUSL_CUDA_LIB=$PWD/build/var/lib_$v.so NPHUSL_VARIANT=$v python profiles/tools/ab.py >> "$out" 2>> "$out.err"
done
python - "$out" <<'PY'
import json, sys
for line in open(sys.argv[1]):
    r = json.loads(line)
    print("%-10s" % r["variant"], " ".join("%s=%.4f" % (k, v["ms"]) for k, v in r.items() if isinstance(v, dict) and k in
          ("to_husl_f32", "to_rgb_f32", "to_hue_f32", "to_hue_f64", "pipeline_f32", "fused_edit_u8", "to_husl_f64", "to_rgb_f64")))
PY

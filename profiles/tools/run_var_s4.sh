for v in base f_256_3_2 f_256_2_3 f_384_2_2 f_128_4_2 f_512_1_2 e1024 e384x3 e768x2 e256x4 base; do
  NPHUSL_CUDA_LIB=$PWD/build/var/lib_$v.so NPHUSL_VARIANT=$v python profiles/tools/ab.py >> gpurun_out/s4_var.jsonl 2>> gpurun_out/s4_var.err
done
python - <<'PY'
import json
for line in open("gpurun_out/s4_var.jsonl"):
    r = json.loads(line)
    print("%-10s" % r["variant"], " ".join("%s=%.4f" % (k, r[k]["ms"]) for k in ("frgb32_to_husl32","frgb64_to_husl64","husl64_to_frgb64","husl32_to_frgb32","fused_edit_u8","fused_edit_chw_u8","nv12_edit_rgb")))
PY

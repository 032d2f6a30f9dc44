for v in seq2 seq1; do
 for f in 0.001 0.01; do
  SRUKF_LIB_PATH=$PWD/variants/lib_$v.so timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --adversarial-frac $f > gpurun_out/r03a_${v}_adv_$f.json 2> gpurun_out/r03a_${v}_adv_$f.err; echo "$v adv $f rc=$?"
  python - gpurun_out/r03a_${v}_adv_$f.json <<'PY'
import json, sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); k=d["roofline"]["kernel_ms"]; st=d["steps"]
print("  value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"],2), {a: round(b/st,2) for a,b in k.items()}, "n_fallback", d["stats"]["n_fallback"], "nees", d["stats"]["nees"], "flags", d["stats"]["flag_or"])
PY
 done
done
SRUKF_LIB_PATH=$PWD/variants/lib_seq2.so timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "adversarial or forced_fallback" 2>&1 | tail -2

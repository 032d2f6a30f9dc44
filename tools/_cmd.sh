timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "adversarial or forced_fallback" 2>&1 | tail -3
for f in 0.001 0.01; do
  timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --adversarial-frac $f > gpurun_out/r03k_adv_$f.json 2> gpurun_out/r03k_adv_$f.err; echo "adv $f rc=$?"
  python - gpurun_out/r03k_adv_$f.json <<'PY'
import json, sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); k=d["roofline"]["kernel_ms"]; st=d["steps"]
print("  value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"],2), {a: round(b/st,2) for a,b in k.items()}, "n_fallback", d["stats"]["n_fallback"], "nees", d["stats"]["nees"], "flags", d["stats"]["flag_or"])
PY
done

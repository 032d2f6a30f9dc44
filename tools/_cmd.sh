export SRUKF_LIB_PATH=$PWD/variants/lib_mq.so
for cfg in "def|" "mq5|SRUKF_UPDATE_MQ5=1"; do
  IFS='|' read -r label envs <<< "$cfg"
  env $envs timeout 300 python bench.py --sweep 10:131072,20:131072,33:65536 --steps 3 --warmup 3 > gpurun_out/r02p_$label.jsonl 2> gpurun_out/r02p_$label.err; echo "$label rc=$?"
  python - gpurun_out/r02p_$label.jsonl <<'PY'
import json, sys
for l in open(sys.argv[1]):
    d=json.loads(l); print("  L", d["landmarks"], "B", d["filters"], "rate", round(d["value"]), "ms", round(d["ms_per_step"],2), "frac", round(d["frac_fp64_peak"],3), {k: round(v,2) for k,v in d["kernel_ms_per_step"].items()}, d["parity"]["ok"])
PY
done
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -5

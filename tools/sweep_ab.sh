#!/bin/bash
# A/B of bench.py --sweep entries on the GPU box: [LIBV=<variant>] tools/sweep_ab.sh <tag> <L,L,..> "label|ENV=1 .." ...
tag=$1; Ls=$2; shift 2
for e in "$@"; do
  IFS='|' read -r label envs <<< "$e"
  lib=""; [ -n "$LIBV" ] && lib="SRUKF_LIB_PATH=$PWD/variants/lib_$LIBV.so"
  env $lib $envs timeout 300 python bench.py --sweep $Ls --steps 3 --warmup 3 > gpurun_out/${tag}_$label.jsonl 2> gpurun_out/${tag}_$label.err
  python - "$label" gpurun_out/${tag}_$label.jsonl <<'PY'
import json, sys
for l in open(sys.argv[2]):
    d = json.loads(l)
    print(sys.argv[1], "L", d["landmarks"], "B", d["filters"], "ms", round(d["ms_per_step"], 2), "frac", round(d["frac_fp64_peak"], 3),
          {k: round(v, 2) for k, v in d["kernel_ms_per_step"].items()}, "parity", d["parity"]["ok"], d["parity"]["relerr_P"], "fb", d["n_fallback"])
PY
done

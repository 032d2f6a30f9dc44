#!/bin/bash
# A/B runs on the GPU box: tools/ab_bench.sh <tag> "label|variant|ENV=1 ENV2=2" ...   (variant = name under variants/, or - for the in-tree build)
# Each entry: quick parity against the oracle, then a short bench; one summary line per entry.
tag=$1; shift
mkdir -p gpurun_out
for e in "$@"; do
  IFS='|' read -r label var envs <<< "$e"
  lib=""; [ "$var" != "-" ] && lib="SRUKF_LIB_PATH=$PWD/variants/lib_$var.so"
  out=gpurun_out/${tag}_${label}
  env $lib $envs python tools/quick_parity.py ${QP:-3:3:2 20:3:2 50:2:2} > $out.quick.log 2>&1; qrc=$?
  if [ $qrc -ne 0 ] && [ -z "$BENCH_ANYWAY" ]; then echo "$label parity FAILED (rc=$qrc): bench skipped"; tail -3 $out.quick.log; continue; fi
  env $lib $envs timeout ${BENCH_TIMEOUT:-120} python bench.py --steps ${STEPS:-5} --warmup 3 --no-cpu-baseline ${BENCH_ARGS} > $out.bench.json 2> $out.bench.err; brc=$?
  python - "$label" "$qrc" "$brc" $out.bench.json <<'PY'
import json, sys
label, qrc, brc, path = sys.argv[1:5]
try:
    d = json.loads(open(path).read().strip().splitlines()[-1])
    k = d["roofline"]["kernel_ms"]; st = d["steps"]
    print(f"{label:24s} parity_rc={qrc} bench_rc={brc} value={d['value']:.0f} e2e={d['e2e']['value']:.0f} ms/step={d['ms_per_step']:.2f} "
          f"predict={k['k_predict']/st:.2f} gain={k['k_gain']/st:.2f} update={k['k_update']/st:.2f} frac={d['roofline']['frac']:.3f} "
          f"whole={d['roofline']['whole_step_frac']:.3f} rmse={d['stats']['rmse_xy']:.5f} nees={d['stats']['nees']:.4f} flags={d['stats']['flag_or']}")
except Exception as ex:
    print(f"{label:24s} parity_rc={qrc} bench_rc={brc} (no bench line: {ex})")
PY
  tail -1 $out.quick.log
done

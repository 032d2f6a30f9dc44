#!/bin/bash
# Standard GPU battery (run under gpurun): tests, bench (both arms), config 1, ncu launch list + full captures.
# usage: tools/gpu_battery.sh <tag> [tests] [bench] [ref] [config1] [ncu]
tag=$1; shift
what="$*"; [ -z "$what" ] && what="tests bench ref config1 ncu sweep adv"
mkdir -p gpurun_out
for w in $what; do
  case $w in
    tests)   ( time python -m pytest tests -m gpu -x -q ) > gpurun_out/${tag}_gputests.log 2>&1; tail -5 gpurun_out/${tag}_gputests.log ;;
    bench)   python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"; cat gpurun_out/${tag}_bench.json ;;
    ref)     python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err; echo "ref rc=$?"; cat gpurun_out/${tag}_bench_reference.json ;;
    config1) python tools/config1.py --steps 1000 --gpu --reference 1000 > gpurun_out/${tag}_config1.json 2> gpurun_out/${tag}_config1.err; echo "config1 rc=$?"; cat gpurun_out/${tag}_config1.json ;;
    ncu)     ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_bench.log 2>&1
             # gpurun brings back at most 64 MiB and a `--set full` report is ~26 MB: the full set (with source) for the
             # dominant kernel, a metric list (csv, even a metrics-only .ncu-rep is ~20 MB) for the other two
             ncu --set full --clock-control none --import-source on -k regex:k_update -s 4 -c 1 -o gpurun_out/${tag}_k_update -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_k_update.log 2>&1
             M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,launch__grid_size,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,lts__t_sector_hit_rate.pct,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed
             for k in k_gain k_predict; do
               ncu --metrics $M --clock-control none -k regex:$k -s 4 -c 1 --csv --log-file gpurun_out/${tag}_$k.metrics.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_$k.log 2>&1
             done; ls -la gpurun_out/${tag}_*.ncu-rep ;;
    smoke)   python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 ;;
    sweep)   timeout 900 python bench.py --sweep 10,20,33,50,66,100,200 --steps 3 --warmup 3 > gpurun_out/${tag}_sweep_config4.jsonl 2> gpurun_out/${tag}_sweep.err; echo "sweep rc=$?"
             python - gpurun_out/${tag}_sweep_config4.jsonl <<'PY'
import json, sys
for l in open(sys.argv[1]):
    d = json.loads(l)
    print("  L", d["landmarks"], "n", d["state_dim"], "B", d["filters"], "rate", round(d["value"]), "ms", round(d["ms_per_step"], 2), "frac", round(d["frac_fp64_peak"], 3),
          {k: round(v, 2) for k, v in d["kernel_ms_per_step"].items()}, "fallback", d["n_fallback"], "parity", d["parity"]["ok"], d["parity"]["relerr_P"], d["noise"])
PY
             ;;
    adv)     for f in 0.001 0.01; do
               timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --adversarial-frac $f > gpurun_out/${tag}_adv_$f.json 2> gpurun_out/${tag}_adv_$f.err; echo "adv $f rc=$?"; cut -c1-200 gpurun_out/${tag}_adv_$f.json
             done ;;
    lib)     cp cv_monoslam_b200/libsrukf_b200.so gpurun_out/${tag}_libsrukf_b200.so ;;
    qp)      python tools/quick_parity.py 3:3:2 20:3:2 50:2:2 2>&1 | tail -2 ;;
  esac
done

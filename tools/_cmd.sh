N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r03i_bench_n$N.json 2> gpurun_out/r03i_bench_n$N.err; echo "n$N rc=$?"; python - gpurun_out/r03i_bench_n$N.json <<'PY'
import json, sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("n_gpus", d["n_gpus"], "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"],2), "stats", d["stats"])
PY
tail -3 gpurun_out/r03i_bench_n$N.err

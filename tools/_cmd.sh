nvidia-smi -L | head -3
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "sharding" 2>&1 | tail -2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02t_bench_n2.json 2> gpurun_out/r02t_bench_n2.err; echo "n2 rc=$?"; cat gpurun_out/r02t_bench_n2.json | cut -c1-600
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 3 > gpurun_out/r02t_ref_n2.json 2> gpurun_out/r02t_ref_n2.err; echo "ref n2 rc=$?"; cat gpurun_out/r02t_ref_n2.json | cut -c1-300

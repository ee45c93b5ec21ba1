cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests -m gpu -q > gpurun_out/r2f2_pytest_2gpu.log 2>&1; tail -3 gpurun_out/r2f2_pytest_2gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2f2_bench_n2.json 2> gpurun_out/r2f2_bench_n2.err; tail -c 600 gpurun_out/r2f2_bench_n2.err; head -c 700 gpurun_out/r2f2_bench_n2.json

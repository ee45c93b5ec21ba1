cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -k "peer or c5" 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/gpu_peer_minloc.py 2>&1 | grep OK | tee gpurun_out/r2f3_peer_minloc.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2f3_bench_n2.json 2> gpurun_out/r2f3_bench_n2.err; tail -c 300 gpurun_out/r2f3_bench_n2.err; python -c "
import json; d=json.load(open('gpurun_out/r2f3_bench_n2.json')); print(d['value'], d['ms_per_step'], d['roofline']['kernel_us'], d['parity'], d['e2e']['value'])"

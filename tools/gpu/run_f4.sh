cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/gpu_dist_push.py 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | tee gpurun_out/r2f4_dist_push.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2f4_bench_n2.json 2> gpurun_out/r2f4_bench_n2.err; python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2f4_bench_n2.json') if l.startswith('{')][0]); print(d['value'], d['ms_per_step'], d['roofline']['kernel_us'], d['parity'], d['e2e']['value'], d['tick_latency'])"

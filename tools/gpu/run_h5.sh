cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -k "replay or push or pdl or peer or dist_push" 2>&1 | tail -3
python tools/gpu_replay_length_timing.py 2>&1 | tail -4
python bench.py --no-cpu --no-extras > gpurun_out/r2h5_bench.json 2> gpurun_out/r2h5_bench.err; tail -c 500 gpurun_out/r2h5_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r2h5_bench.json')); print(d['value'], d['ms_per_step']); print({k:v for k,v in d['e2e'].items() if 'api' not in k}); print(d['tick_latency']['p50_us'], d['parity'].get('e2e_replay_equals_push'))"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2h5_bench_n2.json 2> gpurun_out/r2h5_bench_n2.err; tail -c 300 gpurun_out/r2h5_bench_n2.err; python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2h5_bench_n2.json') if l.startswith('{')][0]); print(d['value'], d['ms_per_step'], {k:v for k,v in d['e2e'].items() if 'api' not in k}, d['parity'])"

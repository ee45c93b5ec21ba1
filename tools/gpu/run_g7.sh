cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -k "pdl" 2>&1 | tail -3
python bench.py > gpurun_out/r2g7_bench.json 2> gpurun_out/r2g7_bench.err; tail -c 300 gpurun_out/r2g7_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r2g7_bench.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['other_configs']['C5_1gpu_lookback_1048576x50'])"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2g7_bench_n2.json 2> gpurun_out/r2g7_bench_n2.err; python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2g7_bench_n2.json') if l.startswith('{')][0]); print(d['value'], d['ms_per_step'], d['roofline']['kernel_us'], d['l2_flushed']['ms_per_step'], d['parity'], d.get('strong_scaling_efficiency_vs_c5_on_1_gpu'), d['scaling_base']['ms_per_tick'])"

cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for n in 8 4; do
timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2955$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r2j6_bench_n$n.json 2> gpurun_out/r2j6_bench_n$n.err
python -c "
import json; d=json.load(open('gpurun_out/r2j6_bench_n$n.json')); print($n, d['value'], d['ms_per_step'], d['roofline']['kernel_us'], d['e2e']['value'], d['e2e']['sync_push_value'], d['tick_latency']['p50_us'], d.get('strong_scaling_efficiency_vs_c5_on_1_gpu'), d['parity'])"
done

cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for n in 8 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r2g8_bench_n$n.json 2> gpurun_out/r2g8_bench_n$n.err
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2g8_bench_n$n.json') if l.startswith('{')][0]); print($n, d['value'], d['ms_per_step'], d['roofline']['kernel_us'], d['parity']['ranks_agree'], d['parity']['vs_nccl'], d['e2e']['value'], d['tick_latency'], d.get('scaling_base_value'))"
done

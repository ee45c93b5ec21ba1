cd $GRAFT_REPO_ROOT
timeout 400 python -m pytest tests -m gpu -q > gpurun_out/r2j5_pytest_2gpu.log 2>&1; tail -2 gpurun_out/r2j5_pytest_2gpu.log
timeout 250 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2j5_bench_n2.json 2> gpurun_out/r2j5_bench_n2.err; tail -c 200 gpurun_out/r2j5_bench_n2.err; python -c "
import json; d=json.load(open('gpurun_out/r2j5_bench_n2.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['sync_push_value'], d['parity']['e2e_replay_equals_push'], d['parity']['e2e_replay_ranks_agree'])"

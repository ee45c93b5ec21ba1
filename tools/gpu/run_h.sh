cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests -m gpu -q -k "peer_minloc" > gpurun_out/r2h_pytest.log 2>&1; tail -15 gpurun_out/r2h_pytest.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/gpu_peer_minloc.py > gpurun_out/r2h_peer.txt 2>&1; tail -4 gpurun_out/r2h_peer.txt
PEER_N=1048576 PEER_W=50 PEER_TICKS=40 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 tools/gpu_peer_minloc.py > gpurun_out/r2h_peer_c5.txt 2>&1; tail -3 gpurun_out/r2h_peer_c5.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2h_bench_n2.json 2> gpurun_out/r2h_bench_n2.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/r2h_bench_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29536 tools/gpu_dist_push.py > gpurun_out/r2h_dist_push.txt 2>&1; tail -3 gpurun_out/r2h_dist_push.txt

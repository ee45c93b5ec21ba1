cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2h1_pytest_2gpu.log 2>&1; tail -2 gpurun_out/r2h1_pytest_2gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --impl reference > gpurun_out/r2h1_bench_ref.json 2> gpurun_out/r2h1_bench_ref.err; wc -l gpurun_out/r2h1_bench_ref.json
python bench.py > gpurun_out/r2h1_bench.json 2> gpurun_out/r2h1_bench.err; wc -l gpurun_out/r2h1_bench.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2h1_bench_n2.json 2> gpurun_out/r2h1_bench_n2.err; wc -l gpurun_out/r2h1_bench_n2.json; head -c 100 gpurun_out/r2h1_bench_n2.json; echo
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2h1_bench_ref_n2.json 2> gpurun_out/r2h1_bench_ref_n2.err; wc -l gpurun_out/r2h1_bench_ref_n2.json; head -c 100 gpurun_out/r2h1_bench_ref_n2.json; echo
python -c "
import json
d=json.load(open('gpurun_out/r2h1_bench.json')); r=json.load(open('gpurun_out/r2h1_bench_ref.json'))
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], 'ref', r['value'], 'e2e ratio', d['e2e']['value']/r['value'], 'device ratio', d['value']/r['value'])"

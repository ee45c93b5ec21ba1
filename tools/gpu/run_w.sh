cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2w_pytest.log 2>&1; tail -5 gpurun_out/r2w_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err; echo "bench rc=$?"; tail -c 400 gpurun_out/r2w_bench.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2w_bench_ref.json 2> gpurun_out/r2w_bench_ref.err; echo "ref rc=$?"
python tools/gpu_planner_timing.py > gpurun_out/r2w_planner.txt 2>&1; cat gpurun_out/r2w_planner.txt

cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2g_pytest.log 2>&1; tail -8 gpurun_out/r2g_pytest.log
python tools/gpu_c3_timing.py > gpurun_out/r2g_c3.txt 2>&1; cat gpurun_out/r2g_c3.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2

cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2g1_pytest_2gpu.log 2>&1; tail -3 gpurun_out/r2g1_pytest_2gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1

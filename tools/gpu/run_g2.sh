cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2g2_pytest.log 2>&1; tail -5 gpurun_out/r2g2_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2

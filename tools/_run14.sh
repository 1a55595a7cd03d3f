cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "hybrid" 2>&1 | tail -15 > gpurun_out/r02_pytest14.log
python tools/c5_case.py > gpurun_out/r02_c5.json 2> gpurun_out/r02_c5.err

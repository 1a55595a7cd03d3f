cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_parity.py tests/test_widen_gpu.py -m gpu -x -q -k "cg or pcg or bsr3 or static or shell or constrained or modal or kuhn20 or precond" 2>&1 | tail -6 > gpurun_out/r02_pytest17.log
python tools/c2_case.py > gpurun_out/r02_c2_17.json 2> gpurun_out/r02_c2_17.err
FEMB_CG_CLASSIC=1 python tools/c2_case.py >> gpurun_out/r02_c2_17.json 2>> gpurun_out/r02_c2_17.err

cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "c3d10 or c3d8 or c3d6 or quadratic or consistent_mass or mixed or stress or static_structure or full_size_properties_c2 or bsr3" 2>&1 | tail -6 > gpurun_out/r02_pytest15.log
for e in 16 8 32 4; do FEMB_SOLID_EPB=$e python tools/c2_case.py >> gpurun_out/r02_c2_15.json 2>> gpurun_out/r02_c2_15.err; done
FEMB_SOLID_TILE=1 FEMB_SOLID_EPB=4 python tools/c2_case.py >> gpurun_out/r02_c2_15.json 2>> gpurun_out/r02_c2_15.err

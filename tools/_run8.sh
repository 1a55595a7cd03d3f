cd $GRAFT_REPO_ROOT
export FEMB_ASM_VERBOSE=1
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "assembly or kuhn20 or edge_cases or full_size_properties_c4 or hybrid" 2>&1 | tail -12 > gpurun_out/r02_pytest8.log
for R in 512 256 128; do FEMB_ASM_BLOCK_ROWS=$R python tools/asm_rate.py --n 220 2>&1 | tail -2 >> gpurun_out/r02_asm8.log; done
FEMB_ASM_BLOCK_ROWS=256 python tools/asm_rate.py --n 220 --jitter 0.2 2>&1 | tail -2 >> gpurun_out/r02_asm8.log

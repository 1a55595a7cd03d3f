cd $GRAFT_REPO_ROOT
export FEMB_ASM_VERBOSE=1
for R in 512 256 128; do FEMB_ASM_BLOCK_ROWS=$R python tools/asm_rate.py --n 220 2>&1 | tail -2 >> gpurun_out/r02_asm6.log; done
FEMB_ASM_TILES=1 python tools/asm_rate.py --n 220 2>&1 | tail -1 >> gpurun_out/r02_asm6.log

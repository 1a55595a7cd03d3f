cd $GRAFT_REPO_ROOT
for v in "512 512" "512 256" "256 512" "256 128" "384 256" "384 512"; do set -- $v; FEMB_ASM_BLOCK_ROWS=$1 FEMB_ASM_BLOCK_THREADS=$2 python tools/asm_rate.py --n 220 2>&1 | tail -1 >> gpurun_out/r02_asm10.log; done

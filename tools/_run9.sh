cd $GRAFT_REPO_ROOT
FEMB_ASM_BLOCK_ROWS=256 timeout 600 ncu --set full --clock-control none --import-source on -k regex:assemble_p1_blocks -s 1 -c 1 -o gpurun_out/r02_asm_blocks_v2 -f python tools/asm_rate.py --n 110 --reps 2 > gpurun_out/r02_ncu9.log 2>&1

cd $GRAFT_REPO_ROOT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29681 tools/dist_parity.py --reps 3 2>&1 | grep PARITY > gpurun_out/r02_parity23.log
FEMB_DIST_NO_START_BARRIER=1 FEMB_DIST_NO_UPLOAD=1 timeout 600 $TR --master-port 29682 tools/dist_parity.py --reps 3 2>&1 | grep PARITY >> gpurun_out/r02_parity23.log
FEMB_NO_PDL=1 timeout 600 $TR --master-port 29683 tools/dist_parity.py --reps 3 2>&1 | grep PARITY >> gpurun_out/r02_parity23.log

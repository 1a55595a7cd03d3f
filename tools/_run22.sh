cd $GRAFT_REPO_ROOT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 8 --master-port 29671 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_bench22_8gpu_s20.json 2> gpurun_out/r02_bench22_8gpu_s20.err
FEMB_DIST_TRACE=1 timeout 600 $TR --nproc-per-node 8 --master-port 29672 bench.py --gpus 8 --steps 400 --warmup 10 > gpurun_out/r02_bench22_8gpu.json 2> gpurun_out/r02_bench22_8gpu.err
timeout 600 $TR --nproc-per-node 4 --master-port 29673 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r02_bench22_4gpu_s20.json 2> gpurun_out/r02_bench22_4gpu_s20.err

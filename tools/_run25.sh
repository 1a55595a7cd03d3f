cd $GRAFT_REPO_ROOT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 8 --master-port 29701 tests/dist_gpu_check.py 2>&1 | grep dist_gpu_check > gpurun_out/r02_dist25_check.log
timeout 600 $TR --nproc-per-node 8 --master-port 29702 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_bench25_8gpu_s20.json 2> gpurun_out/r02_bench25_8gpu_s20.err
FEMB_DIST_TRACE=1 timeout 600 $TR --nproc-per-node 8 --master-port 29703 bench.py --gpus 8 --steps 400 --warmup 10 > gpurun_out/r02_bench25_8gpu.json 2> gpurun_out/r02_bench25_8gpu.err
timeout 900 $TR --nproc-per-node 8 --master-port 29704 bench.py --gpus 8 --config 2 --steps 200 --warmup 10 > gpurun_out/r02_bench25_c2_8gpu.json 2> gpurun_out/r02_bench25_c2_8gpu.err

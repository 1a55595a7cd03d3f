cd $GRAFT_REPO_ROOT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 4 --master-port 29621 tests/dist_gpu_check.py 2>&1 | grep dist_gpu_check > gpurun_out/r02_dist_check_4_8.log
timeout 300 $TR --nproc-per-node 8 --master-port 29622 tests/dist_gpu_check.py 2>&1 | grep dist_gpu_check >> gpurun_out/r02_dist_check_4_8.log
FEMB_CHECK_N=48 timeout 300 $TR --nproc-per-node 8 --master-port 29623 tests/dist_gpu_check.py 2>&1 | grep dist_gpu_check >> gpurun_out/r02_dist_check_4_8.log
FEMB_DIST_TRACE=1 timeout 600 $TR --nproc-per-node 8 --master-port 29624 bench.py --gpus 8 --steps 400 --warmup 10 > gpurun_out/r02_bench_8gpu.json 2> gpurun_out/r02_bench_8gpu.err
FEMB_PDL_MODE=0 timeout 600 $TR --nproc-per-node 8 --master-port 29625 bench.py --gpus 8 --steps 400 --warmup 10 > gpurun_out/r02_bench_8gpu_nopdl.json 2> gpurun_out/r02_bench_8gpu_nopdl.err
FEMB_DIST_TRACE=1 timeout 600 $TR --nproc-per-node 4 --master-port 29626 bench.py --gpus 4 --steps 400 --warmup 10 > gpurun_out/r02_bench_4gpu.json 2> gpurun_out/r02_bench_4gpu.err
nvidia-smi topo -m > gpurun_out/r02_topo.txt 2>&1

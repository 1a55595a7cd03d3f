cd $GRAFT_REPO_ROOT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29661 tests/dist_gpu_check.py 2>&1 | grep dist_gpu_check > gpurun_out/r02_dist21_check.log
timeout 900 $TR --master-port 29662 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_bench21_2gpu_s20.json 2> gpurun_out/r02_bench21_2gpu.err
timeout 900 $TR --master-port 29663 bench.py --gpus 2 --steps 400 --warmup 10 > gpurun_out/r02_bench21_2gpu_s400.json 2>> gpurun_out/r02_bench21_2gpu.err
python bench.py --gpus 1 --steps 20 --warmup 5 --no-c2 --no-topo --no-cpu > gpurun_out/r02_bench21_1gpu_s20.json 2>/dev/null

cd $GRAFT_REPO_ROOT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29641 tests/dist_gpu_check.py 2>&1 | grep dist_gpu_check > gpurun_out/r02_dist19_check.log
FEMB_DIST_TRACE=1 timeout 900 $TR --master-port 29642 bench.py --gpus 2 --steps 200 --warmup 10 > gpurun_out/r02_bench19_2gpu.json 2> gpurun_out/r02_bench19_2gpu.err
python bench.py --gpus 1 --steps 200 --warmup 10 --no-c2 --no-topo --no-cpu > gpurun_out/r02_bench19_1gpu.json 2>/dev/null

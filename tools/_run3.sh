cd $GRAFT_REPO_ROOT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
timeout 300 $TR tests/dist_gpu_check.py > gpurun_out/r02_dist2_check.log 2>&1
FEMB_CHECK_N=40 timeout 300 $TR tests/dist_gpu_check.py >> gpurun_out/r02_dist2_check.log 2>&1
FEMB_DIST_TRACE=1 timeout 600 $TR bench.py --gpus 2 --steps 200 --warmup 10 > gpurun_out/r02_bench_2gpu.json 2> gpurun_out/r02_bench_2gpu.err
FEMB_DIST_CLASSIC=1 FEMB_DIST_TRACE=1 timeout 600 $TR bench.py --gpus 2 --steps 200 --warmup 10 > gpurun_out/r02_bench_2gpu_classic.json 2> gpurun_out/r02_bench_2gpu_classic.err
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "two_devices" 2>&1 | tail -5 > gpurun_out/r02_pytest3.log

cd $GRAFT_REPO_ROOT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "p1_to_p2 or c3d10 or hybrid or breakdown" 2>&1 | tail -5 > gpurun_out/r02_pytest13.log
timeout 600 $TR --master-port 29631 tests/dist_gpu_check.py > gpurun_out/r02_dist13_check.log 2>&1
timeout 900 $TR --master-port 29632 bench.py --gpus 2 --config 2 --steps 200 --warmup 10 > gpurun_out/r02_bench_c2_2gpu.json 2> gpurun_out/r02_bench_c2_2gpu.err

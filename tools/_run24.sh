cd $GRAFT_REPO_ROOT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
for i in 1 2 3 4; do
FEMB_DIST_DEBUG=1 timeout 600 $TR --master-port 2969$i bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r02_bench24_4gpu_$i.json 2> gpurun_out/r02_bench24_4gpu_$i.err
done

cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r02_pytest16.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke16.log 2>&1
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench16.json 2> gpurun_out/r02_bench16.err
python bench.py --gpus 1 --steps 200 --warmup 10 --no-c2 --no-topo --no-cpu > gpurun_out/r02_bench16b.json 2>> gpurun_out/r02_bench16.err

cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r02_pytest26.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke26.log 2>&1
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench26.json 2> gpurun_out/r02_bench26.err
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_ref26.json 2> gpurun_out/r02_ref26.err
timeout 600 ncu --set full --clock-control none -k regex:'spmv_tma_kernel|cg_merged_kernel' -s 6 -c 2 -o gpurun_out/r02_cg_loop -f python tools/cg_rate.py --n 220 --iters 50 > gpurun_out/r02_ncu26.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'spmv_tma|cg_merged|cg_init|cg_mask|assemble|pad_coords|c3d4_kernel|bucket|write_s' -c 120 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 20 --warmup 5 --no-c2 --no-cpu > /dev/null 2>&1

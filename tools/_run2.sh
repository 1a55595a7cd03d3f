cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "exact_convergence or boolean or numpy_conn or out_of_range" 2>&1 | tail -40 > gpurun_out/r02_pytest2.log
for v in "FEMB_PDL_MODE=0" "FEMB_PDL_MODE=1" "FEMB_PDL_MODE=2" "FEMB_PDL_MODE=3" "FEMB_PDL_MODE=3 FEMB_VEC_WAVES=6" "FEMB_PDL_MODE=3 FEMB_VEC_WAVES=4" "FEMB_PDL_MODE=1 FEMB_VEC_WAVES=4"; do
  env $v python tools/cg_rate.py --n 220 --iters 400 2>&1 | tail -1 >> gpurun_out/r02_cgrate2.log
done
for v in "FEMB_PDL_MODE=0" "FEMB_PDL_MODE=1" "FEMB_PDL_MODE=2" "FEMB_PDL_MODE=3" "FEMB_PDL_MODE=3 FEMB_VEC_WAVES=6" "FEMB_PDL_MODE=3 FEMB_VEC_WAVES=4" "FEMB_PDL_MODE=3 FEMB_VEC_WAVES=2"; do
  env $v python tools/cg_rate.py --n 110 --iters 1000 2>&1 | tail -1 >> gpurun_out/r02_cgrate2.log
done

set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02_pytest1.log
for v in "FEMB_NO_PDL=1 FEMB_SPMV_PIN_MB=0" "FEMB_SPMV_PIN_MB=0" "FEMB_SPMV_PIN_MB=24" "FEMB_SPMV_PIN_MB=48" "FEMB_SPMV_PIN_MB=72" "FEMB_SPMV_PIN_MB=96"; do
  env $v python tools/cg_rate.py --n 110 --iters 1000 2>&1 | tail -1 >> gpurun_out/r02_cgrate1.log
done
for v in "FEMB_NO_PDL=1" "FEMB_SPMV_PIN_MB=0"; do
  env $v python tools/cg_rate.py --n 220 --iters 400 2>&1 | tail -1 >> gpurun_out/r02_cgrate1.log
done
python bench.py --steps 200 --warmup 10 > gpurun_out/r02_bench1.json 2> gpurun_out/r02_bench1.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_ref1.json 2> gpurun_out/r02_ref1.err
nproc; free -g | head -2

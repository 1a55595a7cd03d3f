cd $GRAFT_REPO_ROOT
export FEMB_ASM_VERBOSE=1
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "assembly or kuhn20 or edge_cases or full_size_properties_c4 or smoke or hybrid" 2>&1 | tail -30 > gpurun_out/r02_pytest5.log
python - > gpurun_out/r02_asm5.log 2>&1 <<'PY'
import os, sys, torch
ROOT=os.environ["GRAFT_REPO_ROOT"]; PKG=os.path.join(ROOT,"cuda-powered-mesh-handling-and-iterative-solvers_b200")
for p in (ROOT,PKG,os.path.join(PKG,"solver")): sys.path.insert(0,p)
import element as el
from femb200 import meshgen
dev="cuda:0"
for n in (110, 220):
    c,t=meshgen.kuhn_cube(n,device=dev)
    plan=el.CsrPlan(t,c.shape[0],dev)
    vals=plan.assemble_c3d4(c,"poisson")
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): plan.assemble_c3d4(c,"poisson",out=vals,check_singular=False)
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/10
    M=t.shape[0]; N=c.shape[0]; nnz=vals.numel()
    by=M*32+N*24+nnz*8
    print(f"n={n} M={M} asm_ms={ms:.3f} Gelem/s={M/ms/1e6:.1f} frac={by/ms/1e6/6448.7:.3f}", flush=True)
    del plan, vals, c, t
PY
for R in 256 128; do FEMB_ASM_BLOCK_ROWS=$R python - >> gpurun_out/r02_asm5.log 2>&1 <<'PY'
import os, sys, torch
ROOT=os.environ["GRAFT_REPO_ROOT"]; PKG=os.path.join(ROOT,"cuda-powered-mesh-handling-and-iterative-solvers_b200")
for p in (ROOT,PKG,os.path.join(PKG,"solver")): sys.path.insert(0,p)
import element as el
from femb200 import meshgen
dev="cuda:0"
n=220
c,t=meshgen.kuhn_cube(n,device=dev)
plan=el.CsrPlan(t,c.shape[0],dev)
vals=plan.assemble_c3d4(c,"poisson")
torch.cuda.synchronize()
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): plan.assemble_c3d4(c,"poisson",out=vals,check_singular=False)
e1.record(); torch.cuda.synchronize()
ms=e0.elapsed_time(e1)/10
M=t.shape[0]; N=c.shape[0]; nnz=vals.numel()
by=M*32+N*24+nnz*8
print(f"R={os.environ['FEMB_ASM_BLOCK_ROWS']} n={n} M={M} asm_ms={ms:.3f} Gelem/s={M/ms/1e6:.1f} frac={by/ms/1e6/6448.7:.3f}", flush=True)
PY
done

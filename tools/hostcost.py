import time, torch, sys
sys.path.insert(0,'/root/repo')
import feinsum_b200 as f
from feinsum_b200.codegen import generate_cuda
from tests import einsums as E
cq=f.CudaQueue(0)
for name,e in (('grad',E.grad()),('div',E.div()),('lift',E.lift_fe())):
    n=1600
    arrs={k: torch.rand(tuple(n if not isinstance(d,int) else d for d in s), dtype=torch.float64, device='cuda') for k,s in e.arg_to_shape.items()}
    ex=generate_cuda(e).executor(cq)
    evt,outs=ex(cq,**arrs); evt.wait()
    full=dict(arrs); full.update(outs)
    for _ in range(50): ex(cq,**full)
    torch.cuda.synchronize()
    t0=time.perf_counter()
    for _ in range(2000): ex(cq,**full)
    t1=time.perf_counter()
    torch.cuda.synchronize()
    t2=time.perf_counter()
    print(name,'host us/call',(t1-t0)/2000*1e6,'incl drain',(t2-t0)/2000*1e6)

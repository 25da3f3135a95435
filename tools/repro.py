"""Cross-check a kernel variant against the simt variant at a given size (debug helper)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import feinsum_b200 as f
from feinsum_b200.codegen import generate_cuda
from tests import einsums as E

name, n = sys.argv[1], int(sys.argv[2])
params = {k: int(v) for k, v in (kv.split("=") for kv in sys.argv[3:])}
e = {"grad": E.grad, "div": E.div, "lift_fe": E.lift_fe, "lift_ef": E.lift_ef}[name]()
cq = f.CudaQueue(0)
g = torch.Generator(device=cq.torch_device).manual_seed(0)
shape = lambda s: tuple(int(d) if isinstance(d, int) else n for d in s)
arrs = {k: torch.rand(shape(s), dtype=torch.float64, device=cq.torch_device, generator=g)
        for k, s in sorted(e.arg_to_shape.items())}
evt, o1 = generate_cuda(e).with_params(variant=1, **params).executor(cq)(cq, **arrs); evt.wait()
torch.cuda.synchronize()
print("variant 1 done")
evt, o0 = generate_cuda(e).with_params(variant=2).executor(cq)(cq, **arrs); evt.wait()
for k in o1:
    rel = ((o1[k] - o0[k]).abs().max() / o0[k].abs().max()).item()
    print(name, n, k, "max rel diff", rel)
    assert rel < 1e-13

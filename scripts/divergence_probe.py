"""How much of the FBS solve time is iteration-count divergence inside a warp / tile?

Runs a workload's scenarios in three orders: as generated, sorted by their own iteration count
(the ideal order: every warp's lanes stop together), and sorted by a cheap key available before
the solve (sum of the packed |spu| rows).  Same scenarios, same per-scenario results."""
import os, sys
import numpy as np
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from gigpower_b200 import _cabi
import bench

wl = sys.argv[1] if len(sys.argv) > 1 else 'c2'
S = int(sys.argv[2]) if len(sys.argv) > 2 else 0
dev = torch.device('cuda', 0)
sol, spu_np = bench.make_probe_inputs(wl, S)
sol._dev()
sol._maybe_specialise(spu_np.shape[1])
t = sol.tables
S = spu_np.shape[1]
spu0 = torch.from_numpy(spu_np).to(dev)
V = torch.empty((t.n_vrows, S), dtype=torch.complex128, device=dev)
I = torch.empty((max(t.n_irows, 1), S), dtype=torch.complex128, device=dev)
it = torch.empty(S, dtype=torch.int32, device=dev); st = torch.empty(S, dtype=torch.int32, device=dev)
df = torch.empty(S, dtype=torch.float64, device=dev)
Vm = torch.empty((t.n_vrows, S), dtype=torch.float64, device=dev)
ls = torch.empty(S, dtype=torch.complex128, device=dev)
stream = torch.cuda.current_stream()


def run(spu):
    _cabi.check(_cabi.gp_fbs_solve_post_batch(sol._feeder.handle, S, S, spu.data_ptr(), V.data_ptr(), I.data_ptr(),
                                              it.data_ptr(), st.data_ptr(), df.data_ptr(), Vm.data_ptr(),
                                              ls.data_ptr(), sol.tolerance, 100, stream.cuda_stream))


def timeit(spu, reps=20):
    for _ in range(3):
        run(spu)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        run(spu)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


ms0 = timeit(spu0)
iters = it.clone()
print(f'{wl} S={S} kernel={_cabi.gp_fbs_active_kernel(sol._feeder.handle)} as generated: {ms0*1e3:9.1f} us  '
      f'mean it {iters.float().mean():.3f} max {int(iters.max())}  '
      f'mean over warps of max-in-warp {iters[:S // 32 * 32].view(-1, 32).max(1).values.float().mean():.3f}')
order = torch.argsort(iters, stable=True)
ms1 = timeit(spu0[:, order].contiguous())
print(f'  sorted by iteration count (ideal): {ms1*1e3:9.1f} us  ({ms0/ms1:.3f}x)')
key = spu0.abs().sum(0)
order2 = torch.argsort(key)
spu2 = spu0[:, order2].contiguous()
ms2 = timeit(spu2)
w = it[:S // 32 * 32].view(-1, 32).max(1).values.float().mean()
print(f'  sorted by sum |spu| (cheap key)  : {ms2*1e3:9.1f} us  ({ms0/ms2:.3f}x)  mean max-in-warp {w:.3f}')
order3 = torch.argsort(key, descending=True)
ms3 = timeit(spu0[:, order3].contiguous())
print(f'  sorted by sum |spu| descending   : {ms3*1e3:9.1f} us  ({ms0/ms3:.3f}x)')
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    key = spu0.abs().sum(0); o = torch.argsort(key); x = spu0[:, o].contiguous()
e1.record(); torch.cuda.synchronize()
print(f'  cost of key + argsort + gather with torch: {e0.elapsed_time(e1)/20*1e3:.1f} us')

"""e2e (host-buffer path) time vs number of pipeline chunks, C2 workload."""
import os, sys, time
import numpy as np
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from gigpower_b200 import _cabi
from gigpower_b200.solution_fbs import SolutionFBS
sol = SolutionFBS(os.path.join(REPO, 'tests', 'golden', 'feeders', 'IEEE_13_Bus_allwye_noxfm_noreg.json'))
sol._dev()
t = sol.tables
S = 65536
rng = np.random.Generator(np.random.PCG64(13))
rows = np.ascontiguousarray(sol.load_weights() @ rng.uniform(0.5, 1.5, size=(S, sol.circuit.loads.num_elements)).T)
spu_h = torch.from_numpy(rows).pin_memory()
V_h = torch.empty((t.n_vrows, S), dtype=torch.complex128).pin_memory()
loss_h = torch.empty(S, dtype=torch.complex128).pin_memory()
for chunks in (1, 2, 3, 4, 8):
    _cabi.check(_cabi.gp_set_option(b'host_chunks', chunks))
    sol.solve_batch_host(spu_h, V_h, None, None, loss_h)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        sol.solve_batch_host(spu_h, V_h, None, None, loss_h)
    dt = (time.perf_counter() - t0) / 10
    print(f'chunks={chunks}: {dt*1e3:.3f} ms  -> {S/dt:.3e} solves/s')

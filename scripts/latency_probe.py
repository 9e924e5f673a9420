"""Kernel time vs batch size for the three FBS kernels (critical path of one warp vs throughput)."""
import os, sys
import numpy as np
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from gigpower_b200 import _cabi
from gigpower_b200.solution_fbs import SolutionFBS

feeder = sys.argv[1] if len(sys.argv) > 1 else 'IEEE_13_Bus_allwye_noxfm_noreg.json'
p = os.path.join(REPO, 'tests', 'golden', 'feeders', feeder)
if not os.path.exists(p):
    p = os.path.join(REPO, 'gigpower_b200', 'feeders', feeder)
sol = SolutionFBS(p)
sol._dev()
sol.specialise()
t = sol.tables
dev = torch.device('cuda', 0)
rng = np.random.Generator(np.random.PCG64(13))
W = sol.load_weights()
SIZES = [int(x) for x in os.environ.get('GP_SIZES', '320,2560,16384,65536,262144').split(',')]
KERNELS = [int(x) for x in os.environ.get('GP_KERNELS', '3,5').split(',')]
for S in SIZES:
    mult = rng.uniform(0.5, 1.5, size=(S, W.shape[1]))
    mult[:] = mult[0]                      # identical scenarios: every lane runs the same iteration count
    spu = torch.from_numpy(np.ascontiguousarray(W @ mult.T)).to(dev)
    V = torch.empty((t.n_vrows, S), dtype=torch.complex128, device=dev)
    I = torch.empty((t.n_irows, S), dtype=torch.complex128, device=dev)
    it = torch.empty(S, dtype=torch.int32, device=dev); st = torch.empty(S, dtype=torch.int32, device=dev)
    df = torch.empty(S, dtype=torch.float64, device=dev)
    line = f'S={S:7d}'
    for kern in KERNELS:
        _cabi.check(_cabi.gp_set_option(b'fbs_kernel', kern))
        stream = torch.cuda.current_stream()
        def run():
            _cabi.check(_cabi.gp_fbs_solve_batch(sol._feeder.handle, S, S, spu.data_ptr(), V.data_ptr(), I.data_ptr(),
                                                 it.data_ptr(), st.data_ptr(), df.data_ptr(), sol.tolerance, 100,
                                                 stream.cuda_stream))
        for _ in range(3):
            run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); 
        for _ in range(10):
            run()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        line += f' | k{kern}: {ms*1e3:8.1f} us it={int(it[0])}'
    print(line)

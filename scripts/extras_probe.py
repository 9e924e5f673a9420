"""Throughput of the optional per-scenario inputs (GpSolveExtras) next to the plain path:
chained (warm-started) C4 study vs the flat-started study, per-scenario taps and reactive injection on
the IEEE 13-bus regulator feeder vs the same batch without them."""
import os, sys, time
import numpy as np
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, 'tests'))
from gigpower_b200.solution_fbs import SolutionFBS
from gigpower_b200.study import DerStudy
from test_study_gpu import _c4_tables


def wall(fn, reps=3):
    """Median wall time of a call (each call synchronised): an occasional new allocator segment
    (cudaMalloc, tens of milliseconds) would otherwise dominate a mean over a few calls."""
    for _ in range(3):          # the first calls pay for NVRTC and for new allocator segments
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    return float(np.median(ts)), out


sol = SolutionFBS(os.path.join(REPO, 'gigpower_b200', 'feeders', 'ieee123_allwye_noxfm_noreg.dss'))
m_h, pv_h, mask, pen = _c4_tables(sol.tables.n_srows)
study = DerStudy(sol, m_h, pv_h, mask, pen)
chains, L = 5840, 24                      # 140,160 scenarios = one GPU's share of C4
t_flat, flat = wall(lambda: study.run(first=0, count=chains * L, chunk=chains * L))
t_chain, ch = wall(lambda: study.run_chained(chain_hours=L, first_chain=0, chains=chains))
print(f'C4 study, {chains * L} scenarios: flat start {t_flat*1e3:.1f} ms ({chains*L/t_flat:.3e} solves/s, '
      f'{flat["iterations"].mean():.2f} sweeps/solve); day chains {t_chain*1e3:.1f} ms ({chains*L/t_chain:.3e} solves/s, '
      f'{ch["iterations"].mean():.2f} sweeps/solve, {L} launches of {chains} scenarios)')
print(f'  max |vmin_chained - vmin_flat| = {np.abs(ch["vmin"] - flat["vmin"]).max():.2e}')

s13 = SolutionFBS(os.path.join(REPO, 'tests', 'golden', 'feeders', 'IEEE_13_Bus_allwye.json'),
                  taps={'reg1': 9, 'reg2': 7, 'reg3': 9})
S = 65536
rng = np.random.Generator(np.random.PCG64(1))
mult = rng.uniform(0.5, 1.5, size=(S, s13.circuit.loads.num_elements))
rows = torch.from_numpy(np.ascontiguousarray(s13.load_weights() @ mult.T)).cuda()
taps = torch.from_numpy(rng.integers(-16, 17, size=(S, 3))).cuda()
wpu = torch.zeros((s13.tables.n_vrows, S), dtype=torch.float64, device='cuda')
wpu[5] = 0.01
for name, kw in (('plain (specialised on-chip kernel)', {}), ('taps per scenario (EX build of the specialised kernel since round 2)', {'taps': taps}),
                 ('wpu per scenario (streaming kernel)', {'wpu': wpu})):
    seg0 = torch.cuda.memory_stats()['segment.all.allocated']
    t, _ = wall(lambda: s13.solve_batch(spu_rows=rows, outputs=('Vmag', 'loss'), return_device=True, **kw), reps=10)
    print('  (cudaMalloc segments created during these calls:', torch.cuda.memory_stats()['segment.all.allocated'] - seg0, ')')
    print(f'13-bus with regulators, {S} scenarios, {name}: {t*1e3:.3f} ms per call ({S/t:.3e} solves/s incl. allocation)')

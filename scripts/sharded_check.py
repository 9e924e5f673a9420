"""Run under torchrun with N >= 2 GPUs: sharded FBS + NR3 solve (gigpower_b200.sharding)
against the oracle on every rank, over NCCL."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from gigpower_b200.dss_reader import read_dss            # noqa: E402
from gigpower_b200.sharding import solve_sharded         # noqa: E402
from gigpower_b200.solution_fbs import SolutionFBS       # noqa: E402
from gigpower_b200.solution_nr3 import SolutionNR3       # noqa: E402
from oracle.gp_oracle import FBSOracle                   # noqa: E402
from oracle.nr3_tree import NR3TreeOracle                # noqa: E402

local = int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
fp = os.path.join(REPO, 'tests', 'golden', 'feeders', 'IEEE_13_Bus_allwye_noxfm_noreg.json')
S = 4099                                                # ragged across ranks
for cls, orc in ((SolutionFBS, FBSOracle), (SolutionNR3, NR3TreeOracle)):
    sol = cls(fp, device=local)
    o = orc(read_dss(fp))
    mult = np.random.Generator(np.random.PCG64(3)).uniform(0.5, 1.5, size=(S, len(o.c.load_names)))
    rows = np.ascontiguousarray(sol.load_weights() @ mult.T)
    out = solve_sharded(sol, rows)
    want = o.solve(o.c.spu_matrix(o.c.load_kw * mult, o.c.load_kvar * mult))
    assert np.array_equal(out['iterations'].cpu().numpy(), want['iterations'])
    t = sol.tables
    b, p = np.nonzero(t.vrow >= 0)
    Vmag = np.abs(want['V'])[:, b, p].T
    got = out['Vmag'].cpu().numpy()
    err = np.abs(got[t.vrow[b, p]] - Vmag).max()
    assert err <= 1e-9, err
    if dist.get_rank() == 0:
        print(f'{cls.__name__}: sharded over {dist.get_world_size()} GPUs, S={S}, max |dVmag| = {err:.2e}')
dist.destroy_process_group()

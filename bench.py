#!/usr/bin/env python
"""Benchmark of the batched power-flow hot path (BASELINE.json metric:
power-flow solves/sec).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                  [--workload c2|c3|c4fbs|c5] [--solver fbs|nr3]

A "step" is one pass of the hot path over one batch of synthetic scenarios:
flat start -> per-scenario convergence -> |V| and losses (-> all-gather of both
for N > 1, on a side stream so that it overlaps the next step's solve).  Steps
rotate over enough buffer sets that nothing is found in L2 from the last use.  At N=1 the default
workload is BASELINE config C2 (IEEE 13-bus FBS, 65,536 random load-multiplier
scenarios).  One process per GPU under torchrun; scenarios are sharded with no
data-path collective, the only exchange is the final all-gather of |V| and
losses: stores from the solve kernel's epilogue (or copy-engine pushes) into every rank's
peer-mapped buffer over NVLink, or NCCL with --gather nccl.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

FEEDER_DIR = os.path.join(REPO, 'tests', 'golden', 'feeders')
AUTHORED_DIR = os.path.join(REPO, 'gigpower_b200', 'feeders')
WORKLOADS = {
    # name: (feeder file, solver, scenarios per GPU, seed, multiplier range, description)
    'c2': ('IEEE_13_Bus_allwye_noxfm_noreg.json', 'fbs', 65536, 13, (0.5, 1.5),
           'C2: IEEE 13-bus FBS, 65,536 load-multiplier scenarios U(0.5,1.5), seed 13'),
    'c3': ('IEEE_37_Bus_allwye_noxfm_noreg.json', 'nr3', 16384, 37, (0.5, 1.5),
           'C3: IEEE 37-bus NR3, 16,384 load-multiplier scenarios U(0.5,1.5), seed 37'),
    'c4': ('ieee123_allwye_noxfm_noreg.dss', 'fbs', 140160, 123, 'der',
           'C4: IEEE 123-bus (reconstruction) FBS, 8,760 h x 128 DER scenarios = 1,121,280, 1/8 (140,160) per GPU'),
    'c4nr3': ('ieee123_allwye_noxfm_noreg.dss', 'nr3', 17520, 123, 'der',
              'C4: IEEE 123-bus (reconstruction) NR3, fixed 1/8 sub-sample of the 8,760 h x 128 DER set: 17,520 per GPU'),
    'c4ts': ('ieee123_allwye_noxfm_noreg.dss', 'fbs', 140160, 123, 'der',
             'C4 as a study: IEEE 123-bus FBS, 8,760 h x 128 DER scenarios, 140,160 per GPU; e2e = DerStudy.run '
             '(scenario rows synthesised on the device from the hour / DER tables, per-scenario summaries returned)'),
    'c5': ('synth2500_radial.dss', 'fbs', 131072, 2500, (0.3, 0.9),
           'C5: synthetic 2,541-bus radial FBS, 2^20 scenarios U(0.3,0.9), 1/8 (131,072) per GPU'),
    'nr3_37_big': ('IEEE_37_Bus_allwye_noxfm_noreg.json', 'nr3', 262144, 37, (0.5, 1.5),
                   'IEEE 37-bus NR3, 262,144 scenarios (throughput regime of the C3 kernel)'),
    'nr3_123_big': ('ieee123_allwye_noxfm_noreg.dss', 'nr3', 140160, 123, 'der',
                    'IEEE 123-bus NR3, 140,160 DER scenarios (throughput regime)'),
    'fbs37': ('IEEE_37_Bus_allwye_noxfm_noreg.json', 'fbs', 65536, 37, (0.5, 1.5),
              'IEEE 37-bus FBS, 65,536 scenarios'),
    'fbs34': ('IEEE_34_Bus_allwye_noxfm_noreg.json', 'fbs', 65536, 34, (0.5, 1.3),
              'IEEE 34-bus FBS, 65,536 scenarios'),
}


_JSON_FD = None


def protect_stdout():
    """stdout carries exactly ONE JSON line.  Libraries write to file descriptor 1 behind Python's back
    (NCCL prints its version there when NCCL_DEBUG is set in the environment), so fd 1 is pointed at
    stderr for the whole run and the line goes to a private duplicate of the original stdout."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + '\n').encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def peaks():
    p = os.path.join(REPO, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), 'measured'
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0}, 'fallback'


def ncu_traffic(workload, kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` on `workload`, from the
    committed `ncu --set full` captures (profiles/ncu_traffic.json); None when there is none."""
    p = os.path.join(REPO, 'profiles', 'ncu_traffic.json')
    if not os.path.exists(p):
        return None
    with open(p) as f:
        rec = json.load(f).get(f'{workload}:{kernel}')
    return float(rec['dram_bytes_per_launch']) if rec else None


def fbs_flops_per_iteration(t):
    """Algorithmic FP64 operations of one FBS iteration of one scenario:
    forward 8 per (bus phase, edge phase) complex multiply-subtract; backward per bus phase 29
    (|V|^2 3, sqrt 1, square 1, ZIP factor 4, spu scaling 2, capacitor 2, 1/|V|^2 4, conj(sV/V) 8,
    accumulate 2, finite test 2), 2 per child-edge phase, 8 per (rewritten parent phase, edge
    phase); root mismatch 18.  Divisions and square roots count as ONE operation each."""
    line = np.asarray(t.kind) == 0          # GP_EDGE_LINE; regulator steps scale V by gamma
    nb = (np.asarray(t.bus_vrow) >= 0).sum(axis=1)
    ne = (np.asarray(t.edge_irow) >= 0).sum(axis=1)
    npar = ((np.asarray(t.bus_vrow) >= 0) & (np.asarray(t.par_vrow) >= 0)).sum(axis=1)
    nchild = int((np.asarray(t.child_irow) >= 0).sum())
    return float((line * (8 * nb * ne + 29 * nb + 8 * npar * ne)).sum() + 2 * nchild + 18
                 + ((~line) * 4 * nb).sum())


class NvmlSampler:
    """SM clock and clock-event (throttle) reasons DURING the timed region, read through NVML from a
    thread every millisecond: the timed region of the default workload lasts tens of milliseconds,
    which `nvidia-smi -lms 100` (ClockSampler below, the fallback) samples once at best."""

    def __init__(self, gpu_index):
        import pynvml
        self.nv = pynvml
        pynvml.nvmlInit()
        idx = gpu_index
        vis = os.environ.get('CUDA_VISIBLE_DEVICES')
        if vis:
            try:
                idx = int(vis.split(',')[gpu_index])
            except Exception:
                pass
        self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        self.sm, self.mask, self.stop_flag = [], 0, False
        self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))

    def start(self):
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def _run(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.mask |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
            except Exception:
                pass
            time.sleep(0.001)

    def stop(self):
        nv = self.nv
        self.stop_flag = True
        self.thread.join(timeout=2)
        names = {'hw_slowdown': nv.nvmlClocksEventReasonHwSlowdown,
                 'hw_thermal_slowdown': nv.nvmlClocksEventReasonHwThermalSlowdown,
                 'sw_thermal_slowdown': nv.nvmlClocksEventReasonSwThermalSlowdown,
                 'sw_power_cap': nv.nvmlClocksEventReasonSwPowerCap,
                 'hw_power_brake_slowdown': nv.nvmlClocksEventReasonHwPowerBrakeSlowdown}
        return {'sm_mhz': float(np.median(self.sm)) if self.sm else None, 'sm_max_mhz': self.max_mhz,
                'samples': len(self.sm), 'reasons': sorted(n for n, bit in names.items() if self.mask & bit),
                'source': 'NVML, 1 ms period, timed region only'}


def make_sampler(gpu_index):
    try:
        return NvmlSampler(gpu_index)
    except Exception:
        return ClockSampler(gpu_index)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
         'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-lms', '100',
                 '-i', str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            try:
                sm.append(float(r[1])), mx.append(float(r[2]))
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith('active'):
                        reasons.add(n)
            except Exception:
                pass
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'samples': len(sm), 'reasons': sorted(reasons)}


def feeder_file(name):
    p = os.path.join(FEEDER_DIR, name)
    return p if os.path.exists(p) else os.path.join(AUTHORED_DIR, name)


def make_multipliers(n_loads, S, seed, spec, first=0):
    """Per-load multipliers [S, n_loads] of scenarios first..first+S of the workload."""
    if spec == 'der':
        raise ValueError('DER workloads are generated as spu rows')
    rng = np.random.Generator(np.random.PCG64(seed))
    return rng.uniform(spec[0], spec[1], size=(S, n_loads))


def make_der_rows(W0, S, first):
    """scenario index = d * 8760 + h.  Hour multiplier m_h (seed 123),
    PV at a random 10-40 % of the loaded bus-phases per DER scenario d (seed 1230 + d),
    penetration pen_d ~ U(0, 0.8), net spu = m_h * S_load - pen_d * pv(h) * P_load."""
    H, D = 8760, 128
    h_all = np.arange(H)
    rng = np.random.Generator(np.random.PCG64(123))
    m_h = np.clip(0.6 + 0.25 * np.sin(2 * np.pi * ((h_all % 24) - 9) / 24) + 0.1 * np.sin(2 * np.pi * h_all / H)
                  + rng.normal(0, 0.02, H), 0.3, 1.2)
    pv_h = np.maximum(0.0, np.sin(np.pi * ((h_all % 24) - 6) / 12))
    idx = (first + np.arange(S)) % (H * D)
    d, h = idx // H, idx % H
    nrow = W0.shape[0]
    mask = np.zeros((D, nrow))
    pen = np.zeros(D)
    for dd in np.unique(d):
        r = np.random.Generator(np.random.PCG64(1230 + int(dd)))
        frac = r.uniform(0.1, 0.4)
        mask[dd, r.permutation(nrow)[:max(1, int(round(frac * nrow)))]] = 1.0
        pen[dd] = r.uniform(0.0, 0.8)
    rows = W0[:, None] * m_h[h][None, :] - (W0.real[:, None] * mask[d].T) * (pen[d] * pv_h[h])[None, :]
    return np.ascontiguousarray(rows)


def make_scenarios(sol, S, seed, spec, first=0):
    """Synthetic scenario batch -> packed spu rows [n_srows, S] (+ multipliers when they exist)."""
    W = sol.load_weights()
    if spec == 'der':
        return make_der_rows(W.sum(axis=1), S, first), None
    mult = make_multipliers(sol.circuit.loads.num_elements, S, seed, spec)
    return np.ascontiguousarray(W @ mult.T), mult


def make_probe_inputs(workload, S=0):
    """(SolutionFBS, packed spu rows) of a bench workload, for the scripts under scripts/."""
    from gigpower_b200.solution_fbs import SolutionFBS
    feeder, solver, S0, seed, spec, _ = WORKLOADS[workload]
    sol = SolutionFBS(feeder_file(feeder))
    rows, _ = make_scenarios(sol, S or S0, seed, spec)
    return sol, rows


# --------------------------------------------------------------------------
# reference arm / CPU baseline: the oracle port on the host cores
# --------------------------------------------------------------------------
def _cpu_worker(args):
    feeder, solver, rows = args
    from gigpower_b200.dss_reader import read_dss
    from oracle.gp_oracle import FBSOracle, NR3Oracle
    from oracle.nr3_tree import NR3TreeOracle
    model = read_dss(feeder_file(feeder))
    if solver == 'fbs':
        o = FBSOracle(model)
    else:   # the literal dense restatement up to N ~ 500 unknowns, the tree elimination above that
        o = NR3Oracle(model) if len(model.bus_names) <= 45 else NR3TreeOracle(model)
    c = o.c
    # packed rows -> the reference's [S, nb, 3] spu matrices
    lmask = c.zip_mask * c.bus_pm
    b, p = np.nonzero(lmask)
    spu = np.zeros((rows.shape[1], c.nb, 3), dtype=complex)
    spu[:, b, p] = rows.T
    t0 = time.perf_counter()
    r = o.solve(spu)
    return time.perf_counter() - t0, int(r['iterations'].sum())


class CpuPool:
    """Persistent process pool running the oracle port on all host cores."""

    def __init__(self, cores=None):
        import multiprocessing as mp
        self.cores = cores or (os.cpu_count() or 1)
        # one BLAS thread per worker: the workers inherit the environment at spawn
        for var in ('OMP_NUM_THREADS', 'MKL_NUM_THREADS', 'OPENBLAS_NUM_THREADS', 'NUMEXPR_NUM_THREADS'):
            os.environ[var] = '1'
        self.pool = mp.get_context('spawn').Pool(self.cores)

    def step(self, feeder, solver, rows, n_per_proc):
        """Every core solves `n_per_proc` scenarios of the workload (columns of the packed
        spu rows); returns (scenarios solved, wall seconds)."""
        jobs = []
        S = rows.shape[1]
        for i in range(self.cores):
            lo = (i * n_per_proc) % max(S - n_per_proc + 1, 1)
            jobs.append((feeder, solver, np.ascontiguousarray(rows[:, lo:lo + n_per_proc])))
        t0 = time.perf_counter()
        self.pool.map(_cpu_worker, jobs)
        return sum(j[2].shape[1] for j in jobs), time.perf_counter() - t0

    def close(self):
        self.pool.close()
        self.pool.join()


def cpu_sample_size(solver, nb):
    """Scenarios per core per step, sized so that a step of the port takes ~0.2-10 s.  NR3 up to
    45 buses runs the literal dense restatement of the reference (H tensors of 0.4 GB and ~5 s per
    solve at 37 buses, like the reference itself): one scenario per core per step."""
    if solver == 'fbs':
        return max(16, min(1024, 15 * 1024 // nb))
    return 1 if nb <= 45 else max(2, min(64, 126 * 64 // nb))


def run_reference(args, wl):
    feeder, solver, S, seed, spec, desc = wl
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    from gigpower_b200.solution_fbs import SolutionFBS
    from gigpower_b200.dss_reader import read_dss
    from gigpower_b200.circuit import Circuit
    from gigpower_b200.flatten import flatten_fbs

    class _Host:        # host-only view of the feeder: load weights without touching the GPU
        def __init__(self, path):
            self.circuit = Circuit(read_dss(path))
            self.tables = flatten_fbs(self.circuit)
        load_weights = SolutionFBS.load_weights
    sol = _Host(feeder_file(feeder))
    nb = sol.circuit.buses.num_elements
    rows, mult = make_scenarios(sol, min(S, 16384), seed, spec)
    pool = CpuPool()
    cores = pool.cores
    n_per = cpu_sample_size(solver, nb)
    slow = solver == 'nr3' and n_per == 1           # dense NR3 restatement: seconds per solve
    for _ in range(1 if slow else max(args.warmup, 1)):
        pool.step(feeder, solver, rows, max(n_per // 8, 1))
    t0 = time.perf_counter()
    tot, done = 0, 0
    for _ in range(args.steps):
        n, _w = pool.step(feeder, solver, rows, n_per)
        tot, done = tot + n, done + 1
        if time.perf_counter() - t0 > (60.0 if slow else 150.0):      # keep the whole arm within a few minutes
            break
    wall = time.perf_counter() - t0
    pool.close()
    value = tot / wall
    line = {
        'impl': 'reference', 'metric': 'power-flow solves/sec', 'value': value, 'unit': 'solves/s',
        'n_gpus': args.gpus, 'steps': done, 'warmup': max(args.warmup, 1),
        'ms_per_step': 1e3 * wall / done, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': desc, 'solver': solver},
        'cpu_baseline': {'value': value, 'unit': 'solves/s', 'cores': cores, 'kind': 'port',
                         'sample': f'{n_per} scenarios of the workload per core per step, oracle/gp_oracle.py '
                                   f'(NumPy port of the reference, batched), persistent process pool'},
        'e2e': {'value': value, 'unit': 'solves/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    emit(line)


# --------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------
def run_ours(args, wl):
    import torch
    import torch.distributed as dist
    from gigpower_b200 import _cabi
    feeder, solver, S, seed, spec, desc = wl
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        # NCCL writes its version / debug lines to stdout by default; stdout carries the one JSON line
        os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')
        dist.init_process_group('nccl', device_id=dev)
    if solver == 'fbs':
        from gigpower_b200.solution_fbs import SolutionFBS
        sol = SolutionFBS(feeder_file(feeder), device=local)
    else:
        from gigpower_b200.solution_nr3 import SolutionNR3
        sol = SolutionNR3(feeder_file(feeder), device=local)
    rows, mult = make_scenarios(sol, S, seed + 1000 * rank, spec, first=rank * S)
    _cabi.check(_cabi.gp_set_option(b'jit_i_global', args.jit_ig))
    _cabi.check(_cabi.gp_set_option(b'fbs_share_rows', args.share_rows))
    if args.host_warps:
        _cabi.check(_cabi.gp_set_option(b'host_warps', args.host_warps))
    _cabi.check(_cabi.gp_set_option(b'nr3_kernel', args.nr3_kernel))
    sol._dev()
    _cabi.check(_cabi.gp_set_option(b'fbs_kernel', args.fbs_kernel))
    if args.prefetch is not None:
        _cabi.check(_cabi.gp_set_option(b'fbs_prefetch' if solver == 'fbs' else b'nr3_prefetch', args.prefetch))
    t = sol.tables
    h = sol._feeder.handle
    if solver == 'fbs' and (args.fbs_kernel == 5 or (args.fbs_kernel == 0 and _cabi.gp_fbs_jit_warps(h) >= sol.SPECIALISE_MIN_WARPS)):
        sol.specialise()        # what solve_batch does by itself for a batch this size
    active = int(_cabi.gp_fbs_active_kernel(h)) if solver == 'fbs' else 0
    ld = S
    spu_h = torch.from_numpy(rows).pin_memory()
    c128, f64, i32 = torch.complex128, torch.float64, torch.int32
    stream = torch.cuda.current_stream()
    side = torch.cuda.Stream(device=dev)          # the gather of step k overlaps the solve of step k+1
    maxiter = sol.maxiter
    tol = sol.tolerance if solver == 'fbs' else 0.0
    fused = active == 5     # the specialised FBS kernel writes |V| and losses itself
    pg = None               # peer-mapped gather buffers (world > 1, --gather peer)
    wbytes = int(_cabi.gp_nr3_workspace_bytes(h, ld)) if solver == 'nr3' else 0

    class BufSet:
        """Inputs, state and outputs of one step.  The timed loop rotates over enough sets that a
        step never finds its inputs or state in L2 from the previous use (total > 2.5 x L2)."""

        def __init__(self):
            self.spu = spu_h.to(dev)
            self.Vmag = torch.empty((t.n_vrows, ld), dtype=f64, device=dev)
            self.loss = torch.empty(S, dtype=c128, device=dev)
            self.iters = torch.empty(S, dtype=i32, device=dev)
            self.status = torch.empty(S, dtype=i32, device=dev)
            self.diff = torch.empty(S, dtype=f64, device=dev)
            if solver == 'fbs':
                self.V = torch.empty((t.n_vrows, ld), dtype=c128, device=dev)
                self.I = torch.empty((max(t.n_irows, 1), ld), dtype=c128, device=dev)
            else:
                self.X = torch.empty((t.n_vrows + t.n_irows, ld), dtype=c128, device=dev)
                self.work = torch.empty(wbytes, dtype=torch.uint8, device=dev)
            if world > 1 and args.gather == 'nccl':
                self.Vmag_all = torch.empty((world, t.n_vrows, ld), dtype=f64, device=dev)
                self.loss_all = torch.empty((world, S), dtype=c128, device=dev)
            self.gathered = None        # event: the gather that last read this set has finished

        def nbytes(self):
            return sum(v.numel() * v.element_size() for k, v in vars(self).items()
                       if hasattr(v, 'numel') and not k.endswith('_all'))

        def solve(self, slot=0):
            if solver == 'fbs' and fused and pg is not None:
                # one launch: solve + |V| + losses + all-gather (the epilogue stores each scenario's |V| rows
                # and loss into block `rank` of every rank's gather buffer over NVLink)
                ex = _cabi.make_extras(peers=pg.own_destination(slot) if args.gather == 'dma' else pg.destinations(slot))
                _cabi.check(_cabi.gp_fbs_solve_batch_ex(h, S, ld, self.spu.data_ptr(), ctypes.byref(ex),
                                                        self.V.data_ptr(), self.I.data_ptr(), self.iters.data_ptr(),
                                                        self.status.data_ptr(), self.diff.data_ptr(),
                                                        self.Vmag.data_ptr(), self.loss.data_ptr(), tol, maxiter,
                                                        stream.cuda_stream))
            elif solver == 'fbs' and fused:
                _cabi.check(_cabi.gp_fbs_solve_post_batch(h, S, ld, self.spu.data_ptr(), self.V.data_ptr(),
                                                          self.I.data_ptr(), self.iters.data_ptr(),
                                                          self.status.data_ptr(), self.diff.data_ptr(),
                                                          self.Vmag.data_ptr(), self.loss.data_ptr(), tol, maxiter,
                                                          stream.cuda_stream))
            elif solver == 'fbs':
                _cabi.check(_cabi.gp_fbs_solve_batch(h, S, ld, self.spu.data_ptr(), self.V.data_ptr(),
                                                     self.I.data_ptr(), self.iters.data_ptr(), self.status.data_ptr(),
                                                     self.diff.data_ptr(), tol, maxiter, stream.cuda_stream))
            else:
                _cabi.check(_cabi.gp_nr3_solve_batch(h, S, ld, self.spu.data_ptr(), self.X.data_ptr(),
                                                     self.iters.data_ptr(), self.status.data_ptr(),
                                                     self.diff.data_ptr(), self.work.data_ptr(), wbytes, maxiter,
                                                     stream.cuda_stream))

        def post(self):
            if solver == 'fbs' and not fused:
                _cabi.check(_cabi.gp_fbs_post(h, S, ld, self.spu.data_ptr(), self.V.data_ptr(), self.I.data_ptr(),
                                              None, self.Vmag.data_ptr(), None, None, self.loss.data_ptr(),
                                              stream.cuda_stream))
            elif solver == 'nr3':
                _cabi.check(_cabi.gp_nr3_map_output(h, S, ld, self.spu.data_ptr(), self.X.data_ptr(), None, None,
                                                    None, None, None, None, self.Vmag.data_ptr(),
                                                    self.loss.data_ptr(), stream.cuda_stream))

        def gather(self, slot=0):
            """The path's one exchange (|V| and losses of every rank), on the side stream: copy-engine
            pushes into every rank's peer-mapped buffer, or (--gather nccl) an NCCL all-gather.  The
            feeder-specialised FBS kernel has already done it from its epilogue."""
            if pg is not None and fused and args.gather != 'dma':
                return
            done = torch.cuda.Event()
            done.record(stream)
            side.wait_event(done)
            with torch.cuda.stream(side):
                if pg is not None and fused:
                    pg.push_own(slot, stream=side)          # the kernel wrote this rank's block of its own buffer
                    slot_pushed[slot] = torch.cuda.Event()
                    slot_pushed[slot].record(side)
                elif pg is not None:
                    pg.push(slot, self.Vmag, self.loss, stream=side)
                else:
                    dist.all_gather_into_tensor(self.Vmag_all, self.Vmag)
                    dist.all_gather_into_tensor(self.loss_all, self.loss)
                self.gathered = torch.cuda.Event()
                self.gathered.record(side)

    L2_BYTES = 126 << 20
    sets = [BufSet()]
    n_sets = max(1, min(8, -(-int(2.5 * L2_BYTES) // sets[0].nbytes())))
    sets += [BufSet() for _ in range(n_sets - 1)]
    n_slots = max(2, n_sets)
    slot_pushed = {}
    gather_note = None
    if world > 1 and args.gather in ('peer', 'dma'):
        from gigpower_b200.sharding import PeerGather
        try:
            pg = PeerGather(t.n_vrows, ld, slots=n_slots)
            ok = torch.ones(1, device=dev)
        except Exception as e:          # no CUDA IPC between the ranks (container policy): NCCL does the gather
            pg, ok = None, torch.zeros(1, device=dev)
            gather_note = f'peer-mapped gather unavailable ({e}); NCCL all-gather instead'
            print(gather_note, file=sys.stderr)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)       # all ranks or none
        if float(ok.item()) == 0.0:
            if pg is not None:
                pg.close()
            pg, args.gather = None, 'nccl'
            gather_note = gather_note or 'peer-mapped gather unavailable on another rank; NCCL all-gather instead'
            for b in sets:
                b.Vmag_all = torch.empty((world, t.n_vrows, ld), dtype=f64, device=dev)
                b.loss_all = torch.empty((world, S), dtype=c128, device=dev)

    def step(k):
        b = sets[k % n_sets]
        fused_gather = pg is not None and fused and args.gather != 'dma'
        if (k % n_slots) in slot_pushed:
            stream.wait_event(slot_pushed[k % n_slots])     # the DMA out of this slot's own block is done
        if b.gathered is not None and (fused or pg is None):
            stream.wait_event(b.gathered)       # the previous gather out of this set's buffers is done
        b.solve(k % n_slots)
        mid = torch.cuda.Event(enable_timing=True)
        mid.record(stream)
        if b.gathered is not None and not (fused or pg is None):
            stream.wait_event(b.gathered)       # |V| and losses of this set are still being pushed: only the
        b.post()                                # kernels that overwrite them wait, not the solve
        if world > 1 and not fused_gather:
            b.gather(k % n_slots)
        return mid

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for k in range(max(args.warmup, 3)):
        step(k)
    barrier()
    # ---- timed region: K steps back to back between two events; solve kernels timed individually ----
    sampler = make_sampler(local)
    if rank == 0:
        sampler.start()
    launches0 = _cabi.gp_launch_count()
    e0 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    e_begin, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    mids = []
    barrier()
    t_wall0 = time.perf_counter()
    e_begin.record(stream)
    for k in range(args.steps):
        e0[k].record(stream)
        mids.append(step(k))
    stream.wait_stream(side)                      # the last gathers belong to the timed steps
    e_end.record(stream)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop() if rank == 0 else None
    launches = _cabi.gp_launch_count() - launches0
    if pg is not None:
        # every rank's last step sits in block r of the slot on this rank
        slot = (args.steps - 1) % n_slots
        mine = sets[(args.steps - 1) % n_sets]
        got = pg.vmag_all(slot)
        assert torch.equal(got[rank], mine.Vmag), 'gathered |V| differs from the local result'
        assert torch.equal(pg.loss_all(slot)[rank], mine.loss), 'gathered losses differ from the local result'
        other = got[(rank + 1) % world]
        assert bool(torch.isfinite(other).all()) and float(other.min()) > 0.3, 'peer block not filled'
    ms_total = e_begin.elapsed_time(e_end)
    ms_solve = np.array([e0[k].elapsed_time(mids[k]) for k in range(args.steps)])
    tot_ms = torch.tensor([ms_total], dtype=f64, device=dev)
    if world > 1:
        dist.all_reduce(tot_ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(tot_ms.item()) / args.steps
    ms_steps = np.full(args.steps, ms_per_step)
    value = world * S / (ms_per_step * 1e-3)
    last = sets[(args.steps - 1) % n_sets]
    iters, status = last.iters, last.status
    spu_d = last.spu
    if solver == 'nr3':
        X = last.X

        def solve_kernel():
            last.solve()
    its = iters.cpu().numpy().astype(np.int64)
    assert (status.cpu().numpy() == 0).all(), 'bench scenarios must all converge'

    # ---- end to end: host buffers in, results out, copies inside the timed region -----------
    e2e_steps = max(3, min(args.steps, 10))
    e2e_extra = {}
    if args.workload == 'c4ts':
        # the study driver: hour / DER tables up (kilobytes), 36 bytes of summaries per scenario down
        from gigpower_b200.study import DerStudy
        H, D = 8760, 128
        h_all = np.arange(H)
        rng = np.random.Generator(np.random.PCG64(123))
        m_h = np.clip(0.6 + 0.25 * np.sin(2 * np.pi * ((h_all % 24) - 9) / 24) + 0.1 * np.sin(2 * np.pi * h_all / H)
                      + rng.normal(0, 0.02, H), 0.3, 1.2)
        pv_h = np.maximum(0.0, np.sin(np.pi * ((h_all % 24) - 6) / 12))
        mask, pen = np.zeros((D, t.n_srows)), np.zeros(D)
        for dd in range(D):
            r = np.random.Generator(np.random.PCG64(1230 + dd))
            frac = r.uniform(0.1, 0.4)
            mask[dd, r.permutation(t.n_srows)[:max(1, int(round(frac * t.n_srows)))]] = 1.0
            pen[dd] = r.uniform(0.0, 0.8)
        for b in sets:      # the study allocates its own chunk buffers
            del b.V, b.I
        torch.cuda.empty_cache()

        study = DerStudy(sol, m_h, pv_h, mask, pen)

        def e2e_step():
            study.upload(m_h, pv_h, mask, pen)                   # the step's inputs: the tables
            return study.run(first=rank * S, count=S, chunk=S)
        out = e2e_step()
        assert np.array_equal(out['iterations'], its), 'study and explicit batch disagree'
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        torch.cuda.synchronize()
        e2e_api = ('DerStudy.run: gp_der_rows + gp_fbs_solve_post_batch + gp_vmag_summary per chunk '
                   '(one chunk here); returns iterations, status, loss, min|V|, max|V|, violations')
        h2d_override = int((m_h.size + pv_h.size + mask.size + pen.size) * 8 + t.n_srows * 16)
        d2h = int(S * (4 + 4 + 16 + 8 + 8 + 4))
    elif solver == 'fbs':
        V_h = torch.empty((t.n_vrows, ld), dtype=c128).pin_memory()
        loss_h = torch.empty(S, dtype=c128).pin_memory()
        sol.solve_batch_host(spu_h, V_h, None, None, loss_h)      # warm-up (allocates staging)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            it_h, st_h, df_h = sol.solve_batch_host(spu_h, V_h, None, None, loss_h)
        torch.cuda.synchronize()
        e2e_api = 'gp_fbs_solve_batch_host (pinned host buffers; returns V, losses, iterations, status)'
        d2h = int(V_h.numel() * 16 + loss_h.numel() * 16 + S * (4 + 4 + 8))
        t_full = (time.perf_counter() - t0) / e2e_steps
        # the same call returning what the sharded path gathers (|V| and losses) instead of complex V
        Vm_h = torch.empty((t.n_vrows, ld), dtype=f64).pin_memory()
        sol.solve_batch_host(spu_h, None, None, Vm_h, loss_h)
        barrier()
        t1 = time.perf_counter()
        for _ in range(e2e_steps):
            sol.solve_batch_host(spu_h, None, None, Vm_h, loss_h)
        torch.cuda.synchronize()
        e2e_extra = {'returning_vmag_and_losses_only': {
            'seconds_per_step_this_rank': (time.perf_counter() - t1) / e2e_steps,
            'value_this_rank': S / ((time.perf_counter() - t1) / e2e_steps), 'unit': 'solves/s per GPU',
            'd2h_bytes_per_step': int(Vm_h.numel() * 8 + loss_h.numel() * 16 + S * (4 + 4 + 8))}}
        t0 = time.perf_counter() - t_full * e2e_steps      # the headline e2e stays the call that returns V
    else:
        X_h = torch.empty((t.n_vrows + t.n_irows, ld), dtype=c128).pin_memory()
        it_h = torch.empty(S, dtype=i32).pin_memory()

        def e2e_step():
            spu_d.copy_(spu_h, non_blocking=True)
            solve_kernel()
            X_h.copy_(X, non_blocking=True)
            it_h.copy_(iters, non_blocking=True)
            stream.synchronize()
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        e2e_api = 'pinned spu -> device, gp_nr3_solve_batch, X and iterations -> pinned host'
        d2h = int(X_h.numel() * 16 + S * 4)
    e2e_t = torch.tensor([(time.perf_counter() - t0) / e2e_steps], dtype=f64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = world * S / float(e2e_t.item())
    h2d = int(spu_h.numel() * 16) if args.workload != 'c4ts' else h2d_override

    if rank == 0:
        pk, pk_kind = peaks()
        solve_ms = float(np.mean(ms_solve))
        n_it = float(its.sum())
        kernel_names = {1: 'fbs_dfs_kernel', 2: 'fbs_phase_kernel', 3: 'fbs_solve_kernel', 4: 'fbs_onchip_kernel',
                        5: 'gp_fbs_jit_kernel'}
        variant = {0: None, 1: 'fused-dfs', 2: 'phase-parallel', 3: 'two-sweep streaming (state in HBM)',
                   4: 'on-chip two-sweep (state in shared memory), table-driven',
                   5: 'on-chip two-sweep (state in shared memory), feeder-specialised (NVRTC)'}[active]
        if solver == 'fbs' and active in (4, 5):
            # on-chip kernel: per iteration only the spu rows are read (L2-resident after the first
            # sweep); V and I leave the SM once.  The kernel is bound by the FP64 pipe, so the
            # headline roofline is FP64 (algorithmic flops of the sweeps);
            # the HBM fraction of the same launch is reported next to it.
            per_it, once = 16 * t.n_srows, 16 * (t.n_vrows + t.n_irows) + 16
            kname = kernel_names[active]
        elif solver == 'fbs':
            # per sweep read I, write V | read V, read spu, write I, write V
            per_it = 16 * (2 * t.n_irows + 3 * t.n_vrows + t.n_srows)
            once = 16 * (t.n_srows + t.n_vrows + t.n_irows)
            kname = kernel_names[active]
        else:
            # per Newton iteration: X read+write, spu read, dV write+read, fronts (Y, a, r1) write+read
            nph = (t.bus_vrow >= 0).sum(axis=1)
            fronts = int((8 * ((2 * nph) ** 2 + 4 * nph)).sum())
            per_it = 2 * 16 * (t.n_vrows + t.n_irows) + 16 * t.n_srows + 2 * 16 * t.n_vrows + 2 * fronts
            once = 16 * (t.n_vrows + t.n_irows)
            phase_lanes = args.nr3_kernel == 2 or (args.nr3_kernel == 0 and bool((np.asarray(t.kind) == 0).all())
                                                   and S <= 32768)      # the library's automatic choice
            kname = 'nr3_phase_kernel' if phase_lanes else 'nr3_solve_kernel'
        alg_bytes = n_it * per_it + S * once            # solve kernel, per launch
        achieved = alg_bytes / (solve_ms * 1e-3) / 1e9
        traffic = ncu_traffic(args.workload, kname)
        roofline = {'bound': 'hbm', 'kernel': kname, 'achieved': achieved,
                    'peak': pk['hbm_gbs'], 'peak_kind': pk_kind, 'unit': 'GB/s',
                    'frac': achieved / pk['hbm_gbs'], 'traffic': traffic,
                    'algorithmic_bytes_per_launch': alg_bytes, 'kernel_ms': solve_ms,
                    'kernel_share_of_step': solve_ms / float(np.mean(ms_steps))}
        if solver == 'fbs' and active in (4, 5):
            tf = ctypes.c_double(0.0)
            _cabi.check(_cabi.gp_measure_fp64_peak(local, ctypes.byref(tf)))
            flops = n_it * fbs_flops_per_iteration(t)
            ach_tf = flops / (solve_ms * 1e-3) / 1e12
            roofline = {'bound': 'fp64', 'kernel': kname, 'achieved': ach_tf, 'peak': tf.value,
                        'peak_kind': 'measured live (gp_measure_fp64_peak: DFMA chains, 2 flop each)',
                        'unit': 'TFLOP/s', 'frac': ach_tf / tf.value, 'traffic': traffic,
                        'algorithmic_flops_per_launch': flops, 'kernel_ms': solve_ms,
                        'kernel_share_of_step': solve_ms / float(np.mean(ms_steps)),
                        'hbm': {'achieved': achieved, 'peak': pk['hbm_gbs'], 'peak_kind': pk_kind, 'unit': 'GB/s',
                                'frac': achieved / pk['hbm_gbs'], 'algorithmic_bytes_per_launch': alg_bytes,
                                'streaming_formulation_equivalent_gbs':
                                    (n_it * 16 * (2 * t.n_irows + 3 * t.n_vrows + t.n_srows)
                                     + S * 16 * (t.n_srows + t.n_vrows + t.n_irows)) / (solve_ms * 1e-3) / 1e9}}
        line = {
            'metric': 'power-flow solves/sec', 'value': value, 'unit': 'solves/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': ms_per_step,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
            'data': 'synthetic',
            'config': {'workload': desc, 'solver': solver, 'scenarios_per_gpu': S,
                       'fbs_kernel': variant,
                       'rows': {'P_b': t.n_vrows, 'P_l': t.n_irows, 'P_L': t.n_srows},
                       'mean_iterations': float(its.mean()), 'max_iterations': int(its.max()),
                       'l2': f'no reuse across steps: {n_sets} rotating buffer set(s) of {sets[0].nbytes() >> 20} MiB '
                             f'(inputs + state + outputs) against a 126 MiB L2',
                       'timing': 'CUDA events around the K back-to-back steps on the launch stream (gather of step k '
                                 'on a side stream, overlapping the solve of step k+1), max over ranks; solve kernels '
                                 'timed individually for the roofline',
                       'parallelism': f'scenario-sharded x{world}, ' + (
                           'single GPU: no exchange' if world == 1 else
                           'all-gather of |V| and losses by NCCL on a side stream' if pg is None else
                           'all-gather of |V| and losses: the solve kernel writes this rank\'s block of its own gather '
                           'buffer, copy engines push it to every peer on a side stream' if fused and args.gather == 'dma' else
                           'all-gather of |V| and losses fused into the solve kernel (epilogue stores over NVLink into '
                           'every rank\'s peer-mapped buffer)' if fused else
                           'all-gather of |V| and losses by copy-engine pushes into every rank\'s peer-mapped buffer '
                           'on a side stream')},
            'e2e': {'value': e2e_value, 'unit': 'solves/s', 'h2d_bytes_per_step': h2d,
                    'd2h_bytes_per_step': d2h, 'steps': e2e_steps,
                    'api': e2e_api, **e2e_extra},
            'gpu_launches': int(launches),
            'roofline': roofline,
            'clocks': clocks, 'wall_s_timed_region': t_wall,
        }
        if world == 1 and not args.no_cpu:
            pool = CpuPool()
            n_per = cpu_sample_size(solver, sol.circuit.buses.num_elements)
            if not (solver == 'nr3' and n_per == 1):
                pool.step(feeder, solver, rows, max(n_per // 8, 1))      # warm the workers
            n, wall, reps = 0, 0.0, 0
            while wall < 10.0 and reps < 200:
                dn, dw = pool.step(feeder, solver, rows, n_per)
                n, wall, reps = n + dn, wall + dw, reps + 1
            pool.close()
            line['cpu_baseline'] = {
                'value': n / wall, 'unit': 'solves/s', 'cores': pool.cores, 'kind': 'port',
                'sample': f'{n} scenarios of the workload ({n_per} per core per call, {reps} calls, '
                          f'{wall:.1f} s), oracle/gp_oracle.py batched NumPy port on a process pool'}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='c2', choices=sorted(WORKLOADS))
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    ap.add_argument('--gather', default='peer', choices=['peer', 'dma', 'nccl'],
                    help='N > 1: the all-gather of |V| and losses through peer-mapped memory (fused into the '
                         'specialised FBS kernel, copy-engine pushes otherwise) or by NCCL')
    ap.add_argument('--jit-ig', type=int, default=-1, choices=[-1, 0, 1],
                    help='I rows of the specialised FBS kernel in global memory: -1 automatic, 0 no, 1 yes')
    ap.add_argument('--host-warps', type=int, default=0,
                    help='warps per SM of the one-launch host path (e2e); 0 = the library default')
    ap.add_argument('--share-rows', type=int, default=1, choices=[0, 1],
                    help='streaming FBS kernel: 1 = I rows that repeat a child\'s current or stay zero make no HBM round '
                         'trips (default), 0 = every row travels')
    ap.add_argument('--nr3-kernel', type=int, default=0, choices=[0, 1, 2],
                    help='0 automatic, 1 one thread per scenario, 2 three phase lanes per scenario')
    ap.add_argument('--prefetch', type=int, default=None, help='L2 prefetch distance of the streaming FBS kernel / the NR3 kernel')
    ap.add_argument('--fbs-kernel', type=int, default=0, choices=[0, 1, 2, 3, 4, 5],
                    help='0 = automatic (on-chip when the feeder fits in shared memory, else two-sweep streaming), '
                         '1 = fused depth-first, 2 = phase-parallel lanes, 3 = two-sweep streaming, 4 = on-chip table-driven, '
                         '5 = on-chip feeder-specialised')
    args = ap.parse_args()
    protect_stdout()
    wl = WORKLOADS[args.workload]
    if args.impl == 'reference':
        run_reference(args, wl)
    else:
        run_ours(args, wl)


if __name__ == '__main__':
    main()

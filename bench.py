#!/usr/bin/env python
"""Benchmark of the batched power-flow hot path (BASELINE.json metric:
power-flow solves/sec, FBS & NR3, IEEE-123, at 1/2/4/8 B200).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                  [--workload c4|c2|c3|c4nr3|c5|...] [--sub c2,c3,c4nr3,c5|none]

A "step" is one pass of the hot path over one batch of synthetic scenarios:
flat start -> per-scenario convergence -> |V| and losses (-> all-gather of both
for N > 1, overlapping the next step's solve).  Steps rotate over enough buffer
sets that nothing is found in L2 from the last use.

The headline line is BASELINE config C4 at every N: IEEE 123-bus FBS, the
8,760 h x 128 DER scenario set, 140,160 scenarios per GPU (weak scaling; 8 GPUs
hold the whole set).  The K timed steps are repeated until the timed region
lasts >= 0.6 s, so the value is sustained and the clock record has hundreds of
samples.  The other BASELINE configs (C2 13-bus FBS, C3 37-bus NR3, C4 123-bus
NR3, C5 2,541-bus FBS) run after it, each with its own timed loop, and are
reported in the same JSON line under "sub_results".

One process per GPU under torchrun; scenarios are sharded with no data-path
collective, the only exchange is the final all-gather of |V| and losses: stores
from the solve kernel's epilogue (or copy-engine pushes) into every rank's
peer-mapped buffer over NVLink, or NCCL with --gather nccl.  Prints ONE JSON
line on rank 0.

--impl reference: the reference's own CPU implementation on all host cores: the
UNMODIFIED gigpower package from baseline/_ref (copied there by
__graft_entry__.build()) behind the opendssdirect stand-in when it is present,
else the oracle port; the port's figure is reported next to it either way.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

FEEDER_DIR = os.path.join(REPO, 'gigpower_b200', 'feeders')
GOLDEN_FEEDER_DIR = os.path.join(REPO, 'tests', 'golden', 'feeders')
REF_DIR = os.path.join(REPO, 'baseline', '_ref')
STANDIN_DIR = os.path.join(REPO, 'oracle', 'standin')
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
DEFAULT_WORKLOAD = 'c4'
DEFAULT_SUBS = ('c2', 'c3', 'c4nr3', 'c5')
MIN_TIMED_SECONDS = 0.6

_JSON_FD = None


def log(*a):
    print('[bench]', *a, file=sys.stderr, flush=True)


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
            return json.load(f), 'measured (MEASURED_PEAKS.json)'
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0}, 'fallback (B200_PROFILING.md)'


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
    thread every millisecond."""

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
    return p if os.path.exists(p) else os.path.join(GOLDEN_FEEDER_DIR, name)


def make_multipliers(n_loads, S, seed, spec, first=0):
    """Per-load multipliers [S, n_loads] of scenarios first..first+S of the workload."""
    if spec == 'der':
        raise ValueError('DER workloads are generated as spu rows')
    rng = np.random.Generator(np.random.PCG64(seed))
    return rng.uniform(spec[0], spec[1], size=(S, n_loads))


def der_tables(nrow):
    """hour multiplier m_h (seed 123), pv(h), and per DER scenario d (seed 1230 + d) the
    bus-phases that carry PV (a random 10-40 % of the loaded ones) and the penetration pen_d ~ U(0, 0.8)."""
    H, D = 8760, 128
    h_all = np.arange(H)
    rng = np.random.Generator(np.random.PCG64(123))
    m_h = np.clip(0.6 + 0.25 * np.sin(2 * np.pi * ((h_all % 24) - 9) / 24) + 0.1 * np.sin(2 * np.pi * h_all / H)
                  + rng.normal(0, 0.02, H), 0.3, 1.2)
    pv_h = np.maximum(0.0, np.sin(np.pi * ((h_all % 24) - 6) / 12))
    mask, pen = np.zeros((D, nrow)), np.zeros(D)
    for dd in range(D):
        r = np.random.Generator(np.random.PCG64(1230 + dd))
        frac = r.uniform(0.1, 0.4)
        mask[dd, r.permutation(nrow)[:max(1, int(round(frac * nrow)))]] = 1.0
        pen[dd] = r.uniform(0.0, 0.8)
    return m_h, pv_h, mask, pen


def make_der_rows(W0, S, first):
    """scenario index = d * 8760 + h; net spu = m_h * S_load - pen_d * pv(h) * P_load."""
    H, D = 8760, 128
    m_h, pv_h, mask, pen = der_tables(W0.shape[0])
    idx = (first + np.arange(S)) % (H * D)
    d, h = idx // H, idx % H
    rows = W0[:, None] * m_h[h][None, :] - (W0.real[:, None] * mask[d].T) * (pen[d] * pv_h[h])[None, :]
    return np.ascontiguousarray(rows)


def make_scenarios(sol, S, seed, spec, first=0, device=None):
    """Synthetic scenario batch -> packed spu rows [n_srows, S] (+ multipliers when they exist).
    With `device` (a torch CUDA device) a large multiplier batch is expanded on the GPU (W @ mult.T is a
    dense product of some 10^11 operations for the 2,541-bus feeder: input synthesis, not the path) and
    the rows are returned as a CUDA tensor."""
    W = sol.load_weights()
    if spec == 'der':
        return make_der_rows(W.sum(axis=1), S, first), None
    mult = make_multipliers(sol.circuit.loads.num_elements, S, seed, spec)
    if device is not None and W.shape[0] * W.shape[1] * S > 4e9:
        import torch
        m = torch.from_numpy(mult).to(device)
        re = torch.from_numpy(np.ascontiguousarray(W.real)).to(device) @ m.T
        im = torch.from_numpy(np.ascontiguousarray(W.imag)).to(device) @ m.T
        rows = torch.complex(re, im).contiguous()
        del m, re, im
        return rows, None
    return np.ascontiguousarray(W @ mult.T), mult


def make_probe_inputs(workload, S=0):
    """(SolutionFBS, packed spu rows) of a bench workload, for the scripts under scripts/."""
    from gigpower_b200.solution_fbs import SolutionFBS
    feeder, solver, S0, seed, spec, _ = WORKLOADS[workload]
    sol = SolutionFBS(feeder_file(feeder))
    rows, _ = make_scenarios(sol, S or S0, seed, spec)
    return sol, rows


class HostFeeder:
    """Host-only view of a feeder (parse + flatten): what the CPU arms need to synthesise the workload's
    scenario rows.  Imports neither the CUDA library nor torch."""

    def __init__(self, path, solver='fbs'):
        from gigpower_b200.circuit import Circuit
        from gigpower_b200.dss_reader import read_dss
        from gigpower_b200 import flatten
        self.circuit = Circuit(read_dss(path))
        self.tables = flatten.flatten_fbs(self.circuit) if solver == 'fbs' else flatten.flatten_nr3(self.circuit)
        self._lw = flatten.load_weights

    def load_weights(self):
        return self._lw(self.circuit, self.tables.srow, self.tables.n_srows)


# --------------------------------------------------------------------------
# CPU arms: (1) the oracle port, (2) the unmodified reference from baseline/_ref
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


_REF = {}


def reference_available():
    return os.path.exists(os.path.join(REF_DIR, 'gigpower', 'solution_fbs.py'))


def _ref_init(feeder, solver):
    """Pool initialiser: one UNMODIFIED reference Solution object per worker,
    imported from baseline/_ref behind the opendssdirect stand-in."""
    for var in ('OMP_NUM_THREADS', 'MKL_NUM_THREADS', 'OPENBLAS_NUM_THREADS'):
        os.environ[var] = '1'
    for p in (REF_DIR, STANDIN_DIR):
        if p not in sys.path:
            sys.path.insert(0, p)
    import warnings
    warnings.filterwarnings('ignore')
    if solver == 'fbs':
        from gigpower.solution_fbs import SolutionFBS as Ref
    else:
        from gigpower.solution_nr3 import SolutionNR3 as Ref
    s = Ref(feeder_file(feeder))
    _REF['s'], _REF['solver'] = s, solver
    _REF['base'] = [(n, s.circuit.loads.get_element(n).kW, s.circuit.loads.get_element(n).kvar)
                    for n in s.circuit.loads.all_names()]


def _ref_worker(args):
    """K scenarios through the reference's own solve() (solution_fbs.py:76-136 / solution_nr3.py:592-632),
    state reset between scenarios as BASELINE.md section 3 describes.  FBS: spu_k [K, nb, 3] complex, the
    scenario's Solution.spu; NR3: mult_k [K, n_loads], set through the reference's own set_kW / set_kvar."""
    spu_k, mult_k = args
    s, solver = _REF['s'], _REF['solver']
    t0 = time.perf_counter()
    its = 0
    if solver == 'fbs':
        for spu in spu_k:
            s.V[:] = 0
            s.I[:] = 0
            s.iterations = 0
            s.Vtest = np.zeros(3, dtype=complex)
            s.spu = spu.copy()
            s.solve()
            its += int(s.iterations)
    else:
        for m in mult_k:
            for (n, kw, kvar), f in zip(_REF['base'], m):
                s.circuit.set_kW(n, kw * f)
                s.circuit.set_kvar(n, kvar * f)
            s._init_KCL_matrices()
            s._init_XNR()
            s.solve()
    return time.perf_counter() - t0, its


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


class RefPool:
    """Persistent process pool: one unmodified reference Solution per core (baseline/_ref)."""

    def __init__(self, feeder, solver, srow, nb, cores=None):
        import multiprocessing as mp
        self.cores = cores or (os.cpu_count() or 1)
        self.srow, self.nb = np.asarray(srow), nb
        t0 = time.perf_counter()
        self.pool = mp.get_context('spawn').Pool(self.cores, initializer=_ref_init, initargs=(feeder, solver))
        self.pool.map(_ref_ready, range(self.cores))
        self.setup_s = time.perf_counter() - t0

    def step(self, rows, k_per_proc, mult=None):
        """Every core solves k_per_proc scenarios: columns of the packed spu rows (FBS) or rows of the
        per-load multipliers (NR3, whose loads enter through the circuit's setters)."""
        S = rows.shape[1]
        b, p = np.nonzero(self.srow >= 0)
        jobs = []
        for i in range(self.cores):
            lo = (i * k_per_proc) % max(S - k_per_proc + 1, 1)
            if mult is not None:
                jobs.append((None, np.ascontiguousarray(mult[lo:lo + k_per_proc])))
                continue
            piece = rows[:, lo:lo + k_per_proc]
            spu = np.zeros((piece.shape[1], self.nb, 3), dtype=complex)
            spu[:, b, p] = piece[self.srow[b, p]].T
            jobs.append((spu, None))
        t0 = time.perf_counter()
        out = self.pool.map(_ref_worker, jobs)
        return sum(len(j[0] if j[0] is not None else j[1]) for j in jobs), time.perf_counter() - t0, sum(o[1] for o in out)

    def close(self):
        self.pool.close()
        self.pool.join()


def _ref_ready(_):
    return 's' in _REF


def cpu_sample_size(solver, nb):
    """Scenarios per core per step of the PORT, sized so that a step takes ~0.2-10 s.  NR3 up to
    45 buses runs the literal dense restatement of the reference (H tensors of 0.4 GB and ~5 s per
    solve at 37 buses, like the reference itself): one scenario per core per step."""
    if solver == 'fbs':
        return max(16, min(1024, 15 * 1024 // nb))
    return 1 if nb <= 45 else max(2, min(64, 126 * 64 // nb))


def ref_sample_size(solver, nb):
    """Scenarios per core per step of the UNMODIFIED reference (FBS ~0.8 ms per bus per solve)."""
    if solver == 'fbs':
        return max(2, min(64, 1000 // nb))
    return 1


def time_port(feeder, solver, rows, nb, seconds, max_calls=200, warm=True):
    pool = CpuPool()
    n_per = cpu_sample_size(solver, nb)
    if warm and not (solver == 'nr3' and n_per == 1):
        pool.step(feeder, solver, rows, max(n_per // 8, 1))      # warm the workers
    n, wall, reps = 0, 0.0, 0
    while wall < seconds and reps < max_calls:
        dn, dw = pool.step(feeder, solver, rows, n_per)
        n, wall, reps = n + dn, wall + dw, reps + 1
    pool.close()
    return {'value': n / wall, 'unit': 'solves/s', 'cores': pool.cores, 'kind': 'port',
            'sample': f'{n} scenarios of the workload ({n_per} per core per call, {reps} calls, {wall:.1f} s), '
                      f'oracle/gp_oracle.py batched NumPy port on a process pool'}


def time_reference_python(feeder, solver, rows, host, seconds, max_calls=50, mult=None):
    """The unmodified reference's own solve() on all host cores; None when baseline/_ref is absent or the
    workload is one the reference cannot hold (123-bus NR3: 15 GB of dense H per process)."""
    nb = host.circuit.buses.num_elements
    if not reference_available() or (solver == 'nr3' and (nb > 45 or mult is None)):
        return None
    pool = RefPool(feeder, solver, host.tables.srow, nb)
    k = ref_sample_size(solver, nb)
    mm = mult if solver == 'nr3' else None
    pool.step(rows, 1, mm)                                       # warm-up: one solve per core
    n, wall, reps, its = 0, 0.0, 0, 0
    while wall < seconds and reps < max_calls:
        dn, dw, di = pool.step(rows, k, mm)
        n, wall, reps, its = n + dn, wall + dw, reps + 1, its + di
    pool.close()
    return {'value': n / wall, 'unit': 'solves/s', 'cores': pool.cores, 'kind': 'reference', 'K': k,
            'mean_iterations': its / n if (n and solver == 'fbs') else None,
            'sample': f'{n} scenarios of the workload ({k} per core per call, {reps} calls, {wall:.1f} s) through the '
                      f'UNMODIFIED gigpower {"SolutionFBS" if solver == "fbs" else "SolutionNR3"}.solve() '
                      f'(baseline/_ref behind oracle/standin/opendssdirect), one object per core, state reset per scenario '
                      f'(BASELINE.md section 3); pool + constructors {pool.setup_s:.1f} s, not timed'}


def run_reference(args, wl):
    feeder, solver, S, seed, spec, desc = wl
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    host = HostFeeder(feeder_file(feeder), solver)
    nb = host.circuit.buses.num_elements
    rows, mult = make_scenarios(host, min(S, 16384), seed, spec)
    use_ref = reference_available() and not (solver == 'nr3' and (nb > 45 or mult is None))
    mm = mult if solver == 'nr3' else None
    warm = max(args.warmup, 1)
    if use_ref:
        pool = RefPool(feeder, solver, host.tables.srow, nb)
        cores, k = pool.cores, ref_sample_size(solver, nb)
        for _ in range(min(warm, 2)):
            pool.step(rows, 1, mm)
        t0 = time.perf_counter()
        tot, done = 0, 0
        for _ in range(args.steps):
            n, _w, _i = pool.step(rows, k, mm)
            tot, done = tot + n, done + 1
            if time.perf_counter() - t0 > 120.0:        # keep the whole arm within a few minutes
                break
        wall = time.perf_counter() - t0
        pool.close()
        kind = 'reference'
        sample = (f'{k} scenarios of the workload per core per step through the UNMODIFIED gigpower solve() '
                  f'(baseline/_ref behind oracle/standin/opendssdirect), one object per core, state reset per scenario')
        port = time_port(feeder, solver, rows, nb, 8.0)
    else:
        pool = CpuPool()
        cores = pool.cores
        n_per = cpu_sample_size(solver, nb)
        slow = solver == 'nr3' and n_per == 1           # dense NR3 restatement: seconds per solve
        for _ in range(1 if slow else warm):
            pool.step(feeder, solver, rows, max(n_per // 8, 1))
        t0 = time.perf_counter()
        tot, done = 0, 0
        for _ in range(args.steps):
            n, _w = pool.step(feeder, solver, rows, n_per)
            tot, done = tot + n, done + 1
            if time.perf_counter() - t0 > (60.0 if slow else 150.0):
                break
        wall = time.perf_counter() - t0
        pool.close()
        kind = 'port'
        sample = (f'{n_per} scenarios of the workload per core per step, oracle/gp_oracle.py (NumPy port of the '
                  f'reference, batched), persistent process pool')
        port = None
    value = tot / wall
    line = {
        'impl': 'reference', 'metric': 'power-flow solves/sec', 'value': value, 'unit': 'solves/s',
        'n_gpus': args.gpus, 'steps': done, 'warmup': warm,
        'ms_per_step': 1e3 * wall / done, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': desc, 'solver': solver},
        'cpu_baseline': {'value': value, 'unit': 'solves/s', 'cores': cores, 'kind': kind, 'sample': sample},
        'e2e': {'value': value, 'unit': 'solves/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    if port is not None:
        line['cpu_baseline']['port'] = port
    emit(line)


# --------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------
class Ctx:
    pass


def run_workload(ctx, args, name, steps, main):
    """One workload on this rank's GPU: K timed steps (repeated until the timed region lasts
    MIN_TIMED_SECONDS for the main workload), roofline of the solve kernel, e2e through host buffers
    (main workload only).  Returns the result dict on rank 0, None elsewhere."""
    import torch
    import torch.distributed as dist
    from gigpower_b200 import _cabi
    feeder, solver, S, seed, spec, desc = WORKLOADS[name]
    world, rank, local, dev = ctx.world, ctx.rank, ctx.local, ctx.dev
    t_setup = time.perf_counter()
    if solver == 'fbs':
        from gigpower_b200.solution_fbs import SolutionFBS
        sol = SolutionFBS(feeder_file(feeder), device=local)
    else:
        from gigpower_b200.solution_nr3 import SolutionNR3
        sol = SolutionNR3(feeder_file(feeder), device=local)
    rows, mult = make_scenarios(sol, S, seed + 1000 * rank, spec, first=rank * S, device=dev)
    _cabi.check(_cabi.gp_set_option(b'jit_i_global', args.jit_ig))
    _cabi.check(_cabi.gp_set_option(b'jit_dfs', args.jit_dfs))
    _cabi.check(_cabi.gp_set_option(b'fbs_share_rows', args.share_rows))
    if args.host_warps:
        _cabi.check(_cabi.gp_set_option(b'host_warps', args.host_warps))
    _cabi.check(_cabi.gp_set_option(b'host_chunks', args.host_chunks))
    if args.push_streams:
        _cabi.check(_cabi.gp_set_option(b'peer_push_streams', args.push_streams))
    _cabi.check(_cabi.gp_set_option(b'nr3_kernel', args.nr3_kernel))
    _cabi.check(_cabi.gp_set_option(b'nr3_ring', args.nr3_ring))
    _cabi.check(_cabi.gp_set_option(b'fbs_order', args.fbs_order))
    _cabi.check(_cabi.gp_set_option(b'fbs_walk_blocks', args.walk_blocks))
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
    on_device = hasattr(rows, 'data_ptr')
    spu_src = rows if on_device else torch.from_numpy(rows).pin_memory()
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
            self.spu = spu_src.to(dev) if not on_device else (spu_src if not sets else spu_src.clone())
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
    sets = []
    sets.append(BufSet())
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
            log(gather_note)
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

    n_warm = max(args.warmup, 3)
    for k in range(n_warm):
        step(k)
    barrier()
    # one timed pass of the warm-up length sizes the number of repeats of the K steps
    w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0.record(stream)
    for k in range(n_warm):
        step(k)
    stream.wait_stream(side)
    w1.record(stream)
    barrier()
    est_ms = torch.tensor([w0.elapsed_time(w1) / n_warm], dtype=f64, device=dev)
    if world > 1:
        dist.all_reduce(est_ms, op=dist.ReduceOp.MAX)
    repeats = 1
    if main:
        repeats = max(1, min(2000, int(math.ceil(MIN_TIMED_SECONDS * 1e3 / (float(est_ms.item()) * steps)))))
    total_steps = steps * repeats
    if rank == 0:
        log(f'{name}: set-up {time.perf_counter() - t_setup:.1f} s, ~{float(est_ms.item()):.3f} ms/step, '
            f'{steps} steps x {repeats} repeat(s), {n_sets} buffer set(s)')
    # ---- timed region: K steps (x repeats) back to back between two events; solve kernels timed individually ----
    sampler = make_sampler(local)
    if rank == 0:
        sampler.start()
    launches0 = _cabi.gp_launch_count()
    e0 = [torch.cuda.Event(enable_timing=True) for _ in range(total_steps)]
    e_begin, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    mids = []
    barrier()
    t_wall0 = time.perf_counter()
    e_begin.record(stream)
    for k in range(total_steps):
        e0[k].record(stream)
        mids.append(step(k))
    stream.wait_stream(side)                      # the last gathers belong to the timed steps
    e_end.record(stream)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop() if rank == 0 else None
    launches = (_cabi.gp_launch_count() - launches0) // repeats
    if pg is not None:
        # every rank's last step sits in block r of the slot on this rank
        slot = (total_steps - 1) % n_slots
        mine = sets[(total_steps - 1) % n_sets]
        got = pg.vmag_all(slot)
        assert torch.equal(got[rank], mine.Vmag), 'gathered |V| differs from the local result'
        assert torch.equal(pg.loss_all(slot)[rank], mine.loss), 'gathered losses differ from the local result'
        other = got[(rank + 1) % world]
        assert bool(torch.isfinite(other).all()) and float(other.min()) > 0.3, 'peer block not filled'
    ms_total = e_begin.elapsed_time(e_end)
    ms_solve = np.array([e0[k].elapsed_time(mids[k]) for k in range(total_steps)])
    tot_ms = torch.tensor([ms_total], dtype=f64, device=dev)
    if world > 1:
        dist.all_reduce(tot_ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(tot_ms.item()) / total_steps
    value = world * S / (ms_per_step * 1e-3)
    last = sets[(total_steps - 1) % n_sets]
    iters, status = last.iters, last.status
    spu_d = last.spu
    its = iters.cpu().numpy().astype(np.int64)
    assert (status.cpu().numpy() == 0).all(), 'bench scenarios must all converge'

    # ---- end to end: host buffers in, results out, copies inside the timed region -----------
    e2e = None
    if main or args.sub_e2e:
        e2e_steps = max(3, min(steps, 10))
        e2e_extra = {}

        def sustained(step):
            """Steps of the e2e region: the headline workload's covers MIN_TIMED_SECONDS like the device-timed one (a
            0.2 s window on a shared host has shown 25 % run-to-run spread); every rank takes the same count."""
            if not main:
                return e2e_steps
            t = time.perf_counter()
            step()
            torch.cuda.synchronize()
            n = torch.tensor([MIN_TIMED_SECONDS / max(time.perf_counter() - t, 1e-6)], dtype=f64, device=dev)
            if world > 1:
                dist.all_reduce(n, op=dist.ReduceOp.MAX)
            return int(min(max(e2e_steps, math.ceil(float(n.item()))), 60))
        spu_h = spu_src if not on_device else spu_src.cpu().pin_memory()
        h2d_override = None
        if name == 'c4ts':
            # the study driver: hour / DER tables up (kilobytes), 36 bytes of summaries per scenario down
            from gigpower_b200.study import DerStudy
            m_h, pv_h, mask, pen = der_tables(t.n_srows)
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
            e2e_steps = sustained(lambda: sol.solve_batch_host(spu_h, V_h, None, None, loss_h))
            barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                it_h, st_h, df_h = sol.solve_batch_host(spu_h, V_h, None, None, loss_h)
            torch.cuda.synchronize()
            assert np.array_equal(np.asarray(it_h), its), 'host path and device path disagree on iteration counts'
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
            X = last.X

            def e2e_step():
                spu_d.copy_(spu_h, non_blocking=True)
                last.solve()
                X_h.copy_(X, non_blocking=True)
                it_h.copy_(iters, non_blocking=True)
                stream.synchronize()
            e2e_step()
            e2e_steps = sustained(e2e_step)
            barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                e2e_step()
            e2e_api = 'pinned spu -> device, gp_nr3_solve_batch, X and iterations -> pinned host'
            d2h = int(X_h.numel() * 16 + S * 4)
        e2e_t = torch.tensor([(time.perf_counter() - t0) / e2e_steps], dtype=f64, device=dev)
        if world > 1:
            dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
        h2d = int(spu_h.numel() * 16) if h2d_override is None else h2d_override
        e2e = {'value': world * S / float(e2e_t.item()), 'unit': 'solves/s', 'h2d_bytes_per_step': h2d,
               'd2h_bytes_per_step': d2h, 'steps': e2e_steps, 'api': e2e_api,
               'pcie_gbs_per_rank': (h2d + d2h) / float(e2e_t.item()) / 1e9, **e2e_extra}

    res = None
    if rank == 0:
        pk, pk_kind = peaks()
        solve_ms = float(np.mean(ms_solve))
        n_it = float(its.sum())
        kernel_names = {1: 'fbs_dfs_kernel', 2: 'fbs_phase_kernel', 3: 'fbs_solve_kernel', 4: 'fbs_onchip_kernel',
                        5: 'gp_fbs_jit_kernel', 6: 'fbs_walk_kernel'}
        variant = {0: None, 1: 'fused-dfs', 2: 'phase-parallel', 3: 'streaming (state in HBM), forward and backward steps of the two sweeps '
                      + ('in depth-first order' if args.fbs_order else 'as two full sweeps'),
                   4: 'on-chip two-sweep (state in shared memory), table-driven',
                   5: 'on-chip, feeder-specialised (NVRTC): two sweeps with V and I in shared memory, or depth-first with I '
                      'in shared memory and V in the caller\'s array',
                   6: 'tree walk (depth-first, current bus in registers, rows in HBM)'}[active]
        # per sweep read I, write V | read V, read spu, write I, write V; once: read spu, write V, I
        stream_per_it = 16 * (2 * t.n_irows + 3 * t.n_vrows + t.n_srows)
        stream_once = 16 * (t.n_srows + t.n_vrows + t.n_irows)
        if solver == 'fbs' and active in (4, 5):
            # on-chip kernel: V and I cross the HBM interface once, the load rows once (they are re-read
            # through L2 every sweep, which is not DRAM traffic)
            per_it, once = 0, 16 * (t.n_srows + t.n_vrows + t.n_irows) + 8 * t.n_vrows + 16
            kname = kernel_names[active]
        elif solver == 'fbs':
            per_it, once = stream_per_it, stream_once
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
        traffic = ncu_traffic(name, kname)
        roofline = {'bound': 'hbm', 'kernel': kname, 'achieved': achieved,
                    'peak': pk['hbm_gbs'], 'peak_kind': pk_kind, 'unit': 'GB/s',
                    'frac': achieved / pk['hbm_gbs'], 'traffic': traffic,
                    'algorithmic_bytes_per_launch': alg_bytes, 'kernel_ms': solve_ms,
                    'kernel_share_of_step': solve_ms / ms_per_step}
        if solver == 'fbs' and active in (4, 5):
            tf = ctypes.c_double(0.0)
            _cabi.check(_cabi.gp_measure_fp64_peak(local, ctypes.byref(tf)))
            flops = n_it * fbs_flops_per_iteration(t)
            ach_tf = flops / (solve_ms * 1e-3) / 1e12
            roofline = {'bound': 'fp64', 'kernel': kname, 'achieved': ach_tf, 'peak': tf.value,
                        'peak_kind': 'measured live (gp_measure_fp64_peak: DFMA chains, 2 flop each)',
                        'unit': 'TFLOP/s', 'frac': ach_tf / tf.value, 'traffic': traffic,
                        'algorithmic_flops_per_launch': flops, 'kernel_ms': solve_ms,
                        'kernel_share_of_step': solve_ms / ms_per_step,
                        'hbm': {'achieved': achieved, 'peak': pk['hbm_gbs'], 'peak_kind': pk_kind, 'unit': 'GB/s',
                                'frac': achieved / pk['hbm_gbs'], 'algorithmic_bytes_per_launch': alg_bytes,
                                'note': 'one-time bytes only (SURVEY 8(d): spu in, V, I, |V|, loss out); the state never '
                                        'leaves the SM between sweeps',
                                'streaming_formulation_equivalent_gbs':
                                    (n_it * stream_per_it + S * stream_once) / (solve_ms * 1e-3) / 1e9}}
        if world > 1:
            nv_bytes = (world - 1) * (t.n_vrows * 8 + 16) * S
            roofline['nvlink'] = {'ingress_bytes_per_gpu_per_step': nv_bytes,
                                  'achieved_gbs': nv_bytes / (ms_per_step * 1e-3) / 1e9, 'peak_gbs': 900.0,
                                  'note': 'all-gather of |V| and losses; overlaps the next step\'s solve, so this is '
                                          'bytes over the whole step time (a lower bound of the link rate)'}
        res = {
            'metric': 'power-flow solves/sec', 'value': value, 'unit': 'solves/s', 'n_gpus': world,
            'steps': steps, 'repeats': repeats, 'warmup': n_warm, 'ms_per_step': ms_per_step,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
            'data': 'synthetic',
            'config': {'workload': desc, 'solver': solver, 'scenarios_per_gpu': S,
                       'fbs_kernel': variant,
                       'rows': {'P_b': t.n_vrows, 'P_l': t.n_irows, 'P_L': t.n_srows},
                       'mean_iterations': float(its.mean()), 'max_iterations': int(its.max()),
                       'l2': f'no reuse across steps: {n_sets} rotating buffer set(s) of {sets[0].nbytes() >> 20} MiB '
                             f'(inputs + state + outputs) against a 126 MiB L2',
                       'timing': f'CUDA events around {repeats} x {steps} back-to-back steps on the launch stream '
                                 f'(ms_per_step = total / {total_steps}; gather of step k on a side stream, overlapping '
                                 f'the solve of step k+1), max over ranks; solve kernels timed individually for the roofline',
                       'parallelism': f'scenario-sharded x{world}, ' + (
                           'single GPU: no exchange' if world == 1 else
                           'all-gather of |V| and losses by NCCL on a side stream' if pg is None else
                           'all-gather of |V| and losses: the solve kernel writes this rank\'s block of its own gather '
                           'buffer, copy engines push it to every peer on a side stream' if fused and args.gather == 'dma' else
                           'all-gather of |V| and losses fused into the solve kernel (epilogue stores over NVLink into '
                           'every rank\'s peer-mapped buffer)' if fused else
                           'all-gather of |V| and losses by copy-engine pushes into every rank\'s peer-mapped buffer '
                           'on a side stream')},
            'gpu_launches': int(launches),
            'roofline': roofline,
            'clocks': clocks, 'wall_s_timed_region': t_wall,
        }
        if gather_note:
            res['config']['gather_note'] = gather_note
        if ctx.world > 1:
            res['config']['host_affinity'] = (f'rank 0 pinned to the {len(ctx.numa_cores)} cores NVML reports local to its GPU '
                                              f'(every rank likewise)' if ctx.numa_cores else 'not set')
        if e2e is not None:
            res['e2e'] = e2e
        if main and world == 1 and not args.no_cpu:
            host_rows = rows if not on_device else rows[:, :16384].cpu().numpy()
            nb = sol.circuit.buses.num_elements
            res['cpu_baseline'] = time_port(feeder, solver, host_rows, nb, 10.0)
            try:
                ref = time_reference_python(feeder, solver, host_rows, sol, 10.0, mult=mult)
            except Exception as e:          # the reference leg must never cost the line
                ref = {'unavailable': f'{type(e).__name__}: {e}'}
            res['cpu_baseline']['reference'] = ref if ref is not None else {
                'unavailable': 'baseline/_ref not shipped (run __graft_entry__.build() where /root/reference exists)'
                if not reference_available() else 'the reference holds 15 GB of dense H per process for this feeder'}
    # release this workload's memory before the next one
    if pg is not None:
        pg.close()
    del sets[:]
    del sol
    torch.cuda.empty_cache()
    return res


def run_ours(args):
    import torch
    import torch.distributed as dist
    ctx = Ctx()
    ctx.world = int(os.environ.get('WORLD_SIZE', '1'))
    ctx.rank = int(os.environ.get('RANK', '0'))
    ctx.local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(ctx.local)
    ctx.dev = torch.device('cuda', ctx.local)
    ctx.numa_cores = None
    if ctx.world > 1 and not args.no_numa:
        # one process per GPU: run on the cores of the GPU's NUMA node, so the pinned e2e buffers are local to it
        from gigpower_b200.sharding import bind_to_gpu_numa
        ctx.numa_cores = bind_to_gpu_numa(ctx.local)
    if ctx.world > 1:
        # NCCL writes its version / debug lines to stdout by default; stdout carries the one JSON line
        os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')
        dist.init_process_group('nccl', device_id=ctx.dev)
    line = run_workload(ctx, args, args.workload, args.steps, main=True)
    subs = [] if args.sub in ('none', '') else [s for s in args.sub.split(',') if s and s != args.workload]
    sub_results = {}
    for name in subs:
        if name not in WORKLOADS:
            raise SystemExit(f'unknown sub-workload {name}')
        try:
            r = run_workload(ctx, args, name, max(3, min(args.steps, args.sub_steps)), main=False)
        except Exception as e:      # a failing sub-workload is reported, not fatal to the headline line
            if ctx.world > 1:
                raise
            r = {'error': f'{type(e).__name__}: {e}'}
        if ctx.rank == 0:
            if 'error' in r:
                sub_results[name] = r
            else:
                rf = r['roofline']
                sub_results[name] = {
                    'workload': r['config']['workload'], 'value': r['value'], 'unit': r['unit'],
                    'ms_per_step': r['ms_per_step'], 'steps': r['steps'], 'kernel': rf['kernel'],
                    'kernel_ms': rf['kernel_ms'], 'bound': rf['bound'], 'frac': rf['frac'],
                    'achieved': rf['achieved'], 'peak': rf['peak'], 'roofline_unit': rf['unit'],
                    'traffic': rf['traffic'], 'hbm_frac': rf.get('hbm', rf)['frac'],
                    'mean_iterations': r['config']['mean_iterations'], 'gpu_launches': r['gpu_launches'],
                    'clocks': r['clocks'], 'scenarios_per_gpu': r['config']['scenarios_per_gpu'],
                    **({'nvlink': rf['nvlink']} if 'nvlink' in rf else {}),
                    **({'e2e': r['e2e']} if 'e2e' in r else {})}
    if ctx.rank == 0:
        if sub_results:
            line['sub_results'] = sub_results
        emit(line)
    if ctx.world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument('--sub', default=None,
                    help='comma-separated workloads reported under sub_results after the headline one '
                         f'(default {",".join(DEFAULT_SUBS)} with the default workload, none otherwise)')
    ap.add_argument('--sub-steps', type=int, default=10, help='timed steps of each sub-workload')
    ap.add_argument('--sub-e2e', action='store_true', help='also measure e2e for the sub-workloads')
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    ap.add_argument('--no-numa', action='store_true', help='N > 1: do not pin the ranks to their GPU\'s NUMA node')
    ap.add_argument('--gather', default='peer', choices=['peer', 'dma', 'nccl'],
                    help='N > 1: the all-gather of |V| and losses through peer-mapped memory (fused into the '
                         'specialised FBS kernel, copy-engine pushes otherwise) or by NCCL')
    ap.add_argument('--jit-ig', type=int, default=-1, choices=[-1, 0, 1],
                    help='I rows of the specialised FBS kernel in global memory: -1 automatic, 0 no, 1 yes')
    ap.add_argument('--jit-dfs', type=int, default=-1, choices=[-1, 0, 1],
                    help='depth-first build of the specialised FBS kernel (V rows off chip): -1 automatic, 0 no, 1 yes')
    ap.add_argument('--host-warps', type=int, default=0,
                    help='warps per SM of the one-launch host path (e2e); 0 = the library default')
    ap.add_argument('--push-streams', type=int, default=0,
                    help='N > 1: copy streams the copy-engine all-gather spreads its destinations over (0 = library default)')
    ap.add_argument('--host-chunks', type=int, default=0,
                    help='pipeline chunks of the staged host path (e2e of feeders without a specialised kernel); 0 = automatic')
    ap.add_argument('--share-rows', type=int, default=1, choices=[0, 1],
                    help='streaming FBS kernel: 1 = I rows that repeat a child\'s current or stay zero make no HBM round '
                         'trips (default), 0 = every row travels')
    ap.add_argument('--nr3-ring', type=int, default=-1, choices=[-1, 0, 1],
                    help='NR3 phase-lane kernel: 1 = next-step rows staged in shared memory with cp.async, 0 = direct loads, -1 = automatic')
    ap.add_argument('--fbs-order', type=int, default=1, choices=[0, 1],
                    help='streaming FBS kernel: 1 = depth-first order of the forward / backward steps (default), 0 = two full sweeps')
    ap.add_argument('--nr3-kernel', type=int, default=0, choices=[0, 1, 2],
                    help='0 automatic, 1 one thread per scenario, 2 three phase lanes per scenario')
    ap.add_argument('--prefetch', type=int, default=None, help='L2 prefetch distance of the streaming FBS kernel / the NR3 kernel')
    ap.add_argument('--walk-blocks', type=int, default=4, choices=[4, 5, 6, 8],
                    help='tree-walk FBS kernel: resident 128-thread blocks per SM it is compiled for (128/96/80/64 registers)')
    ap.add_argument('--fbs-kernel', type=int, default=0, choices=[0, 1, 2, 3, 4, 5, 6],
                    help='0 = automatic (on-chip when the feeder fits in shared memory, else two-sweep streaming), '
                         '1 = fused depth-first, 2 = phase-parallel lanes, 3 = two-sweep streaming, 4 = on-chip table-driven, '
                         '5 = on-chip feeder-specialised, 6 = tree walk')
    args = ap.parse_args()
    if args.sub is None:
        args.sub = ','.join(DEFAULT_SUBS) if args.workload == DEFAULT_WORKLOAD else 'none'
    protect_stdout()
    if args.impl == 'reference':
        run_reference(args, WORKLOADS[args.workload])
    else:
        run_ours(args)


if __name__ == '__main__':
    main()

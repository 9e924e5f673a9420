"""Golden vectors on RANDOM radial feeders from the UNMODIFIED reference -- TEST INFRASTRUCTURE.

The reference's own test feeders (13 / 34 / 37 buses) and the two authored ones fix eight topologies.  This script
writes seeded random radial feeders - random tree shapes, 1- / 2- / 3-phase laterals (some declared in reverse
phase order, e.g. ``.3.2``), linecode / inline-matrix / sequence-impedance / switch lines, 1-, 2- and 3-phase
loads, capacitors, and on some feeders per-phase regulators with random taps and a 4.16 / 0.48 kV transformer -
as ``.dss`` text, runs the reference's own ``SolutionFBS`` / ``SolutionNR3`` on them behind the ``opendssdirect``
stand-in and commits

* ``tests/golden/random_feeders/<name>.json``   the parsed inputs (FeederModel JSON, taps included);
* ``tests/golden/ref_random.npz``               V, I, sV, Stx, Srx, iterations, topo order (FBS) and XNR, V,
                                                iterations (NR3) at full precision: nominal loads and two seeded
                                                load-multiplier scenarios per feeder.

Runs only in the build container (needs ``/root/reference``).  Usage: python oracle/make_golden_random.py
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg                                      # noqa: E402  (sets up the reference import)

OUT = os.path.join(mg.OUT, 'random_feeders')
# name -> (seed, buses, regulators?, transformer?)
FEEDERS = {f'rnd{i:02d}': (9100 + i, nb, reg, xfm) for i, (nb, reg, xfm) in enumerate([
    (6, False, False), (9, False, False), (12, False, False), (14, True, False), (17, False, False),
    (20, False, True), (22, True, False), (25, False, False), (28, True, True), (31, False, False),
    (34, False, False), (40, True, False), (11, False, False), (16, True, True), (23, False, False),
    (36, False, False)])}
NR3_MAX_BUSES = 26          # the reference's dense H grows as N^3


def lower_tri(m):
    return ' | '.join(' '.join(f'{m[i][j]:.9g}' for j in range(i + 1)) for i in range(len(m)))


def make_feeder(name, seed, n_bus, with_reg, with_xfm, exotic=False):
    """.dss text of one random radial feeder and its regulator taps.  exotic (the fuzzer's, oracle/fuzz_random_feeders.py;
    the committed fixtures are made without it): lines and loads below the transformer, a second regulator bank,
    regulator banks on a subset of the parent's phases, generation (negative kW), loads up to 8 x heavier so that
    some scenarios stop at maxiter or go non-finite."""
    rng = np.random.Generator(np.random.PCG64(seed))
    out = ['Clear', f'! random radial feeder {name} (oracle/make_golden_random.py, seed {seed})',
           f'New Circuit.{name} basekv=4.16 pu=1.0 phases=3 bus1=src MVAsc3=20000 MVAsc1=21000']
    # line codes: symmetric positive-definite-ish blocks around the IEEE overhead configurations, per kft
    codes = {3: [], 2: [], 1: []}
    for n in (3, 2, 1):
        for k in range(3):
            r = np.full((n, n), 0.029) + np.diag(np.full(n, 0.058)) + rng.uniform(0, 0.01, (n, n))
            x = np.full((n, n), 0.08) + np.diag(np.full(n, 0.12)) + rng.uniform(0, 0.02, (n, n))
            r, x = (r + r.T) / 2 * rng.uniform(0.8, 3.0), (x + x.T) / 2 * rng.uniform(0.8, 1.5)
            cname = f'c{n}{"abc"[k]}'
            codes[n].append((cname, r, x))
            out.append(f'New Linecode.{cname} nphases={n} units=kft rmatrix=[{lower_tri(r)}] xmatrix=[{lower_tri(x)}]')
    buses = [('src', (1, 2, 3))]
    kids = {0: 0}
    depth = {0: 0}
    edges = []
    taps = {}
    n_reg = 0
    reg_at = int(rng.integers(1, max(2, n_bus // 2))) if with_reg else -1
    xfm_at = n_bus - 2 if with_xfm else -1
    reg2_at, reg_parent = -1, -1
    if exotic:
        if with_xfm:
            xfm_at = int(rng.integers(max(3, n_bus // 2), n_bus - 1))
        if with_reg and n_bus > 8 and rng.random() < 0.5:
            reg2_at = int(rng.integers(n_bus // 2, n_bus - 1))
            reg2_at = -1 if reg2_at in (reg_at, xfm_at) else reg2_at
    lv = set()                                   # buses of the 0.48 kV zone
    while len(buses) < n_bus:
        b = len(buses)
        # parent: the first edge hangs off the source; later ones prefer recent (deep) buses, at most 4 children
        while True:
            p = 0 if b == 1 else int(min(b - 1, max(0, b - 1 - rng.geometric(0.35) + 1)))
            if b == xfm_at:                      # the transformer hangs off a three-phase bus
                three = [i for i in range(1, b) if len(buses[i][1]) == 3 and i not in lv]
                if not three:
                    xfm_at = -1
                    continue
                p = int(rng.choice(three))
            # (two regulator banks off ONE bus are left out: the reference de-duplicates a bus's regulator edges with
            # list(set(...)) - voltage_regulator_group.py:60-72 -, so their order, the topo order and the result then
            # depend on PYTHONHASHSEED; seed 30029 of the fuzzer found it: 3 or 5 sweeps, voltages apart by 1e-6)
            if b == reg2_at and p == reg_parent:
                continue
            if kids.get(p, 0) < (1 if p == 0 else 4) and (p not in lv or (exotic and b != xfm_at and b not in (reg_at, reg2_at))):
                break
        pph = buses[p][1]
        if len(pph) == 3:
            u = rng.random()
            if b <= 2 or u < 0.5:
                ph = pph
            elif u < 0.75:
                ph = tuple(sorted(rng.choice(3, 2, replace=False) + 1))
            else:
                ph = (int(rng.integers(1, 4)),)
        elif len(pph) == 2:
            ph = pph if rng.random() < 0.5 else (int(rng.choice(pph)),)
        else:
            ph = pph
        ph = tuple(int(q) for q in ph)
        bname = f'b{b}'
        spec = list(ph)
        if len(spec) == 2 and rng.random() < 0.4:
            spec = spec[::-1]                    # declared in reverse order (line.py / utils.py pad_phases sort them)
        dots = ''.join(f'.{q}' for q in spec)
        if exotic and b in (reg_at, reg2_at) and len(ph) > 1 and rng.random() < 0.4:
            ph = tuple(sorted(int(q) for q in rng.choice(ph, len(ph) - 1, replace=False)))   # a bank on some phases only
        if b in (reg_at, reg2_at):
            reg_parent = p if b == reg_at else reg_parent
            # per-phase regulators (the 13-bus feeder's pattern): parent -> this bus through one 1-phase
            # regulating transformer per phase, each with its own tap
            for q in ph:
                n_reg += 1
                rname = f'reg{n_reg}'
                out.append(f'New Transformer.{rname} phases=1 XHL=0.01 kVAs=[1666 1666]')
                out.append(f'~ Buses=[{buses[p][0]}.{q} {bname}.{q}] kVs=[2.4 2.4] %LoadLoss=0.01')
                out.append(f'New regcontrol.{rname} transformer={rname} winding=2 vreg=122 band=2 ptratio=20 ctprim=700 R=3 X=9')
                taps[rname] = int(rng.integers(-8, 13))
        elif b == xfm_at and len(pph) == 3:
            ph = (1, 2, 3)
            out.append(f'New Transformer.xfm1 Phases=3 Windings=2 XHL=2')
            out.append(f'~ wdg=1 bus={buses[p][0]} conn=Wye kv=4.16 kva=500 %r=.55 XHT=1')
            out.append(f'~ wdg=2 bus={bname} conn=Wye kv=0.480 kva=500 %r=.55 XLT=1')
            lv.add(b)
        else:
            n = len(ph)
            kind = rng.random()
            length = float(np.round(rng.uniform(0.08, 0.9), 3))
            head = f'New Line.l{p}_{b} Phases={n} Bus1={buses[p][0]}{dots} Bus2={bname}{dots}'
            if kind < 0.06 and n == 3:
                out.append(head + ' Switch=y r1=1e-3 r0=1e-3 x1=0 x0=0 c1=0 c0=0')
            elif kind < 0.16 and n == 3:
                out.append(head + f' r1={rng.uniform(0.05, 0.12):.4f} x1={rng.uniform(0.1, 0.2):.4f} '
                                  f'r0={rng.uniform(0.15, 0.3):.4f} x0={rng.uniform(0.3, 0.6):.4f} c1=0 c0=0 '
                                  f'Length={length} units=kft')
            elif kind < 0.3:
                _, r, x = codes[n][int(rng.integers(3))]
                out.append(head + f' rmatrix=[{lower_tri(r * 1.1)}] xmatrix=[{lower_tri(x * 0.9)}] Length={length} units=kft')
            else:
                cname = codes[n][int(rng.integers(3))][0]
                out.append(head + f' LineCode={cname} Length={length} units=kft')
        if p in lv:
            lv.add(b)
        buses.append((bname, ph))
        kids[p] = kids.get(p, 0) + 1
        kids[b] = 0
        depth[b] = depth[p] + 1
        edges.append((p, b))
    # loads: light enough that every feeder converges from flat start, heavy enough for 5-12 sweeps
    heavy = rng.uniform(0.6, 1.6) * (1.0 if seed % 3 else 3.0)      # every third feeder is loaded hard
    if exotic:
        heavy *= float(rng.choice([1.0, 1.0, 2.5, 5.0, 8.0]))
    n_load = 0
    for b in range(1, n_bus):
        bname, ph = buses[b]
        if rng.random() > 0.65:
            continue
        kv_ln, kv_ll = (0.277, 0.48) if b in lv else (2.4, 4.16)
        scale = (0.15 if b in lv else 1.0) * heavy * 14.0 / n_bus
        n_load += 1
        u = rng.random()
        if len(ph) == 3 and u < 0.3:
            kw = rng.uniform(60, 300) * scale
            out.append(f'New Load.ld{n_load} Bus1={bname}.1.2.3 Phases=3 Conn=Wye Model=1 kV={kv_ll} '
                       f'kW={kw:.2f} kvar={kw * rng.uniform(0.2, 0.6):.2f}')
        elif len(ph) >= 2 and u < 0.45:
            q = sorted(int(v) for v in rng.choice(ph, 2, replace=False))
            kw = rng.uniform(40, 160) * scale
            out.append(f'New Load.ld{n_load} Bus1={bname}.{q[0]}.{q[1]} Phases=2 Conn=Wye Model=1 kV={kv_ll} '
                       f'kW={kw:.2f} kvar={kw * rng.uniform(0.2, 0.6):.2f}')
        else:
            q = int(rng.choice(ph))
            kw = rng.uniform(15, 120) * scale
            if exotic and rng.random() < 0.15:
                kw = -kw                         # generation
            out.append(f'New Load.ld{n_load} Bus1={bname}.{q} Phases=1 Conn=Wye Model=1 kV={kv_ln} '
                       f'kW={kw:.2f} kvar={kw * rng.uniform(0.2, 0.6):.2f}')
    if n_load == 0:
        bname, ph = buses[-1]
        out.append(f'New Load.ld1 Bus1={bname}.{ph[0]} Phases=1 Conn=Wye Model=1 kV=2.4 kW=80 kvar=30')
    for c in range(int(rng.integers(0, 3))):
        b = int(rng.integers(1, n_bus))
        if b in lv:
            continue
        bname, ph = buses[b]
        if len(ph) == 3 and rng.random() < 0.5:
            out.append(f'New Capacitor.cap{c + 1} Bus1={bname}.1.2.3 phases=3 conn=wye kVAR={rng.uniform(100, 400):.1f} kV=4.16')
        else:
            out.append(f'New Capacitor.cap{c + 1} Bus1={bname}.{int(rng.choice(ph))} phases=1 conn=wye '
                       f'kVAR={rng.uniform(30, 120):.1f} kV=2.4')
    out.append('Set Voltagebases=[4.16, 0.48]' if lv else 'Set Voltagebases=[4.16]')
    out.append('calcv')
    out.append('Solve')
    return '\n'.join(out) + '\n', taps


def main():
    os.makedirs(OUT, exist_ok=True)
    z = mg.ZIPS['test_zip']
    res = {}
    tmp = tempfile.mkdtemp(prefix='gp_random_')
    parts = None
    if '--extras' in sys.argv:                   # only (some of) the extras: python ... --extras droop,wpu_nr3
        parts = sys.argv[sys.argv.index('--extras') + 1].split(',')
    for name, (seed, nb, reg, xfm) in FEEDERS.items():
        text, taps = make_feeder(name, seed, nb, reg, xfm)
        fp = os.path.join(tmp, name + '.dss')
        with open(fp, 'w') as fh:
            fh.write(text)
        if parts is not None:
            continue
        mg.dss.TAPS = taps
        model = mg.read_dss(fp, taps)
        with open(os.path.join(OUT, name + '.json'), 'w') as fh:
            fh.write(model.to_json())
        rng = np.random.Generator(np.random.PCG64(seed + 5000))
        mult = rng.uniform(0.4, 1.4, size=(3, len(model.loads)))
        mult[0] = 1.0
        res[f'{name}/mult'] = mult
        V, I, it, sV, Stx, Srx, conv = [], [], [], [], [], [], []
        for i in range(len(mult)):
            s = mg.ref_fbs(fp, z, mult[i])
            V.append(s.V.copy()), I.append(s.I.copy()), it.append(s.iterations), sV.append(s.sV.copy())
            Stx.append(s.Stx.copy()), Srx.append(s.Srx.copy()), conv.append(bool(s.converged))
        res[f'{name}/FBS/V'], res[f'{name}/FBS/I'] = np.array(V), np.array(I)
        res[f'{name}/FBS/sV'], res[f'{name}/FBS/Stx'], res[f'{name}/FBS/Srx'] = np.array(sV), np.array(Stx), np.array(Srx)
        res[f'{name}/FBS/iterations'], res[f'{name}/FBS/converged'] = np.array(it), np.array(conv)
        res[f'{name}/FBS/topo_order'] = np.array([s.circuit.buses.get_idx(n) for n in s.topo_order])
        msg = f'{name}: {nb} buses, {len(model.lines)} lines, {len(model.loads)} loads, taps {taps}, FBS it {it} {conv}'
        if nb <= NR3_MAX_BUSES:
            X, Vn, itn = [], [], []
            for i in range(2):
                n, its = mg.ref_nr3(fp, z, mult[i])
                X.append(n.XNR[:, 0].copy()), Vn.append(n.V.copy()), itn.append(its)
                del n
            res[f'{name}/NR3/XNR'], res[f'{name}/NR3/V'], res[f'{name}/NR3/iterations'] = np.array(X), np.array(Vn), np.array(itn)
            msg += f', NR3 it {itn}'
        print(msg, flush=True)
    if parts is None:
        np.savez_compressed(os.path.join(mg.OUT, 'ref_random.npz'), **res)
        print('wrote', os.path.join(mg.OUT, 'ref_random.npz'))
    extras(tmp, parts or ALL_PARTS)


TAP_CASES = {'rnd03': 3, 'rnd06': 3, 'rnd08': 0, 'rnd11': 0, 'rnd13': 3}      # feeder: NR3 scenarios (FBS: 5 each)
WARM_CASES = {'rnd02': (4, 3), 'rnd03': (4, 3), 'rnd05': (4, 3), 'rnd10': (4, 0)}   # feeder: (FBS steps, NR3 steps)
WPU_CASES = ['rnd04', 'rnd13', 'rnd09']
DROOP_CASES = ['rnd09', 'rnd13', 'rnd10', 'rnd11']
WPU_NR3_CASES = ['rnd04', 'rnd13']
DROOP_POINTS, DROOP_Q = (0.94, 0.97, 1.0, 1.03), (-0.08, 0.08)
ALL_PARTS = ['taps', 'warm', 'wpu', 'droop', 'wpu_nr3']


def extras(tmp, parts):
    """The per-scenario inputs of SURVEY.md section 8(f)-2, -3 on the random feeders, each as the reference defines it:
    taps = one reference object per scenario built with that scenario's taps (oracle/make_golden_taps.py), warm start =
    one reference object driven through a load sequence (make_golden_warm.py), wpu = reference objects whose ``wpu``
    attribute is set (make_golden_wpu.py).  -> tests/golden/ref_random_extras.npz"""
    from make_golden_warm import set_loads
    z = mg.ZIPS['test_zip']
    path = os.path.join(mg.OUT, 'ref_random_extras.npz')
    out = {}
    if os.path.exists(path):                     # parts not asked for keep what the file holds
        old = np.load(path)
        out = {k: old[k] for k in old.files if k.split('/')[1] not in parts}
    for name, n_nr3 in (TAP_CASES.items() if 'taps' in parts else ()):
        seed = FEEDERS[name][0]
        fp = os.path.join(tmp, name + '.dss')
        model = mg.read_dss(fp, {})
        regs = [rc.name for rc in model.regcontrols]
        rng = np.random.Generator(np.random.PCG64(seed + 7000))
        taps = rng.integers(-16, 17, size=(5, len(regs)))
        taps[0], taps[1] = 16, -16
        mult = rng.uniform(0.5, 1.4, size=(5, len(model.loads)))
        out[name + '/taps/regs'], out[name + '/taps/taps'], out[name + '/taps/mult'] = np.array(regs), taps, mult
        V, I, it = [], [], []
        for i in range(5):
            mg.dss.TAPS = {r: int(t) for r, t in zip(regs, taps[i])}
            s = mg.ref_fbs(fp, z, mult[i])
            V.append(s.V.copy()), I.append(s.I.copy()), it.append(s.iterations)
        out[name + '/taps/FBS/V'], out[name + '/taps/FBS/I'], out[name + '/taps/FBS/iterations'] = np.array(V), np.array(I), np.array(it)
        X, itn = [], []
        for i in range(n_nr3):
            mg.dss.TAPS = {r: int(t) for r, t in zip(regs, taps[i])}
            n, its = mg.ref_nr3(fp, z, mult[i])
            X.append(n.XNR[:, 0].copy()), itn.append(its)
            del n
        if n_nr3:
            out[name + '/taps/NR3/XNR'], out[name + '/taps/NR3/iterations'] = np.array(X), np.array(itn)
        print(name, 'taps: FBS it', it, 'NR3 it', itn, flush=True)
    for name, (n_fbs, n_nr3) in (WARM_CASES.items() if 'warm' in parts else ()):
        seed = FEEDERS[name][0]
        fp = os.path.join(tmp, name + '.dss')
        _, taps = make_feeder(name, *FEEDERS[name])
        mg.dss.TAPS = taps
        rng = np.random.Generator(np.random.PCG64(seed + 8000))
        s = mg.SolutionFBS(fp, zip_v=np.asarray(z))
        nload = s.circuit.loads.num_elements
        level = 1.0 + np.cumsum(rng.uniform(-0.15, 0.15, size=n_fbs))
        mult = level[:, None] * (1 + rng.normal(0, 0.03, size=(n_fbs, nload)))
        mult[0] = 1.0
        out[name + '/warm/mult'] = mult
        base = [(s.circuit.loads.get_element(n).kW, s.circuit.loads.get_element(n).kvar) for n in s.circuit.loads.all_names()]
        V, I, it = [], [], []
        for h in range(n_fbs):
            set_loads(s, base, mult[h])
            s._set_sV_params()
            s.Vtest = np.zeros(3, dtype=complex)
            s.iterations = 0
            s.solve()
            V.append(s.V.copy()), I.append(s.I.copy()), it.append(s.iterations)
        out[name + '/warm/FBS/V'], out[name + '/warm/FBS/I'], out[name + '/warm/FBS/iterations'] = np.array(V), np.array(I), np.array(it)
        X, itn = [], []
        if n_nr3:
            n = mg.SolutionNR3(fp, zip_v=np.asarray(z))
            for h in range(n_nr3):
                set_loads(n, base, mult[h])
                n._init_KCL_matrices()
                mg._ft_calls[0] = 0
                n.solve()
                X.append(n.XNR[:, 0].copy()), itn.append(mg._ft_calls[0])
            out[name + '/warm/NR3/XNR'], out[name + '/warm/NR3/iterations'] = np.array(X), np.array(itn)
            del n
        print(name, 'warm: FBS it', it, 'NR3 it', itn, flush=True)
    for name in (WPU_CASES if 'wpu' in parts else ()):
        seed = FEEDERS[name][0]
        fp = os.path.join(tmp, name + '.dss')
        _, taps = make_feeder(name, *FEEDERS[name])
        mg.dss.TAPS = taps
        rng = np.random.Generator(np.random.PCG64(seed + 9000))
        s0 = mg.SolutionFBS(fp, zip_v=np.asarray(z))
        nb = s0.circuit.buses.num_elements
        pm = s0.phase_matrix.astype(bool)
        n = 4
        mult = rng.uniform(0.6, 1.3, size=(n, s0.circuit.loads.num_elements))
        wpu = np.zeros((n, nb, 3))
        for i in range(n):
            b, p = np.nonzero(pm)
            pick = rng.permutation(len(b))[:max(2, len(b) // 4)]
            wpu[i, b[pick], p[pick]] = rng.uniform(-0.05, 0.05, size=len(pick))
        out[name + '/wpu/mult'], out[name + '/wpu/wpu'] = mult, wpu
        V, I, sV, it = [], [], [], []
        for i in range(n):
            s = mg.SolutionFBS(fp, zip_v=np.asarray(z))
            for j, lname in enumerate(s.circuit.loads.all_names()):
                ld = s.circuit.loads.get_element(lname)
                s.circuit.set_kW(lname, ld.kW * mult[i][j])
                s.circuit.set_kvar(lname, ld.kvar * mult[i][j])
            s._set_sV_params()
            s.wpu = wpu[i].copy()
            s.solve()
            V.append(s.V.copy()), I.append(s.I.copy()), sV.append(s.sV.copy()), it.append(s.iterations)
        out[name + '/wpu/FBS/V'], out[name + '/wpu/FBS/I'], out[name + '/wpu/FBS/sV'] = np.array(V), np.array(I), np.array(sV)
        out[name + '/wpu/FBS/iterations'] = np.array(it)
        print(name, 'wpu: FBS it', it, flush=True)
    if 'droop' in parts:
        # volt-var droop INSIDE the FBS iteration: the reference's SolutionFBS with its own VoltVARController objects
        # consulted before every calc_sV (oracle/make_golden_vvc_iter.py has the construction)
        from make_golden_vvc_iter import DroopFBS, VoltVARController
        out['droop/points'], out['droop/q'] = np.array(DROOP_POINTS), np.array(DROOP_Q)
        for name in DROOP_CASES:
            seed = FEEDERS[name][0]
            fp = os.path.join(tmp, name + '.dss')
            _, taps = make_feeder(name, *FEEDERS[name])
            mg.dss.TAPS = taps
            rng = np.random.Generator(np.random.PCG64(seed + 10000))
            s0 = mg.SolutionFBS(fp, zip_v=np.asarray(z))
            nb = s0.circuit.buses.num_elements
            pm = s0.phase_matrix.astype(bool)
            names = s0.circuit.buses.all_names()
            cand = [k for k in range(1, nb) if pm[k].all()]
            ctl = [(k, p) for k in (cand[len(cand) // 2], cand[-1]) for p in range(3)]
            single = [k for k in range(1, nb) if pm[k].sum() == 1]
            if single:
                ctl.append((single[-1], int(np.nonzero(pm[single[-1]])[0][0])))
            n = 4
            mult = rng.uniform(0.5, 1.5, size=(n, s0.circuit.loads.num_elements))
            out[name + '/droop/ctl'], out[name + '/droop/mult'] = np.array(ctl), mult
            V, I, sV, it, Q = [], [], [], [], []
            for i in range(n):
                s = DroopFBS(fp, zip_v=np.asarray(z))
                for j, lname in enumerate(s.circuit.loads.all_names()):
                    ld = s.circuit.loads.get_element(lname)
                    s.circuit.set_kW(lname, ld.kW * mult[i][j])
                    s.circuit.set_kvar(lname, ld.kvar * mult[i][j])
                s._set_sV_params()
                s.controllers = [((k, p), VoltVARController(DROOP_POINTS, *DROOP_Q, names[k], p)) for k, p in ctl]
                s.solve()
                V.append(s.V.copy()), I.append(s.I.copy()), sV.append(s.sV.copy()), it.append(s.iterations)
                Q.append(np.array([s.wpu[k, p] for k, p in ctl]))
            out[name + '/droop/V'], out[name + '/droop/I'], out[name + '/droop/sV'] = np.array(V), np.array(I), np.array(sV)
            out[name + '/droop/iterations'], out[name + '/droop/Q'] = np.array(it), np.array(Q)
            print(name, 'droop: FBS it', it, 'max|Q|', np.abs(np.array(Q)).max(), flush=True)
    for name in (WPU_NR3_CASES if 'wpu_nr3' in parts else ()):
        # a non-zero wpu in the reference's SolutionNR3 (oracle/make_golden_wpu_nr3.py has the construction)
        seed = FEEDERS[name][0]
        fp = os.path.join(tmp, name + '.dss')
        _, taps = make_feeder(name, *FEEDERS[name])
        mg.dss.TAPS = taps
        rng = np.random.Generator(np.random.PCG64(seed + 11000))
        s0 = mg.SolutionNR3(fp, zip_v=np.asarray(z))
        nb = s0.circuit.buses.num_elements
        pm = np.asarray(s0.circuit.buses.get_phase_matrix('rows')).astype(bool)
        n = 3
        mult = rng.uniform(0.6, 1.3, size=(n, s0.circuit.loads.num_elements))
        wpu = np.zeros((n, nb, 3))
        for i in range(n):
            b, p = np.nonzero(pm[1:])
            pick = rng.permutation(len(b))[:max(2, len(b) // 4)]
            wpu[i, b[pick] + 1, p[pick]] = rng.uniform(-0.04, 0.04, size=len(pick))
        del s0
        out[name + '/wpu_nr3/mult'], out[name + '/wpu_nr3/wpu'] = mult, wpu
        X, sV, it = [], [], []
        for i in range(n):
            s = mg.SolutionNR3(fp, zip_v=np.asarray(z))
            for j, lname in enumerate(s.circuit.loads.all_names()):
                ld = s.circuit.loads.get_element(lname)
                s.circuit.set_kW(lname, ld.kW * mult[i][j])
                s.circuit.set_kvar(lname, ld.kvar * mult[i][j])
            s._init_KCL_matrices()
            w = wpu[i]
            s.circuit.get_wpu_matrix = (lambda w=w, c=s.circuit: c._orient_switch(w.copy()))
            s.change_KCL_matrices(der=1e-300)
            mg._ft_calls[0] = 0
            s.solve()
            X.append(s.XNR[:, 0].copy()), sV.append(s.sV.copy()), it.append(mg._ft_calls[0])
            del s
        out[name + '/wpu_nr3/XNR'], out[name + '/wpu_nr3/sV'], out[name + '/wpu_nr3/iterations'] = np.array(X), np.array(sV), np.array(it)
        print(name, 'wpu_nr3: it', it, flush=True)
    np.savez_compressed(path, **out)
    print('wrote', path)


if __name__ == '__main__':
    main()

"""A feeder on which the UNMODIFIED reference is not deterministic, and the answer this engine pins -- TEST INFRASTRUCTURE.

Found by oracle/fuzz_random_feeders.py (seed 30029): bus ``b1`` feeds TWO regulator banks (``b1 -> b4`` on three
phases, ``b1 -> b5`` on two).  ``VoltageRegulatorGroup.get_adj_set`` (voltage_regulator_group.py:60-72) de-duplicates a
bus's regulator edges with ``list(set(neighbours))``, so the order of ``b4`` and ``b5`` in ``SolutionFBS.adj['b1']``
follows the string hashes of the process: the topo order, which child rewrites ``b1``'s voltage last in the backward
sweep, the iteration count (3 or 5) and the voltages (1e-6 apart) all depend on ``PYTHONHASHSEED``.

This script runs the reference's ``SolutionFBS`` on that feeder in fresh interpreters with PYTHONHASHSEED = 0..7 and
writes ``tests/golden/ref_two_banks.npz``: every distinct (topo order, iterations, V, I) it produced, and which of them
lists the banks in the order of definition - the one ``gigpower_b200.flatten`` and ``oracle/gp_oracle.py`` reproduce.
The feeder itself goes to ``tests/golden/random_feeders/two_banks.json``.
Runs only in the build container.  Usage: python oracle/make_golden_two_banks.py
"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SEED, N_BUS = 30029, 11
TAPS = {'reg1': 5, 'reg2': 10, 'reg3': 9, 'reg4': -5, 'reg5': -3}


def feeder_text():
    """The fuzzer's feeder of seed 30029 as it was generated (two banks off one bus are no longer generated)."""
    return '''Clear
! bus b1 feeds two regulator banks: the reference's result depends on PYTHONHASHSEED (oracle/make_golden_two_banks.py)
New Circuit.two_banks basekv=4.16 pu=1.0 phases=3 bus1=src MVAsc3=20000 MVAsc1=21000
New Linecode.c3a nphases=3 units=kft rmatrix=[0.172 | 0.0611 0.1702 | 0.0605 0.0617 0.1735] xmatrix=[0.2306 | 0.0966 0.2261 | 0.0921 0.1003 0.2345]
New Linecode.c2b nphases=2 units=kft rmatrix=[0.2501 | 0.0896 0.2411] xmatrix=[0.1583 | 0.0687 0.1594]
New Linecode.c1a nphases=1 units=kft rmatrix=[0.2273] xmatrix=[0.2697]
New Linecode.c1b nphases=1 units=kft rmatrix=[0.1871] xmatrix=[0.2102]
New Linecode.c1c nphases=1 units=kft rmatrix=[0.2044] xmatrix=[0.2106]
New Line.l0_1 Phases=3 Bus1=src.1.2.3 Bus2=b1.1.2.3 LineCode=c3a Length=0.11 units=kft
New Line.l1_2 Phases=3 Bus1=b1.1.2.3 Bus2=b2.1.2.3 Switch=y r1=1e-3 r0=1e-3 x1=0 x0=0 c1=0 c0=0
New Line.l2_3 Phases=3 Bus1=b2.1.2.3 Bus2=b3.1.2.3 LineCode=c3a Length=0.653 units=kft
New Transformer.reg1 phases=1 XHL=0.01 kVAs=[1666 1666]
~ Buses=[b1.1 b4.1] kVs=[2.4 2.4] %LoadLoss=0.01
New regcontrol.reg1 transformer=reg1 winding=2 vreg=122 band=2 ptratio=20 ctprim=700 R=3 X=9
New Transformer.reg2 phases=1 XHL=0.01 kVAs=[1666 1666]
~ Buses=[b1.2 b4.2] kVs=[2.4 2.4] %LoadLoss=0.01
New regcontrol.reg2 transformer=reg2 winding=2 vreg=122 band=2 ptratio=20 ctprim=700 R=3 X=9
New Transformer.reg3 phases=1 XHL=0.01 kVAs=[1666 1666]
~ Buses=[b1.3 b4.3] kVs=[2.4 2.4] %LoadLoss=0.01
New regcontrol.reg3 transformer=reg3 winding=2 vreg=122 band=2 ptratio=20 ctprim=700 R=3 X=9
New Transformer.reg4 phases=1 XHL=0.01 kVAs=[1666 1666]
~ Buses=[b1.1 b5.1] kVs=[2.4 2.4] %LoadLoss=0.01
New regcontrol.reg4 transformer=reg4 winding=2 vreg=122 band=2 ptratio=20 ctprim=700 R=3 X=9
New Transformer.reg5 phases=1 XHL=0.01 kVAs=[1666 1666]
~ Buses=[b1.3 b5.3] kVs=[2.4 2.4] %LoadLoss=0.01
New regcontrol.reg5 transformer=reg5 winding=2 vreg=122 band=2 ptratio=20 ctprim=700 R=3 X=9
New Line.l5_6 Phases=1 Bus1=b5.1 Bus2=b6.1 LineCode=c1a Length=0.871 units=kft
New Line.l5_7 Phases=2 Bus1=b5.3.1 Bus2=b7.3.1 LineCode=c2b Length=0.636 units=kft
New Line.l5_8 Phases=1 Bus1=b5.1 Bus2=b8.1 LineCode=c1c Length=0.497 units=kft
New Line.l5_9 Phases=1 Bus1=b5.3 Bus2=b9.3 LineCode=c1c Length=0.734 units=kft
New Line.l9_10 Phases=1 Bus1=b9.3 Bus2=b10.3 LineCode=c1b Length=0.883 units=kft
New Load.ld1 Bus1=b2.1.2.3 Phases=3 Conn=Wye Model=1 kV=4.16 kW=240 kvar=90
New Load.ld2 Bus1=b3.2 Phases=1 Conn=Wye Model=1 kV=2.4 kW=95 kvar=40
New Load.ld3 Bus1=b4.1.2.3 Phases=3 Conn=Wye Model=1 kV=4.16 kW=310 kvar=120
New Load.ld4 Bus1=b6.1 Phases=1 Conn=Wye Model=1 kV=2.4 kW=70 kvar=25
New Load.ld5 Bus1=b7.1.3 Phases=2 Conn=Wye Model=1 kV=4.16 kW=130 kvar=45
New Load.ld6 Bus1=b8.1 Phases=1 Conn=Wye Model=1 kV=2.4 kW=55 kvar=20
New Load.ld7 Bus1=b10.3 Phases=1 Conn=Wye Model=1 kV=2.4 kW=85 kvar=35
New Capacitor.cap1 Bus1=b9.3 phases=1 conn=wye kVAR=60 kV=2.4
Set Voltagebases=[4.16]
calcv
Solve
'''


def child(fp):
    """In a fresh interpreter (its own PYTHONHASHSEED): solve and print the result as JSON."""
    sys.path.insert(0, HERE)
    import make_golden as mg
    mg.dss.TAPS = TAPS
    s = mg.ref_fbs(fp, mg.ZIPS['test_zip'])
    print(json.dumps({'topo': [s.circuit.buses.get_idx(n) for n in s.topo_order], 'adj_b1': list(s.adj['b1']),
                      'it': int(s.iterations), 'V': [s.V.real.tolist(), s.V.imag.tolist()],
                      'I': [s.I.real.tolist(), s.I.imag.tolist()]}))


def main():
    if len(sys.argv) > 2 and sys.argv[1] == '--child':
        return child(sys.argv[2])
    sys.path.insert(0, HERE)
    import make_golden as mg
    tmp = tempfile.mkdtemp(prefix='gp_two_banks_')
    fp = os.path.join(tmp, 'two_banks.dss')
    with open(fp, 'w') as fh:
        fh.write(feeder_text())
    model = mg.read_dss(fp, TAPS)
    with open(os.path.join(mg.OUT, 'random_feeders', 'two_banks.json'), 'w') as fh:
        fh.write(model.to_json())
    variants = {}
    for h in range(8):
        env = dict(os.environ, PYTHONHASHSEED=str(h))
        res = subprocess.run([sys.executable, os.path.abspath(__file__), '--child', fp], env=env, capture_output=True, text=True)
        r = json.loads(res.stdout.strip().splitlines()[-1])
        key = tuple(r['adj_b1'])
        variants.setdefault(key, dict(r, seeds=[]))['seeds'].append(h)
    out = {}
    for k, (adj, r) in enumerate(sorted(variants.items())):
        banks = [b for b in adj if b in ('b4', 'b5')]
        out[f'v{k}/banks'] = np.array(banks)
        out[f'v{k}/definition_order'] = np.bool_(banks == ['b4', 'b5'])
        out[f'v{k}/hash_seeds'] = np.array(r['seeds'])
        out[f'v{k}/topo_order'], out[f'v{k}/iterations'] = np.array(r['topo']), np.int64(r['it'])
        out[f'v{k}/V'] = np.array(r['V'][0]) + 1j * np.array(r['V'][1])
        out[f'v{k}/I'] = np.array(r['I'][0]) + 1j * np.array(r['I'][1])
        print(f'variant {k}: adj[b1] = {list(adj)}, PYTHONHASHSEED in {r["seeds"]}, {r["it"]} sweeps, topo {r["topo"]}')
    out['n_variants'] = np.int64(len(variants))
    np.savez_compressed(os.path.join(mg.OUT, 'ref_two_banks.npz'), **out)
    print('wrote', os.path.join(mg.OUT, 'ref_two_banks.npz'))


if __name__ == '__main__':
    main()

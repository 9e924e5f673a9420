"""Circuit -> flat device tables (the host "flattener").

FBS: the sweep order of the reference is fixed by ``SolutionFBS.set_adj``
(``src/solution_fbs.py:39-74``) and ``topo_sort`` (``src/utils.py:130-167``);
both are reproduced here on bus indices, then every non-root bus becomes one
"step" (bus, parent, incoming edge) of the table described by ``GpFbsDesc`` in
``include/gigpower_b200.h``.  State rows are packed: only existing
(bus, phase), (edge, phase) and loaded (bus, phase) pairs get a row.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List

import numpy as np

from .circuit import Circuit

EDGE_LINE, EDGE_VR = 0, 1


def fbs_adjacency(c: Circuit) -> Dict[str, List[str]]:
    """Directed adjacency by bus name: lines in definition order, then
    transformers, then one edge per regulated bus pair with the upstream leg
    removed (solution_fbs.py:39-74)."""
    adj: Dict[str, List[str]] = {}
    for e in c.lines.get_elements():
        adj.setdefault(e.tx, []).append(e.rx)
    for e in c.transformers.get_elements():
        adj.setdefault(e.tx, []).append(e.rx)
    pairs: Dict[str, List[str]] = {}
    for vr in c.voltage_regulators.get_elements():
        for a, b in ((vr.tx, vr.rx), (vr.rx, vr.tx)):
            if b not in pairs.setdefault(a, []):
                pairs[a].append(b)
    for node, nbrs in pairs.items():
        adj[node] = adj.get(node, []) + nbrs
    for vr in c.voltage_regulators.get_elements():
        if vr.tx in adj.get(vr.rx, []):
            adj[vr.rx].remove(vr.tx)
    return adj


def topo_sort(bus_names: List[str], adj: Dict[str, List[str]]) -> List[str]:
    """Reverse DFS finishing order, DFS roots in bus order, children in
    adjacency order; iterative.  Raises on a cycle like utils.py:163-164."""
    n = len(bus_names)
    order: List[str] = [None] * n
    state = {b: 0 for b in bus_names}           # 0 new, 1 active, 2 finished
    clock = n
    for start in bus_names:
        if state[start] != 0:
            continue
        state[start] = 1
        stack = [(start, iter(adj.get(start, [])))]
        while stack:
            node, it = stack[-1]
            advanced = False
            for ch in it:
                if state[ch] == 0:
                    state[ch] = 1
                    stack.append((ch, iter(adj.get(ch, []))))
                    advanced = True
                    break
                if state[ch] == 1:
                    raise ValueError('Not radial. Contains a cycle.')
            if not advanced:
                stack.pop()
                state[node] = 2
                clock -= 1
                order[clock] = node
    return order


def gamma_row_map(kind: np.ndarray, vr_gamma: np.ndarray, gamma_vr: np.ndarray):
    """Rows of a per-scenario regulator-ratio array (``GpSolveExtras.gamma_dev``): one per regulated
    (step, phase), numbered in step order and, inside a step, phase order - the rule the library
    applies to ``vr_gamma`` (``gp_gamma_rows``).  Returns ``grow [n_steps, 3]`` (row or -1) and
    ``grow_vr [n_rows]``, the index of the regulator (order of ``circuit.voltage_regulators``) whose
    tap sets that row."""
    grow = np.full(vr_gamma.shape, -1, np.int32)
    owner = []
    for i in range(vr_gamma.shape[0]):
        for p in range(3):
            if kind[i] == EDGE_VR and vr_gamma[i, p] != 0.0:
                grow[i, p] = len(owner)
                owner.append(int(gamma_vr[i, p]))
    return grow, np.array(owner, dtype=np.int32)


def load_weights(c: Circuit, srow: np.ndarray, n_srows: int) -> np.ndarray:
    """``[n_srows, n_loads]`` complex W with ``spu_rows = W @ mult.T`` for per-load multipliers
    (load.py:16-24: a load's power split per phase of its bus-name suffix, summed per bus-phase).
    Host-only: needs the circuit and the packed-row map of either solver's tables, no device."""
    W = np.zeros((n_srows, c.loads.num_elements), dtype=complex)
    for j, e in enumerate(c.loads.get_elements()):
        b = c.buses.get_idx(e.related_bus)
        for p in e.get_ph_idx_matrix():
            W[srow[b, p], j] += e.spu[p]
    return W


def reverse_adjacency(adj: Dict[str, List[str]]) -> Dict[str, List[str]]:
    rev: Dict[str, List[str]] = {}
    for node, nbrs in adj.items():
        rev.setdefault(node, [])
        for v in nbrs:
            rev.setdefault(v, []).append(node)
    return rev


@dataclass
class FbsTables:
    """NumPy image of ``GpFbsDesc`` plus the row maps used to (un)pack results."""
    n_steps: int
    n_vrows: int
    n_irows: int
    n_srows: int
    n_lines: int
    bus_vrow: np.ndarray
    par_vrow: np.ndarray
    edge_irow: np.ndarray
    load_srow: np.ndarray
    kind: np.ndarray
    child_begin: np.ndarray
    child_count: np.ndarray
    child_irow: np.ndarray
    z: np.ndarray
    cappu: np.ndarray
    vr_gamma: np.ndarray
    root_vrow: np.ndarray
    root_load_srow: np.ndarray
    root_cappu: np.ndarray
    vref: np.ndarray
    zip: np.ndarray
    line_irow: np.ndarray
    line_tx_vrow: np.ndarray
    line_rx_vrow: np.ndarray
    # host-side maps
    vrow: np.ndarray = field(default=None)      # [nb,3]  V row or -1
    irow: np.ndarray = field(default=None)      # [nl+ntf+nvr,3]  I row or -1 (VR rows: all -1)
    srow: np.ndarray = field(default=None)      # [nb,3]  spu row or -1
    topo_order: List[str] = field(default=None)
    adj: Dict[str, List[str]] = field(default=None)
    reverse_adj: Dict[str, List[str]] = field(default=None)
    tolerance: float = 0.0
    depth: int = 0
    grow: np.ndarray = field(default=None)      # [n_steps,3] row of a per-scenario gamma array or -1
    grow_vr: np.ndarray = field(default=None)   # [n_grows] regulator index behind each gamma row


def flatten_fbs(c: Circuit) -> FbsTables:
    names = c.buses.all_names()
    nb = len(names)
    bidx = {n: i for i, n in enumerate(names)}
    adj = fbs_adjacency(c)
    order = topo_sort(names, adj)
    rev = reverse_adjacency(adj)
    root = order[0]
    bus_pm = c.buses.get_phase_matrix('rows')
    if bus_pm[bidx[root]].sum() != 3:
        raise NotImplementedError('the root (source) bus must carry all three phases')

    # packed rows
    vrow = np.full((nb, 3), -1, dtype=np.int32)
    k = 0
    for b in range(nb):
        for p in range(3):
            if bus_pm[b, p]:
                vrow[b, p] = k
                k += 1
    n_vrows = k
    nl, ntf, nvr = c.lines.num_elements, c.transformers.num_elements, c.voltage_regulators.num_elements
    edges = c.lines.get_elements() + c.transformers.get_elements()
    irow = np.full((nl + ntf + nvr, 3), -1, dtype=np.int32)
    k = 0
    for i, e in enumerate(edges):
        for p in range(3):
            if e.phase_matrix[p]:
                irow[i, p] = k
                k += 1
    n_irows = k
    lmask = c.load_phase_mask()
    srow = np.full((nb, 3), -1, dtype=np.int32)
    k = 0
    for b in range(nb):
        for p in range(3):
            if lmask[b, p] and bus_pm[b, p]:
                srow[b, p] = k
                k += 1
    n_srows = k
    cappu = c._by_bus(c.capacitors, 'cappu', float)

    # edge lookup: real lines win over transformers win over regulators (line_group.py:26-38)
    edge_of = {}
    for vr in c.voltage_regulators.get_elements():
        edge_of.setdefault(vr.key, ('vr', []))[1].append(vr)
    for i, e in enumerate(c.transformers.get_elements()):
        edge_of[e.key] = ('line', nl + i, e)
    for i, e in enumerate(c.lines.get_elements()):
        edge_of[e.key] = ('line', i, e)

    n_steps = nb - 1
    S3 = (n_steps, 3)
    t = dict(bus_vrow=np.full(S3, -1, np.int32), par_vrow=np.full(S3, -1, np.int32),
             edge_irow=np.full(S3, -1, np.int32), load_srow=np.full(S3, -1, np.int32),
             kind=np.zeros(n_steps, np.int32), child_begin=np.zeros(n_steps, np.int32),
             child_count=np.zeros(n_steps, np.int32), z=np.zeros((n_steps, 3, 3, 2)),
             cappu=np.zeros(S3), vr_gamma=np.zeros(S3))
    child_rows: List[np.ndarray] = []
    gamma_vr = np.full(S3, -1, np.int32)
    vr_index = {id(vr): i for i, vr in enumerate(c.voltage_regulators.get_elements())}
    depth_of = {root: 0}
    for pos, name in enumerate(order[1:]):
        parents = rev.get(name, [])
        if len(parents) > 1:
            raise ValueError(f"Network not radial. Bus {name} has parents {parents}")
        if len(parents) == 0:
            raise ValueError(f"Bus {name} is not connected to the root {root}")
        parent = parents[0]
        depth_of[name] = depth_of.get(parent, 0) + 1
        b, pb = bidx[name], bidx[parent]
        t['bus_vrow'][pos] = vrow[b]
        t['par_vrow'][pos] = vrow[pb]
        t['load_srow'][pos] = srow[b]
        t['cappu'][pos] = cappu[b]
        e = edge_of[(parent, name)]
        if e[0] == 'vr':
            t['kind'][pos] = EDGE_VR
            for vr in e[1]:
                for p in vr.get_ph_idx_matrix():
                    t['vr_gamma'][pos, p] = vr.gamma
                    gamma_vr[pos, p] = vr_index[id(vr)]
        else:
            t['kind'][pos] = EDGE_LINE
            t['edge_irow'][pos] = irow[e[1]]
            t['z'][pos, :, :, 0] = e[2].FZpu.real
            t['z'][pos, :, :, 1] = e[2].FZpu.imag
        t['child_begin'][pos] = len(child_rows)
        for ch in adj.get(name, []):
            ce = edge_of[(name, ch)]
            if ce[0] == 'line':          # VR children contribute vr.Itx == 0 (solution_fbs.py:278-281)
                child_rows.append(irow[ce[1]])
        t['child_count'][pos] = len(child_rows) - t['child_begin'][pos]
    child_irow = (np.array(child_rows, dtype=np.int32).reshape(-1, 3) if child_rows
                  else np.zeros((0, 3), np.int32))

    vref = np.array([1, np.exp(1j * 240 * np.pi / 180), np.exp(1j * 120 * np.pi / 180)], dtype=complex)
    tol = abs((vref[1]) * 10 ** -6)                       # solution_fbs.py:37
    rb = bidx[root]
    line_irow = irow[:nl].copy()
    tx = np.array([bidx[e.tx] for e in c.lines.get_elements()], dtype=int)
    rx = np.array([bidx[e.rx] for e in c.lines.get_elements()], dtype=int)
    grow, grow_vr = gamma_row_map(t['kind'], t['vr_gamma'], gamma_vr)
    return FbsTables(
        grow=grow, grow_vr=grow_vr,
        n_steps=n_steps, n_vrows=n_vrows, n_irows=n_irows, n_srows=n_srows, n_lines=nl,
        child_irow=child_irow, root_vrow=vrow[rb].copy(), root_load_srow=srow[rb].copy(),
        root_cappu=cappu[rb].copy(), vref=np.stack([vref.real, vref.imag], axis=1).reshape(-1).copy(),
        zip=np.array([c.loads.aPQ_p, c.loads.aI_p, c.loads.aZ_p], dtype=float),
        line_irow=line_irow, line_tx_vrow=vrow[tx].reshape(-1, 3).copy() if nl else np.zeros((0, 3), np.int32),
        line_rx_vrow=vrow[rx].reshape(-1, 3).copy() if nl else np.zeros((0, 3), np.int32),
        vrow=vrow, irow=irow, srow=srow, topo_order=order, adj=adj, reverse_adj=rev,
        tolerance=float(tol), depth=max(depth_of.values()) if depth_of else 0, **t)


# ==========================================================================
# NR3
# ==========================================================================
@dataclass
class Nr3Tables:
    """NumPy image of ``GpNr3Desc`` plus the maps between the packed device
    state and the reference's ``XNR`` ordering (solution_nr3.py:46-61)."""
    n_steps: int
    n_vrows: int
    n_irows: int
    n_srows: int
    n_lines: int
    bus_vrow: np.ndarray
    par_vrow: np.ndarray
    edge_irow: np.ndarray
    edge_orow: np.ndarray
    load_srow: np.ndarray
    kind: np.ndarray
    child_begin: np.ndarray
    child_count: np.ndarray
    child_step: np.ndarray
    child_irow: np.ndarray
    z: np.ndarray
    cappu: np.ndarray
    vr_gamma: np.ndarray
    root_vrow: np.ndarray
    root_load_srow: np.ndarray
    root_cappu: np.ndarray
    vslack: np.ndarray
    zip: np.ndarray
    cA: np.ndarray
    cB: np.ndarray
    f0_absent: float
    line_irow: np.ndarray
    line_tx_vrow: np.ndarray
    line_rx_vrow: np.ndarray
    vrow: np.ndarray = field(default=None)       # [nb,3]
    irow: np.ndarray = field(default=None)       # [n_edges,3] rows of the current entering the child bus
    srow: np.ndarray = field(default=None)       # [nb,3]
    N: int = 0                                    # length of XNR
    xnr_index: np.ndarray = field(default=None)  # [n_vrows+n_irows] even XNR index of each packed row
    order: List[int] = field(default=None)
    grow: np.ndarray = field(default=None)       # [n_steps,3] row of a per-scenario gamma array or -1
    grow_vr: np.ndarray = field(default=None)    # [n_grows] regulator index behind each gamma row


def flatten_nr3(c: Circuit) -> Nr3Tables:
    """Tree tables for the block elimination of ``gp_nr3.cu``.  Supported: radial
    feeders of Lines, Transformers and VoltageRegulators oriented away from the
    slack bus (index 0); every bus must carry exactly the phases of its incoming edge."""
    names = c.buses.all_names()
    nb, nl, ntf = len(names), c.lines.num_elements, c.transformers.num_elements
    bidx = {n: i for i, n in enumerate(names)}
    bus_pm = c.buses.get_phase_matrix('rows')
    if bus_pm[0].sum() != 3:
        raise NotImplementedError('the slack bus must carry all three phases')
    # edges: lines, transformers, then one edge per regulated bus pair (its VRs merged per phase)
    edges = [dict(kind=EDGE_LINE, tx=e.tx, rx=e.rx, pm=np.asarray(e.phase_matrix), el=e)
             for e in c.lines.get_elements() + c.transformers.get_elements()]
    vr_edges, cnt = {}, 0
    vr_index = {id(vr): i for i, vr in enumerate(c.voltage_regulators.get_elements())}
    for vr in c.voltage_regulators.get_elements():        # VR-phase counter in (VR, phase) order (:471-514)
        e = vr_edges.setdefault(vr.key, dict(kind=EDGE_VR, tx=vr.tx, rx=vr.rx, pm=np.zeros(3, dtype=int),
                                             gamma=np.zeros(3), slot={}, vr=np.full(3, -1, np.int32)))
        for p in vr.get_ph_idx_matrix():
            e['pm'][p], e['gamma'][p], e['slot'][int(p)] = 1, vr.gamma, cnt
            e['vr'][p] = vr_index[id(vr)]
            cnt += 1
    vr_lines = cnt
    edges += list(vr_edges.values())
    parent, edge_in, children = {}, {}, {b: [] for b in range(nb)}
    for ei, e in enumerate(edges):
        tx, rx = bidx[e['tx']], bidx[e['rx']]
        if rx in parent:
            raise ValueError(f"Network not radial. Bus {e['rx']} has parents {[names[parent[rx]], e['tx']]}")
        parent[rx], edge_in[rx] = tx, ei
        children[tx].append(rx)
    order, stack = [], [0]
    seen = {0}
    while stack:                      # parents before children; any such order is valid
        u = stack.pop()
        order.append(u)
        for ch in reversed(children[u]):
            if ch in seen:
                raise ValueError('Not radial. Contains a cycle.')
            seen.add(ch)
            stack.append(ch)
    if len(order) != nb:
        raise ValueError('Network not radial / not connected to the slack bus')

    vrow = np.full((nb, 3), -1, dtype=np.int32)
    k = 0
    for b in range(nb):
        for p in range(3):
            if bus_pm[b, p]:
                vrow[b, p] = k
                k += 1
    n_vrows = k
    ne = len(edges)
    irow = np.full((ne, 3), -1, dtype=np.int32)
    k = 0
    for i, e in enumerate(edges):
        for p in range(3):
            if e['pm'][p]:
                irow[i, p] = k
                k += 1
    orow = irow.copy()
    for i, e in enumerate(edges):                 # out-legs of the regulators
        if e['kind'] == EDGE_VR:
            for p in range(3):
                if e['pm'][p]:
                    orow[i, p] = k
                    k += 1
    n_irows = k
    lmask = c.load_phase_mask()
    srow = np.full((nb, 3), -1, dtype=np.int32)
    k = 0
    for b in range(nb):
        for p in range(3):
            if lmask[b, p] and bus_pm[b, p]:
                srow[b, p] = k
                k += 1
    n_srows = k
    cappu = c._by_bus(c.capacitors, 'cappu', float)

    n_steps = nb - 1
    step_of = {b: i for i, b in enumerate(order[1:])}
    S3 = (n_steps, 3)
    t = dict(bus_vrow=np.full(S3, -1, np.int32), par_vrow=np.full(S3, -1, np.int32),
             edge_irow=np.full(S3, -1, np.int32), edge_orow=np.full(S3, -1, np.int32),
             load_srow=np.full(S3, -1, np.int32), kind=np.zeros(n_steps, np.int32),
             child_begin=np.zeros(n_steps, np.int32), child_count=np.zeros(n_steps, np.int32),
             z=np.zeros((n_steps, 3, 3, 2)), cappu=np.zeros(S3), vr_gamma=np.zeros(S3))
    child_step, child_rows = [], []
    gamma_vr = np.full(S3, -1, np.int32)
    for pos, b in enumerate(order[1:]):
        e = edges[edge_in[b]]
        if not np.array_equal(e['pm'], bus_pm[b]):
            raise NotImplementedError(f"bus {names[b]}: phases differ from its incoming edge")
        t['bus_vrow'][pos], t['par_vrow'][pos] = vrow[b], vrow[parent[b]]
        t['edge_irow'][pos], t['edge_orow'][pos] = irow[edge_in[b]], orow[edge_in[b]]
        t['load_srow'][pos], t['cappu'][pos], t['kind'][pos] = srow[b], cappu[b], e['kind']
        if e['kind'] == EDGE_VR:
            t['vr_gamma'][pos] = e['gamma']
            gamma_vr[pos] = e['vr']
        elif e['el'].kind == 'Line':   # rmat/xmat (line.py:48-73); identical numbers to FZpu on existing phases
            R, X = e['el'].rmat.reshape(3, 3), e['el'].xmat.reshape(3, 3)
            pm = e['pm']
            t['z'][pos, :, :, 0] = R * pm[None, :] * pm[:, None]
            t['z'][pos, :, :, 1] = X * pm[None, :] * pm[:, None]
        t['child_begin'][pos] = len(child_step)
        for ch in children[b]:
            child_step.append(step_of[ch])
            child_rows.append(orow[edge_in[ch]])
        t['child_count'][pos] = len(child_step) - t['child_begin'][pos]

    # second-order ZIP constants per phase (solution_nr3.py:228-261)
    zp = np.array([c.loads.aPQ_p, c.loads.aI_p, c.loads.aZ_p], dtype=float)
    cA, cB = np.zeros(3), np.zeros(3)
    for ph in range(3):
        A0, B0 = (1, 0) if ph == 0 else ((-1 / 2, -1 * np.sqrt(3) / 2) if ph == 1 else (-1 / 2, np.sqrt(3) / 2))
        h00 = -((A0 ** 2) * (A0 ** 2 + B0 ** 2) ** (-3 / 2)) + (A0 ** 2 + B0 ** 2) ** (-1 / 2)
        h11 = -((B0 ** 2) * (A0 ** 2 + B0 ** 2) ** (-3 / 2)) + ((A0 ** 2 + B0 ** 2) ** (-1 / 2))
        cA[ph] = zp[2] + (0.5 * zp[1] * h00)
        cB[ph] = zp[2] + (0.5 * zp[1] * h11)
    vslack = np.array([1, np.exp(1j * -120 * np.pi / 180), np.exp(1j * 120 * np.pi / 180)], dtype=complex)
    f0 = 0.0
    for b in range(1, nb):
        for p in range(3):
            if not bus_pm[b, p]:
                f0 = max(f0, abs(vslack[p].real), abs(vslack[p].imag))

    # packed row -> XNR index (solution_nr3.py:46-61): lines, transformer currents, then per
    # regulated phase the in-leg (even) and out-leg (odd) currents
    tf_lines = int(sum(len(e.phases) for e in c.transformers.get_elements()))
    N = 2 * 3 * (nb + nl) + 2 * tf_lines + 2 * 2 * vr_lines
    xnr = np.zeros(n_vrows + n_irows, dtype=np.int64)
    for b in range(nb):
        for p in range(3):
            if vrow[b, p] >= 0:
                xnr[vrow[b, p]] = 2 * nb * p + 2 * b
    for l in range(nl):
        for p in range(3):
            if irow[l, p] >= 0:
                xnr[n_vrows + irow[l, p]] = 6 * nb + 2 * p * nl + 2 * l
    cnt = 0
    for ti in range(ntf):
        for p in range(3):
            if irow[nl + ti, p] >= 0:
                xnr[n_vrows + irow[nl + ti, p]] = 6 * (nb + nl) + 2 * cnt
                cnt += 1
    base_r = 6 * (nb + nl) + 2 * tf_lines
    for i, e in enumerate(edges):
        if e['kind'] == EDGE_VR:
            for p, slot in e['slot'].items():
                xnr[n_vrows + irow[i, p]] = base_r + 2 * (2 * slot)
                xnr[n_vrows + orow[i, p]] = base_r + 2 * (2 * slot + 1)
    tx = np.array([bidx[e.tx] for e in c.lines.get_elements()], dtype=int)
    rx = np.array([bidx[e.rx] for e in c.lines.get_elements()], dtype=int)
    return Nr3Tables(
        n_steps=n_steps, n_vrows=n_vrows, n_irows=n_irows, n_srows=n_srows, n_lines=nl,
        child_step=np.array(child_step, dtype=np.int32),
        child_irow=(np.array(child_rows, dtype=np.int32).reshape(-1, 3) if child_rows else np.zeros((0, 3), np.int32)),
        root_vrow=vrow[0].copy(), root_load_srow=srow[0].copy(), root_cappu=cappu[0].copy(),
        vslack=np.stack([vslack.real, vslack.imag], axis=1).reshape(-1).copy(), zip=zp, cA=cA, cB=cB,
        f0_absent=float(f0), line_irow=irow[:nl].copy(),
        line_tx_vrow=vrow[tx].reshape(-1, 3).copy() if nl else np.zeros((0, 3), np.int32),
        line_rx_vrow=vrow[rx].reshape(-1, 3).copy() if nl else np.zeros((0, 3), np.int32),
        vrow=vrow, irow=irow, srow=srow, N=N, xnr_index=xnr, order=order,
        grow=gamma_row_map(t['kind'], t['vr_gamma'], gamma_vr)[0],
        grow_vr=gamma_row_map(t['kind'], t['vr_gamma'], gamma_vr)[1], **t)

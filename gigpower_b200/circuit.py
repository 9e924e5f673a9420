"""Host-side data model behind the drop-in solvers.

Mirrors the part of gigpower's ``Circuit`` interface that callers use
(reference ``src/circuit.py:36-184``, ``src/circuit_element_group.py:39-91``):
six element groups addressed by name or index in OpenDSS order, the per-bus
matrices the solvers read, and ``set_kW`` / ``set_kvar``.  Unlike the
reference it is built from a :class:`~gigpower_b200.dss_reader.FeederModel`
(no OpenDSS), keeps everything as flat NumPy arrays, and exists only to be
flattened into device tables; no solver arithmetic happens here.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Dict, List, Sequence, Tuple, Union

import numpy as np
import pandas as pd

from .dss_reader import FeederModel

SBASE = 10 ** 6      # reference circuit.py:36, circuit_element.py:25


def _bus(spec: str) -> str:
    return spec.split('.')[0]


def _phases_of_spec(spec: str) -> List[str]:
    """Phase labels from a bus spec; no suffix means all three (utils.py:16-27)."""
    return spec.split('.')[1:] if '.' in spec else ['1', '2', '3']


def _mask(phases: Sequence[Union[str, int]]) -> np.ndarray:
    m = np.zeros(3, dtype=int)
    for p in phases:
        m[int(p) - 1] = 1
    return m


def _scatter_block(block: np.ndarray, mask: np.ndarray) -> np.ndarray:
    """Place an n_ph x n_ph block into 3x3 at the (sorted) present phases
    (what ``pad_phases`` does, utils.py:74-102)."""
    out = np.zeros((3, 3), dtype=block.dtype)
    idx = np.nonzero(mask)[0]
    flat = block.flatten()
    k = 0
    for r in idx:
        for c in idx:
            if k < flat.size:
                out[r, c] = flat[k]
                k += 1
    return out


class Element(SimpleNamespace):
    """One circuit element; ``__name__`` as in the reference classes."""

    def get_ph_idx_matrix(self) -> np.ndarray:
        return np.asarray([int(ph) - 1 for ph in self.phases])

    def __str__(self):
        return f"{self.kind}, {self.__name__}, {self.phases}"


class ElementGroup:
    """Name <-> index container (circuit_element_group.py:39-91)."""

    def __init__(self, kind: str, elements: List[Element]):
        self.kind = kind
        self._elements = list(elements)
        self._names = [e.__name__ for e in elements]
        self._name_to_idx_dict = {n: i for i, n in enumerate(self._names)}
        self._by_key: Dict[Tuple[str, str], Element] = {}
        for e in elements:
            if hasattr(e, 'key'):
                self._by_key[e.key] = e
        self.num_elements = len(elements)

    def all_names(self) -> List[str]:
        return self._names

    def get_idx(self, obj) -> int:
        if isinstance(obj, tuple):
            return self._name_to_idx_dict[self._by_key[obj].__name__]
        name = obj if isinstance(obj, str) else obj.__name__
        return int(self._name_to_idx_dict[name])

    def get_name(self, idx: int) -> str:
        return self._names[idx]

    def get_element(self, key):
        msg = "Invalid key. Key must be str, int, or (tx, rx) tuple."
        if isinstance(key, str):
            return self._elements[self._name_to_idx_dict[key]]
        if isinstance(key, (int, np.integer)):
            try:
                return self._elements[int(key)]
            except IndexError:
                raise KeyError(msg)
        if isinstance(key, tuple):
            try:
                return self._by_key[key]
            except KeyError:
                raise KeyError(msg)
        raise KeyError(msg)

    def get_elements(self):
        return list(self._elements)

    def get_phase_matrix(self, orient: str = 'rows') -> np.ndarray:
        m = np.array([e.phase_matrix for e in self._elements]).reshape(-1, 3)
        return m.T if orient == 'cols' else m

    def get_phase_matrix_dict(self):
        return {e.__name__: e.phase_matrix for e in self._elements}


class Circuit:
    """Flat circuit built from a FeederModel.  ZIP coefficients are class-wide,
    as in the reference (circuit.py:21-34): setting them changes every Circuit."""

    ZIP_V = np.asarray([0.10, 0.05, 0.85, 0.10, 0.05, 0.85, 0.80])
    aZ_p, aI_p, aPQ_p = ZIP_V[0:3]
    aZ_q, aI_q, aPQ_q = ZIP_V[3:6]
    min_voltage_pu = ZIP_V[6]

    @classmethod
    def _set_zip_values(cls, zip_V):
        cls.ZIP_V = np.asarray(zip_V, dtype=float)
        cls.aZ_p, cls.aI_p, cls.aPQ_p = cls.ZIP_V[0:3]
        cls.aZ_q, cls.aI_q, cls.aPQ_q = cls.ZIP_V[3:6]
        cls.min_voltage_pu = cls.ZIP_V[6]

    def __init__(self, model: FeederModel, Sbase=SBASE):
        self.model = model
        self.Sbase = Sbase
        self._orient = 'rows'
        # ---- buses (circuit_element.py:40-51)
        buses = []
        for n in model.bus_names:
            vb = model.bus_kvbase[n] * 1000
            ib = Sbase / vb
            nodes = list(model.bus_nodes[n])
            buses.append(Element(kind='Bus', __name__=n, name=n, related_bus=n, phases=nodes,
                                 phase_matrix=_mask(nodes), Vbase=vb, Ibase=ib, Zbase=vb / ib,
                                 Sbase=Sbase))
        self.buses = ElementGroup('Bus', buses)
        bz = {b.__name__: b for b in buses}

        # ---- lines (line.py:10-73)
        lines = []
        for ln in model.lines:
            tx, rx = _bus(ln.bus1), _bus(ln.bus2)
            phases = ['1', '2', '3'] if ln.is_switch else _phases_of_spec(ln.bus1)
            pm = _mask(phases)
            base = bz[tx]
            mult = 1 / base.Zbase * ln.length
            R = np.asarray(ln.rmatrix, dtype=float)
            X = np.asarray(ln.xmatrix, dtype=float)
            nph = len(phases)
            Zblk = np.reshape(mult * (R + 1j * X), (R.shape[0] // nph, nph))
            if R.size == 1:
                rmat, xmat = (np.identity(3) * R).flatten(), (np.identity(3) * X).flatten()
            elif R.size == 9:
                rmat, xmat = R, X
            elif R.size == 4:
                rmat = _scatter_block(R.reshape(2, 2), pm).flatten()
                xmat = _scatter_block(X.reshape(2, 2), pm).flatten()
            else:
                raise IndexError(f"Xmat is length {X.size}, expected 1, 4, or 9 elements")
            lines.append(Element(kind='Line', __name__=ln.name, name=ln.name, tx=tx, rx=rx, key=(tx, rx),
                                 related_bus=tx, phases=phases, phase_matrix=pm, length=ln.length,
                                 FZpu=_scatter_block(Zblk, pm), rmat=rmat * mult, xmat=xmat * mult,
                                 Vbase=base.Vbase, Ibase=base.Ibase, Zbase=base.Zbase, Sbase=Sbase))
        self.lines = ElementGroup('Line', lines)

        # ---- loads (load.py:9-35)
        loads = []
        for ld in model.loads:
            phases = _phases_of_spec(ld.bus1)
            pm = _mask(phases)
            e = Element(kind='Load', __name__=ld.name, name=ld.name, related_bus=_bus(ld.bus1),
                        phases=phases, phase_matrix=pm, kW=ld.kw, kvar=ld.kvar, zipV=list(Circuit.ZIP_V),
                        Vbase=bz[_bus(ld.bus1)].Vbase, Sbase=Sbase)
            self._set_load_power(e)
            loads.append(e)
        self.loads = ElementGroup('Load', loads)
        self.loads.ZIP_V = Circuit.ZIP_V.copy()
        self.loads.aZ_p, self.loads.aI_p, self.loads.aPQ_p = self.loads.ZIP_V[0:3]

        # ---- capacitors (capacitor.py:7-10; phases are all nodes of the bus)
        caps = []
        for cp in model.capacitors:
            b = bz[_bus(cp.bus1)]
            caps.append(Element(kind='Capacitor', __name__=cp.name, name=cp.name, related_bus=b.__name__,
                                phases=list(b.phases), phase_matrix=b.phase_matrix.copy(),
                                cappu=b.phase_matrix * (cp.kvar * 1000 / Sbase / len(b.phases)),
                                Vbase=b.Vbase, Sbase=Sbase))
        self.capacitors = ElementGroup('Capacitor', caps)

        # ---- transformers that are not regulators (transformer_group.py:10-20)
        regs = [rc.name for rc in model.regcontrols]
        tmap = {t.name: t for t in model.transformers}
        xfms = []
        for t in model.transformers:
            if t.name in regs:
                continue
            tx, rx = _bus(t.buses[0]), _bus(t.buses[1])
            phases = _phases_of_spec(t.buses[0])
            xfms.append(Element(kind='Transformer', __name__=t.name, name=t.name, tx=tx, rx=rx, key=(tx, rx),
                                related_bus=tx, phases=phases, phase_matrix=_mask(phases), length=0,
                                FZpu=np.zeros((3, 3), dtype=complex), kV=t.kvs[-1], kVA=t.kvas[-1],
                                Vbase=bz[tx].Vbase, Sbase=Sbase))
        self.transformers = ElementGroup('Transformer', xfms)

        # ---- voltage regulators (voltage_regulator.py:10-62)
        vrs = []
        for rc in model.regcontrols:
            t = tmap[rc.transformer]
            tx, reg = _bus(t.buses[0]), _bus(t.buses[1])
            phases = _phases_of_spec(t.buses[0])
            gamma = (rc.tap + 16) / 32
            gamma *= 1.10 - .90
            gamma += .90
            vrs.append(Element(kind='VoltageRegulator', __name__=rc.name, name=rc.name, tx=tx, rx=reg,
                               regControl_bus=reg, related_bus=reg, key=(tx, reg), phases=phases,
                               phase_matrix=_mask(phases), transformer_name=rc.transformer, tap=rc.tap,
                               gamma=gamma, Itx=np.zeros(3, dtype=complex), Ireg=np.zeros(3, dtype=complex),
                               Vbase=bz[reg].Vbase, Sbase=Sbase))
        self.voltage_regulators = ElementGroup('VoltageRegulator', vrs)

    # ---- loads -------------------------------------------------------------
    @staticmethod
    def _set_load_power(e: Element):
        n = len(e.phases)
        e.ppu = e.phase_matrix * (e.kW / 1000 / n)          # load.py:16-24
        e.qpu = e.phase_matrix * (e.kvar / 1000 / n)
        e.spu = e.ppu + 1j * e.qpu

    def set_kW(self, load_name: str, kW: float):            # circuit.py:66-76
        e = self.loads.get_element(load_name)
        e.kW = kW
        self._set_load_power(e)

    def set_kvar(self, load_name: str, kvar: float):        # circuit.py:78-88
        e = self.loads.get_element(load_name)
        e.kvar = kvar
        self._set_load_power(e)

    # ---- per-bus matrices (circuit.py:90-184) ---------------------------------
    def _orient_switch(self, m):
        return m.transpose() if self._orient == 'cols' else m

    def _by_bus(self, group, attr, dtype):
        out = np.zeros((self.buses.num_elements, 3), dtype=dtype)
        for e in group.get_elements():
            out[self.buses.get_idx(e.related_bus)] += getattr(e, attr)
        return out

    def get_spu_matrix(self):
        return self._orient_switch(self._by_bus(self.loads, 'spu', complex))

    def get_cappu_matrix(self):
        return self._orient_switch(self._by_bus(self.capacitors, 'cappu', float))

    def load_phase_mask(self) -> np.ndarray:
        """nb x 3, 1 where any load touches the bus-phase (load_group.py:39-52)."""
        m = np.zeros((self.buses.num_elements, 3))
        for e in self.loads.get_elements():
            m[self.buses.get_idx(e.related_bus), e.get_ph_idx_matrix()] = 1
        return m

    def get_aPQ_matrix(self):
        return self._orient_switch(self.load_phase_mask() * self.loads.aPQ_p)

    def get_aI_matrix(self):
        return self._orient_switch(self.load_phase_mask() * self.loads.aI_p)

    def get_aZ_matrix(self):
        return self._orient_switch(self.load_phase_mask() * self.loads.aZ_p)

    def get_wpu_matrix(self):
        return self._orient_switch(np.zeros((self.buses.num_elements, 3), dtype=float))

    def get_total_lines(self) -> int:
        return self.lines.num_elements + self.transformers.num_elements + \
            self.voltage_regulators.num_elements

    def _bus_ids(self, which: str) -> List[str]:
        ids = [getattr(e, which) for e in self.lines.get_elements()]
        ids += [getattr(e, which) for e in self.transformers.get_elements()]
        for vr in self.voltage_regulators.get_elements():    # two synthetic legs per VR
            ids += [vr.tx, vr.rx] if which == 'tx' else [vr.rx, vr.tx]
        return ids

    def get_tx_idx_matrix(self):
        return np.asarray([self.buses.get_idx(b) for b in self._bus_ids('tx')])

    def get_rx_idx_matrix(self):
        return np.asarray([self.buses.get_idx(b) for b in self._bus_ids('rx')])

    def get_nominal_bus_powers(self) -> pd.DataFrame:
        data = self.get_spu_matrix() - 1j * self.get_cappu_matrix()
        index, cols = self.buses.all_names(), ['A', 'B', 'C']
        if self._orient == 'cols':
            index, cols = cols, list(index)
        return pd.DataFrame(data=data, index=index, columns=cols, dtype=complex)

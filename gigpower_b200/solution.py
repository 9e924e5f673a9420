"""Base class of the drop-in solvers: constructor, ZIP class state, result
matrices and DataFrame accessors with the reference's names and shapes
(reference ``src/solution.py:20-315``)."""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import pandas as pd

from .circuit import Circuit
from .dss_reader import FeederModel, read_dss


class Solution:
    SLACKIDX = 0
    VSLACK = np.array([1, np.exp(1j * -120 * np.pi / 180), np.exp(1j * 120 * np.pi / 180)], dtype=complex)
    maxiter = 100

    SOLUTION_PARAMS = {
        'V': ['buses', ['A', 'B', 'C'], complex],
        'I': ['lines', ['A', 'B', 'C'], complex],
        'sV': ['buses', ['A', 'B', 'C'], complex],
        'Vmag': ['buses', ['A', 'B', 'C'], float],
        'Stx': ['lines', ['A', 'B', 'C'], complex],
        'Srx': ['lines', ['A', 'B', 'C'], complex],
    }

    @classmethod
    def set_zip_values(cls, zip_v):
        """Class-wide, like the reference (solution.py:40-53): also sets Circuit's."""
        Solution.ZIP_V = np.asarray(zip_v, dtype=float)
        Solution.aZ_p, Solution.aI_p, Solution.aPQ_p = Solution.ZIP_V[0:3]
        Solution.aZ_q, Solution.aI_q, Solution.aPQ_q = Solution.ZIP_V[3:6]
        Solution.min_voltage_pu = Solution.ZIP_V[6]
        Circuit._set_zip_values(zip_v)

    def __init__(self, dss_fp, zip_v=np.asarray([0.10, 0.05, 0.85, 0.10, 0.05, 0.85, 0.80]),
                 taps: Optional[Dict[str, int]] = None, device: Optional[int] = None):
        """``dss_fp``: path of a ``.dss`` master file (or a feeder-model ``.json``,
        or a FeederModel).  ``taps``: regulator tap numbers by regcontrol name
        (OpenDSS would supply them, voltage_regulator.py:17).  ``device``: CUDA
        ordinal, default the current torch device."""
        self.set_zip_values(zip_v)
        self.model = dss_fp if isinstance(dss_fp, FeederModel) else read_dss(dss_fp, taps)
        self.circuit = Circuit(self.model)
        self._orient = 'rows'
        self.circuit._orient = self._orient
        self.device = device
        self.iterations = 0
        self.solution_tolerance = -1
        self.convergence_diff = -1
        self.converged = False
        self._init_solution_matrices()
        self.Vtest = np.zeros(3, dtype='complex')
        self._set_sV_params()

    def _set_sV_params(self):
        c = self.circuit
        self.phase_matrix = c.buses.get_phase_matrix(self._orient)
        self.spu = c.get_spu_matrix()
        self.aPQ, self.aI, self.aZ = c.get_aPQ_matrix(), c.get_aI_matrix(), c.get_aZ_matrix()
        self.cappu, self.wpu = c.get_cappu_matrix(), c.get_wpu_matrix()

    def _init_solution_matrices(self):
        for param, (group, cols, dtype) in self.SOLUTION_PARAMS.items():
            rows = self.circuit.get_total_lines() if group == 'lines' \
                else getattr(self.circuit, group).num_elements
            setattr(self, param, np.zeros((rows, len(cols)), dtype=dtype))

    # ---- per-scenario regulator taps ---------------------------
    def gamma_rows_from_taps(self, taps, S, dev):
        """Per-scenario tap numbers -> the ``[n_gamma_rows, S]`` float64 device tensor behind
        ``GpSolveExtras.gamma_dev``.  ``taps``: ``[S, n_regulators]`` integers in the order of
        ``circuit.voltage_regulators`` (NumPy or torch), or a dict ``{regcontrol name: [S] or scalar}``
        (regulators left out keep the tap they were built with).  gamma = 0.9 + 0.2 (tap + 16) / 32
        with the reference's operation order (voltage_regulator.py:23-39), evaluated on the device."""
        import torch
        t = self.tables
        vrs = self.circuit.voltage_regulators.get_elements()
        if len(t.grow_vr) == 0:
            return None
        if isinstance(taps, dict):
            low = {str(k).lower(): v for k, v in taps.items()}
            unknown = set(low) - {vr.__name__ for vr in vrs}
            if unknown:
                raise KeyError(f'no such regulator(s): {sorted(unknown)}')
            cols = [torch.as_tensor(np.asarray(low.get(vr.__name__, vr.tap)), dtype=torch.float64).expand(S)
                    if np.ndim(low.get(vr.__name__, vr.tap)) == 0
                    else torch.as_tensor(np.asarray(low[vr.__name__]), dtype=torch.float64) for vr in vrs]
            taps = torch.stack(cols, dim=1)
        taps = torch.as_tensor(taps).to(device=dev, dtype=torch.float64)
        if taps.dim() != 2 or tuple(taps.shape) != (S, len(vrs)):
            raise ValueError(f'taps must be [S={S}, n_regulators={len(vrs)}], got {tuple(taps.shape)}')
        g = (taps + 16) / 32
        g = g * (1.10 - .90)
        g = g + .90
        rows = torch.as_tensor(np.asarray(t.grow_vr, dtype=np.int64), device=dev)
        return g.t().index_select(0, rows).contiguous()

    # ---- accessors (solution.py:127-175, 235-315) ------------------------------
    def get_data_frame(self, param: str, orient: str = '') -> pd.DataFrame:
        if not orient:
            orient = self._orient
        try:
            group, cols, dtype = self.SOLUTION_PARAMS[param]
        except KeyError:
            print(f"Not a valid solution parameter. Valid parameters: {self.SOLUTION_PARAMS.keys()}")
            return None
        index = list(getattr(self.circuit, group).all_names())
        data = getattr(self, param)
        if group == 'lines' and self.__class__.__name__ == 'SolutionFBS':
            data = data[0:self.circuit.lines.num_elements]
        if self.__class__.__name__ == 'SolutionNR3':
            data = data.transpose()
        if orient == 'cols':
            data = data.transpose()
            index, cols = list(cols), index
        return pd.DataFrame(data=data, index=index, columns=list(cols), dtype=dtype)

    def get_nominal_bus_powers(self, orient: str = ''):
        data = np.asarray(self.circuit.get_nominal_bus_powers()) * 1000
        index, cols = list(self.circuit.buses.all_names()), ['A', 'B', 'C']
        if self.__class__.__name__ == 'SolutionNR3':
            data = data.transpose()
        if orient == 'cols':
            data = data.transpose()
            index, cols = cols, index
        return pd.DataFrame(data=data, index=index, columns=cols, dtype=complex)

    @classmethod
    def get_params(cls) -> Dict:
        return {'slack_idx': cls.SLACKIDX, 'v_slack': cls.VSLACK, 'max_iter': cls.maxiter,
                'zip_v': Solution.ZIP_V}

    def _set_orient(self, orient: str):
        self._orient = orient
        self.circuit._orient = orient

    def get_V(self, orient=''):
        return self.get_data_frame('V', orient)

    def get_Vmag(self, orient=''):
        return self.get_data_frame('Vmag', orient)

    def get_I(self, orient=''):
        return self.get_data_frame('I', orient)

    def get_Stx(self, orient=''):
        return self.get_data_frame('Stx', orient)

    def get_Srx(self, orient=''):
        return self.get_data_frame('Srx', orient)

    def get_sV(self, orient=''):
        return self.get_data_frame('sV', orient)

    def print_solution(self):
        pd.options.display.float_format = '{:,.3f}'.format
        print("\nParameters:")
        for p, v in self.get_params().items():
            print(f'{p}: {v}')
        for title, getter in (('V', self.get_V), ('Vmag', self.get_Vmag), ('I', self.get_I),
                              ('Stx', self.get_Stx), ('Srx', self.get_Srx), ('sV', self.get_sV)):
            print('-' * 80)
            print(f"\n{title} solution")
            print(getter('rows'))
        print('-' * 80)

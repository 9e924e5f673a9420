"""FP64 tensor-core probe: DMMA alone, DFMA alone, both interleaved (gp_measure_dmma)."""
import ctypes as C
import json
import sys
sys.path.insert(0, '.')
from gigpower_b200 import _cabi
out = (C.c_double * 7)()
_cabi.check(_cabi.gp_measure_dmma(0, out))
peak = C.c_double()
_cabi.check(_cabi.gp_measure_fp64_peak(0, C.byref(peak)))
print(json.dumps({'dmma_tflops': out[0], 'dfma_tflops_8_chains': out[1], 'dmma_tflops_while_interleaved': out[2],
                  'dfma_tflops_while_interleaved': out[3], 'ms_dmma': out[4], 'ms_dfma': out[5], 'ms_both': out[6],
                  'dfma_peak_tflops_gp_measure_fp64_peak': peak.value}))

"""Stand-in for the third-party ``opendssdirect`` package -- TEST INFRASTRUCTURE.

gigpower imports ``opendssdirect`` at module import (reference
``src/solution.py:9``) and reads every input through it (call sites listed in
SURVEY.md section 8(b2)).  The real package (OpenDSSDirect.py >= 0.5.0,
``gigpower/requirements.txt:2``) is not installable here, so this package
answers exactly those calls from a :class:`gigpower_b200.dss_reader.FeederModel`.
It lets the UNMODIFIED reference run in this container to generate golden
vectors (``oracle/make_golden.py``).  It contains no solver arithmetic and is
never imported by the product.

Regulator taps are explicit: set ``opendssdirect.TAPS = {'reg1': 9, ...}``
before constructing a reference Solution.
"""
from . import _state
from . import (Circuit, Bus, Lines, Loads, Capacitors, Transformers,
               RegControls, CktElement, Solution)

TAPS = {}


def run_command(cmd: str):
    parts = cmd.strip().split(None, 1)
    if parts and parts[0].lower() in ('redirect', 'compile'):
        _state.load(parts[1].strip(), TAPS)
    return ''

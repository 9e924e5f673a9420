"""Shared state of the stand-in: the parsed feeder and the 'active' cursors."""
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
_repo = os.path.abspath(os.path.join(_here, '..', '..', '..'))
if _repo not in sys.path:
    sys.path.insert(0, _repo)

from gigpower_b200.dss_reader import read_dss  # noqa: E402  (input extraction only)

model = None
active_bus = None
active = {}          # class -> active record
active_elem = None   # last activated circuit element (Line/Load/Cap/Xfm record)
zipv = {}            # load name -> list


def load(path, taps):
    global model, active_bus, active, active_elem, zipv
    model = read_dss(path, taps)
    active_bus, active, active_elem, zipv = None, {}, None, {}


def find(records, name):
    name = name.lower()
    for r in records:
        if r.name == name:
            return r
    raise KeyError(name)

from . import _state as S


def BusNames():
    e = S.active_elem
    if hasattr(e, 'buses'):
        return list(e.buses)
    if hasattr(e, 'bus2'):
        return [e.bus1, e.bus2]
    return [e.bus1]


def NumPhases():
    return S.active_elem.nphases

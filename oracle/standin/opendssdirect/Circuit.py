from . import _state as S


def AllBusNames():
    return list(S.model.bus_names)


def SetActiveBus(name):
    S.active_bus = name.lower().split('.')[0]
    return S.model.bus_names.index(S.active_bus)

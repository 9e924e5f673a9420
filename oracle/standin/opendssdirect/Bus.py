from . import _state as S


def kVBase():
    return S.model.bus_kvbase[S.active_bus]


def Nodes():
    return list(S.model.bus_nodes[S.active_bus])


def Name():
    return S.active_bus

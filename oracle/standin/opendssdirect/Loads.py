from . import _state as S


def AllNames():
    return [r.name for r in S.model.loads]


def Name(*a):
    if a:
        S.active['load'] = S.active_elem = S.find(S.model.loads, a[0])
        return None
    return S.active['load'].name


def Model(*a):
    return 8


def ZipV(*a):
    if a:
        S.zipv[S.active['load'].name] = list(a[0])
        return None
    return S.zipv.get(S.active['load'].name, [])


def kW():
    return S.active['load'].kw


def kvar():
    return S.active['load'].kvar

from . import _state as S


def AllNames():
    return [r.name for r in S.model.capacitors]


def Name(*a):
    if a:
        S.active['cap'] = S.active_elem = S.find(S.model.capacitors, a[0])
        return None
    return S.active['cap'].name


def kvar():
    return S.active['cap'].kvar

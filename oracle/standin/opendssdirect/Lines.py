from . import _state as S


def AllNames():
    return [r.name for r in S.model.lines]


def Name(*a):
    if a:
        S.active['line'] = S.active_elem = S.find(S.model.lines, a[0])
        return None
    return S.active['line'].name


def Length():
    return S.active['line'].length


def Bus1():
    return S.active['line'].bus1


def Bus2():
    return S.active['line'].bus2


def IsSwitch():
    return S.active['line'].is_switch


def RMatrix():
    return list(S.active['line'].rmatrix)


def XMatrix():
    return list(S.active['line'].xmatrix)

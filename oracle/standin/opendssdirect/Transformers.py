from . import _state as S


def AllNames():
    return [r.name for r in S.model.transformers]


def Name(*a):
    if a:
        S.active['xfm'] = S.active_elem = S.find(S.model.transformers, a[0])
        return None
    return S.active['xfm'].name


def kV():
    return S.active['xfm'].kvs[-1]


def kVA():
    return S.active['xfm'].kvas[-1]

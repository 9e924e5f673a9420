from . import _state as S


def AllNames():
    return [r.name for r in S.model.regcontrols]


def Name(*a):
    if a:
        rc = S.find(S.model.regcontrols, a[0])
        S.active['reg'] = rc
        # activating a regcontrol makes its transformer the active element
        S.active['xfm'] = S.active_elem = S.find(S.model.transformers, rc.transformer)
        return None
    rc = S.active.get('reg')
    return rc.name if rc is not None else ''


def Transformer():
    return S.active['reg'].transformer


def TapNumber():
    return S.active['reg'].tap

"""Reader for the OpenDSS ``.dss`` subset that gigpower's feeders use.

gigpower never parses ``.dss`` itself: every constructor goes through the
third-party ``opendssdirect`` package (reference ``src/solution.py:71-78``,
call sites enumerated in SURVEY.md section 8(b2)).  This module answers the
same questions from the text of the file so that the drop-in classes have no
third-party dependency.  It is host-side input extraction only; nothing in
here is on the solve path.

Supported: ``Clear``, ``New <class>.<name> k=v ...`` (also ``New
object=<class>.<name>``), ``~`` continuation lines, ``Redirect``/``Compile``,
``Set VoltageBases=...``; classes circuit, linecode, line, load, capacitor,
transformer, regcontrol; comments ``!``, ``//`` and ``/* ... */``.  Everything
else (``Solve``, ``CalcVoltageBases``, ``BusCoords``, ``Set Controlmode`` and
property edits such as ``Load.x.vminpu=.85``) does not change what gigpower
reads and is skipped.

Regulator tap positions are an explicit input (``taps=``): OpenDSS chooses
them in its control loop and they cannot be derived from the file (reference
``src/voltage_regulator.py:15-17``).
"""
from __future__ import annotations

import json
import math
import os
import re
from dataclasses import dataclass, field, asdict
from typing import Dict, List, Optional, Sequence

# length of one unit, in metres (OpenDSS LineUnits)
_UNIT_M = {
    "mi": 1609.344, "kft": 304.8, "km": 1000.0, "m": 1.0, "ft": 0.3048,
    "in": 0.0254, "cm": 0.01, "mm": 0.001,
}
SQRT3 = math.sqrt(3.0)


@dataclass
class LineRec:
    name: str
    bus1: str            # spec as written, with .n suffixes, lower-case
    bus2: str
    nphases: int
    length: float        # in the line's own length unit
    is_switch: bool
    rmatrix: List[float]  # flat nphases^2, ohm per length unit of the line
    xmatrix: List[float]


@dataclass
class LoadRec:
    name: str
    bus1: str
    nphases: int
    kw: float
    kvar: float


@dataclass
class CapRec:
    name: str
    bus1: str
    nphases: int
    kvar: float


@dataclass
class XfmRec:
    name: str
    buses: List[str]
    kvs: List[float]
    kvas: List[float]
    nphases: int


@dataclass
class RegRec:
    name: str
    transformer: str
    tap: int = 0


@dataclass
class FeederModel:
    """Everything gigpower reads from OpenDSS, in OpenDSS index order."""
    name: str = ""
    source_bus: str = "sourcebus"
    base_kv: float = 115.0
    voltage_bases: List[float] = field(default_factory=list)
    bus_names: List[str] = field(default_factory=list)
    bus_nodes: Dict[str, List[int]] = field(default_factory=dict)
    bus_kvbase: Dict[str, float] = field(default_factory=dict)   # line-to-neutral kV
    lines: List[LineRec] = field(default_factory=list)
    loads: List[LoadRec] = field(default_factory=list)
    capacitors: List[CapRec] = field(default_factory=list)
    transformers: List[XfmRec] = field(default_factory=list)
    regcontrols: List[RegRec] = field(default_factory=list)

    # ---- serialisation (feeder fixtures travel as JSON) -------------------
    def to_json(self) -> str:
        return json.dumps(asdict(self), indent=None, separators=(",", ":"))

    @staticmethod
    def from_json(text: str) -> "FeederModel":
        d = json.loads(text)
        m = FeederModel(
            name=d["name"], source_bus=d["source_bus"], base_kv=d["base_kv"],
            voltage_bases=list(d["voltage_bases"]), bus_names=list(d["bus_names"]),
            bus_nodes={k: list(v) for k, v in d["bus_nodes"].items()},
            bus_kvbase=dict(d["bus_kvbase"]))
        m.lines = [LineRec(**x) for x in d["lines"]]
        m.loads = [LoadRec(**x) for x in d["loads"]]
        m.capacitors = [CapRec(**x) for x in d["capacitors"]]
        m.transformers = [XfmRec(**x) for x in d["transformers"]]
        m.regcontrols = [RegRec(**x) for x in d["regcontrols"]]
        return m


# --------------------------------------------------------------------------
# tokenising
# --------------------------------------------------------------------------
def _strip_comments(text: str) -> List[str]:
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    out = []
    for raw in text.splitlines():
        line = raw
        for mark in ("!", "//"):
            k = line.find(mark)
            if k >= 0:
                line = line[:k]
        line = line.strip()
        if line:
            out.append(line)
    return out


_OPEN = {"(": ")", "[": "]", '"': '"', "'": "'", "{": "}"}


def _split_tokens(s: str) -> List[str]:
    """Split a command on blanks/commas, keeping bracketed/quoted groups and
    ``key = value`` (blanks around ``=``) together."""
    s = re.sub(r"\s*=\s*", "=", s)
    toks, cur, close = [], [], None
    for ch in s:
        if close is not None:
            cur.append(ch)
            if ch == close:
                close = None
        elif ch in _OPEN:
            close = _OPEN[ch]
            cur.append(ch)
        elif ch.isspace() or ch == ",":
            if cur:
                toks.append("".join(cur))
                cur = []
        else:
            cur.append(ch)
    if cur:
        toks.append("".join(cur))
    return toks


def _unwrap(v: str) -> str:
    v = v.strip()
    while len(v) >= 2 and v[0] in _OPEN and v[-1] == _OPEN[v[0]]:
        v = v[1:-1].strip()
    return v


def _rpn(v: str) -> float:
    """OpenDSS in-line RPN, e.g. ``(8 1000 /)``."""
    stack: List[float] = []
    for t in _unwrap(v).split():
        if t in "+-*/" and len(t) == 1:
            b, a = stack.pop(), stack.pop()
            stack.append({"+": a + b, "-": a - b, "*": a * b, "/": a / b}[t])
        else:
            stack.append(float(t))
    return stack[-1]


def _num(v: str) -> float:
    u = _unwrap(v)
    try:
        return float(u)
    except ValueError:
        return _rpn(u)


def _numlist(v: str) -> List[float]:
    return [float(t) for t in re.split(r"[\s,|]+", _unwrap(v)) if t]


def _strlist(v: str) -> List[str]:
    return [t.lower() for t in re.split(r"[\s,]+", _unwrap(v)) if t]


def _lower_tri(v: str, n: int) -> List[List[float]]:
    """``(a | b c | d e f)`` lower-triangular (or full) rows -> symmetric n x n."""
    rows = [[float(t) for t in re.split(r"[\s,]+", r.strip()) if t]
            for r in _unwrap(v).split("|")]
    m = [[0.0] * n for _ in range(n)]
    for i, r in enumerate(rows[:n]):
        for j, x in enumerate(r[:n]):
            m[i][j] = x
            if len(r) <= i + 1:      # triangular row: mirror
                m[j][i] = x
    return m


def _yes(v: str) -> bool:
    return _unwrap(v).lower() in ("y", "yes", "true", "t")


# --------------------------------------------------------------------------
# the reader
# --------------------------------------------------------------------------
class _Reader:
    def __init__(self):
        self.model = FeederModel()
        self.linecodes: Dict[str, dict] = {}
        self.objs: List[tuple] = []        # (class, name, props-in-order) as defined
        self.active: Optional[list] = None

    # -- file level ----------------------------------------------------------
    def read_file(self, path: str):
        with open(path, "r", errors="replace") as f:
            text = f.read()
        base = os.path.dirname(os.path.abspath(path))
        for cmd in _strip_comments(text):
            self.command(cmd, base)

    def command(self, cmd: str, base: str):
        toks = _split_tokens(cmd)
        if not toks:
            return
        verb = toks[0].lower()
        if verb == "~" or verb.startswith("~") or verb == "more":
            rest = toks[1:] if verb in ("~", "more") else [toks[0][1:]] + toks[1:]
            if self.active is not None:
                self.active.extend(self._props(rest))
        elif verb == "new":
            self._new(toks[1:])
        elif verb in ("redirect", "compile"):
            target = _unwrap(" ".join(toks[1:]))
            self.read_file(target if os.path.isabs(target) else os.path.join(base, target))
        elif verb == "set":
            for key, val in self._props(toks[1:]):
                if key == "voltagebases":
                    self.model.voltage_bases = _numlist(val)
        else:
            self.active = None if verb == "clear" else self.active
        # everything else (solve, calcv, buscoords, edits) is irrelevant here

    @staticmethod
    def _props(toks: Sequence[str]) -> List[tuple]:
        out = []
        for t in toks:
            if "=" in t:
                k, v = t.split("=", 1)
                out.append((k.strip().lower(), v.strip()))
            elif t:
                out.append(("", t))
        return out

    def _new(self, toks: List[str]):
        if not toks:
            return
        first = toks[0]
        if first.lower().startswith("object="):
            first = first.split("=", 1)[1]
        if "." not in first:
            self.active = None
            return
        cls, name = first.split(".", 1)
        props = self._props(toks[1:])
        self.objs.append((cls.lower(), name.lower(), props))
        self.active = props

    # -- building ------------------------------------------------------------
    def build(self, taps: Optional[Dict[str, int]]) -> FeederModel:
        m = self.model
        taps = {k.lower(): int(v) for k, v in (taps or {}).items()}
        order: List[str] = []
        nodes: Dict[str, List[int]] = {}

        def touch(spec: str, nph: int):
            spec = spec.lower()
            parts = spec.split(".")
            bus = parts[0]
            ns = [int(p) for p in parts[1:]] if len(parts) > 1 else list(range(1, nph + 1))
            if bus not in nodes:
                nodes[bus] = []
                order.append(bus)
            for n in ns:
                if n != 0 and n not in nodes[bus]:
                    nodes[bus].append(n)

        # linecodes first: lines may refer to codes defined later in the file
        for cls, name, props in self.objs:
            if cls == "linecode":
                self.linecodes[name] = self._linecode(props)

        for cls, name, props in self.objs:
            p = {}
            for k, v in props:
                p[k] = v            # last assignment wins, except wdg groups (below)
            if cls == "circuit":
                m.name = name
                m.base_kv = _num(p.get("basekv", "115"))
                m.source_bus = _unwrap(p.get("bus1", "sourcebus")).lower().split(".")[0]
                touch(m.source_bus, 3)
            elif cls == "line":
                rec = self._line(name, props)
                touch(rec.bus1, rec.nphases)
                touch(rec.bus2, rec.nphases)
                m.lines.append(rec)
            elif cls == "load":
                nph = int(_num(p.get("phases", "3")))
                rec = LoadRec(name, _unwrap(p["bus1"]).lower(), nph,
                              _num(p.get("kw", "10")), _num(p.get("kvar", "5")))
                touch(rec.bus1, nph)
                m.loads.append(rec)
            elif cls == "capacitor":
                nph = int(_num(p.get("phases", "3")))
                rec = CapRec(name, _unwrap(p["bus1"]).lower(), nph, _num(p.get("kvar", "1200")))
                touch(rec.bus1, nph)
                m.capacitors.append(rec)
            elif cls == "transformer":
                rec = self._transformer(name, props)
                for b in rec.buses:
                    touch(b, rec.nphases)
                m.transformers.append(rec)
            elif cls == "regcontrol":
                tr = _unwrap(p.get("transformer", name)).lower()
                m.regcontrols.append(RegRec(name, tr, taps.get(name, 0)))
        if m.source_bus not in nodes:
            touch(m.source_bus, 3)
        # source bus is always index 0 (the Vsource is the first circuit element)
        order.remove(m.source_bus)
        m.bus_names = [m.source_bus] + order
        m.bus_nodes = {b: nodes[b] for b in m.bus_names}
        self._voltage_bases()
        return m

    def _linecode(self, props) -> dict:
        p = dict(props)
        n = int(_num(p.get("nphases", "3")))
        code = {"n": n, "units": _unwrap(p.get("units", "none")).lower()}
        if "rmatrix" in p:
            code["r"] = _lower_tri(p["rmatrix"], n)
            code["x"] = _lower_tri(p["xmatrix"], n)
        else:
            code["seq"] = tuple(_num(p.get(k, d)) for k, d in
                                (("r1", ".058"), ("x1", ".1206"), ("r0", ".1784"), ("x0", ".4047")))
        return code

    @staticmethod
    def _seq_matrix(r1, x1, r0, x0, n):
        rs, rm = (2 * r1 + r0) / 3.0, (r0 - r1) / 3.0
        xs, xm = (2 * x1 + x0) / 3.0, (x0 - x1) / 3.0
        r = [[rs if i == j else rm for j in range(n)] for i in range(n)]
        x = [[xs if i == j else xm for j in range(n)] for i in range(n)]
        return r, x

    def _line(self, name: str, props) -> LineRec:
        p = dict(props)
        nph = int(_num(p.get("phases", "3")))
        is_switch = _yes(p.get("switch", "n"))
        units = _unwrap(p.get("units", "none")).lower()
        length = _num(p["length"]) if "length" in p else 1.0
        scale = 1.0
        if "linecode" in p:
            code = self.linecodes[_unwrap(p["linecode"]).lower()]
            nph = code["n"] if "phases" not in p else nph
            if "r" in code:
                r, x = code["r"], code["x"]
            else:
                r, x = self._seq_matrix(*code["seq"], nph)
            cu = code["units"]
            if cu != "none" and units != "none" and cu != units:
                # ohm per code-unit -> ohm per line-unit
                scale = _UNIT_M[units] / _UNIT_M[cu]
        elif "rmatrix" in p:
            r, x = _lower_tri(p["rmatrix"], nph), _lower_tri(p["xmatrix"], nph)
        else:
            seq = [_num(p.get(k, d)) for k, d in
                   (("r1", ".058"), ("x1", ".1206"), ("r0", ".1784"), ("x0", ".4047"))]
            r, x = self._seq_matrix(*seq, nph)
        if is_switch:
            length = 0.001
        rflat = [r[i][j] * scale for i in range(nph) for j in range(nph)]
        xflat = [x[i][j] * scale for i in range(nph) for j in range(nph)]
        return LineRec(name, _unwrap(p["bus1"]).lower(), _unwrap(p["bus2"]).lower(),
                       nph, length, is_switch, rflat, xflat)

    @staticmethod
    def _transformer(name: str, props) -> XfmRec:
        nph, buses, kvs, kvas = 3, {}, {}, {}
        w = 1
        for k, v in props:
            if k == "phases":
                nph = int(_num(v))
            elif k == "wdg":
                w = int(_num(v))
            elif k == "bus":
                buses[w] = _unwrap(v).lower()
            elif k == "kv":
                kvs[w] = _num(v)
            elif k == "kva":
                kvas[w] = _num(v)
            elif k == "buses":
                for i, b in enumerate(_strlist(v)):
                    buses[i + 1] = b
            elif k == "kvs":
                for i, x in enumerate(_numlist(v)):
                    kvs[i + 1] = x
            elif k == "kvas":
                for i, x in enumerate(_numlist(v)):
                    kvas[i + 1] = x
        ws = sorted(buses)
        return XfmRec(name, [buses[i] for i in ws], [kvs.get(i, 12.47) for i in ws],
                      [kvas.get(i, 1000.0) for i in ws], nph)

    def _voltage_bases(self):
        """kVBase per bus = LN kV of its voltage zone: the source's zone is the
        circuit basekv, a zone is constant along lines and changes across a
        transformer to the winding kV (1-phase windings are rated LN); each zone
        is snapped to the nearest ``Set VoltageBases`` entry."""
        m = self.model
        adj: Dict[str, List[tuple]] = {}
        for ln in m.lines:
            a, b = ln.bus1.split(".")[0], ln.bus2.split(".")[0]
            adj.setdefault(a, []).append((b, None))
            adj.setdefault(b, []).append((a, None))
        for t in m.transformers:
            bs = [b.split(".")[0] for b in t.buses]
            kv = [k * (SQRT3 if t.nphases == 1 else 1.0) for k in t.kvs]
            for i in range(len(bs)):
                for j in range(len(bs)):
                    if i != j:
                        adj.setdefault(bs[i], []).append((bs[j], kv[j]))
        kvll = {m.source_bus: m.base_kv}
        stack = [m.source_bus]
        while stack:
            u = stack.pop()
            for v, kv in adj.get(u, []):
                if v not in kvll:
                    kvll[v] = kvll[u] if kv is None else kv
                    stack.append(v)
        for b in m.bus_names:
            kv = kvll.get(b, m.base_kv)
            if m.voltage_bases:
                kv = min(m.voltage_bases, key=lambda vb: abs(vb - kv))
            m.bus_kvbase[b] = kv / SQRT3


def read_dss(path: str, taps: Optional[Dict[str, int]] = None) -> FeederModel:
    """Parse a ``.dss`` master file (or a feeder-model ``.json`` written by
    :meth:`FeederModel.to_json`) into a :class:`FeederModel`."""
    path = str(path)
    if path.lower().endswith(".json"):
        with open(path) as f:
            model = FeederModel.from_json(f.read())
        if taps:
            low = {k.lower(): int(v) for k, v in taps.items()}
            for rc in model.regcontrols:
                rc.tap = low.get(rc.name, rc.tap)
        return model
    rd = _Reader()
    rd.read_file(path)
    return rd.build(taps)

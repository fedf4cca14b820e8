"""Parser for TF2 network-config headers (`<net>.h`).

The reference selects a network at compile time by including a generated header
(`Runtime_Engine/cnn/host/inc/cnn.h:29-35`) that defines per-layer constant tables named
`k*[NUM_CONVOLUTIONS]` (`Runtime_Engine/cnn/host/inc/resnet50.h:119-1368`).  This module reads
such a header as *data* so a TransForm_Kit / TF2_auto_config emitted header drops in unchanged:
it evaluates the `#define` constants and the brace initialisers (which may use the helper macros
of `defines.h:45-49`: CEIL, NEXT_DIVISIBLE, NEXT_POWER_OF_2, MYMAX2) with the FPGA vector widths
of `archs.h:25-43` and returns plain Python lists.
"""
from __future__ import annotations

import re
from typing import Dict, List, Union

# FPGA architecture vector widths (archs.h:25-43); only needed to evaluate table initialisers
ARCH = {
    "IMAGE_BATCH_SIZE": 1, "N_VECTOR": 16, "C_VECTOR": 16, "OW_VECTOR": 5, "FW_VECTOR": 3,
    "NARROW_N_VECTOR": 16, "W_VECTOR": 7, "DOUBLE_BUFFER_DIM": 2, "NUM_IMAGES": 1,
}


def _ceil(x, y):
    return (x - 1) // y + 1


def _next_divisible(x, y):
    return x if x % y == 0 else x + y - x % y


def _next_pow2(x):
    v = x - 1
    for s in (1, 2, 4, 8, 16):
        v |= v >> s
    return v + 1


_FUNCS = {"CEIL": _ceil, "NEXT_DIVISIBLE": _next_divisible, "NEXT_POWER_OF_2": _next_pow2,
          "MYMAX2": max, "true": 1, "false": 0}

_COMMENT_BLOCK = re.compile(r"/\*.*?\*/", re.S)
_COMMENT_LINE = re.compile(r"//[^\n]*")
_DEFINE = re.compile(r"^[ \t]*#define[ \t]+([A-Za-z_]\w*)[ \t]+(.+?)[ \t]*$", re.M)
_TABLE = re.compile(r"CONSTANT\s+(bool|int|char|short)\s+(\w+)\s*\[\s*([^\]]*)\]\s*=\s*\{(.*?)\}\s*;", re.S)
_SCALAR = re.compile(r"CONSTANT\s+(bool|int|char|short)\s+(\w+)\s*=\s*([^;{]+);")


def _py_expr(expr: str) -> str:
    # C integer division / ternaries appear only inside the helper macros we re-implement;
    # plain "/" in table initialisers is integer division.
    expr = expr.replace("/", "//")
    expr = re.sub(r"(\d+)[uUlL]+\b", r"\1", expr)
    return expr


class HeaderTables(dict):
    """dict name -> int | list[int]; scalars from #define / CONSTANT scalars, tables as lists."""


def parse_header(text: str) -> HeaderTables:
    text = _COMMENT_BLOCK.sub(" ", text)
    text = _COMMENT_LINE.sub(" ", text)
    text = text.replace("\\\n", " ")
    env: Dict[str, Union[int, List[int]]] = dict(ARCH)
    # defines.h:86-87 (pool edge offsets used by a few initialisers)
    raw_defines: Dict[str, str] = {"POOL_OFFSET_P": "(POOL_WINDOW_MAX-1)", "POOL_OFFSET_Q": "(POOL_WINDOW_MAX-1)"}
    for m in _DEFINE.finditer(text):
        name, val = m.group(1), m.group(2)
        if "(" in name:
            continue
        raw_defines[name] = raw_defines.get(name, val) if name not in ("POOL_OFFSET_P", "POOL_OFFSET_Q") else val

    def ev(expr: str):
        scope = dict(_FUNCS)
        scope.update({k: v for k, v in env.items() if isinstance(v, int)})
        return eval(_py_expr(expr), {"__builtins__": {}}, _Lazy(scope, raw_defines, ev))

    out = HeaderTables()
    for name, val in raw_defines.items():
        try:
            v = ev(val)
        except Exception:
            continue
        if isinstance(v, (int, bool)):
            env[name] = int(v)
            out[name] = int(v)
    for m in _TABLE.finditer(text):
        name, body = m.group(2), m.group(4)
        items = _split_top_level(body.replace("\n", " "))
        vals = [int(ev(s)) for s in items if s]
        out[name] = vals
        env[name] = vals
    for m in _SCALAR.finditer(text):
        name, val = m.group(2), m.group(3)
        try:
            out[name] = int(ev(val))
        except Exception:
            pass
    return out


def _split_top_level(body: str) -> List[str]:
    items, depth, cur = [], 0, []
    for ch in body:
        if ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
        if ch == "," and depth == 0:
            items.append("".join(cur).strip())
            cur = []
        else:
            cur.append(ch)
    items.append("".join(cur).strip())
    return items


class _Lazy(dict):
    """Name lookup that evaluates not-yet-evaluated #defines on demand."""

    def __init__(self, scope, raw, ev):
        super().__init__(scope)
        self._raw, self._ev, self._busy = raw, ev, set()

    def __missing__(self, key):
        if key in self._raw and key not in self._busy:
            self._busy.add(key)
            try:
                v = self._ev(self._raw[key])
            finally:
                self._busy.discard(key)
            self[key] = v
            return v
        raise KeyError(key)


def parse_header_file(path: str) -> HeaderTables:
    with open(path, "r", errors="replace") as f:
        return parse_header(f.read())

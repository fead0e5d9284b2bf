"""oracle/onnx_min.py — minimal ONNX protobuf reader (TEST INFRASTRUCTURE ONLY).

Part of the CPU oracle: nothing in the product path (fastdet_b200/) may import this file.  It is an
independent second reader of the same ``.onnx`` bytes the native loader (csrc/onnx_reader.cc) parses;
field numbers follow the public onnx.proto (ModelProto.graph=7, GraphProto.node=1/initializer=5/
input=11/output=12, NodeProto.input=1/output=2/op_type=4/attribute=5, TensorProto.dims=1/data_type=2/
float_data=4/int32_data=5/int64_data=7/name=8/raw_data=9/double_data=10).  The reference itself hands the
file to onnxruntime (reference server/detector.py:118), which is not installed here.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from typing import Dict, List, Tuple

import numpy as np


def _read_varint(b: bytes, i: int) -> Tuple[int, int]:
    v = 0
    shift = 0
    while True:
        c = b[i]
        i += 1
        v |= (c & 0x7F) << shift
        if not (c & 0x80):
            return v, i
        shift += 7


def _fields(b: bytes):
    i = 0
    n = len(b)
    while i < n:
        key, i = _read_varint(b, i)
        fno, wt = key >> 3, key & 7
        if wt == 0:
            v, i = _read_varint(b, i)
        elif wt == 1:
            v = b[i:i + 8]
            i += 8
        elif wt == 2:
            ln, i = _read_varint(b, i)
            v = b[i:i + ln]
            i += ln
        elif wt == 5:
            v = b[i:i + 4]
            i += 4
        else:
            raise ValueError(f"unsupported wire type {wt}")
        yield fno, wt, v


def _signed(v: int) -> int:
    return v - (1 << 64) if v >= (1 << 63) else v


def _packed_varints(b: bytes) -> List[int]:
    out = []
    i = 0
    while i < len(b):
        v, i = _read_varint(b, i)
        out.append(_signed(v))
    return out


_DTYPES = {1: np.float32, 6: np.int32, 7: np.int64, 10: np.float16, 11: np.float64}


def parse_tensor(b: bytes) -> Tuple[str, np.ndarray]:
    dims: List[int] = []
    dtype = 1
    name = ""
    raw = None
    floats: List[float] = []
    ints: List[int] = []
    doubles: List[float] = []
    for fno, wt, v in _fields(b):
        if fno == 1:
            dims.extend(_packed_varints(v) if wt == 2 else [_signed(v)])
        elif fno == 2:
            dtype = v
        elif fno == 4:
            floats.extend(struct.unpack(f"<{len(v) // 4}f", v))
        elif fno in (5, 7):
            ints.extend(_packed_varints(v) if wt == 2 else [_signed(v)])
        elif fno == 8:
            name = v.decode()
        elif fno == 9:
            raw = bytes(v)
        elif fno == 10:
            doubles.extend(struct.unpack(f"<{len(v) // 8}d", v))
    np_dtype = _DTYPES[dtype]
    if raw is not None:
        arr = np.frombuffer(raw, dtype=np.dtype(np_dtype).newbyteorder("<")).astype(np_dtype)
    elif dtype == 1:
        arr = np.array(floats, np.float32)
    elif dtype == 11:
        arr = np.array(doubles, np.float64)
    elif dtype == 10:
        arr = np.array(ints, np.uint16).view(np.float16)
    else:
        arr = np.array(ints, np_dtype)
    return name, arr.reshape(dims)


@dataclass
class Node:
    op: str
    name: str
    inputs: List[str]
    outputs: List[str]
    attrs: Dict[str, object] = field(default_factory=dict)


def _parse_attr(b: bytes):
    name = ""
    val = None
    floats: List[float] = []
    ints: List[int] = []
    has_list = None
    atype = 0
    f = i = s = t = None
    for fno, wt, v in _fields(b):
        if fno == 1:
            name = v.decode()
        elif fno == 2:
            f = struct.unpack("<f", v)[0]
        elif fno == 3:
            i = _signed(v)
        elif fno == 4:
            s = v.decode()
        elif fno == 5:
            t = parse_tensor(v)[1]
        elif fno == 7:
            floats.extend(struct.unpack(f"<{len(v) // 4}f", v))
            has_list = "f"
        elif fno == 8:
            ints.extend(_packed_varints(v) if wt == 2 else [_signed(v)])
            has_list = "i"
        elif fno == 20:
            atype = v
    if atype == 1 or (atype == 0 and f is not None):
        val = f
    elif atype == 2 or (atype == 0 and i is not None):
        val = i
    elif atype == 3 or (atype == 0 and s is not None):
        val = s
    elif atype == 4 or (atype == 0 and t is not None):
        val = t
    elif atype == 6 or has_list == "f":
        val = floats
    elif atype == 7 or has_list == "i":
        val = ints
    return name, val


def _parse_node(b: bytes) -> Node:
    n = Node("", "", [], [])
    for fno, wt, v in _fields(b):
        if fno == 1:
            n.inputs.append(v.decode())
        elif fno == 2:
            n.outputs.append(v.decode())
        elif fno == 3:
            n.name = v.decode()
        elif fno == 4:
            n.op = v.decode()
        elif fno == 5:
            k, val = _parse_attr(v)
            n.attrs[k] = val
    return n


def _value_info_name(b: bytes) -> str:
    for fno, wt, v in _fields(b):
        if fno == 1:
            return v.decode()
    return ""


@dataclass
class Graph:
    nodes: List[Node]
    initializers: Dict[str, np.ndarray]
    inputs: List[str]
    outputs: List[str]
    opset: int = 0


def load(data: bytes) -> Graph:
    graph_bytes = None
    opset = 0
    for fno, wt, v in _fields(data):
        if fno == 7:
            graph_bytes = v
        elif fno == 8:
            dom, ver = "", 0
            for f2, w2, v2 in _fields(v):
                if f2 == 1:
                    dom = v2.decode()
                elif f2 == 2:
                    ver = v2
            if dom in ("", "ai.onnx"):
                opset = ver
    if graph_bytes is None:
        raise ValueError("no graph in model")
    g = Graph([], {}, [], [], opset)
    for fno, wt, v in _fields(graph_bytes):
        if fno == 1:
            g.nodes.append(_parse_node(v))
        elif fno == 5:
            name, arr = parse_tensor(v)
            g.initializers[name] = arr
        elif fno == 11:
            g.inputs.append(_value_info_name(v))
        elif fno == 12:
            g.outputs.append(_value_info_name(v))
    g.inputs = [n for n in g.inputs if n not in g.initializers]
    return g

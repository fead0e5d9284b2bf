"""oracle/ref_graph.py — CPU fp32/fp64 (and bf16-operand) executor for the ONNX operator subset of the YOLOv3 graphs.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
legs, never by the product path.

This restates, with ONNX-specification semantics, what ``self.model.run(None, {'input': a})`` does at
reference server/detector.py:135.  The arithmetic there lives in a third-party dependency that is
neither vendored nor version-pinned (``onnxruntime-gpu``, reference requirements.txt:3) and is not
installable in this image, and the reference holds no test vectors for it: PARITY UNPINNED at the
onnxruntime boundary.  What pins this file instead: (i) tests/test_oracle_graph.py checks it against
torch eager modules exported to ONNX by torch's own serializer (an independent writer + an independent
implementation of the same operators); (ii) it reads the same ``.onnx`` bytes the native loader reads.

Operators (ONNX opset 9-13 forms): Conv, BatchNormalization (inference), LeakyRelu, Relu, Add, Mul, MaxPool,
Resize / Upsample (nearest, asymmetric, floor), Concat, Pad (constant), Constant, Identity and the integer
shape-arithmetic ops torch's exporter emits around Resize/Pad (Shape, Gather, Cast, Slice, Unsqueeze,
Squeeze, Floor, Div, Sub, ConstantOfShape, Reshape, Transpose).

``dtype="bf16"`` evaluates the arithmetic BASELINE.json's north_star prescribes for the new implementation —
bf16 operands, fp32 accumulation — on the CPU: BatchNormalization is folded into the preceding Conv in fp32
(w' = w * gamma / sqrt(var + eps), b' = beta + (b - mean) * gamma / sqrt(var + eps)), the folded weights and every
convolution input are rounded to bf16 (round-to-nearest-even), products are summed in fp32, and every stored
activation (the value after LeakyRelu, and after a residual Add) is rounded to bf16; graph outputs stay fp32.
It separates the two sources of difference from the fp32 oracle: what the prescribed operand precision costs
(bf16 oracle vs fp32 oracle, a CPU-only comparison: tests/test_oracle_graph.py::test_bf16_operand_floor) and what
the CUDA kernels add on top (GPU vs bf16 oracle: fp32 summation order only).
"""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn.functional as F

from . import onnx_min


def _bf16(t: torch.Tensor) -> torch.Tensor:
    """Round-to-nearest-even to bf16, kept in fp32 storage."""
    return t.to(torch.bfloat16).to(torch.float32)


def _t(x, dtype=None):
    if isinstance(x, torch.Tensor):
        return x
    a = np.asarray(x)
    t = torch.from_numpy(a.copy())
    return t


class GraphExecutor:
    """Runs a parsed graph on torch-CPU.  ``dtype`` = torch.float32 (what ORT computes in), float64, or "bf16"
    (bf16 operands / fp32 accumulation, see the module docstring)."""

    def __init__(self, onnx_bytes: bytes, dtype=torch.float32, accumulate=torch.float32, fp32_tail: int = 0):
        """accumulate (bf16 mode): dtype the products are summed in — float64 gives the exactly-summed variant of the same
        bf16-operand arithmetic.  fp32_tail (bf16 mode): the last `fp32_tail` convolutions in front of every graph output
        keep fp32 weights and pass fp32 (unrounded) activations between them."""
        self.graph = onnx_min.load(onnx_bytes)
        self.bf16 = dtype == "bf16"
        self.accumulate = accumulate
        self.fp32_tail = int(fp32_tail)
        if self.bf16:
            dtype = torch.float32
        self.dtype = dtype
        self.consts: Dict[str, torch.Tensor] = {}
        for k, v in self.graph.initializers.items():
            t = torch.from_numpy(np.ascontiguousarray(v))
            if t.is_floating_point():
                t = t.to(dtype)
            self.consts[k] = t
        if "input" not in self.graph.inputs:
            # reference server/detector.py:135 feeds {'input': a}; any other name makes ORT raise
            raise KeyError(f"graph has no input named 'input' (inputs: {self.graph.inputs})")
        self.folded: Dict[str, tuple] = {}   # Conv output name -> (bf16-rounded folded weight, fp32 bias)
        self.bn_identity = set()             # BatchNormalization nodes folded into their producer
        self.round_after = set()             # node outputs that are stored activations (rounded to bf16)
        if self.bf16:
            self._prepare_bf16()

    def _prepare_bf16(self):
        g = self.graph
        consumers: Dict[str, list] = {}
        for n in g.nodes:
            for i in n.inputs:
                consumers.setdefault(i, []).append(n)
        outputs = set(g.outputs)

        def sole(name):
            c = consumers.get(name, [])
            return c[0] if len(c) == 1 and name not in outputs else None

        # the last fp32_tail convolutions in front of every graph output (walking back through the fused chain)
        producer = {n.outputs[0]: n for n in g.nodes}
        tail = set()
        for o in g.outputs:
            name, left = o, self.fp32_tail
            while left > 0 and name in producer:
                n = producer[name]
                if n.op == "Conv":
                    tail.add(id(n))
                    left -= 1
                name = n.inputs[0]
        self.tail = tail
        for n in g.nodes:
            if n.op == "Conv":
                w = self.consts[n.inputs[1]].to(torch.float32)
                b = self.consts[n.inputs[2]].to(torch.float32) if len(n.inputs) > 2 and n.inputs[2] else torch.zeros(w.shape[0])
                last = n
                nx = sole(n.outputs[0])
                if nx is not None and nx.op == "BatchNormalization" and nx.inputs[0] == n.outputs[0]:
                    sc, bb, mean, var = (self.consts[k].to(torch.float32) for k in nx.inputs[1:5])
                    eps = torch.tensor(nx.attrs.get("epsilon", 1e-5), dtype=torch.float32)
                    inv = sc / torch.sqrt(var + eps)
                    w = w * inv.view(-1, 1, 1, 1)
                    b = bb + (b - mean) * inv
                    self.bn_identity.add(id(nx))
                    last = nx
                self.folded[n.outputs[0]] = (w if id(n) in tail else _bf16(w), b)
                nx = sole(last.outputs[0])
                if nx is not None and nx.op in ("LeakyRelu", "Relu"):
                    last = nx
                # the value after the activation is what the conv kernel's epilogue rounds and stores; a graph output
                # (head tensor) is written in fp32
                feeds_tail = any(id(c) in tail for c in consumers.get(last.outputs[0], []))
                if last.outputs[0] not in outputs and not (id(n) in tail and feeds_tail):
                    self.round_after.add(last.outputs[0])
            elif n.op == "Add":
                if n.inputs[0] not in self.consts and n.inputs[1] not in self.consts and n.outputs[0] not in outputs:
                    self.round_after.add(n.outputs[0])  # residual add: bf16 + bf16, rounded once

    # -- single-op semantics -------------------------------------------------
    def _conv(self, n, x, w, b=None):
        a = n.attrs
        if a.get("group", 1) != 1:
            raise NotImplementedError("grouped conv")
        k = a.get("kernel_shape", list(w.shape[2:]))
        strides = a.get("strides", [1, 1])
        dil = a.get("dilations", [1, 1])
        auto = a.get("auto_pad", "NOTSET")
        if auto in ("NOTSET", ""):
            pads = a.get("pads", [0, 0, 0, 0])
        elif auto == "VALID":
            pads = [0, 0, 0, 0]
        else:  # SAME_UPPER / SAME_LOWER
            pads = [0, 0, 0, 0]
            for i, (inp, kk, s) in enumerate(zip(x.shape[2:], k, strides)):
                out = -(-inp // s)
                total = max((out - 1) * s + kk - inp, 0)
                lo = total // 2 if auto == "SAME_UPPER" else total - total // 2
                pads[i], pads[i + 2] = lo, total - lo
        t, l, bt, r = pads
        if (t, l) == (bt, r):
            return F.conv2d(x, w, b, stride=strides, padding=(t, l), dilation=dil)
        x = F.pad(x, (l, r, t, bt))
        return F.conv2d(x, w, b, stride=strides, dilation=dil)

    def _maxpool(self, n, x):
        a = n.attrs
        k = a["kernel_shape"]
        strides = a.get("strides", [1, 1])
        pads = a.get("pads", [0, 0, 0, 0])
        if a.get("ceil_mode", 0):
            raise NotImplementedError("ceil_mode")
        t, l, bt, r = pads
        if any(pads):
            x = F.pad(x, (l, r, t, bt), value=float("-inf"))  # ONNX MaxPool pads never win the max
        return F.max_pool2d(x, kernel_size=k, stride=strides)

    def _resize(self, n, x, ins):
        mode = n.attrs.get("mode", "nearest")
        if mode != "nearest":
            raise NotImplementedError(f"Resize mode {mode}")
        scales = None
        sizes = None
        if n.op == "Upsample":
            scales = ins[1] if len(ins) > 1 else torch.tensor(n.attrs["scales"])
        else:
            if len(ins) == 2:  # opset 10
                scales = ins[1]
            else:
                if len(ins) > 2 and ins[2] is not None and ins[2].numel() > 0:
                    scales = ins[2]
                if len(ins) > 3 and ins[3] is not None and ins[3].numel() > 0:
                    sizes = ins[3]
            ctm = n.attrs.get("coordinate_transformation_mode", "half_pixel")
            nm = n.attrs.get("nearest_mode", "round_prefer_floor")
            if len(ins) > 2 and not (ctm == "asymmetric" and nm == "floor"):
                # for integer upscaling of 2 the common modes coincide except half_pixel+round_prefer_floor,
                # which also equals pixel repetition for scale 2; accept and treat as repetition
                pass
        if sizes is not None:
            oh, ow = int(sizes[2]), int(sizes[3])
        else:
            oh, ow = int(x.shape[2] * float(scales[2])), int(x.shape[3] * float(scales[3]))
        fy, fx = oh // x.shape[2], ow // x.shape[3]
        if fy * x.shape[2] != oh or fx * x.shape[3] != ow:
            raise NotImplementedError("non-integer resize")
        return x.repeat_interleave(fy, dim=2).repeat_interleave(fx, dim=3)

    def _pad(self, n, ins):
        x = ins[0]
        if len(ins) > 1:
            pads = [int(v) for v in ins[1]]
            value = float(ins[2]) if len(ins) > 2 and ins[2] is not None and ins[2].numel() else 0.0
        else:
            pads = n.attrs["pads"]
            value = n.attrs.get("value", 0.0)
        if n.attrs.get("mode", "constant") != "constant":
            raise NotImplementedError("Pad mode")
        r = len(pads) // 2
        tp = []
        for d in reversed(range(r)):
            tp += [pads[d], pads[d + r]]
        return F.pad(x, tp, value=value)

    # -- graph walk ----------------------------------------------------------
    def run(self, x: np.ndarray, keep: Optional[List[str]] = None, all_values: bool = False):
        """x: [N,3,H,W] float array.  Returns the graph outputs as float32 numpy arrays (graph order),
        or a dict name->array when `keep`/`all_values` is given."""
        vals: Dict[str, torch.Tensor] = dict(self.consts)
        vals["input"] = torch.from_numpy(np.ascontiguousarray(x)).to(self.dtype)
        vals[""] = None
        with torch.no_grad():
            for n in self.graph.nodes:
                ins = [vals.get(i) if i else None for i in n.inputs]
                op = n.op
                if op == "Conv" and self.bf16:
                    w16, bias = self.folded[n.outputs[0]]
                    xin = ins[0] if id(n) in self.tail else _bf16(ins[0])
                    if self.accumulate == torch.float64:
                        out = self._conv(n, xin.double(), w16.double(), bias.double()).to(torch.float32)
                    else:
                        out = self._conv(n, xin, w16, bias)
                elif op == "Conv":
                    out = self._conv(n, *ins)
                elif op == "BatchNormalization" and id(n) in self.bn_identity:
                    out = ins[0]
                elif op == "BatchNormalization":
                    xx, sc, bb, mean, var = ins
                    eps = n.attrs.get("epsilon", 1e-5)
                    shp = (1, -1, 1, 1)
                    out = sc.view(shp) * (xx - mean.view(shp)) / torch.sqrt(var.view(shp) + eps) + bb.view(shp)
                elif op == "LeakyRelu":
                    out = F.leaky_relu(ins[0], n.attrs.get("alpha", 0.01))
                elif op == "Relu":
                    out = F.relu(ins[0])
                elif op == "Add":
                    out = ins[0] + ins[1]
                elif op == "Sub":
                    out = ins[0] - ins[1]
                elif op == "Mul":
                    out = ins[0] * ins[1]
                elif op == "Div":
                    if not ins[0].is_floating_point() and not ins[1].is_floating_point():
                        out = torch.div(ins[0], ins[1], rounding_mode="trunc")
                    else:
                        out = ins[0] / ins[1]
                elif op == "Floor":
                    out = torch.floor(ins[0])
                elif op == "MaxPool":
                    out = self._maxpool(n, ins[0])
                elif op in ("Resize", "Upsample"):
                    out = self._resize(n, ins[0], ins)
                elif op == "Concat":
                    axis = n.attrs.get("axis", 1)
                    out = torch.cat([i for i in ins], dim=axis)
                elif op == "Pad":
                    out = self._pad(n, ins)
                elif op == "Constant":
                    v = n.attrs["value"]
                    out = torch.from_numpy(np.ascontiguousarray(v))
                    if out.is_floating_point():
                        out = out.to(self.dtype)
                elif op == "Identity":
                    out = ins[0]
                elif op == "Shape":
                    out = torch.tensor(list(ins[0].shape), dtype=torch.int64)
                elif op == "Gather":
                    axis = n.attrs.get("axis", 0)
                    idx = ins[1].to(torch.int64)
                    out = torch.index_select(ins[0], axis, idx.reshape(-1)).reshape(
                        list(ins[0].shape[:axis]) + list(idx.shape) + list(ins[0].shape[axis + 1:]))
                elif op == "Cast":
                    to = n.attrs["to"]
                    out = ins[0].to({1: self.dtype, 6: torch.int32, 7: torch.int64, 11: torch.float64}[to])
                elif op == "Slice":
                    if len(ins) > 1:
                        starts, ends = [int(v) for v in ins[1]], [int(v) for v in ins[2]]
                        axes = [int(v) for v in ins[3]] if len(ins) > 3 and ins[3] is not None else list(range(len(starts)))
                        steps = [int(v) for v in ins[4]] if len(ins) > 4 and ins[4] is not None else [1] * len(starts)
                    else:
                        starts, ends = n.attrs["starts"], n.attrs["ends"]
                        axes = n.attrs.get("axes", list(range(len(starts))))
                        steps = [1] * len(starts)
                    arr = ins[0].numpy()
                    index = [slice(None)] * arr.ndim
                    for s, e, ax, st in zip(starts, ends, axes, steps):
                        # numpy slicing clamps exactly like the ONNX Slice specification (INT64 extremes included)
                        index[ax] = slice(int(np.clip(s, -2**62, 2**62)), int(np.clip(e, -2**62, 2**62)), st)
                        if st < 0 and e < -arr.shape[ax]:
                            index[ax] = slice(int(np.clip(s, -2**62, 2**62)), None, st)
                    out = torch.from_numpy(np.ascontiguousarray(arr[tuple(index)]))
                elif op == "Unsqueeze":
                    axes = n.attrs.get("axes") or [int(v) for v in ins[1]]
                    out = ins[0]
                    for ax in sorted(axes):
                        out = out.unsqueeze(ax)
                elif op == "Squeeze":
                    axes = n.attrs.get("axes") or ([int(v) for v in ins[1]] if len(ins) > 1 else None)
                    out = ins[0]
                    if axes is None:
                        out = out.squeeze()
                    else:
                        for ax in sorted(axes, reverse=True):
                            out = out.squeeze(ax)
                elif op == "ConstantOfShape":
                    v = n.attrs.get("value")
                    fill = torch.from_numpy(np.ascontiguousarray(v)).reshape(-1)[0] if v is not None else torch.tensor(0.0)
                    out = torch.full([int(d) for d in ins[0]], fill.item(), dtype=fill.dtype)
                elif op == "Reshape":
                    out = ins[0].reshape([int(d) for d in ins[1]])
                elif op == "Transpose":
                    out = ins[0].permute(n.attrs["perm"])
                else:
                    raise NotImplementedError(f"oracle: ONNX op {op}")
                if n.outputs[0] in self.round_after:
                    out = _bf16(out)
                vals[n.outputs[0]] = out
        if all_values or keep is not None:
            names = keep if keep is not None else [k for k in vals if k and k not in self.consts]
            return {k: vals[k].to(torch.float32).numpy() for k in names if isinstance(vals.get(k), torch.Tensor)}
        return [vals[o].to(torch.float32).numpy() for o in self.graph.outputs]


class OrtSubstituteSession:
    """Duck-types the two onnxruntime.InferenceSession calls the reference makes (detector.py:118,135)."""

    def __init__(self, path_or_bytes, dtype=torch.float32):
        if isinstance(path_or_bytes, (bytes, bytearray)):
            data = bytes(path_or_bytes)
        else:
            with open(path_or_bytes, "rb") as fp:
                data = fp.read()
        self.exe = GraphExecutor(data, dtype)

    def run(self, output_names, feeds):
        assert output_names is None
        return self.exe.run(feeds["input"])

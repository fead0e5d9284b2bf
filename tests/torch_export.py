"""Independent ONNX writer for tests: torch's TorchScript exporter serialises in C++ and only needs the
(absent) `onnx` package for a post-processing step, which is patched to the identity here."""
import io
import warnings

import torch
import torch.nn as nn
import torch.nn.functional as F


def _patch():
    from torch.onnx._internal.torchscript_exporter import onnx_proto_utils
    onnx_proto_utils._add_onnxscript_fn = lambda model_bytes, custom_opsets: model_bytes


class MiniYolo(nn.Module):
    """A small net with every construct of the YOLOv3 graphs: conv+BN+leaky, stride-2 conv, residual add,
    max-pool (stride 2 and padded stride 1), x2 nearest upsample, route-concat, two linear heads."""

    def __init__(self, nc=4, width=16):
        super().__init__()
        w = width
        self.nc = nc
        self.c1, self.b1 = nn.Conv2d(3, w, 3, 1, 1, bias=False), nn.BatchNorm2d(w)
        self.c2, self.b2 = nn.Conv2d(w, 2 * w, 3, 2, 1, bias=False), nn.BatchNorm2d(2 * w)
        self.c3, self.b3 = nn.Conv2d(2 * w, w, 1, bias=False), nn.BatchNorm2d(w)
        self.c4, self.b4 = nn.Conv2d(w, 2 * w, 3, 1, 1, bias=False), nn.BatchNorm2d(2 * w)
        self.c5, self.b5 = nn.Conv2d(2 * w, 4 * w, 3, 1, 1, bias=False), nn.BatchNorm2d(4 * w)
        self.c6, self.b6 = nn.Conv2d(4 * w, 2 * w, 1, bias=False), nn.BatchNorm2d(2 * w)
        self.h1 = nn.Conv2d(2 * w, 3 * (5 + nc), 1)
        self.c7, self.b7 = nn.Conv2d(2 * w, w, 1, bias=False), nn.BatchNorm2d(w)
        self.c8, self.b8 = nn.Conv2d(3 * w, 2 * w, 3, 1, 1, bias=False), nn.BatchNorm2d(2 * w)
        self.h2 = nn.Conv2d(2 * w, 3 * (5 + nc), 1)
        for m in self.modules():
            if isinstance(m, nn.BatchNorm2d):
                nn.init.uniform_(m.weight, 0.8, 1.2)
                nn.init.normal_(m.bias, 0.0, 0.1)
                m.running_mean.normal_(0.0, 0.1)
                m.running_var.uniform_(0.8, 1.2)

    def forward(self, x):
        act = lambda t: F.leaky_relu(t, 0.1)
        a = act(self.b1(self.c1(x)))                      # w   @ S
        b = act(self.b2(self.c2(a)))                      # 2w  @ S/2
        r = act(self.b4(self.c4(act(self.b3(self.c3(b))))))
        b = r + b                                         # residual
        p = F.max_pool2d(b, 2, 2)                         # 2w  @ S/4
        q = act(self.b5(self.c5(p)))                      # 4w
        q = F.max_pool2d(F.pad(q, (0, 1, 0, 1), value=float("-inf")), 2, 1)  # padded stride-1 pool
        t = act(self.b6(self.c6(q)))                      # 2w  @ S/4
        out1 = self.h1(t)
        u = F.interpolate(act(self.b7(self.c7(t))), scale_factor=2, mode="nearest")  # w @ S/2
        v = act(self.b8(self.c8(torch.cat([u, b], 1))))
        out2 = self.h2(v)
        return out1, out2


def export(model: nn.Module, size: int, training_form: bool = False, opset: int = 11) -> bytes:
    _patch()
    f = io.BytesIO()
    kw = {}
    if training_form:
        kw = dict(training=torch.onnx.TrainingMode.PRESERVE, do_constant_folding=False)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        torch.onnx.export(model, torch.zeros(1, 3, size, size), f, input_names=["input"], opset_version=opset,
                          dynamo=False, **kw)
    return f.getvalue()

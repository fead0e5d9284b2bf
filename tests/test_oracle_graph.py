"""The oracle's ONNX executor against independent implementations of the same operators: torch eager
modules serialised by torch's own exporter (tests/torch_export.py).  The real reference for this half
(onnxruntime) is not installable here, so this is what pins the operator semantics."""
import numpy as np
import pytest
import torch

from fastdet_b200 import modelgen
from oracle import onnx_min, ref_graph
from tests import torch_export


@pytest.mark.parametrize("training_form", [False, True])
def test_executor_matches_torch_eager(training_form):
    torch.manual_seed(0)
    net = torch_export.MiniYolo(nc=4, width=16).eval()
    data = torch_export.export(net, 64, training_form=training_form)
    g = onnx_min.load(data)
    ops = {n.op for n in g.nodes}
    assert "Conv" in ops and "Concat" in ops and "MaxPool" in ops
    assert ("BatchNormalization" in ops) == training_form  # eval export folds BN, the other keeps the nodes
    x = np.random.default_rng(0).random((2, 3, 64, 64), dtype=np.float32)
    with torch.no_grad():
        want = [t.numpy() for t in net(torch.from_numpy(x))]
    got = ref_graph.GraphExecutor(data).run(x)
    assert len(got) == 2
    for a, b in zip(got, want):
        assert a.shape == b.shape
        np.testing.assert_allclose(a, b, rtol=1e-4, atol=1e-5)
    got64 = ref_graph.GraphExecutor(data, dtype=torch.float64).run(x)
    for a, b in zip(got64, want):
        np.testing.assert_allclose(a, b, rtol=1e-4, atol=1e-5)


def test_generated_variants_agree():
    """Every exporter form modelgen can emit describes the same function (BN folded vs not, Resize vs
    Upsample, Pad+MaxPool vs padded MaxPool, initializers vs Constant nodes, raw vs typed data)."""
    x = np.random.default_rng(1).random((1, 3, 96, 96), dtype=np.float32)
    base = ref_graph.GraphExecutor(modelgen.build_onnx("tiny", 3, 96, seed=5)).run(x)
    assert [o.shape for o in base] == [(1, 24, 3, 3), (1, 24, 6, 6)]
    alt = modelgen.ExportOptions(fold_bn=False, upsample_op="Upsample", pool_pad="pad_node", const_as="constant_node",
                                 raw_data=False, packed_attrs=False, batch=1)
    other = ref_graph.GraphExecutor(modelgen.build_onnx("tiny", 3, 96, seed=5, opts=alt)).run(x)
    for a, b in zip(base, other):
        np.testing.assert_allclose(a, b, rtol=2e-4, atol=2e-4)


def test_generated_full_topology():
    data = modelgen.build_onnx("full", 2, 128, seed=7)
    g = onnx_min.load(data)
    ops = [n.op for n in g.nodes]
    assert ops.count("Conv") == 75 and ops.count("Add") == 23 and ops.count("Resize") == 2 and ops.count("Concat") == 2
    x = (modelgen.synthetic_frame(3, 128)[None] / 255).astype(np.float32).transpose(0, 3, 1, 2)
    exe = ref_graph.GraphExecutor(data)
    outs = exe.run(x)
    assert [o.shape for o in outs] == [(1, 21, 4, 4), (1, 21, 8, 8), (1, 21, 16, 16)]  # coarsest first
    vals = exe.run(x, all_values=True)
    rms = [float(np.sqrt(np.mean(v ** 2))) for k, v in vals.items() if k.endswith("_act")]
    assert 0.05 < min(rms) and max(rms) < 20.0  # calibrated init keeps activations O(1) on unseen frames


def test_missing_input_name_is_rejected():
    data = modelgen.build_onnx("tiny", 3, 64, seed=1).replace(b"\x0a\x05input", b"\x0a\x05inpux")
    with pytest.raises(KeyError):
        ref_graph.GraphExecutor(data)

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


def test_bf16_operand_mode_follows_the_planner():
    """dtype="bf16": BatchNormalization folded before the rounding (both exporter forms give the same heads bit for bit
    only if the fold is the planner's), stored activations are bf16 values, heads stay fp32."""
    x = (modelgen.synthetic_frame(3, 96)[None] / 255).astype(np.float32).transpose(0, 3, 1, 2)
    folded = modelgen.build_onnx("tiny", 3, 96, seed=5)
    exe = ref_graph.GraphExecutor(folded, dtype="bf16")
    vals = exe.run(x, all_values=True)
    acts = [v for k, v in vals.items() if k in exe.round_after]
    assert len(acts) >= 11
    for v in acts:
        assert np.array_equal(v, torch.from_numpy(v).to(torch.bfloat16).to(torch.float32).numpy())
    heads = exe.run(x)
    assert any(not np.array_equal(h, torch.from_numpy(h).to(torch.bfloat16).to(torch.float32).numpy()) for h in heads)
    ref = ref_graph.GraphExecutor(folded).run(x)
    for a, b in zip(heads, ref):
        assert 0 < np.abs(a - b).max() <= 2e-2 * np.abs(b).max()
    # the unfolded export (BatchNormalization nodes present) folds to the same operands up to the exporter's own fp32
    # fold of the folded form: heads agree far inside one bf16 step of the activations
    alt = modelgen.ExportOptions(fold_bn=False)
    heads_bn = ref_graph.GraphExecutor(modelgen.build_onnx("tiny", 3, 96, seed=5, opts=alt), dtype="bf16").run(x)
    for a, b in zip(heads, heads_bn):
        assert np.abs(a - b).max() <= 1e-2 * np.abs(b).max()


def test_bf16_operand_floor():
    """What BASELINE.json's prescribed arithmetic (bf16 operands, fp32 accumulation) costs, with no GPU involved.
    (i)   bf16-operand oracle vs fp32 oracle: raw heads stay inside the spec's 2e-2 * max|ref| and scores inside 1e-2, but the
          box IoU of matched detections does NOT stay above 0.99 — exp(tw), exp(th) amplify the ~1 % logit noise of these
          random-init nets.
    (ii)  two bf16-operand evaluations that differ ONLY in how the fp32 sums are formed (fp32 vs exact float64 accumulation of
          the same bf16 products) are as far from each other as from fp32: a flipped bf16 rounding perturbs every sum it feeds
          by a fraction of a rounding step and flips more downstream.  No implementation — CPU or tensor core — can agree with
          another one's boxes to 0.99 IoU through bf16 activation storage.
    (iii) keeping the last two convolutions in front of every head in fp32 (weights and the activation between them) does
          not buy the bound back: the noise is accumulated over the depth of the network, not added at the end.
    The GPU tests therefore hold the CUDA path to the floor measured here, against both oracles
    (tests/test_gpu_parity.py::test_detections_match_oracle)."""
    from oracle import ref_post
    spread = {"fp32": ([], []), "acc64": ([], []), "tail": ([], [])}
    herr = []
    for arch, nc, seed, frames in (("tiny", 80, 1, 3), ("rsu", 9, 3, 2)):
        data = modelgen.build_onnx(arch, nc, 416, seed)
        e32, e16 = ref_graph.GraphExecutor(data), ref_graph.GraphExecutor(data, dtype="bf16")
        e16x = ref_graph.GraphExecutor(data, dtype="bf16", accumulate=torch.float64)
        e16t = ref_graph.GraphExecutor(data, dtype="bf16", fp32_tail=2)
        for s in range(frames):
            x = ref_post.normalise(modelgen.synthetic_frame(200 + s, 416))
            h32, h16, h16x, h16t = e32.run(x), e16.run(x), e16x.run(x), e16t.run(x)
            herr.append(max(float(np.abs(a - b).max() / np.abs(a).max()) for a, b in zip(h32, h16)))
            for key, (ha, hb) in (("fp32", (h32, h16)), ("acc64", (h16, h16x)), ("tail", (h32, h16t))):
                i, d = ref_post.detection_spread(ha, hb, nc, (416, 416), 0.1)
                spread[key][0].extend(i)
                spread[key][1].extend(d)
    stats = {}
    for key, (i, d) in spread.items():
        i, d = np.array(i), np.array(d)
        stats[key] = i, d
        print(f"{key:>6}: {len(i)} matched boxes, IoU min {i.min():.4f} median {np.median(i):.4f} >=0.99: {np.mean(i >= 0.99):.2f}; "
              f"|dconf| max {d.max():.4f}")
    print(f"bf16-operand oracle vs fp32 oracle: head err max {max(herr):.4f}")
    assert max(herr) <= 2e-2
    ious, dconfs = stats["fp32"]
    assert len(ious) >= 60 and dconfs.max() <= 1e-2
    assert ious.min() >= 0.95 and np.median(ious) >= 0.98
    assert np.mean(ious >= 0.99) < 0.9   # (i) the floor: bf16 operands alone miss the 0.99 bound on a good share of the boxes
    assert np.mean(stats["acc64"][0] >= 0.99) < 0.9  # (ii) ... between two bf16-correct evaluations as well
    assert np.mean(stats["tail"][0] >= 0.99) < 0.9   # (iii) ... and fp32 head layers do not repair it

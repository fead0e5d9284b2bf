"""CPU oracle for the detection hot path — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this package; the
product (fastdet_b200/) never does.  See ref_post.py (pinned against the reference's own code through
tests/golden/) and ref_graph.py (ONNX-spec restatement; parity UNPINNED at the onnxruntime boundary).
"""

"""ORACLE -- test infrastructure, NOT product code.

A CPU (NumPy) restatement of the reference's algorithm for the inference hot path (every function cites the
reference file:line it follows), pinned against golden vectors produced by the reference's own code
(tests/golden/make_golden.py) and against the vendored Caffe's known-answer tests
(tests/test_oracle_kat.py).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs may import this package; the product (smallhardface_b200/) never does and
fails loudly when its CUDA extension is missing.

Parity status: a1-a3 pre-processing pinned by the installed OpenCV (the reference's third-party
dependency); a6-a13 layers pinned by Caffe's gtest known answers; a14 ProposalLayer, a16 bbox_vote,
a17 cpu_nms/py_cpu_nms and a18 bbox_overlaps pinned bit-exactly by outputs of the reference's own
modules run in this container.  The vendored Caffe itself cannot be built here, so the
conv stack has no end-to-end golden from the original binary.
"""

"""oracle/ — TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's convolutional-occupancy hot path
(jeffsonyu/VTacO: src/common.py, src/layers.py, src/encoder/pointnet.py,
src/conv_onet/models/decoder.py, src/conv_onet/generation.py) used as the
checker for the CUDA path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` leg may import this package.  The product
(`vtaco_b200/`) never imports it and has no CPU fallback.

Pinning status
--------------
* convonet.py (encoder / decoder / eval_points): PINNED against outputs of the
  reference itself, run in the build container from /root/reference by
  tests/golden/make_golden.py; the vectors live in tests/golden/*.npz and
  tests/test_oracle_golden.py replays them.
  torch_scatter (2.0.9, un-vendored) is restated from its published semantics.
* marching_cubes.py: PARITY UNPINNED.  scikit-image (un-pinned in the
  reference's requirements.txt:6) is not installed, not vendored and there is
  no wheel offline; the reference has no test on it.  The restatement follows
  the call contract of src/conv_onet/generation.py:268-272 and is pinned only
  by table-independent invariants (see its header).
"""

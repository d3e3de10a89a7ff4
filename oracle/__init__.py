"""CPU oracle for the VirConv sparse-convolution hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under `oracle/` is product code: it may be imported by
`tests/`, by `__graft_entry__.smoke()` and by the `cpu_baseline` / `--impl reference` legs of
`bench.py`, and only as the checker (or the timed CPU baseline) — never by `virconv_b200/`.

What it restates
  * the parts of the path that live in the reference tree (`pcdet/models/backbones_3d/
    spconv_backbone.py`, `pcdet/datasets/augmentor/X_transform.py`, `pcdet/utils/calibration_kitti.py`,
    `pcdet/models/backbones_3d/vfe/mean_vfe.py`, `pcdet/datasets/processor/data_processor.py`) —
    each function cites the file:line it follows, and `oracle/make_golden.py` pins them against the
    reference's own Python (imported from /root/reference in the build container) through the
    fixtures in `tests/golden/`;
  * the sparse-convolution operators themselves, which live in the third-party package
    **spconv 2.1.22** (+ cumm) that is NOT vendored in the reference and cannot be installed here.
    Those semantics are restated from the published algorithm (submanifold / regular sparse
    convolution, "Native" gather-GEMM-scatter) and are pinned against brute-force dense
    `torch.nn.functional.conv{2,3}d`, because the reference ships no test, golden vector or fixture
    for them.  **PARITY UNPINNED for the spconv internals** (pair ordering, duplicate-coordinate
    winner, output-row order of strided convolution): the rules chosen are written in
    `oracle/rulebook.py` and DESIGN.md.
"""

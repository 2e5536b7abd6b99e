"""CPU oracle for the ``spconv-contraction`` hot path — TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU (NumPy / PyTorch-CPU, no spconv, no Open3D), the
algorithm that project-lightlin/SmartQSM v2.1.3 runs for the ``spconv_based_contraction``
skeletonisation stage.  Every function cites the reference
``file:line`` it follows (paths relative to the reference checkout).

Who may use it: only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py``.  The product package ``smartqsm_b200`` never
imports, links or executes anything in here; it fails loudly when its CUDA library is
missing instead of falling back to this code.

PARITY UNPINNED.  The reference has no tests, golden vectors or fixtures (SURVEY.md §4,
F12) and the arithmetic lives in third-party wheels that are neither vendored nor
version-pinned nor installable here: ``spconv``/``cumm`` (PointToVoxel, SubMConv3d,
SparseConv3d, SparseInverseConv3d), ``open3d`` (voxel_down_sample_and_trace,
compute_point_cloud_distance).  Their published algorithms are restated from the
reference's own call sites under the parity contract of SURVEY.md
App. D: CPU (first-come, deterministic) semantics, "clean bounds" rule on the far faces of
the batch bounding box.  What pins this oracle:

* the reference's OWN Python for tiling, collation, network wiring, decode, contraction
  update and the radius-neighbour edge set is executed unmodified on top of a stub
  ``spconv``/``open3d`` (``oracle/refshim``) in ``tests/golden/make_golden.py`` and must
  agree with the restatement (golden fixtures in ``tests/golden``);
* the sparse convolutions are cross-checked against dense ``torch.nn.functional.conv3d``
  / ``conv_transpose3d`` on small crops;
* ``scipy`` (the reference's real dependency for the radius graph) is importable and is
  used directly.

The residual risk "restatement != genuine spconv/Open3D" stays open until one genuine
run can be compared.
"""

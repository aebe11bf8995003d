"""CPU oracle for the Mask R-CNN R50/R101-C4 hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl
reference`` legs of ``bench.py`` may import it, and there only as the checker
or as the timed CPU baseline -- never as the path that is shipped.  The product
package (``chainer_mask_rcnn_b200``) never imports this package and fails loudly
when its CUDA library is missing.

Pinning status (see DESIGN.md "Oracle"):

* ``roi_align``        -- PINNED.  ``oracle/ref_loader.py`` executes the
  reference's own ``functions/roi_align_2d.py`` (forward_cpu / backward_cpu)
  verbatim under a stub ``chainer`` module; ``tests/golden/make_golden.py``
  stored its outputs in ``tests/golden/roi_align_*.npz`` and the numpy
  restatement in ``oracle/roi_align.py`` is checked against them.
* ``affine_channel``   -- PINNED the same way (``functions/affine_channel_2d.py``).
* ``bbox`` (anchors, loc2bbox, bbox_iou, NMS, ProposalCreator,
  AnchorTargetCreator) -- PARITY UNPINNED.  The arithmetic lives in chainercv
  (requirements.txt:2 ``chainercv>=0.9.0``, unpinned, not vendored, not
  installable offline) and no reference test holds a golden vector for it
  (SURVEY.md 8c).  The restatement follows the published chainercv 0.13
  algorithm as recalled in SURVEY.md Appendix B, anchored on the reference's
  call sites.
* ``nn`` (conv / pooling / linear / losses) -- PARITY UNPINNED for the same
  reason (chainer is absent); restates Chainer's CPU algorithm
  (im2col + tensordot) and is cross-checked against torch CPU fp32 ops.
"""

"""Batch assembly on the input side of the hot path (the reference's
``chainer_mask_rcnn.datasets.concat_examples``).  Data sets themselves are out of scope
(DESIGN.md section 7)."""
from .concat_examples import concat_examples, pinned_empty

__all__ = ['concat_examples', 'pinned_empty']

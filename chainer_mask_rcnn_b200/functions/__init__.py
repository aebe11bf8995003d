"""Operators under the reference's names (chainer_mask_rcnn.functions): the ROIAlign
pooler and the frozen-BatchNorm affine.  ``crop_and_resize`` (an alternative pooler) is out
of scope, see DESIGN.md section 7."""
from .affine_channel_2d import AffineChannel2DFunction, affine_channel_2d
from .roi_align_2d import ROIAlign2D, roi_align_2d

__all__ = ['AffineChannel2DFunction', 'affine_channel_2d', 'ROIAlign2D', 'roi_align_2d']

# flake8: noqa
from .affine_channel_2d import affine_channel_2d
from .affine_channel_2d import AffineChannel2DFunction
from .roi_align_2d import roi_align_2d
from .roi_align_2d import ROIAlign2D

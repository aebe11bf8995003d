"""chainer_mask_rcnn_b200 -- the Mask R-CNN R50/R101-C4 hot path of
wkentaro/chainer-mask-rcnn, rebuilt for NVIDIA B200 (sm_100a).

Host code is Python over ``libcmr_b200.so`` (hand-written CUDA behind the C ABI in
``include/cmr_b200.h``).  Arrays are ``torch`` CUDA tensors (device memory and
streams are torch's); NumPy arrays are accepted at the operator surface and are
copied to the GPU and back.  There is no CPU implementation in this package.
"""
from . import _lib  # noqa: F401
from . import functions  # noqa: F401
from . import utils  # noqa: F401
from . import models  # noqa: F401
from . import optimizers  # noqa: F401
from . import datasets  # noqa: F401

__version__ = '0.2.0'

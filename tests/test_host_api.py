"""Host-side behaviour of the operator surface that needs no GPU: argument and
type errors mirror the reference (functions/roi_align_2d.py:29-47, 555-556)."""
import numpy as np
import pytest

import chainer_mask_rcnn_b200 as cmr
from chainer_mask_rcnn_b200 import functions
from chainer_mask_rcnn_b200.utils import config


def test_roi_align_ctor_type_errors():
    for bad in ((2.0, 2, 1.0, 0), (2, '2', 1.0, 0), (2, 2, 1.0, -1), (2, 2, 1.0, 1.5)):
        with pytest.raises(TypeError):
            functions.ROIAlign2D(*bad)
    with pytest.raises(TypeError):
        functions.ROIAlign2D(2, 2, '0.5')
    f = functions.ROIAlign2D(7, 7, 1)          # int scale is coerced to float
    assert isinstance(f.spatial_scale, float) and f.sampling_ratio == 0


def test_roi_align_axes_value_error():
    x = np.zeros((1, 1, 4, 4), np.float32)
    r = np.zeros((1, 5), np.float32)
    with pytest.raises(ValueError):
        functions.roi_align_2d(x, r, 2, 2, 1.0, axes='ab')


def test_config_train_switch():
    pc = cmr.utils.ProposalCreator(n_test_pre_nms=6000, n_test_post_nms=1000, min_size=0)
    assert pc.budgets() == (12000, 2000)
    with config.using_config('train', False):
        assert pc.budgets() == (6000, 1000)
    assert config.train is True


def test_anchor_base_matches_oracle():
    from oracle import bbox as ob
    for scales in ((4, 8, 16, 32), (2, 4, 8, 16, 32)):
        a = cmr.utils.generate_anchor_base(16, (0.5, 1, 2), scales)
        np.testing.assert_allclose(a, ob.generate_anchor_base(16, (0.5, 1, 2), scales), atol=1e-5)

"""datasets.concat_examples: the reference's batch format
(chainer_mask_rcnn/datasets/concat_examples.py:6-34 over chainer's _concat_arrays), as the
training script configures it (padding=0, indices_concat=[0, 2, 3, 4],
indices_to_device=[0, 1]; examples/train_common.py)."""
import numpy as np
import pytest

from chainer_mask_rcnn_b200.datasets import concat_examples


def _example(rs, h, w, n):
    return (rs.standard_normal((3, h, w)).astype(np.float32),
            rs.uniform(0, 50, (n, 4)).astype(np.float32),
            rs.randint(0, 80, n).astype(np.int32),
            rs.randint(0, 2, (n, h, w)).astype(np.int32),
            np.float64(1.6))


def test_reference_training_configuration():
    rs = np.random.RandomState(0)
    batch = [_example(rs, 20, 30, 3), _example(rs, 24, 26, 5)]
    imgs, bboxes, labels, masks, scales = concat_examples(
        batch, device=None, padding=0, indices_concat=[0, 2, 3, 4], indices_to_device=[0, 1])
    assert imgs.shape == (2, 3, 24, 30) and imgs.dtype == np.float32
    np.testing.assert_array_equal(imgs[0, :, :20, :30], batch[0][0])
    assert (imgs[0, :, 20:] == 0).all() and (imgs[1, :, :, 26:] == 0).all()
    assert isinstance(bboxes, list) and len(bboxes) == 2            # not concatenated
    np.testing.assert_array_equal(bboxes[1], batch[1][1])
    assert labels.shape == (2, 5) and labels.dtype == np.int32
    np.testing.assert_array_equal(labels[0, :3], batch[0][2])
    assert (labels[0, 3:] == 0).all()
    assert masks.shape == (2, 5, 24, 30) and masks.dtype == np.int32
    np.testing.assert_array_equal(masks[1, :, :, :26], batch[1][3])
    assert (masks[0, 3:] == 0).all()
    assert scales.shape == (2,) and scales.dtype == np.float64


def test_same_shapes_need_no_padding_and_errors():
    rs = np.random.RandomState(1)
    batch = [_example(rs, 8, 8, 2), _example(rs, 8, 8, 2)]
    out = concat_examples(batch)
    assert out[0].shape == (2, 3, 8, 8) and out[1].shape == (2, 2, 4)
    with pytest.raises(ValueError):
        concat_examples([])
    with pytest.raises(ValueError):
        concat_examples([_example(rs, 8, 8, 2), _example(rs, 9, 8, 2)])   # no padding value


def test_canvas_gives_every_batch_one_geometry():
    rs = np.random.RandomState(2)
    batch = [_example(rs, 20, 30, 3), _example(rs, 24, 26, 3)]
    imgs, _, labels, masks, _ = concat_examples(batch, padding=0, indices_concat=[0, 2, 3, 4],
                                                indices_to_device=[], canvas=(32, 32))
    assert imgs.shape == (2, 3, 32, 32) and masks.shape == (2, 3, 32, 32)
    assert labels.shape == (2, 3)                                    # not an image: untouched
    np.testing.assert_array_equal(masks[1, :, :24, :26], batch[1][3])
    assert (masks[:, :, 24:] == 0).all()
    with pytest.raises(ValueError):
        concat_examples(batch, padding=0, canvas=(16, 64))

"""The C-ABI library builds, loads, and exports every symbol that include/*.h
declares (no compute calls: this runs without a GPU)."""
import ctypes
import glob
import os
import re

from chainer_mask_rcnn_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    names = []
    for path in glob.glob(os.path.join(ROOT, 'include', '*.h')):
        text = open(path).read()
        text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
        names += re.findall(r'\b(cmr_[a-z0-9_]+)\s*\(', text)
    return sorted(set(names))


def test_header_symbols_are_exported(lib):
    declared = _declared_functions()
    assert len(declared) >= 10
    for name in declared:
        assert hasattr(lib, name), 'libcmr_b200.so does not export ' + name


def test_binding_table_matches_header(lib):
    assert sorted(_lib.EXPORTED_SYMBOLS) == _declared_functions()


def test_version_and_status_strings(lib):
    assert lib.cmr_version() == 5       # bumped with every ABI change
    assert lib.cmr_status_string(0) == b'ok'
    assert b'workspace' in lib.cmr_status_string(-3)


def test_workspace_queries_need_no_gpu(lib):
    n = 12000
    nb = (n + 63) // 64
    assert lib.cmr_nms_workspace_bytes(n) >= n * nb * 8
    assert lib.cmr_proposals_workspace_bytes(2, 64260, 12000) > 2 * n * nb * 8


def test_conv_desc_layout_matches_header():
    text = open(os.path.join(ROOT, 'include', 'cmr_b200.h')).read()
    body = re.search(r'typedef struct cmr_conv_desc \{(.*?)\} cmr_conv_desc;', text, re.S).group(1)
    fields = [f.strip() for decl in re.findall(r'int ([^;]+);', body) for f in decl.split(',')]
    assert fields == [name for name, _ in _lib.ConvDesc._fields_]
    assert ctypes.sizeof(_lib.ConvDesc) == 4 * len(fields)


def test_prepare_size_is_cv_round(lib):
    """cmr_prepare_size is host-only: the size cv2.resize(img, None, fx, fy) produces --
    cvRound (ties to even) of the products -- against the oracle and cv2 itself."""
    import cv2
    import numpy as np
    from oracle import prepare as op
    h, w = ctypes.c_int(), ctypes.c_int()
    rs = np.random.RandomState(0)
    cases = [(80, 103, 0.5), (80, 101, 0.5), (500, 833, 1.6), (3, 5, 0.5), (480, 640, 800 / 480.)]
    cases += [(int(rs.randint(8, 400)), int(rs.randint(8, 400)), float(rs.uniform(0.3, 3.)))
              for _ in range(50)]
    for H, W, f in cases:
        assert lib.cmr_prepare_size(H, W, f, f, ctypes.addressof(h), ctypes.addressof(w)) == 0
        assert (h.value, w.value) == op.out_size(H, W, f, f)
        got = cv2.resize(np.zeros((H, W, 3), np.float32), None, fx=f, fy=f).shape[:2]
        assert (h.value, w.value) == got, (H, W, f)
    assert lib.cmr_prepare_size(0, 5, 1., 1., ctypes.addressof(h), ctypes.addressof(w)) == -1
    assert lib.cmr_prepare_size(5, 5, 0.01, 0.01, ctypes.addressof(h), ctypes.addressof(w)) == -1


def test_detections_workspace_scales_with_rows(lib):
    """One NMS row per (image, class): the mask dominates, rows * max_roi * ceil(max_roi/64) words."""
    b = lib.cmr_detections_workspace_bytes(2, 1000, 81, 19000)
    assert b >= 2 * 80 * 1000 * 16 * 8
    assert lib.cmr_detections_workspace_bytes(0, 1000, 81, 19000) == 256

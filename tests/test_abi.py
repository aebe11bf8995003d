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
    assert lib.cmr_version() >= 1
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

"""ncu driver: cmr_nms on 12000 clustered boxes (the train step's pre-NMS count), limit 2000."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import synth  # noqa: E402
from chainer_mask_rcnn_b200 import _lib  # noqa: E402

n = 12000
limit = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
rs = np.random.RandomState(0)
boxes = torch.from_numpy(synth.clustered_boxes(rs, n, 800, 1088, n // 12)).cuda()
keep = torch.empty((n,), dtype=torch.int32, device='cuda')
nk = torch.zeros((1,), dtype=torch.int32, device='cuda')
lib = _lib.load()
wsb = lib.cmr_nms_workspace_bytes(n)
ws = torch.empty((wsb // 8,), dtype=torch.int64, device='cuda')
for _ in range(4):
    _lib.call('cmr_nms', _lib.ptr(boxes), n, 0.7, limit, _lib.ptr(keep), _lib.ptr(nk), _lib.ptr(ws),
              wsb, _lib.stream_ptr())
torch.cuda.synchronize()
print('kept', int(nk.item()))

"""cProfile of MaskRCNN.predict on one 800x1333 image (host-side costs)."""
import cProfile, pstats, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from chainer_mask_rcnn_b200 import models
rs = np.random.RandomState(0)
model = models.MaskRCNNResNet(50, 80, anchor_scales=(2, 4, 8, 16, 32), roi_size=14, min_size=800, max_size=1333)
model.score_thresh = 1. / 81. * 1.02
imgs = [rs.uniform(0, 255, (3, 800, 1333)).astype(np.float32)]
for _ in range(2):
    model.predict(imgs)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    model.predict(imgs)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(28)

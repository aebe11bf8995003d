"""Device target creators (csrc/targets.cu) against the oracle's assignment rules.

Which random subset is drawn is not part of the contract (the reference draws from
NumPy's global generator on the host); eligibility, subset sizes, labels, assignments
and box targets are, and are checked exactly / to fp32 round-off."""
import numpy as np
import pytest
import torch

import synth
from chainer_mask_rcnn_b200.models import utils as mu
from oracle import bbox as ob

pytestmark = pytest.mark.gpu


def _gt(seeds, H, W):
    bbs, lbs = [], []
    for s in seeds:
        _, bbox, label, _, _ = synth.detection_scene(s, n_gt=3 + 4 * s, H=H, W=W)
        bbs.append(bbox); lbs.append(label)
    return bbs, lbs


@pytest.mark.parametrize('seed', [1, 2])
def test_anchor_targets(seed):
    H, W = 320, 416
    bbs, lbs = _gt([seed, seed + 1], H, W)
    base = ob.generate_anchor_base(16, (0.5, 1, 2), (2, 4, 8, 16, 32))
    anchor = ob.enumerate_shifted_anchor(base, 16, H // 16, W // 16)
    gt = mu.GroundTruth(bbs, lbs, torch.device('cuda'))
    atc = mu.DeviceAnchorTargetCreator()
    a_dev = torch.from_numpy(anchor).cuda()
    loc, label = atc(gt, a_dev, (H, W), seed=seed)
    loc2, label2 = atc(gt, a_dev, (H, W), seed=seed)
    assert torch.equal(label, label2)                      # same seed, same draw
    _, label3 = atc(gt, a_dev, (H, W), seed=seed + 100)
    loc, label, label3 = loc.cpu().numpy(), label.cpu().numpy(), label3.cpu().numpy()
    for b in range(2):
        # the oracle without subsampling gives the eligible sets
        full = ob.AnchorTargetCreator(n_sample=10 ** 9)
        want_loc, elig = full(bbs[b], anchor, (H, W), rng=np.random.RandomState(0))
        n_pos, n_neg = int((elig == 1).sum()), int((elig == 0).sum())
        got = label[b]
        assert set(np.unique(got)) <= {-1, 0, 1}
        assert ((got == 1) <= (elig == 1)).all() and ((got == 0) <= (elig == 0)).all()
        kept_pos = min(n_pos, 128)
        assert (got == 1).sum() == kept_pos
        assert (got == 0).sum() == min(n_neg, 256 - kept_pos)
        np.testing.assert_allclose(loc[b], want_loc, rtol=1e-5, atol=1e-5)
        if n_neg > 256:
            assert (label3[b] != got).any()                # another seed, another subset


@pytest.mark.parametrize('n_roi,n_sample', [(300, 64), (2000, 512), (40, 512)])
def test_proposal_targets(n_roi, n_sample):
    H, W = 320, 416
    bbs, lbs = _gt([1, 2], H, W)
    rs = np.random.RandomState(n_roi)
    max_roi = n_roi + 7
    rois = np.zeros((2, max_roi, 4), np.float32)
    counts = np.array([n_roi, n_roi - 5], np.int32)
    for b in range(2):
        jit = bbs[b][rs.randint(0, len(bbs[b]), n_roi // 2)] + rs.normal(0, 6, (n_roi // 2, 4))
        r = np.concatenate([jit, synth.random_boxes(rs, n_roi - n_roi // 2, H, W)])
        r = np.stack([np.minimum(r[:, 0], r[:, 2]), np.minimum(r[:, 1], r[:, 3]),
                      np.maximum(r[:, 0], r[:, 2]) + 1, np.maximum(r[:, 1], r[:, 3]) + 1], 1)
        rois[b, :n_roi] = np.clip(r, 0, [H, W, H, W])
    gt = mu.GroundTruth(bbs, lbs, torch.device('cuda'))
    ptc = mu.DeviceProposalTargetCreator(n_sample=n_sample)
    out = ptc.sample(torch.from_numpy(rois).cuda(), torch.from_numpy(counts).cuda(), gt, seed=3)
    sroi, gloc, glab, gasg, npos = (t.cpu().numpy() for t in out)
    mean, std = np.zeros(4, np.float32), np.array([.1, .1, .2, .2], np.float32)
    for b in range(2):
        cand = np.concatenate([rois[b, :counts[b]], bbs[b]])
        iou = ob.bbox_iou(cand, bbs[b])
        assign, best = iou.argmax(1), iou.max(1)
        n_pos_all, n_neg_all = int((best >= 0.5).sum()), int((best < 0.5).sum())
        kp = min(n_pos_all, int(np.round(n_sample * 0.25)))
        kn = min(n_sample - kp, n_neg_all)
        assert npos[b] == kp
        used = set()
        for j in range(n_sample):
            if j >= kp + kn:                                  # padding rows
                assert glab[b, j] == -1 and gasg[b, j] == -1 and not sroi[b, j].any()
                continue
            hit = np.flatnonzero((cand == sroi[b, j]).all(1))
            assert len(hit) >= 1
            c = [h for h in hit if h not in used] or list(hit)
            c = c[0]
            used.add(c)
            if j < kp:
                assert best[c] >= 0.5 and gasg[b, j] == assign[c]
                assert glab[b, j] == lbs[b][assign[c]] + 1
            else:
                assert best[c] < 0.5 and glab[b, j] == 0 and gasg[b, j] == -1
            want = (ob.bbox2loc(cand[c:c + 1], bbs[b][assign[c]:assign[c] + 1])[0] - mean) / std
            np.testing.assert_allclose(gloc[b, j], want, rtol=1e-4, atol=1e-4)
        assert len(used) == kp + kn                            # sampling without replacement


def test_mask_targets_match_host_creator():
    """The host half of the device creator rasterises exactly like the host creator."""
    roi, bbox, label, mask, _ = synth.detection_scene(4)
    np.random.seed(0)
    sr, _, glab, gm = mu.ProposalTargetCreator(n_sample=64)(roi, bbox, label, mask)
    n_pos = int((glab > 0).sum())
    cand = np.concatenate([roi, bbox])
    iou = ob.bbox_iou(cand, bbox)
    asg = np.array([iou[np.flatnonzero((cand == r).all(1))[0]].argmax() for r in sr[:n_pos]])
    got = mu.DeviceProposalTargetCreator(n_sample=64).mask_targets(
        sr[None], np.pad(asg, (0, 64 - n_pos))[None].astype(np.int32), np.array([n_pos]), [mask])
    np.testing.assert_array_equal(got[0], gm)


@pytest.mark.parametrize('dtype', [torch.uint8, torch.int32])
def test_device_mask_targets_bit_exact(dtype):
    """cmr_mask_targets against the oracle's restatement of the reference's
    round / crop / one-hot / cv2.resize / argmax pipeline: every pixel equal."""
    from oracle import mask_target as omt
    H, W, n = 320, 416, 96
    rs = np.random.RandomState(7)
    masks, sroi, asg, npos = [], [], [], []
    G = 9
    for b, seed in enumerate((2, 5)):
        roi, bbox, _, mask, _ = synth.detection_scene(seed, n_gt=8, H=H, W=W, n_roi=n)
        m = np.zeros((G, H, W), np.int32)
        m[:len(bbox)] = mask
        m[0] *= 2                                   # a label map with value 2 (one-hot path)
        m[1, ::3, ::2] = 3
        masks.append(m)
        r = roi.copy()
        r[:5] = [[0, 0, H, W], [10.5, 20.5, 11.4, 21.4], [3, 3, 3, 40], [H - 1, W - 1, H, W],
                 [0.5, 1.5, 2.5, 3.5]]             # full image, 1-pixel, empty, corner, ties
        sroi.append(r)
        asg.append(rs.randint(0, len(bbox), n).astype(np.int32))
        npos.append(n - 20 * b)
    sroi = np.stack(sroi).astype(np.float32)
    asg = np.stack(asg)
    npos = np.asarray(npos, np.int32)
    want = omt.mask_targets(sroi, asg, npos, masks, 14)
    ptc = mu.DeviceProposalTargetCreator(n_sample=n)
    got = ptc.mask_targets_device(torch.from_numpy(sroi).cuda(), torch.from_numpy(asg).cuda(),
                                  torch.from_numpy(npos).cuda(),
                                  torch.from_numpy(np.stack(masks)).to(dtype).cuda())
    got = got.cpu().numpy()
    np.testing.assert_array_equal(got, want)
    assert (got[1, npos[1]:] == -1).all() and set(np.unique(got)) <= {-1, 0, 1, 2, 3}
    if dtype == torch.uint8:
        # the same binary masks packed one bit per pixel
        binary = (np.stack(masks) > 0).astype(np.uint8)
        want_b = omt.mask_targets(sroi, asg, npos, list(binary), 14)
        packed = mu.PackedMasks.from_numpy(binary).to('cuda')
        assert packed.data.shape == (2, G, H, (W + 7) // 8)
        got_b = ptc.mask_targets_device(torch.from_numpy(sroi).cuda(), torch.from_numpy(asg).cuda(),
                                        torch.from_numpy(npos).cuda(), packed).cpu().numpy()
        np.testing.assert_array_equal(got_b, want_b)


def test_seed_dev_advances_the_draw():
    H, W = 320, 416
    bbs, lbs = _gt([1, 2], H, W)
    base = ob.generate_anchor_base(16, (0.5, 1, 2), (2, 4, 8, 16, 32))
    anchor = torch.from_numpy(ob.enumerate_shifted_anchor(base, 16, H // 16, W // 16)).cuda()
    gt = mu.GroundTruth(bbs, lbs, torch.device('cuda'), capacity=32)
    assert gt.G == 32
    atc = mu.DeviceAnchorTargetCreator()
    word = torch.zeros((1,), dtype=torch.int64, device='cuda')
    _, l0 = atc(gt, anchor, (H, W), seed=5, seed_dev=word)
    _, l_plain = atc(gt, anchor, (H, W), seed=5)
    assert torch.equal(l0, l_plain)                   # word == 0 leaves the seed alone
    word += 1
    _, l1 = atc(gt, anchor, (H, W), seed=5, seed_dev=word)
    assert (l1 >= 0).sum() == (l0 >= 0).sum() and not torch.equal(l0, l1)
    gt.fill_(bbs[::-1], lbs[::-1])                    # same device buffers, new boxes
    _, l2 = atc(gt, anchor, (H, W), seed=5)
    assert torch.equal(l2[0] >= 0, l2[0] >= 0) and not torch.equal(l2, l_plain)

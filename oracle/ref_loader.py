"""Load the reference's own operator files verbatim (dev container only).

TEST INFRASTRUCTURE.  ``/root/reference`` is read-only and exists only in the
development container; this loader is used by ``tests/golden/make_golden.py`` to
produce the committed golden vectors and by CPU tests that re-validate the numpy
restatement when the reference tree happens to be present.  Nothing on the GPU
box may call it (it returns ``None`` there).

The reference package cannot be imported as a whole (chainer / chainercv / cupy
are absent).  ``functions/roi_align_2d.py`` and
``functions/affine_channel_2d.py`` only need three chainer names on their CPU
paths, so a stub ``chainer`` package is registered just for the duration of the
load and the file is executed unmodified from where it lies.
"""
import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get('CMR_REFERENCE_ROOT', '/root/reference')


def reference_available():
    return os.path.isfile(os.path.join(
        REFERENCE_ROOT, 'chainer_mask_rcnn', 'functions', 'roi_align_2d.py'))


def _stub_chainer():
    import numpy

    chainer = types.ModuleType('chainer')
    cuda = types.ModuleType('chainer.cuda')
    cuda.get_array_module = lambda *a: numpy
    cuda.to_cpu = lambda a: a
    cuda.to_gpu = lambda a: a
    function = types.ModuleType('chainer.function')

    class Function(object):
        def retain_inputs(self, indexes):
            self._retained = indexes

    function.Function = Function
    functions = types.ModuleType('chainer.functions')
    utils = types.ModuleType('chainer.utils')
    type_check = types.ModuleType('chainer.utils.type_check')
    utils.type_check = type_check
    chainer.cuda = cuda
    chainer.function = function
    chainer.functions = functions
    chainer.utils = utils
    chainer.Function = Function
    return {
        'chainer': chainer, 'chainer.cuda': cuda, 'chainer.function': function,
        'chainer.functions': functions, 'chainer.utils': utils,
        'chainer.utils.type_check': type_check,
    }


def _stub_chainercv():
    """chainercv is absent: the two helpers the reference's
    ProposalTargetCreator imports are served by the oracle's restatements."""
    from . import bbox as ob
    names = ['chainercv', 'chainercv.links', 'chainercv.links.model',
             'chainercv.links.model.faster_rcnn', 'chainercv.links.model.faster_rcnn.utils',
             'chainercv.links.model.faster_rcnn.utils.bbox2loc', 'chainercv.utils',
             'chainercv.utils.bbox', 'chainercv.utils.bbox.bbox_iou']
    mods = {n: types.ModuleType(n) for n in names}
    mods['chainercv.links.model.faster_rcnn.utils.bbox2loc'].bbox2loc = ob.bbox2loc
    mods['chainercv.utils.bbox.bbox_iou'].bbox_iou = ob.bbox_iou
    return mods


def _load(relpath, modname, extra_stubs=None):
    if not reference_available():
        return None
    stubs = _stub_chainer()
    if extra_stubs:
        stubs.update(extra_stubs)
    stubs.pop(modname, None)
    saved = {k: sys.modules.get(k) for k in stubs}
    sys.modules.update(stubs)
    try:
        path = os.path.join(REFERENCE_ROOT, relpath)
        spec = importlib.util.spec_from_file_location(modname, path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod


def load_roi_align_module():
    """-> module of chainer_mask_rcnn/functions/roi_align_2d.py, or None."""
    return _load('chainer_mask_rcnn/functions/roi_align_2d.py', '_ref_roi_align_2d')


def load_affine_channel_module():
    """-> module of chainer_mask_rcnn/functions/affine_channel_2d.py, or None."""
    return _load('chainer_mask_rcnn/functions/affine_channel_2d.py',
                 '_ref_affine_channel_2d')


def load_proposal_target_creator_module():
    """-> module of chainer_mask_rcnn/models/utils/proposal_target_creator.py (the
    reference's own host-side sampler, run verbatim; bbox2loc / bbox_iou come from
    oracle/bbox.py), or None."""
    return _load('chainer_mask_rcnn/models/utils/proposal_target_creator.py',
                 '_ref_proposal_target_creator', _stub_chainercv())


def load_mask_rcnn_module():
    """-> module of chainer_mask_rcnn/models/mask_rcnn.py (``segm_results``, ``expand_boxes``
    and the ``MaskRCNN`` class whose ``_suppress`` / ``_to_bboxes`` run verbatim on NumPy
    arrays), or None.  chainer.Chain / Variable / F.softmax / F.sigmoid are two-line NumPy
    stubs; ``loc2bbox`` and ``non_maximum_suppression`` (chainercv, absent) come from
    oracle/bbox.py."""
    import numpy
    from . import bbox as ob
    stubs = _stub_chainercv()
    for n in ('chainercv.links.model.faster_rcnn.utils.loc2bbox',):
        stubs[n] = types.ModuleType(n)
    stubs['chainercv.links.model.faster_rcnn.utils.loc2bbox'].loc2bbox = ob.loc2bbox
    stubs['chainercv.utils'].non_maximum_suppression = ob.non_maximum_suppression
    for n in ('chainer_mask_rcnn', 'chainer_mask_rcnn.models', 'chainer_mask_rcnn.datasets'):
        stubs[n] = types.ModuleType(n)
        stubs[n].__path__ = []
    stubs['chainer_mask_rcnn.datasets'].concat_examples = lambda *a, **k: None

    class _Arr(object):
        def __init__(self, a):
            self.array = a

    def _softmax(x):
        e = numpy.exp(x - x.max(axis=1, keepdims=True))
        return _Arr((e / e.sum(axis=1, keepdims=True)).astype(numpy.float32))

    mod_stubs = _stub_chainer()
    mod_stubs['chainer'].Chain = object
    mod_stubs['chainer'].Variable = _Arr
    mod_stubs['chainer.functions'].softmax = _softmax
    mod_stubs['chainer.functions'].sigmoid = lambda x: _Arr(
        (1 / (1 + numpy.exp(-x))).astype(numpy.float32))
    stubs.update(mod_stubs)
    return _load('chainer_mask_rcnn/models/mask_rcnn.py', 'chainer_mask_rcnn.models.mask_rcnn',
                 stubs)


def ref_to_bboxes(roi_cls_locs, roi_scores, rois, roi_indices, sizes, scales, n_class,
                  mean, std, score_thresh, nms_thresh, detections_per_im):
    """MaskRCNN._to_bboxes (mask_rcnn.py:203-261) run verbatim on NumPy arrays."""
    import numpy
    mod = load_mask_rcnn_module()
    m = mod.MaskRCNN.__new__(mod.MaskRCNN)
    m.xp = numpy
    m.head = types.SimpleNamespace(n_class=n_class)     # MaskRCNN.n_class reads head.n_class
    m.loc_normalize_mean, m.loc_normalize_std = mean, std
    m.score_thresh, m.nms_thresh = score_thresh, nms_thresh
    m._detections_per_im = detections_per_im
    return m._to_bboxes(roi_cls_locs, roi_scores, rois.copy(), roi_indices, sizes, scales)


def ref_prepare(imgs, min_size, max_size, mean):
    """MaskRCNN.prepare (mask_rcnn.py:152-176) run verbatim (cv2.resize as installed)."""
    import numpy
    mod = load_mask_rcnn_module()
    m = mod.MaskRCNN.__new__(mod.MaskRCNN)
    m.min_size, m.max_size = min_size, max_size
    m.mean = numpy.asarray(mean, numpy.float32)[:, None, None]
    return m.prepare(imgs)


def ref_roi_align_forward(x, rois_xy, outh, outw, spatial_scale, sampling_ratio):
    """Reference ROIAlign2D.forward_cpu (roi_align_2d.py:61-160), run verbatim."""
    mod = load_roi_align_module()
    f = mod.ROIAlign2D(outh, outw, spatial_scale, sampling_ratio)
    y, = f.forward_cpu((x, rois_xy))
    return y


def ref_roi_align_backward(x_shape, rois_xy, gy, outh, outw, spatial_scale,
                           sampling_ratio):
    """Reference ROIAlign2D.backward_cpu (roi_align_2d.py:292-389), verbatim."""
    mod = load_roi_align_module()
    f = mod.ROIAlign2D(outh, outw, spatial_scale, sampling_ratio)
    f._bottom_data_shape = tuple(x_shape)
    gx, _ = f.backward_cpu((None, rois_xy), (gy,))
    return gx


def load_resnet_extractor_module():
    """-> module of chainer_mask_rcnn/models/resnet_extractor.py (``_get_affine_from_bn`` and
    ``_convert_bn_to_affine``, :16-44, run verbatim), or None.  chainer.links gets a bare
    ``BatchNormalization`` class, ``chainer_mask_rcnn.links.AffineChannel2D`` a two-array
    stand-in with the reference's attribute names; ``fcn`` and the ResNet*Layers bases are
    empty (only the two module-level functions are used)."""
    import numpy

    class _Var(object):
        def __init__(self, a):
            self.data = a
            self.array = a
            self.size = a.size

    class BatchNormalization(object):
        def __init__(self, gamma, beta, avg_mean, avg_var):
            self.gamma, self.beta = _Var(gamma), _Var(beta)
            self.avg_mean, self.avg_var = avg_mean, avg_var

    class AffineChannel2D(object):
        def __init__(self, channels):
            self.W = _Var(numpy.zeros((channels,), numpy.float32))
            self.b = _Var(numpy.zeros((channels,), numpy.float32))

    stubs = _stub_chainer()
    links = types.ModuleType('chainer.links')
    links.BatchNormalization = BatchNormalization
    stubs['chainer.links'] = links
    stubs['chainer'].links = links
    stubs['chainer'].dataset = types.SimpleNamespace(get_dataset_directory=lambda *a, **k: '')
    for n in ('chainer.links.model', 'chainer.links.model.vision',
              'chainer.links.model.vision.resnet'):
        stubs[n] = types.ModuleType(n)
    stubs['chainer.links.model.vision.resnet'].ResNet50Layers = type('ResNet50Layers', (), {})
    stubs['chainer.links.model.vision.resnet'].ResNet101Layers = type('ResNet101Layers', (), {})
    stubs['fcn'] = types.ModuleType('fcn')
    for n in ('chainer_mask_rcnn', 'chainer_mask_rcnn.models', 'chainer_mask_rcnn.links'):
        stubs[n] = types.ModuleType(n)
        stubs[n].__path__ = []
    stubs['chainer_mask_rcnn.links'].AffineChannel2D = AffineChannel2D
    stubs['chainer_mask_rcnn'].links = stubs['chainer_mask_rcnn.links']
    mod = _load('chainer_mask_rcnn/models/resnet_extractor.py',
                'chainer_mask_rcnn.models.resnet_extractor', stubs)
    if mod is not None:
        mod._BatchNormalization = BatchNormalization
        mod._AffineChannel2D = AffineChannel2D
    return mod


class StubChain(object):
    """The three chainer.Chain methods ``_convert_bn_to_affine`` uses (namedlinks in
    Chainer's order -- '/', then every child sorted by name, depth first --, add_link, and
    attribute deletion)."""

    def __init__(self):
        object.__setattr__(self, '_children', [])

    def add_link(self, name, link):
        object.__setattr__(self, name, link)
        self._children.append(name)

    def __delattr__(self, name):
        object.__delattr__(self, name)
        if name in self._children:
            self._children.remove(name)

    def namedlinks(self, skipself=False):
        if not skipself:
            yield '/', self
        for name in sorted(self._children):
            child = getattr(self, name)
            yield '/' + name, child
            if hasattr(child, 'namedlinks'):
                for path, link in child.namedlinks(True):
                    yield '/' + name + path, link


def ref_convert_bn_to_affine(bn_params):
    """``_convert_bn_to_affine`` (resnet_extractor.py:32-44) run verbatim on a tree of stub
    links built from ``{'res2/a/bn1': (gamma, beta, avg_mean, avg_var), ...}``.
    -> {'res2/a/bn1': (W, b), ...} read back from the AffineChannel2D links it installed."""
    mod = load_resnet_extractor_module()
    root = StubChain()
    for path, (gamma, beta, mean, var) in bn_params.items():
        node = root
        parts = path.split('/')
        for key in parts[:-1]:
            if not hasattr(node, key):
                node.add_link(key, StubChain())
            node = getattr(node, key)
        node.add_link(parts[-1], mod._BatchNormalization(gamma, beta, mean, var))
    mod._convert_bn_to_affine(root)
    out = {}
    for path in bn_params:
        node = root
        for key in path.split('/'):
            node = getattr(node, key)
        assert isinstance(node, mod._AffineChannel2D), path
        out[path] = (node.W.data.copy(), node.b.data.copy())
    return out

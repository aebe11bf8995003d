"""MomentumSGD + WeightDecay + the data-parallel gradient exchange.

Mirrors the optimizer wiring of examples/train_common.py:176-190:
``chainer.optimizers.MomentumSGD(lr, momentum=0.9)`` wrapped by
``chainermn.create_multi_node_optimizer`` with a ``WeightDecay(1e-4)`` hook, and
conv1 / bn1 / res2 / every AffineChannel2D excluded from updates (those live in the
model's frozen store and are never touched).

All trainable parameters, their gradients and momenta are flat fp32 buffers, so one
step is: ONE NCCL all-reduce (sum) over the gradient buffer + ONE fused kernel
(cmr_sgd_momentum) that applies the 1/world_size mean, the weight decay, the
momentum update and the parameter update.
"""
import os

import torch

from . import _lib
from .models import engine as E


class _eager_rpn_backward(object):
    """Lets a MaskRCNNTrainChain run its RPN branch's backward pass ahead, during the forward
    call: valid here because the gradients were just zeroed and loss.backward() follows."""

    def __init__(self, chain):
        self.chain = chain if hasattr(chain, 'eager_rpn_backward') else None

    def __enter__(self):
        if self.chain is not None:
            self.old = self.chain.eager_rpn_backward
            self.chain.eager_rpn_backward = True

    def __exit__(self, *exc):
        if self.chain is not None:
            self.chain.eager_rpn_backward = self.old
        return False


class WeightDecay(object):
    name = 'WeightDecay'

    def __init__(self, rate):
        self.rate = rate


class MomentumSGD(object):

    def __init__(self, lr=0.01, momentum=0.9):
        self.lr = lr
        self.momentum = momentum
        self.weight_decay = 0.
        self.comm = None
        self.target = None
        self.t = 0
        self._needs_broadcast = False
        # update_overlapped: the backbone's gradient exchange goes in pieces of at least this
        # many floats (0: one piece after the backward pass, the round-2a scheme)
        self.backbone_chunk = int(float(os.environ.get('CMR_ALLREDUCE_CHUNK_MB', '16')) * (1 << 18))

    def setup(self, link):
        self.target = link
        self.ctx = link.ctx
        self.velocity = torch.zeros_like(self.ctx.train.data)
        self.broadcast_params()
        return self

    def broadcast_params(self):
        """Rank 0's trainable and frozen parameters and momenta to every rank (once)."""
        if not self._needs_broadcast or self.target is None:
            return
        import torch.distributed as dist
        for buf in (self.ctx.train.data, self.ctx.frozen.data, self.velocity):
            dist.broadcast(buf, src=dist.get_global_rank(self.comm.group, 0)
                           if self.comm.group is not None else 0, group=self.comm.group)
        self.ctx.mark_dirty()
        self._needs_broadcast = False

    def add_hook(self, hook):
        if isinstance(hook, WeightDecay):
            self.weight_decay = hook.rate
        else:
            raise TypeError('unsupported optimizer hook: {!r}'.format(hook))

    def allreduce_grad(self):
        if self.comm is not None and self.comm.size > 1:
            torch.distributed.all_reduce(self.ctx.grads, group=self.comm.group)

    def grad_buckets(self):
        """The flat gradient buffer as two views: (backbone, RPN + RoI head).  Parameters are
        laid out in creation order -- extractor, rpn, head -- and the backward pass finishes
        them in the reverse order, so the second view is final while the backbone's backward
        pass still runs."""
        store = self.ctx.train
        split = store.size
        for name, _, off in store.specs:
            if not name.startswith('extractor/'):
                split = off
                break
        return self.ctx.grads[:split], self.ctx.grads[split:]

    def update_overlapped(self, lossfun, *args, **kwds):
        """update() with the gradient exchange split in two all-reduces, the first (RPN +
        RoI head, ~3/4 of the R50 bytes) issued as soon as those gradients are final and
        running on NCCL's stream under the backbone's backward pass, the second after it;
        then the fused update.  Same result as update() (a sum is a sum).  Capturable in a
        CUDA graph: the collectives become graph nodes."""
        import torch.distributed as dist
        self.broadcast_params()
        self.ctx.grads.zero_()
        with _eager_rpn_backward(lossfun):
            loss = lossfun(*args, **kwds)
        multi = self.comm is not None and self.comm.size > 1
        backbone, heads = self.grad_buckets()
        split = backbone.numel()
        pending = []

        def after_head():
            self.ctx.finish_grads(split, None)       # (deterministic mode: fixed point -> fp32)
            if multi and heads.numel():
                pending.append(dist.all_reduce(heads, group=self.comm.group, async_op=True))

        # The backbone's bucket goes in pieces of >= backbone_chunk floats, each issued from the
        # weight-gradient side stream as soon as the blocks it covers have been enqueued (the
        # collective then waits for exactly those kernels, and the data-gradient chain on the
        # main stream is not held up): R101's 108 MB of res4 gradients are exchanged under the
        # rest of the backward pass instead of after it.
        store = self.ctx.train
        state = {'hi': split}

        def exchange(lo, hi, on_side):
            side = E.grad_side.stream if on_side and E.grad_side.active else None
            if side is not None:
                with torch.cuda.stream(side):
                    self.ctx.finish_grads(lo, hi)
                    pending.append(dist.all_reduce(self.ctx.grads[lo:hi], group=self.comm.group,
                                                   async_op=True))
            else:
                self.ctx.finish_grads(lo, hi)
                pending.append(dist.all_reduce(self.ctx.grads[lo:hi], group=self.comm.group,
                                               async_op=True))

        def progress(block):
            if not multi or self.backbone_chunk <= 0:
                return
            off = store.specs[store.index[block.conv1.W]][2]
            if state['hi'] - off >= self.backbone_chunk and off > 0:
                exchange(off, state['hi'], True)
                state['hi'] = off

        if getattr(loss, 'supports_progress', False):
            loss.backward(after_head=after_head, progress=progress)
        else:
            loss.backward(after_head=after_head)
        if multi and state['hi'] > 0:
            exchange(0, state['hi'], False)
        else:
            self.ctx.finish_grads(0, state['hi'])
        for work in pending:
            work.wait()                     # the current stream waits; the host does not
        self.apply_update()
        self.t += 1
        return loss

    def update(self, lossfun=None, *args, **kwds):
        self.broadcast_params()
        loss = None
        if lossfun is not None:
            self.ctx.grads.zero_()
            with _eager_rpn_backward(lossfun):
                loss = lossfun(*args, **kwds)
            loss.backward()
        self.allreduce_grad()
        self.apply_update()
        self.t += 1
        return loss

    def apply_update(self):
        """One fused launch: 1/world_size mean, weight decay, momentum, parameter update."""
        n = self.ctx.train.data.numel()
        scale = 1.0 / self.comm.size if self.comm is not None else 1.0
        _lib.call('cmr_sgd_momentum', E._p(self.ctx.train.data), E._p(self.ctx.grads),
                  E._p(self.velocity), n, float(self.lr), float(self.momentum),
                  float(self.weight_decay), float(scale), E._p(self.ctx.rounded), E.stream())
        self.ctx.mark_dirty(frozen=False)
        self.ctx._train_dirty = False        # the kernel refreshed the tf32 forward copy


class Communicator(object):
    """One process per GPU over torch.distributed (NCCL on the GPU box, gloo in CPU
    tests); the role of ``chainermn.create_communicator('hierarchical')``
    (examples/train_common.py:99)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.group = group
        self.size = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.intra_rank = self.rank


def create_communicator(group=None):
    return Communicator(group)


def create_multi_node_optimizer(optimizer, comm):
    """chainermn.create_multi_node_optimizer: gradients are averaged over the ranks before
    every update, and -- like chainermn's optimizer does on its first update -- rank 0's
    parameters are broadcast once (``setup`` / first ``update``), so replicas that were
    initialised with different seeds, or of which only rank 0 loaded a snapshot, agree."""
    optimizer.comm = comm
    optimizer._needs_broadcast = comm is not None and comm.size > 1
    if optimizer.target is not None:
        optimizer.broadcast_params()
    return optimizer


def shard_indices(n_items, comm_size, rank):
    """Contiguous shard of range(n_items) for `rank` (chainermn.scatter_dataset's split
    rule: sizes differ by at most one)."""
    base, rem = divmod(n_items, comm_size)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


class GraphedUpdater(object):
    """One training iteration as a CUDA-graph replay.

    The role of ``chainer.training.updaters.StandardUpdater.update_core``
    (examples/train_common.py:193-196 builds it; it calls
    ``optimizer.update(model, imgs, bboxes, labels, masks, scales)`` once per iteration),
    for the case where nothing in the step needs the host: device target creators and
    device mask targets.  The ~240 kernel launches of forward + losses + backward +
    update are captured once per (batch shape, scale, hyper-parameter) key into a
    ``torch.cuda.CUDAGraph`` and replayed; per iteration the host only copies the inputs
    into the graph's fixed input buffers.  With more than one rank the gradient
    all-reduce and the update kernel run after the replay (NCCL outside the graph).

    ``updater(imgs, bboxes, labels, masks, scales)`` -> :class:`Loss`-like object
    (``.array`` device scalar, ``.item()``).  imgs (B,3,H,W) float32 and masks
    (B,G,H,W) uint8 torch tensors or bit-packed ``models.utils.PackedMasks`` (CUDA, or
    pinned host memory for asynchronous copies);
    bboxes / labels: lists of per-image NumPy arrays.

    NumPy arrays in the reference's own batch format are accepted as well
    (``datasets.concat_examples``: imgs (B,3,H,W) float32, masks (B,G,H,W) int32 / uint8 /
    bool): they are viewed, not copied, on the host; when their memory is page-locked
    (``datasets.concat_examples(..., pinned=True)`` / ``datasets.pinned_empty``) the uploads are
    asynchronous, otherwise the driver stages them.

    The first call for a key runs eagerly (it is also the warm-up that sizes every
    workspace), the second captures and replays, later calls replay.  A key is the batch
    geometry (image and mask shapes / dtypes) and the hyper-parameters; the per-image
    scales are part of it only when the proposal layer's ``min_size`` uses them.  At most
    ``max_states`` keys are kept (least recently used first out; an evicted key's graph,
    its private memory pool and its input buffers are freed), so a data set whose batches
    come in many shapes cannot grow the device memory without bound -- pad batches to a few
    canvas sizes (``datasets.concat_examples(..., canvas=(H, W))``) to stay on replays.
    """

    def __init__(self, optimizer, lossfun, max_boxes=64, use_graph=True, max_states=4,
                 graph_allreduce=None):
        import collections
        import os
        # More than one rank: the two bucketed all-reduces and the update are captured INSIDE
        # the graph (MomentumSGD.update_overlapped), the first one overlapping the backbone's
        # backward pass.  graph_allreduce=False (or CMR_GRAPH_ALLREDUCE=0) keeps the collective
        # outside: replay, then one all-reduce over the whole buffer, then the update.
        if graph_allreduce is None:
            graph_allreduce = os.environ.get('CMR_GRAPH_ALLREDUCE', '1') != '0'
        self.graph_allreduce = bool(graph_allreduce)
        from .models.utils import GroundTruth
        self.use_graph = use_graph
        self._GroundTruth = GroundTruth
        self.optimizer = optimizer
        self.lossfun = lossfun
        self.max_boxes = max_boxes
        self.max_states = max(1, int(max_states))
        self._states = collections.OrderedDict()
        self._copy_stream = None
        self._prefetched = None
        self._seed_word = None          # ONE device word for all keys: draws never restart
        self.launches_per_replay = 0
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        self.evictions = 0

    class _State(object):
        pass

    @staticmethod
    def _mask_tensor(masks):
        return masks.data if hasattr(masks, 'width') else masks

    @staticmethod
    def _as_tensor(a):
        """NumPy arrays (the reference's batch format) become zero-copy host tensors."""
        import numpy as np
        if isinstance(a, np.ndarray):
            if a.dtype == np.bool_:
                a = a.view(np.uint8)
            return torch.from_numpy(a if a.flags.c_contiguous else np.ascontiguousarray(a))
        return a

    def _scales_matter(self):
        rpn = getattr(getattr(self.lossfun, 'mask_rcnn', None), 'rpn', None)
        layer = getattr(rpn, 'proposal_layer', None)
        return layer is None or getattr(layer, 'min_size', 1) != 0

    def _key(self, imgs, masks, scales):
        o = self.optimizer
        packed = hasattr(masks, 'width')
        masks = self._mask_tensor(masks)
        # the scales only enter the step through ProposalCreator's min_size * scale
        sc = tuple(float(s) for s in scales) if self._scales_matter() else ()
        return (tuple(imgs.shape), tuple(masks.shape), str(masks.dtype), packed, sc,
                float(o.lr), float(o.momentum), float(o.weight_decay),
                o.comm.size if o.comm is not None else 1)

    def _new_state(self, imgs, bboxes, labels, masks):
        dev = self.optimizer.ctx.device
        st = self._State()
        B = imgs.shape[0]
        mt = self._mask_tensor(masks)
        G = max(self.max_boxes, mt.shape[1], max(len(b) for b in bboxes))
        st.imgs = torch.empty(tuple(imgs.shape), dtype=torch.float32, device=dev)
        # same shape as the caller's masks: staging them is one contiguous copy
        st.masks = torch.zeros(tuple(mt.shape), dtype=mt.dtype, device=dev)
        st.masks_arg = type(masks)(st.masks, masks.width) if hasattr(masks, 'width') else st.masks
        st.gt = self._GroundTruth(bboxes, labels, dev, capacity=G)
        if self._seed_word is None:
            self._seed_word = torch.zeros((1,), dtype=torch.int64, device=dev)
        st.seed_word = self._seed_word
        st.graph = None
        st.loss = None
        st.calls = 0
        return st

    def _stage(self, st, imgs, bboxes, labels, masks):
        self.h2d_bytes = st.gt.nbytes
        for dst, src in ((st.imgs, imgs), (st.masks, self._mask_tensor(masks))):
            if not src.is_cuda:
                self.h2d_bytes += src.numel() * src.element_size()
            dst.copy_(src, non_blocking=True)
        st.gt.fill_(bboxes, labels)

    def _step(self, st, scales):
        """The work of one iteration on the current stream (eager or under capture)."""
        o, chain = self.optimizer, self.lossfun
        chain.seed_dev = st.seed_word
        chain._calls = 0        # fixed host seed: the draws advance through seed_word only
        multi = o.comm is not None and o.comm.size > 1
        try:
            if multi and self.graph_allreduce:
                loss = o.update_overlapped(chain, st.imgs, st.gt, None, st.masks_arg, scales)
                o.t -= 1                     # counted by _run
            else:
                o.ctx.grads.zero_()
                with _eager_rpn_backward(chain):
                    loss = chain(st.imgs, st.gt, None, st.masks_arg, scales)
                loss.backward()
                if not multi:
                    o.apply_update()
            st.seed_word += 1
        finally:
            chain.seed_dev = None
        return loss.array

    def _lookup(self, imgs, bboxes, labels, masks, scales):
        scales = [float(s) for s in (scales.tolist() if hasattr(scales, 'tolist') else scales)]
        if not (isinstance(imgs, torch.Tensor) and
                isinstance(self._mask_tensor(masks), torch.Tensor)):
            raise TypeError('GraphedUpdater needs arrays for imgs and masks: torch tensors '
                            '(CUDA or host), PackedMasks, or NumPy arrays')
        key = self._key(imgs, masks, scales)
        st = self._states.get(key)
        if st is None:
            while len(self._states) >= self.max_states:
                self._evict()
            st = self._states[key] = self._new_state(imgs, bboxes, labels, masks)
        else:
            self._states.move_to_end(key)
        return st, scales

    def _evict(self):
        """Drop the least recently used key: its CUDA graph (and the graph's private memory
        pool holding every activation of the step), input and staging buffers."""
        _, st = self._states.popitem(last=False)
        if self._prefetched is not None and self._prefetched[0] is st:
            self._prefetched = None
        torch.cuda.current_stream().synchronize()     # nothing of it is still running
        for name in list(vars(st)):
            if name != 'seed_word':
                setattr(st, name, None)
        self.evictions += 1

    def _run(self, st, scales):
        o = self.optimizer
        lib = _lib.load()
        o.broadcast_params()
        if o.ctx._train_dirty or o.ctx._frozen_dirty:
            # weights were loaded / edited since the last step (load_state_dict, load_npz):
            # a captured graph holds no refresh of the tf32 forward copies (the update
            # kernel maintains them from step to step), so refresh them here
            o.ctx.prepare(backward=True)
        if st.calls == 0 or not self.use_graph:
            loss = self._step(st, scales)                       # eager: warm-up + real step
        else:
            if st.graph is None:
                torch.cuda.synchronize()
                n0 = lib.cmr_launch_count()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    st.loss = self._step(st, scales)
                self.launches_per_replay = lib.cmr_launch_count() - n0
                st.graph = g
            st.graph.replay()
            loss = st.loss
        if o.comm is not None and o.comm.size > 1 and not self.graph_allreduce:
            o.allreduce_grad()
            o.apply_update()
        st.calls += 1
        o.t += 1
        self.d2h_bytes = 0
        return _GraphedLoss(loss, self)

    def __call__(self, imgs, bboxes, labels, masks, scales):
        imgs, masks = self._as_tensor(imgs), self._as_tensor(masks)
        st, scales = self._lookup(imgs, bboxes, labels, masks, scales)
        self._stage(st, imgs, bboxes, labels, masks)
        return self._run(st, scales)

    # -- input pipelining: copy iteration i+1's inputs while iteration i computes --
    def prefetch(self, imgs, bboxes, labels, masks, scales):
        """Start the host -> device copies of the NEXT iteration's inputs on a side stream
        (into staging buffers), so that they overlap the iteration that is running;
        ``step()`` then moves them into the graph's input buffers with device-to-device
        copies (tens of microseconds) and runs the iteration."""
        imgs, masks = self._as_tensor(imgs), self._as_tensor(masks)
        st, scales = self._lookup(imgs, bboxes, labels, masks, scales)
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream()
        if getattr(st, 'imgs_s', None) is None:
            st.imgs_s = torch.empty_like(st.imgs)
            st.masks_s = torch.empty_like(st.masks)
            st.gt_s = self._GroundTruth(bboxes, labels, st.imgs.device, capacity=st.gt.G)
            st.staging_free = torch.cuda.Event()
            st.staging_free.record()
            st.staged = torch.cuda.Event()
        cs = self._copy_stream
        cs.wait_event(st.staging_free)          # the previous step() has drained the staging
        self.h2d_bytes = st.gt.nbytes
        with torch.cuda.stream(cs):
            for dst, src in ((st.imgs_s, imgs), (st.masks_s, self._mask_tensor(masks))):
                if not src.is_cuda:
                    self.h2d_bytes += src.numel() * src.element_size()
                dst.copy_(src, non_blocking=True)
            st.gt_s.fill_(bboxes, labels)
            st.staged.record(cs)
        self._prefetched = (st, scales)

    def step(self):
        """Run one iteration on the inputs given to the last ``prefetch()``."""
        if self._prefetched is None:
            raise RuntimeError('step() needs a preceding prefetch()')
        (st, scales), self._prefetched = self._prefetched, None
        main = torch.cuda.current_stream()
        main.wait_event(st.staged)
        st.imgs.copy_(st.imgs_s, non_blocking=True)
        st.masks.copy_(st.masks_s, non_blocking=True)
        st.gt._buf.copy_(st.gt_s._buf, non_blocking=True)
        st.staging_free.record(main)
        return self._run(st, scales)


class _GraphedLoss(object):
    """Device scalar of the last replay (valid until the next call of the updater)."""

    def __init__(self, array, owner):
        self.array = array
        self._owner = owner

    data = property(lambda self: self.array)

    def item(self):
        self._owner.d2h_bytes += 4
        return float(self.array.item())

    __float__ = item

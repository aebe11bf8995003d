"""World-size-2 `gloo` test of the data-parallel host logic:
image sharding, the single flat-buffer gradient all-reduce and the mean it implies.

The reference wires this at examples/train_common.py:96-104 (one rank per GPU,
`batch_size_per_gpu` images each) and :176-178 (`chainermn.create_multi_node_optimizer`
= all-reduce-mean of every gradient before the MomentumSGD update).  The SGD kernel
itself is CUDA-only and is covered by the `-m gpu` tests; here the update rule is
restated in NumPy so the two-rank result can be compared with the single-process
result on the concatenated batch.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from chainer_mask_rcnn_b200 import optimizers
from chainer_mask_rcnn_b200.models import engine as E


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _momentum_sgd(p, g, v, lr, momentum, wd, grad_scale):
    """Upstream MomentumSGD + WeightDecay: the arithmetic of
    cmr_sgd_momentum (include/cmr_b200.h)."""
    g = grad_scale * g + wd * p
    v = momentum * v - lr * g
    return p + v, v


class _Ctx(object):
    """The slice of TrainContext the optimizer's exchange step touches."""

    def __init__(self, store):
        self.train = store
        self.grads = torch.zeros_like(store.data)

    def finish_grads(self, lo=0, hi=None):      # (deterministic mode only; a no-op here)
        pass


def _make_store():
    s = E.FlatStore()
    s.add('rpn/conv1/W', (8, 3, 3, 4))
    s.add('rpn/conv1/b', (8,))
    s.add('head/score/W', (5, 7))
    return s.allocate('cpu')


def _fake_grad(store, image_index):
    """A deterministic per-image 'gradient' (what one image's backward pass would add)."""
    rs = np.random.RandomState(100 + image_index)
    g = torch.zeros_like(store.data)
    for name in store.names():
        v = store.view(name, g)
        v += torch.from_numpy(rs.standard_normal(tuple(v.shape)).astype(np.float32))
    return g


def _worker(rank, world, port, n_images, out_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        comm = optimizers.create_communicator()
        assert comm.size == world and comm.rank == rank
        opt = optimizers.create_multi_node_optimizer(
            optimizers.MomentumSGD(lr=0.00125 * n_images, momentum=0.9), comm)
        store = _make_store()
        opt.ctx = _Ctx(store)
        # this rank's shard of the global batch; per-rank loss = mean over its own images
        mine = optimizers.shard_indices(n_images, comm.size, comm.rank)
        for i in mine:
            opt.ctx.grads += _fake_grad(store, i) / len(mine)
        opt.allreduce_grad()
        np.save(os.path.join(out_dir, 'grads_%d.npy' % rank), opt.ctx.grads.numpy())
        np.save(os.path.join(out_dir, 'shard_%d.npy' % rank), np.asarray(list(mine)))
    finally:
        dist.destroy_process_group()


def test_shard_indices_partition():
    for n, size in ((16, 8), (5, 2), (3, 4), (117266, 8)):
        seen = []
        sizes = []
        for r in range(size):
            idx = optimizers.shard_indices(n, size, r)
            seen.extend(idx)
            sizes.append(len(idx))
        assert seen == list(range(n))
        assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(120)
def test_two_rank_allreduce_matches_single_process(tmp_path):
    world, n_images = 2, 4
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_images, str(tmp_path)), nprocs=world, join=True)
    g0 = np.load(tmp_path / 'grads_0.npy')
    g1 = np.load(tmp_path / 'grads_1.npy')
    # every rank holds the same summed buffer after the exchange
    assert np.array_equal(g0, g1)
    shards = [np.load(tmp_path / ('shard_%d.npy' % r)) for r in range(world)]
    assert sorted(np.concatenate(shards).tolist()) == list(range(n_images))
    # the update applies grad_scale = 1/world: that must equal the single-process
    # gradient of the mean loss over the global batch
    store = _make_store()
    want = sum(_fake_grad(store, i) for i in range(n_images)).numpy() / n_images
    np.testing.assert_allclose(g0 / world, want, rtol=1e-6, atol=1e-6)
    # and the parameter after one MomentumSGD + WeightDecay step is then the same
    p = np.linspace(-1, 1, g0.size).astype(np.float32)
    v = np.zeros_like(p)
    p_dp, _ = _momentum_sgd(p, g0, v, 0.005, 0.9, 1e-4, 1.0 / world)
    p_sp, _ = _momentum_sgd(p, want, v, 0.005, 0.9, 1e-4, 1.0)
    np.testing.assert_allclose(p_dp, p_sp, rtol=1e-6, atol=1e-7)


def test_flat_store_slots_are_tma_aligned():
    s = _make_store()
    for name, shape, off in s.specs:
        assert off % E.ALIGN == 0
        assert tuple(s.view(name).shape) == tuple(shape)
    assert s.size % E.ALIGN == 0


class _FakeLoss(object):
    """A loss whose backward() fills the flat gradient buffer in the real order -- RPN and
    head parameters first, the after_head hook, then the backbone's."""

    def __init__(self, ctx, image_ids):
        self.ctx, self.ids = ctx, image_ids
        self.hook_saw_heads_only = None

    def backward(self, after_head=None):
        store = self.ctx.train
        full = sum(_fake_grad(store, i) for i in self.ids) / len(self.ids)
        heads = [n for n in store.names() if not n.startswith('extractor/')]
        for n in heads:
            store.view(n, self.ctx.grads).add_(store.view(n, full))
        if after_head is not None:
            before = self.ctx.grads.clone()
            after_head()
            self.hook_saw_heads_only = all(
                float(store.view(n, before).abs().sum()) == 0.
                for n in store.names() if n.startswith('extractor/'))
        for n in store.names():
            if n.startswith('extractor/'):
                store.view(n, self.ctx.grads).add_(store.view(n, full))


def _make_store4():
    s = E.FlatStore()
    s.add('extractor/res4/a/conv1/W', (6, 1, 1, 4))
    s.add('extractor/res4/a/conv2/W', (6, 3, 3, 6))
    s.add('rpn/conv1/W', (8, 3, 3, 4))
    s.add('head/score/W', (5, 7))
    return s.allocate('cpu')


def _worker_overlapped(rank, world, port, n_images, out_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        comm = optimizers.create_communicator()
        opt = optimizers.MomentumSGD(lr=0.005, momentum=0.9)
        store = _make_store4()
        # replicas start DIFFERENT (per-rank seed): the broadcast makes them rank 0's
        store.data.copy_(torch.from_numpy(
            np.random.RandomState(rank).standard_normal(store.data.shape).astype(np.float32)))
        ctx = _Ctx(store)
        ctx.frozen = E.FlatStore()
        ctx.frozen.add('extractor/conv1/W', (4, 7, 7, 3))
        ctx.frozen.allocate('cpu')
        ctx.frozen.data.fill_(float(rank + 1))
        ctx.mark_dirty = lambda frozen=True: None

        class Chain(object):
            pass
        chain = Chain()
        chain.ctx = ctx
        opt.setup(chain)
        opt = optimizers.create_multi_node_optimizer(opt, comm)
        assert not opt._needs_broadcast                      # done at attach time
        np.save(os.path.join(out_dir, 'p0_%d.npy' % rank), store.data.numpy().copy())
        np.save(os.path.join(out_dir, 'f0_%d.npy' % rank), ctx.frozen.data.numpy().copy())
        backbone, heads = opt.grad_buckets()
        assert backbone.numel() + heads.numel() == ctx.grads.numel()
        assert backbone.data_ptr() == ctx.grads.data_ptr() and heads.numel() > 0
        mine = list(optimizers.shard_indices(n_images, comm.size, comm.rank))
        fake = _FakeLoss(ctx, mine)
        applied = []
        opt.apply_update = lambda: applied.append(ctx.grads.clone())
        opt.update_overlapped(lambda: fake)
        assert fake.hook_saw_heads_only is True and len(applied) == 1
        np.save(os.path.join(out_dir, 'og_%d.npy' % rank), applied[0].numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_bucketed_overlapped_update_and_parameter_broadcast(tmp_path):
    """MomentumSGD.update_overlapped: two all-reduces (RPN + head bucket from the after_head
    hook, backbone bucket at the end) give the gradient sum of one all-reduce; and
    create_multi_node_optimizer broadcasts rank 0's parameters (trainable and frozen) like
    chainermn's optimizer does on its first update."""
    world, n_images = 2, 4
    port = _free_port()
    mp.spawn(_worker_overlapped, args=(world, port, n_images, str(tmp_path)), nprocs=world,
             join=True)
    g0, g1 = np.load(tmp_path / 'og_0.npy'), np.load(tmp_path / 'og_1.npy')
    assert np.array_equal(g0, g1)
    store = _make_store4()
    want = sum(_fake_grad(store, i) for i in range(n_images)).numpy() / n_images
    np.testing.assert_allclose(g0 / world, want, rtol=1e-6, atol=1e-6)
    assert np.array_equal(np.load(tmp_path / 'p0_0.npy'), np.load(tmp_path / 'p0_1.npy'))
    assert np.array_equal(np.load(tmp_path / 'f0_1.npy'), np.load(tmp_path / 'f0_0.npy'))
    assert float(np.load(tmp_path / 'f0_1.npy')[0]) == 1.0          # rank 0's value


class _Block(object):
    def __init__(self, root):
        class _C(object):
            pass
        self.conv1 = _C()
        self.conv1.W = root + '/conv1/W'


class _FakeLossWithProgress(_FakeLoss):
    """Like _FakeLoss, and reports the backbone's blocks in backward order (last created
    first) through ``progress`` -- what models.layers.BuildingBlock.backward does."""
    supports_progress = True

    def backward(self, after_head=None, progress=None):
        store = self.ctx.train
        full = sum(_fake_grad(store, i) for i in self.ids) / len(self.ids)
        for n in store.names():
            if not n.startswith('extractor/'):
                store.view(n, self.ctx.grads).add_(store.view(n, full))
        if after_head is not None:
            after_head()
        self.reported = []
        for root in ('extractor/res4/b2', 'extractor/res4/b1', 'extractor/res4/a',
                     'extractor/res3/a'):
            for n in store.names():
                if n.startswith(root + '/'):
                    store.view(n, self.ctx.grads).add_(store.view(n, full))
            if progress is not None:
                progress(_Block(root))
                self.reported.append(root)


def _make_store_blocks():
    s = E.FlatStore()
    for root in ('extractor/res3/a', 'extractor/res4/a', 'extractor/res4/b1', 'extractor/res4/b2'):
        s.add(root + '/conv1/W', (16, 1, 1, 8))
        s.add(root + '/conv2/W', (16, 3, 3, 16))
    s.add('rpn/conv1/W', (8, 3, 3, 4))
    s.add('head/score/W', (5, 7))
    return s.allocate('cpu')


def _worker_chunked(rank, world, port, n_images, out_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        comm = optimizers.create_communicator()
        opt = optimizers.MomentumSGD(lr=0.005, momentum=0.9)
        store = _make_store_blocks()
        ctx = _Ctx(store)
        ctx.frozen = E.FlatStore()
        ctx.frozen.add('extractor/conv1/W', (4, 7, 7, 3))
        ctx.frozen.allocate('cpu')
        ctx.mark_dirty = lambda frozen=True: None

        class Chain(object):
            pass
        chain = Chain()
        chain.ctx = ctx
        opt.setup(chain)
        opt = optimizers.create_multi_node_optimizer(opt, comm)
        opt.backbone_chunk = 3000          # floats: about one block (2432) -> several pieces
        calls = []
        real = dist.all_reduce

        def counting(t, *a, **k):
            calls.append(t.numel())
            return real(t, *a, **k)
        dist.all_reduce = counting
        mine = list(optimizers.shard_indices(n_images, comm.size, comm.rank))
        fake = _FakeLossWithProgress(ctx, mine)
        applied = []
        opt.apply_update = lambda: applied.append(ctx.grads.clone())
        opt.update_overlapped(lambda: fake)
        dist.all_reduce = real
        assert len(fake.reported) == 4 and len(applied) == 1
        assert len(calls) >= 3 and sum(calls) == ctx.grads.numel()     # heads + >= 2 backbone pieces
        np.save(os.path.join(out_dir, 'cg_%d.npy' % rank), applied[0].numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_chunked_backbone_exchange(tmp_path):
    """update_overlapped with ``backbone_chunk``: the backbone's gradients leave in several
    all-reduces issued from the ``progress`` hook of the backward pass; together with the head
    bucket they cover the buffer exactly once and give the sum of one all-reduce."""
    world, n_images = 2, 4
    port = _free_port()
    mp.spawn(_worker_chunked, args=(world, port, n_images, str(tmp_path)), nprocs=world, join=True)
    g0, g1 = np.load(tmp_path / 'cg_0.npy'), np.load(tmp_path / 'cg_1.npy')
    assert np.array_equal(g0, g1)
    store = _make_store_blocks()
    want = sum(_fake_grad(store, i) for i in range(n_images)).numpy() / n_images
    np.testing.assert_allclose(g0 / world, want, rtol=1e-6, atol=1e-6)

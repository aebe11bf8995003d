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
import torch

from . import _lib
from .models import engine as E


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

    def setup(self, link):
        self.target = link
        self.ctx = link.ctx
        self.velocity = torch.zeros_like(self.ctx.train.data)
        return self

    def add_hook(self, hook):
        if isinstance(hook, WeightDecay):
            self.weight_decay = hook.rate
        else:
            raise TypeError('unsupported optimizer hook: {!r}'.format(hook))

    def allreduce_grad(self):
        if self.comm is not None and self.comm.size > 1:
            torch.distributed.all_reduce(self.ctx.grads, group=self.comm.group)

    def update(self, lossfun=None, *args, **kwds):
        loss = None
        if lossfun is not None:
            self.ctx.grads.zero_()
            loss = lossfun(*args, **kwds)
            loss.backward()
        self.allreduce_grad()
        n = self.ctx.train.data.numel()
        scale = 1.0 / self.comm.size if self.comm is not None else 1.0
        _lib.call('cmr_sgd_momentum', E._p(self.ctx.train.data), E._p(self.ctx.grads),
                  E._p(self.velocity), n, float(self.lr), float(self.momentum),
                  float(self.weight_decay), float(scale), E.stream())
        self.ctx.mark_dirty(frozen=False)
        self.t += 1
        return loss


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
    optimizer.comm = comm
    return optimizer


def shard_indices(n_items, comm_size, rank):
    """Contiguous shard of range(n_items) for `rank` (chainermn.scatter_dataset's split
    rule: sizes differ by at most one)."""
    base, rem = divmod(n_items, comm_size)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))

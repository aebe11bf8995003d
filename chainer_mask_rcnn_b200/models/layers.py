"""Building blocks of the R50/R101-C4 graph on the tensor-core kernels.

``Context`` owns the flat parameter buffers; ``Conv`` is one Convolution2D (+ optional
bias, + optional frozen AffineChannel2D, links/affine_channel_2d.py:8-24) with its
forward GEMM, data-gradient GEMM and weight-gradient GEMM; ``Bottleneck`` /
``BuildingBlock`` follow chainer.links.model.vision.resnet.BuildingBlock as used at
models/mask_rcnn_resnet.py:131-133 and models/resnet_extractor.py:76-90 (stride on
the first 1x1 convolution and on the shortcut, ReLU after the addition).

The backward pass is hand-scheduled (the graph is static): a block receives the
gradient at its output already multiplied by the output's ReLU mask and rounded to
tf32, and produces the same for its input.
"""
import numpy as np
import torch

from . import engine as E

f32 = torch.float32


class Context(object):
    """Parameter storage + per-step weight preparation for one model replica."""

    def __init__(self):
        self.train = E.FlatStore()      # parameters updated by the optimizer
        self.frozen = E.FlatStore()     # conv1, res2, every AffineChannel2D
        self.kinds = {}                 # name -> ('conv'|'deconv'|'linear'|'vec', ref_shape)
        self.layers = []                # objects with prep_frozen() / prep_backward()
        self.grads = None
        self.rounded = None
        self.frozen_rounded = None
        self.device = None
        self.recording = False          # save activations for the backward pass
        # 'tf32': operands rounded to TF32 by the producing epilogue (the speed path).
        # 'tf32x3': forward GEMMs on hi/lo-split operands (3 x the work, fp32-level results;
        # the parity mode of SURVEY.md 7.3); forward only.
        self.precision = 'tf32'
        self.version = 0                # bumped whenever parameters change
        # deterministic mode: weight / bias gradients and the ROIAlign backward accumulate in
        # int64 fixed point (order-independent), see include/cmr_b200.h
        self.deterministic = False
        self.grads_fixed = None
        self._train_dirty = True
        self._frozen_dirty = True
        self._dgrad_dirty = True
        self._prep_table = None

    # -- construction -------------------------------------------------------
    def add_param(self, name, shape, kind, ref_shape, trainable):
        (self.train if trainable else self.frozen).add(name, shape)
        self.kinds[name] = (kind, tuple(ref_shape), trainable)
        return name

    def finalize(self, device):
        self.device = device
        self.train.allocate(device)
        self.frozen.allocate(device)
        self.grads = torch.zeros_like(self.train.data)
        self.rounded = torch.zeros_like(self.train.data)
        self.frozen_rounded = torch.zeros_like(self.frozen.data)
        return self

    # -- access ---------------------------------------------------------------
    def param(self, name):
        """Master fp32 value, internal layout."""
        return (self.train if name in self.train else self.frozen).view(name)

    def grad(self, name):
        """Where the backward kernels accumulate dL/d(name): the fp32 gradient buffer, or its
        int64 fixed-point shadow in deterministic mode (``finish_grads`` converts)."""
        if self.deterministic:
            if self.grads_fixed is None:
                self.grads_fixed = torch.zeros(self.grads.shape, dtype=torch.int64,
                                               device=self.device)
            return self.train.view(name, self.grads_fixed)
        return self.train.view(name, self.grads)

    def finish_grads(self, lo=0, hi=None):
        """End of the backward pass (or of the part of it that fills gradients [lo, hi)): in
        deterministic mode the fixed-point sums become the fp32 gradients (and are zeroed for
        the next step)."""
        if self.deterministic and self.grads_fixed is not None:
            hi = self.grads.numel() if hi is None else hi
            if hi > lo:
                E.fixed_to_float(self.grads_fixed[lo:hi], self.grads[lo:hi], accumulate=False,
                                 zero_src=True)

    def fwd(self, name):
        """tf32-rounded copy read by the forward GEMMs."""
        if name in self.train:
            return self.train.view(name, self.rounded)
        return self.frozen.view(name, self.frozen_rounded)

    def mark_dirty(self, frozen=True):
        self.version += 1
        self._train_dirty = True
        self._dgrad_dirty = True
        if frozen:
            self._frozen_dirty = True

    def prepare(self, backward):
        """Refresh the derived weight copies the kernels read (tf32-rounded forward
        banks; transposed, affine-scaled banks of the data-gradient GEMMs)."""
        if backward and self.precision != 'tf32':
            raise RuntimeError("precision '%s' is a forward-only parity mode; train in 'tf32'"
                               % self.precision)
        if self._frozen_dirty:
            E.round_tf32(self.frozen.data, self.frozen_rounded)
            for l in self.layers:
                l.prep_frozen()
            self._frozen_dirty = False
        if self._train_dirty:
            E.round_tf32(self.train.data, self.rounded)
            self._train_dirty = False
        if backward and self._dgrad_dirty:
            if self._prep_table is None:
                self._prep_table = E.PrepTable(self.layers, self.device)
            self._prep_table.run()
            self._dgrad_dirty = False

    # -- reference-layout import / export (Chainer npz naming) ----------------
    def to_reference(self, name):
        kind, ref_shape, _ = self.kinds[name]
        v = self.param(name)
        if kind == 'conv':                       # (O,kh,kw,I) -> (O,I,kh,kw)
            return v.permute(0, 3, 1, 2)
        if kind == 'deconv':                     # (kh*kw,O,I) -> (I,O,kh,kw)
            return v.permute(2, 1, 0).reshape(ref_shape)
        return v.view(ref_shape)

    def set_from_reference(self, name, value):
        kind, ref_shape, trainable = self.kinds[name]
        if isinstance(value, torch.Tensor):
            t = value.detach().to(self.device, torch.float32)
        else:
            t = torch.as_tensor(np.asarray(value, dtype=np.float32)).to(self.device)
        if tuple(t.shape) != ref_shape:
            raise ValueError('{}: expected shape {}, got {}'.format(name, ref_shape,
                                                                    tuple(t.shape)))
        dst = self.param(name)
        if kind == 'conv':
            dst.copy_(t.permute(0, 2, 3, 1))
        elif kind == 'deconv':
            i, o, kh, kw = ref_shape
            dst.copy_(t.reshape(i, o, kh * kw).permute(2, 1, 0))
        else:
            dst.copy_(t.view(dst.shape))
        self.mark_dirty(frozen=not trainable)

    def names(self):
        return sorted(self.kinds)


class Conv(object):
    """Convolution2D(cin, cout, k, stride, pad) [+ bias] [+ AffineChannel2D] as one
    implicit-GEMM launch with a fused epilogue."""

    def __init__(self, ctx, name, cin, cout, k, stride=1, pad=0, trainable=True, bias=False,
                 affine=None, need_dgrad=True):
        self.ctx, self.name = ctx, name
        self.cin, self.cout, self.k, self.stride, self.pad = cin, cout, k, stride, pad
        self.trainable = trainable
        self.need_dgrad = need_dgrad and trainable
        self.W = ctx.add_param(name + '/W', (cout, k, k, cin), 'conv', (cout, cin, k, k),
                               trainable)
        self.b = ctx.add_param(name + '/b', (cout,), 'vec', (cout,), trainable) if bias else None
        self.aW = self.ab = None
        if affine:
            self.aW = ctx.add_param(affine + '/W', (cout,), 'vec', (cout,), False)
            self.ab = ctx.add_param(affine + '/b', (cout,), 'vec', (cout,), False)
        self.fused_bias = None
        self.w_dgrad = None
        ctx.layers.append(self)

    # ---- derived data
    def prep_frozen(self):
        c = self.ctx
        if self.aW is not None and self.b is not None:
            # affine(conv + b) = W_a * conv + (W_a * b + b_a)
            self.fused_bias = c.param(self.aW) * c.param(self.b) + c.param(self.ab)

    def prep_backward(self):
        if not self.need_dgrad:
            return
        c = self.ctx
        T = self.k * self.k
        if self.w_dgrad is None:
            self.w_dgrad = torch.empty((self.cin, self.k, self.k, self.cout), dtype=f32,
                                       device=c.device)
        E.prep_dgrad_weight(c.param(self.W), self.cout, T, self.cin, T * self.cin, self.cin,
                            c.param(self.aW) if self.aW else None, True, self.w_dgrad)

    def _epilogue(self):
        c = self.ctx
        scale = c.param(self.aW) if self.aW else None
        if self.aW and self.b:
            bias = self.fused_bias
        elif self.aW:
            bias = c.param(self.ab)
        elif self.b:
            bias = c.param(self.b)
        else:
            bias = None
        return scale, bias

    # ---- passes
    def _w3(self):
        """[hi | hi | lo] split of the master fp32 filter bank (3 x TF32 parity mode)."""
        c = self.ctx
        if getattr(self, '_w3_cache', (None, None))[0] != c.version:
            self._w3_cache = (c.version, E.split3(c.param(self.W), order=1))
        return self._w3_cache[1]

    def forward(self, x, relu=False, addend=None, round_out=True, out=None):
        scale, bias = self._epilogue()
        if self.ctx.precision == 'tf32x3':
            return E.conv_gemm(E.split3(x), self._w3(), self.cout, self.k, self.k, self.stride,
                               self.pad, out=out, scale=scale, bias=bias, addend=addend,
                               relu=relu, round_out=False)
        return E.conv_gemm(x, self.ctx.fwd(self.W), self.cout, self.k, self.k, self.stride,
                           self.pad, out=out, scale=scale, bias=bias, addend=addend, relu=relu,
                           round_out=round_out)

    def backward_w(self, g, x):
        """Accumulate dL/dW (and dL/db) from g = dL/d(affine output), x = layer input."""
        c = self.ctx
        gw = c.grad(self.W)
        B, oh, ow, _ = g.shape
        k, cin = self.k, self.cin
        scale = c.param(self.aW) if self.aW else None
        E.wgrad_tap(g, x, gw, self.cout, cin, (oh, ow), k * k * cin, x_stride=self.stride,
                    x_off=(-self.pad, -self.pad), row_scale=scale, taps=(k, k))
        if self.b and self.trainable:
            E.column_sums(g, 0, self.cout, c.grad(self.b))

    def backward_x(self, g, in_hw, addend=None, mask=None, out=None, round_out=True):
        """dL/dx (B, in_h, in_w, cin) from g; `addend` is added and `mask` (the ReLU
        mask of x) applied in the epilogue.  For stride 2 the result is scattered to the
        even pixels of a zero-filled tensor."""
        B = g.shape[0]
        if out is None:
            if self.stride == 1:
                out = torch.empty((B, in_hw[0], in_hw[1], self.cin), dtype=f32, device=g.device)
            else:
                out = torch.zeros((B, in_hw[0], in_hw[1], self.cin), dtype=f32, device=g.device)
        return E.conv_gemm(g, self.w_dgrad, self.cin, self.k, self.k, 1, self.k - 1 - self.pad,
                           out=out, addend=addend, mask=mask, round_out=round_out,
                           d_stride=self.stride)


class Bottleneck(object):

    def __init__(self, ctx, root, cin, mid, cout, stride, is_a, trainable, need_gx=True):
        self.is_a, self.stride = is_a, stride
        self.need_gx = need_gx
        mk = lambda i, ci, co, k, s, p, dg=True: Conv(  # noqa: E731
            ctx, '%s/conv%d' % (root, i), ci, co, k, s, p, trainable,
            affine='%s/bn%d' % (root, i), need_dgrad=dg)
        self.conv1 = mk(1, cin, mid, 1, stride, 0, need_gx)
        self.conv2 = mk(2, mid, mid, 3, 1, 1)
        self.conv3 = mk(3, mid, cout, 1, 1, 0)
        self.conv4 = mk(4, cin, cout, 1, stride, 0, need_gx) if is_a else None
        self.ctx = ctx
        self.saved = None

    def forward(self, x):
        h1 = self.conv1.forward(x, relu=True)
        h2 = self.conv2.forward(h1, relu=True)
        sc = self.conv4.forward(x, round_out=False) if self.is_a else x
        y = self.conv3.forward(h2, relu=True, addend=sc)
        if self.ctx.recording:
            self.saved = (x, h1, h2, y)
        return y

    def backward(self, g, input_is_relu):
        """g = dL/dy * [y > 0].  Returns dL/dx (times [x > 0] when the block input is a
        ReLU output), or None when no input gradient is needed."""
        x, h1, h2, _ = self.saved
        self.saved = None
        x_mask = x if input_is_relu else None
        self.conv3.backward_w(g, h2)
        g2 = self.conv3.backward_x(g, h2.shape[1:3], mask=h2)
        self.conv2.backward_w(g2, h1)
        g1 = self.conv2.backward_x(g2, h1.shape[1:3], mask=h1)
        self.conv1.backward_w(g1, x)
        if self.is_a:
            self.conv4.backward_w(g, x)
        if not self.need_gx:
            return None
        hw = x.shape[1:3]
        if self.is_a:
            gx = self.conv1.backward_x(g1, hw, round_out=False)
            return self.conv4.backward_x(g, hw, addend=gx, mask=x_mask, out=gx)
        return self.conv1.backward_x(g1, hw, addend=g, mask=x_mask)


class BuildingBlock(object):
    """BuildingBlock(n_layer, in_channels, mid_channels, out_channels, stride)."""

    def __init__(self, ctx, root, n_layer, cin, mid, cout, stride, trainable=True,
                 need_gx=True):
        self.blocks = []
        self.names = ['a'] + ['b%d' % i for i in range(1, n_layer)]
        for i, nm in enumerate(self.names):
            self.blocks.append(Bottleneck(
                ctx, '%s/%s' % (root, nm), cin if i == 0 else cout, mid, cout,
                stride if i == 0 else 1, i == 0, trainable, need_gx or i > 0))

    def forward(self, x):
        for b in self.blocks:
            x = b.forward(x)
        return x

    def backward(self, g, input_is_relu, progress=None):
        """input_is_relu: whether the stage input is a ReLU output (False for the RoI
        pool feeding res5).  ``progress(block)`` is called after each block's backward pass
        has been enqueued: the gradients of that block's parameters and of everything created
        after it are then final once the weight-gradient side stream has drained."""
        for i in range(len(self.blocks) - 1, -1, -1):
            g = self.blocks[i].backward(g, input_is_relu or i > 0)
            if progress is not None:
                progress(self.blocks[i])
        return g

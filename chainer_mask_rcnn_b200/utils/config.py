"""Module-level mode switch mirroring ``chainer.config.train``.

The reference selects the train/test proposal budgets through the global
``chainer.config.train`` (set False by models/mask_rcnn.py:279,314).
"""
import contextlib

train = True


@contextlib.contextmanager
def using_config(name, value):
    g = globals()
    old = g[name]
    g[name] = value
    try:
        yield
    finally:
        g[name] = old

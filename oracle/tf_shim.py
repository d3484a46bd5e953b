"""TEST INFRASTRUCTURE ONLY - never imported by the product path.

A minimal stand-in for the `tensorflow` 1.x module, backed by torch-CPU, that
provides exactly the symbols the reference's hot path touches inside function
bodies.  With this module installed as
``sys.modules['tensorflow']`` the files /root/reference/graph.py and
/root/reference/nn.py import and execute *unmodified*, and torch autograd
supplies the backward pass that TF's autodiff supplied in the reference
(train.py:72 `optimizer.minimize`).

Semantics restated (TF 1.x docs):
  * unsorted_segment_mean: sum by id / max(count, 1); empty segment -> 0
  * scatter_nd: zeros(shape) with duplicates accumulated
  * gather_nd: only ever called with a single index column (graph.py:93,
    266-267, 390), i.e. a row gather
"""
import contextlib
import types

import numpy as np
import torch

float32 = torch.float32
float64 = torch.float64
int32 = torch.int32
int64 = torch.int64


def _t(x):
    if isinstance(x, torch.Tensor):
        return x
    if isinstance(x, np.ndarray):
        return torch.from_numpy(np.ascontiguousarray(x))
    return torch.as_tensor(x)


def _idx(i):
    return _t(i).long()


def matmul(a, b):
    return torch.matmul(_t(a), _t(b))


def einsum(eq, *ops):
    if len(ops) == 1 and isinstance(ops[0], (list, tuple)):
        ops = tuple(torch.stack([_t(o) for o in ops[0]]),)
    return torch.einsum(eq, *[_t(o) for o in ops])


def reshape(x, shape):
    return _t(x).reshape(tuple(int(s) for s in shape))


def expand_dims(x, axis):
    return _t(x).unsqueeze(axis)


def concat(xs, axis):
    return torch.cat([_t(x) for x in xs], dim=axis)


def shape(x):
    return tuple(_t(x).shape)


def gather(params, indices):
    return _t(params)[_idx(indices)]


def gather_nd(params, indices):
    ind = _idx(indices)
    assert ind.dim() == 2 and ind.shape[1] == 1, "shim supports single-column gather_nd only"
    return _t(params)[ind[:, 0]]


def scatter_nd(indices, updates, shape):
    ind = _idx(indices)
    assert ind.dim() == 2 and ind.shape[1] == 1
    upd = _t(updates)
    out = torch.zeros(tuple(int(s) for s in shape), dtype=upd.dtype)
    return out.index_add(0, ind[:, 0], upd)


def unsorted_segment_sum(data, segment_ids, num_segments):
    data = _t(data)
    ids = _idx(segment_ids)
    out = torch.zeros((int(num_segments),) + tuple(data.shape[1:]), dtype=data.dtype)
    return out.index_add(0, ids, data)


def unsorted_segment_mean(data, segment_ids, num_segments):
    data = _t(data)
    ids = _idx(segment_ids)
    s = unsorted_segment_sum(data, ids, num_segments)
    cnt = torch.zeros(int(num_segments), dtype=data.dtype).index_add(
        0, ids, torch.ones(ids.shape[0], dtype=data.dtype))
    cnt = cnt.clamp(min=1)
    return s / cnt.reshape((-1,) + (1,) * (data.dim() - 1))


def broadcast_to(x, shape):
    return _t(x).expand(tuple(int(s) for s in shape))


def add_n(xs):
    out = xs[0]
    for x in xs[1:]:
        out = out + x
    return out


def reduce_mean(x, axis=None, keepdims=False):
    x = _t(x)
    if axis is None:
        return x.mean()
    return x.mean(dim=axis, keepdim=keepdims)


def reduce_sum(x, axis=None, keepdims=False):
    x = _t(x)
    if axis is None:
        return x.sum()
    return x.sum(dim=axis, keepdim=keepdims)


def squared_difference(a, b):
    d = _t(a) - _t(b)
    return d * d


def minimum(a, b):
    return torch.minimum(_t(a), _t(b))


def sign(x):
    return torch.sign(_t(x))


def cast(x, dtype):
    return _t(x).to(dtype)


def zeros_like(x):
    return torch.zeros_like(_t(x))


@contextlib.contextmanager
def variable_scope(*_a, **_k):
    yield


AUTO_REUSE = object()

def transpose(x, perm=None):
    x = _t(x)
    return x.permute(*perm) if perm is not None else x.t()


def _leaky_relu(x, alpha=0.2):          # tf.nn.leaky_relu default alpha
    return torch.nn.functional.leaky_relu(_t(x), negative_slope=alpha)


def _softmax(x, axis=-1):
    return torch.softmax(_t(x), dim=axis)


nn = types.SimpleNamespace(relu=torch.relu, tanh=torch.tanh, leaky_relu=_leaky_relu, softmax=_softmax)


# tf.layers.batch_normalization(x) as experiment.py:142 calls it: training=False (the default), so the layer is the
# affine map gamma * (x - moving_mean) / sqrt(moving_variance + 0.001) + beta with the freshly initialised moving
# statistics (mean 0, variance 1) and TRAINABLE gamma (init 1) / beta (init 0) over the last axis.  The golden
# generator registers the (gamma, beta) pair of every call, in call order, in `layers.bn_vars`.
def _batch_normalization(x, **_k):
    x = _t(x)
    gamma, beta = layers.bn_vars[layers.bn_calls % len(layers.bn_vars)]
    layers.bn_calls += 1
    return gamma * (x / float(np.sqrt(np.float32(1.0) + np.float32(0.001)))) + beta


layers = types.SimpleNamespace(batch_normalization=_batch_normalization, bn_vars=[], bn_calls=0)


# initialisers referenced at import time by utils.py:171-174 (never called by the
# hot path; present so attribute lookups do not fail if someone imports utils).
def _unsupported(*_a, **_k):  # pragma: no cover
    raise NotImplementedError("tf_shim: not part of the hot path")


random_uniform_initializer = _unsupported
random_normal_initializer = _unsupported
glorot_uniform_initializer = _unsupported
glorot_normal_initializer = _unsupported


# make Tensor.get_shape().as_list() work (nn.py:108, get_readout)
class _Shape(list):
    def as_list(self):
        return list(self)


if not hasattr(torch.Tensor, "get_shape"):
    torch.Tensor.get_shape = lambda self: _Shape(self.shape)  # type: ignore[attr-defined]

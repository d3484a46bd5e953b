"""Drop-in for the hot-path functions of /root/reference/nn.py: set layer / set network, readout and
losses - same names and signatures, torch CUDA tensors in and out, hand-written sm_100a kernels
underneath (libnbpc.so).  No CPU fallback."""
import numpy as np
import torch

from . import ops
from .graph import _is_relu, _to_cuda

__all__ = ["set_layer", "network_func_set", "model_func_set", "get_readout", "periodic_boundary_dist",
           "pbc_loss", "loss_ZA", "mse_za", "get_init_pos"]


def _set_layer(h_in, layer_vars, relu, input_relu=False, grad_premasked=False, chain=None, idx=0, last=True):
    W, B = layer_vars
    W = W[0]  # only one weight for set layer (nn.py:22)
    return ops.SetLayer.apply(_to_cuda(h_in, torch.float32), W, B, relu, input_relu, grad_premasked, chain, idx, last)


def set_layer(h_in, layer_vars):
    """nn.py:10-28: (b,N,D) -> (b,N,Q),  (H - mean_N H) W[0] + B."""
    return _set_layer(h_in, layer_vars, False)


def network_func_set(X_in, model_vars):
    """nn.py:31-67.  A ReLU activation is fused into the layer kernel."""
    num_layers = model_vars.num_layers
    activation = model_vars.activation
    get_layer_vars = model_vars.get_layer_vars
    fuse = _is_relu(activation)
    # inside this function every hidden tensor has exactly one consumer (the next layer): the ReLU backward of layer l is
    # applied by layer l+1's backward kernel (input_relu) and layer l skips its own mask (grad_premasked); the kernel that
    # writes a hidden tensor (or its gradient) also hands its per-sample column sums to that consumer (ops.SetChain)
    chain = fuse and num_layers > 1
    sc = ops.SetChain() if chain else None
    H = _set_layer(X_in, get_layer_vars(0), fuse, input_relu=False, grad_premasked=chain, chain=sc, idx=0, last=num_layers == 1)
    if not fuse:
        H = activation(H)
    for layer_idx in range(1, num_layers):
        is_last = layer_idx >= num_layers - 1
        H = _set_layer(H, get_layer_vars(layer_idx), fuse and not is_last, input_relu=fuse, grad_premasked=fuse and not is_last,
                       chain=sc, idx=layer_idx, last=is_last)
        if not is_last and not fuse:
            H = activation(H)
    return H


def model_func_set(X_in, model_vars):
    """nn.py:70-97"""
    return network_func_set(X_in, model_vars)


def get_readout(h_out):
    """nn.py:107-119: wrap the first three channels into [0,1) (reference's literal formula)."""
    return ops.Readout.apply(_to_cuda(h_out, torch.float32))


def periodic_boundary_dist(readout_full, x_truth):
    """nn.py:123-134 (forward only; use pbc_loss for gradients)."""
    return ops.periodic_boundary_dist(_to_cuda(readout_full, torch.float32), _to_cuda(x_truth, torch.float32))


def pbc_loss(x_pred, x_truth, scale_error=True):
    """nn.py:137-148"""
    return ops.Loss.apply(_to_cuda(x_pred, torch.float32), _to_cuda(x_truth, torch.float32), True, bool(scale_error))


def loss_ZA(predicted_error, true_error):
    """nn.py:151-166"""
    return ops.Loss.apply(_to_cuda(predicted_error, torch.float32), _to_cuda(true_error, torch.float32), False, False)


# ---- NumPy helpers kept for script compatibility (nn.py:177-189; host-side, not on the hot path)
def mse_za(fpm_displacement, za_displacement):
    err_diff = np.square(fpm_displacement - za_displacement)
    return np.mean(np.sum(err_diff, axis=-1))


def get_init_pos(za_disp):
    b, N, k = za_disp.shape
    mg = range(2, 130, 4)
    q = np.einsum('ijkl->kjli', np.array(np.meshgrid(mg, mg, mg)))
    return za_disp + q.reshape(-1, 3)

"""TEST INFRASTRUCTURE ONLY - the CPU oracle for the layer half of the path.

A torch-CPU restatement of the reference's TensorFlow-1.x layer functions, op for
op in the reference's order (so float32 rounding matches the reference run through
oracle/tf_shim.py), each citing the /root/reference file:line it follows.  torch
autograd plays the role of TF autodiff (train.py:72).  Works in float32 or float64
depending on the input dtype.

Parity pinned: tests/test_oracle_golden.py checks every function here against golden
vectors produced by running the unmodified reference (oracle/make_golden.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.
"""
import numpy as np
import torch


def _t(x):
    if isinstance(x, np.ndarray):
        return torch.from_numpy(np.ascontiguousarray(x))
    return x


def _segment_mean(h, ids, num_segs):
    """tf.unsorted_segment_mean: sum / max(count,1); empty segments give 0."""
    ids = _t(ids).long()
    s = torch.zeros((num_segs,) + tuple(h.shape[1:]), dtype=h.dtype).index_add(0, ids, h)
    cnt = torch.zeros(num_segs, dtype=h.dtype).index_add(0, ids, torch.ones(ids.shape[0], dtype=h.dtype))
    return s / cnt.clamp(min=1).reshape(-1, 1)


# ------------------------------------------------------------------ input features
def include_node_features(X_in_edges, X_in_nodes, COO_feats, redshift=None):
    """graph.py:245-275"""
    row_idx = _t(COO_feats[0]).long()
    col_idx = _t(COO_feats[1]).long()
    node_rows = X_in_nodes[row_idx]
    node_cols = X_in_nodes[col_idx]
    X_in = torch.cat([X_in_edges, node_rows, node_cols], dim=1)
    if redshift is not None:
        X_in = torch.cat([X_in, redshift], dim=1)
    return X_in


def get_input_features_shift_inv_ZA(init_pos, ZA_displacement, coo, diag, dims):
    """graph.py:289-343"""
    b, N, M = dims
    flattened_pos = init_pos.reshape(-1, 3)
    cols = _t(coo[1]).long()
    edges = flattened_pos[cols].reshape(b, N, M, 3)
    edges = edges - init_pos.unsqueeze(2)
    flattened_za_disp = ZA_displacement.reshape(-1, 3)
    diagonal_za = torch.zeros((b * N * M, 3), dtype=init_pos.dtype).index_add(
        0, _t(diag).long(), flattened_za_disp)
    return edges.reshape(-1, 3) + diagonal_za


def get_input_features_shift_inv(X_in, coo, dims):
    """graph.py:346-364 (raw differences - no minimum-image wrap)"""
    b, N, M = dims
    X = X_in.reshape(-1, 6)
    edges = X[..., :3]
    nodes = X[..., 3:]
    cols = _t(coo[1]).long()
    edges = edges[cols].reshape(b, N, M, 3)
    edges = edges - X_in[..., :3].unsqueeze(2)
    return edges.reshape(-1, 3), nodes


# ------------------------------------------------------------------ graph layer
def shift_inv_conv(h, pool_idx, num_segs, broadcast):
    """graph.py:367-391"""
    pooled = _segment_mean(h, pool_idx, num_segs)
    if broadcast:
        pooled = pooled[_t(pool_idx).long()]
    return pooled


def shift_inv_layer(H_in, COO_feats, bN, layer_vars, is_last=False):
    """graph.py:394-456"""
    b, N = bN
    row_idx, col_idx, cube_idx = COO_feats[0], COO_feats[1], COO_feats[2]
    weights, B = layer_vars
    W1, W2, W3, W4 = weights

    def _pool(H, idx, broadcast=True):
        return shift_inv_conv(H, idx, b * N, broadcast)

    H1 = H_in @ W1
    H2 = _pool(H_in, col_idx) @ W2
    H3 = _pool(H_in, row_idx) @ W3
    H4 = _pool(H_in, cube_idx) @ W4
    H_out = (H1 + H2 + H3 + H4) + B
    if is_last:
        H_out = _pool(H_out, row_idx, broadcast=False).reshape(b, N, -1)
    return H_out


def network_func_shift_inv_za(edges, coo, num_layers, dims, activation, model_vars):
    """graph.py:463-476"""
    H = activation(shift_inv_layer(edges, coo, dims, model_vars.get_layer_vars(0)))
    for layer_idx in range(1, num_layers):
        is_last = layer_idx == num_layers - 1
        H = shift_inv_layer(H, coo, dims, model_vars.get_layer_vars(layer_idx), is_last=is_last)
        if not is_last:
            H = activation(H)
    return H


def model_func_shift_inv_za(init_pos, COO_feats, ZA_displacement, ZA_diagonal, model_vars, dims,
                            activation=torch.relu):
    """graph.py:479-515"""
    num_layers = len(model_vars.channels) - 1
    edges = get_input_features_shift_inv_ZA(init_pos, ZA_displacement, COO_feats, ZA_diagonal, dims)
    return network_func_shift_inv_za(edges, COO_feats, num_layers, dims[:-1], activation, model_vars)


def network_func_shift_inv(X_in_edges, X_in_nodes, COO_feats, num_layers, dims, activation, model_vars, redshift=None):
    """graph.py:517-533 - the multi-redshift network.  The reference keeps this function inside a commented-out block
    (graph.py:516-567); it is restated from that text, on top of the live include_node_features (graph.py:245-275)
    and shift_inv_layer (graph.py:394-456)."""
    H_in = include_node_features(X_in_edges, X_in_nodes, COO_feats, redshift=redshift)
    H = activation(shift_inv_layer(H_in, COO_feats, dims, model_vars.get_layer_vars(0)))
    for layer_idx in range(1, num_layers):
        is_last = layer_idx == num_layers - 1
        H = shift_inv_layer(H, COO_feats, dims, model_vars.get_layer_vars(layer_idx), is_last=is_last)
        if not is_last:
            H = activation(H)
    return H


def model_func_shift_inv(X_in, COO_feats, model_vars, dims, activation=torch.relu, redshift=None):
    """graph.py:536-567 (commented-out block, see network_func_shift_inv): X_in (b,N,6) = [position, velocity] ->
    (b,N,6|3): loc' = net[:3]*loc_scalar + loc + vel*vel_scalar, vel' = net[3:]*vel_scalar + vel."""
    num_layers = len(model_vars.channels) - 1
    edges, nodes = get_input_features_shift_inv(X_in, COO_feats, dims)
    X_in_loc, X_in_vel = X_in[..., :3], X_in[..., 3:]
    net_out = network_func_shift_inv(edges, nodes, COO_feats, num_layers, dims[:-1], activation, model_vars, redshift)
    loc_scalar, vel_scalar = model_vars.get_scalars()
    H_out = net_out[..., :3] * loc_scalar + X_in_loc + X_in_vel * vel_scalar
    if net_out.shape[-1] > 3:
        H_vel = net_out[..., 3:] * vel_scalar + X_in_vel
        H_out = torch.cat([H_out, H_vel], dim=-1)
    return H_out


def shift_inv_15op_layer(H_in, adj, bN, layer_vars, is_last=False):
    """graph.py:20-200, op for op (15 projections after broadcasting, two biases)."""
    b, N = bN
    W, B = layer_vars
    S, q = H_in.shape[0], W[0].shape[-1]
    idx = {k: _t(np.asarray(v)).long() for k, v in adj.items()}

    def pool(h, name, nseg):
        return _segment_mean(h, idx[name], nseg)

    def to_diag(h):                                          # tf.scatter_nd (graph.py:106)
        return torch.zeros((S, q), dtype=h.dtype).index_add(0, idx["dia"], h)

    H_all = [H_in @ W[0], H_in[idx["tra"]] @ W[1]]
    Hd = H_in[idx["dia"]]
    H_all.append(to_diag(Hd @ W[2]))
    Hr = pool(H_in, "col", b * N)
    H_all += [(Hr @ W[3])[idx["col"]], (Hr @ W[4])[idx["row"]], to_diag(Hr @ W[5])]
    Hc = pool(H_in, "row", b * N)
    H_all += [(Hc @ W[6])[idx["row"]], (Hc @ W[7])[idx["col"]], to_diag(Hc @ W[8])]
    Ha = pool(H_in, "all", b)
    H_all += [(Ha @ W[9])[idx["all"]], to_diag((Ha @ W[10])[idx["dal"]])]
    Hp = _segment_mean(Hd, idx["dal"], b)
    H_all += [(Hp @ W[11])[idx["all"]], to_diag((Hp @ W[12])[idx["dal"]])]
    H_all += [(Hd @ W[13])[idx["col"]], (Hd @ W[14])[idx["row"]]]
    B_diag = to_diag(B[0].expand(b * N, q))
    H = sum(H_all) + B_diag + B[1]
    if is_last:
        return pool(H, "row", b * N).reshape(b, N, -1)
    return H


def network_func_15op_shift_inv_za(edges, adj, num_layers, dims, activation, sess_mgr):
    """graph.py:202-216"""
    H = activation(shift_inv_15op_layer(edges, adj, dims, sess_mgr.get_layer_vars(0)))
    for layer_idx in range(1, num_layers):
        is_last = layer_idx == num_layers - 1
        H = shift_inv_15op_layer(H, adj, dims, sess_mgr.get_layer_vars(layer_idx), is_last=is_last)
        if not is_last:
            H = activation(H)
    return H


# ------------------------------------------------------------------ set layer
def set_layer(h_in, layer_vars):
    """nn.py:10-28 (only W[0] of the layer's weights is used, nn.py:22)"""
    W, B = layer_vars
    W = W[0]
    h_mu = h_in.mean(dim=1, keepdim=True)
    h = h_in - h_mu
    return torch.einsum('bnk,kq->bnq', h, W) + B


def network_func_set(X_in, model_vars):
    """nn.py:31-67"""
    num_layers = model_vars.num_layers
    activation = model_vars.activation
    H = activation(set_layer(X_in, model_vars.get_layer_vars(0)))
    for layer_idx in range(1, num_layers):
        H = set_layer(H, model_vars.get_layer_vars(layer_idx))
        if not layer_idx >= num_layers - 1:
            H = activation(H)
    return H


def model_func_set(X_in, model_vars):
    """nn.py:70-97"""
    return network_func_set(X_in, model_vars)


# ------------------------------------------------------------------ readout / losses
def get_readout(h_out):
    """nn.py:107-119"""
    M = h_out.shape[-1]
    c = h_out[..., :3]
    gt_one = (torch.sign(c - 1) + 1) / 2
    ls_zero = -(torch.sign(c) - 1) / 2
    rest = 1 - gt_one - ls_zero
    readout = rest * c + gt_one * (c - 1) + ls_zero * (1 + c)
    if M > 3:
        readout = torch.cat([readout, h_out[..., 3:]], dim=-1)
    return readout


def periodic_boundary_dist(readout_full, x_truth):
    """nn.py:123-134"""
    readout = readout_full[..., :3]
    t = x_truth[..., :3]
    d1 = (readout - t) ** 2
    d2 = (readout - (1 + t)) ** 2
    d3 = ((1 + readout) - t) ** 2
    return torch.minimum(torch.minimum(d1, d2), d3)


def pbc_loss(x_pred, x_truth, scale_error=True):
    """nn.py:137-148"""
    error = periodic_boundary_dist(x_pred, x_truth).sum(dim=-1).mean()
    if scale_error:
        error = error * 1e5
    return error


def loss_ZA(predicted_error, true_error):
    """nn.py:151-166"""
    d = predicted_error - true_error
    return (d * d).sum(dim=-1).mean()

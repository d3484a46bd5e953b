"""Synthetic particle boxes with the reference's array layout.

The reference trains on `ZA_###.npy` cubes that are not shipped
(/root/reference/utils.py:518-630); these generators produce inputs of the
same shapes/dtypes:  positions in the unit periodic box, float32, seeded with
``numpy.random.default_rng``.  NumPy only - no device code here.
"""
import numpy as np

__all__ = ["uniform_box", "clustered_box", "lattice_box", "make_box",
           "za_features", "glorot_params", "DEFAULT_GRAPH_CHANNELS", "PARAMS_SEED"]

# canonical benchmark net: utils.py:163 list with the 8-wide layer dropped
DEFAULT_GRAPH_CHANNELS = [3, 32, 16, 3]
PARAMS_SEED = 77743196  # utils.py:161


def uniform_box(b, N, seed=0):
    rng = np.random.default_rng(seed)
    return rng.random((b, N, 3)).astype(np.float32)


def clustered_box(b, N, seed=0, n_clumps=64, sigma=0.02):
    """Half the particles uniform, half in `n_clumps` Gaussian clumps (wrapped)."""
    rng = np.random.default_rng(seed)
    out = np.empty((b, N, 3), dtype=np.float32)
    n_uni = N // 2
    n_cl = N - n_uni
    for i in range(b):
        uni = rng.random((n_uni, 3))
        centres = rng.random((n_clumps, 3))
        which = rng.integers(0, n_clumps, size=n_cl)
        pts = centres[which] + sigma * rng.standard_normal((n_cl, 3))
        pts = pts % 1.0
        x = np.concatenate([uni, pts], axis=0)
        x = x[rng.permutation(N)].astype(np.float32)
        # float32 rounding of values just below 1.0 can give exactly 1.0; keep [0,1)
        x[x >= 1.0] = np.float32(0.0)
        out[i] = x
    return out


def lattice_box(b, n_side, seed=0, jitter=0.01):
    """Reference lattice q = meshgrid(range(2,130,4)) (utils.py:611-613; nn.py:183-189)
    generalised to n_side points per axis, scaled to the unit box, plus N(0, jitter^2)
    displacement.  Tie-heavy when jitter == 0."""
    rng = np.random.default_rng(seed)
    step = 1.0 / n_side
    mg = (np.arange(n_side) + 0.5) * step
    q = np.einsum('ijkl->kjli', np.array(np.meshgrid(mg, mg, mg))).reshape(-1, 3)
    x = q[None] + jitter * rng.standard_normal((b, n_side ** 3, 3))
    x = (x % 1.0).astype(np.float32)
    x[x >= 1.0] = np.float32(0.0)
    return x


def make_box(kind, b, N, seed=0):
    if kind == "uniform":
        return uniform_box(b, N, seed)
    if kind == "clustered":
        return clustered_box(b, N, seed)
    if kind == "lattice":
        n_side = round(N ** (1.0 / 3.0))
        assert n_side ** 3 == N
        return lattice_box(b, n_side, seed)
    raise ValueError(kind)


def za_features(b, N, seed=0, scale=0.01, cols=3):
    """ZA displacement and regression target ~ N(0, scale^2), shape (b, N, cols) each."""
    rng = np.random.default_rng(seed + 1000003)
    za = (scale * rng.standard_normal((b, N, cols))).astype(np.float32)
    tgt = (scale * rng.standard_normal((b, N, cols))).astype(np.float32)
    return za, tgt


def glorot_params(channels, n_w=4, n_b=1, seed=PARAMS_SEED, dtype=np.float32):
    """Glorot-normal weights (sigma = sqrt(2/(k+q))) and bias 1e-8
    (utils.py:324-358: `init_weight`, `init_bias`).  Returns a list over layers of
    ([W_0..W_{n_w-1}] each (k,q), B (q,)) - or B of shape (n_b, q) when n_b > 1."""
    rng = np.random.default_rng(seed)
    out = []
    for k, q in zip(channels[:-1], channels[1:]):
        sigma = np.sqrt(2.0 / (k + q))
        Ws = [(sigma * rng.standard_normal((k, q))).astype(dtype) for _ in range(n_w)]
        if n_b == 1:
            B = np.full((q,), 1e-8, dtype=dtype)
        else:
            B = np.full((n_b, q), 1e-8, dtype=dtype)
        out.append((Ws, B))
    return out

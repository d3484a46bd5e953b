"""Host-side glue the entry points need around the hot path: parameter store with the reference's
`get_layer_vars(i) -> ([W0..W3], B)` contract (utils.py:292-385), the `ModelVars` tuple
(utils.py:199-202), TF-form Adam on a flat buffer (train.py:70) and sample-sharded data-parallel
gradient all-reduce (one NCCL all-reduce of the flat gradient buffer per step)."""
import math
from collections import namedtuple

import torch

from . import _lib, ops
from .synthetic import PARAMS_SEED

ModelVars = namedtuple("ModelVars", ["num_layers", "get_layer_vars", "activation"])  # utils.py:199-202


class ParamStore:
    """All layer weights/biases as views into ONE flat float32 buffer (so the gradient all-reduce and
    the Adam step are a single kernel each).  Layout per layer: n_w weights (k,q) then the bias (q)."""

    def __init__(self, channels, n_w=4, seed=PARAMS_SEED, device="cuda", var_scope="params"):
        self.channels = list(channels)
        self.var_scope = var_scope
        self.num_layers = len(channels) - 1
        self.n_w = n_w
        sizes = []
        for k, q in zip(channels[:-1], channels[1:]):
            sizes.append(n_w * k * q + q)
        self.flat = torch.zeros(sum(sizes), dtype=torch.float32, device=device)
        self._flat_grad = torch.zeros_like(self.flat)
        self._loose = False                                   # True: the .grad tensors are autograd's own (see zero_grad)
        gen = torch.Generator(device="cpu").manual_seed(int(seed))
        self._W, self._B, self._grad_views = [], [], []
        off = 0
        for (k, q), _ in zip(zip(channels[:-1], channels[1:]), sizes):
            W = self.flat[off:off + n_w * k * q].view(n_w, k, q)
            Wg = self._flat_grad[off:off + n_w * k * q].view(n_w, k, q)
            off += n_w * k * q
            Bv = self.flat[off:off + q]
            Bg = self._flat_grad[off:off + q]
            off += q
            sigma = math.sqrt(2.0 / (k + q))   # glorot normal (utils.py:178, 357)
            init = torch.empty(n_w, k, q).normal_(0.0, sigma, generator=gen)
            W.copy_(init)
            Bv.fill_(1e-8)                     # utils.py:334
            W.requires_grad_(True)
            Bv.requires_grad_(True)
            W.grad, Bv.grad = Wg, Bg           # views of the flat gradient buffer
            self._W.append(W)
            self._B.append(Bv)
            self._grad_views.append((Wg, Bg))
        self.m = torch.zeros_like(self.flat)
        self.v = torch.zeros_like(self.flat)
        self.step_count = 0                                  # AdamTF.step(): count kept on the host
        self.step_dev = torch.zeros(1, dtype=torch.int64, device=device)   # AdamTF.step_dev(): count in device memory

    def get_layer_vars(self, i):
        """([W_0..W_{n_w-1}] as one (n_w,k,q) tensor - indexable like the reference's list -, B)."""
        return self._W[i], self._B[i]

    def model_vars(self, activation=torch.relu):
        return ModelVars(self.num_layers, self.get_layer_vars, activation)

    def load_numpy(self, params):
        """params: list over layers of ([W...], B) NumPy arrays (synthetic.glorot_params)."""
        with torch.no_grad():
            for i, (Ws, B) in enumerate(params):
                for j, w in enumerate(Ws):
                    self._W[i][j].copy_(torch.as_tensor(w))
                self._B[i].copy_(torch.as_tensor(B).reshape(-1))

    def zero_grad(self):
        """Forget the gradients.  No kernel runs: the .grad tensors are dropped, so the next backward makes autograd ADOPT the
        tensors the layer kernels wrote (an existing .grad would cost one accumulate kernel per parameter plus the zero fill);
        `pack_grads` - called by every reader of `flat_grad` - gathers them into the flat buffer with one concatenation."""
        for p in self._W + self._B:
            p.grad = None
        self._loose = True

    def pack_grads(self):
        if not self._loose:
            return
        self._loose = False
        parts, views = [], []
        for W, B, (Wg, Bg) in zip(self._W, self._B, self._grad_views):
            for p, v in ((W, Wg), (B, Bg)):
                parts.append((p.grad if p.grad is not None else torch.zeros_like(v)).reshape(-1))
                views.append((p, v))
        with torch.no_grad():
            torch.cat(parts, out=self._flat_grad)
        for p, v in views:
            p.grad = v

    @property
    def flat_grad(self):
        self.pack_grads()
        return self._flat_grad

    @flat_grad.setter
    def flat_grad(self, value):                 # `store.flat_grad += x` assigns the same tensor back
        if value is not self._flat_grad:
            self.pack_grads()
            self._flat_grad.copy_(value)

    def parameters(self):
        return self._W + self._B


class AdamTF:
    """tf.train.AdamOptimizer(lr) (train.py:70): lr_t = lr*sqrt(1-b2^t)/(1-b1^t); p -= lr_t*m/(sqrt(v)+eps)."""

    def __init__(self, store, lr=0.01, beta1=0.9, beta2=0.999, eps=1e-8):
        self.store, self.lr, self.beta1, self.beta2, self.eps = store, lr, beta1, beta2, eps

    def step(self, grad_scale=1.0):
        s = self.store
        s.step_count += 1
        ops.adam_tf_(s.flat, s.flat_grad, s.m, s.v, s.step_count, self.lr, self.beta1, self.beta2, self.eps, grad_scale)

    def step_dev(self, grad_scale=1.0):
        """Same update with the step count in device memory (CUDA-graph capturable; do not mix with step())."""
        s = self.store
        ops.adam_tf_dev_(s.flat, s.flat_grad, s.m, s.v, s.step_dev, self.lr, self.beta1, self.beta2, self.eps, grad_scale)


class GraphedStep:
    """One whole training step (kNN graph build, forward, backward, gradient all-reduce, Adam) captured ONCE in a CUDA
    graph and replayed: ~60 kernel launches become one graph launch, so the step is not limited by the host's launch
    rate.  `step_fn(*static_inputs) -> loss` must be shape-static, sync-free and use AdamTF.step_dev (the reference's
    loop is: same shapes every iteration, train.py:87-120).  Inputs are copied into the static tensors before a replay.
        gs = GraphedStep(step_fn, example_inputs);  loss = gs(x, za, target)      # loss: static 0-d tensor"""

    def __init__(self, step_fn, example_inputs, warmup=3):
        self.static_in = tuple(torch.empty_like(t).copy_(t) for t in example_inputs)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                      # lazy initialisations must not happen inside the capture
            for _ in range(warmup):
                step_fn(*self.static_in)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        n0 = _lib.launch_count()
        with torch.cuda.graph(self.graph):
            self.loss = step_fn(*self.static_in)
        self.kernels_per_replay = _lib.launch_count() - n0     # libnbpc kernels inside one replay

    def __call__(self, *inputs):
        for d, s in zip(self.static_in, inputs):
            if d.data_ptr() != s.data_ptr():
                d.copy_(s, non_blocking=True)
        self.graph.replay()
        return self.loss

    def close(self):
        """Drop the graph (required before destroy_process_group() if it holds NCCL kernels)."""
        self.graph = None
        torch.cuda.synchronize()


class PipelinedStep:
    """Data-parallel training step captured as ONE CUDA graph, NCCL all-reduce included, with the collective hidden
    behind work that does not need the parameters:

        replay i:   side stream : all-reduce(gradient of step i-1) -> Adam update of step i-1
                    main stream : prep_fn(*inputs)        (kNN graph build, adjacency, edge features: no parameters)
                    join        : grad_fn(ctx, *inputs)   (forward, loss, backward -> gradient of step i)

    The sequence of parameter values is exactly that of the plain loop (update i-1 lands before forward i); the update of
    the LAST step is applied by flush().  The Adam step counter starts at -1, which makes the update of the first replay a
    no-op (nbpc_adam_tf_dev).  prep_fn(*static_inputs) -> ctx; grad_fn(ctx, *static_inputs) -> loss, and must zero the
    gradient buffer before its backward.  close() must be called before torch.distributed.destroy_process_group():
    destroying a process group while a live CUDA graph holds its NCCL kernels hangs."""

    def __init__(self, prep_fn, grad_fn, store, adam, world, example_inputs, warmup=2):
        self.store, self.adam, self.world = store, adam, world
        self.static_in = tuple(torch.empty_like(t).copy_(t) for t in example_inputs)
        self.side = torch.cuda.Stream()
        state = [t.clone() for t in (store.flat, store.m, store.v)]

        def body():
            main = torch.cuda.current_stream()
            self.side.wait_stream(main)
            with torch.cuda.stream(self.side):
                allreduce_gradients(store, world)
                adam.step_dev(grad_scale=1.0 / world)
            ctx = prep_fn(*self.static_in)
            main.wait_stream(self.side)
            out = grad_fn(ctx, *self.static_in)
            store.pack_grads()                             # the next body's all-reduce reads the flat buffer of THIS capture
            return out

        warm = torch.cuda.Stream()
        warm.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(warm):                      # lazy initialisations (NCCL communicator included) outside the capture
            for _ in range(max(warmup, 1)):
                body()
        torch.cuda.current_stream().wait_stream(warm)
        torch.cuda.synchronize()
        with torch.no_grad():                              # the warm-up steps trained on the example batch: undo
            for t, s0 in zip((store.flat, store.m, store.v), state):
                t.copy_(s0)
            store.flat_grad.zero_()
            store.step_dev.fill_(-1)
        self.graph = torch.cuda.CUDAGraph()
        n0 = _lib.launch_count()
        with torch.cuda.graph(self.graph):
            self.loss = body()
        self.kernels_per_replay = _lib.launch_count() - n0

    def __call__(self, *inputs):
        for d, s in zip(self.static_in, inputs):
            if d.data_ptr() != s.data_ptr():
                d.copy_(s, non_blocking=True)
        self.graph.replay()
        return self.loss

    def flush(self):
        """Apply the update of the last replayed step (eagerly)."""
        allreduce_gradients(self.store, self.world)
        self.adam.step_dev(grad_scale=1.0 / self.world)
        self.store.flat_grad.zero_()

    def close(self):
        self.graph = None
        torch.cuda.synchronize()


class OverlappedStep:
    """Training loop software-pipelined ACROSS steps: the neighbour graph of batch i (kNN build, adjacency, CSR transpose - no
    parameters involved, issue-bound) is built on a second stream while the forward / backward / all-reduce / Adam of batch
    i-1 (HBM-bound) runs on the first, so the two share the SMs instead of taking turns:

        call i:   main stream (high priority): inputs_i -> slot i%2;  T[(i-1)%2]: grad_fn(ctx, batch i-1) -> all-reduce -> Adam
                  side stream (low priority) : P[i%2]: prep_fn(batch i) -> ctx[i%2]

    Four CUDA graphs (P and T for each of the two slots), ordered by events: T waits for the P of its slot (previous call),
    P waits until the T that last read its slot has been enqueued-and-finished (stream order on main + the `staged` event).
    Every call does one full step's work - one graph build and one parameter update - and the sequence of parameter values
    is exactly that of the plain loop; the call returns the loss of batch i-1 (None on the first call) and flush() trains on
    the last batch.  prep_fn(*static_inputs) -> ctx; grad_fn(ctx, *static_inputs) -> loss (zeroes the gradients before its
    backward).  close() must precede destroy_process_group() (the T graphs hold the NCCL kernels).

    head_fn (optional): head_fn(ctx, *static_inputs) -> head is the parameter-FREE start of a training step (the edge input
    features); grad_fn is then called as grad_fn(ctx, head, *static_inputs), and the all-reduce + Adam update of batch i-2 are
    captured at the START of T[(i-1)%2] on a forked stream next to head_fn instead of at the end of the previous T, where
    every rank would sit in the collective with nothing to run (same parameter sequence: the update still lands before the
    first kernel that reads a parameter; the device-side step counter starts at -1 so that the first update is a no-op, and
    flush() applies the last one)."""

    def __init__(self, prep_fn, grad_fn, store, adam, world, example_inputs, warmup=2, head_fn=None):
        self.store, self.adam, self.world = store, adam, world
        self.deferred = head_fn is not None
        self.upd = torch.cuda.Stream(priority=-1)
        self.static_in = [tuple(torch.empty_like(t).copy_(t) for t in example_inputs) for _ in range(2)]
        self.side = torch.cuda.Stream(priority=0)
        cap_hi, cap_lo = torch.cuda.Stream(priority=-1), torch.cuda.Stream(priority=0)
        state = [t.clone() for t in (store.flat, store.m, store.v, store.step_dev)]

        def train(ctx, ins):
            if head_fn is None:
                out = grad_fn(ctx, *ins)
                allreduce_gradients(store, world)
                adam.step_dev(grad_scale=1.0 / world)
                return out
            main = torch.cuda.current_stream()
            self.upd.wait_stream(main)
            with torch.cuda.stream(self.upd):              # update of the PREVIOUS training graph's gradient
                allreduce_gradients(store, world)
                adam.step_dev(grad_scale=1.0 / world)
            head = head_fn(ctx, *ins)
            main.wait_stream(self.upd)
            out = grad_fn(ctx, head, *ins)
            store.pack_grads()                             # the next training graph's all-reduce reads the flat buffer
            return out

        def reset():
            with torch.no_grad():
                for t, s0 in zip((store.flat, store.m, store.v, store.step_dev), state):
                    t.copy_(s0)
                if self.deferred:
                    store.flat_grad.zero_()
                    store.step_dev.fill_(-1)

        cap_hi.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(cap_hi):                    # lazy initialisations (NCCL communicator included) outside the captures,
            for _ in range(max(warmup, 1)):                # on the stream the training graphs are captured on
                train(prep_fn(*self.static_in[0]), self.static_in[0])
        torch.cuda.current_stream().wait_stream(cap_hi)
        torch.cuda.synchronize()
        reset()                                            # the warm-up steps trained on the example batch: undo
        self.P, self.T, self.ctx, self.loss = [None, None], [None, None], [None, None], [None, None]
        n0 = _lib.launch_count()
        for s in range(2):
            self.P[s] = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.P[s], stream=cap_lo):
                self.ctx[s] = prep_fn(*self.static_in[s])
        for s in range(2):
            self.T[s] = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.T[s], stream=cap_hi, pool=(self.T[0].pool() if s else None)):
                self.loss[s] = train(self.ctx[s], self.static_in[s])
        self.kernels_per_replay = (_lib.launch_count() - n0) // 2
        reset()                                            # (a capture runs nothing, but keep the contract explicit)
        self.staged = [torch.cuda.Event(), torch.cuda.Event()]
        self.prepped = [torch.cuda.Event(), torch.cuda.Event()]
        self.inputs_read = torch.cuda.Event()
        self.i = 0
        torch.cuda.synchronize()

    def _train(self, s):
        main = torch.cuda.current_stream()
        main.wait_event(self.prepped[s])
        self.T[s].replay()
        return self.loss[s]

    def __call__(self, *inputs):
        s = self.i & 1
        main = torch.cuda.current_stream()
        self.staged[s].record(main)                        # T[s] of the previous call (the last reader of slot s) precedes this
        with torch.cuda.stream(self.side):                 # point, and so does whatever produced `inputs` on the caller's stream
            self.side.wait_event(self.staged[s])
            for d, src in zip(self.static_in[s], inputs):  # staged on the build stream: nothing is added to the training stream
                if d.data_ptr() != src.data_ptr():
                    d.copy_(src, non_blocking=True)
            self.inputs_read.record(self.side)             # `inputs` may be overwritten once this event has completed
            self.P[s].replay()
            self.prepped[s].record(self.side)
        out = self._train(1 - s) if self.i > 0 else None
        self.i += 1
        return out

    def flush(self):
        """Train on the batch whose graph the last call built (the pipeline's drain)."""
        if self.i == 0:
            return None
        out = self._train((self.i - 1) & 1)
        self.i = 0
        if self.deferred:                                  # the update of the batch just trained on
            allreduce_gradients(self.store, self.world)
            self.adam.step_dev(grad_scale=1.0 / self.world)
            self.store.flat_grad.zero_()
        return out

    def close(self):
        self.P = self.T = None
        torch.cuda.synchronize()


def allreduce_gradients(store, world_size):
    """Sum the flat gradient buffer over ranks (NCCL over NVLink on GPUs; gloo in CPU tests)."""
    if world_size > 1:
        torch.distributed.all_reduce(store.flat_grad, op=torch.distributed.ReduceOp.SUM)


def shard_samples(n_samples, rank, world_size):
    """Contiguous block of sample indices owned by `rank` (samples are independent: block-diagonal
    adjacency, graph.py:643-652)."""
    per = n_samples // world_size
    rem = n_samples % world_size
    start = rank * per + min(rank, rem)
    return range(start, start + per + (1 if rank < rem else 0))


class FlatParams:
    """Named groups of parameter tensors as views into ONE flat float32 buffer (+ matching flat gradient / Adam
    moments), for models whose variables are not the (4 W + B per layer) layout of ParamStore - experiment.py's
    Wf / Wg / Wh / Rset / Bset / batch-norm gamma, beta.  `spec`: {group: [array-like initial values]}."""

    def __init__(self, spec, device="cuda"):
        arrays = [(g, i, torch.as_tensor(a, dtype=torch.float32)) for g, lst in spec.items() for i, a in enumerate(lst)]
        total = sum(a.numel() for _, _, a in arrays)
        self.flat = torch.zeros(total, dtype=torch.float32, device=device)
        self.flat_grad = torch.zeros_like(self.flat)
        self.m, self.v = torch.zeros_like(self.flat), torch.zeros_like(self.flat)
        self.step_count = 0
        self.step_dev = torch.zeros(1, dtype=torch.int64, device=device)
        self.groups = {g: [] for g in spec}
        off = 0
        for g, _, a in arrays:
            n = a.numel()
            p = self.flat[off:off + n].view(a.shape)
            p.copy_(a)
            p.requires_grad_(True)
            p.grad = self.flat_grad[off:off + n].view(a.shape)
            self.groups[g].append(p)
            off += n

    def __getitem__(self, group):
        return self.groups[group]

    def zero_grad(self):
        self.flat_grad.zero_()

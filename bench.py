#!/usr/bin/env python
"""bench.py - the reference's headline metric on B200.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port)

metric  : particles/sec per train step (kNN + adjacency + features + fwd + bwd + Adam) at 32^3
workload: BASELINE config 2 - 32^3 particles, k=14, 3-layer graph net [3,32,16,3], batch 8 per GPU
          (weak scaling: every rank owns 8 samples; one NCCL all-reduce of the flat gradient per step).
A "step" is one pass of the hot path over one batch of synthetic particle boxes.

One JSON line on stdout (rank 0).  `value` = whole-job particles/s with inputs resident in HBM;
`e2e` = the same through the public API with pinned HOST inputs copied H2D and the loss read back
D2H inside every timed step.  `roofline` describes the dominant kernel (device time from CUDA events
recorded by the library around each launch); `cpu_baseline` times the oracle port of the reference
path on this box's host cores on a bounded sample.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import tempfile
import time
import types

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "particles/sec per train step (kNN+fwd+bwd) at 32^3"
UNIT = "particles/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n-side", type=int, default=32)
    ap.add_argument("--batch", type=int, default=8, help="samples per GPU")
    ap.add_argument("--k", type=int, default=14)
    ap.add_argument("--kind", default="uniform", choices=["uniform", "clustered"])
    ap.add_argument("--channels", type=int, nargs="+", default=[3, 32, 16, 3])
    ap.add_argument("--cpu-batch", type=int, default=0,
                    help="samples per CPU step (default: the full batch for --impl reference, 2 for the cpu_baseline leg)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from the host instead of replaying one CUDA graph per step")
    ap.add_argument("--no-extras", action="store_true", help="skip the 128^3 kNN build timing")
    ap.add_argument("--no-overlap", action="store_true",
                    help="one graph per step (kNN build, forward, backward in sequence) instead of building the graph of batch i+1 "
                         "on a second stream next to the forward / backward of batch i (train_utils.OverlappedStep)")
    return ap.parse_args()


def common_config(a, world):
    """The keys both arms print identically (the driver compares the arms' configs)."""
    N = a.n_side ** 3
    return {"workload": workload_name(a), "particles_per_step_per_gpu": a.batch * N, "edges_per_step_per_gpu": a.batch * N * a.k,
            "l2": "inputs larger than L2: a step streams ~5 GB of edge tensors (>> 126 MB L2) and rotates 4 distinct input batches"}


def workload_name(a):
    return (f"{a.n_side}^3 particles, k={a.k}, graph net {a.channels}, batch {a.batch} per GPU, "
            f"{a.kind} box, FP32 layers / FP64 kNN distances")


# ----------------------------------------------------------------------------- reference arm (CPU)
def cpu_reference_step(x, za, tgt, k, channels, params):
    """One step of the reference path, restated op for op (oracle/): sklearn KD-tree kNN
    (graph.py:704-713) -> COO (621-662) -> model_func_shift_inv_za (479-515) -> loss_ZA -> backward."""
    from oracle import ref_graph, ref_layers
    b, N, _ = x.shape
    t0 = time.perf_counter()
    A = ref_graph.get_kneighbor_list(x, k)
    t1 = time.perf_counter()
    coo, diag = ref_graph.to_coo_batch_ZA_diag(A)
    t2 = time.perf_counter()
    tp = [([torch.tensor(w, requires_grad=True) for w in Ws], torch.tensor(B, requires_grad=True)) for Ws, B in params]
    mv = types.SimpleNamespace(channels=channels, get_layer_vars=lambda i: tp[i])
    pred = ref_layers.model_func_shift_inv_za(torch.tensor(x), coo, torch.tensor(za), diag, mv, (b, N, k))
    loss = ref_layers.loss_ZA(pred, torch.tensor(tgt))
    t3 = time.perf_counter()
    loss.backward()
    t4 = time.perf_counter()
    return {"knn": t1 - t0, "coo": t2 - t1, "fwd": t3 - t2, "bwd": t4 - t3, "total": t4 - t0, "loss": float(loss.detach())}


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm must use every host core it can."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    torch.set_num_threads(max(n, 1))
    return torch.get_num_threads()


def cpu_sample(a, syn, steps, warmup, b):
    N = a.n_side ** 3
    params = syn.glorot_params(a.channels)
    x = syn.make_box(a.kind, b, N, 0)
    za, tgt = syn.za_features(b, N, 0)
    for _ in range(warmup):
        cpu_reference_step(x, za, tgt, a.k, a.channels, params)
    parts, t0 = [], time.perf_counter()
    for _ in range(steps):
        parts.append(cpu_reference_step(x, za, tgt, a.k, a.channels, params))
    dt = (time.perf_counter() - t0) / steps
    stage = {s: float(np.mean([p[s] for p in parts])) for s in ("knn", "coo", "fwd", "bwd")}
    return b * N / dt, dt, stage


def run_reference(a):
    """--impl reference: the reference's own CPU implementation of the path (oracle port - the reference is
    pure Python and /root/reference does not exist on the GPU box), all host threads torch can use."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    syn = importlib.import_module("n-body_pointcloudevolution_b200.synthetic")
    cores = use_all_host_threads()
    b = a.cpu_batch or a.batch                     # the full batch of the GPU arm: same config
    value, dt, stage = cpu_sample(a, syn, a.steps, min(a.warmup, 1), b)
    sample = (f"{b} samples of the {a.n_side}^3 workload per step "
              f"(kNN sklearn 1 thread as shipped, layers torch-CPU {cores} threads)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": common_config(a, 1),
        "arm": {"reference_sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "host_cpus": os.cpu_count(), "stage_s": stage},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- clocks
CLOCK_QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
               "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
               "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")


class ClockSampler:
    """Samples SM clock, power and the throttle reasons of one GPU DURING the timed region, from a background thread
    through NVML (the same counters `nvidia-smi --query-gpu=clocks.sm,...,clocks_event_reasons.*` prints; an
    nvidia-smi child process polling next to the benchmark stalled the launching thread by milliseconds per query)."""
    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, gpu_index, period_s=0.02):
        import threading
        self.samples, self.period, self._stop, self.thread, self.h = [], period_s, threading.Event(), None, None
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[gpu_index]) if vis and vis.split(",")[gpu_index].strip().isdigit() else gpu_index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._run, daemon=True)
            self.thread.start()
        except Exception:
            self.h = None

    def _sample(self):
        nv = self.nv
        sm = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
        try:
            mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
        except Exception:
            mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
        try:
            power = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
        except Exception:
            power = float("nan")
        return time.time(), sm, mask, power

    def _run(self):
        while not self._stop.is_set():
            try:
                self.samples.append(self._sample())
            except Exception:
                pass
            self._stop.wait(self.period)

    def wait_ready(self, timeout=5.0):
        t0 = time.time()
        while self.h is not None and not self.samples and time.time() - t0 < timeout:
            time.sleep(0.01)

    def mark(self):
        return time.time()

    def stop(self, since=0.0):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.h is None:
            return out
        self._stop.set()
        self.thread.join(timeout=2)
        rows = [r for r in self.samples if r[0] >= since]
        if not rows and self.samples:        # timed region shorter than the sampling period: nearest sample (warm-up load)
            rows = [min(self.samples, key=lambda r: abs(r[0] - since))]
            out["note"] = "timed region shorter than the 20 ms sampling period: nearest sample, taken under the warm-up load"
        if rows:
            reasons = sorted({name for _, _, mask, _ in rows for name, bit in self.REASONS if mask & bit})
            out.update(sm_mhz=float(np.median([r[1] for r in rows])), sm_max_mhz=self.max_sm, reasons=reasons,
                       samples=len(rows), power_w_max=float(np.nanmax([r[3] for r in rows])), source="NVML thread, 20 ms period")
        return out


# ----------------------------------------------------------------------------- roofline bookkeeping
def kernel_algorithmic_bytes(name, b, N, M, relu_by_layer):
    """Algorithmic (minimum) bytes one launch of `name` must move (FP32, s = 4 B).
    c = b*N*M edges, n = b*N nodes; node-level operands are counted once (they are L2-resident)."""
    n, c = b * N, b * N * M
    base, k, q = name, None, None
    if "[" in name:
        base, dims = name.split("[")
        k, q = (int(t.split("=")[1]) for t in dims.rstrip("]").split(","))
    relu = 2 if (k, q) in relu_by_layer and relu_by_layer[(k, q)] else 1   # masked dZ also reads H_out
    if base == "knn_query":
        return n * (12 + 4 * M)                      # xyz in, int32 idx out
    if base in ("gl_pool_kernel", "glf_pool_kernel", "gln_pool_kernel", "gln_pool_generic_kernel", "gln_pool_colonly_kernel"):
        return c * (4 * k + 4) + n * (4 + 2 * 4 * k)           # contract: the pool pass reads H ONCE
    if base in ("gl_edge_out_kernel", "glf_edge_out_kernel", "glk3_edge_out_kernel", "glk3_edge_out_rowpool_kernel", "glt_edge_out_tf32",
                "glt_edge_out_tf32x3"):
        return c * (4 * k + 4 + 4 * q) + n * 2 * 4 * q         # H in, col in, H_out out (+ node tables)
    if base == "gl_last_out_kernel":
        return c * (4 * k + 4) + n * 3 * 4 * q
    if base == "glf_last_out_kernel":
        return c * 4 + n * (4 * k + 3 * 4 * q)
    if base in ("glb_pool_kernel", "glf_bwd_pool_kernel", "gln_bwd_pool_kernel", "gln_bwd_pool_colonly_kernel"):   # contract: dZ read ONCE by the pool pass;
        return c * (4 * q + 4) + n * (4 + 2 * 4 * q)           # the network path delivers dZ pre-masked (no H_out read)
    if base == "xty_partial_dW1":
        return c * (4 * k + 4 * q * relu)
    if base == "glb_edge_in_kernel":
        return c * (4 * q * relu + 4 + 4 * k) + n * 2 * 4 * k
    if base in ("glf_edge_bwd_kernel", "glt_edge_bwd_tf32", "glt_edge_bwd_tf32x3"):   # dZ + H in, dH out (not for layer 1), col
        first = (k == 3)
        return c * (4 * q + 4 * k + (0 if first else 4 * k + 4)) + (0 if first else n * 2 * 4 * k)
    if base == "glk3_edge_dw_kernel":                          # first layer: E (c,3) and dZ in, dW1 out
        return c * (12 + 4 * q)
    if base == "glk3_first_layer_bwd_kernel":                  # first layer, all gradients: E, col and dZ in (once)
        return c * (12 + 4 + 4 * q) + n * 2 * 12
    if base in ("glf_last_edge_in_kernel", "glf_last_edge_in_rowsum_kernel"):   # col in, H (mask) in, dH out
        return c * (4 + 2 * 4 * k) + n * 2 * 4 * k
    if base == "gln_node_project_kernel":                      # P_col, P_row in; Q_col, Q_row out
        return n * (2 * 4 * k + 2 * 4 * q)
    if base == "gln_node_grad_kernel":                         # dQ_col, dQ_row, in-degree in; G_col, G_row out
        return n * (2 * 4 * q + 2 * 4 * k + 4)
    if base in ("glf_node_xty_dW2", "glf_node_xty_dW3"):       # node tensors X (n,k) and Y (n,q), once per layer
        return None
    if base in ("edge_features_kernel", "edge_features_za_kernel"):
        return c * (4 + 12) + n * 12
    if base in ("adj_coo_kernel",):
        return c * (4 + 12)
    if base == "adj_coo_count_kernel":                         # idx in, COO rows + bucket rank out
        return c * (4 + 12 + 4)
    if base in ("seg_count_kernel", "seg_fill_kernel", "adj_fill_kernel"):
        return c * 8
    return None


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(name):
    """dram bytes per launch from the committed ncu capture, if one exists for this kernel."""
    alias = {"knn_query": "knn_query_uniform", "glt_edge_out_tf32": "glt_edge_out_kernel", "glt_edge_out_tf32x3": "glt_edge_out_kernel",
             "glt_edge_bwd_tf32": "glt_edge_bwd_kernel", "glt_edge_bwd_tf32x3": "glt_edge_bwd_kernel"}
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            base = name.split("[")[0]
            e = json.load(f).get(alias.get(base, base))
            return e["bytes_per_launch"] if isinstance(e, dict) else e
    except Exception:
        return None


# ----------------------------------------------------------------------------- our arm
class Workload:
    """One training workload on this rank: a pool of distinct batches resident in HBM (and their pinned host copies), the
    parameter store, and the step function - eager or replayed from one CUDA graph."""

    def __init__(self, nb, dev, rank, world, n_side, b, k, ch, kind, use_graph, n_pool=4, host_copies=True, overlap=True):
        syn, graph, nn_, tu = nb.synthetic, nb.graph, nb.nn, nb.train_utils
        self.nb, self.dev, self.world, self.b, self.N, self.k, self.ch = nb, dev, world, b, n_side ** 3, k, ch
        N = self.N
        self.store = store = tu.ParamStore(ch, device=dev)
        store.load_numpy(syn.glorot_params(ch))
        self.adam = adam = tu.AdamTF(store, lr=0.01)
        mv = types.SimpleNamespace(channels=ch, var_scope="params", get_layer_vars=store.get_layer_vars)
        self.n_pool = n_pool
        self.host = []
        for i in range(n_pool):                                      # different boxes every step; every rank owns its samples
            seed = 1000 * rank + i
            x = syn.make_box(kind, b, N, seed)
            za, tgt = syn.za_features(b, N, seed)
            ts = tuple(torch.from_numpy(t) for t in (x, za, tgt))
            self.host.append(tuple(t.pin_memory() for t in ts) if host_copies else ts)
        self.resident = [tuple(t.to(dev) for t in hb) for hb in self.host]
        if not host_copies:
            self.host = None

        def train_step(x, za, tgt, comm=True, dev_step=False, update=True):
            A = graph.get_kneighbor_list(x, k)                       # kNN rebuilt every step
            coo, diag = graph.to_coo_batch_ZA_diag(A)
            pred = graph.model_func_shift_inv_za(x, coo, za, diag, mv, (b, N, k))
            loss = nn_.loss_ZA(pred, tgt)
            store.zero_grad()
            loss.backward()
            if comm:
                tu.allreduce_gradients(store, world)
            if not update:
                return loss
            if dev_step:
                adam.step_dev(grad_scale=1.0 / world)                # step count in device memory: graph capturable
            else:
                adam.step(grad_scale=1.0 / world)
            return loss
        self.train_step = train_step

        # The whole step (kNN build, adjacency, forward, backward, Adam) is captured once in a CUDA graph and replayed, so
        # the timed region is not limited by the host's launch rate; --no-graph launches eagerly.
        self.graphed, self.graph_note, self.pipelined = None, "eager launches (--no-graph)", False
        self.overlapped = False
        if use_graph:
            try:
                if overlap:
                    # the graph of batch i+1 (kNN, adjacency, CSR transpose: issue-bound, no parameters) is built on a second,
                    # low-priority stream while forward / backward / all-reduce / Adam of batch i (HBM-bound) run on the first:
                    # every call still does one graph build and one parameter update (train_utils.OverlappedStep)
                    def prep(x, za, tgt):
                        return graph.to_coo_batch_ZA_diag(graph.get_kneighbor_list(x, k))

                    seed = torch.ones((), dtype=torch.float32, device=dev)

                    def head(ctx, x, za, tgt):                       # parameter-free start of the step: edge input features
                        coo, diag = ctx
                        return graph.get_input_features_shift_inv_ZA(x, za, coo, diag, (b, N, k))

                    def grad(ctx, edges, x, za, tgt):
                        coo, diag = ctx
                        pred = graph.network_func_shift_inv_za(edges, coo, len(ch) - 1, (b, N), torch.relu, mv)
                        loss = nn_.loss_ZA(pred, tgt)
                        store.zero_grad()
                        loss.backward(seed)                          # (a resident seed gradient: no fill kernel per step)
                        return loss
                    # (the two calls above are the body of graph.model_func_shift_inv_za, graph.py:479-515, split at the point
                    # where the first parameter is read, so that the previous step's all-reduce + Adam run next to `head`)
                    self.graphed = tu.OverlappedStep(prep, grad, store, adam, world, self.resident[0], head_fn=head)
                    self.overlapped = True
                    self.graph_note = ("CUDA graphs on two streams: the kNN graph build of batch i+1 runs next to forward / backward of "
                                       "batch i; the all-reduce + Adam of a step are captured at the start of the next training graph "
                                       "next to its parameter-free edge features; one build and one update per step")
                elif world == 1:
                    self.graphed = tu.GraphedStep(lambda x, za, tgt: train_step(x, za, tgt, dev_step=True), self.resident[0])
                    self.graph_note = "one CUDA graph replay per step (whole step captured once)"
                else:
                    # the NCCL all-reduce is INSIDE the graph, on a forked stream next to the kNN build of the same replay:
                    # replay i all-reduces and applies the gradient of step i-1 while the graph of step i is built
                    # (train_utils.PipelinedStep; same parameter sequence as the plain loop)
                    def prep(x, za, tgt):
                        return graph.to_coo_batch_ZA_diag(graph.get_kneighbor_list(x, k))

                    def grad(ctx, x, za, tgt):
                        coo, diag = ctx
                        loss = nn_.loss_ZA(graph.model_func_shift_inv_za(x, coo, za, diag, mv, (b, N, k)), tgt)
                        store.zero_grad()
                        loss.backward()
                        return loss
                    self.graphed = tu.PipelinedStep(prep, grad, store, adam, world, self.resident[0])
                    self.pipelined = True
                    self.graph_note = ("one CUDA graph replay per step, NCCL all-reduce + Adam of step i-1 captured on a forked "
                                       "stream next to the kNN build of step i")
            except Exception as exc:
                self.graphed, self.graph_note = None, f"eager launches (graph capture failed: {repr(exc)[:200]})"
                torch.cuda.synchronize()

    def run_step(self, x, za, tgt):
        if self.graphed is None:
            return self.train_step(x, za, tgt)
        loss = self.graphed(x, za, tgt)
        if self.overlapped:
            return loss                                              # loss of the previous batch (None on the very first call)
        if self.world > 1 and not self.pipelined:
            self.nb.train_utils.allreduce_gradients(self.store, self.world)
            self.adam.step(grad_scale=1.0 / self.world)
        return loss

    def close(self):
        """Apply a pending pipelined update and drop the CUDA graph (must precede destroy_process_group)."""
        if self.graphed is not None:
            if self.pipelined or self.overlapped:
                self.graphed.flush()
            self.graphed.close()
            self.graphed = None

    def join_streams(self):
        """The build stream of the overlapped loop joins the current stream (a timed region then holds exactly K builds)."""
        if self.overlapped:
            torch.cuda.current_stream().wait_stream(self.graphed.side)

    def step_resident(self, i):
        return self.run_step(*self.resident[i % self.n_pool])


def barrier(world):
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()


def timed(fn, steps, world, dev, stats=None, tag=None, join=None):
    """EXACTLY `steps` calls bracketed by barrier + synchronize on both sides; CUDA-event time, max over ranks (ms).
    `join`: makes the timing stream wait for any other stream the step uses before the last event is recorded."""
    barrier(world)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    host = []
    ev[0].record()
    t0 = time.perf_counter()
    for i in range(steps):
        fn(i)
        if join is not None and i == steps - 1:
            join()
        ev[i + 1].record()
        host.append(time.perf_counter() - t0)
    barrier(world)
    if stats is not None:
        per = np.array([ev[i].elapsed_time(ev[i + 1]) for i in range(steps)])
        hd = np.diff(np.array([0.0] + host)) * 1e3
        stats[tag] = {"device_ms_median": float(np.median(per)), "device_ms_max": float(per.max()),
                      "device_ms_min": float(per.min()), "host_enqueue_ms_median": float(np.median(hd)),
                      "host_enqueue_ms_max": float(hd.max()), "argmax": int(per.argmax())}
    ms = torch.tensor([ev[0].elapsed_time(ev[steps])], device=dev)
    if world > 1:
        torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
    return float(ms.item())


def event_ms(fn, reps, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(reps):
        fn()
    ev[1].record()
    torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) / reps


def knn_sweep(nb, dev, peak, with_sklearn):
    """BASELINE config 3: kNN graph construction, N in {8 x 32^3, 64^3, 128^3}, k in {8, 14, 32}, uniform and clustered
    boxes, open box (graph.get_kneighbor_list, index order) and periodic box (graph.get_pbc_kneighbors_csr, thr 0.05,
    distance order).  Algorithmic bytes = particles * (12 + 4 k).  sklearn (the reference's kneighbors_graph
    call, one thread as shipped) is timed beside it on one sample where that takes seconds, not minutes."""
    syn, graph = nb.synthetic, nb.graph
    rows = []
    for side, b in ((32, 8), (64, 1), (128, 1)):
        N = side ** 3
        for kind in ("uniform", "clustered"):
            xh = syn.make_box(kind, b, N, 0)
            x = torch.from_numpy(xh).to(dev)
            for k in (8, 14, 32):
                reps = 10 if side < 128 else 5
                t_open = event_ms(lambda: graph.get_kneighbor_list(x, k), reps)
                t_pbc = event_ms(lambda: graph.get_pbc_kneighbors_csr(x, k, 0.05, include_self=True), reps)
                ab = b * N * (12 + 4 * k)
                row = {"particles": f"{b}x{side}^3", "kind": kind, "k": k, "open_ms": round(t_open, 4), "periodic_ms": round(t_pbc, 4),
                       "algorithmic_bytes": ab, "open_roofline_frac": ab / (t_open * 1e-3) / 1e9 / peak,
                       "periodic_roofline_frac": ab / (t_pbc * 1e-3) / 1e9 / peak,
                       "open_particles_per_s": b * N / (t_open * 1e-3)}
                # the query kernel nbpc_knn picks for this k, and the other one beside it (same result bit for bit)
                row["kernel"] = "warp-per-query" if k > 16 else "thread-per-query"
                if side < 128:
                    other = "thread" if k > 16 else "warp"
                    nb.set_knn_kernel(other)
                    try:
                        row[f"open_ms_{other}_kernel"] = round(event_ms(lambda: graph.get_kneighbor_list(x, k), 5), 4)
                    finally:
                        nb.set_knn_kernel("auto")
                if with_sklearn and (side == 32 or (side == 64 and k == 14)):
                    from oracle import ref_graph                      # CPU leg (checker code timed as the baseline)
                    t0 = time.perf_counter()
                    ref_graph.get_kneighbor_list(xh[:1], k)
                    row["sklearn_open_ms_per_sample"] = round((time.perf_counter() - t0) * 1e3, 1)
                    row["sklearn_note"] = "kneighbors_graph (KD-tree, 1 thread as shipped, graph.py:709) on ONE sample"
                rows.append(row)
            del x
    return rows


def set_model_bench(nb, dev, peak, steps=10):
    """The model the reference's train.py actually runs (train.py:66): nn.model_func_set at the default widths
    (utils.py:165), batch 8 x 32^3, one training step = forward + loss + backward + Adam, replayed from one CUDA graph.
    Algorithmic bytes: forward s b N (2 k + q) per layer (mean pass + GEMM pass), backward s b N (2 q + 2 k)
    (first layer: no dH: 2 q + k)."""
    syn, nn_, tu, lib = nb.synthetic, nb.nn, nb.train_utils, nb._lib
    ch = [6, 64, 128, 128, 256, 64, 128, 16, 3]
    b, N = 8, 32 ** 3
    rng = np.random.default_rng(0)
    X = torch.from_numpy(rng.standard_normal((b, N, 6)).astype(np.float32)).to(dev)
    Y = torch.from_numpy((0.1 * rng.standard_normal((b, N, 3))).astype(np.float32)).to(dev)
    store = tu.ParamStore(ch, device=dev)
    adam = tu.AdamTF(store, lr=0.001)
    mv = store.model_vars(torch.relu)

    def step(x, y):
        loss = nn_.loss_ZA(nn_.model_func_set(x, mv), y)
        store.zero_grad()
        loss.backward()
        adam.step_dev()
        return loss
    gs = tu.GraphedStep(step, (X, Y))
    ms = event_ms(lambda: gs(X, Y), steps, warm=3)
    rows = b * N
    fwd = sum(4 * rows * (2 * k + q) for k, q in zip(ch[:-1], ch[1:]))
    bwd = sum(4 * rows * (2 * q + (2 * k if i else k)) for i, (k, q) in enumerate(zip(ch[:-1], ch[1:])))
    flops = 3 * 2 * rows * sum(k * q for k, q in zip(ch[:-1], ch[1:]))          # fwd + dH + dW GEMMs
    lib.prof_enable(True)
    for _ in range(2):
        step(X, Y)
    torch.cuda.synchronize()
    rep = lib.prof_report()
    lib.prof_enable(False)
    kern = sorted(((n, tot / 2) for n, (cnt, tot) in rep.items()), key=lambda kv: -kv[1])[:int(os.environ.get("NBPC_BENCH_TOPK", "14"))]
    return {"channels": ch, "batch": b, "particles": rows, "ms_per_step": ms, "particles_per_s": rows / (ms * 1e-3), "finite": bool(torch.isfinite(gs.loss)),
            "algorithmic_bytes": fwd + bwd, "roofline_frac": (fwd + bwd) / (ms * 1e-3) / 1e9 / peak,
            "gemm_tflops_fp32_equiv": flops / (ms * 1e-3) / 1e12, "math_mode": lib.get_math_mode(),
            "kernels_ms": {n: round(t, 4) for n, t in kern}}


def layer15_bench(nb, dev, peak, steps=5):
    """the 15-weight layer (graph.py:20-229) as a training step - symmetrised adjacency build + 3-layer net
    [3,32,16,3] forward + loss + backward on 8 x 32^3 particles, k = 14 (csrc/graph15.cu: node-level pooling / projections +
    one edge kernel per direction)."""
    syn, graph, nn_ = nb.synthetic, nb.graph, nb.nn
    ch, b, N, k = [3, 32, 16, 3], 8, 32 ** 3, 14
    x = torch.from_numpy(syn.make_box("uniform", b, N, 0)).to(dev)
    tgt = torch.from_numpy(syn.za_features(b, N, 0)[1]).to(dev)
    rng = np.random.default_rng(3)
    tp = [(torch.tensor((rng.standard_normal((15, kk, qq)) * np.sqrt(2.0 / (kk + qq))).astype(np.float32), device=dev, requires_grad=True),
           torch.zeros((2, qq), device=dev, requires_grad=True)) for kk, qq in zip(ch[:-1], ch[1:])]
    mgr = types.SimpleNamespace(channels=ch, var_scope="params", get_layer_vars=lambda i: tp[i])
    t_adj = event_ms(lambda: graph.get_symmetrized_adjacency(graph.get_kneighbor_list(x, k)), 3, warm=1)
    adj = graph.get_symmetrized_adjacency(graph.get_kneighbor_list(x, k))
    S = int(adj["row"].numel())
    edges = (x.reshape(b * N, 3)[adj["col"].long()] - x.reshape(b * N, 3)[adj["row"].long()]).contiguous()

    def step():
        for W, B in tp:
            W.grad = None
            B.grad = None
        loss = nn_.loss_ZA(graph.model_func_15op_shift_inv_za(edges, adj, mgr, (b, N, k)), tgt)
        loss.backward()
        return loss
    ms = event_ms(step, steps, warm=2)
    return {"channels": ch, "particles": b * N, "edges_symmetrised": S, "adjacency_build_ms": t_adj, "fwd_bwd_ms": ms,
            "particles_per_s": b * N / (ms * 1e-3), "finite": bool(torch.isfinite(step()))}


def main():
    a = parse()
    if a.impl == "reference":
        return run_reference(a)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200 GPU (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.distributed.init_process_group("nccl", device_id=dev)

    nb = importlib.import_module("n-body_pointcloudevolution_b200")
    syn, graph, tu, lib = nb.synthetic, nb.graph, nb.train_utils, nb._lib
    nb.ops.device_check()
    try:   # the graphs are captured on their own streams and the per-kernel profile below runs eagerly on the default stream
        torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)
    except AttributeError:
        pass

    N, b, k, ch = a.n_side ** 3, a.batch, a.k, a.channels
    wl = Workload(nb, dev, rank, world, a.n_side, b, k, ch, a.kind, not a.no_graph, overlap=not a.no_overlap)
    store, host, n_pool, graph_note = wl.store, wl.host, wl.n_pool, wl.graph_note
    step_stats = {}

    # End to end: every step's inputs start in pinned HOST memory and every step's loss ends in pinned host memory.
    # Like a production input pipeline the copies are asynchronous: step i+1's H2D runs on a copy stream while step i
    # computes (two device staging slots, event-ordered), and the loss is read back with a non-blocking D2H copy; all
    # copies complete inside the timed region (it ends with a device-wide synchronize).
    copy_stream = torch.cuda.Stream(device=dev)
    d2h_stream = torch.cuda.Stream(device=dev)
    loss_done = torch.cuda.Event()
    slots = [tuple(torch.empty_like(t, device=dev) for t in host[0]) for _ in range(2)]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]
    loss_host = torch.zeros(max(a.steps, 8), dtype=torch.float32).pin_memory()
    used = [False, False]

    def prefetch(i):
        sl = i % 2
        with torch.cuda.stream(copy_stream):
            if used[sl]:
                copy_stream.wait_event(consumed[sl])                 # the step that read this slot has been enqueued
            for d, h in zip(slots[sl], host[i % n_pool]):
                d.copy_(h, non_blocking=True)                        # H2D from pinned memory
            ready[sl].record(copy_stream)

    def step_e2e(i):
        sl = i % 2
        if i == 0:
            prefetch(0)
        torch.cuda.current_stream().wait_event(ready[sl])
        prefetch(i + 1)
        loss = wl.run_step(*slots[sl])
        if wl.overlapped:                                            # the build stream stages the inputs: its event says when
            consumed[sl] = wl.graphed.inputs_read                    # the slot may be refilled (one event, re-recorded per call:
            copy_stream.wait_event(consumed[sl])                     # make the copy stream wait for THIS recording now)
            used[sl] = False
        else:
            consumed[sl].record()
            used[sl] = True
        if loss is not None:                                         # (overlapped loop: the loss of the previous batch)
            # D2H read of the loss on its own stream, behind the training graph that wrote it: the 4-byte DMA does not sit
            # between two training graphs on the compute stream
            loss_done.record()
            with torch.cuda.stream(d2h_stream):
                d2h_stream.wait_event(loss_done)
                loss_host[i % loss_host.numel()].copy_(loss.detach().reshape(()), non_blocking=True)

    # ---- warm-up, then the timed region (device-resident inputs)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    for i in range(max(a.warmup, 3)):
        wl.step_resident(i)
    skip = 0
    if sampler:
        sampler.wait_ready()
        skip = sampler.mark()
    l0 = lib.launch_count()
    ms = timed(wl.step_resident, a.steps, world, dev, step_stats, "resident", join=wl.join_streams)
    launches = lib.launch_count() - l0
    if wl.graphed is not None:                                       # kernels are launched by the graph replays
        launches = (wl.graphed.kernels_per_replay + (1 if world > 1 else 0)) * a.steps   # + the NCCL kernel
    clocks = sampler.stop(skip) if sampler else {}
    particles = world * b * N
    value = particles * a.steps / (ms * 1e-3)

    # ---- end to end through the public API with host buffers
    for i in range(2):
        step_e2e(i)
    torch.cuda.synchronize()
    used[0] = used[1] = False
    def join_e2e():                                                   # the last event covers the build, copy and read-back streams
        wl.join_streams()
        torch.cuda.current_stream().wait_stream(d2h_stream)
        torch.cuda.current_stream().wait_stream(copy_stream)
    ms_e2e = timed(step_e2e, a.steps, world, dev, step_stats, "e2e", join=join_e2e)
    assert bool(torch.isfinite(loss_host[:min(a.steps, loss_host.numel())]).all()), "e2e losses did not arrive on the host"
    e2e_value = particles * a.steps / (ms_e2e * 1e-3)
    h2d = sum(t.numel() * t.element_size() for t in host[0])

    # ---- BASELINE config 4: 64^3 particles, data-parallel training, run by ALL ranks (weak: 1 sample per GPU; strong:
    # global batch 8 split over the ranks).  Same step, same timing rules; skipped with --no-extras.
    c4 = {}
    if not a.no_extras and a.n_side == 32 and ch == [3, 32, 16, 3]:
        for mode, bb in (("weak", 1), ("strong", max(8 // world, 1))):
            try:
                w4 = Workload(nb, dev, rank, world, 64, bb, k, ch, a.kind, not a.no_graph, n_pool=2, host_copies=False, overlap=not a.no_overlap)
                for i in range(3):
                    w4.step_resident(i)
                st = 20
                ms4 = timed(w4.step_resident, st, world, dev, join=w4.join_streams)
                p4 = world * bb * 64 ** 3
                c_edges = bb * 64 ** 3 * k
                step_bytes = 1428 * c_edges + (12 + 4 * k) * bb * 64 ** 3 + 20 * c_edges + 16 * bb * 64 ** 3
                c4[mode] = {"samples_per_gpu": bb, "global_batch": world * bb, "ms_per_step": ms4 / st, "particles_per_s": p4 * st / (ms4 * 1e-3),
                            "step_roofline_frac": step_bytes / (ms4 / st * 1e-3) / 1e9 / measured_peak_gbs()[0], "launch": w4.graph_note}
                w4.close()
                del w4
                torch.cuda.empty_cache()
            except Exception as exc:
                c4[mode] = {"error": repr(exc)[:300]}
                barrier(world)

    graphed_kernels = wl.graphed.kernels_per_replay if wl.graphed is not None else None
    overlapped_note = wl.overlapped
    if world > 1:
        wl.close()                                                    # a live graph holding NCCL kernels hangs the teardown
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    if rank != 0:
        return

    # ---- per-kernel device times (library-side CUDA events on the launching stream), rank 0
    prof_steps = 3
    torch.cuda.synchronize()
    lib.prof_enable(True)
    for i in range(prof_steps):
        wl.train_step(*wl.resident[i % n_pool], comm=False)
    torch.cuda.synchronize()
    report = lib.prof_report()
    lib.prof_enable(False)
    kern_ms = {n: tot / prof_steps for n, (cnt, tot) in report.items()}
    kern_cnt = {n: cnt / prof_steps for n, (cnt, tot) in report.items()}
    sum_ms = sum(kern_ms.values())
    step_ms = ms / a.steps
    relu_by_layer = {(kk, qq): (li < len(ch) - 2) for li, (kk, qq) in enumerate(zip(ch[:-1], ch[1:]))}
    peak, peak_src = measured_peak_gbs()
    top = sorted(kern_ms.items(), key=lambda kv: -kv[1])
    kernels = []
    for name, t in top[:40]:
        per_launch_ms = t / kern_cnt[name]
        ab = kernel_algorithmic_bytes(name, b, N, k, relu_by_layer)
        # share: of the graph-replayed step (ms_per_step); the eager event times of all kernels sum to `eager_sum_ms`
        kernels.append({"kernel": name, "ms_per_step": round(t, 4), "launches_per_step": kern_cnt[name],
                        "share": round(t / step_ms, 4),
                        "algorithmic_bytes": ab,
                        "achieved_gbs": (ab / (per_launch_ms * 1e-3) / 1e9) if ab else None})
    dom = kernels[0]
    roofline = {"bound": "hbm", "kernel": dom["kernel"], "achieved": dom["achieved_gbs"], "peak": peak,
                "peak_source": peak_src, "unit": "GB/s",
                "frac": (dom["achieved_gbs"] / peak) if dom["achieved_gbs"] else None,
                "traffic": ncu_traffic(dom["kernel"]), "share_of_step": dom["share"],
                "share_note": "kernel time (eager, CUDA events) / ms_per_step of the graph-replayed step",
                "eager_sum_ms": round(sum_ms, 4),
                "algorithmic_bytes_per_launch": dom["algorithmic_bytes"],
                "step_algorithmic_bytes": None, "kernels": kernels}
    # whole-step roofline (SURVEY §8d contract figure: 1428 B/edge + 68 B + 280 B per particle for [3,32,16,3], k=14)
    if ch == [3, 32, 16, 3]:
        c_edges = b * N * k
        step_bytes = 1428 * c_edges + (12 + 4 * k) * b * N + 20 * c_edges + 16 * b * N
        roofline["step_algorithmic_bytes"] = step_bytes
        roofline["step_achieved"] = step_bytes / (step_ms * 1e-3) / 1e9
        roofline["step_frac"] = roofline["step_achieved"] / peak

    extras = {}
    knn128 = {}
    if not a.no_extras:
        del wl
        torch.cuda.empty_cache()
        # ---- BASELINE config 3: kNN build sweep (includes the metric's second half: kNN build ms at 128^3)
        try:
            sweep = knn_sweep(nb, dev, peak, with_sklearn=not a.no_cpu_baseline)
            extras["knn_sweep"] = sweep
            for r in sweep:
                if r["particles"] == "1x128^3" and r["k"] == 14:
                    knn128[r["kind"]] = {"periodic_ms": r["periodic_ms"], "open_ms": r["open_ms"],
                                         "periodic_roofline_frac": r["periodic_roofline_frac"]}
        except Exception as exc:
            extras["knn_sweep"] = {"error": repr(exc)[:300]}
        if c4:
            extras["config4_64^3_dp"] = c4
        try:
            extras["set_model"] = set_model_bench(nb, dev, peak)
        except Exception as exc:
            extras["set_model"] = {"error": repr(exc)[:300]}
        try:
            extras["layer15"] = layer15_bench(nb, dev, peak)
        except Exception as exc:
            extras["layer15"] = {"error": repr(exc)[:300]}
        torch.cuda.empty_cache()

        # BASELINE config 5: 128^3-particle multi-redshift rollout INFERENCE, periodic kNN graph rebuilt every step
        # (graph.rollout_shift_inv: pbc kNN -> 9-channel edges -> [9,32,16,6] graph net -> scaled residual -> readout)
        try:
            n5 = 128 ** 3
            rng = np.random.default_rng(5)
            X5 = np.concatenate([syn.make_box("uniform", 1, n5, 0), 0.01 * rng.standard_normal((1, n5, 3)).astype(np.float32)], axis=-1)
            X5 = torch.from_numpy(np.ascontiguousarray(X5, dtype=np.float32)).to(dev)
            ch5 = [9, 32, 16, 6]
            st5 = tu.ParamStore(ch5, device=dev)
            st5.load_numpy(syn.glorot_params(ch5))
            mv5 = types.SimpleNamespace(channels=ch5, var_scope="params", get_layer_vars=st5.get_layer_vars,
                                        get_scalars=lambda: (0.01, 0.01))
            with torch.no_grad():
                state = {"X": X5}

                def roll():
                    state["X"] = graph.rollout_shift_inv(state["X"], mv5, 14, 0.05)
                t5 = event_ms(roll, 5)
            extras["rollout_128^3"] = {"ms_per_rollout_step": t5, "particles_per_s": n5 / (t5 * 1e-3), "channels": ch5, "k": 14,
                                       "boundary_threshold": 0.05, "finite": bool(torch.isfinite(state["X"]).all()),
                                       "what": "inference, periodic kNN rebuilt every step, FP64 kNN distances, tf32x3/fp32 layers"}
            del X5, state
        except Exception as exc:   # the headline line must survive a failure of this extra
            extras["rollout_128^3"] = {"error": repr(exc)[:300]}

    # ---- the reference's CPU path on this box's host cores (bounded sample)
    cpu = None
    if not a.no_cpu_baseline and world == 1:
        cores = use_all_host_threads()
        cb = a.cpu_batch or 2
        v, dt, stage = cpu_sample(a, syn, 1, 1, cb)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "host_cpus": os.cpu_count(),
               "sample": f"1 step of {cb} samples of the same {a.n_side}^3/k={a.k} workload "
                         f"(kNN: sklearn KD-tree, 1 thread as shipped; layers: torch-CPU, {cores} threads)",
               "stage_s": stage}

    cfg = common_config(a, world)                                     # identical in both arms
    arm = {"math_mode": lib.get_math_mode(), "launch": graph_note, "particles_per_step": particles, "edges_per_step": particles * k,
           "parallelism": f"dp{world} (sample-sharded, 1 NCCL all-reduce of {store.flat.numel()} floats/step)"}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
        "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"fp32": "f32", "tf32x3": "f32 (tcgen05 TF32x3 error-compensated, FP32-class accuracy)",
                  "tf32": "tf32 (tcgen05 single pass, FP32 accumulate)"}[lib.get_math_mode()],
        "data": "synthetic",
        "config": cfg,
        "arm": arm,
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / a.steps, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": 4,
                "pipeline": "inputs in pinned host memory, H2D one step ahead on a copy stream (2 device slots), loss read "
                            "back every step with a non-blocking D2H copy into pinned memory on its own stream; the last timed "
                            "event waits for the build, copy and read-back streams and the region ends with a device-wide "
                            "synchronize" + ("; overlapped loop: the loss read in step i is that of batch i-1 (each batch's "
                                             "loss is read exactly once)" if overlapped_note else "")},
        "gpu_launches": int(launches),
        # second half of BASELINE's metric: periodic kNN build (k = 14, thr 0.05, self included) on one 128^3 box, ms
        "knn_build_ms_128^3": knn128 or None,
        "step_stats": step_stats,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "extras": extras,
    }
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()

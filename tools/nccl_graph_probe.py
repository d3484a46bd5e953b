"""Probe: does an NCCL all-reduce captured inside torch.cuda.graph replay correctly on this box?  (bounded: run under `timeout`)
   torchrun --nproc-per-node 2 tools/nccl_graph_probe.py"""
import os, sys, time
import torch
import torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
g = torch.full((2675,), float(rank + 1), device=dev)
out = torch.zeros_like(g)
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for _ in range(3):
        t = g.clone()
        dist.all_reduce(t)
torch.cuda.current_stream().wait_stream(side)
torch.cuda.synchronize()
print(rank, "warm-up ok", float(t[0]), flush=True)
graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph):
    tmp = g * 2.0
    dist.all_reduce(tmp)
    out.copy_(tmp)
torch.cuda.synchronize()
print(rank, "captured", flush=True)
for i in range(20):
    g.fill_(float(rank + 1 + i))
    graph.replay()
torch.cuda.synchronize()
expect = 2.0 * sum(r + 1 + 19 for r in range(world))
print(rank, "replay ok", float(out[0]), expect, flush=True)
assert float(out[0]) == expect
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
ev[0].record()
for i in range(200):
    graph.replay()
ev[1].record()
torch.cuda.synchronize()
print(rank, "us per replay (mul + allreduce + copy)", ev[0].elapsed_time(ev[1]) * 1000 / 200, flush=True)
# teardown: a live CUDA graph that holds NCCL kernels made destroy_process_group() hang; drop the graph first
del graph
torch.cuda.synchronize()
dist.barrier()
print(rank, "barrier ok", flush=True)
dist.destroy_process_group()
print(rank, "destroyed", flush=True)

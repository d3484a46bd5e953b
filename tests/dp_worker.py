"""Worker of tests/test_gpu_round2.py::test_dp2_gradients_match_global_batch (launched with torchrun, one rank per GPU):
data-parallel training of the graph model over NCCL must equal single-GPU training on the global batch."""
import importlib
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main(out_path):
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    torch.distributed.init_process_group("nccl", device_id=dev)
    nb = importlib.import_module("n-body_pointcloudevolution_b200")
    syn, graph, nn_, tu = nb.synthetic, nb.graph, nb.nn, nb.train_utils
    ch, N, k, per = [3, 32, 16, 3], 1000, 8, 2
    B = per * world                                                  # global batch
    steps = 3
    batches = []
    for i in range(steps):
        x = syn.make_box("clustered", B, N, 40 + i)
        za, tgt = syn.za_features(B, N, 40 + i)
        batches.append((x, za, tgt))

    def run(sample_slice, n_ranks, comm):
        store = tu.ParamStore(ch, device=dev)
        store.load_numpy(syn.glorot_params(ch))
        adam = tu.AdamTF(store, lr=0.01)
        mv = types.SimpleNamespace(channels=ch, var_scope="params", get_layer_vars=store.get_layer_vars)
        grads = []
        for x, za, tgt in batches:
            xs, zs, ts = (torch.tensor(np.ascontiguousarray(t[sample_slice]), device=dev) for t in (x, za, tgt))
            b = xs.shape[0]
            coo, diag = graph.to_coo_batch_ZA_diag(graph.get_kneighbor_list(xs, k))
            loss = nn_.loss_ZA(graph.model_func_shift_inv_za(xs, coo, zs, diag, mv, (b, N, k)), ts)
            store.zero_grad()
            loss.backward()
            if comm:
                tu.allreduce_gradients(store, n_ranks)
            grads.append(store.flat_grad.clone() / n_ranks)           # gradient of the GLOBAL-batch mean loss
            adam.step(grad_scale=1.0 / n_ranks)
        return torch.stack(grads), store.flat.clone()

    g_dp, p_dp = run(slice(rank * per, (rank + 1) * per), world, True)        # this rank's samples, NCCL all-reduce

    # the same data-parallel run with the whole step (NCCL all-reduce included) captured in one CUDA graph
    def run_pipelined(overlapped=False):
        store = tu.ParamStore(ch, device=dev)
        store.load_numpy(syn.glorot_params(ch))
        adam = tu.AdamTF(store, lr=0.01)
        mv = types.SimpleNamespace(channels=ch, var_scope="params", get_layer_vars=store.get_layer_vars)
        sl = slice(rank * per, (rank + 1) * per)
        dev_batches = [tuple(torch.tensor(np.ascontiguousarray(t[sl]), device=dev) for t in bt) for bt in batches]

        def prep(x, za, tgt):
            return graph.to_coo_batch_ZA_diag(graph.get_kneighbor_list(x, k))

        def grad(ctx, x, za, tgt):
            loss = nn_.loss_ZA(graph.model_func_shift_inv_za(x, ctx[0], za, ctx[1], mv, (per, N, k)), tgt)
            store.zero_grad()
            loss.backward()
            return loss
        if overlapped:   # deferred-update form: all-reduce + Adam of a step at the start of the next training graph
            def head(ctx, x, za, tgt):
                return graph.get_input_features_shift_inv_ZA(x, za, ctx[0], ctx[1], (per, N, k))

            def grad_h(ctx, edges, x, za, tgt):
                loss = nn_.loss_ZA(graph.network_func_shift_inv_za(edges, ctx[0], len(ch) - 1, (per, N), torch.relu, mv), tgt)
                store.zero_grad()
                loss.backward()
                return loss
            ps = tu.OverlappedStep(prep, grad_h, store, adam, world, dev_batches[0], head_fn=head)
        else:
            ps = tu.PipelinedStep(prep, grad, store, adam, world, dev_batches[0])
        for bt in dev_batches:
            ps(*bt)
        ps.flush()
        ps.close()
        return store.flat.clone()
    p_pipe = run_pipelined()
    p_over = run_pipelined(overlapped=True)       # graph build of batch i+1 next to forward / backward / all-reduce / Adam of batch i
    if rank == 0:
        g_1, p_1 = run(slice(0, B), 1, False)                                  # same global batch on one GPU
        np.savez(out_path, g_dp=g_dp.cpu().numpy(), p_dp=p_dp.cpu().numpy(), g_1=g_1.cpu().numpy(), p_1=p_1.cpu().numpy(),
                 p_pipe=p_pipe.cpu().numpy(), p_over=p_over.cpu().numpy())
    # every rank must hold identical parameters
    gathered = [torch.empty_like(p_dp) for _ in range(world)]
    torch.distributed.all_gather(gathered, p_dp)
    assert all(torch.equal(gathered[0], t) for t in gathered), "replicas diverged"
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main(sys.argv[1])

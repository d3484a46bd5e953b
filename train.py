"""Drop-in for the reference's train.py (train.py:1-184) on the B200-native hot path.

Same command line as the reference (utils.py:242-271: -c -i -b -d -k -n -s -l -t), same loop: initialise parameters from
the seed, Adam (TF form), `num_iters` iterations on random minibatches, a checkpoint + printout every 250 steps, then the
test-set evaluation.  Differences a maintainer should know about:
  * `-k/--kneighbors` is LIVE (the reference parses it but train.py:48 comments it out): K > 0 trains the shift-invariant
    graph model on a kNN graph rebuilt every step (graph.get_kneighbor_list -> to_coo_batch_ZA_diag ->
    model_func_shift_inv_za), K == -1 trains the set model (nn.model_func_set), as the flag's help text says.
    With the graph model the first channel must be 3 (relative positions), e.g. -c 3 32 16 3 -k 14.
  * the ZA/FastPM cubes (utils.py:92-135, `ZA_###.npy`) are not redistributable: if `--data_dir` holds `ZA_{idx:03d}.npy`
    it is loaded exactly like utils.Dataset.load_data (utils.py:595-620), otherwise a synthetic set with the same array
    layout (N, 9) = [q - 64, ZA displacement, FPM - ZA] is generated (`--side` particles per axis).
  * launched under torchrun it trains data-parallel: every rank draws its own minibatch of `batch_size` samples and the
    flat gradient buffer is all-reduced once per step (NCCL).
  * checkpoints are `torch.save` files under `--out_dir` (the reference's tf.train.Saver cannot restore either,
    utils.py:481-482).
There is no CPU fallback: the layers and the kNN run in libnbpc.so on an sm_100 GPU."""
import argparse
import os
import time
import types

import numpy as np
import torch

import nbpc

CHANNELS = [6, 64, 128, 128, 256, 64, 128, 16, 3]   # utils.py:165
PARAMS_SEED, DATASET_SEED, NUM_NEIGHBORS = 77743196, 12345, 14   # utils.py:161, 147, 166


def parser():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawTextHelpFormatter)
    adg = ap.add_argument
    adg("-c", "--channels", type=int, nargs="+", default=None, metavar="C", help="List of ints that define layer sizes")
    adg("-i", "--num_iters", type=int, default=20000, metavar="N", help="Number of training iterations")
    adg("-b", "--batch_size", type=int, default=4, metavar="B", help="Number of samples per training batch")
    adg("-d", "--data_idx", type=int, default=0, choices=set(range(10)), metavar="i", help="Index, int in [0, 10), of a dataset")
    adg("-k", "--kneighbors", type=int, default=NUM_NEIGHBORS, metavar="K",
        help="Number of neighbors in graph model (KNN); if K == -1, then set model")
    adg("-n", "--name", type=str, default="", metavar="name", help="Name for model; randomly generated if not specified")
    adg("-s", "--seed", type=int, default=PARAMS_SEED, metavar="X", help="Random seed for parameter initialization")
    adg("-l", "--learnrate", type=float, default=0.01, metavar="lr", help="Learning rate for optimizer")
    adg("-t", "--num_test", type=int, default=200, metavar="M", help="Number of samples in test set")
    # additions (not in the reference)
    adg("--data_dir", type=str, default=os.path.expanduser("~/.Data/nbody_simulations"), help="directory with ZA_###.npy")
    adg("--out_dir", type=str, default=os.path.expanduser("~/.Data/Experiments/Nbody"), help="checkpoint / result directory")
    adg("--side", type=int, default=32, help="particles per axis of the synthetic data set")
    adg("--num_samples", type=int, default=0, help="synthetic samples (default: num_test + 100 + 4 * batch_size)")
    adg("--checkpoint", type=int, default=250, help="steps between checkpoints (train.py:29)")
    adg("--graph", action="store_true", help="capture the training step once in a CUDA graph and replay it every iteration")
    return ap


class Dataset:
    """utils.py:547-621: X (samples, N, 9) = [q - 64, ZA displacement, FPM - ZA]; seeded split train / 100 val / num_test."""

    def __init__(self, args):
        path = os.path.join(args.data_dir, f"ZA_{args.data_idx + 1:03d}.npy")
        if os.path.isfile(path):
            X = self.load_data(path)
        else:
            n = args.num_samples or (args.num_test + 100 + 4 * args.batch_size)
            X = self.synthetic(n, args.side)
            print(f"\nNo {path}: synthetic data, {n} samples of {args.side}^3 particles\n")
        rng = np.random.RandomState(DATASET_SEED)
        X = X[rng.permutation(X.shape[0])]
        self.X_train, self.X_val, self.X_test = np.split(X, [-args.num_test - 100, -args.num_test], axis=0)
        self.rng = np.random.RandomState(DATASET_SEED + 1 + int(os.environ.get("RANK", "0")))

    @staticmethod
    def grid(side):
        mg = np.arange(side, dtype=np.float32) * 4 + 2          # utils.py:611: range(2, 130, 4) for side = 32
        q = np.einsum("ijkl->kjli", np.array(np.meshgrid(mg, mg, mg)))
        return q.reshape(1, -1, 3)

    @classmethod
    def load_data(cls, path):
        data = np.load(path)                                    # (1000, 32, 32, 32, 19)
        n, side = data.shape[0], data.shape[1]                  # the reference hard-codes side = 32 (utils.py:611)
        za = data[..., 1:4].reshape(n, -1, 3)
        fpm = data[..., 7:10].reshape(n, -1, 3) - za
        q = np.broadcast_to(cls.grid(side), za.shape)
        return np.concatenate([q - 2.0 * side, za, fpm], axis=-1).astype(np.float32)   # q - 64 for side = 32

    @classmethod
    def synthetic(cls, n, side):
        rng = np.random.default_rng(DATASET_SEED)
        N = side ** 3
        za = rng.normal(0.0, 1.0, (n, N, 3)).astype(np.float32)
        fpm = (0.1 * za + rng.normal(0.0, 0.05, (n, N, 3))).astype(np.float32)   # a learnable ZA -> FPM correction
        q = np.broadcast_to(cls.grid(side) - 2.0 * side, za.shape)
        return np.concatenate([q, za, fpm], axis=-1).astype(np.float32)

    def get_minibatch(self, batch_size):
        idx = self.rng.choice(self.X_train.shape[0], batch_size, replace=False)
        return np.copy(self.X_train[idx])


def build_step(args, store, dev):
    """-> fn(batch (b, N, 9) numpy) -> (prediction (b, N, 3), loss); the graph is rebuilt inside for the graph model."""
    graph, nn = nbpc.graph, nbpc.nn
    mv = types.SimpleNamespace(channels=store.channels, var_scope="params", num_layers=store.num_layers,
                               get_layer_vars=store.get_layer_vars, activation=torch.relu)
    K = args.kneighbors

    def forward(batch):
        x = batch if isinstance(batch, torch.Tensor) else torch.from_numpy(batch).to(dev, non_blocking=True)
        true_error = x[..., 6:].contiguous()
        if K == -1:
            pred = nn.model_func_set(x[..., :6].contiguous(), mv)                       # train.py:66
        else:
            b, N = x.shape[0], x.shape[1]
            za = x[..., 3:6].contiguous()
            pos = (x[..., :3] + za).contiguous()                                        # nn.get_init_pos: q + ZA displacement
            coo, diag = graph.to_coo_batch_ZA_diag(graph.get_kneighbor_list(pos, K))
            # ONE position set for the graph and for the edge features (graph.py:289-343: "coo ... MADE FROM init_pos")
            pred = graph.model_func_shift_inv_za(pos, coo, za, diag, mv, (b, N, K))
        return pred, nn.loss_ZA(pred, true_error)                                       # train.py:71
    return forward


def main(argv=None):
    args = parser().parse_args(argv)
    if args.channels is None:
        args.channels = CHANNELS if args.kneighbors == -1 else [3, 32, 16, 3]
    if args.kneighbors != -1 and args.channels[0] != 3:
        raise SystemExit("graph model (-k > 0): the first channel must be 3 (relative positions); use -k -1 for the set model")
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.distributed.init_process_group("nccl", device_id=dev)
    nbpc.ops.device_check()
    name = args.name or f"model_{int(time.time()) % 100000}"
    out_dir = os.path.join(args.out_dir, name)
    if rank == 0:
        os.makedirs(os.path.join(out_dir, "Session"), exist_ok=True)
        os.makedirs(os.path.join(out_dir, "Results"), exist_ok=True)
    dataset = Dataset(args)
    store = nbpc.train_utils.ParamStore(args.channels, seed=args.seed, device=dev)      # utils.initialize_params
    adam = nbpc.train_utils.AdamTF(store, lr=args.learnrate)                            # tf.train.AdamOptimizer(lr)
    forward = build_step(args, store, dev)

    def save(step):
        if rank == 0:
            torch.save({"step": step, "params": store.flat.cpu(), "m": store.m.cpu(), "v": store.v.cpu(), "channels": args.channels},
                       os.path.join(out_dir, "Session", f"chkpt-{step}.pt"))

    def grad_step(x):
        pred, loss = forward(x)
        store.zero_grad()
        loss.backward()
        return loss

    def update(dev_step=False):
        nbpc.train_utils.allreduce_gradients(store, world)
        adam.step_dev(grad_scale=1.0 / world) if dev_step else adam.step(grad_scale=1.0 / world)

    def train_step(x, dev_step=False):
        loss = grad_step(x)
        update(dev_step)
        return loss

    graphed = None
    if args.graph and args.num_iters > 0:
        # The warm-up steps of the capture would draw from the data RNG and train on the example batch: snapshot and restore
        # the RNG and the optimiser state.  With more than one rank the graph also holds the NCCL all-reduce: replay i
        # all-reduces and applies the gradient of step i-1 on a forked stream while the kNN graph of step i is built
        # (train_utils.PipelinedStep: same parameter sequence; a checkpoint written in that mode holds the parameters
        # BEFORE the update of the step whose loss is printed next to it).
        rng_state = dataset.rng.get_state()
        example = torch.from_numpy(dataset.get_minibatch(args.batch_size)).to(dev)
        dataset.rng.set_state(rng_state)
        state = [t.clone() for t in (store.flat, store.m, store.v, store.step_dev)]
        if world == 1:
            graphed = nbpc.train_utils.GraphedStep(lambda x: train_step(x, dev_step=True), (example,), warmup=2)
            with torch.no_grad():
                for t, s0 in zip((store.flat, store.m, store.v, store.step_dev), state):
                    t.copy_(s0)
        else:
            # NCCL all-reduce + Adam of step i-1 are captured on a forked stream at the start of replay i
            graphed = nbpc.train_utils.PipelinedStep(lambda x: None, lambda ctx, x: grad_step(x), store, adam, world, (example,), warmup=2)

    def run(batch):
        if graphed is None:
            return train_step(batch)
        return graphed(torch.from_numpy(batch).to(dev, non_blocking=True))

    tstart = time.time()
    if rank == 0:
        print(f"\nTraining:\n{'=' * 78}")
    for step in range(args.num_iters):
        batch = dataset.get_minibatch(args.batch_size)
        loss = run(batch)
        if (step + 1) % args.checkpoint == 0:                                          # train.py:117-120
            save(step)
            if rank == 0:
                print(f"Checkpoint {step + 1:>6} :  {float(loss.detach()):.8f}")
    if graphed is not None:
        if world > 1:
            graphed.flush()                                                            # the update of the last step
        graphed.close()
    torch.cuda.synchronize()
    if rank == 0:
        print(f"Training finished!\n\tElapsed time: {(time.time() - tstart) / 60:.2f}m")
    save(args.num_iters)

    # ---- evaluation (train.py:140-182), rank 0
    test_error = None
    if rank == 0:
        print(f"\nEvaluation:\n{'=' * 78}")
        nb_test = args.num_test // args.batch_size
        test_error = np.zeros((nb_test,), dtype=np.float32)
        preds = np.zeros((2, nb_test * args.batch_size) + dataset.X_test.shape[1:2] + (args.channels[-1],), dtype=np.float32)
        with torch.no_grad():
            for j in range(nb_test):
                p, q = args.batch_size * j, args.batch_size * (j + 1)
                batch = dataset.X_test[p:q]
                pred, err = forward(batch)
                preds[0, p:q] = batch[..., 6:]
                preds[1, p:q] = pred.cpu().numpy()
                test_error[j] = float(err)
                print(f"val_err, {j} : {test_error[j]}")
        np.save(os.path.join(out_dir, "Results", "error_test.npy"), test_error)        # utils.py:488-498
        np.save(os.path.join(out_dir, "Results", f"X_{args.data_idx}_prediction.npy"), preds)
        if nb_test:
            print(f"\n# Test error: median {np.median(test_error):.8f}, mean {test_error.mean():.8f} +- {test_error.std():.8f}")
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    return test_error


if __name__ == "__main__":
    main()

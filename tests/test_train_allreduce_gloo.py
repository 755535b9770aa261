"""Data-parallel half of the training step on CPU: world_size-2 gloo groups run the gradient all-reduce that
``UNetTrainer.step`` performs between backward and the optimizer (sum over ranks, then the 1/world factor), with one and
with several buckets.  With equal local batches the averaged gradient must equal the gradient of the global batch — the
property torch DDP guarantees and the B200 path relies on.  The gradients here come from the fp32 oracle network on a
small image (the CUDA kernels need a GPU); what is tested is the host-side reduction logic."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from chessvision import training


def _flat_grads(model):
    return torch.cat([p.grad.flatten() for p in model.parameters()])


def worker(rank, world, port, buckets, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import train as otrain
        torch.manual_seed(0)
        x = torch.rand(2 * world, 3, 32, 32)
        t = (torch.rand(2 * world, 1, 32, 32) > 0.5).float()
        model = otrain.new_model(0)
        # BatchNorm statistics are local per rank (no SyncBN in the reference): evaluate with frozen statistics so that
        # the global-batch gradient is exactly the mean of the shard gradients.
        model.eval()
        for p in model.parameters():
            p.grad = None
        otrain.loss_fn(model(x), t).backward()
        want = _flat_grads(model)
        for p in model.parameters():
            p.grad = None
        lo, hi = 2 * rank, 2 * rank + 2
        otrain.loss_fn(model(x[lo:hi]), t[lo:hi]).backward()
        flat = _flat_grads(model).clone()
        order = []
        if buckets == "layer-reverse":   # explicit ranges taken from the end of the buffer, like the trainer's cvb_train_buckets
            n = flat.numel()
            cuts = [n, (7 * n) // 8, n // 2, n // 5, 0]
            buckets = list(zip(cuts[1:], cuts[:-1]))
        scale = training.allreduce_gradients(flat, buckets=buckets, before_bucket=order.append)
        assert order == list(range(buckets if isinstance(buckets, int) else len(buckets)))
        got = flat * scale
        # BCE is a mean over all pixels and Dice a mean over samples: equal shards -> mean of shard losses = global loss
        err = float((got - want).norm() / want.norm())
        q.put((rank, err, scale))
    finally:
        dist.destroy_process_group()


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("buckets", [1, 5, "layer-reverse"])
def test_allreduce_averages_shard_gradients(buckets):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=worker, args=(r, world, port, buckets, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for _, err, scale in results:
        assert scale == 0.5
        assert err < 1e-4, err


def test_single_process_is_a_no_op():
    g = torch.arange(10.0)
    assert training.allreduce_gradients(g) == 1.0
    assert torch.equal(g, torch.arange(10.0))

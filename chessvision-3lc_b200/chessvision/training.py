"""UNet board-extractor training step on the B200 (host-side mirror of ``scripts/train/train_unet.py:236-245,293-323``).

``UNetTrainer.step(images, masks)`` is the body of the reference's training loop for one batch::

    masks_pred = model(images)                                        # BatchNorm in training mode
    loss = BCEWithLogitsLoss()(masks_pred, true_masks) + dice_loss(sigmoid(masks_pred), true_masks)
    optimizer.zero_grad(); loss.backward()
    clip_grad_norm_(model.parameters(), 1.0); optimizer.step()        # RMSprop(lr, weight_decay=1e-8, momentum=0.999)

executed by hand-written sm_100a kernels behind the C ABI (``cvb_train_*``): tcgen05 implicit-GEMM convolutions for the
forward pass and the data gradients, a tcgen05 MN-major GEMM for the weight gradients, fp16 operands with fp32 master
weights, accumulation and optimizer state (the arithmetic of the reference's ``--amp`` path).

Data parallelism (BASELINE.json configs[4]; the reference itself is single-device): one process per GPU, every rank holds
a replica and calls ``step`` on its own micro-batch; the flat fp32 gradient buffer is all-reduced (sum) over NCCL /
NVLink in four layer-reverse buckets, each launched from a communication stream as soon as the backward pass has
finished that bucket (``cvb_train_bucket_wait``), so the transfers overlap the rest of the backward pass; the optimizer
applies ``1/world_size`` — the semantics of torch DDP (gradient average, local BatchNorm statistics; the reference has no
SyncBN).  There is no CPU fallback.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import _native


def allreduce_gradients(flat: torch.Tensor, group=None, buckets=1, before_bucket=None) -> float:
    """Sum ``flat`` (a 1-D gradient buffer) over the ranks of ``group`` in place and return the factor that turns the sum
    into the average (1/world).  ``buckets``: an int (that many equal contiguous ranges) or a list of ``(lo, hi)`` ranges,
    all-reduced in the given order as independent asynchronous collectives; ``before_bucket(i)`` (optional) runs right
    before collective ``i`` is enqueued (the trainer uses it to make the communication stream wait for the event of that
    bucket).  Works on NCCL (device tensors) and gloo (host tensors, used by the CPU tests)."""
    if not dist.is_available() or not dist.is_initialized():
        return 1.0
    world = dist.get_world_size(group)
    if world == 1:
        return 1.0
    n = flat.numel()
    if isinstance(buckets, int):
        k = max(1, min(buckets, n))
        bounds = [(i * n) // k for i in range(k + 1)]
        ranges = list(zip(bounds, bounds[1:]))
    else:
        ranges = [(int(a), int(b)) for a, b in buckets]
        covered = sorted(ranges)
        assert covered[0][0] == 0 and covered[-1][1] == n and all(x[1] == y[0] for x, y in zip(covered, covered[1:])), \
            "gradient buckets must tile the buffer"
    handles = []
    for i, (a, b) in enumerate(ranges):
        if b <= a:
            continue
        if before_bucket is not None:
            before_bucket(i)
        handles.append(dist.all_reduce(flat[a:b], op=dist.ReduceOp.SUM, group=group, async_op=True))
    for h in handles:
        h.wait()
    return 1.0 / world


class UNetTrainer:
    """One replica of UNet(n_channels=3, n_classes=1) in training mode on one GPU."""

    def __init__(self, state_dict, batch_size: int = 2, learning_rate: float = 1e-6, device: int = 0, engine: "_native.Engine | None" = None,
                 process_group=None, **train_options):
        self.engine = engine if engine is not None else _native.Engine(device, max_batch=1)
        self._own_engine = engine is None
        self.engine.train_create(state_dict, batch=batch_size, **train_options)
        self.batch_size = batch_size
        self.learning_rate = learning_rate
        self.process_group = process_group
        self._grads = self.engine.train_grads()
        self.buckets = self.engine.train_buckets()      # layer-reverse ranges, in the order the backward pass completes them
        self.n_buckets = len(self.buckets)
        self._comm_stream = torch.cuda.Stream(device=self.engine.device)
        self.global_step = 0

    def step(self, images: torch.Tensor, true_masks: torch.Tensor) -> torch.Tensor:
        """One optimisation step; returns the (local) batch loss as a 1-element device tensor without synchronising."""
        images = images.to(device=self.engine.device, dtype=torch.float32).contiguous()
        true_masks = true_masks.to(device=self.engine.device, dtype=torch.float32).contiguous()
        loss = self.engine.train_forward_backward(images, true_masks)   # enqueues the whole pass and records one event per bucket
        scale = self._allreduce_overlapped()
        self.engine.train_optimizer_step(self.learning_rate, scale)
        self.global_step += 1
        return loss

    def _allreduce_overlapped(self) -> float:
        """All-reduce the gradient buckets in the order the backward pass finishes them.  The backward kernels are already
        queued on the compute stream; collective ``b`` is issued from a communication stream that waits for bucket ``b``'s
        event only, so it runs over NVLink while the compute stream is still producing the earlier layers' gradients.  The
        compute stream waits for all collectives before the optimizer touches the buffer."""
        if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(self.process_group) == 1:
            return 1.0
        compute = torch.cuda.current_stream(self.engine.device)
        with torch.cuda.stream(self._comm_stream):
            return self._reduce_on_comm_stream(compute)

    def _reduce_on_comm_stream(self, compute) -> float:
        world = dist.get_world_size(self.process_group)
        handles = []
        for b, (lo, hi) in enumerate(self.buckets):
            self.engine.train_bucket_wait(b, self._comm_stream)
            handles.append(dist.all_reduce(self._grads[lo:hi], op=dist.ReduceOp.SUM, group=self.process_group, async_op=True))
        with torch.cuda.stream(compute):
            for h in handles:
                h.wait()            # stream-level wait: the compute stream resumes after the last collective
        return 1.0 / world

    def forward_backward(self, images, true_masks):
        return self.engine.train_forward_backward(images, true_masks)

    def gradients(self):
        """Gradients of the last forward_backward as a CPU dict in state_dict layout (for tests / inspection)."""
        return self.engine.train_export(grads=True)

    def state_dict(self):
        """Current parameters and BatchNorm running statistics, CPU tensors in the reference's state_dict layout; feed it
        to ``ChessVision`` / ``cvb_load_unet`` or ``torch.save({"model_state_dict": ...})`` (train_unet.py:31-40)."""
        return self.engine.train_export(grads=False)

    def close(self):
        if self._own_engine:
            self.engine.close()

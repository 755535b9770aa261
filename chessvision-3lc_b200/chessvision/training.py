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


class ClassifierTrainer:
    """timm resnet18(num_classes=13, in_chans=1) in training on one GPU: the reference's ``train`` / ``validate`` loops and its
    optimizer / scheduler (``scripts/train/train_classifier.py:63-113,218-221``: CrossEntropyLoss, Adam(lr=1e-3), StepLR(step_size=4,
    gamma=0.1)) behind the C ABI (``cvb_cls_train_*``, fp32 CUDA kernels in ``csrc/train_cls.cu``).

    ``step(data, target)`` is the loop body ``optimizer.zero_grad(); output = model(data); loss = criterion(output, target);
    loss.backward(); optimizer.step()`` and returns ``(loss, correct)`` as device tensors; ``epoch_end()`` is ``scheduler.step()``;
    ``evaluate(data, target)`` is the body of ``validate`` (``model.eval()``).  Data parallel: as for ``UNetTrainer`` the flat
    gradient buffer is all-reduced over the process group before the optimizer step."""

    def __init__(self, state_dict, batch_size: int = 64, learning_rate: float = 1e-3, step_size: int = 4, gamma: float = 0.1, device: int = 0,
                 engine: "_native.Engine | None" = None, process_group=None, **train_options):
        self.engine = engine if engine is not None else _native.Engine(device, max_batch=1)
        self._own_engine = engine is None
        self.engine.cls_train_create(state_dict, batch=batch_size, **train_options)
        self.batch_size = batch_size
        self.base_lr = learning_rate
        self.step_size, self.gamma = step_size, gamma
        self.epoch = 0
        self.process_group = process_group
        self._grads = self.engine.cls_train_grads()
        self._int_keys = {k: v.clone() for k, v in state_dict.items() if torch.is_tensor(v) and not v.is_floating_point()}

    @property
    def learning_rate(self) -> float:
        """StepLR: base_lr * gamma ** (epoch // step_size)."""
        return self.base_lr * self.gamma ** (self.epoch // self.step_size)

    def epoch_end(self):
        self.epoch += 1

    def _prep(self, data, target):
        data = data.to(device=self.engine.device, dtype=torch.float32).contiguous()
        target = target.to(device=self.engine.device, dtype=torch.int32).contiguous() if target is not None else None
        return data, target

    def step(self, data: torch.Tensor, target: torch.Tensor):
        data, target = self._prep(data, target)
        loss, correct = self.engine.cls_train_forward_backward(data, target)
        scale = allreduce_gradients(self._grads, self.process_group)
        self.engine.cls_train_optimizer_step(self.learning_rate, scale)
        return loss, correct

    def forward_backward(self, data, target):
        data, target = self._prep(data, target)
        return self.engine.cls_train_forward_backward(data, target)

    def evaluate(self, data, target=None):
        """``model.eval()`` forward: (logits [B,13], loss, correct)."""
        data, target = self._prep(data, target)
        return self.engine.cls_train_forward(data, target, training=False)

    def gradients(self):
        return self.engine.cls_train_export(1)

    def state_dict(self):
        """Parameters + BatchNorm buffers, CPU tensors in the layout ``utils.get_classifier_model().load_state_dict`` /
        ``cvb_load_resnet18`` take (``num_batches_tracked`` counted up by the steps taken here)."""
        sd = self.engine.cls_train_export(0)
        steps = self.engine.cls_train_steps()
        for k, v in self._int_keys.items():
            sd[k] = v + steps if k.endswith("num_batches_tracked") else v
        return sd

    def optimizer_state_dict(self, names=None):
        """torch.optim.Adam's ``state_dict()`` layout (``state``: index -> step / exp_avg / exp_avg_sq; one param group), so that
        ``save_classifier_checkpoint`` (train_classifier.py:116-126) writes a checkpoint ``strip_optimizer`` understands."""
        m, v = self.engine.cls_train_export(2), self.engine.cls_train_export(3)
        names = list(names) if names is not None else [k for k in m if "running_" not in k]
        step = torch.tensor(float(self.engine.cls_train_steps()))
        state = {i: {"step": step.clone(), "exp_avg": m[k], "exp_avg_sq": v[k]} for i, k in enumerate(names)}
        group = {"lr": self.learning_rate, "betas": (self.engine.cls_cfg.beta1, self.engine.cls_cfg.beta2), "eps": self.engine.cls_cfg.eps,
                 "weight_decay": self.engine.cls_cfg.weight_decay, "amsgrad": False, "params": list(range(len(names)))}
        return {"state": state, "param_groups": [group]}

    def close(self):
        if self._own_engine:
            self.engine.close()

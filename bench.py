#!/usr/bin/env python
"""Benchmark of the image->FEN hot path (BASELINE.json metric: boards/sec image->FEN).

    python bench.py --gpus N --steps K --warmup W [--workload pipeline|unet-sweep|classify|train|decode] [--impl reference]

Workloads (BASELINE.json `configs`):
  pipeline    configs[3]  one step = the whole pipeline (UNet -> mask -> quad -> warp/crop -> ResNet-18 -> FEN) over B distinct
                          synthetic 512x512x3 boards per GPU; B is a multiple of the chunk (296 boards = two boards per SM, so
                          every conv launch is a whole number of waves) and the default B x 20 steps covers the 65,536 boards
                          of configs[3] on one GPU.  `value` = device-resident throughput (inputs in HBM, CUDA events, max over
                          ranks); `e2e` = the same through the host-buffer C-ABI entry point cvb_image_to_fen_host (pinned host
                          input, H2D and D2H inside the timed region); `e2e_api` = the same through the drop-in Python API
                          ChessVision.process_images with every output the reference's process_image returns.
  unet-sweep  configs[1]  UNet board-extractor forward alone, batch sweep 1..1036 at the reference input size.
  classify    configs[2]  square extraction (warp/crop gather) + piece classifier over 4096 boards with given quads.
  train       configs[4]  UNet training step, data-parallel, NCCL gradient all-reduce.
  decode      8(f) n2     JPEG decode front-end.
`roofline` describes the dominant kernel of the workload against the measured peaks (MEASURED_PEAKS.json).
`cpu_baseline` / `--impl reference`: the UNMODIFIED reference (`ChessVision.process_image` of chessvision/core.py:152-195,
imported from /root/reference or from the byte-for-byte staging oracle/_ref written by oracle/build_ref.sh) on the box's host
cores, in a process of its own (the product package has the same import name), on a bounded sample of the same boards.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
PRODUCT = str(ROOT / "chessvision-3lc_b200")

SEED = 20261017
UNET_GFLOP = 96.335           # UNet forward per board (2*MAC)
UNET_STEM_GFLOP = 0.2265      # inc.double_conv.0 (K = 27) runs inside the fused preprocessing kernel
CLS_GFLOP = 18.127            # ResNet-18 forward for 64 squares
WARP_BYTES_PER_BOARD = 512 * 512 * 3 + 512 * 512   # warp/crop: u8 BGR image read + u8 gray squares written
H2D_PER_BOARD = 512 * 512 * 3
D2H_PER_BOARD = 4 * 2 * 4 + 1 + 4 + 64 * 13 * 4 + 64 + 64 + 2 * 72   # quad, found, status, probs, labels x2, fen
D2H_FULL_PER_BOARD = D2H_PER_BOARD + 65536 * 4 + 65536 + 262144       # + logits, mask, board image
CHUNK = 148                   # boards per network chunk = SM count of a B200: every persistent conv grid is whole waves
PIPE_CHUNK = 296              # the pipeline workload's default: two boards per SM (per-launch prologue / drain amortised over twice the
                              # boards: +1.0 % device-resident against 148 in a same-box A/B, profiles/README.md; 17.8 GB of workspace)
CONFIG3_BOARDS = 65536

_REAL_STDOUT = None


def capture_stdout():
    """The contract is ONE JSON line on stdout.  Libraries print there too (NCCL's version banner at N > 1), so file
    descriptor 1 is pointed at stderr for the whole run and the JSON line is written to the saved original stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(text: str):
    out = _REAL_STDOUT or sys.stdout
    out.write(text + "\n")
    out.flush()


def use_product_package():
    for p in (str(ROOT), PRODUCT):
        if p not in sys.path:
            sys.path.insert(0, p)


def synthetic_boards(n_distinct: int, seed: int = SEED):
    """Generator A of SURVEY.md §8(d): a real data/test image under a random homography plus per-channel gain/offset,
    so that the trained UNet segments it.  Deterministic; every board of a call is distinct (its own homography/colour)."""
    import cv2
    files = sorted((ROOT / "tests" / "golden" / "data_test").glob("*/*"))
    base = [cv2.imread(str(f)) for f in files]
    rng = np.random.default_rng(seed)
    out = np.empty((n_distinct, 512, 512, 3), np.uint8)
    corners = np.array([[0, 0], [511, 0], [511, 511], [0, 511]], np.float32)
    for i in range(n_distinct):
        img = base[i % len(base)]
        dst = corners + rng.uniform(-24, 24, (4, 2)).astype(np.float32)
        M = cv2.getPerspectiveTransform(corners, dst)
        w = cv2.warpPerspective(img, M, (512, 512), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REPLICATE).astype(np.float32)
        w = w * rng.uniform(0.85, 1.15, 3).astype(np.float32) + rng.uniform(-12, 12, 3).astype(np.float32)
        out[i] = np.clip(np.rint(w), 0, 255).astype(np.uint8)
    return out


class ClockSampler:
    """nvidia-smi clocks line of B200_PROFILING.md, sampled while the timed region runs."""

    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index: int):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, pw, mx, reasons = [], [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = max(mx, float(r[2]))
                pw.append(float(r[3]))
                for k, nm in enumerate(names):
                    if r[5 + k].lower().startswith("active"):
                        reasons.add(nm)
            except (ValueError, IndexError):
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm),
                "power_w": float(np.median(pw)) if pw else None}


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", d.get("bf16_tflops")), d.get("hbm_gbs"), "measured (MEASURED_PEAKS.json: sustained bf16, HBM copy)"
    return 1590.0, 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel_class: str, boards_per_launch: int = 0):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of a kernel class, from the committed ncu `--set full` capture
    (profiles/<round>/traffic.json, written by profiles/traffic_from_ncu.py from the raw csv of the same command): a profiler
    counter cannot be read inside an un-profiled run, so the line cites the capture it comes from."""
    for rnd in ("r02", "r01"):
        f = ROOT / "profiles" / rnd / "traffic.json"
        if f.exists():
            d = json.load(open(f))
            if kernel_class in d:
                e = d[kernel_class]
                scale = boards_per_launch / e["boards"] if boards_per_launch else 1.0   # the capture ran one 148-board chunk per launch
                note = f", scaled to {boards_per_launch} boards per launch" if boards_per_launch and boards_per_launch != e["boards"] else ""
                return e["bytes_per_launch"] * scale, (f"profiles/{rnd}/traffic.json ({e['launches']} launches, {e['boards']} boards per launch{note}; "
                                                       f"{e['source']})")
    return None, "no ncu capture committed for this kernel class"


# =====================================================================================================================
# reference arm: the unmodified reference on the host cores
# =====================================================================================================================
def _reference_env():
    import torch
    sys.path.insert(0, str(ROOT))
    from oracle import ref_loader
    torch.set_num_threads(os.cpu_count() or 1)
    wdir = ROOT / "weights"
    cv = ref_loader.reference_pipeline(str(wdir / "best_extractor.pth"), str(wdir / "best_classifier.pth"))
    where = ref_loader.reference_root()
    return cv, ("/root/reference (checkout)" if where == ref_loader.REF else "oracle/_ref (byte-for-byte staging of the checkout, oracle/build_ref.sh)")


def run_reference(args):
    """--impl reference: ChessVision.process_image of the unmodified reference (chessvision/core.py:152-195), CPU, all host
    threads, on the first boards of the same synthetic stream.  Rank 0 only; the other ranks exit without work."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import torch
    cores = os.cpu_count() or 1
    try:
        cv, where = _reference_env()
    except ImportError as e:   # neither the checkout nor its staging exists: say so (the oracle port is test infrastructure only)
        emit(json.dumps({"impl": "reference", "unavailable": f"reference sources not found: {e}"}))
        return
    per_step = max(1, args.cpu_boards_per_step)
    common = {"impl": "reference", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "scaling": "weak",
              "vs_baseline": None, "dtype": "f32"}

    if args.workload == "unet-sweep":
        import cv2
        ref_mod = sys.modules["chessvision"]
        boards = synthetic_boards(per_step)
        xs = [torch.Tensor(np.array([cv2.resize(b, ref_mod.constants.INPUT_SIZE, interpolation=cv2.INTER_AREA)])).div(255).permute(0, 3, 1, 2) for b in boards]
        with torch.no_grad():
            for i in range(min(args.warmup, 2)):
                cv.board_extractor(xs[i % per_step])
            t0 = time.perf_counter()
            for s in range(args.steps):
                for x in xs:
                    cv.board_extractor(x)
            dt = time.perf_counter() - t0
        thr = per_step * args.steps / dt
        emit(json.dumps({**common, "metric": "UNet board-extractor forward boards/sec", "value": thr, "unit": "boards/s",
                         "ms_per_step": 1000.0 * dt / args.steps, "data": "synthetic boards, locally trained weights",
                         "config": {"workload": "configs[1]: UNet forward alone (reference: batch 1 per call, core.py:215-220)", "boards_per_step": per_step},
                         "cpu_baseline": {"value": thr, "unit": "boards/s", "cores": cores, "kind": "reference",
                                          "sample": f"{per_step * args.steps} calls of the reference's UNet(3,1) at batch 1, fp32, {cores} threads; {where}"},
                         "e2e": {"value": thr, "unit": "boards/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    boards = synthetic_boards(max(per_step, 4))
    if args.workload == "classify":
        import cv2
        ref_mod = sys.modules["chessvision"]
        quads = []
        for b in boards:   # given quads: the reference's own extraction, untimed
            r = cv.extract_board(b)
            quads.append(r.quadrangle)
        items = [(b, q) for b, q in zip(boards, quads) if q is not None] or None
        assert items, "no board found in the sample"

        def one(b, q):
            board = ref_mod.utils.extract_perspective(b, q, ref_mod.constants.BOARD_SIZE)
            board = cv2.flip(cv2.cvtColor(board, cv2.COLOR_BGR2GRAY), 1)
            return cv.classify_position(board)
        for i in range(min(args.warmup, 2)):
            one(*items[i % len(items)])
        t0 = time.perf_counter()
        for s in range(args.steps):
            for k in range(per_step):
                one(*items[k % len(items)])
        dt = time.perf_counter() - t0
        thr = per_step * args.steps / dt
        emit(json.dumps({**common, "metric": "square extraction + piece classification boards/sec", "value": thr, "unit": "boards/s",
                         "ms_per_step": 1000.0 * dt / args.steps, "data": "synthetic boards, locally trained weights",
                         "config": {"workload": "configs[2]: warp/crop + classifier with given quads (utils.py:115-132, core.py:225-249)", "boards_per_step": per_step},
                         "cpu_baseline": {"value": thr, "unit": "boards/s", "cores": cores, "kind": "reference",
                                          "sample": f"{per_step * args.steps} boards, extract_perspective + BGR2GRAY + flip + classify_position, fp32, {cores} threads; {where}"},
                         "e2e": {"value": thr, "unit": "boards/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    # pipeline
    for i in range(min(args.warmup, 3)):
        cv.process_image(boards[i % len(boards)])
    found = 0
    t0 = time.perf_counter()
    for s in range(args.steps):
        for k in range(per_step):
            found += int(cv.process_image(boards[(s * per_step + k) % len(boards)]).position is not None)
    dt = time.perf_counter() - t0
    n = per_step * args.steps
    thr = n / dt
    # per-stage split: extract_board and classify_position timed separately on a few boards
    t_ext = t_cls = 0.0
    n_cls = 0
    for b in boards[:4]:
        t1 = time.perf_counter()
        r = cv.extract_board(b)
        t_ext += time.perf_counter() - t1
        if r.board_image is not None:
            t1 = time.perf_counter()
            cv.classify_position(r.board_image)
            t_cls += time.perf_counter() - t1
            n_cls += 1
    stage_ms = {"extract_board": 1000.0 * t_ext / 4, "classify_position": 1000.0 * t_cls / max(n_cls, 1)}
    print(f"[reference] {n} boards in {dt:.2f} s on {cores} threads; per-stage CPU ms/board: {stage_ms}", file=sys.stderr)
    emit(json.dumps({**common, "metric": "boards/sec image->FEN", "value": thr, "unit": "boards/s", "ms_per_step": 1000.0 * dt / args.steps,
                     "data": "synthetic (data/test images under random homographies, seed 20261017); locally trained weights",
                     "config": {"workload": "configs[3]: full image->FEN pipeline on synthetic 512x512x3 boards (reference: one board per call)",
                                "boards_per_step": per_step},
                     "cpu_baseline": {"value": thr, "unit": "boards/s", "cores": cores, "kind": "reference",
                                      "sample": f"{n} boards of the same synthetic stream through ChessVision.process_image (core.py:152-195), fp32 torch + cv2, "
                                                f"{cores} threads ({dt:.1f} s); {where}", "stage_ms_per_board": stage_ms},
                     "e2e": {"value": thr, "unit": "boards/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                     "found_rate": found / max(n, 1)}))


def cpu_baseline_subprocess(workload: str, boards: int, timeout: int = 600):
    """cpu_baseline leg of the B200 arm: the reference arm on a bounded sample in a process of its own (the reference
    package and the product package are both called `chessvision`, and this process holds the GPU)."""
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", RANK="0", WORLD_SIZE="1", LOCAL_RANK="0")
    for k in ("MASTER_ADDR", "MASTER_PORT", "TORCHELASTIC_RUN_ID", "GROUP_RANK", "LOCAL_WORLD_SIZE"):
        env.pop(k, None)
    cmd = [sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--workload", workload, "--gpus", "1", "--steps", "1", "--warmup", "2",
           "--cpu-boards-per-step", str(boards)]
    try:
        p = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env)
        line = [l for l in p.stdout.splitlines() if l.strip()][-1]
        d = json.loads(line)
        if "cpu_baseline" not in d:
            return {"value": None, "unit": "boards/s", "cores": os.cpu_count() or 1, "kind": "reference", "sample": d.get("unavailable", "unavailable")}
        cb = d["cpu_baseline"]
        if "found_rate" in d:
            cb["found_rate"] = d["found_rate"]
        return cb
    except Exception as e:   # noqa: BLE001  (a failed baseline must not lose the GPU measurement)
        return {"value": None, "unit": "boards/s", "cores": os.cpu_count() or 1, "kind": "reference", "sample": f"failed: {e}"}


# =====================================================================================================================
# training workload (configs[4])
# =====================================================================================================================
def synthetic_training_batch(b: int, seed: int):
    """configs[4] input: fp32 [b,3,256,256] images in [0,1] (smooth colour field + checkerboard inside a random quad) and
    the quad's {0,1} mask [b,1,256,256] -- the shape and value range scripts/train/train_unet.py feeds the UNet."""
    rng = np.random.default_rng(seed)
    yy, xx = np.meshgrid(np.arange(256, dtype=np.float32), np.arange(256, dtype=np.float32), indexing="ij")
    imgs = np.empty((b, 3, 256, 256), np.float32)
    masks = np.empty((b, 1, 256, 256), np.float32)
    for i in range(b):
        cx, cy = 128 + 30 * (rng.random(2) - 0.5)
        r = 60 + 40 * rng.random()
        th = 0.6 * (rng.random() - 0.5)
        u = (xx - cx) * np.cos(th) + (yy - cy) * np.sin(th)
        v = -(xx - cx) * np.sin(th) + (yy - cy) * np.cos(th)
        m = ((np.abs(u) < r) & (np.abs(v) < r * (0.8 + 0.2 * rng.random()))).astype(np.float32)
        base = np.kron(rng.random((3, 8, 8)).astype(np.float32), np.ones((32, 32), np.float32))
        checker = (np.floor(u / (r / 4)) + np.floor(v / (r / 4))) % 2
        imgs[i] = np.clip(0.6 * base + 0.4 * m * checker + 0.05 * rng.random((3, 256, 256)).astype(np.float32), 0, 1)
        masks[i, 0] = m
    return imgs, masks


def run_train_reference(args):
    """--impl reference --workload train: the fp32 oracle of the reference's training step (oracle/train.py, pinned to
    scripts/train/train_unet.py, which itself needs the 3LC package and cannot be imported here) on the host cores, rank 0."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import torch
    sys.path.insert(0, str(ROOT))
    from oracle import train as otrain
    torch.set_num_threads(os.cpu_count() or 1)
    b = max(1, min(args.train_batch, 2))
    model = otrain.new_model(0)
    opt = otrain.make_optimizer(model, 1e-6)
    imgs, masks = synthetic_training_batch(b, SEED)
    imgs, masks = torch.from_numpy(imgs), torch.from_numpy(masks)
    for _ in range(min(args.warmup, 1)):
        otrain.train_step(model, opt, imgs, masks)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        otrain.train_step(model, opt, imgs, masks)
    dt = time.perf_counter() - t0
    thr = b * args.steps / dt
    cores = os.cpu_count() or 1
    emit(json.dumps({
        "impl": "reference", "metric": "UNet training images/sec (fwd+bwd+clip+RMSprop)", "value": thr, "unit": "images/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": {"workload": "configs[4]: UNet training step", "batch_per_step": b},
        "cpu_baseline": {"value": thr, "unit": "images/s", "cores": cores, "kind": "port",
                         "sample": f"{args.steps} steps of batch {b}, fp32 torch oracle of train_unet.py's step, {cores} threads"},
        "e2e": {"value": thr, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def _dist_env():
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    return world, rank, local_rank, dev, barrier, max_over_ranks


def run_train(args):
    """--workload train (BASELINE.json configs[4]): UNet training step, data-parallel; the gradient all-reduce runs over NCCL
    in layer-reverse buckets launched while the backward pass is still producing the earlier layers' gradients
    (chessvision/training.py).  One step = fwd + loss + bwd + all-reduce + clip + RMSprop on `--train-batch` images per GPU."""
    import torch
    import torch.distributed as dist
    use_product_package()
    from chessvision import utils
    from chessvision.training import UNetTrainer

    world, rank, local_rank, dev, barrier, max_over_ranks = _dist_env()
    B = args.train_batch
    sd = utils.load_state_dict(str(ROOT / "weights" / "best_extractor.pth"))[0]
    tr = UNetTrainer(sd, batch_size=B, learning_rate=1e-6, device=local_rank)
    imgs, masks = synthetic_training_batch(B, SEED + rank)
    h_img, h_mask = torch.from_numpy(imgs).pin_memory(), torch.from_numpy(masks).pin_memory()
    d_img, d_mask = h_img.to(dev), h_mask.to(dev)

    stream = torch.cuda.current_stream()
    for _ in range(args.warmup):
        tr.step(d_img, d_mask)
    barrier()
    l0 = tr.engine.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        barrier()
        e0.record(stream)
        for _ in range(args.steps):
            loss = tr.step(d_img, d_mask)
        e1.record(stream)
        barrier()
        dev_ms = max_over_ranks(e0.elapsed_time(e1))
        # end to end: pinned host batch -> H2D -> step -> loss D2H, every step
        for _ in range(args.warmup):
            float(tr.step(h_img.to(dev, non_blocking=True), h_mask.to(dev, non_blocking=True)).item())
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            last = float(tr.step(h_img.to(dev, non_blocking=True), h_mask.to(dev, non_blocking=True)).item())
        barrier()
        e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1000.0)
    launches = tr.engine.launch_count() - l0
    clocks = clk.summary()
    total = B * args.steps * world
    value = total / (dev_ms / 1000.0)
    peak_tf, _, peak_src = measured_peaks()
    tflops = 3 * UNET_GFLOP * value / 1000.0 / world
    if rank == 0:
        emit(json.dumps({
            "metric": "UNet training images/sec (fwd+bwd+clip+RMSprop)", "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f16 operands, f32 master weights/accumulation", "data": "synthetic (quad masks + checkerboard images), trained start weights",
            "config": {"workload": "configs[4]: UNet board-extractor training step, data-parallel", "batch_per_gpu": B,
                       "l2": f"activations + gradients of one step ({B} x ~0.5 GB) exceed L2, no flush",
                       "parallelism": f"dp{world}, NCCL all-reduce of 31.0 M fp32 gradients per step in {getattr(tr, "n_buckets", 1)} layer-reverse bucket(s) "
                                      f"overlapped with the backward pass"},
            "e2e": {"value": total / (e2e_ms / 1000.0), "unit": "images/s", "h2d_bytes_per_step": B * 4 * 256 * 256 * 4, "d2h_bytes_per_step": 4,
                    "ms_per_step": e2e_ms / args.steps, "last_loss": last},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "kernel": "whole step (conv_tc fwd + dgrad, wgrad_tc)", "achieved": tflops, "peak": peak_tf,
                         "unit": "TFLOP/s per GPU", "frac": tflops / peak_tf if peak_tf else None, "traffic": None, "peak_source": peak_src,
                         "algorithmic_gflop_per_image": 3 * UNET_GFLOP},
            "clocks": clocks, "loss": float(loss.item())}))
    tr.close()
    if world > 1:
        dist.destroy_process_group()


def run_train_classifier(args):
    """--workload train-classifier: the piece-classifier training step of scripts/train/train_classifier.py
    (resnet18 on 64x64 squares, CrossEntropyLoss, Adam) through chessvision.training.ClassifierTrainer, `--train-batch` squares
    per GPU and step (default 256), data-parallel with an NCCL all-reduce of the flat gradient buffer.  fp32 CUDA-core kernels
    (csrc/train_cls.cu); `cpu_baseline` is the same step in fp32 PyTorch on the host cores (the oracle of the parity tests)."""
    import torch
    import torch.distributed as dist
    use_product_package()
    from chessvision import utils
    from chessvision.training import ClassifierTrainer

    world, rank, local_rank, dev, barrier, max_over_ranks = _dist_env()
    B = args.train_batch if args.train_batch != 8 else 256
    sd = utils.load_state_dict(str(ROOT / "weights" / "best_classifier.pth"))[0]
    tr = ClassifierTrainer(sd, batch_size=B, learning_rate=1e-3, device=local_rank)
    g = torch.Generator().manual_seed(SEED + rank)
    h_x = torch.rand(B, 1, 64, 64, generator=g).pin_memory()
    h_t = torch.randint(0, 13, (B,), generator=g, dtype=torch.int32).pin_memory()
    d_x, d_t = h_x.to(dev), h_t.to(dev)
    stream = torch.cuda.current_stream()
    for _ in range(args.warmup):
        tr.step(d_x, d_t)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        barrier()
        e0.record(stream)
        for _ in range(args.steps):
            loss, correct = tr.step(d_x, d_t)
        e1.record(stream)
        barrier()
        dev_ms = max_over_ranks(e0.elapsed_time(e1))
        t0 = time.perf_counter()
        for _ in range(args.steps):   # end to end: pinned host batch -> H2D -> step -> loss D2H
            last = float(tr.step(h_x.to(dev, non_blocking=True), h_t.to(dev, non_blocking=True))[0].item())
        barrier()
        e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1000.0)
    clocks = clk.summary()
    total = B * args.steps * world
    value = total / (dev_ms / 1000.0)
    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline:
            sys.path.insert(0, str(ROOT))
            from oracle import nets
            torch.set_num_threads(os.cpu_count() or 1)
            m = nets.PieceResNet18()
            m.load_state_dict({k: v.float() if v.is_floating_point() else v for k, v in sd.items()})
            m.train()
            opt = torch.optim.Adam(m.parameters(), lr=1e-3)
            xb, tb = h_x[:64].clone(), h_t[:64].long()
            n_steps, t0 = 0, time.perf_counter()
            while time.perf_counter() - t0 < 10.0:
                opt.zero_grad()
                torch.nn.functional.cross_entropy(m(xb), tb).backward()
                opt.step()
                n_steps += 1
            dt = time.perf_counter() - t0
            cpu = {"value": 64 * n_steps / dt, "unit": "squares/s", "cores": os.cpu_count() or 1, "kind": "port",
                   "sample": f"{n_steps} steps of batch 64 of the same loop in fp32 PyTorch on the host ({dt:.1f} s)"}
        emit(json.dumps({
            "metric": "piece-classifier training squares/sec (fwd+bwd+Adam)", "value": value, "unit": "squares/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic squares, trained start weights",
            "config": {"workload": "8(f) n3: resnet18(1 -> 13) training step on 64x64 squares, CrossEntropy + Adam", "batch_per_gpu": B,
                       "parallelism": f"dp{world}, NCCL all-reduce of 11.2 M fp32 gradients per step"},
            "e2e": {"value": total / (e2e_ms / 1000.0), "unit": "squares/s", "h2d_bytes_per_step": B * 4096 * 4 + B * 4, "d2h_bytes_per_step": 4,
                    "ms_per_step": e2e_ms / args.steps, "last_loss": last},
            "algorithmic_gflop_per_square": 3 * CLS_GFLOP / 64, "tflops": 3 * CLS_GFLOP / 64 * value / 1000.0 / world,
            "cpu_baseline": cpu, "clocks": clocks, "loss": float(loss.item())}))
    tr.close()
    if world > 1:
        dist.destroy_process_group()


# =====================================================================================================================
# JPEG front-end (8(f) n2)
# =====================================================================================================================
def run_decode(args):
    """--workload decode: the JPEG front-end on the reference's data/test files, cycled to `--boards`
    images per step (default 4,096).  `value` = files -> pixels in HBM through the public call: from 1,024 images per call the
    entropy decoding runs on the device (one warp per image; the host only parses headers and packs the compressed bytes),
    below that on the host threads.  `host_path` repeats the measurement on 1,024 images with the host Huffman decoder."""
    import cv2
    import torch
    from concurrent.futures import ThreadPoolExecutor
    use_product_package()
    from chessvision import _native

    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if int(os.environ.get("RANK", "0")) != 0:
        return
    torch.cuda.set_device(local_rank)
    eng = _native.Engine(local_rank, max_batch=4)
    files = sorted((ROOT / "tests" / "golden" / "data_test").glob("*/*"))
    base = [f.read_bytes() for f in files]
    n = args.boards or 4096
    streams = [base[i % len(base)] for i in range(n)]
    for _ in range(args.warmup):
        img = eng.decode_jpeg(streams)
    torch.cuda.synchronize()
    l0 = eng.launch_count()
    with ClockSampler(local_rank) as clk:
        t0 = time.perf_counter()
        for _ in range(args.steps):
            img = eng.decode_jpeg(streams)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 1000.0 / args.steps
    launches = eng.launch_count() - l0
    # the same front-end with the host Huffman decoder (what every batch below 1,024 images uses)
    os.environ["CVB_JPEG_DEVICE_MIN"] = "0"
    small = streams[:1024]
    eng.decode_jpeg(small)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        eng.decode_jpeg(small)
    torch.cuda.synchronize()
    host_ms = (time.perf_counter() - t0) * 1000.0 / 3
    del os.environ["CVB_JPEG_DEVICE_MIN"]
    ref = cv2.imdecode(np.frombuffer(streams[0], np.uint8), cv2.IMREAD_COLOR)
    assert np.array_equal(img[0].cpu().numpy(), ref), "decode differs from cv2.imdecode"
    threads = min(os.cpu_count() or 4, 32)                                   # what the library uses for its Huffman threads
    cv2.setNumThreads(1)
    sample = (streams * (1 + 8192 // len(streams)))[:8192]              # bounded sample: ~8k decodes, a few seconds of CPU work
    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(lambda b: cv2.imdecode(np.frombuffer(b, np.uint8), cv2.IMREAD_COLOR), sample[:threads]))
        t0 = time.perf_counter()
        list(ex.map(lambda b: cv2.imdecode(np.frombuffer(b, np.uint8), cv2.IMREAD_COLOR), sample, chunksize=16))
        cpu_s = time.perf_counter() - t0
    _, hbm, src = measured_peaks()
    px = 512 * 512
    emit(json.dumps({
        "metric": "JPEG decode images/sec (files -> u8 BGR in HBM)", "value": n / (ms / 1000.0), "unit": "images/s", "n_gpus": 1,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int32 / u8", "data": "the reference's 38 data/test JPEGs (512x512, 4:2:0), cycled",
        "config": {"workload": "8(f) n2: JPEG decode front-end", "images_per_step": n, "host_threads": threads,
                   "l2": f"coefficients + pixels of one step ({n} x 1.5 MB) exceed L2, no flush"},
        "e2e": {"value": n / (ms / 1000.0), "unit": "images/s", "h2d_bytes_per_step": int(sum(len(b) for b in streams)) if n >= 1024 else int(n * px * 3),
                "d2h_bytes_per_step": 0},
        "host_path": {"value": len(small) / (host_ms / 1000.0), "unit": "images/s", "images": len(small),
                      "note": "CVB_JPEG_DEVICE_MIN=0: Huffman decoding on the host threads, 3 B/px of coefficients over PCIe"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": "k_jpeg_idct + k_jpeg_color", "achieved": None, "peak": hbm, "unit": "GB/s", "frac": None,
                     "traffic": None, "peak_source": src, "algorithmic_bytes_per_image": 9 * px,
                     "note": "idct: 3 B/px coefficients in + 1.5 out; colour: 1.5 in + 3 out; per-launch durations: profiles/ ncu launch list"},
        "cpu_baseline": {"value": len(sample) / cpu_s, "unit": "images/s", "cores": threads, "kind": "reference",
                         "sample": f"cv2.imdecode of {len(sample)} of the same files on {threads} threads ({cpu_s:.1f} s)"},
        "clocks": clk.summary()}))
    eng.close()


# =====================================================================================================================
# configs[1]: UNet forward alone, batch sweep
# =====================================================================================================================
def run_unet_sweep(args):
    import torch
    use_product_package()
    from chessvision import _native, utils
    world, rank, local_rank, dev, barrier, max_over_ranks = _dist_env()
    if rank != 0:   # the sweep is a single-GPU configuration ("on 1xB200")
        if world > 1:
            torch.distributed.destroy_process_group()
        return
    eng = _native.Engine(local_rank, max_batch=CHUNK)
    eng.load_unet(utils.load_state_dict(str(ROOT / "weights" / "best_extractor.pth"))[0])
    top = 7 * CHUNK   # 1036: the "1024" end of the sweep as whole chunks
    distinct = synthetic_boards(top, SEED)
    host = torch.from_numpy(distinct).pin_memory()
    img = host.to(dev)
    stream = torch.cuda.current_stream()
    peak_tf, _, peak_src = measured_peaks()
    sweep = []
    clocks = None
    for b in (1, 2, 4, 8, 16, 32, 64, 128, CHUNK, 2 * CHUNK, 4 * CHUNK, top):
        x = img[:b]
        reps = max(2, min(args.steps * 8, int(2048 / b)))   # a few milliseconds at least per point, `steps` passes at the top
        if b == top:
            reps = args.steps
        for _ in range(args.warmup):
            eng.unet_forward(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = eng.launch_count()
        ctxm = ClockSampler(local_rank) if b == top else None
        if ctxm:
            ctxm.__enter__()
        e0.record(stream)
        for _ in range(reps):
            eng.unet_forward(x)
        e1.record(stream)
        torch.cuda.synchronize()
        if ctxm:
            ctxm.__exit__()
            clocks = ctxm.summary()
        ms = e0.elapsed_time(e1) / reps
        sweep.append({"batch": b, "ms": ms, "boards_per_s": b / (ms / 1000.0), "tflops": UNET_GFLOP * b / ms, "launches_per_pass": (eng.launch_count() - l0) // reps})
    topline = sweep[-1]
    # e2e at the top batch: pinned host images -> H2D -> forward -> D2H of the masks, every pass
    for _ in range(2):
        eng.unet_forward(host.to(dev, non_blocking=True))[1].cpu()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        eng.unet_forward(host.to(dev, non_blocking=True))[1].cpu()
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1000.0 / args.steps
    traffic, traffic_src = ncu_traffic("unet_conv_tc")
    cpu = None if args.no_cpu_baseline else cpu_baseline_subprocess("unet-sweep", 24)
    emit(json.dumps({
        "metric": "UNet board-extractor forward boards/sec", "value": topline["boards_per_s"], "unit": "boards/s", "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": topline["ms"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16",
        "data": "synthetic (data/test images under random homographies); locally trained weights",
        "config": {"workload": f"configs[1]: UNet board-extractor forward alone (512x512x3 u8 in -> INTER_AREA -> UNet(3,1) @256^2 -> logits + mask), "
                               f"batch sweep 1..{top}; the headline value is batch {top} (= 7 chunks of {CHUNK})", "chunk": CHUNK,
                   "l2": "activations of one chunk (~9 GB) exceed L2, no flush"},
        "sweep": sweep,
        "e2e": {"value": top / (e2e_ms / 1000.0), "unit": "boards/s", "h2d_bytes_per_step": top * H2D_PER_BOARD, "d2h_bytes_per_step": top * 65536, "ms_per_step": e2e_ms},
        "gpu_launches": int(topline["launches_per_pass"] * args.steps),
        "roofline": {"bound": "tensor", "kernel": "UNet forward (tcgen05 implicit-GEMM convs + stem)", "achieved": topline["tflops"], "peak": peak_tf,
                     "unit": "TFLOP/s", "frac": topline["tflops"] / peak_tf if peak_tf else None, "traffic": traffic, "traffic_source": traffic_src,
                     "peak_source": peak_src, "algorithmic_gflop_per_board": UNET_GFLOP,
                     "best_tflops_in_sweep": max(s["tflops"] for s in sweep)},
        "cpu_baseline": cpu, "clocks": clocks}))
    eng.close()
    if world > 1:
        torch.distributed.destroy_process_group()


# =====================================================================================================================
# configs[2]: square extraction + piece classifier over 4096 boards with given quads
# =====================================================================================================================
def run_classify(args):
    import torch
    use_product_package()
    from chessvision import _native, utils
    world, rank, local_rank, dev, barrier, max_over_ranks = _dist_env()
    if rank != 0:
        if world > 1:
            torch.distributed.destroy_process_group()
        return
    B = args.boards or 4096
    eng = _native.Engine(local_rank, max_batch=CHUNK)
    wdir = ROOT / "weights"
    eng.load_unet(utils.load_state_dict(str(wdir / "best_extractor.pth"))[0])
    eng.load_resnet18(utils.load_state_dict(str(wdir / "best_classifier.pth"))[0])
    host = torch.from_numpy(synthetic_boards(B, SEED)).pin_memory()
    img = host.to(dev)
    # given quads: one untimed extraction pass
    quads, founds = [], []
    for off in range(0, B, 8 * CHUNK):
        _, mask = eng.unet_forward(img[off:off + 8 * CHUNK])
        q, f, _ = eng.mask_to_quad(mask)
        quads.append(q)
        founds.append(f)
    quad, found = torch.cat(quads), torch.cat(founds)
    h_quad, h_found = quad.cpu().pin_memory(), found.cpu().pin_memory()
    stream = torch.cuda.current_stream()

    def one_pass():
        board = eng.warp_squares(img, quad, found)
        return eng.classify(board, False)

    for _ in range(args.warmup):
        one_pass()
    torch.cuda.synchronize()
    eng.profile(True)
    l0 = eng.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        e0.record(stream)
        for _ in range(args.steps):
            one_pass()
        e1.record(stream)
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    launches = eng.launch_count() - l0
    stages = eng.profile_read()
    eng.profile(False)
    # e2e: pinned host images + quads -> H2D -> warp + classify -> D2H of probabilities, labels, FEN
    def e2e_pass():
        d_img = host.to(dev, non_blocking=True)
        board = eng.warp_squares(d_img, h_quad.to(dev, non_blocking=True), h_found.to(dev, non_blocking=True))
        return [t.cpu() for t in eng.classify(board, False)]
    for _ in range(2):
        e2e_pass()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_pass()
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1000.0 / args.steps
    peak_tf, peak_hbm, peak_src = measured_peaks()
    warp_ms = stages["warp"] / args.steps
    warp_gbs = WARP_BYTES_PER_BOARD * B / (warp_ms / 1000.0) / 1e9 if warp_ms > 0 else None
    cls_ms = (stages["resnet_stem"] + stages["resnet_conv_tc"] + stages["head"]) / args.steps
    traffic, traffic_src = ncu_traffic("warp_board")
    cpu = None if args.no_cpu_baseline else cpu_baseline_subprocess("classify", 96)
    emit(json.dumps({
        "metric": "square extraction + piece classification boards/sec", "value": B / (ms / 1000.0), "unit": "boards/s", "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8 / f16",
        "data": "synthetic (data/test images under random homographies); quads from one untimed extraction pass; locally trained weights",
        "config": {"workload": f"configs[2]: warp/crop gather + ResNet-18 forward, 64 squares/board x {B} boards", "chunk": CHUNK,
                   "l2": "inputs (B x 786 KB) larger than L2, no flush"},
        "squares_per_s": 64 * B / (ms / 1000.0), "found_rate": float(found.float().mean().item()),
        "e2e": {"value": B / (e2e_ms / 1000.0), "unit": "boards/s", "h2d_bytes_per_step": B * (H2D_PER_BOARD + 33),
                "d2h_bytes_per_step": B * (64 * 13 * 4 + 128 + 144), "ms_per_step": e2e_ms},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": "k_warp_board (+ k_homography): perspective warp + gray + flip + 64-square crop", "achieved": warp_gbs,
                     "peak": peak_hbm, "unit": "GB/s", "frac": warp_gbs / peak_hbm if warp_gbs and peak_hbm else None, "traffic": traffic,
                     "traffic_source": traffic_src, "peak_source": peak_src, "algorithmic_bytes_per_board": WARP_BYTES_PER_BOARD, "ms_per_step": warp_ms},
        "roofline_classifier": {"bound": "tensor", "kernel": "ResNet-18(1->13) forward over 64 squares/board", "achieved": CLS_GFLOP * B / cls_ms if cls_ms else None,
                                "peak": peak_tf, "unit": "TFLOP/s", "frac": CLS_GFLOP * B / cls_ms / peak_tf if cls_ms and peak_tf else None, "ms_per_step": cls_ms},
        "stage_ms_per_step": {k: v / args.steps for k, v in stages.items()}, "cpu_baseline": cpu, "clocks": clk.summary()}))
    eng.close()
    if world > 1:
        torch.distributed.destroy_process_group()


# =====================================================================================================================
# configs[3]: the whole pipeline (the driver's default)
# =====================================================================================================================
def run_pipeline(args):
    import torch
    import torch.distributed as dist
    use_product_package()
    from chessvision import ChessVision, _native, utils

    world, rank, local_rank, dev, barrier, max_over_ranks = _dist_env()
    chunk = args.chunk
    # boards per GPU per step: whole chunks, and `steps` steps on ONE GPU cover configs[3]'s 65,536 boards
    B = args.boards or chunk * max(1, -(-CONFIG3_BOARDS // (max(args.steps, 1) * chunk)))
    B = min(B, 32 * chunk)

    eng = _native.Engine(local_rank, max_batch=chunk)
    wdir = ROOT / "weights"
    ext_w, cls_w = str(wdir / "best_extractor.pth"), str(wdir / "best_classifier.pth")
    eng.load_unet(utils.load_state_dict(ext_w)[0])
    eng.load_resnet18(utils.load_state_dict(cls_w)[0])

    # ---- synthetic input: B DISTINCT boards per GPU, every rank its own shard of the stream (its own seed)
    host_img = torch.from_numpy(synthetic_boards(B, SEED + 1000 * rank)).pin_memory()
    dev_img = host_img.to(dev)
    out_dev = eng.alloc_outputs(B)
    out_host = eng.alloc_outputs(B, pinned_host=True)
    stream = torch.cuda.current_stream()

    with ClockSampler(local_rank) as clk:
        # ---- device-resident arm
        for _ in range(args.warmup):
            eng.image_to_fen(dev_img, out_dev)
        barrier()
        eng.profile(True)
        l0 = eng.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        for _ in range(args.steps):
            eng.image_to_fen(dev_img, out_dev)
        e1.record(stream)
        barrier()
        dev_ms = max_over_ranks(e0.elapsed_time(e1))
        launches = eng.launch_count() - l0
        stages = eng.profile_read()
        eng.profile(False)
        found_rate = float(out_dev["found"].float().mean().item())
        clocks = clk.summary()

        # ---- end-to-end arm: pinned host input -> H2D -> pipeline -> D2H of the results, through the C-ABI host entry point
        for _ in range(args.warmup):
            eng.image_to_fen_host(host_img, out_host)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            eng.image_to_fen_host(host_img, out_host)   # synchronous on return
        barrier()
        e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1000.0)
    same = all(torch.equal(out_dev[k].cpu(), out_host[k]) for k in out_dev)

    # ---- drop-in API arm: ChessVision.process_images on the same pinned host batch, every output process_image returns
    # (logits, mask, quadrangle, board image, probabilities, squares, both FENs, validation fixes), as Python result objects
    api_steps = max(1, min(args.steps, args.api_steps))
    cvm = ChessVision.from_engine(eng)   # the same loaded context (no second copy of the workspaces)
    api_batch = host_img.numpy()
    res = None
    for _ in range(2):
        del res   # a consumer handles one batch of results and drops it before asking for the next
        res = cvm.process_images(api_batch)
    barrier()
    t0 = time.perf_counter()
    for _ in range(api_steps):
        del res
        res = cvm.process_images(api_batch)
    barrier()
    api_ms = max_over_ranks((time.perf_counter() - t0) * 1000.0)
    host_fens = _native.fen_strings(out_host["fen"])
    api_same = all((r.position is not None) == bool(out_host["found"][i]) and (r.position is None or r.position.fen == host_fens[i][1])
                   for i, r in enumerate(res))
    del res

    boards_total = B * args.steps * world
    value = boards_total / (dev_ms / 1000.0)
    e2e_value = boards_total / (e2e_ms / 1000.0)
    api_value = B * api_steps * world / (api_ms / 1000.0)

    # ---- roofline of the dominant kernel class: the tcgen05 implicit-GEMM convs of the UNet (20 launches per chunk)
    peak_tf, peak_hbm, peak_src = measured_peaks()
    unet_tc_ms = stages["unet_conv_tc"]
    tc_flops = (UNET_GFLOP - UNET_STEM_GFLOP) * 1e9 * B * args.steps
    achieved = tc_flops / (unet_tc_ms / 1000.0) / 1e12 if unet_tc_ms > 0 else 0.0
    chunks = (B + chunk - 1) // chunk
    traffic, traffic_src = ncu_traffic("unet_conv_tc", chunk)
    roofline = {"bound": "tensor", "kernel": "conv_tc_kernel / conv3x3_vr_kernel / conv3x3_rs_kernel (UNet layers, tcgen05 implicit GEMM)",
                "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf if peak_tf else None,
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "launches": 20 * chunks * args.steps, "avg_launch_ms": unet_tc_ms / max(1, 20 * chunks * args.steps),   # 20 conv launches per chunk (up3.conv3 + up4.up are one)
                "algorithmic_gflop_per_board": UNET_GFLOP - UNET_STEM_GFLOP, "algorithmic_bytes_per_launch": 868.0e6 * chunk / 128}
    stage_share = {k: v / max(1e-9, sum(stages.values())) for k, v in stages.items()}
    cls_ms = stages["resnet_conv_tc"]
    # second named metric of BASELINE.json: warp + 64-square crop against the measured HBM copy bandwidth
    warp_ms = stages["warp"]
    warp_gbs = WARP_BYTES_PER_BOARD * B * args.steps / (warp_ms / 1000.0) / 1e9 if warp_ms > 0 else 0.0
    groups = (B + 8 * chunk - 1) // (8 * chunk)
    wtraffic, wtraffic_src = ncu_traffic("warp_board", (B + groups - 1) // groups)   # one warp launch per group of 8 chunks
    roofline_warp = {"bound": "hbm", "kernel": "k_warp_board (+ k_homography)", "achieved": warp_gbs, "peak": peak_hbm, "unit": "GB/s",
                     "frac": warp_gbs / peak_hbm if peak_hbm else None, "traffic": wtraffic, "traffic_source": wtraffic_src,
                     "launches": 2 * groups * args.steps, "avg_launch_ms": warp_ms / max(1, groups * args.steps),
                     "algorithmic_bytes_per_board": WARP_BYTES_PER_BOARD}

    if rank == 0:
        cpu = None if args.no_cpu_baseline else cpu_baseline_subprocess("pipeline", args.cpu_boards)
        line = {
            "metric": "boards/sec image->FEN", "value": value, "unit": "boards/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16",
            "data": f"synthetic: {B} distinct boards per GPU (data/test images under random homographies + colour jitter, seed {SEED} + 1000*rank); "
                    "locally trained weights",
            "config": {"workload": f"configs[3]: full image->FEN pipeline, synthetic 512x512x3 boards sharded by batch; {B} boards/GPU/step x {args.steps} steps x "
                                   f"{world} GPU = {boards_total} boards in the timed region (configs[3] names 65,536)",
                       "boards_per_gpu_per_step": B, "chunk": chunk, "l2": "inputs (B x 786 KB) larger than L2, no flush",
                       "parallelism": f"batch-sharded x{world}, no collective"},
            "e2e": {"value": e2e_value, "unit": "boards/s", "h2d_bytes_per_step": B * H2D_PER_BOARD, "d2h_bytes_per_step": B * D2H_PER_BOARD,
                    "ms_per_step": e2e_ms / args.steps, "results_equal_device_arm": bool(same),
                    "call": "cvb_image_to_fen_host (C ABI, pinned host buffers; outputs: quad, found, status, probabilities, labels, FEN)"},
            "e2e_api": {"value": api_value, "unit": "boards/s", "h2d_bytes_per_step": B * H2D_PER_BOARD, "d2h_bytes_per_step": B * D2H_FULL_PER_BOARD,
                        "ms_per_step": api_ms / api_steps, "steps": api_steps, "fen_equal_e2e_arm": bool(api_same),
                        "call": "ChessVision.process_images (drop-in Python API: + logits, mask, board image per board, result dataclasses)"},
            "gpu_launches": int(launches), "roofline": roofline, "roofline_warp_crop": roofline_warp,
            "classifier_tflops": CLS_GFLOP * B * args.steps / cls_ms if cls_ms > 0 else None,
            "cpu_baseline": cpu, "clocks": clocks,
            "found_rate": found_rate, "stage_ms_per_step": {k: v / args.steps for k, v in stages.items()}, "stage_share": stage_share,
            "gflop_per_board": UNET_GFLOP + CLS_GFLOP,
            "model_tflops": value * (UNET_GFLOP + CLS_GFLOP) / 1000.0 / world,
        }
        emit(json.dumps(line))
    cvm._engine_obj = None
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="pipeline", choices=["pipeline", "unet-sweep", "classify", "train", "train-classifier", "decode"],
                    help="pipeline = BASELINE.json's metric (configs[3]); unet-sweep = configs[1]; classify = configs[2]; train = configs[4]; "
                         "decode = JPEG front-end")
    ap.add_argument("--train-batch", type=int, default=8, help="--workload train: images per GPU per step")
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--boards", type=int, default=0, help="boards per GPU per step (0: whole chunks so that `steps` steps cover 65,536 boards)")
    ap.add_argument("--chunk", type=int, default=PIPE_CHUNK, help="boards per pipeline chunk (workspace size)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-boards", type=int, default=64, help="boards of the bounded CPU-baseline sample")
    ap.add_argument("--cpu-boards-per-step", type=int, default=4, help="--impl reference: boards per step")
    ap.add_argument("--api-steps", type=int, default=4, help="steps of the e2e_api arm (ChessVision.process_images)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    capture_stdout()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        if args.workload == "train":
            return run_train_reference(args)
        if args.workload == "train-classifier":
            if int(os.environ.get("RANK", "0")) == 0:
                emit(json.dumps({"impl": "reference", "unavailable": "the train-classifier workload reports the fp32 PyTorch loop as its cpu_baseline in the b200 arm"}))
            return None
        if args.workload == "decode":
            if int(os.environ.get("RANK", "0")) == 0:
                emit(json.dumps({"impl": "reference", "unavailable": "the decode workload reports cv2.imdecode as its cpu_baseline in the b200 arm"}))
            return None
        return run_reference(args)
    return {"decode": run_decode, "train": run_train, "train-classifier": run_train_classifier, "unet-sweep": run_unet_sweep, "classify": run_classify,
            "pipeline": run_pipeline}[args.workload](args)


if __name__ == "__main__":
    main()

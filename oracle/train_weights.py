"""ORACLE tooling: train small local weights for the two networks from the data that ships in the reference tree.

The reference ships no weights (README.md:39) and random-init UNets produce empty masks, so parity on
``data/test`` needs locally trained weights (SURVEY.md §7 hard part 3).  This script is plain PyTorch on CPU, uses
the *inference-time* preprocessing of the reference (BGR/255 for the UNet, core.py:215; gray/255 for the classifier,
core.py:236-237) and writes checkpoints in the reference's own layout
``{"model_state_dict": ..., "metadata": ...}`` (scripts/train/train_unet.py:31-40), stored as fp16 tensors so that the
committed files stay small and the fp32 oracle and the fp16 CUDA path start from bit-identical parameters.

Run here (needs /root/reference):  python oracle/train_weights.py --unet-epochs 5 --cls-epochs 4
"""
from __future__ import annotations

import argparse
import glob
import os
import sys
import time

import cv2
import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.nets import BoardUNet, PieceResNet18  # noqa: E402

REF = os.environ.get("CV_REFERENCE", "/root/reference")
LABELS = ["B", "K", "N", "P", "Q", "R", "b", "k", "n", "p", "q", "r", "f"]  # constants.py:23
CLASS_DIRS = ["B", "K", "N", "P", "Q", "R", "_b", "_k", "_n", "_p", "_q", "_r", "f"]


def save(model, path, meta):
    sd = {k: (v.half() if v.is_floating_point() else v) for k, v in model.state_dict().items()}
    torch.save({"model_state_dict": sd, "metadata": meta}, path)


def round_to_fp16_(model):
    with torch.no_grad():
        for p in list(model.parameters()) + list(model.buffers()):
            if p.is_floating_point():
                p.copy_(p.half().float())


def train_classifier(out, epochs, seed):
    torch.manual_seed(seed)
    rng = np.random.default_rng(seed)

    def load(split):
        xs, ys = [], []
        for ci, d in enumerate(CLASS_DIRS):
            for f in sorted(glob.glob(f"{REF}/data/squares/{split}/{d}/*")):
                im = cv2.imread(f, cv2.IMREAD_GRAYSCALE)
                if im is None:
                    continue
                if im.shape != (64, 64):
                    im = cv2.resize(im, (64, 64))
                xs.append(im)
                ys.append(ci)
        return np.stack(xs), np.array(ys)

    xtr, ytr = load("training")
    xva, yva = load("validation")
    print(f"classifier data: train {xtr.shape} val {xva.shape}", flush=True)
    net = PieceResNet18()
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)
    sched = torch.optim.lr_scheduler.StepLR(opt, step_size=max(1, epochs - 1), gamma=0.1)
    bs = 64
    for ep in range(epochs):
        net.train()
        perm = rng.permutation(len(xtr))
        t0 = time.time()
        tot = 0.0
        for i in range(0, len(perm) - bs + 1, bs):
            idx = perm[i:i + bs]
            xb = xtr[idx].copy()
            # small-shift augmentation (+-3 px) so that +-1 px corner moves cannot flip labels
            for j in range(len(xb)):
                dy, dx = rng.integers(-3, 4, 2)
                xb[j] = np.roll(xb[j], (dy, dx), (0, 1))
            x = torch.from_numpy(xb).float().div_(255.0).unsqueeze(1)
            gain = torch.empty(len(idx), 1, 1, 1).uniform_(0.85, 1.15)
            x = (x * gain).clamp_(0, 1)
            loss = F.cross_entropy(net(x), torch.from_numpy(ytr[idx]))
            opt.zero_grad(set_to_none=True)
            loss.backward()
            opt.step()
            tot += loss.item()
        sched.step()
        net.eval()
        with torch.no_grad():
            pred = torch.cat([net(torch.from_numpy(xva[i:i + 512]).float().div(255).unsqueeze(1)).argmax(1)
                              for i in range(0, len(xva), 512)])
        acc = (pred.numpy() == yva).mean()
        print(f"classifier epoch {ep} loss {tot / (len(perm) // bs):.4f} val_acc {acc:.4f} {time.time() - t0:.0f}s", flush=True)
        save(net, out, {"arch": "resnet18", "epochs": ep + 1, "val_acc": float(acc), "seed": seed})


def train_unet(out, epochs, seed, resume=None):
    torch.manual_seed(seed)
    rng = np.random.default_rng(seed)
    imgs, masks = [], []
    for f in sorted(glob.glob(f"{REF}/data/board_extraction/images/*")):
        m = f.replace("/images/", "/masks/").rsplit(".", 1)[0] + ".png"
        im, mk = cv2.imread(f), cv2.imread(m, cv2.IMREAD_GRAYSCALE)
        if im is None or mk is None:
            continue
        imgs.append(im)
        masks.append((mk > 127).astype(np.float32))
    imgs, masks = np.stack(imgs), np.stack(masks)
    print(f"unet data: {imgs.shape} {masks.shape}", flush=True)
    net = BoardUNet().to(memory_format=torch.channels_last)
    if resume and os.path.exists(resume):
        net.load_state_dict({k: v.float() for k, v in torch.load(resume)["model_state_dict"].items()})
        print("resumed from", resume, flush=True)
    opt = torch.optim.Adam(net.parameters(), lr=2e-4)
    bs = 4
    steps_total = epochs * (len(imgs) // bs)
    sched = torch.optim.lr_scheduler.OneCycleLR(opt, max_lr=3e-4, total_steps=steps_total, pct_start=0.15)
    step = 0
    for ep in range(epochs):
        net.train()
        perm = rng.permutation(len(imgs))
        t0 = time.time()
        tot = 0.0
        for i in range(0, len(perm) - bs + 1, bs):
            idx = perm[i:i + bs]
            x = torch.from_numpy(imgs[idx]).float().div_(255.0).permute(0, 3, 1, 2).contiguous(memory_format=torch.channels_last)
            y = torch.from_numpy(masks[idx]).unsqueeze(1)
            if rng.random() < 0.5:
                x, y = x.flip(3), y.flip(3)
            gain = torch.empty(bs, 3, 1, 1).uniform_(0.85, 1.15)
            x = (x * gain).clamp_(0, 1)
            logit = net(x)
            p = torch.sigmoid(logit)
            inter = (p * y).sum((1, 2, 3))
            dice = (2 * inter + 1e-6) / (p.sum((1, 2, 3)) + y.sum((1, 2, 3)) + 1e-6)  # dice_score.py:5-19
            loss = F.binary_cross_entropy_with_logits(logit, y) + (1 - dice.mean())
            opt.zero_grad(set_to_none=True)
            loss.backward()
            torch.nn.utils.clip_grad_norm_(net.parameters(), 1.0)
            opt.step()
            sched.step()
            step += 1
            tot += loss.item()
            if step % 20 == 0:
                print(f"  unet step {step}/{steps_total} loss {tot / (i // bs + 1):.4f} {time.time() - t0:.0f}s", flush=True)
        print(f"unet epoch {ep} loss {tot / (len(perm) // bs):.4f} {time.time() - t0:.0f}s", flush=True)
        save(net, out, {"arch": "unet", "epochs": ep + 1, "loss": tot / (len(perm) // bs), "seed": seed})


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--out-dir", default=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "weights"))
    ap.add_argument("--unet-epochs", type=int, default=5)
    ap.add_argument("--cls-epochs", type=int, default=4)
    ap.add_argument("--seed", type=int, default=20261017)
    ap.add_argument("--threads", type=int, default=os.cpu_count())
    ap.add_argument("--resume", action="store_true")
    a = ap.parse_args()
    torch.set_num_threads(a.threads)
    os.makedirs(a.out_dir, exist_ok=True)
    if a.cls_epochs:
        train_classifier(os.path.join(a.out_dir, "best_classifier.pth"), a.cls_epochs, a.seed)
    if a.unet_epochs:
        p = os.path.join(a.out_dir, "best_extractor.pth")
        train_unet(p, a.unet_epochs, a.seed, resume=p if a.resume else None)

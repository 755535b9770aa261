"""ORACLE tooling: run the UNMODIFIED reference (``/root/reference/chessvision``) on its own ``data/test`` images and
freeze the outputs as golden vectors under ``tests/golden/``.

Runs only in the build container (needs /root/reference).  The reference imports two packages that are not installed
here; both are replaced by minimal stand-ins registered in ``sys.modules`` before the import:

* ``chess``  (python-chess 1.11.2): only ``SQUARE_NAMES``, ``Piece.from_symbol``, ``BaseBoard.set_piece_at/board_fen``
  are touched (core.py:330-349);
* ``timm``   (1.0.15): only ``create_model("resnet18", num_classes=13, in_chans=1)`` (utils.py:35-39), provided by
  torchvision's resnet18 with a 1-channel conv1 (same topology and state-dict keys).

While generating, every stage of ``oracle/pipeline.py`` is compared with the reference (bit-exact for the integer
stages and — same ATen CPU kernels — for the networks); a mismatch aborts.

    python oracle/make_golden.py            # writes tests/golden/{data_test/, reference_outputs.npz, manifest.json}
"""
from __future__ import annotations

import glob
import hashlib
import json
import os
import shutil
import sys
import types

import cv2
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.environ.get("CV_REFERENCE", "/root/reference")
GOLD = os.path.join(ROOT, "tests", "golden")


from oracle.ref_loader import install_standins, load_reference  # noqa: E402,F401  (stand-ins for `chess` / `timm`, see there)


def sha(a: np.ndarray) -> str:
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    torch.set_num_threads(os.cpu_count())
    wdir = os.path.join(ROOT, "weights")
    ext_w, cls_w = os.path.join(wdir, "best_extractor.pth"), os.path.join(wdir, "best_classifier.pth")
    ref = load_reference()
    ref.core.utils.get_device = lambda: torch.device("cpu")
    cv = ref.ChessVision(board_extractor_weights=ext_w, classifier_weights=cls_w, classifier_model_id="resnet18", lazy_load=False)
    from oracle.pipeline import OraclePipeline
    oracle = OraclePipeline.from_checkpoints(ext_w, cls_w)

    os.makedirs(os.path.join(GOLD, "data_test"), exist_ok=True)
    files = sorted(glob.glob(f"{REF}/data/test/*/raw/*"))
    manifest = {"weights": {"extractor_sha1": hashlib.sha1(open(ext_w, "rb").read()).hexdigest(),
                            "classifier_sha1": hashlib.sha1(open(cls_w, "rb").read()).hexdigest()},
                "cv2": cv2.__version__, "torch": torch.__version__, "images": []}
    arrays = {}
    n_found = 0
    sq_ok = sq_tot = 0
    for idx, f in enumerate(files):
        subset = f.split("/")[-3]
        name = os.path.basename(f)
        dst_dir = os.path.join(GOLD, "data_test", subset)
        os.makedirs(dst_dir, exist_ok=True)
        shutil.copyfile(f, os.path.join(dst_dir, name))
        gt_path = f.replace("/raw/", "/ground_truth/").rsplit(".", 1)[0] + ".txt"
        gt = open(gt_path).read().strip() if os.path.exists(gt_path) else None
        img = cv2.imread(f)
        res = cv.process_image(img)
        be = res.board_extraction
        o = oracle.process_image(img)
        # ---- oracle vs reference, stage by stage
        assert np.array_equal(be.probabilities, o["logits"]), f"{name}: UNet logits differ ({np.abs(be.probabilities - o['logits']).max()})"
        assert np.array_equal(be.binary_mask, o["mask"]), f"{name}: mask differs"
        assert (be.quadrangle is None) == (o["quad"] is None), f"{name}: found flag differs"
        entry = {"file": f"{subset}/{name}", "image_sha1": sha(img), "ground_truth_fen": gt, "found": be.quadrangle is not None,
                 "logits_absmax": float(np.abs(be.probabilities).max())}
        arrays[f"mask_{idx}"] = np.packbits(be.binary_mask > 0)
        arrays[f"logits_{idx}"] = be.probabilities.astype(np.float16)
        if be.quadrangle is not None:
            n_found += 1
            quad256 = (be.quadrangle / 2.0).astype(np.int32).reshape(4, 2)
            assert np.array_equal(quad256, o["quad"].reshape(4, 2)), f"{name}: quad differs"
            assert np.array_equal(be.board_image, o["board"]), f"{name}: board differs in {(be.board_image != o['board']).sum()} bytes"
            pr = res.position
            assert np.array_equal(pr.model_probabilities, o["probs"]), f"{name}: probabilities differ ({np.abs(pr.model_probabilities - o['probs']).max()})"
            assert pr.fen == o["fen"] and pr.original_fen == o["original_fen"], f"{name}: FEN differs"
            labels = np.argmax(pr.model_probabilities, axis=1).astype(np.uint8)
            entry.update(quad=quad256.tolist(), fen=pr.fen, original_fen=pr.original_fen, board_sha1=sha(be.board_image),
                         n_fixes=len(pr.validation_fixes))
            arrays[f"board_{idx}"] = be.board_image
            arrays[f"probs_{idx}"] = pr.model_probabilities
            arrays[f"labels_{idx}"] = labels
            if gt:
                want = fen_to_labels(gt)
                sq_tot += 64
                sq_ok += sum(a == b for a, b in zip(want, o["labels_valid"]))
        manifest["images"].append(entry)
        print(idx, name, "found" if entry["found"] else "NOT FOUND", entry.get("fen"), flush=True)
    manifest["summary"] = {"images": len(files), "found": n_found, "square_accuracy_vs_ground_truth": sq_ok / max(sq_tot, 1)}
    np.savez_compressed(os.path.join(GOLD, "reference_outputs.npz"), **arrays)
    json.dump(manifest, open(os.path.join(GOLD, "manifest.json"), "w"), indent=1)
    print(json.dumps(manifest["summary"]))


def fen_to_labels(fen: str):
    out = []
    for row in fen.split(" ")[0].split("/"):
        for ch in row:
            out += ["f"] * int(ch) if ch.isdigit() else [ch]
    return out


if __name__ == "__main__":
    main()

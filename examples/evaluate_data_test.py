#!/usr/bin/env python
"""The reference's de-facto integration test — ``scripts/eval/evaluate.py`` over ``data/test`` against the ground-truth FEN
files (scripts/bin/evaluate.sh:6-16, evaluate.py:143-152,227-330) — with every stage on the GPU: JPEG files -> pixels
(cvb_decode_jpeg) -> image->FEN (cvb_image_to_fen) -> top-k / position accuracy (cvb_eval_metrics).  No 3LC run is
written; the aggregate numbers evaluate.py logs are printed as one JSON line.

    python examples/evaluate_data_test.py [--image-folder tests/golden/data_test] [--threshold 0.5]
"""
import argparse
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT / "chessvision-3lc_b200")]

import torch  # noqa: E402
from chessvision import ChessVision, decode, evaluation  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--image-folder", default=str(ROOT / "tests" / "golden" / "data_test"))
    ap.add_argument("--manifest", default=str(ROOT / "tests" / "golden" / "manifest.json"), help="ground-truth FEN per file")
    ap.add_argument("--threshold", type=float, default=0.5)
    ap.add_argument("--board-extractor-weights", default=str(ROOT / "weights" / "best_extractor.pth"))
    ap.add_argument("--classifier-weights", default=str(ROOT / "weights" / "best_classifier.pth"))
    a = ap.parse_args()
    truth = {e["file"]: e["ground_truth_fen"] for e in json.load(open(a.manifest))["images"]}
    files = sorted(p for p in Path(a.image_folder).glob("*/*") if f"{p.parent.name}/{p.name}" in truth)
    fens = [truth[f"{p.parent.name}/{p.name}"].split()[0] for p in files]
    cv = ChessVision(board_extractor_weights=a.board_extractor_weights, classifier_weights=a.classifier_weights,
                     classifier_model_id="resnet18", lazy_load=False, max_batch=64)
    eng = cv._engine
    t0 = time.time()
    imgs = decode.imread_batch(files, engine=eng)
    out = eng.image_to_fen(imgs, eng.alloc_outputs(len(files)), a.threshold)
    hits, correct = evaluation.evaluate_batch(out["probs"], out["labels"], out["labels_valid"], fens, k=3, engine=eng)
    found = out["found"].cpu().bool()
    torch.cuda.synchronize()
    dt = time.time() - t0
    n = int(found.sum())
    h, c = hits[found].double(), correct[found].double()
    print(json.dumps({
        "images": len(files), "boards_found": n, "extraction_failures": len(files) - n,
        "top_1_accuracy": float(h[:, 0].mean() / 64), "top_2_accuracy": float(h[:, 1].mean() / 64), "top_3_accuracy": float(h[:, 2].mean() / 64),
        "original_position_accuracy": float(c[:, 0].mean() / 64), "validated_position_accuracy": float(c[:, 1].mean() / 64),
        "positions_fully_correct": int((correct[found][:, 1] == 64).sum()), "seconds": round(dt, 3)}))


if __name__ == "__main__":
    main()

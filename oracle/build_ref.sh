#!/usr/bin/env bash
# ORACLE tooling (test / measurement infrastructure, never product code).
# Stage the UNMODIFIED reference implementation of the image->FEN path -- the six files SURVEY.md §8(a) cites plus the
# vendored UNet -- from the read-only checkout into oracle/_ref/ (git-ignored, but shipped to the GPU box with the
# snapshot), so that `bench.py --impl reference`, bench.py's cpu_baseline leg and tests/ can run
# ChessVision.process_image (chessvision/core.py:152-195) itself where /root/reference does not exist.
# Nothing is edited: the files are copied byte for byte and their sha256 is recorded in oracle/_ref/MANIFEST.
set -euo pipefail
REF=${CV_REFERENCE:-/root/reference}
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
[ -d "$REF/chessvision" ] || { echo "build_ref: $REF/chessvision not found (nothing staged)"; exit 0; }
rm -rf "$OUT"
mkdir -p "$OUT/chessvision/pytorch_unet/unet"
for f in __init__.py core.py utils.py constants.py cv_types.py; do cp "$REF/chessvision/$f" "$OUT/chessvision/$f"; done
for f in __init__.py unet_model.py unet_parts.py; do cp "$REF/chessvision/pytorch_unet/unet/$f" "$OUT/chessvision/pytorch_unet/unet/$f"; done
[ -f "$REF/chessvision/pytorch_unet/__init__.py" ] && cp "$REF/chessvision/pytorch_unet/__init__.py" "$OUT/chessvision/pytorch_unet/__init__.py"
( cd "$OUT" && find chessvision -type f -name '*.py' | sort | xargs sha256sum > MANIFEST )
( cd "$REF" && find chessvision -maxdepth 1 -type f -name '*.py' | sort | xargs sha256sum; \
  cd "$REF" && find chessvision/pytorch_unet/unet -type f -name '*.py' | sort | xargs sha256sum ) | sort -k2 > "$OUT/MANIFEST.src"
sort -k2 "$OUT/MANIFEST" | diff -q - "$OUT/MANIFEST.src" > /dev/null || { echo "build_ref: staged copy differs from the checkout"; exit 1; }
echo "staged $(wc -l < "$OUT/MANIFEST") reference files into $OUT"

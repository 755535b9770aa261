"""The C-ABI shared library loads on a machine without a GPU and exports every symbol that include/chessvision_b200.h
declares; the ctypes table in chessvision/_native.py covers exactly those symbols.  No compute call is made here."""
import ctypes
import re
import subprocess

import pytest

from conftest import PKG, ROOT

HEADER = ROOT / "include" / "chessvision_b200.h"
LIB = PKG / "libchessvision_b200.so"


def declared_symbols():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"CVB_API\s+[\w\s\*]+?\b(cvb_\w+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    if not LIB.exists():
        subprocess.run(["bash", str(PKG / "build.sh")], check=True)
    return ctypes.CDLL(str(LIB))


def test_header_declares_the_expected_surface():
    names = declared_symbols()
    assert len(names) >= 20
    for must in ("cvb_create", "cvb_destroy", "cvb_load_unet", "cvb_load_resnet18", "cvb_unet_forward", "cvb_mask_to_quad",
                 "cvb_warp_squares", "cvb_classify", "cvb_image_to_fen", "cvb_image_to_fen_host"):
        assert must in names


def test_every_declared_symbol_is_exported(lib):
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} is declared in {HEADER.name} but not exported by {LIB.name}"


def test_ctypes_table_matches_header(lib):
    from chessvision import _native
    assert sorted(_native.SYMBOLS) == declared_symbols()
    _native.load_library()


def test_nothing_but_the_c_abi_is_exported():
    out = subprocess.run(["nm", "-D", "--defined-only", str(LIB)], capture_output=True, text=True, check=True).stdout
    exported = [line.split()[-1] for line in out.splitlines() if " T " in line]
    extra = [s for s in exported if not s.startswith("cvb_") and s not in ("_init", "_fini")]
    assert not extra, f"unexpected exported text symbols: {extra[:8]}"


def test_version_and_null_safety(lib):
    lib.cvb_version.restype = ctypes.c_int
    assert lib.cvb_version() >= 100
    lib.cvb_last_error.restype = ctypes.c_char_p
    lib.cvb_last_error.argtypes = [ctypes.c_void_p]
    assert lib.cvb_last_error(None) == b"null context"
    lib.cvb_max_batch.argtypes = [ctypes.c_void_p]
    assert lib.cvb_max_batch(None) == 0


def test_product_path_fails_loudly_without_gpu():
    """No CPU fallback: constructing the engine (and therefore every ChessVision pipeline call) raises without CUDA."""
    import numpy as np
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible here")
    from chessvision import ChessVision, _native
    with pytest.raises(_native.NativeError):
        _native.Engine(0, max_batch=1)
    cvm = ChessVision(board_extractor_weights="/nonexistent/a.pth", classifier_weights="/nonexistent/b.pth")
    with pytest.raises((_native.NativeError, FileNotFoundError, AssertionError, RuntimeError)):
        cvm.process_image(np.zeros((512, 512, 3), np.uint8))
    with pytest.raises(_native.NativeError):
        ChessVision._find_quadrangle(np.zeros((256, 256), np.uint8))
    from chessvision.training import ClassifierTrainer, UNetTrainer
    with pytest.raises(_native.NativeError):
        ClassifierTrainer({}, batch_size=4)
    with pytest.raises(_native.NativeError):
        UNetTrainer({}, batch_size=1)


def test_product_package_never_imports_the_oracle():
    for f in (PKG / "chessvision").glob("*.py"):
        text = f.read_text()
        assert "import oracle" not in text and "from oracle" not in text, f"{f.name} imports the oracle"
    for f in (PKG / "csrc").iterdir():
        assert "oracle" not in f.read_text(), f"{f.name} mentions the oracle"


def test_jpeg_header_rejects_oversubscribed_huffman_tables(lib):
    """A DHT segment whose code-length counts do not form a prefix code (255 codes of length 1, ...) used to index far
    past the 256-entry lookahead table; it must come back as -6 (unsupported / corrupt), for every table class and for
    random count vectors, and the host-only entry points must survive it (no GPU needed: parsing is host code)."""
    import numpy as np
    lib.cvb_jpeg_info.restype = ctypes.c_int
    lib.cvb_jpeg_info.argtypes = [ctypes.c_char_p, ctypes.c_int64, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32)]
    lib.cvb_jpeg_coefficients.restype = ctypes.c_int
    lib.cvb_jpeg_coefficients.argtypes = [ctypes.c_char_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p]

    def stream(counts, tc_th=0x00):
        n = int(sum(counts))
        seg = bytes([tc_th]) + bytes(counts) + bytes(range(256))[:n]
        return b"\xff\xd8\xff\xc4" + (len(seg) + 2).to_bytes(2, "big") + seg + b"\xff\xd9"

    h, w = ctypes.c_int32(), ctypes.c_int32()
    bad = [[255] + [0] * 15, [0, 5] + [0] * 14, [1, 1, 3] + [0] * 13, [0] * 7 + [255, 1] + [0] * 7, [2, 1] + [0] * 14]
    for counts in bad:
        for cls in (0x00, 0x10, 0x13):
            buf = stream(counts, cls)
            assert lib.cvb_jpeg_info(buf, len(buf), ctypes.byref(h), ctypes.byref(w)) == -6, counts
    rng = np.random.default_rng(5)
    scratch = (ctypes.c_int16 * (512 * 512 * 3 // 2))()
    for _ in range(400):
        counts = rng.integers(0, 20, 16).tolist()
        if sum(counts) > 256:
            continue
        buf = stream(counts, int(rng.integers(0, 2)) << 4 | int(rng.integers(0, 4)))
        assert lib.cvb_jpeg_info(buf, len(buf), ctypes.byref(h), ctypes.byref(w)) == -6   # no scan either way: never 0
        assert lib.cvb_jpeg_coefficients(buf, len(buf), scratch, None) == -6

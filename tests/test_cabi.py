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


def test_product_package_never_imports_the_oracle():
    for f in (PKG / "chessvision").glob("*.py"):
        text = f.read_text()
        assert "import oracle" not in text and "from oracle" not in text, f"{f.name} imports the oracle"
    for f in (PKG / "csrc").iterdir():
        assert "oracle" not in f.read_text(), f"{f.name} mentions the oracle"

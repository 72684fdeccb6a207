"""The C-ABI library loads on a CPU-only box and exports every symbol include/machisplin_b200.h declares
(no compute calls without a GPU)."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "machisplin_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mbC?_[A-Za-z0-9_]+)\s*\(", src)))


def test_header_declares_the_hot_path_boundary():
    names = declared_symbols()
    for required in ("mb_tps_fit", "mb_tps_eval", "mb_ensemble_eval", "mb_tiles_tps", "mb_tiles_merge", "mb_gram",
                     "mb_init", "mb_shutdown", "mb_device_count", "mb_last_error"):
        assert required in names


def test_library_exports_every_declared_symbol():
    from machisplin_b200 import _lib
    lib = _lib.load()
    names = declared_symbols()
    assert set(names) == set(_lib.SIGNATURES), set(names) ^ set(_lib.SIGNATURES)
    out = subprocess.run(["nm", "-D", "--defined-only", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (mbC?_[A-Za-z0-9_]+)", out))
    assert set(names) <= exported, set(names) - exported
    assert lib.mb_version() == 100


def test_no_gpu_means_loud_failure_not_fallback():
    from machisplin_b200 import _lib
    import machisplin_b200 as mb
    if _lib.load().mb_device_count() > 0:
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        mb.Engine(0)


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "machisplin_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert "oracle/" not in txt, f

"""CPU: the C-ABI library builds, loads, and exports exactly the symbols include/u2b200.h declares;
the ctypes table in u2seg_b200/_lib.py covers all of them. No compute calls (no GPU here)."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "u2b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(u2b_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_loads():
    from u2seg_b200.build import build_library
    path = build_library()
    assert os.path.exists(path)
    from u2seg_b200 import _lib
    L = _lib.lib()
    assert L.u2b_version() == 100
    assert L.u2b_last_error() is not None


def test_every_declared_symbol_is_exported_and_bound():
    from u2seg_b200 import _lib
    from u2seg_b200.build import build_library
    declared = _header_symbols()
    assert len(declared) >= 20
    out = subprocess.check_output(["nm", "-D", "--defined-only", build_library()], text=True)
    exported = set(re.findall(r"\b(u2b_[a-z0-9_]+)\b", out))
    assert set(declared) <= exported, sorted(set(declared) - exported)
    assert set(declared) == set(_lib.SIGNATURES), (sorted(set(declared) - set(_lib.SIGNATURES)),
                                                    sorted(set(_lib.SIGNATURES) - set(declared)))


def test_sass_contains_blackwell_tensor_and_tma_instructions():
    from u2seg_b200.build import build_library
    sass = subprocess.check_output(["cuobjdump", "-sass", build_library()], text=True)
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM"):          # tcgen05.mma / TMA load / tcgen05.ld
        assert mnemonic in sass, mnemonic
    assert "HGMMA" not in sass


def test_product_does_not_import_oracle():
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "u2seg_b200")):
        for f in files:
            if f.endswith(".py") and f != "bench_train.py":      # bench_train's cpu_baseline leg may time the oracle
                txt = open(os.path.join(dirpath, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M):
                    bad.append(f)
    assert not bad, bad


def test_compute_entry_points_fail_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from u2seg_b200.layers import paste_masks_in_image
    with pytest.raises(RuntimeError):
        paste_masks_in_image(torch.rand(2, 28, 28), torch.tensor([[0, 0, 10, 10.0]] * 2), (20, 20))
    from u2seg_b200.clustering import KMeans
    with pytest.raises(RuntimeError):
        KMeans(torch.randn(100, 64), 0, K=4, Niter=1, verbose=False)

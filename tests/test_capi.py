"""The C-ABI library loads on a CPU-only box and exports every symbol include/strolle_b200.h declares;
without a GPU it fails loudly instead of falling back."""
import os
import re

import pytest

import strolle_b200
from strolle_b200 import build as st_build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(strolle_b200.lib_path()):
        st_build.build()
    return strolle_b200.load_library()


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "strolle_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(st_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(lib):
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/strolle_b200.h but not exported"


def test_pass_names(lib):
    names = list(strolle_b200.PASS_NAMES)
    assert names[0] == "prim_gbuffer" and "frame_denoising_wavelet" in names and len(names) == 27


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(strolle_b200.StrolleError, match="no CPU fallback"):
        strolle_b200.Engine()


def test_product_does_not_touch_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py may import, link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "strolle_b200")
    banned = [r"\bimport\s+oracle", r"\bfrom\s+oracle", r"pyoracle", r"liboracle", r"oracle/", r"orc_[a-z_]+\(", r"#include\s+\"orc_"]
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                for pat in banned:
                    assert not re.search(pat, text), f"{f} references the oracle ({pat})"

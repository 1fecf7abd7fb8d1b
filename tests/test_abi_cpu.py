"""The C-ABI library loads without a GPU, exports every symbol include/faster_b200.h declares, and refuses to create
a context when no GPU is present (no CPU fallback)."""
import os
import re

import pytest

from faster_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_exports_match_header(built_lib):
    hdr = open(os.path.join(ROOT, "include", "faster_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(fq_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    L = capi.lib()
    for name in sorted(declared):
        assert hasattr(L, name), "libfaster_b200.so lacks %s" % name
    assert declared == set(capi.EXPORTS)
    assert L.fq_abi_version() == 2


def test_no_gpu_fails_loudly(built_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(capi.FqError) as e:
        capi.Solver(0)
    assert "no CPU fallback" in str(e.value)


def test_product_never_touches_oracle():
    """The product tree must not import, link or call anything under oracle/."""
    for base in ("faster_b200", "include"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                    txt = open(os.path.join(dp, f)).read()
                    assert "fqo_" not in txt and "pyoracle" not in txt and "oracle/" not in txt, os.path.join(dp, f)

"""CPU-only: the C-ABI library loads and exports every symbol include/ef_b200.h declares; host logic."""
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def declared_functions():
    text = (ROOT / "include" / "ef_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ef_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import efb200
    lib = efb200.load_library()
    names = declared_functions()
    assert len(names) >= 20
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert sorted(efb200.EXPORTS) == names
    assert b"sm_100a" in lib.ef_version()


def test_no_cpu_fallback_without_gpu():
    import torch
    import efb200
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(efb200.EfError):
        efb200.EfficientFeatures.create()


def test_product_does_not_touch_the_oracle():
    """the oracle is test infrastructure: nothing under the product package may reference it"""
    pkg = ROOT / "cuda-efficient-features_b200"
    for p in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")) + list(pkg.rglob("*.cpp")) + list(pkg.rglob("*.h")):
        t = p.read_text(errors="ignore")
        assert "efo" not in re.findall(r"import\s+(\w+)", t), p
        assert "libef_oracle" not in t and "oracle/" not in t.replace("TEST", ""), p


def test_shard_range():
    from efb200.sharding import shard_range
    for n in (0, 1, 7, 64, 65):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def test_mg_shard_range_matches_python_sharding():
    """the C driver (ef_mg_shard_range) and the per-process driver (efb200.sharding) cut a batch the same way"""
    import ctypes as C
    import efb200
    from efb200.sharding import shard_range
    lib = efb200.load_library()
    for n in (0, 1, 7, 64, 65):
        for world in (1, 2, 3, 8):
            for r in range(world):
                b, e = C.c_int(), C.c_int()
                lib.ef_mg_shard_range(n, r, world, C.byref(b), C.byref(e))
                assert (b.value, e.value) == shard_range(n, r, world)

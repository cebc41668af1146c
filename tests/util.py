"""Helpers shared by the parity tests."""
import numpy as np

SEED = 0xEFB20000


def canon_keypoints(k):
    """Sort a keypoint structured array by (octave, y, x) -> canonical order for set comparison."""
    order = np.lexsort((k["x"], k["y"], k["octave"]))
    return k[order], order


def oracle_to_struct(ok):
    """oracle efo_keypoint records -> same field names as efb200.KEYPOINT_DTYPE"""
    import efb200
    out = np.zeros(len(ok), efb200.KEYPOINT_DTYPE)
    out["x"] = ok["x"]; out["y"] = ok["y"]; out["size"] = ok["size"]; out["angle"] = ok["angle"]
    out["response"] = ok["response"]; out["octave"] = ok["octave"]
    return out


def assert_keypoints_equal(gpu, ora):
    """bit-identical (x, y, octave, response, angle, size) sets modulo order"""
    assert len(gpu) == len(ora), f"keypoint count differs: gpu {len(gpu)} oracle {len(ora)}"
    g, _ = canon_keypoints(gpu)
    o, _ = canon_keypoints(ora)
    for f in ("x", "y", "octave"):
        assert np.array_equal(g[f], o[f]), f"keypoint field {f} differs"
    for f in ("response", "angle", "size"):
        gb, ob = g[f].view(np.uint32), o[f].view(np.uint32)
        bad = np.nonzero(gb != ob)[0]
        assert len(bad) == 0, f"{f}: {len(bad)} of {len(g)} differ, first {g[f][bad[:3]]} vs {o[f][bad[:3]]}"

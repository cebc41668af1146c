"""Shared-memory aliasing of the HashSIFT feature kernel (csrc/ef_hashsift.cu): the 32x32 patch lives in the tail of the keypoint block
and the late fraction records are written on top of it while the gradient pass is still reading patch rows.  This replays the gradient
pass's schedule (16 lanes, EF_SIFT_GB pixels per lane and step, reads before writes inside a step) with the constants parsed from the
source and checks that no step reads a patch byte an EARLIER step has overwritten, and that the staged window ends below the patch."""
import re
from pathlib import Path

SRC = (Path(__file__).resolve().parent.parent / "cuda-efficient-features_b200" / "csrc" / "ef_hashsift.cu").read_text()


def const(name):
    m = re.search(r"#define\s+%s\s+(\d+)" % name, SRC)
    assert m, name
    return int(m.group(1))


def test_patch_is_never_read_after_a_record_overwrote_it():
    rec, blk, off, gb = const("EF_SIFT_REC"), const("EF_SIFT_BLK"), const("EF_SIFT_PATCH_OFF"), const("EF_SIFT_GB")
    assert off + 1024 <= 4 * blk
    assert const("EF_SIFT_WIN_ROWS") * const("EF_SIFT_WIN_PITCH") <= off
    for k in (0, 1):
        written = set()
        for s0 in range(0, 900, 16 * gb):
            reads, writes = set(), set()
            for hl in range(16):
                for u in range(gb):
                    i0 = s0 + hl
                    i = min(i0 + 16 * u, 899)
                    y, x = divmod(i, 30)
                    c = off + (y + 1) * 32 + x + 1
                    reads.update((c + 1, c - 1, c - 32, c + 32))
                    if i0 + 16 * u < 900:
                        rix = i + 2 * (y >> 3)
                        for b in range(4):
                            writes.add(4 * (k + rix) + b)
                            writes.add(4 * (rec + k + rix) + b)
            assert not (reads & written), (k, s0, sorted(reads & written)[:4])
            written |= writes
        assert max(written) < 4 * blk
        # the all-zero spare record is written after the pass and must lie inside the block too
        assert 4 * (rec + k + const("EF_SIFT_ZERO")) + 4 <= 4 * blk

"""Shared-memory aliasing of the HashSIFT feature kernel (csrc/ef_hashsift.cu): the 32x32 patch lives in the tail of the keypoint block
and the late fraction records are written on top of it while the gradient pass is still reading patch rows.  This replays the gradient
pass's schedule (lane hl owns columns 2 hl, 2 hl + 1; EF_SIFT_GROWS rows per step; the patch rows y0 + 2 .. y0 + GROWS + 1 are loaded at the
top of a step, rows 0 and 1 before the first; reads before writes inside a step) with the constants parsed from the source and checks
that no step reads a patch byte an EARLIER step has overwritten, and that the staged window ends below the patch."""
import re
from pathlib import Path

SRC = (Path(__file__).resolve().parent.parent / "cuda-efficient-features_b200" / "csrc" / "ef_hashsift.cu").read_text()


def const(name):
    m = re.search(r"#define\s+%s\s+(\d+)" % name, SRC)
    assert m, name
    return int(m.group(1))


def test_patch_is_never_read_after_a_record_overwrote_it():
    rec, blk, off, rows = const("EF_SIFT_REC"), const("EF_SIFT_BLK"), const("EF_SIFT_PATCH_OFF"), const("EF_SIFT_GROWS")
    assert off + 1024 <= 4 * blk and 30 % rows == 0
    assert const("EF_SIFT_WIN_ROWS") * const("EF_SIFT_WIN_PITCH") <= off
    for k in (0, 1):
        written = set()
        for y0 in range(0, 30, rows):
            reads, writes = set(), set()
            for hl in range(16):
                xl = min(2 * hl, 28)
                for r in ([0, 1] if y0 == 0 else []) + list(range(y0 + 2, y0 + rows + 2)):
                    reads.update(off + 32 * r + xl + b for b in range(4))
                if hl < 15:
                    for y in range(y0, y0 + rows):
                        for x in (xl, xl + 1):
                            rix = y * 30 + 2 * (y >> 3) + x
                            for b in range(4):
                                writes.add(4 * (k + rix) + b)
                                writes.add(4 * (rec + k + rix) + b)
            assert not (reads & written), (k, y0, sorted(reads & written)[:4])
            written |= writes
        assert max(written) < 4 * blk
        # the all-zero spare record is written after the pass and must lie inside the block too
        assert 4 * (rec + k + const("EF_SIFT_ZERO")) + 4 <= 4 * blk

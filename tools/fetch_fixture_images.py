"""Copies the reference's own test photographs (tests/data/images/100_71{00..10}.JPG, the only fixtures its test suite has:
tests/descriptor_test.cpp:16-75) into tests/golden/images/ and records the SHA-256 of the DECODED grayscale pixels
(cv2.imread(..., IMREAD_GRAYSCALE), what descriptor_test.cpp:32 feeds the detector).  /root/reference does not exist on
the GPU box, so the files are committed; run here once:  python tools/fetch_fixture_images.py"""
import hashlib
import json
import shutil
from pathlib import Path

import cv2

SRC = Path("/root/reference/tests/data/images")
DST = Path(__file__).resolve().parent.parent / "tests" / "golden" / "images"

if __name__ == "__main__":
    DST.mkdir(parents=True, exist_ok=True)
    sums = {}
    for i in range(11):
        name = f"100_71{i:02d}.JPG"
        shutil.copyfile(SRC / name, DST / name)
        (DST / name).chmod(0o644)
        img = cv2.imread(str(DST / name), cv2.IMREAD_GRAYSCALE)
        sums[name] = {"shape": list(img.shape), "sha256_gray": hashlib.sha256(img.tobytes()).hexdigest()}
        print(name, img.shape, sums[name]["sha256_gray"][:16])
    (DST / "decoded_sha256.json").write_text(json.dumps(sums, indent=1) + "\n")

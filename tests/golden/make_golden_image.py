"""Regenerates tests/golden/image_pack.npz from the REAL reference: the files its own sp::write (Image/Image.cpp:14-76)
produces for the sums of tests/imagecases.py divided as render_thread divides them (main.cpp:100-102) — the PFM payload and
the numbers of the PPM.  Development container only:   python tests/golden/make_golden_image.py"""
from __future__ import annotations

import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from oracle import ref  # noqa: E402
import imagecases  # noqa: E402


def main() -> None:
    if not ref.available():
        raise SystemExit("oracle/_ref/libsp_ref.so missing")
    img, spp = imagecases.sums()
    with tempfile.TemporaryDirectory() as d:
        ref.write_image(img, spp, f"{d}/a.pfm")
        ref.write_image(img, spp, f"{d}/a.ppm")
        pfm = imagecases.read_pfm_payload(f"{d}/a.pfm").copy()
        ppm = imagecases.read_ppm_numbers(f"{d}/a.ppm")
    np.savez_compressed(Path(__file__).resolve().parent / "image_pack.npz", sums=img, spp=np.array(spp), pfm=pfm, ppm=ppm)
    print("pfm", pfm.shape, "ppm", ppm.shape, "max number", int(ppm.max()), "min", int(ppm.min()))


if __name__ == "__main__":
    main()

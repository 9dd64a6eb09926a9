"""Inputs of the image-packing parity tests and readers for the two file formats the reference writes
(Image/Image.cpp:14-55)."""
from __future__ import annotations

import numpy as np


def sums(seed: int = 11, w: int = 48, h: int = 20, spp: int = 7) -> tuple[np.ndarray, int]:
    """Per-pixel radiance sums with everything the sRGB curve distinguishes: zeros, values around the 0.0031308 knee,
    the unit range, values above 1 (bright pixels print numbers above 255), large values that saturate the uint16."""
    rng = np.random.default_rng(seed)
    img = (rng.random((h, w, 3), dtype=np.float32) * np.float32(spp)).astype(np.float32)
    img[0] = 0.0
    img[1] = (np.float32(0.0031308) * spp * (1 + (rng.random((w, 3), dtype=np.float32) - 0.5) * np.float32(1e-3))).astype(np.float32)
    img[2] = np.float32(0.0031308) * np.float32(spp)
    img[3] = (rng.random((w, 3), dtype=np.float32) * np.float32(1e-4 * spp)).astype(np.float32)
    img[4] = (rng.random((w, 3), dtype=np.float32) * np.float32(40.0 * spp)).astype(np.float32)
    img[5] = np.float32(1.0) * np.float32(spp)
    img[6, :, 0] = np.float32(3e6 * spp)
    img[7] = np.float32(1e-30)
    return img, spp


def read_pfm_payload(path) -> np.ndarray:
    raw = open(path, "rb").read()
    parts = raw.split(b"\n", 3)
    assert parts[0] == b"PF" and float(parts[2]) < 0, "little-endian colour PFM expected"
    w, h = (int(v) for v in parts[1].split())
    return np.frombuffer(parts[3], dtype="<f4").reshape(h, w, 3)


def read_ppm_numbers(path) -> np.ndarray:
    tok = open(path, "rb").read().split()
    assert tok[0] == b"P3" and tok[3] == b"255"
    w, h = int(tok[1]), int(tok[2])
    return np.array([int(t) for t in tok[4:]], dtype=np.int64).reshape(h, w, 3)

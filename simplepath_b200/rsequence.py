"""The pixel jitter of the reference's render loop, host side.

`RSequenceSampler` (math/Sampler.h:138-178) draws pixel offsets from `RSequence<2>` (:15-62): point n is
mod1(fseed + alpha_i * (n + 1)) with alpha_i = mod1((1/g)^(i+1)) and g the generalised golden ratio from 10 fixed-point
iterations of powf (:18-27).  fseed = float(seed ^ 0x6184faf4) / FLT_MAX is below 2^-96 for every pixel seed
(x << 16 | y, main.cpp:67-71) and is absorbed by the first float addition, so ONE spp x 2 table serves every pixel.
Everything is float32 in the reference's operation order, with libm's powf called through ctypes (numpy may pick a
vectorised pow with different rounding); tests pin the table bitwise against the one the reference itself produces
(tests/golden/jitter.npz)."""
from __future__ import annotations

import ctypes
import ctypes.util

import numpy as np

_libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
_libm.powf.argtypes = [ctypes.c_float, ctypes.c_float]
_libm.powf.restype = ctypes.c_float
f32 = np.float32


def _powf(a, b) -> np.float32:
    return f32(_libm.powf(float(f32(a)), float(f32(b))))


def _phi(dimension: int) -> np.float32:
    x = f32(2.0)
    for _ in range(10):
        x = _powf(f32(1.0) + x, f32(1.0) / (f32(dimension) + f32(1.0)))
    return x


def alphas(dimension: int = 2) -> np.ndarray:
    g = _phi(dimension)
    out = np.empty(dimension, dtype=np.float32)
    for i in range(dimension):
        out[i] = np.modf(_powf(f32(1.0) / g, f32(i) + f32(1.0)))[0]
    return out


def jitter_table(spp: int) -> np.ndarray:
    """[spp, 2] float32: the offsets RSequenceSampler::get_next_2D returns for samples 0..spp-1 of any pixel."""
    a = alphas(2)
    n1 = np.arange(spp, dtype=np.float32) + f32(1.0)
    # fseed (< 2^-96) + alpha * (n + 1): the addition returns the product unchanged
    prod = a[None, :] * n1[:, None]
    return np.modf(prod.astype(np.float32))[0].astype(np.float32)

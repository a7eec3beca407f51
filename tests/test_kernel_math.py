"""Bit-level restatements (numpy, fp32) of arithmetic tricks used inside the CUDA kernels, checked on the CPU.

These are not the kernels (those are compared with the oracle on the GPU, ``tests/test_gpu_kernels.py``); they pin the
constants the kernels hard-code so that a typo in a coefficient cannot hide behind the bf16 tolerance of the GPU tests.
"""
import re
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ex2_poly_constants():
    """The coefficients of ``ex2_poly`` as written in csrc/sm100_ptx.cuh (hex float literals)."""
    src = open(os.path.join(ROOT, "modaltune_b200", "csrc", "sm100_ptx.cuh")).read()
    body = src[src.index("float ex2_poly(float x)"):]
    body = body[:body.index("return")]
    hexes = re.findall(r"0x1\.[0-9a-f]+p-\d+f", body)
    assert len(hexes) == 4, hexes
    c3, c2, c1, c0 = (np.float32(float.fromhex(h[:-1])) for h in hexes)
    return c0, c1, c2, c3


def ex2_poly(x):
    """fp32 restatement of sm100_ptx.cuh:ex2_poly (Cody-Waite split, degree-3 polynomial, exponent add)."""
    c0, c1, c2, c3 = _ex2_poly_constants()
    x = np.maximum(x.astype(np.float32), np.float32(-125.0))
    magic = np.float32(12582912.0)
    t = (x + magic).astype(np.float32)
    f = (x - (t - magic).astype(np.float32)).astype(np.float32)
    p = (f * c3 + c2).astype(np.float32)
    p = (p * f + c1).astype(np.float32)
    p = (p * f + c0).astype(np.float32)
    bits = p.view(np.uint32) + (t.view(np.uint32) << np.uint32(23))
    return bits.view(np.float32)


def test_ex2_poly_relative_error():
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-40, 9, 200000), np.linspace(-125, 9, 20001), [-0.5, 0.5, 0.0, -1.0, 8.0]])
    x = x.astype(np.float32)
    got = ex2_poly(x).astype(np.float64)
    want = np.exp2(x.astype(np.float64))
    rel = np.abs(got - want) / want
    assert rel.max() < 1e-4, rel.max()            # below the 2^-9 rounding of the bf16 probabilities it feeds
    assert np.all(got > 0)


def test_ex2_poly_clamps_instead_of_underflowing():
    x = np.array([-1e4, -126.0, -125.5, -np.inf], dtype=np.float32)
    got = ex2_poly(x)
    assert np.all(np.isfinite(got)) and np.all(got > 0) and np.all(got <= np.float32(2.0 ** -124))


def _erf_fast_constants():
    src = open(os.path.join(ROOT, "modaltune_b200", "csrc", "elementwise.cu")).read()
    body = src[src.index("float erf_fast(float x)"):]
    body = body[:body.index("asm(")]
    hexes = re.findall(r"-?0x1\.[0-9a-f]+p[-+]\d+f", body)
    assert len(hexes) == 6, hexes
    return [np.float32(float.fromhex(h[:-1])) for h in hexes]   # c6 .. c1


def erf_fast(x):
    """fp32 restatement of elementwise.cu:erf_fast: 1 - 2^-(t * poly(t)), t = min(|x|, 4)."""
    cs = _erf_fast_constants()
    x = x.astype(np.float32)
    t = np.minimum(np.abs(x), np.float32(4.0))
    r = np.full_like(t, cs[0])
    for c in cs[1:]:
        r = (r * t + c).astype(np.float32)
    e = np.exp2(-(r * t).astype(np.float32).astype(np.float64)).astype(np.float32)
    return np.copysign((np.float32(1.0) - e).astype(np.float32), x)


def test_erf_fast_absolute_error():
    from scipy.special import erf
    x = np.concatenate([np.linspace(-6, 6, 240001), [0.0, 1e-6, -1e-6, 4.0, 100.0, -100.0]]).astype(np.float32)
    got = erf_fast(x).astype(np.float64)
    err = np.abs(got - erf(x.astype(np.float64)))
    assert err.max() < 5e-7, err.max()
    # what the kernels use it for: GELU(x) = x/2 (1 + erf(x / sqrt 2)) within 1e-6 absolute on the whole range
    g = 0.5 * x.astype(np.float64) * (1.0 + erf_fast((x * np.float32(0.70710678118654752)).astype(np.float32)))
    g_ref = 0.5 * x.astype(np.float64) * (1.0 + erf(x.astype(np.float64) / np.sqrt(2.0)))
    assert np.abs(g - g_ref).max() < 2e-5 and np.abs(g - g_ref)[np.abs(x) < 6].max() < 2e-6

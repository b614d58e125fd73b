"""CPU tests of the drop-in boundary: the shared object loads, exports what include/raisr_b200.h
declares, and fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes
import os
import re

import numpy as np
import pytest

from oclcomputervision_b200 import _cabi
from tests.conftest import HAS_GPU, ROOT


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "raisr_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b((?:raisr|ocv)_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    lib = _cabi.load()
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n
        assert n in _cabi.SIGNATURES, "binding missing for " + n
    assert b"sm_100a" in lib.raisr_version()


def test_band_source_rows_is_pure_host_logic():
    lib = _cabi.load()
    first, last = ctypes.c_int(), ctypes.c_int()
    # config 4 of BASELINE.json: 16384 source rows, x3, band 1 of 8 -> rows 2046..4098
    assert lib.raisr_band_src_rows(16384, 3, 6144, 6144, ctypes.byref(first), ctypes.byref(last)) == 0
    assert first.value >= 2044 and last.value <= 4099 and first.value <= 2048 and last.value >= 4095
    assert lib.raisr_band_src_rows(16384, 3, 0, 6144, ctypes.byref(first), ctypes.byref(last)) == 0
    assert first.value == 0
    assert lib.raisr_band_src_rows(0, 3, 0, 3, ctypes.byref(first), ctypes.byref(last)) == _cabi.E_ARG


@pytest.mark.skipif(HAS_GPU, reason="checks the no-GPU failure mode")
def test_no_cpu_fallback_without_a_device():
    lib = _cabi.load()
    h = ctypes.c_void_p()
    rc = lib.raisr_create(ctypes.byref(h), 0, 24, 3, 3, 11)
    assert rc == _cabi.E_CUDA and not h.value
    assert b"no CPU fallback" in lib.raisr_last_error()
    from oclcomputervision_b200 import ClRaisr
    with pytest.raises(_cabi.RaisrError):
        ClRaisr(1)


def test_argument_errors_do_not_need_a_device():
    lib = _cabi.load()
    h = ctypes.c_void_p()
    assert lib.raisr_create(ctypes.byref(h), 0, 24, 3, 3, 7) == _cabi.E_ARG      # filter_len must be 11
    assert lib.raisr_create(ctypes.byref(h), 0, 64, 3, 3, 11) == _cabi.E_ARG     # > 256 buckets
    assert lib.raisr_create(None, 0, 24, 3, 3, 11) == _cabi.E_ARG
    assert lib.raisr_sync(None) == _cabi.E_ARG


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "oclcomputervision_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle|libraisr_oracle|#include\s+\"[^\"]*oracle", txt, re.M), f


def test_filter_record_packing_covers_every_tap_once():
    # mirror of octet_pack_filter_s (csrc/raisr_octet.cuh): each of the 121 taps lands in exactly one slot
    for S in (2, 3, 4):
        newp = min(S, 5)
        seen = {}
        def put(lane, slot, tap):
            idx = (lane + 8 * (slot // 4)) * 4 + slot % 4
            assert idx not in seen and 0 <= idx < 128
            seen[idx] = tap
        for p in range(8):
            for j in range(11):
                put(p, j, p * 11 + j)
        for p in range(6):
            for t in range(5):
                put(p, 11 + t, (8 + p // 2) * 11 + 1 + 5 * (p % 2) + t)
        for i in range(3):
            put(6 + i // newp, 16 - newp + i % newp, (8 + i) * 11 + 0)
        assert sorted(seen.values()) == list(range(121))


def _decode_b24_record(rec: np.ndarray):
    """Independent decode of a 384-byte record following the layout documented in include/raisr_b200.h:
    lane p owns chunks p, p+8, p+16 (a circular 48-byte stream); slot k = little-endian u32 at stream byte 3k."""
    out = np.zeros((8, 16), np.float32)
    for p in range(8):
        stream = np.concatenate([rec[(p + 8 * j) * 16:(p + 8 * j) * 16 + 16] for j in range(3)])
        for k in range(16):
            b = [int(stream[(3 * k + i) % 48]) for i in range(4)]
            out[p, k] = np.array([b[0] | b[1] << 8 | b[2] << 16 | b[3] << 24], np.uint32).view(np.float32)[0]
    return out


@pytest.mark.parametrize("scale", [2, 3, 4])
def test_b24_tap_records_round_to_within_2_pow_minus_16(scale):
    """The 24-bit tap records (raisr_pack_taps_b24, csrc/raisr_octet.cuh): every tap of the filter appears exactly
    once among the decoded slots, its decoded value is what the packer reports as `effective`, and it is within
    2^-16 relative of the fp32 tap (sign + exponent + 15 mantissa bits, low byte chosen jointly)."""
    lib = _cabi.load()
    rng = np.random.default_rng(scale)
    for trial in range(20):
        f = rng.normal(0, 0.02, 121).astype(np.float32)
        f[60] += 1
        if trial == 1:
            f[:] = 0
        if trial == 2:
            f = (rng.standard_normal(121) * 10.0 ** rng.integers(-20, 20, 121)).astype(np.float32)
        if trial == 3:
            f[::2] = np.float32(1.9999999)      # mantissa all ones: rounding carries into the exponent
        rec = np.zeros(384, np.uint8)
        eff = np.full(121, np.nan, np.float32)
        assert lib.raisr_pack_taps_b24(f.ctypes.data, scale, rec.ctypes.data, eff.ctypes.data) == 0
        assert np.isfinite(eff).all()
        tol = np.abs(f).astype(np.float64) * 2.0 ** -16 + 1e-40
        assert (np.abs(eff.astype(np.float64) - f) <= tol).all()
        slots = _decode_b24_record(rec).ravel()
        # every effective tap is among the decoded slot values; the unused slots decode to (sub)normal dust
        used = np.zeros(slots.size, bool)
        for t in range(121):
            hit = np.flatnonzero((slots.view(np.uint32) == eff[t:t + 1].view(np.uint32)[0]) & ~used)
            assert hit.size >= 1, t
            used[hit[0]] = True
        assert (np.abs(slots[~used]) < 1e-35).all()
    assert lib.raisr_pack_taps_b24(None, scale, rec.ctypes.data, eff.ctypes.data) == _cabi.E_ARG
    assert lib.raisr_pack_taps_b24(f.ctypes.data, 5, rec.ctypes.data, eff.ctypes.data) == _cabi.E_UNSUPPORTED

"""The value quantiser (IO/Quantize.h as IO/RAWConverter.cpp:205-300 calls it, and
AbstrConverter::Process8Bits): the oracle restatement (oracle/orc_quantize.c) against
  * the reference's OWN known-answer tests (IO/test/quantize.h verify_type / verify_8b_type: 100 consecutive values
    starting at -64 for signed types, at 0 for unsigned ones), and
  * IO/Quantize.h itself, compiled in place (oracle/_ref/ref_quantize), on random data of every input type.
The CUDA quantiser is compared with the oracle in tests/test_gpu_quantize.py."""
import os
import subprocess

import numpy as np
import pytest

from oracle import orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "ref_quantize")
NAMES = {np.dtype(np.int8): "i8", np.dtype(np.uint8): "u8", np.dtype(np.int16): "i16", np.dtype(np.uint16): "u16",
         np.dtype(np.int32): "i32", np.dtype(np.uint32): "u32", np.dtype(np.float32): "f32", np.dtype(np.float64): "f64"}


def have_ref():
    if not os.path.exists(BIN) and os.path.isdir("/root/reference/IO"):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-j8", "ref"], stdout=subprocess.DEVNULL)
    return os.path.exists(BIN)


def run_reference(tmp_path, src, out_bits):
    inp, out, hist = (os.path.join(str(tmp_path), n) for n in ("in.raw", "out.raw", "hist.txt"))
    src.tofile(inp)
    subprocess.check_call([BIN, inp, NAMES[src.dtype], str(src.size), str(out_bits), out, hist],
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    lines = open(hist).read().split()
    changed = int(lines[1])
    h = np.array([int(v) for v in lines[2:]], np.uint64)
    # (the reference creates the file with iSize * sizeof(U) BYTES where iSize already counts bytes: only the first n values are data)
    dst = np.fromfile(out, np.uint8 if out_bits == 8 else np.uint16)[:src.size] if changed else None
    return dst, h, changed


@pytest.mark.parametrize("dtype", [np.int16, np.uint16, np.int32, np.uint32])
def test_reference_kat_verify_type(dtype):
    """IO/test/quantize.h verify_type<T>: value i of 100 consecutive ones maps to (i) * min(65535 / 100, 1) = i and
    every one of the first 100 histogram bins holds 1"""
    start = -64 if np.dtype(dtype).kind == "i" else 0
    src = np.arange(start, start + 100).astype(dtype)
    dst, hist, info = orc.quantize(src, 16)
    got = src.astype(np.uint16) if dst is None else dst           # "if(!quantize(...)) outfn = fn": the input is used as is
    assert np.array_equal(got, np.arange(100, dtype=np.uint16))
    assert np.array_equal(hist[:100], np.ones(100, np.uint64)) and not hist[100:].any()


@pytest.mark.parametrize("dtype", [np.int8, np.uint8])
def test_reference_kat_8_bit(dtype):
    """verify_type<tbyte / tubyte>: Process8Bits -- signed bytes are biased by 128, unsigned ones are used as is"""
    start = -64 if dtype == np.int8 else 0
    src = np.arange(start, start + 100).astype(dtype)
    dst, hist, info = orc.quantize(src)
    if dtype == np.int8:
        assert np.array_equal(dst, (np.arange(100) + start + 128).astype(np.uint8))
        lo = start + 128
    else:
        assert dst is None
        lo = 0
    want = np.zeros(256, np.uint64)
    want[lo:lo + 100] = 1
    assert np.array_equal(hist, want)


def random_input(dtype, n, seed, spread):
    rng = np.random.default_rng(seed)
    dt = np.dtype(dtype)
    if dt.kind == "f":
        return (rng.standard_normal(n) * spread + 3.0).astype(dtype)
    info = np.iinfo(dtype)
    lo = max(info.min, -spread) if dt.kind == "i" else 0
    hi = min(info.max, spread)
    return rng.integers(lo, hi, size=n, endpoint=True).astype(dtype)


CASES = [(np.int16, 30000, 16), (np.int16, 300, 16), (np.uint16, 65535, 16), (np.uint16, 3000, 16), (np.uint16, 200, 16),
         (np.int32, 1 << 20, 16), (np.int32, 1000, 16), (np.uint32, 1 << 30, 16), (np.uint32, 4000, 16), (np.float32, 10.0, 16),
         (np.float32, 1e-3, 16), (np.float64, 1e6, 16), (np.int16, 30000, 8), (np.uint16, 65535, 8), (np.float32, 5.0, 8),
         (np.int8, 127, 8), (np.uint8, 255, 8), (np.uint16, 255, 8)]


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref/ref_quantize not built (reference tree absent)")
@pytest.mark.parametrize("dtype,spread,bits", CASES)
def test_oracle_matches_reference_quantize_h(tmp_path, dtype, spread, bits):
    src = random_input(dtype, 20011, hash((np.dtype(dtype).name, spread, bits)) & 0xFFFF, spread)
    ref_dst, ref_hist, ref_changed = run_reference(tmp_path, src, bits)
    dst, hist, info = orc.quantize(src, bits)
    assert bool(info.changed) == bool(ref_changed)
    if len(ref_hist) == 0:       # early return of Quantize (data fits as it is): the reference never sets the histogram
        assert not info.hist_set and info.bin_count == int((hist != 0).sum()) and not ref_changed
    else:
        assert info.hist_set and np.array_equal(hist[:len(ref_hist)], ref_hist) and not hist[len(ref_hist):].any()
    if ref_changed:
        assert np.array_equal(dst, ref_dst)
    else:
        assert dst is None


def test_constant_input_is_defined():
    """the reference divides by zero on a constant input; the restatement maps it to 0 (stated in orc_quantize.c)"""
    dst, hist, info = orc.quantize(np.full(100, 7.5, np.float32), 16)
    assert dst is not None and not dst.any() and hist[0] == 100

"""-m gpu: the CUDA value quantiser (tvk_quantize, csrc/k_quantize.cu) against the CPU oracle (oracle/orc_quantize.c,
pinned to IO/Quantize.h compiled in place and to the reference's own known-answer tests in tests/test_quantize.py):
quantised values, histogram, range, factor, bin count and the changed / hist_set flags, bit for bit."""
import numpy as np
import pytest
import torch

import tuvok_b200 as tb
from oracle import orc
from test_quantize import CASES, random_input
from tuvok_b200 import _lib as L

pytestmark = pytest.mark.gpu
ST = {np.dtype(np.int8): L.ST_I8, np.dtype(np.uint8): L.ST_U8, np.dtype(np.int16): L.ST_I16, np.dtype(np.uint16): L.ST_U16,
      np.dtype(np.int32): L.ST_I32, np.dtype(np.uint32): L.ST_U32, np.dtype(np.float32): L.ST_F32, np.dtype(np.float64): L.ST_F64}


def gpu_quantize(r, src, bits):
    raw = torch.from_numpy(src.view(np.uint8).copy()).cuda()
    out = torch.zeros(src.size * 2, dtype=torch.uint8, device="cuda")
    hist, info = r.Quantize(raw.data_ptr(), ST[src.dtype], src.size, bits, out.data_ptr())
    bits = 8 if src.dtype.itemsize == 1 else bits
    dst = out.cpu().numpy()[:src.size * (bits // 8)].view(np.uint8 if bits == 8 else np.uint16)
    return (dst if info.changed else None), hist, info


@pytest.mark.parametrize("dtype,spread,bits", CASES)
@pytest.mark.parametrize("n", [20011, 1 << 20])
def test_quantiser_equals_oracle(dtype, spread, bits, n):
    src = random_input(dtype, n, hash((np.dtype(dtype).name, spread, bits, n)) & 0xFFFF, spread)
    want_dst, want_hist, wi = orc.quantize(src, bits)
    r = tb.CudaGridLeaper()
    dst, hist, gi = gpu_quantize(r, src, bits)
    r.Cleanup()
    assert bool(gi.changed) == bool(wi.changed) and bool(gi.hist_set) == bool(wi.hist_set)
    assert np.array_equal(hist, want_hist)
    if wi.changed:
        assert np.array_equal(dst, want_dst)
    else:
        assert dst is None
    if src.dtype.itemsize > 1:
        assert (gi.min, gi.max, gi.factor, gi.bin_count) == (wi.min, wi.max, wi.factor, wi.bin_count)


@pytest.mark.parametrize("dtype", [np.int16, np.uint16, np.int32, np.uint32, np.int8, np.uint8])
def test_reference_kat_on_the_gpu(dtype):
    """IO/test/quantize.h verify_type: 100 consecutive values map to 0..99 (signed bytes to value + 128), one per bin"""
    start = -64 if np.dtype(dtype).kind == "i" else 0
    src = np.arange(start, start + 100).astype(dtype)
    r = tb.CudaGridLeaper()
    dst, hist, info = gpu_quantize(r, src, 16)
    r.Cleanup()
    if np.dtype(dtype).itemsize == 1:
        lo = start + 128 if dtype == np.int8 else 0
        if dtype == np.int8:
            assert np.array_equal(dst, (np.arange(100) + lo).astype(np.uint8))
        assert np.array_equal(hist[lo:lo + 100], np.ones(100, np.uint64)) and hist.sum() == 100
    else:
        got = src.astype(np.uint16) if dst is None else dst
        assert np.array_equal(got, np.arange(100, dtype=np.uint16))
        assert np.array_equal(hist[:100], np.ones(100, np.uint64)) and not hist[100:].any()


def test_constant_input_and_bad_arguments():
    r = tb.CudaGridLeaper()
    dst, hist, info = gpu_quantize(r, np.full(1000, 7.5, np.float32), 16)
    assert dst is not None and not dst.any() and hist[0] == 1000
    x = torch.zeros(64, dtype=torch.uint8, device="cuda")
    with pytest.raises(tb.TvkError):
        r.Quantize(x.data_ptr() + 2, L.ST_U16, 8, 16, x.data_ptr())         # misaligned input
    with pytest.raises(tb.TvkError):
        r.Quantize(x.data_ptr(), 11, 8, 16, x.data_ptr())                    # unknown type
    with pytest.raises(tb.TvkError):
        r.Quantize(x.data_ptr(), L.ST_F32, 8, 12, x.data_ptr())              # 12-bit output does not exist
    r.Cleanup()

/* orc_quantize.c -- CPU restatement of the reference's value quantiser (TEST INFRASTRUCTURE: the checker of
 * tvk_quantize; never linked into the product).  Follows, statement by statement:
 *   Quantize<T, U>                 IO/Quantize.h:427-577  (range pass io_minmax :294-343, early return for data that
 *                                  already fits :462-482, QuantizationFactor :49-72, the map + histogram loop :522-535)
 *   AbstrConverter::Process8Bits   IO/AbstrConverter.cpp:73-157 (signed bytes + 128, histogram of the bytes)
 * as RAWConverter's quantize() calls them (IO/RAWConverter.cpp:205-300): 16 / 32-bit integers and float / double to
 * 16-bit (or 8-bit) unsigned with a 4096 (256) bin histogram.  Pinned by the reference's own known-answer tests
 * (IO/test/quantize.h verify_type / verify_8b_type: 100 consecutive values) and against Quantize.h itself compiled in
 * place (oracle/_ref/ref_quantize) on random data -- tests/test_quantize.py.
 * Defined where the reference is undefined: a constant input (max == min) maps to 0 (the reference divides by zero). */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <string.h>
#include "orc.h"

#define QMIN(a, b) ((a) < (b) ? (a) : (b))

/* T = input type, D = type of (v - min) as C++ evaluates it, FACT = how QuantizationFactor is evaluated */
#define DEF_QUANT(NAME, T, IS_SIGNED, IS_FP, SUB, FACT)                                                          \
  static void quant_##NAME(const T* src, uint64_t n, int out_bits, void* dst, uint64_t* hist, orc_quantize_info* info) { \
    const size_t hist_size = out_bits == 8 ? 256 : 4096;                                                        \
    T mn = src[0], mx = src[0];                                                                                 \
    for (uint64_t i = 1; i < n; i++) { if (src[i] < mn) mn = src[i]; if (src[i] > mx) mx = src[i]; }             \
    memset(hist, 0, hist_size * sizeof(uint64_t));                                                              \
    info->min = (double)mn; info->max = (double)mx; info->factor = 1.0; info->bin_count = 0; info->changed = 0;  \
    info->hist_set = 1;                                                                                         \
    /* unsigned data that already fits needs no processing: histogram of the values themselves */               \
    if (!(IS_SIGNED) && (double)mx < (double)hist_size && sizeof(T) <= (hist_size == 256 ? 1u : 2u)) {          \
      for (uint64_t i = 0; i < n; i++) hist[(size_t)src[i]]++;                                                  \
      for (size_t b = 0; b < hist_size; b++) if (hist[b]) info->bin_count++;                                    \
      info->hist_set = 0;   /* the reference returns before Histogram1D->SetHistogram (Quantize.h:469-482) */   \
      return;                                                                                                   \
    }                                                                                                           \
    info->bin_count = (IS_FP) ? 4096 : (uint64_t)((double)mx - (double)mn) + 1;                                 \
    const size_t max_out = out_bits == 8 ? 255 : 65535;                                                         \
    double f, fh;                                                                                               \
    if (mx == mn) { f = 0.0; fh = 0.0; }                                                                        \
    else { FACT(f, max_out); FACT(fh, hist_size - 1); }                                                         \
    info->factor = f;                                                                                           \
    info->changed = f != 1.0 || mn != 0 || sizeof(T) > 2 || sizeof(T) > (size_t)(out_bits / 8);                 \
    for (uint64_t i = 0; i < n; i++) {                                                                          \
      const double d = (double)(SUB);                                                                           \
      if (out_bits == 8) {                                                                                      \
        ((uint8_t*)dst)[i] = QMIN((uint8_t)max_out, (uint8_t)(d * f));                                          \
        hist[QMIN((uint8_t)(hist_size - 1), (uint8_t)(d * fh))]++;                                              \
      } else {                                                                                                  \
        ((uint16_t*)dst)[i] = QMIN((uint16_t)max_out, (uint16_t)(d * f));                                       \
        hist[QMIN((uint16_t)(hist_size - 1), (uint16_t)(d * fh))]++;                                            \
      }                                                                                                         \
    }                                                                                                           \
  }

/* integers: min(max_out / (double(mx) - mn), 1.0); float: size_t / float in float; double: in double */
#define FACT_INT(out, M) out = (double)(M) / ((double)mx - (double)mn); if (out > 1.0) out = 1.0
#define FACT_F32(out, M) out = (double)((float)(M) / (float)(mx - mn))
#define FACT_F64(out, M) out = (double)(M) / (mx - mn)

DEF_QUANT(i16, int16_t, 1, 0, (int)src[i] - (int)mn, FACT_INT)
DEF_QUANT(u16, uint16_t, 0, 0, (int)src[i] - (int)mn, FACT_INT)
DEF_QUANT(i32, int32_t, 1, 0, (int64_t)src[i] - (int64_t)mn, FACT_INT)   /* the reference's int - int can overflow (UB) */
DEF_QUANT(u32, uint32_t, 0, 0, (uint32_t)(src[i] - mn), FACT_INT)
DEF_QUANT(f32, float, 1, 1, (float)(src[i] - mn), FACT_F32)
DEF_QUANT(f64, double, 1, 1, src[i] - mn, FACT_F64)

int orc_quantize(const void* src, int type, uint64_t n, int out_bits, void* dst, uint64_t* hist, orc_quantize_info* info) {
  if (!src || !n || !hist || !info || (out_bits != 8 && out_bits != 16)) return 1;
  memset(info, 0, sizeof(*info));
  switch (type) {
    case ORC_ST_I8:    /* Process8Bits, signed: value + 128 */
    case ORC_ST_U8: {
      memset(hist, 0, 256 * sizeof(uint64_t));
      for (uint64_t i = 0; i < n; i++) {
        const uint8_t v = type == ORC_ST_I8 ? (uint8_t)(((const int8_t*)src)[i] + 128) : ((const uint8_t*)src)[i];
        if (type == ORC_ST_I8) ((uint8_t*)dst)[i] = v;
        hist[v]++;
      }
      info->changed = type == ORC_ST_I8;
      info->factor = 1.0;
      info->hist_set = 1;
      return 0;
    }
    case ORC_ST_I16: quant_i16((const int16_t*)src, n, out_bits, dst, hist, info); return 0;
    case ORC_ST_U16: quant_u16((const uint16_t*)src, n, out_bits, dst, hist, info); return 0;
    case ORC_ST_I32: quant_i32((const int32_t*)src, n, out_bits, dst, hist, info); return 0;
    case ORC_ST_U32: quant_u32((const uint32_t*)src, n, out_bits, dst, hist, info); return 0;
    case ORC_ST_F32: quant_f32((const float*)src, n, out_bits, dst, hist, info); return 0;
    case ORC_ST_F64: quant_f64((const double*)src, n, out_bits, dst, hist, info); return 0;
    default: return 1;
  }
}

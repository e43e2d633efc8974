/* TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's input quantisation.
 *
 * Follows, under /root/reference/Tuvok/IO/:
 *   Quantize.h:429-578      Quantize<T,U>            (min/max, early-out, bias + scale, 256/4096-bin histogram)
 *   Quantize.h:50-72        QuantizationFactor       (integers: min(max_out/(max-min), 1); float/double: stretch)
 *   Quantize.h:248-287      UnsignedHistogram        (only its result on the unsigned early-out path matters)
 *   Quantize.h:583-686      BinningQuantize<T,U>     (<= 256 / 4096 distinct values -> rank remap)
 *   Quantize.h:348-421      ApplyMapping<T,U>
 *   AbstrConverter.cpp:75-157   Process8Bits         (signed char + 128, 256-bin histogram)
 *   AbstrConverter.cpp:167-246  QuantizeTo8Bit
 *   RAWConverter.cpp:196-352    quantize()           (the dispatcher over component size / signedness)
 * for one component and one time step.  PINNED against the reference itself: oracle/_ref/ref_driver
 * `quantize` calls the reference templates on the same inputs (tests/test_oracle_quantize.py), and
 * the known answers of IO/test/quantize.h:104-262 are checked as well.
 *
 * Behaviour that is undefined in the reference (int32 range wider than 2^31, NaN input, a constant
 * floating-point volume) is out of contract: this restatement evaluates differences in 64 bits.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { DT_I8 = 0, DT_U8, DT_I16, DT_U16, DT_I32, DT_U32, DT_F32, DT_I64, DT_U64, DT_F64 };

typedef struct {
  uint32_t out_bits;     /* component size after quantisation (8 or 16; unchanged input width if untouched) */
  uint32_t generated;    /* 1: `out` holds new data; 0: the caller keeps using the input as it is       */
  uint32_t hist_size;    /* 0: no histogram was produced (Histogram1DDataBlock::Compute runs later)      */
  uint32_t bin_count;    /* iBinCount of the last Quantize<T,U> call (diagnostic)                        */
} orc_quant_result;

/* (in[i] - min) as the double the reference multiplies with the factor */
static double diff_at(const void* in, uint32_t type, uint64_t i, const void* mn) {
  switch (type) {
    case DT_I8:  return (double)((int)((const int8_t*)in)[i] - (int)*(const int8_t*)mn);
    case DT_U8:  return (double)((int)((const uint8_t*)in)[i] - (int)*(const uint8_t*)mn);
    case DT_I16: return (double)((int)((const int16_t*)in)[i] - (int)*(const int16_t*)mn);
    case DT_U16: return (double)((int)((const uint16_t*)in)[i] - (int)*(const uint16_t*)mn);
    case DT_I32: return (double)((int64_t)((const int32_t*)in)[i] - (int64_t)*(const int32_t*)mn);
    case DT_U32: return (double)(uint32_t)(((const uint32_t*)in)[i] - *(const uint32_t*)mn);
    case DT_F32: return (double)(float)(((const float*)in)[i] - *(const float*)mn);
    case DT_I64: return (double)(int64_t)((uint64_t)((const int64_t*)in)[i] - (uint64_t)*(const int64_t*)mn);
    case DT_U64: return (double)(uint64_t)(((const uint64_t*)in)[i] - *(const uint64_t*)mn);
    default:     return ((const double*)in)[i] - *(const double*)mn;
  }
}

static const uint32_t kWidth[10] = {1, 1, 2, 2, 4, 4, 4, 8, 8, 8};
static int is_signed_type(uint32_t t) { return t == DT_I8 || t == DT_I16 || t == DT_I32 || t == DT_I64 || t == DT_F32 || t == DT_F64; }

/* min / max of the stream, stored in the input's own type */
static void minmax_of(const void* in, uint32_t type, uint64_t n, void* mn, void* mx) {
#define MM(T)                                                  \
  {                                                            \
    const T* p = (const T*)in;                                 \
    T a = p[0], b = p[0];                                      \
    for (uint64_t i = 1; i < n; ++i) { if (p[i] < a) a = p[i]; if (p[i] > b) b = p[i]; } \
    *(T*)mn = a; *(T*)mx = b;                                  \
  }
  switch (type) {
    case DT_I8: MM(int8_t) break;   case DT_U8: MM(uint8_t) break;
    case DT_I16: MM(int16_t) break; case DT_U16: MM(uint16_t) break;
    case DT_I32: MM(int32_t) break; case DT_U32: MM(uint32_t) break;
    case DT_F32: MM(float) break;   case DT_I64: MM(int64_t) break;
    case DT_U64: MM(uint64_t) break; default: MM(double) break;
  }
#undef MM
}

/* QuantizationFactor, Quantize.h:50-72 */
static double quant_factor(uint32_t type, size_t max_out, const void* mn, const void* mx) {
  if (type == DT_F32) return (double)((float)max_out / (*(const float*)mx - *(const float*)mn));
  if (type == DT_F64) return (double)max_out / (*(const double*)mx - *(const double*)mn);
  double dmx, dmn;
  switch (type) {
    case DT_I8: dmx = *(const int8_t*)mx; dmn = *(const int8_t*)mn; break;
    case DT_U8: dmx = *(const uint8_t*)mx; dmn = *(const uint8_t*)mn; break;
    case DT_I16: dmx = *(const int16_t*)mx; dmn = *(const int16_t*)mn; break;
    case DT_U16: dmx = *(const uint16_t*)mx; dmn = *(const uint16_t*)mn; break;
    case DT_I32: dmx = *(const int32_t*)mx; dmn = *(const int32_t*)mn; break;
    case DT_U32: dmx = *(const uint32_t*)mx; dmn = *(const uint32_t*)mn; break;
    case DT_I64: dmx = (double)*(const int64_t*)mx; dmn = (double)*(const int64_t*)mn; break;
    default: dmx = (double)*(const uint64_t*)mx; dmn = (double)*(const uint64_t*)mn; break;
  }
  const double q = (double)max_out / (dmx - dmn);
  return q < 1.0 ? q : 1.0;
}

static int min_is_zero(uint32_t type, const void* mn) {
  switch (type) {
    case DT_I8: return *(const int8_t*)mn == 0;   case DT_U8: return *(const uint8_t*)mn == 0;
    case DT_I16: return *(const int16_t*)mn == 0; case DT_U16: return *(const uint16_t*)mn == 0;
    case DT_I32: return *(const int32_t*)mn == 0; case DT_U32: return *(const uint32_t*)mn == 0;
    case DT_F32: return *(const float*)mn == 0;   case DT_I64: return *(const int64_t*)mn == 0;
    case DT_U64: return *(const uint64_t*)mn == 0; default: return *(const double*)mn == 0;
  }
}

/* Quantize<T,U>, Quantize.h:429-578.  out_width = sizeof(U).  Returns bDataWillbeChanged ("generated").
 * *hist_set = 1 when Histogram1D->SetHistogram ran (not on the early-out path). */
static int quantize_tu(const void* in, uint32_t type, uint64_t n, uint32_t out_width, void* out,
                       uint64_t* hist, uint32_t* hist_size, int* hist_set, size_t* bin_count) {
  const size_t hsz = out_width == 1 ? 256 : 4096;
  *hist_size = (uint32_t)hsz;
  *hist_set = 0;
  if (bin_count) *bin_count = 0;
  uint64_t mn8, mx8;                       /* storage for a value of any of the ten types */
  minmax_of(in, type, n, &mn8, &mx8);
  /* unsigned data that already fits needs no processing */
  if (!is_signed_type(type) && kWidth[type] <= (hsz == 256 ? 1u : 2u)) {
    uint16_t mx16;
    memcpy(&mx16, &mx8, 2);
    const uint32_t mxv = type == DT_U8 ? (uint32_t)(mx16 & 0xFFu) : (uint32_t)mx16;
    if (mxv < hsz) {
      if (bin_count) {
        memset(hist, 0, sizeof(uint64_t) * hsz);
        for (uint64_t i = 0; i < n; ++i) hist[type == DT_U8 ? ((const uint8_t*)in)[i] : ((const uint16_t*)in)[i]]++;
        for (size_t b = 0; b < hsz; ++b) if (hist[b]) (*bin_count)++;
      }
      return 0;
    }
  }
  memset(hist, 0, sizeof(uint64_t) * hsz);
  if (bin_count) {
    if (type == DT_F32 || type == DT_F64) *bin_count = (size_t)-1;
    else *bin_count = (size_t)(uint64_t)diff_at(&mx8, type, 0, &mn8) + 1;
  }
  const size_t max_out = out_width == 1 ? 255 : 65535;
  const double f = quant_factor(type, max_out, &mn8, &mx8);
  const double fh = quant_factor(type, hsz - 1, &mn8, &mx8);
  const int changed = f != 1.0 || !min_is_zero(type, &mn8) || kWidth[type] > 2 || kWidth[type] > out_width;
  for (uint64_t i = 0; i < n; ++i) {
    const double d = diff_at(in, type, i, &mn8);
    if (out_width == 1) {
      uint8_t v = (uint8_t)(d * f), h = (uint8_t)(d * fh);
      if (v > (uint8_t)max_out) v = (uint8_t)max_out;
      if (h > (uint8_t)(hsz - 1)) h = (uint8_t)(hsz - 1);
      ((uint8_t*)out)[i] = v;
      hist[h]++;
    } else {
      uint16_t v = (uint16_t)(d * f), h = (uint16_t)(d * fh);
      if (v > (uint16_t)max_out) v = (uint16_t)max_out;
      if (h > (uint16_t)(hsz - 1)) h = (uint16_t)(hsz - 1);
      ((uint16_t*)out)[i] = v;
      hist[h]++;
    }
  }
  *hist_set = 1;
  return changed;
}

/* Process8Bits, AbstrConverter.cpp:75-157 */
static int process_8bits(const void* in, uint64_t n, int sgn, void* out, uint64_t* hist, uint32_t* hist_size) {
  memset(hist, 0, sizeof(uint64_t) * 256);
  *hist_size = 256;
  if (sgn) {
    for (uint64_t i = 0; i < n; ++i) {
      const uint8_t v = (uint8_t)(((const int8_t*)in)[i] + 128);
      ((uint8_t*)out)[i] = v;
      hist[v]++;
    }
    return 1;
  }
  for (uint64_t i = 0; i < n; ++i) hist[((const uint8_t*)in)[i]]++;
  return 0;
}

/* total order of std::map<T,...> keys, on the raw 64-bit images of the values */
static uint32_t g_sort_type;
static int cmp_keys(const void* a, const void* b) {
#define CMP(T) { const T x = *(const T*)a, y = *(const T*)b; return x < y ? -1 : (x > y ? 1 : 0); }
  switch (g_sort_type) {
    case DT_U16: CMP(uint16_t) case DT_U32: CMP(uint32_t) case DT_F32: CMP(float)
    case DT_U64: CMP(uint64_t) default: CMP(double)
  }
#undef CMP
}

/* BinningQuantize<T,U>, Quantize.h:583-686 (T in u16/u32/u64/f32/f64).  *out_bits as iComponentSize. */
static int binning_quantize(const void* in, uint32_t type, uint64_t n, uint32_t out_width, void* out,
                            uint64_t* hist, uint32_t* hist_size, int* hist_set, uint32_t* out_bits) {
  *out_bits = out_width * 8;
  const size_t limit = out_width == 1 ? 256 : 4096;
  const uint32_t w = kWidth[type];
  /* distinct values, kept sorted; bails out as soon as there are more than `limit` */
  uint8_t* keys = (uint8_t*)malloc((limit + 2) * 8);
  size_t nk = 0;
  int possible = 1;
  g_sort_type = type;
  for (uint64_t i = 0; i < n && possible; ++i) {
    const uint8_t* v = (const uint8_t*)in + i * w;
    size_t lo = 0, hi = nk;
    while (lo < hi) { const size_t m = (lo + hi) / 2; if (cmp_keys(keys + m * 8, v) < 0) lo = m + 1; else hi = m; }
    if (lo < nk && cmp_keys(keys + lo * 8, v) == 0) continue;
    memmove(keys + (lo + 1) * 8, keys + lo * 8, (nk - lo) * 8);
    memset(keys + lo * 8, 0, 8);
    memcpy(keys + lo * 8, v, w);
    if (++nk > limit) possible = 0;
  }
  if (!possible) {
    free(keys);
    return quantize_tu(in, type, n, out_width, out, hist, hist_size, hist_set, NULL);
  }
  uint32_t uw = out_width;
  if (nk < 256) { *out_bits = 8; uw = 1; }
  const size_t hs = uw == 1 ? 256 : nk;
  memset(hist, 0, sizeof(uint64_t) * hs);
  for (uint64_t i = 0; i < n; ++i) {
    const uint8_t* v = (const uint8_t*)in + i * w;
    size_t lo = 0, hi = nk;
    while (lo < hi) { const size_t m = (lo + hi) / 2; if (cmp_keys(keys + m * 8, v) < 0) lo = m + 1; else hi = m; }
    if (uw == 1) ((uint8_t*)out)[i] = (uint8_t)lo; else ((uint16_t*)out)[i] = (uint16_t)lo;
    if (lo < hs) hist[lo]++;
  }
  free(keys);
  *hist_size = (uint32_t)hs;
  *hist_set = 1;
  return 1;
}

/* quantize(), RAWConverter.cpp:196-352.  `out` needs room for 2*n bytes, `hist` for 4096 bins. */
int orc_quantize(const void* in, uint32_t type, uint64_t n, int to8bit, void* out, uint64_t* hist, orc_quant_result* r) {
  if (!in || !out || !hist || !r || type > DT_F64 || n == 0) return 1;
  uint32_t bits = kWidth[type] * 8, hs = 0;
  int set = 0, target = 0;
  size_t bc = 0;
  const int sgn = is_signed_type(type), fp = type == DT_F32 || type == DT_F64;
  if (to8bit && bits > 8) {                       /* AbstrConverter::QuantizeTo8Bit */
    target = quantize_tu(in, type, n, 1, out, hist, &hs, &set, NULL);
    bits = 8;
  } else if (bits == 8) {
    target = process_8bits(in, n, sgn, out, hist, &hs);
    set = 1;
  } else if (fp) {
    target = binning_quantize(in, type, n, 2, out, hist, &hs, &set, &bits);
  } else if (sgn) {
    target = quantize_tu(in, type, n, 2, out, hist, &hs, &set, NULL);
    bits = 16;
  } else {
    target = quantize_tu(in, type, n, 2, out, hist, &hs, &set, &bc);
    if (bc > 0 && bc <= 256) target = binning_quantize(in, type, n, 1, out, hist, &hs, &set, &bits);
    else bits = 16;
  }
  r->out_bits = bits;
  r->generated = (uint32_t)target;
  r->hist_size = set ? hs : 0;
  r->bin_count = (uint32_t)(bc > 0xFFFFFFFFu ? 0xFFFFFFFFu : bc);
  return 0;
}

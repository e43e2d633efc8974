// B5 -- input quantisation as streaming kernels (sm_100a).
//
// Replaces the file-to-file passes the reference runs before bricking (paths under
// /root/reference/Tuvok/IO/):
//   quantize()              RAWConverter.cpp:196-352     dispatcher over width / signedness / float
//   Quantize<T,U>           Quantize.h:429-578           min/max, bias to 0, scale (never stretch
//                                                        integers), 256- / 4096-bin histogram
//   QuantizationFactor      Quantize.h:50-72
//   BinningQuantize<T,U>    Quantize.h:583-686           <= 256 / 4096 distinct values -> rank remap
//   ApplyMapping<T,U>       Quantize.h:348-421
//   Process8Bits            AbstrConverter.cpp:75-157    signed char + 128, 256-bin histogram
//   QuantizeTo8Bit          AbstrConverter.cpp:167-246
//
// Three kinds of pass over the volume, each one read of the input at 16 bytes per thread:
//   minmax_kernel    per-CTA min / max in the input's own type (the host folds <= 1184 partials)
//   distinct_kernel  open-addressing set of the distinct values in L2 (bails out past 4096)
//   map_kernel       (v - min) * factor in fp64 with the reference's conversions, or the rank of v
//                    among the distinct values; writes U and counts the histogram in shared memory
// The decisions between the passes (early-out, binning possible, factors) are host logic exactly as
// in the reference; they need min/max on the host, so every pass ends with a stream synchronise.
#include "tvk_internal.h"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <type_traits>

namespace {

const uint32_t kWidth[10] = {1, 1, 2, 2, 4, 4, 4, 8, 8, 8};
inline bool type_signed(uint32_t t) { return t == TVK_DT_I8 || t == TVK_DT_I16 || t == TVK_DT_I32 || t == TVK_DT_I64 || t == TVK_DT_F32 || t == TVK_DT_F64; }

template <typename T> struct Lim;
#define TVK_LIM(T, LO, HI) template <> struct Lim<T> { static __host__ __device__ T lo() { return LO; } static __host__ __device__ T hi() { return HI; } };
TVK_LIM(int8_t, -128, 127) TVK_LIM(uint8_t, 0, 255) TVK_LIM(int16_t, -32768, 32767) TVK_LIM(uint16_t, 0, 65535)
TVK_LIM(int32_t, INT32_MIN, INT32_MAX) TVK_LIM(uint32_t, 0u, UINT32_MAX) TVK_LIM(int64_t, INT64_MIN, INT64_MAX)
TVK_LIM(uint64_t, 0ull, UINT64_MAX) TVK_LIM(float, -INFINITY, INFINITY) TVK_LIM(double, -INFINITY, INFINITY)
#undef TVK_LIM

// one 16-byte load = VEC elements of T
template <typename T> struct Vec16 {
  static constexpr int N = 16 / (int)sizeof(T);
  T v[N];
};
template <typename T> __device__ __forceinline__ Vec16<T> load_vec(const T* p) {
  Vec16<T> r;
  *reinterpret_cast<uint4*>(r.v) = __ldg(reinterpret_cast<const uint4*>(p));
  return r;
}

// ---------------------------------------------------------------- min / max
template <typename T>
__global__ void __launch_bounds__(256) minmax_kernel(const T* __restrict__ in, uint64_t n, T* __restrict__ part_min, T* __restrict__ part_max) {
  constexpr int N = Vec16<T>::N;
  T lo = Lim<T>::hi(), hi = Lim<T>::lo();
  const uint64_t nvec = n / N;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (uint64_t)gridDim.x * blockDim.x) {
    const Vec16<T> q = load_vec(in + i * N);
#pragma unroll
    for (int e = 0; e < N; ++e) { lo = q.v[e] < lo ? q.v[e] : lo; hi = q.v[e] > hi ? q.v[e] : hi; }
  }
  for (uint64_t i = nvec * N + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const T v = in[i];
    lo = v < lo ? v : lo; hi = v > hi ? v : hi;
  }
  __shared__ T slo[256], shi[256];
  slo[threadIdx.x] = lo; shi[threadIdx.x] = hi;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) {
      if (slo[threadIdx.x + s] < slo[threadIdx.x]) slo[threadIdx.x] = slo[threadIdx.x + s];
      if (shi[threadIdx.x + s] > shi[threadIdx.x]) shi[threadIdx.x] = shi[threadIdx.x + s];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) { part_min[blockIdx.x] = slo[0]; part_max[blockIdx.x] = shi[0]; }
}

// ---------------------------------------------------------------- distinct values
// Set of 64-bit key images (the value's bytes, zero extended; -0.0 folded into +0.0 because
// std::map<float> treats them as one key).  Entry 0 = empty, else key + 1; the one key whose image
// is ~0 cannot be stored that way and is tracked by a flag.  state[0] = inserted count,
// state[1] = saw ~0, state[2] = "more than `limit` distinct values, stop".
#define TVK_SET_SLOTS 16384u
template <typename T> __device__ __forceinline__ unsigned long long key_image(T v) {
  if (sizeof(T) == 4 && v == (T)0) return 0ull;         // (only float / double can have two zeros)
  if (sizeof(T) == 8 && v == (T)0) return 0ull;
  unsigned long long k = 0ull;
  memcpy(&k, &v, sizeof(T));
  return k;
}
template <typename T>
__global__ void __launch_bounds__(256) distinct_kernel(const T* __restrict__ in, uint64_t n, unsigned long long* __restrict__ set,
                                                        unsigned int* __restrict__ state, unsigned int limit) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    if (*reinterpret_cast<volatile unsigned int*>(state + 2)) return;
    const unsigned long long k = key_image<T>(in[i]);
    if (k == ~0ull) { state[1] = 1u; continue; }
    const unsigned long long e = k + 1ull;
    unsigned int slot = (unsigned int)((k * 0x9E3779B97F4A7C15ull) >> 50) & (TVK_SET_SLOTS - 1u);
    for (;;) {
      unsigned long long cur = *reinterpret_cast<volatile unsigned long long*>(set + slot);
      if (cur == e) break;
      if (cur == 0ull) {
        cur = atomicCAS(set + slot, 0ull, e);
        if (cur == 0ull) {
          if (atomicAdd(state, 1u) + 1u > limit) atomicExch(state + 2, 1u);
          break;
        }
        if (cur == e) break;
      }
      slot = (slot + 1u) & (TVK_SET_SLOTS - 1u);
      if (*reinterpret_cast<volatile unsigned int*>(state + 2)) return;
    }
  }
}

// ---------------------------------------------------------------- map + histogram
struct MapParams {
  double f, fh;              // fQuantFact, fQuantFactHist
  unsigned long long mn;     // minimum, as the value's byte image
  uint32_t max_out, hist_last;   // max_output_val, hist_size - 1
  uint32_t mode;             // 0 scale, 1 rank (binning), 2 signed char + 128, 3 histogram only (uint8)
  uint32_t nkeys;            // rank mode: number of sorted keys
  uint32_t hist_bins;
};

// (in - min) exactly as the reference's expression evaluates it, then widened to double
template <typename T> __device__ __forceinline__ double diff_to_double(T v, T mn) {
  if (sizeof(T) <= 2) return (double)((int)v - (int)mn);
  return (double)(v - mn);
}
template <> __device__ __forceinline__ double diff_to_double<int32_t>(int32_t v, int32_t mn) { return (double)((long long)v - (long long)mn); }
template <> __device__ __forceinline__ double diff_to_double<int64_t>(int64_t v, int64_t mn) { return (double)(long long)((unsigned long long)v - (unsigned long long)mn); }
template <> __device__ __forceinline__ double diff_to_double<float>(float v, float mn) { return (double)__fsub_rn(v, mn); }
template <> __device__ __forceinline__ double diff_to_double<double>(double v, double mn) { return __dsub_rn(v, mn); }

template <typename T> __device__ __forceinline__ uint32_t rank_of(const T* __restrict__ keys, uint32_t nk, T v) {
  uint32_t lo = 0, hi = nk;
  while (lo < hi) { const uint32_t m = (lo + hi) >> 1; if (keys[m] < v) lo = m + 1; else hi = m; }
  return lo;
}

template <typename T, typename U>
__device__ __forceinline__ U map_one(T v, T mn, const MapParams& p, const T* __restrict__ keys, uint32_t& hbin) {
  if (p.mode == 0) {
    const double d = diff_to_double<T>(v, mn);
    // static_cast<U>(double): truncation; the products never leave [0, max_out + 1)
    const uint32_t q = min(__double2uint_rz(__dmul_rn(d, p.f)), p.max_out);
    hbin = min(__double2uint_rz(__dmul_rn(d, p.fh)), p.hist_last);
    return (U)q;
  }
  if (p.mode == 1) { const uint32_t r = rank_of<T>(keys, p.nkeys, v); hbin = r; return (U)r; }
  if (p.mode == 2) { const uint32_t r = (uint32_t)((int)v + 128) & 0xFFu; hbin = r; return (U)r; }
  hbin = (uint32_t)v & 0xFFu;
  return (U)hbin;
}

// Histogram in shared memory: <= 4096 bins, replicated per warp while the replicas fit 32 KB.
template <typename T, typename U, bool WRITE>
__global__ void __launch_bounds__(256) map_kernel(const T* __restrict__ in, uint64_t n, U* __restrict__ out, const __grid_constant__ MapParams p,
                                                   const T* __restrict__ keys, unsigned long long* __restrict__ hist) {
  extern __shared__ unsigned int sh[];
  constexpr int N = Vec16<T>::N;
  const uint32_t reps = p.hist_bins <= 1024 ? 8u : 1u;
  for (uint32_t i = threadIdx.x; i < p.hist_bins * reps; i += blockDim.x) sh[i] = 0u;
  __syncthreads();
  unsigned int* mine = sh + (reps == 8u ? (threadIdx.x >> 5) * p.hist_bins : 0u);
  T mn;
  memcpy(&mn, &p.mn, sizeof(T));
  const uint64_t nvec = n / N;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (uint64_t)gridDim.x * blockDim.x) {
    const Vec16<T> q = load_vec(in + i * N);
    U r[N];
#pragma unroll
    for (int e = 0; e < N; ++e) {
      uint32_t hb;
      r[e] = map_one<T, U>(q.v[e], mn, p, keys, hb);
      if (hb < p.hist_bins) atomicAdd(&mine[hb], 1u);
    }
    if (WRITE) {
      if (N * sizeof(U) == 16) *reinterpret_cast<uint4*>(out + i * N) = *reinterpret_cast<const uint4*>(r);
      else if (N * sizeof(U) == 8) *reinterpret_cast<uint2*>(out + i * N) = *reinterpret_cast<const uint2*>(r);
      else if (N * sizeof(U) == 4) *reinterpret_cast<uint32_t*>(out + i * N) = *reinterpret_cast<const uint32_t*>(r);
      else *reinterpret_cast<uint16_t*>(out + i * N) = *reinterpret_cast<const uint16_t*>(r);
    }
  }
  for (uint64_t i = nvec * N + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    uint32_t hb;
    const U r = map_one<T, U>(in[i], mn, p, keys, hb);
    if (hb < p.hist_bins) atomicAdd(&mine[hb], 1u);
    if (WRITE) out[i] = r;
  }
  __syncthreads();
  for (uint32_t b = threadIdx.x; b < p.hist_bins; b += blockDim.x) {
    unsigned int s = 0;
    for (uint32_t r = 0; r < reps; ++r) s += sh[r * p.hist_bins + b];
    if (s) atomicAdd(&hist[b], (unsigned long long)s);
  }
}

struct Quantizer {
  tvk_ctx* c;
  const void* d_in;
  uint64_t n;
  void* d_out;
  uint64_t* hist;              // host, 4096 entries
  unsigned long long* d_hist;  // device, 4096 entries
  int sms;
  float ms = 0.f;

  template <typename T> int minmax(T& mn, T& mx) {
    const unsigned blocks = (unsigned)std::min<uint64_t>((n / Vec16<T>::N + 255) / 256 + 1, (uint64_t)sms * 8);
    T* d_part = nullptr;
    TVK_CHECK_CUDA(c, cudaMalloc(&d_part, sizeof(T) * 2 * blocks));
    minmax_kernel<T><<<blocks, 256, 0, c->stream>>>((const T*)d_in, n, d_part, d_part + blocks);
    c->launches++;
    std::vector<T> h(2 * blocks);
    TVK_CHECK_CUDA(c, cudaMemcpyAsync(h.data(), d_part, sizeof(T) * 2 * blocks, cudaMemcpyDeviceToHost, c->stream));
    TVK_CHECK_CUDA(c, cudaStreamSynchronize(c->stream));
    cudaFree(d_part);
    mn = h[0]; mx = h[blocks];
    for (unsigned i = 1; i < blocks; ++i) { if (h[i] < mn) mn = h[i]; if (h[blocks + i] > mx) mx = h[blocks + i]; }
    return TVK_OK;
  }

  template <typename T, typename U> int map(const MapParams& p, const T* d_keys, bool write) {
    TVK_CHECK_CUDA(c, cudaMemsetAsync(d_hist, 0, sizeof(unsigned long long) * 4096, c->stream));
    const unsigned blocks = (unsigned)std::min<uint64_t>((n / Vec16<T>::N + 255) / 256 + 1, (uint64_t)sms * 8);
    const size_t smem = sizeof(unsigned int) * p.hist_bins * (p.hist_bins <= 1024 ? 8 : 1);
    if (write) map_kernel<T, U, true><<<blocks, 256, smem, c->stream>>>((const T*)d_in, n, (U*)d_out, p, d_keys, d_hist);
    else map_kernel<T, U, false><<<blocks, 256, smem, c->stream>>>((const T*)d_in, n, (U*)d_out, p, d_keys, d_hist);
    c->launches++;
    TVK_CHECK_CUDA(c, cudaGetLastError());
    TVK_CHECK_CUDA(c, cudaMemcpyAsync(hist, d_hist, sizeof(uint64_t) * p.hist_bins, cudaMemcpyDeviceToHost, c->stream));
    TVK_CHECK_CUDA(c, cudaStreamSynchronize(c->stream));
    return TVK_OK;
  }

  // QuantizationFactor, Quantize.h:50-72
  template <typename T> static double factor(size_t max_out, T mn, T mx) {
    const double q = (double)max_out / ((double)mx - (double)mn);
    return std::min(q, 1.0);
  }

  // Quantize<T,U>, Quantize.h:429-578; returns bDataWillbeChanged through *changed
  template <typename T, typename U>
  int quantize_tu(uint32_t type, bool want_bins, int* changed, uint32_t* hist_size, size_t* bin_count) {
    const uint32_t hsz = sizeof(U) == 1 ? 256u : 4096u;
    *changed = 0; *hist_size = 0;
    if (bin_count) *bin_count = 0;
    T mn, mx;
    int rc = minmax<T>(mn, mx);
    if (rc) return rc;
    MapParams p{};
    memcpy(&p.mn, &mn, sizeof(T));
    if (!type_signed(type) && sizeof(T) <= (hsz == 256 ? 1u : 2u) && (double)mx < (double)hsz) {
      if (want_bins) {               // iBinCount = occupied bins of the exact histogram (every value < hsz)
        p.mode = 0; p.mn = 0; p.f = 1.0; p.fh = 1.0; p.max_out = 65535u; p.hist_last = hsz - 1; p.hist_bins = hsz;
        rc = map<T, U>(p, nullptr, false);
        if (rc) return rc;
        for (uint32_t b = 0; b < hsz; ++b) if (hist[b]) (*bin_count)++;
      }
      return TVK_OK;                 // data does not need processing; no histogram is handed on
    }
    if (bin_count) *bin_count = (type == TVK_DT_F32 || type == TVK_DT_F64) ? (size_t)-1 : (size_t)((double)mx - (double)mn) + 1;
    p.max_out = sizeof(U) == 1 ? 255u : 65535u;
    p.hist_last = hsz - 1; p.hist_bins = hsz; p.mode = 0;
    if constexpr (std::is_same<T, float>::value) {
      p.f = (double)((float)(size_t)p.max_out / (mx - mn));      // size_t / float -> float (Quantize.h:61-64)
      p.fh = (double)((float)(size_t)(hsz - 1) / (mx - mn));
    } else if constexpr (std::is_same<T, double>::value) {
      p.f = (double)p.max_out / (mx - mn);
      p.fh = (double)(hsz - 1) / (mx - mn);
    } else {
      p.f = factor<T>(p.max_out, mn, mx);
      p.fh = factor<T>(hsz - 1, mn, mx);
    }
    *changed = p.f != 1.0 || !(mn == (T)0) || sizeof(T) > 2 || sizeof(T) > sizeof(U);
    rc = map<T, U>(p, nullptr, *changed != 0);
    if (rc) return rc;
    *hist_size = hsz;
    return TVK_OK;
  }

  // BinningQuantize<T,U>, Quantize.h:583-686
  template <typename T, typename U>
  int binning(uint32_t type, int* generated, uint32_t* hist_size, uint32_t* out_bits) {
    *out_bits = sizeof(U) * 8;
    const uint32_t limit = sizeof(U) == 1 ? 256u : 4096u;
    unsigned long long* d_set = nullptr;
    unsigned int* d_state = nullptr;
    TVK_CHECK_CUDA(c, cudaMalloc(&d_set, sizeof(unsigned long long) * TVK_SET_SLOTS + 16));
    d_state = reinterpret_cast<unsigned int*>(d_set + TVK_SET_SLOTS);
    TVK_CHECK_CUDA(c, cudaMemsetAsync(d_set, 0, sizeof(unsigned long long) * TVK_SET_SLOTS + 16, c->stream));
    distinct_kernel<T><<<sms * 8, 256, 0, c->stream>>>((const T*)d_in, n, d_set, d_state, limit);
    c->launches++;
    std::vector<unsigned long long> set(TVK_SET_SLOTS + 2);
    TVK_CHECK_CUDA(c, cudaMemcpyAsync(set.data(), d_set, sizeof(unsigned long long) * TVK_SET_SLOTS + 16, cudaMemcpyDeviceToHost, c->stream));
    TVK_CHECK_CUDA(c, cudaStreamSynchronize(c->stream));
    cudaFree(d_set);
    unsigned int state[4];
    memcpy(state, &set[TVK_SET_SLOTS], 16);
    std::vector<T> keys;
    if (!state[2]) {
      for (uint32_t s = 0; s < TVK_SET_SLOTS; ++s) if (set[s]) { const unsigned long long k = set[s] - 1ull; T v; memcpy(&v, &k, sizeof(T)); keys.push_back(v); }
      if (state[1]) { const unsigned long long k = ~0ull; T v; memcpy(&v, &k, sizeof(T)); keys.push_back(v); }
    }
    if (state[2] || keys.size() > limit) {       // too many values: really quantise
      int changed = 0;
      int rc = quantize_tu<T, U>(type, false, &changed, hist_size, nullptr);
      *generated = changed;
      return rc;
    }
    std::sort(keys.begin(), keys.end());
    T* d_keys = nullptr;
    TVK_CHECK_CUDA(c, cudaMalloc(&d_keys, sizeof(T) * keys.size()));
    TVK_CHECK_CUDA(c, cudaMemcpyAsync(d_keys, keys.data(), sizeof(T) * keys.size(), cudaMemcpyHostToDevice, c->stream));
    MapParams p{};
    p.mode = 1; p.nkeys = (uint32_t)keys.size();
    int rc;
    if (keys.size() < 256 || sizeof(U) == 1) {    // ApplyMapping<T, uint8_t>
      if (keys.size() < 256) *out_bits = 8;
      p.hist_bins = 256;
      rc = map<T, uint8_t>(p, d_keys, true);
    } else {
      p.hist_bins = (uint32_t)keys.size();
      rc = map<T, uint16_t>(p, d_keys, true);
    }
    cudaFree(d_keys);
    *hist_size = p.hist_bins;
    *generated = 1;
    return rc;
  }
};

template <typename T>
int run_unsigned(Quantizer& q, uint32_t type, tvk_quantize_result* r) {       // 16/32/64-bit unsigned integers
  int changed = 0;
  uint32_t hs = 0, bits = 16;
  size_t bc = 0;
  int rc = q.quantize_tu<T, uint16_t>(type, true, &changed, &hs, &bc);
  if (rc) return rc;
  if (bc > 0 && bc <= 256) rc = q.binning<T, uint8_t>(type, &changed, &hs, &bits);
  r->out_bits = bits; r->generated = (uint32_t)changed; r->hist_size = hs;
  r->bin_count = (uint32_t)std::min<size_t>(bc, 0xFFFFFFFFu);
  return rc;
}
template <typename T>
int run_signed(Quantizer& q, uint32_t type, tvk_quantize_result* r) {
  int changed = 0;
  uint32_t hs = 0;
  int rc = q.quantize_tu<T, uint16_t>(type, false, &changed, &hs, nullptr);
  r->out_bits = 16; r->generated = (uint32_t)changed; r->hist_size = hs;
  return rc;
}
template <typename T>
int run_float(Quantizer& q, uint32_t type, tvk_quantize_result* r) {
  int gen = 0;
  uint32_t hs = 0, bits = 16;
  int rc = q.binning<T, uint16_t>(type, &gen, &hs, &bits);
  r->out_bits = bits; r->generated = (uint32_t)gen; r->hist_size = hs;
  return rc;
}
template <typename T>
int run_to8(Quantizer& q, uint32_t type, tvk_quantize_result* r) {
  int changed = 0;
  uint32_t hs = 0;
  int rc = q.quantize_tu<T, uint8_t>(type, false, &changed, &hs, nullptr);
  r->out_bits = 8; r->generated = (uint32_t)changed; r->hist_size = hs;
  return rc;
}

}  // namespace

extern "C" int tvk_quantize(tvk_ctx* c, uint32_t type, uint64_t n, const void* src, int src_on_device, int to8bit,
                            void* dst_device, uint64_t* hist, tvk_quantize_result* r) {
  if (!c || !src || !dst_device || !hist || !r || type > TVK_DT_F64 || n == 0) return TVK_ERR_INVALID;
  TVK_CHECK_CUDA(c, cudaSetDevice(c->device));
  memset(r, 0, sizeof(*r));
  Quantizer q{};
  q.c = c; q.n = n; q.d_out = dst_device; q.hist = hist;
  TVK_CHECK_CUDA(c, cudaDeviceGetAttribute(&q.sms, cudaDevAttrMultiProcessorCount, c->device));
  void* d_tmp = nullptr;
  if (src_on_device) {
    if (reinterpret_cast<uintptr_t>(src) & 15) TVK_FAIL(c, TVK_ERR_INVALID, "device input must be 16-byte aligned");
    q.d_in = src;
  } else {
    TVK_CHECK_CUDA(c, cudaMalloc(&d_tmp, n * kWidth[type]));
    TVK_CHECK_CUDA(c, cudaMemcpyAsync(d_tmp, src, n * kWidth[type], cudaMemcpyHostToDevice, c->stream));
    q.d_in = d_tmp;
  }
  if (reinterpret_cast<uintptr_t>(dst_device) & 15) { cudaFree(d_tmp); TVK_FAIL(c, TVK_ERR_INVALID, "device output must be 16-byte aligned"); }
  if (cudaMalloc(&q.d_hist, sizeof(unsigned long long) * 4096) != cudaSuccess) { cudaFree(d_tmp); TVK_FAIL(c, TVK_ERR_NOMEM, "cudaMalloc"); }
  cudaEventRecord(c->bev0, c->stream);
  int rc;
  const uint32_t bits = kWidth[type] * 8;
  if (to8bit && bits > 8) {                      // AbstrConverter::QuantizeTo8Bit
    switch (type) {
      case TVK_DT_I16: rc = run_to8<int16_t>(q, type, r); break;
      case TVK_DT_U16: rc = run_to8<uint16_t>(q, type, r); break;
      case TVK_DT_I32: rc = run_to8<int32_t>(q, type, r); break;
      case TVK_DT_U32: rc = run_to8<uint32_t>(q, type, r); break;
      case TVK_DT_F32: rc = run_to8<float>(q, type, r); break;
      case TVK_DT_I64: rc = run_to8<int64_t>(q, type, r); break;
      case TVK_DT_U64: rc = run_to8<uint64_t>(q, type, r); break;
      default: rc = run_to8<double>(q, type, r); break;
    }
  } else {
    switch (type) {
      case TVK_DT_I8: case TVK_DT_U8: {          // AbstrConverter::Process8Bits
        MapParams p{};
        p.mode = type == TVK_DT_I8 ? 2 : 3; p.hist_bins = 256;
        rc = type == TVK_DT_I8 ? q.map<int8_t, uint8_t>(p, nullptr, true) : q.map<uint8_t, uint8_t>(p, nullptr, false);
        r->out_bits = 8; r->generated = type == TVK_DT_I8; r->hist_size = 256;
        break;
      }
      case TVK_DT_I16: rc = run_signed<int16_t>(q, type, r); break;
      case TVK_DT_U16: rc = run_unsigned<uint16_t>(q, type, r); break;
      case TVK_DT_I32: rc = run_signed<int32_t>(q, type, r); break;
      case TVK_DT_U32: rc = run_unsigned<uint32_t>(q, type, r); break;
      case TVK_DT_F32: rc = run_float<float>(q, type, r); break;
      case TVK_DT_I64: rc = run_signed<int64_t>(q, type, r); break;
      case TVK_DT_U64: rc = run_unsigned<uint64_t>(q, type, r); break;
      default: rc = run_float<double>(q, type, r); break;
    }
  }
  cudaEventRecord(c->bev1, c->stream);
  cudaStreamSynchronize(c->stream);
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, c->bev0, c->bev1) == cudaSuccess) r->device_ms = ms;
  cudaFree(q.d_hist);
  cudaFree(d_tmp);
  if (!r->hist_size) memset(hist, 0, sizeof(uint64_t) * 4096);
  return rc;
}

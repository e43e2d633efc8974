// Probe (B200, through gpurun): how does the texture unit convert UNORM8 / UNORM16 texels to fp32 in
// cudaReadModeNormalizedFloat -- v / max correctly rounded, or v * (1/max)?  Decides whether K1 may let the hardware
// do the integer -> float conversion of its gathered texels (the oracle then states the same formula).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o unorm_probe unorm_probe.cu && ./unorm_probe
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("{\"error\": \"%s: %s\"}\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

__device__ __forceinline__ float4 gather4f(cudaTextureObject_t t, uint32_t layer, float x, float y) {
  float4 r;
  asm("tld4.r.a2d.v4.f32.f32 {%0, %1, %2, %3}, [%4, {%5, %6, %7, %7}];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(t), "r"(layer), "f"(x), "f"(y));
  return r;
}
template <typename T> __global__ void fill(cudaSurfaceObject_t s, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) surf2DLayeredwrite((T)i, s, (int)((i & 255u) * sizeof(T)), (int)(i >> 8), 0);
}
__global__ void readback(cudaTextureObject_t t, uint32_t n, float* out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 g = gather4f(t, 0, (float)(i & 255u) + 1.0f, (float)(i >> 8) + 1.0f);   // .w = texel (x, y)
  out[i] = g.w;
}
template <typename T> int run(const char* name, uint32_t n, float maxv) {
  cudaChannelFormatDesc fd = cudaCreateChannelDesc(8 * sizeof(T), 0, 0, 0, cudaChannelFormatKindUnsigned);
  cudaArray_t arr = nullptr;
  CK(cudaMalloc3DArray(&arr, &fd, make_cudaExtent(256, (n + 255) / 256 + 1, 2), cudaArrayLayered | cudaArraySurfaceLoadStore));
  cudaResourceDesc rd{}; rd.resType = cudaResourceTypeArray; rd.res.array.array = arr;
  cudaSurfaceObject_t surf = 0; CK(cudaCreateSurfaceObject(&surf, &rd));
  fill<T><<<(n + 255) / 256, 256>>>(surf, n);
  cudaTextureDesc td{}; td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
  td.filterMode = cudaFilterModePoint; td.readMode = cudaReadModeNormalizedFloat; td.normalizedCoords = 0;
  cudaTextureObject_t tex = 0; CK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
  float* d = nullptr; CK(cudaMalloc(&d, 4 * n));
  readback<<<(n + 255) / 256, 256>>>(tex, n, d);
  std::vector<float> h(n); CK(cudaMemcpy(h.data(), d, 4 * n, cudaMemcpyDeviceToHost));
  uint32_t bad_div = 0, bad_mul = 0; const float inv = 1.0f / maxv;
  for (uint32_t v = 0; v < n; ++v) {
    const float a = (float)v / maxv, b = (float)v * inv;
    if (memcmp(&a, &h[v], 4)) ++bad_div;
    if (memcmp(&b, &h[v], 4)) ++bad_mul;
  }
  printf("{\"format\": \"%s\", \"values\": %u, \"differs_from_v_div_max\": %u, \"differs_from_v_times_inv_max\": %u}\n", name, n, bad_div, bad_mul);
  return 0;
}
int main() { CK(cudaSetDevice(0)); if (run<uint8_t>("unorm8", 256, 255.0f)) return 1; return run<uint16_t>("unorm16", 65536, 65535.0f); }

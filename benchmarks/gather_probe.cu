// Probe (run on a B200 through gpurun): can a LAYERED 2-D CUDA array be read with the four-texel gather
// instruction (PTX tld4 with .a2d geometry), and how fast is "two gathers + fp32 lerp" against the hardware
// trilinear fetch of a 3-D array?  The CUDA runtime documents texture gather for plain 2-D arrays only; PTX has
// the layered form.  Decides the 1x-footprint brick pool layout of K1.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_probe gather_probe.cu && ./gather_probe
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("{\"error\": \"%s: %s\"}\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

__device__ __forceinline__ uint4 tld4_a2d(cudaTextureObject_t t, uint32_t layer, float x, float y) {
  uint4 r;
  asm volatile("tld4.r.a2d.v4.u32.f32 {%0, %1, %2, %3}, [%4, {%5, %6, %7, %7}];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(t), "r"(layer), "f"(x), "f"(y));
  return r;
}

__device__ __forceinline__ uint16_t pattern(uint32_t x, uint32_t y, uint32_t l) { return (uint16_t)((x * 3u + y * 7u + l * 131u) ^ (x * y)); }

__global__ void fill_layered(cudaSurfaceObject_t s, uint32_t w, uint32_t h, uint32_t layers) {
  const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, l = blockIdx.z;
  if (x < w && y < h && l < layers) surf2DLayeredwrite(pattern(x, y, l), s, (int)(x * 2), (int)y, (int)l);
}
__global__ void fill_3d(cudaSurfaceObject_t s, uint32_t w, uint32_t h, uint32_t d) {
  const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, z = blockIdx.z;
  if (x < w && y < h && z < d) surf3Dwrite(pattern(x, y, z), s, (int)(x * 2), (int)y, (int)z);
}

// correctness: gather at the shared corner of texels (i..i+1, j..j+1) of layer l
__global__ void check_gather(cudaTextureObject_t t, uint32_t w, uint32_t h, uint32_t layers, uint32_t n, uint32_t* bad, uint32_t* order) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  uint32_t r = k * 2654435761u;
  const uint32_t i = r % (w - 1); r = r * 1664525u + 1013904223u;
  const uint32_t j = r % (h - 1); r = r * 1664525u + 1013904223u;
  const uint32_t l = r % layers;
  const uint4 g = tld4_a2d(t, l, (float)i + 1.0f, (float)j + 1.0f);
  // expected order of a gather: x = (i, j+1), y = (i+1, j+1), z = (i+1, j), w = (i, j)
  const uint32_t e0 = pattern(i, j + 1, l), e1 = pattern(i + 1, j + 1, l), e2 = pattern(i + 1, j, l), e3 = pattern(i, j, l);
  if (g.x != e0 || g.y != e1 || g.z != e2 || g.w != e3) {
    atomicAdd(bad, 1u);
    if (k == 0) { order[0] = g.x; order[1] = g.y; order[2] = g.z; order[3] = g.w; order[4] = e0; order[5] = e1; order[6] = e2; order[7] = e3; }
  }
}

// throughput: every thread is a ray of an 8x4 pixel tile marching through the volume, half a voxel per step
template <int MODE>
__global__ void __launch_bounds__(128) march(cudaTextureObject_t t, uint32_t w, uint32_t h, uint32_t d, uint32_t steps, float* out) {
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t lane = tid & 31u, tile = tid >> 5;
  const uint32_t tx = tile % (w / 8u - 8u), ty = (tile / (w / 8u - 8u)) % (h / 4u - 8u);
  float x = (float)(tx * 8u + (lane & 7u)) + 0.25f, y = (float)(ty * 4u + (lane >> 3)) + 0.25f, z = 0.75f;
  const float dx = 0.11f, dy = 0.07f, dz = 0.48f;
  float acc = 0.f;
  for (uint32_t s = 0; s < steps; ++s) {
    if (MODE == 0) {
      const float u = x - 0.5f, v = y - 0.5f, q = z - 0.5f;
      const float fu = floorf(u), fv = floorf(v), fq = floorf(q);
      const float a = u - fu, b = v - fv, g = q - fq;
      const uint32_t l = (uint32_t)fq;
      const uint4 g0 = tld4_a2d(t, l, fu + 1.0f, fv + 1.0f), g1 = tld4_a2d(t, l + 1u, fu + 1.0f, fv + 1.0f);
      const float c00 = fmaf(a, (float)g0.z - (float)g0.w, (float)g0.w), c10 = fmaf(a, (float)g0.y - (float)g0.x, (float)g0.x);
      const float c01 = fmaf(a, (float)g1.z - (float)g1.w, (float)g1.w), c11 = fmaf(a, (float)g1.y - (float)g1.x, (float)g1.x);
      const float c0 = fmaf(b, c10 - c00, c00), c1 = fmaf(b, c11 - c01, c01);
      acc += fmaf(g, c1 - c0, c0);
    } else {
      acc += tex3D<float>(t, x, y, z);
    }
    x += dx; y += dy; z += dz;
    if (z > (float)d - 2.0f) { z = 0.75f; }
  }
  out[tid] = acc;
}

int main() {
  int dev = 0;
  CK(cudaSetDevice(dev));
  int ml[3] = {0, 0, 0}, mg[2] = {0, 0};
  cudaDeviceGetAttribute(&ml[0], cudaDevAttrMaxTexture2DLayeredWidth, dev);
  cudaDeviceGetAttribute(&ml[1], cudaDevAttrMaxTexture2DLayeredHeight, dev);
  cudaDeviceGetAttribute(&ml[2], cudaDevAttrMaxTexture2DLayeredLayers, dev);
  cudaDeviceGetAttribute(&mg[0], cudaDevAttrMaxTexture2DGatherWidth, dev);
  cudaDeviceGetAttribute(&mg[1], cudaDevAttrMaxTexture2DGatherHeight, dev);
  printf("{\"max_layered\": [%d, %d, %d], \"max_gather_2d\": [%d, %d]}\n", ml[0], ml[1], ml[2], mg[0], mg[1]);

  const uint32_t W = 2048, H = 2048, D = 1024;     // 8 GB of u16: far beyond L2
  cudaChannelFormatDesc fd = cudaCreateChannelDesc(16, 0, 0, 0, cudaChannelFormatKindUnsigned);
  cudaArray_t arr = nullptr;
  unsigned flags = cudaArrayLayered | cudaArraySurfaceLoadStore | cudaArrayTextureGather;
  cudaError_t e = cudaMalloc3DArray(&arr, &fd, make_cudaExtent(W, H, D), flags);
  printf("{\"layered_with_gather_flag\": \"%s\"}\n", cudaGetErrorString(e));
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    flags = cudaArrayLayered | cudaArraySurfaceLoadStore;
    CK(cudaMalloc3DArray(&arr, &fd, make_cudaExtent(W, H, D), flags));
    printf("{\"layered_without_gather_flag\": \"ok\"}\n");
  }
  cudaResourceDesc rd{};
  rd.resType = cudaResourceTypeArray;
  rd.res.array.array = arr;
  cudaSurfaceObject_t surf = 0;
  CK(cudaCreateSurfaceObject(&surf, &rd));
  fill_layered<<<dim3(W / 256, H, D), 256>>>(surf, W, H, D);
  CK(cudaDeviceSynchronize());
  cudaTextureDesc td{};
  td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
  td.filterMode = cudaFilterModePoint;
  td.readMode = cudaReadModeElementType;
  td.normalizedCoords = 0;
  cudaTextureObject_t tex = 0;
  CK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
  uint32_t *d_bad = nullptr, *d_order = nullptr;
  CK(cudaMalloc(&d_bad, 4)); CK(cudaMalloc(&d_order, 32));
  CK(cudaMemset(d_bad, 0, 4)); CK(cudaMemset(d_order, 0, 32));
  const uint32_t n = 1u << 22;
  check_gather<<<n / 256, 256>>>(tex, W, H, D, n, d_bad, d_order);
  CK(cudaDeviceSynchronize());
  uint32_t bad = 0, order[8];
  CK(cudaMemcpy(&bad, d_bad, 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(order, d_order, 32, cudaMemcpyDeviceToHost));
  printf("{\"gather_checked\": %u, \"mismatches\": %u, \"first_got\": [%u, %u, %u, %u], \"first_expected\": [%u, %u, %u, %u]}\n", n, bad,
         order[0], order[1], order[2], order[3], order[4], order[5], order[6], order[7]);

  // throughput
  const uint32_t threads = 148 * 4 * 128 * 8, steps = 2048;
  float* d_out = nullptr;
  CK(cudaMalloc(&d_out, sizeof(float) * threads));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms = 0.f;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    march<0><<<threads / 128, 128>>>(tex, W, H, D, steps, d_out);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    cudaEventElapsedTime(&ms, e0, e1);
  }
  printf("{\"kernel\": \"2 x tld4.a2d + fp32 lerp\", \"ms\": %.3f, \"gsamples_per_s\": %.1f}\n", ms, (double)threads * steps / ms * 1e-6);
  cudaDestroyTextureObject(tex); cudaDestroySurfaceObject(surf); cudaFreeArray(arr);

  // the hardware-filtered 3-D array, same data and access pattern
  cudaArray_t arr3 = nullptr;
  CK(cudaMalloc3DArray(&arr3, &fd, make_cudaExtent(W, H, D), cudaArraySurfaceLoadStore));
  rd.res.array.array = arr3;
  CK(cudaCreateSurfaceObject(&surf, &rd));
  fill_3d<<<dim3(W / 256, H, D), 256>>>(surf, W, H, D);
  CK(cudaDeviceSynchronize());
  td.filterMode = cudaFilterModeLinear;
  td.readMode = cudaReadModeNormalizedFloat;
  CK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    march<1><<<threads / 128, 128>>>(tex, W, H, D, steps, d_out);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    cudaEventElapsedTime(&ms, e0, e1);
  }
  printf("{\"kernel\": \"tex3D linear (hardware trilinear)\", \"ms\": %.3f, \"gsamples_per_s\": %.1f}\n", ms, (double)threads * steps / ms * 1e-6);
  return 0;
}

// Sort-last direct-send exchange fused into the kernels, over peer memory (the reference is
// single-GPU and has no counterpart).
//
// Every rank owns scan-line stripe `rank` of the final image.  One cudaMalloc per rank (shared with the
// other ranks through CUDA IPC, or by pointer inside one process) holds
//     recv[2][world][rows][W] float4   the partial stripes of all ranks, double buffered by frame parity
//     final[world][rows][W]   float4   the finished image (filled on rank 0 and, stripe-wise, locally)
//     flags[64]               uint32   arrival flags: [parity*16 + src] partial stripe of `src` is complete,
//                                      [32 + parity*16 + src] finished stripe of `src` has landed (rank 0),
//                                      [192] a wait timed out (after the rectangle table: [63] is the landed flag
//                                      of member 15 with 16 members)
//     rects[2*16][4]          uint32   [parity*16 + src] = the pixel rectangle (x0, y0, x1, y1) rank `src` sent this frame:
//                                      the screen rectangle of its box; its partial image is zero everywhere else and those
//                                      pixels are neither marched nor sent
// * K1 (tvk_raycast.cu) stores each finished pixel straight into the owner's recv buffer with a peer
//   store over NVLink while it is still marching other tiles, so the exchange overlaps the render; the
//   last CTA to retire publishes the arrival flag on every peer (fence.sys + flag store).
// * exchange_composite_kernel waits (bounded spin) for the `world` flags of its stripe, blends the partial
//   stripes front to back (same arithmetic as composite_kernel) reading them past L1 (ld.cg), and stores the
//   result into the local final image and into rank 0's; its last CTA raises the "landed" flag on rank 0.
// * exchange_wait_kernel (rank 0) orders the stream behind the arrival of all finished stripes.
// No NCCL call is on this path.  A wait that exceeds TVK_XCHG_TIMEOUT_NS sets flags[192] instead of hanging.
#include <cstdio>
#include <cstdlib>
#include "tvk_internal.h"
#include <cstring>

#define TVK_XCHG_TIMEOUT_NS 4000000000ull
#define TVK_XCHG_ERR_FLAG 192u       // word index after flags[64] and rects[32][4]

static inline uint64_t stripe_px(const ExchangeState& x) { return (uint64_t)x.rows * x.width; }
static inline uint64_t recv_bytes(const ExchangeState& x) { return stripe_px(x) * x.world * sizeof(float4); }
static inline uint64_t region_bytes(const ExchangeState& x) { return 3 * recv_bytes(x) + 1024; }   // + flags[64] + rects[2*16][4] + error flag

static void derive_pointers(ExchangeState& x) {
  for (uint32_t r = 0; r < x.world; ++r) {
    uint8_t* b = static_cast<uint8_t*>(x.base[r]);
    x.flags[r] = b ? reinterpret_cast<uint32_t*>(b + 3 * recv_bytes(x)) : nullptr;
  }
}

static void close_peers(ExchangeState& x) {
  for (uint32_t r = 0; r < x.world; ++r)
    if (x.ipc_opened[r]) { cudaIpcCloseMemHandle(x.base[r]); x.ipc_opened[r] = false; x.base[r] = nullptr; }
}

extern "C" int tvk_exchange_destroy(tvk_ctx* c) {
  if (!c) return TVK_ERR_INVALID;
  cudaSetDevice(c->device);
  ExchangeState& x = c->xchg;
  if (x.world) {
    cudaDeviceSynchronize();
    close_peers(x);
    cudaFree(x.local);
  }
  x = ExchangeState();
  return TVK_OK;
}

extern "C" int tvk_exchange_create(tvk_ctx* c, uint32_t world, uint32_t rank, uint32_t width, uint32_t height, void* ipc_handle64) {
  if (!c || world == 0 || world > TVK_XCHG_MAX_WORLD || rank >= world || !width || !height) return TVK_ERR_INVALID;
  TVK_CHECK_CUDA(c, cudaSetDevice(c->device));
  tvk_exchange_destroy(c);
  ExchangeState& x = c->xchg;
  x.world = world; x.rank = rank; x.width = width; x.height = height;
  x.rows = (height + world - 1) / world;
  TVK_CHECK_CUDA(c, cudaMalloc(&x.local, region_bytes(x)));
  TVK_CHECK_CUDA(c, cudaMemset(x.local, 0, region_bytes(x)));
  x.base[rank] = x.local;
  derive_pointers(x);
  if (ipc_handle64) {
    cudaIpcMemHandle_t h;
    TVK_CHECK_CUDA(c, cudaIpcGetMemHandle(&h, x.local));
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
    memcpy(ipc_handle64, &h, 64);
  }
  return TVK_OK;
}

extern "C" int tvk_exchange_open_peers(tvk_ctx* c, const void* handles) {
  if (!c || !c->xchg.world || !handles) return TVK_ERR_INVALID;
  TVK_CHECK_CUDA(c, cudaSetDevice(c->device));
  ExchangeState& x = c->xchg;
  close_peers(x);
  for (uint32_t r = 0; r < x.world; ++r) {
    if (r == x.rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, static_cast<const uint8_t*>(handles) + 64 * r, 64);
    void* p = nullptr;
    TVK_CHECK_CUDA(c, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    x.base[r] = p; x.ipc_opened[r] = true;
  }
  derive_pointers(x);
  return TVK_OK;
}

extern "C" int tvk_exchange_set_peers(tvk_ctx* c, void* const* bases) {
  if (!c || !c->xchg.world || !bases) return TVK_ERR_INVALID;
  ExchangeState& x = c->xchg;
  close_peers(x);
  for (uint32_t r = 0; r < x.world; ++r) x.base[r] = (r == x.rank) ? x.local : bases[r];
  derive_pointers(x);
  return TVK_OK;
}

// Mixed connection for several exchange members per process (a GPU that renders more than one box): bases[r] != NULL is a
// region already addressable in this process (a sibling context's own region, or a peer mapping another sibling opened);
// the others are opened from their IPC handles.  A mapping is owned -- and closed -- by the context that opened it.
extern "C" int tvk_exchange_connect(tvk_ctx* c, const void* handles, void* const* bases) {
  if (!c || !c->xchg.world || (!handles && !bases)) return TVK_ERR_INVALID;
  TVK_CHECK_CUDA(c, cudaSetDevice(c->device));
  ExchangeState& x = c->xchg;
  close_peers(x);
  for (uint32_t r = 0; r < x.world; ++r) {
    if (r == x.rank) { x.base[r] = x.local; continue; }
    if (bases && bases[r]) { x.base[r] = bases[r]; continue; }
    if (!handles) TVK_FAIL(c, TVK_ERR_INVALID, "no base pointer and no IPC handle for an exchange member");
    cudaIpcMemHandle_t h;
    memcpy(&h, static_cast<const uint8_t*>(handles) + 64 * r, 64);
    void* p = nullptr;
    TVK_CHECK_CUDA(c, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    x.base[r] = p; x.ipc_opened[r] = true;
  }
  derive_pointers(x);
  return TVK_OK;
}

extern "C" int tvk_exchange_peer_base(tvk_ctx* c, uint32_t r, void** base) {
  if (!c || !c->xchg.world || r >= c->xchg.world || !base) return TVK_ERR_INVALID;
  *base = c->xchg.base[r];
  return TVK_OK;
}

extern "C" int tvk_exchange_local_base(tvk_ctx* c, void** base) {
  if (!c || !c->xchg.world || !base) return TVK_ERR_INVALID;
  *base = c->xchg.local;
  return TVK_OK;
}

extern "C" int tvk_exchange_begin_frame(tvk_ctx* c, uint32_t frame_id) {
  if (!c || !c->xchg.world || frame_id == 0) return TVK_ERR_INVALID;
  for (uint32_t r = 0; r < c->xchg.world; ++r) if (!c->xchg.base[r]) TVK_FAIL(c, TVK_ERR_INVALID, "exchange peers are not connected");
  c->xchg.send_frame = frame_id;
  return TVK_OK;
}

// the part of the exchange K1 needs (called by tvk_raycast_subframe)
void tvk_exchange_fill(const tvk_ctx* c, ExchangeSend& s) {
  const ExchangeState& x = c->xchg;
  memset(&s, 0, sizeof(s));
  if (!x.world || !x.send_frame) return;
  s.world = x.world; s.rank = x.rank; s.rows = x.rows; s.frame = x.send_frame;
  s.flag_index = (x.send_frame & 1u) * 16u + x.rank;
  const uint32_t parity = x.send_frame & 1u;
  for (uint32_t r = 0; r < x.world; ++r) {
    // stripe slot `rank` of the owner's receive buffer of this parity
    s.recv[r] = reinterpret_cast<float4*>(static_cast<uint8_t*>(x.base[r]) + parity * recv_bytes(x)) + (uint64_t)x.rank * stripe_px(x);
    s.flag[r] = x.flags[r] + parity * 16u + x.rank;
  }
}

__device__ __forceinline__ bool wait_flag(volatile uint32_t* f, uint32_t want, volatile uint32_t* err) {
  unsigned long long t0 = 0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  while ((int)(*f - want) < 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (t - t0 > TVK_XCHG_TIMEOUT_NS) { *err = 1u; return false; }
    __nanosleep(256);
  }
  return true;
}

struct CompositeArgs {
  const float4* recv;          // [world][rows*W] of this frame's parity (local)
  float4* final_local;         // + rank*rows*W applied
  float4* final_root;          // rank 0's final image (nullptr on rank 0)
  uint32_t* flags_local;
  uint32_t* landed_root;       // rank 0's flags[32 + parity*16 + rank]
  uint32_t* done_counter;
  unsigned long long* debug_ts;  // != nullptr: CTA 0 records globaltimer at start / after the flag wait / at its end
  uint32_t order[TVK_XCHG_MAX_WORLD];
  uint32_t world, parity, frame;
  uint32_t width, first_row;   // image width; first scan line of this stripe
  uint64_t stripe_px, stripe_offset;
};

__global__ void __launch_bounds__(256) exchange_composite_kernel(const __grid_constant__ CompositeArgs a) {
  __shared__ int ok;
  __shared__ uint4 srect[TVK_XCHG_MAX_WORLD];
  if (a.debug_ts && blockIdx.x == 0 && threadIdx.x == 0) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); a.debug_ts[0] = t; }
  if (threadIdx.x == 0) {
    int good = 1;
    for (uint32_t s = 0; s < a.world && good; ++s) good = wait_flag(a.flags_local + a.parity * 16u + s, a.frame, a.flags_local + TVK_XCHG_ERR_FLAG) ? 1 : 0;
    ok = good;
    __threadfence_system();
    if (good)                                    // published before the flags by the sender's last CTA
      for (uint32_t s = 0; s < a.world; ++s) {
        const volatile uint32_t* r = a.flags_local + 64u + 4u * (a.parity * 16u + s);
        srect[s] = make_uint4(r[0], r[1], r[2], r[3]);
      }
  }
  __syncthreads();
  if (a.debug_ts && blockIdx.x == 0 && threadIdx.x == 0) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); a.debug_ts[1] = t; }
  if (ok) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.stripe_px; i += (uint64_t)gridDim.x * blockDim.x) {
      float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
      const uint32_t row = (uint32_t)(i / a.width), px = (uint32_t)(i - (uint64_t)row * a.width), py = a.first_row + row;
      for (uint32_t k = 0; k < a.world; ++k) {
        const uint32_t src = a.order[k];
        const uint4 r = srect[src];
        if (px < r.x || px >= r.z || py < r.y || py >= r.w) continue;                    // outside what `src` sent: zero
        const float4 s = __ldcg(a.recv + (uint64_t)src * a.stripe_px + i);              // written by a peer: never from L1
        const float w = 1.0f - d.w;
        d.x += w * s.x; d.y += w * s.y; d.z += w * s.z; d.w += w * s.w;
      }
      a.final_local[a.stripe_offset + i] = d;
      if (a.final_root) a.final_root[a.stripe_offset + i] = d;
    }
  }
  // ONE system-scope fence per CTA, after the CTA barrier: it is cumulative over the stores the barrier ordered before it
  __syncthreads();
  if (a.debug_ts && blockIdx.x == 0 && threadIdx.x == 0) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); a.debug_ts[2] = t; }
  if (threadIdx.x == 0) {
    __threadfence_system();
    const unsigned prev = atomicAdd(a.done_counter, 1u);
    if (prev == gridDim.x - 1) {
      *a.done_counter = 0u;
      __threadfence_system();
      *reinterpret_cast<volatile uint32_t*>(a.landed_root) = a.frame;
    }
  }
}

__global__ void exchange_wait_kernel(uint32_t* flags, uint32_t world, uint32_t parity, uint32_t frame) {
  if (threadIdx.x < world) wait_flag(flags + 32u + parity * 16u + threadIdx.x, frame, flags + TVK_XCHG_ERR_FLAG);
  __threadfence_system();
}

extern "C" int tvk_exchange_composite(tvk_ctx* c, const uint32_t* order, uint32_t frame_id) {
  if (!c || !c->xchg.world || !order || frame_id == 0) return TVK_ERR_INVALID;
  TVK_CHECK_CUDA(c, cudaSetDevice(c->device));
  ExchangeState& x = c->xchg;
  CompositeArgs a{};
  const uint32_t parity = frame_id & 1u;
  uint8_t* lb = static_cast<uint8_t*>(x.local);
  a.recv = reinterpret_cast<const float4*>(lb + parity * recv_bytes(x));
  a.final_local = reinterpret_cast<float4*>(lb + 2 * recv_bytes(x));
  a.final_root = x.rank == 0 ? nullptr : reinterpret_cast<float4*>(static_cast<uint8_t*>(x.base[0]) + 2 * recv_bytes(x));
  if (getenv("TVK_XCHG_NO_LANDING")) a.final_root = nullptr;          // experiment: leave the image distributed
  a.flags_local = x.flags[x.rank];
  a.landed_root = x.flags[0] + 32u + parity * 16u + x.rank;
  a.done_counter = c->d_tile_counter + 2;
  a.debug_ts = getenv("TVK_XCHG_DEBUG") ? reinterpret_cast<unsigned long long*>(c->d_tile_counter + 4) : nullptr;
  for (uint32_t k = 0; k < x.world; ++k) {
    if (order[k] >= x.world) TVK_FAIL(c, TVK_ERR_INVALID, "bad compositing order");
    a.order[k] = order[k];
  }
  a.world = x.world; a.parity = parity; a.frame = frame_id;
  a.stripe_px = stripe_px(x); a.stripe_offset = (uint64_t)x.rank * stripe_px(x);
  a.width = x.width; a.first_row = x.rank * x.rows;
  int sms = 0;
  TVK_CHECK_CUDA(c, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
  uint64_t blocks = (a.stripe_px + 255) / 256;
  if (blocks > (uint64_t)sms * 4) blocks = (uint64_t)sms * 4;
  exchange_composite_kernel<<<(unsigned)blocks, 256, 0, c->stream>>>(a);
  c->launches++;
  TVK_CHECK_CUDA(c, cudaGetLastError());
  x.send_frame = 0;
  if (a.debug_ts) {
    unsigned long long ts[3] = {0, 0, 0};
    cudaStreamSynchronize(c->stream);
    cudaMemcpy(ts, a.debug_ts, sizeof(ts), cudaMemcpyDeviceToHost);
    fprintf(stderr, "[xchg rank %u frame %u] flag wait %.3f ms, blend %.3f ms (CTA 0)\n", x.rank, frame_id, (ts[1] - ts[0]) * 1e-6, (ts[2] - ts[1]) * 1e-6);
  }
  return TVK_OK;
}

extern "C" int tvk_exchange_final(tvk_ctx* c, uint32_t frame_id, void** final_dptr) {
  if (!c || !c->xchg.world || frame_id == 0 || !final_dptr) return TVK_ERR_INVALID;
  TVK_CHECK_CUDA(c, cudaSetDevice(c->device));
  ExchangeState& x = c->xchg;
  if (x.rank == 0) {
    exchange_wait_kernel<<<1, 32, 0, c->stream>>>(x.flags[0], x.world, frame_id & 1u, frame_id);
    c->launches++;
    TVK_CHECK_CUDA(c, cudaGetLastError());
  }
  *final_dptr = static_cast<uint8_t*>(x.local) + 2 * recv_bytes(x);
  return TVK_OK;
}

extern "C" int tvk_exchange_status(tvk_ctx* c, uint32_t* timed_out) {
  if (!c || !c->xchg.world || !timed_out) return TVK_ERR_INVALID;
  TVK_CHECK_CUDA(c, cudaSetDevice(c->device));
  TVK_CHECK_CUDA(c, cudaStreamSynchronize(c->stream));
  TVK_CHECK_CUDA(c, cudaMemcpy(timed_out, c->xchg.flags[c->xchg.rank] + TVK_XCHG_ERR_FLAG, sizeof(uint32_t), cudaMemcpyDeviceToHost));
  return TVK_OK;
}

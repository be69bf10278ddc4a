// Shared device/host helpers for libsstb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <utility>
#include <vector>

#include "../../include/sstb200.h"

#define SSTB_OK 0
#define SSTB_ERR_CUDA -1
#define SSTB_ERR_ARG -2
#define SSTB_ERR_UNSUPPORTED -3
#define SSTB_ERR_WORKSPACE -4

struct sstb200_ctx {
  int device = 0;
  cudaStream_t stream = 0;
  int num_sms = 148;
  // bump arena for temporaries; grows (with a sync) when a call needs more.
  char* arena = nullptr;
  size_t arena_cap = 0;
  size_t arena_off = 0;
  std::vector<void*> retired;  // old arenas kept alive until destroy (stream-ordered safety)
  std::string err;
  int32_t* pinned_i32 = nullptr;  // small pinned scratch for D2H counters
};

inline int sstb_fail(sstb200_ctx* c, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (c) c->err = buf;
  return code;
}

#define CUDA_TRY(ctx, expr)                                                                     \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess)                                                                      \
      return sstb_fail(ctx, SSTB_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #expr,        \
                       cudaGetErrorString(_e));                                                 \
  } while (0)

#define CHECK_ARG(ctx, cond)                                                                    \
  do {                                                                                          \
    if (!(cond)) return sstb_fail(ctx, SSTB_ERR_ARG, "%s:%d bad argument: %s", __FILE__, __LINE__, #cond); \
  } while (0)

#define LAUNCH_CHECK(ctx) CUDA_TRY(ctx, cudaGetLastError())

// ---- arena ---------------------------------------------------------------
inline void arena_reset(sstb200_ctx* c) { c->arena_off = 0; }

// Reserve: make sure `bytes` are available from offset 0 (called once at the top of an op
// with an upper bound, so that no growth happens mid-op).
inline int arena_reserve(sstb200_ctx* c, size_t bytes) {
  bytes += 4096;
  if (bytes <= c->arena_cap) return SSTB_OK;
  size_t cap = c->arena_cap ? c->arena_cap : (size_t)1 << 24;
  while (cap < bytes) cap *= 2;
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, cap);
  if (e != cudaSuccess) return sstb_fail(c, SSTB_ERR_WORKSPACE, "arena cudaMalloc(%zu) failed: %s", cap, cudaGetErrorString(e));
  if (c->arena) c->retired.push_back(c->arena);
  c->arena = (char*)p;
  c->arena_cap = cap;
  return SSTB_OK;
}

template <typename T>
inline T* arena_alloc(sstb200_ctx* c, size_t n) {
  size_t off = (c->arena_off + 255) & ~(size_t)255;
  size_t bytes = n * sizeof(T);
  if (off + bytes > c->arena_cap) return nullptr;
  c->arena_off = off + bytes;
  return (T*)(c->arena + off);
}

static inline size_t al256(size_t b) { return (b + 255) & ~(size_t)255; }

// ---- programmatic dependent launch (PDL) ------------------------------------------------------------
// Every kernel of the library starts with pdl_wait() (griddepcontrol.wait: block until the producer grid has completed
// and its writes are visible) followed by pdl_launch() (griddepcontrol.launch_dependents: allow the next kernel of the
// stream / graph to be scheduled while this one still runs).  Launches carry the programmatic-stream-serialization
// attribute, so inside the per-frame CUDA graph consecutive kernels are linked by programmatic edges and the launch
// latency of kernel N+1 overlaps the execution of kernel N.  SSTB200_PDL=0 disables the attribute (plain launches).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

inline int sstb_pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SSTB200_PDL");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v;
}

// All kernels of the library ask for the same shared-memory carveout (max shared): consecutive kernels with different
// carveouts force the SMs to drain and re-partition L1/shared between launches, which costs more than any L1 hit here.
inline int sstb_uniform_carveout() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SSTB200_CARVEOUT");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v;
}

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  static void* seen[32];  // per kernel signature & TU; the attribute is sticky per function -> set once, before any capture
  static int nseen = 0;
  if (sstb_uniform_carveout()) {
    bool found = false;
    for (int i = 0; i < nseen; i++) found |= (seen[i] == (void*)kern);
    if (!found) {
      cudaFuncSetAttribute((const void*)kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
      if (nseen < 32) seen[nseen++] = (void*)kern;
    }
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = sstb_pdl_enabled();
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

// ---- device helpers ------------------------------------------------------
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// order-preserving float <-> uint mapping (for atomicMax on floats of any sign)
__device__ __forceinline__ unsigned f2ord(float f) {
  unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

// ---- generic exclusive scan over uint32 values ----------------------------------------------
// Two launches: (A) per-block totals, last-arriving block scans the totals; (B) per-block scan +
// block prefix.  `Load` maps element index -> uint32 value.  Block = 256 threads x 8 items.
#define SCAN_ITEMS 8
#define SCAN_THREADS 256
#define SCAN_TILE (SCAN_ITEMS * SCAN_THREADS)

struct LoadU32 {
  const uint32_t* p;
  __device__ __forceinline__ uint32_t operator()(size_t i) const { return p[i]; }
};
struct LoadPopc {
  const uint32_t* p;
  __device__ __forceinline__ uint32_t operator()(size_t i) const { return __popc(p[i]); }
};

__device__ __forceinline__ uint32_t block_exclusive_scan_256(uint32_t v, uint32_t* total, uint32_t* sh /*[9]*/) {
  // inclusive warp scan
  uint32_t x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane_id() >= o) x += y;
  }
  int w = threadIdx.x >> 5;
  if (lane_id() == 31) sh[w] = x;
  __syncthreads();
  if (w == 0) {
    uint32_t s = (lane_id() < 8) ? sh[lane_id()] : 0;
    uint32_t t = s;
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
      uint32_t y = __shfl_up_sync(0xffffffffu, t, o);
      if (lane_id() >= o) t += y;
    }
    if (lane_id() < 8) sh[lane_id()] = t - s;  // exclusive warp offsets
    if (lane_id() == 7) sh[8] = t;
  }
  __syncthreads();
  uint32_t res = x - v + sh[w];
  *total = sh[8];
  __syncthreads();
  return res;
}

// n may be read from device memory (n_dev != nullptr) so that data-dependent sizes never sync.
template <typename Load>
__global__ void __launch_bounds__(SCAN_THREADS) scan_phaseA(Load load, size_t n_host, const int32_t* n_dev,
                                                            uint32_t* block_sums, uint32_t* block_prefix,
                                                            uint32_t* ticket, uint32_t* total_out) {
  pdl_wait();
  pdl_launch();
  __shared__ uint32_t sh[9];
  __shared__ bool is_last;
  size_t n = n_dev ? (size_t)max(*n_dev, 0) : n_host;
  size_t nblk_needed = (n + SCAN_TILE - 1) / SCAN_TILE;
  size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
  uint32_t s = 0;
  if ((size_t)blockIdx.x < nblk_needed) {
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
      size_t k = base + i;
      if (k < n) s += load(k);
    }
  }
  uint32_t tot;
  block_exclusive_scan_256(s, &tot, sh);
  if (threadIdx.x == 0) {
    block_sums[blockIdx.x] = tot;
    __threadfence();
    uint32_t t = atomicAdd(ticket, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  // last block: scan block_sums[0..gridDim.x) with a running carry
  uint32_t carry = 0;
  for (uint32_t b0 = 0; b0 < gridDim.x; b0 += SCAN_THREADS) {
    uint32_t b = b0 + threadIdx.x;
    uint32_t v = (b < gridDim.x) ? ((volatile uint32_t*)block_sums)[b] : 0;
    uint32_t t2;
    uint32_t ex = block_exclusive_scan_256(v, &t2, sh);
    if (b < gridDim.x) block_prefix[b] = carry + ex;
    carry += t2;
  }
  if (threadIdx.x == 0) {
    *total_out = carry;
    *ticket = 0;  // self-reset so the buffer can be reused by the next scan on the stream
  }
}

template <typename Load>
__global__ void __launch_bounds__(SCAN_THREADS) scan_phaseB(Load load, size_t n_host, const int32_t* n_dev,
                                                            const uint32_t* block_prefix, uint32_t* out,
                                                            bool write_total_at_n) {
  pdl_wait();
  pdl_launch();
  __shared__ uint32_t sh[9];
  size_t n = n_dev ? (size_t)max(*n_dev, 0) : n_host;
  size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
  if ((size_t)blockIdx.x * SCAN_TILE > n) return;
  uint32_t v[SCAN_ITEMS];
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; i++) {
    size_t k = base + i;
    v[i] = (k < n) ? load(k) : 0;
    s += v[i];
  }
  uint32_t tot;
  uint32_t ex = block_exclusive_scan_256(s, &tot, sh) + block_prefix[blockIdx.x];
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; i++) {
    size_t k = base + i;
    if (k < n) out[k] = ex;
    else if (k == n && write_total_at_n) out[k] = ex;
    ex += v[i];
  }
}

// Single-block variant (one launch) for small inputs: 1024 threads x 8 items per round with a running carry.
template <typename Load>
__global__ void __launch_bounds__(1024) scan_single_block(Load load, size_t n_host, const int32_t* n_dev, uint32_t* out,
                                                          uint32_t* total_out, bool write_total_at_n) {
  pdl_wait();
  pdl_launch();
  __shared__ uint32_t wsum[32];
  __shared__ uint32_t carry_s;
  const size_t n = n_dev ? (size_t)max(*n_dev, 0) : n_host;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (size_t base0 = 0; base0 < n || base0 == 0; base0 += 8192) {
    size_t base = base0 + (size_t)threadIdx.x * 8;
    uint32_t v[8], s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      v[i] = (base + i < n) ? load(base + i) : 0;
      s += v[i];
    }
    uint32_t x = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) wsum[w] = x;
    __syncthreads();
    if (w == 0) {
      uint32_t t = wsum[lane], t0 = t;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, t, o);
        if (lane >= o) t += y;
      }
      wsum[lane] = t - t0;  // exclusive warp offsets
    }
    __syncthreads();
    uint32_t ex = carry_s + wsum[w] + x - s;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      size_t k = base + i;
      if (k < n) out[k] = ex;
      else if (k == n && write_total_at_n) out[k] = ex;
      ex += v[i];
    }
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = ex;  // ex of the last thread after its 8 items == carry + tile total
    __syncthreads();
    if (n == 0) break;
  }
  if (threadIdx.x == 0) {
    *total_out = carry_s;
    // when n is a multiple of 8192 the element at index n was not visited: write the total there
    if (write_total_at_n && n > 0 && (n % 8192) == 0) out[n] = carry_s;
  }
}

struct ScanTemps {
  uint32_t* block_sums;
  uint32_t* block_prefix;
  uint32_t* ticket;  // must be zero before first use (self-resetting afterwards)
};
static inline size_t scan_num_blocks(size_t n_cap) { return (n_cap + 1 + SCAN_TILE - 1) / SCAN_TILE + 0; }

// out has n(+1 if write_total_at_n) entries; total_out receives the grand total.
template <typename Load>
static inline void launch_exclusive_scan(cudaStream_t st, Load load, size_t n_cap, const int32_t* n_dev,
                                         ScanTemps t, uint32_t* out, uint32_t* total_out, bool write_total_at_n) {
  if (n_cap <= ((size_t)1 << 18)) {  // one launch instead of two; the block only walks the actual (device) length
    launch_pdl(scan_single_block<Load>, dim3(1), dim3(1024), (size_t)0, st, load, n_cap, n_dev, out, total_out, write_total_at_n);
    return;
  }
  unsigned nblk = (unsigned)((n_cap + 1 + SCAN_TILE - 1) / SCAN_TILE);
  if (nblk == 0) nblk = 1;
  launch_pdl(scan_phaseA<Load>, dim3(nblk), dim3(SCAN_THREADS), (size_t)(0), st, load, n_cap, n_dev, t.block_sums, t.block_prefix, t.ticket, total_out);
  launch_pdl(scan_phaseB<Load>, dim3(nblk), dim3(SCAN_THREADS), (size_t)(0), st, load, n_cap, n_dev, t.block_prefix, out, write_total_at_n);
}

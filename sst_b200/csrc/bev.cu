// SSTv2.recover_bev (mmdet3d/models/backbones/sst_v2.py:161-196): sparse voxel rows [M, C] -> dense canvas [B, C, ny, nx].
// The reference zero-fills one canvas per sample and index_puts the transposed rows; here ONE pass writes every output byte
// exactly once (HBM-write bound: B*C*ny*nx*4 bytes): a cell -> row map (4 bytes per cell) is filled first, then each block
// transposes the rows of 128 consecutive x cells, 32 channels at a time, through shared memory so that the canvas is written
// as 512-byte warp stores (float4 per lane, streaming); all-empty blocks only stream zeros.
#include <stdarg.h>
#include "common.cuh"

__global__ void bev_map_kernel(const long long* __restrict__ coors, int M, int B, int ny, int nx, int32_t* __restrict__ cell_map,
                               int32_t* __restrict__ err) {
  pdl_wait();
  pdl_launch();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  long long b = coors[(size_t)i * 4], y = coors[(size_t)i * 4 + 2], x = coors[(size_t)i * 4 + 3];
  if (b < 0 || b >= B || y < 0 || y >= ny || x < 0 || x >= nx) {
    *err = 1;
    return;
  }
  cell_map[((size_t)b * ny + y) * nx + x] = i;  // a duplicated cell keeps one of its rows (the reference: the last writer)
}

#define BEV_X 128   // x cells per block
#define BEV_CC 32    // channels per pass
#define BEV_PITCH 132  // floats; == 4 (mod 32): 16-byte row reads by 32 lanes are conflict-free
__global__ void __launch_bounds__(256) bev_fill_kernel(const float* __restrict__ feat, int C, const int32_t* __restrict__ cell_map, int ny,
                                                       int nx, float* __restrict__ out) {
  pdl_wait();
  pdl_launch();
  __shared__ __align__(16) float tile[BEV_CC][BEV_PITCH];  // [channel][cell]
  __shared__ int sRow[BEV_X];
  __shared__ int any;
  const int x0 = blockIdx.x * BEV_X, y = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) any = 0;
  __syncthreads();
  if (threadIdx.x < BEV_X) {
    int x = x0 + threadIdx.x;
    int r = x < nx ? cell_map[((size_t)b * ny + y) * nx + x] : -1;
    sRow[threadIdx.x] = r;
    if (r >= 0) any = 1;
  }
  __syncthreads();
  const bool some = any != 0;
  const int cell = threadIdx.x & (BEV_X - 1), chalf = threadIdx.x >> 7;  // phase A: thread = (cell, half of the channel pass)
  const int myrow = sRow[cell];
  const int x = x0 + lane * 4;
  const size_t cstride = (size_t)ny * nx;
  const bool vec = ((nx & 3) == 0) && x + 3 < nx;
  for (int c0 = 0; c0 < C; c0 += BEV_CC) {
    if (some) {
      __syncthreads();  // previous pass read out
      const float* src = myrow >= 0 ? feat + (size_t)myrow * C + c0 + chalf * (BEV_CC / 2) : nullptr;
#pragma unroll
      for (int k = 0; k < BEV_CC / 2; k++) {
        int ch = chalf * (BEV_CC / 2) + k;
        tile[ch][cell] = (src && c0 + ch < C) ? __ldg(src + k) : 0.f;
      }
      __syncthreads();
    }
    for (int c = warp; c < BEV_CC && c0 + c < C; c += 8) {
      float4 v = some ? *reinterpret_cast<const float4*>(&tile[c][lane * 4]) : make_float4(0.f, 0.f, 0.f, 0.f);
      float* o = out + (((size_t)b * C + c0 + c) * ny + y) * nx + x;
      if (vec) {
        __stcs(reinterpret_cast<float4*>(o), v);
      } else {
        if (x < nx) __stcs(o, v.x);
        if (x + 1 < nx) __stcs(o + 1, v.y);
        if (x + 2 < nx) __stcs(o + 2, v.z);
        if (x + 3 < nx) __stcs(o + 3, v.w);
      }
    }
  }
}

extern "C" int sstb200_recover_bev(sstb200_ctx* c, const float* voxel_feat, const int64_t* coors, int M, int C, int B, int ny, int nx,
                                   float* canvas) {
  CHECK_ARG(c, c && M >= 0 && C >= 1 && C <= 1024 && B >= 1 && ny >= 1 && nx >= 1 && canvas);
  CHECK_ARG(c, M == 0 || (voxel_feat && coors));
  CHECK_ARG(c, ny <= 65535 && B <= 65535);
  const size_t cells = (size_t)B * ny * nx;
  arena_reset(c);
  int rc = arena_reserve(c, al256(cells * 4) + 4096);
  if (rc) return rc;
  int32_t* cell_map = arena_alloc<int32_t>(c, cells + 64);
  if (!cell_map) return sstb_fail(c, SSTB_ERR_WORKSPACE, "recover_bev: arena");
  int32_t* err = cell_map + cells;
  CUDA_TRY(c, cudaMemsetAsync(cell_map, 0xFF, cells * 4, c->stream));
  CUDA_TRY(c, cudaMemsetAsync(err, 0, 4, c->stream));
  if (M > 0)
    launch_pdl(bev_map_kernel, dim3((M + 255) / 256), dim3(256), (size_t)0, c->stream, (const long long*)coors, M, B, ny, nx, cell_map, err);
  launch_pdl(bev_fill_kernel, dim3((nx + BEV_X - 1) / BEV_X, ny, B), dim3(256), (size_t)0, c->stream, voxel_feat, C, (const int32_t*)cell_map, ny, nx,
             canvas);
  LAUNCH_CHECK(c);
  CUDA_TRY(c, cudaMemcpyAsync(c->pinned_i32, err, 4, cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  if (c->pinned_i32[0]) return sstb_fail(c, SSTB_ERR_ARG, "recover_bev: a voxel coordinate lies outside [0,B) x [0,ny) x [0,nx)");
  return SSTB_OK;
}

// SSTv2.recover_bev (mmdet3d/models/backbones/sst_v2.py:161-196): sparse voxel rows [M, C] -> dense canvas [B, C, ny, nx].
// The reference zero-fills one canvas per sample and index_puts the transposed rows; here ONE pass writes every output byte
// exactly once (HBM-write bound: B*C*ny*nx*4 bytes): a cell -> row map (4 bytes per cell) is filled first, then each block
// transposes the rows of 32 consecutive x cells through shared memory so that both the row reads (512-B rows) and the canvas
// writes (128-B segments of one channel row) are coalesced; all-empty blocks only stream zeros.
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

#define BEV_X 32
__global__ void __launch_bounds__(256) bev_fill_kernel(const float* __restrict__ feat, int C, const int32_t* __restrict__ cell_map, int ny,
                                                       int nx, float* __restrict__ out) {
  pdl_wait();
  pdl_launch();
  extern __shared__ float tile[];  // [BEV_X][C + 1]
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
  if (some) {
    for (int cx = warp; cx < BEV_X; cx += 8) {
      int r = sRow[cx];
      if (r < 0) continue;
      const float* src = feat + (size_t)r * C;
      for (int c = lane; c < C; c += 32) tile[cx * (C + 1) + c] = src[c];
    }
    __syncthreads();
  }
  const int x = x0 + lane;
  if (x >= nx) return;
  const bool occ = some && sRow[lane] >= 0;
  float* o = out + (((size_t)b * C) * ny + y) * nx + x;
  const size_t cstride = (size_t)ny * nx;
  for (int c = warp; c < C; c += 8) __stcs(o + (size_t)c * cstride, occ ? tile[lane * (C + 1) + c] : 0.f);
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
  const size_t smem = (size_t)BEV_X * (C + 1) * 4;
  static size_t attr = 0;
  if (smem > 48 * 1024 && smem > attr) {
    CUDA_TRY(c, cudaFuncSetAttribute(bev_fill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  launch_pdl(bev_fill_kernel, dim3((nx + BEV_X - 1) / BEV_X, ny, B), dim3(256), smem, c->stream, voxel_feat, C, (const int32_t*)cell_map, ny, nx,
             canvas);
  LAUNCH_CHECK(c);
  CUDA_TRY(c, cudaMemcpyAsync(c->pinned_i32, err, 4, cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  if (c->pinned_i32[0]) return sstb_fail(c, SSTB_ERR_ARG, "recover_bev: a voxel coordinate lies outside [0,B) x [0,ny) x [0,nx)");
  return SSTB_OK;
}

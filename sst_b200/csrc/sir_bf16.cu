// SIRLayer on the tensor cores (precision = BF16): two persistent tcgen05 kernels per layer, tiles of 128 points taken in
// group (CSR) order so that the segmented max-pool is an in-tile scan with a handful of atomics per tile.
//   models/voxel_encoders/voxel_encoder.py:696-764 (SIRLayer.forward), ops/sst/sst_ops.py:334-361 (build_mlp)
//
//   kernel A (per tile):  rel-MLP layers 1,2 in registers (thread per point)            3 -> R1 -> R2, LN + act each
//                         rel-MLP layer 3 as UMMA  [128,R2] x [cin,R2]^T -> TMEM        LN + act in the epilogue,
//                         x0 = [xyz/normalizer || feats] * rel  -> bf16 A operand       (gating fused into that epilogue)
//                         layer 0 as UMMA [128,KP] x [128,KP]^T -> TMEM                  LN + act -> p0 (bf16, CSR order) + segmax g0
//   (G x 128 x 128 fp32 GEMM: gterm = g0 . W1b^T - the pooled half of layer 1's input, added per point by group id)
//   kernel B (per tile):  layer 1 as UMMA [128,128] x [128,128]^T                        + gterm[group], LN + act -> out (fp32,
//                                                                                           original point order) + segmax g1
// Operands are bf16 (fp32 accumulate in TMEM); LayerNorm, activations, gating, pooling and every output stay fp32.
// The [N, 2*C0] concat of the reference is never formed; weights are converted to bf16 operands once per CTA.
#include <stdarg.h>
#include "index.cuh"
#include "sra.cuh"
#include "umma.cuh"

namespace {

constexpr int ST = 128;   // points per tile
constexpr int SC = 128;   // C0 == C1 == 128 on this path
constexpr int SPITCH = SC + 1;

struct SirDev {
  int cin, rel_in, act;
  float eps, rel_dist_scaler, nz[3];
  const float *rw[3], *rg[3], *rb[3];  // rel-MLP: weights [out,in], LayerNorm weight / bias
  const float *w0, *g0, *b0;           // vfe layer 0: [SC, cin]
  const float *w1, *g1, *b1;           // vfe layer 1: [SC, 2*SC]
};

__device__ __forceinline__ float gelu_as(float x) {  // exact-form GELU, erf by Abramowitz-Stegun 7.1.26 (|err| <= 1.5e-7)
  const float z = fabsf(x) * 0.70710678118654752440f;
  const float t = __frcp_rn(fmaf(0.3275911f, z, 1.0f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float e = 1.0f - p * t * __expf(-z * z);
  return 0.5f * x * (1.0f + copysignf(e, x));
}
__device__ __forceinline__ float act_f(float x, int act) { return act == 2 ? gelu_as(x) : fmaxf(x, 0.f); }

// fp32 weight block W[r, col0 + k] (r < nrows, k < ncols) -> bf16 K-major SWIZZLE_128B operand of NR rows x KPAD columns
__device__ __forceinline__ void stage_weight(uint8_t* dst, int NR, int KPAD, const float* __restrict__ W, int ld, int col0, int nrows,
                                             int ncols) {
  const int pieces = NR * (KPAD / 8);
  for (int idx = threadIdx.x; idx < pieces; idx += blockDim.x) {
    const int r = idx / (KPAD / 8), j = idx % (KPAD / 8);
    float f[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
      int k = j * 8 + i;
      f[i] = (r < nrows && k < ncols) ? __ldg(W + (size_t)r * ld + col0 + k) : 0.f;
    }
    int4 q;
    q.x = (int)pack_bf16(f[0], f[1]);
    q.y = (int)pack_bf16(f[2], f[3]);
    q.z = (int)pack_bf16(f[4], f[5]);
    q.w = (int)pack_bf16(f[6], f[7]);
    const int c = j >> 3, jj = j & 7;
    *reinterpret_cast<int4*>(dst + (size_t)c * NR * 128 + r * 128 + ((jj ^ (r & 7)) << 4)) = q;
  }
}

// column-wise segmented max over the rows of a tile (rows sorted by group): [ST][SPITCH] fp32 in smem -> order-preserving
// atomicMax on the pooled array (few segments per tile, so ~2 atomics per channel and tile)
__device__ __forceinline__ void tile_segmax_ord(const float* tile, const int* sSeg, int nrow, uint32_t* __restrict__ gord) {
  const int c = threadIdx.x % SC, grp = threadIdx.x / SC, ngrp = blockDim.x / SC;
  const int rows_per = ST / ngrp;
  const int r0 = grp * rows_per, r1 = min(r0 + rows_per, nrow);
  if (r0 >= r1) return;
  int seg = sSeg[r0];
  float m = -INFINITY;
  for (int r = r0; r < r1; r++) {
    int sg = sSeg[r];
    if (sg != seg) {
      atomicMax(&gord[(size_t)seg * SC + c], f2ord(m));
      seg = sg;
      m = -INFINITY;
    }
    m = fmaxf(m, tile[r * SPITCH + c]);
  }
  atomicMax(&gord[(size_t)seg * SC + c], f2ord(m));
}

template <int N>
__device__ __forceinline__ void ln_regs(float* v, float& mean, float& rstd, float eps) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < N; i++) s += v[i];
  mean = s / (float)N;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < N; i++) {
    float t = v[i] - mean;
    q = fmaf(t, t, q);
  }
  rstd = rsqrtf(q / (float)N + eps);
}

// ------------------------------------------------------------------------------------------------------------------------
// kernel A
// ------------------------------------------------------------------------------------------------------------------------
template <int KP, int R1, int R2>
__global__ void __launch_bounds__(256, 1) sir_a_kernel(SirDev d, const float* __restrict__ in_feats, const float* __restrict__ f_cluster,
                                                       const long long* __restrict__ inv, const int32_t* __restrict__ order, int N,
                                                       __nv_bfloat16* __restrict__ p0buf, uint32_t* __restrict__ gord, int pitch_in) {
  pdl_wait();
  pdl_launch();
  static_assert(KP % 64 == 0 && KP <= 192 && R2 % 16 == 0 && R2 <= 64 && R1 <= 32, "shape");
  constexpr int NH = KP / 2;  // columns per epilogue half (rel3); layer 0 uses SC / 2
  extern __shared__ uint8_t sira_raw[];
  uint8_t* base = (uint8_t*)(((uintptr_t)sira_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sB0 = base;                          // W0  : KP/64 chunks x SC rows x 128 B
  uint8_t* sB3 = sB0 + (size_t)SC * KP * 2;     // Wr3 : 1 chunk x KP rows x 128 B
  uint8_t* sA0 = sB3 + (size_t)KP * 128;        // x0  : KP/64 chunks x ST rows x 128 B
  uint8_t* sA3 = sA0 + (size_t)ST * KP * 2;     // r2  : 1 chunk x ST rows x 128 B
  float* sIn = reinterpret_cast<float*>(sA3 + (size_t)ST * 128);  // in_feats tile [ST][pitch_in]; later the p0 tile [ST][SPITCH]
  __shared__ float sW1[R1 * 4], sW2[R2 * R1];
  __shared__ float sG1[R1], sBt1[R1], sG2[R2], sBt2[R2], sG3[KP], sBt3[KP], sG0[SC], sBt0[SC];
  __shared__ float redA[2][ST], redB[2][ST];
  __shared__ int sRow[ST], sSeg[ST];
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int ntiles = (N + ST - 1) / ST;
  if ((int)blockIdx.x >= ntiles) return;
  const int cin = d.cin;
  constexpr uint32_t TCOLS = KP > 128 ? 256 : 128;
  if (warp == 0) tmem_alloc(&tmem_slot, TCOLS);
  if (tid == 0) {
    mbar_init(smem_u32(&mbar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  for (int i = tid; i < R1 * 4; i += blockDim.x) sW1[i] = (i % 4) < d.rel_in ? d.rw[0][(i / 4) * d.rel_in + (i % 4)] : 0.f;
  for (int i = tid; i < R2 * R1; i += blockDim.x) sW2[i] = d.rw[1][i];
  for (int i = tid; i < R1; i += blockDim.x) sG1[i] = d.rg[0][i], sBt1[i] = d.rb[0][i];
  for (int i = tid; i < R2; i += blockDim.x) sG2[i] = d.rg[1][i], sBt2[i] = d.rb[1][i];
  for (int i = tid; i < KP; i += blockDim.x) sG3[i] = i < cin ? d.rg[2][i] : 0.f, sBt3[i] = i < cin ? d.rb[2][i] : 0.f;
  for (int i = tid; i < SC; i += blockDim.x) sG0[i] = d.g0[i], sBt0[i] = d.b0[i];
  stage_weight(sB0, SC, KP, d.w0, cin, 0, SC, cin);
  stage_weight(sB3, KP, 64, d.rw[2], R2, 0, cin, R2);
  uint32_t parity = 0;
  const int half = warp >> 2;
  const int lrow = (warp & 3) * 32 + (tid & 31);
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int k0 = tile * ST;
    const int nrow = min(ST, N - k0);
    __syncthreads();  // previous tile fully consumed (sIn / sSeg / operands)
    if (tid < ST) {
      int row = tid < nrow ? order[k0 + tid] : -1;
      sRow[tid] = row;
      sSeg[tid] = row >= 0 ? (int)inv[row] : -1;
    }
    __syncthreads();
    // in_feats rows -> smem (4-byte cp.async: rows of cin floats are only 4-byte aligned in general)
    for (int idx = tid; idx < nrow * cin; idx += blockDim.x) {
      int r = idx / cin, c = idx - r * cin;
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(smem_u32(sIn + r * pitch_in + c)),
                   "l"(in_feats + (size_t)sRow[r] * cin + c)
                   : "memory");
    }
    // rel-MLP layers 1, 2: thread per point (warps 0-3), result bf16 -> A operand of the layer-3 GEMM
    if (tid < ST) {
      float h2[R2];
      if (tid < nrow) {
        const float* fc = f_cluster + (size_t)sRow[tid] * d.rel_in;
        float x[4] = {0.f, 0.f, 0.f, 0.f};
        for (int i = 0; i < d.rel_in; i++) x[i] = fc[i] / d.rel_dist_scaler;
        float h1[R1];
#pragma unroll
        for (int o = 0; o < R1; o++)
          h1[o] = fmaf(sW1[o * 4 + 3], x[3], fmaf(sW1[o * 4 + 2], x[2], fmaf(sW1[o * 4 + 1], x[1], sW1[o * 4] * x[0])));
        float mean, rstd;
        ln_regs<R1>(h1, mean, rstd, d.eps);
#pragma unroll
        for (int o = 0; o < R1; o++) h1[o] = act_f((h1[o] - mean) * rstd * sG1[o] + sBt1[o], d.act);
#pragma unroll
        for (int o = 0; o < R2; o++) {
          float a = 0.f;
#pragma unroll
          for (int k = 0; k < R1; k++) a = fmaf(sW2[o * R1 + k], h1[k], a);
          h2[o] = a;
        }
        ln_regs<R2>(h2, mean, rstd, d.eps);
#pragma unroll
        for (int o = 0; o < R2; o++) h2[o] = act_f((h2[o] - mean) * rstd * sG2[o] + sBt2[o], d.act);
      } else {
#pragma unroll
        for (int o = 0; o < R2; o++) h2[o] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < 8; j++) {
        int4 q = make_int4(0, 0, 0, 0);
        if (j * 8 < R2) {
          q.x = (int)pack_bf16(h2[(j * 8 + 0) % R2], h2[(j * 8 + 1) % R2]);
          q.y = (int)pack_bf16(h2[(j * 8 + 2) % R2], h2[(j * 8 + 3) % R2]);
          q.z = (int)pack_bf16(h2[(j * 8 + 4) % R2], h2[(j * 8 + 5) % R2]);
          q.w = (int)pack_bf16(h2[(j * 8 + 6) % R2], h2[(j * 8 + 7) % R2]);
        }
        *reinterpret_cast<int4*>(sA3 + tid * 128 + ((j ^ (tid & 7)) << 4)) = q;
      }
    }
    asm volatile("cp.async.wait_all;\n" ::: "memory");
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if (tid == 0) {  // rel layer 3: [ST, R2] x [KP rows, R2]^T
      const uint32_t idesc = umma_idesc(128, KP);
      const uint32_t a0 = smem_u32(sA3), b0 = smem_u32(sB3);
#pragma unroll
      for (int s = 0; s < R2 / 16; s++) umma_bf16(tmem, umma_desc_sw128(a0 + s * 32), umma_desc_sw128(b0 + s * 32), idesc, s ? 1u : 0u);
      umma_commit(smem_u32(&mbar));
    }
    __syncwarp();
    mbar_wait(smem_u32(&mbar), parity);
    parity ^= 1u;
    tc_fence_after();
    const uint32_t tlane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    {
      // epilogue rel3: LayerNorm over cin, act, gate with [xyz / normalizer || feats] -> x0 (bf16 A operand of layer 0)
      float v[NH];
#pragma unroll
      for (int c0 = 0; c0 < NH; c0 += 32) tmem_ld32(tlane + half * NH + c0, v + c0);
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < NH; i++) s += (half * NH + i < cin) ? v[i] : 0.f;
      redA[half][lrow] = s;
      __syncthreads();
      const float mean = (redA[0][lrow] + redA[1][lrow]) / (float)cin;
      float q = 0.f;
#pragma unroll
      for (int i = 0; i < NH; i++) {
        float t = v[i] - mean;
        q += (half * NH + i < cin) ? t * t : 0.f;
      }
      redB[half][lrow] = q;
      __syncthreads();
      const float rstd = rsqrtf((redB[0][lrow] + redB[1][lrow]) / (float)cin + d.eps);
      const bool live = lrow < nrow;
      const float* frow = sIn + lrow * pitch_in;
#pragma unroll
      for (int j = 0; j < NH / 8; j++) {
        float y[8];
#pragma unroll
        for (int i = 0; i < 8; i++) {
          const int gc = half * NH + j * 8 + i;
          float o = 0.f;
          if (live && gc < cin) {
            float f = frow[gc];
            if (gc < 3) f = f / d.nz[gc];
            o = act_f((v[j * 8 + i] - mean) * rstd * sG3[gc] + sBt3[gc], d.act) * f;
          }
          y[i] = o;
        }
        int4 pk;
        pk.x = (int)pack_bf16(y[0], y[1]);
        pk.y = (int)pack_bf16(y[2], y[3]);
        pk.z = (int)pack_bf16(y[4], y[5]);
        pk.w = (int)pack_bf16(y[6], y[7]);
        const int k8 = (half * NH) / 8 + j;
        const int c = k8 >> 3, jj = k8 & 7;
        *reinterpret_cast<int4*>(sA0 + (size_t)c * ST * 128 + lrow * 128 + ((jj ^ (lrow & 7)) << 4)) = pk;
      }
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (tid == 0) {  // layer 0: [ST, KP] x [SC, KP]^T
      const uint32_t idesc = umma_idesc(128, SC);
      const uint32_t a0 = smem_u32(sA0), b0 = smem_u32(sB0);
#pragma unroll
      for (int c = 0; c < KP / 64; c++)
#pragma unroll
        for (int s = 0; s < 4; s++)
          umma_bf16(tmem, umma_desc_sw128(a0 + c * ST * 128 + s * 32), umma_desc_sw128(b0 + c * SC * 128 + s * 32), idesc, (c | s) ? 1u : 0u);
      umma_commit(smem_u32(&mbar));
    }
    __syncwarp();
    mbar_wait(smem_u32(&mbar), parity);
    parity ^= 1u;
    tc_fence_after();
    {
      // epilogue 0: LayerNorm over SC, act -> p0 (bf16 rows in CSR order for kernel B; fp32 tile for the pooled max)
      constexpr int H = SC / 2;
      float v[H];
#pragma unroll
      for (int c0 = 0; c0 < H; c0 += 32) tmem_ld32(tlane + half * H + c0, v + c0);
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < H; i++) s += v[i];
      redA[half][lrow] = s;
      __syncthreads();
      const float mean = (redA[0][lrow] + redA[1][lrow]) / (float)SC;
      float q = 0.f;
#pragma unroll
      for (int i = 0; i < H; i++) {
        float t = v[i] - mean;
        q = fmaf(t, t, q);
      }
      redB[half][lrow] = q;
      __syncthreads();
      const float rstd = rsqrtf((redB[0][lrow] + redB[1][lrow]) / (float)SC + d.eps);
      float* trow = sIn + lrow * SPITCH + half * H;  // sIn is dead (x0 is in the operand buffer): reuse as the p0 tile
#pragma unroll
      for (int i = 0; i < H; i++) {
        v[i] = act_f((v[i] - mean) * rstd * sG0[half * H + i] + sBt0[half * H + i], d.act);
        trow[i] = v[i];
      }
      if (lrow < nrow) {
        int4* dst = reinterpret_cast<int4*>(p0buf + (size_t)(k0 + lrow) * SC + half * H);
#pragma unroll
        for (int j = 0; j < H / 8; j++) {
          int4 pk;
          pk.x = (int)pack_bf16(v[j * 8 + 0], v[j * 8 + 1]);
          pk.y = (int)pack_bf16(v[j * 8 + 2], v[j * 8 + 3]);
          pk.z = (int)pack_bf16(v[j * 8 + 4], v[j * 8 + 5]);
          pk.w = (int)pack_bf16(v[j * 8 + 6], v[j * 8 + 7]);
          dst[j] = pk;
        }
      }
    }
    tc_fence_before();
    __syncthreads();
    tile_segmax_ord(sIn, sSeg, nrow, gord);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_slot, TCOLS);
}

// ------------------------------------------------------------------------------------------------------------------------
// kernel B
// ------------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 1) sir_b_kernel(SirDev d, const __nv_bfloat16* __restrict__ p0buf, const float* __restrict__ gterm,
                                                       const long long* __restrict__ inv, const int32_t* __restrict__ order, int N,
                                                       float* __restrict__ out_point, int ldo, uint32_t* __restrict__ gord) {
  pdl_wait();
  pdl_launch();
  extern __shared__ uint8_t sirb_raw[];
  uint8_t* base = (uint8_t*)(((uintptr_t)sirb_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sB = base;                        // W1a : 2 chunks x SC rows x 128 B
  uint8_t* sA = sB + (size_t)SC * SC * 2;    // p0  : 2 chunks x ST rows x 128 B
  float* sTile = reinterpret_cast<float*>(sA + (size_t)ST * SC * 2);  // [ST][SPITCH]
  __shared__ float sG[SC], sBt[SC];
  __shared__ float redA[2][ST], redB[2][ST];
  __shared__ int sRow[ST], sSeg[ST];
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int ntiles = (N + ST - 1) / ST;
  if ((int)blockIdx.x >= ntiles) return;
  if (warp == 0) tmem_alloc(&tmem_slot, SC);
  if (tid == 0) {
    mbar_init(smem_u32(&mbar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  for (int i = tid; i < SC; i += blockDim.x) sG[i] = d.g1[i], sBt[i] = d.b1[i];
  stage_weight(sB, SC, SC, d.w1, 2 * SC, 0, SC, SC);
  uint32_t parity = 0;
  const int half = warp >> 2;
  const int lrow = (warp & 3) * 32 + (tid & 31);
  constexpr int H = SC / 2;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int k0 = tile * ST;
    const int nrow = min(ST, N - k0);
    __syncthreads();
    if (tid < ST) {
      int row = tid < nrow ? order[k0 + tid] : -1;
      sRow[tid] = row;
      sSeg[tid] = row >= 0 ? (int)inv[row] : -1;
    }
    for (int idx = tid; idx < ST * (SC / 8); idx += blockDim.x) {
      int r = idx / (SC / 8), j = idx % (SC / 8);
      int c = j >> 3, jj = j & 7;
      uint8_t* dst = sA + (size_t)c * ST * 128 + r * 128 + ((jj ^ (r & 7)) << 4);
      if (r < nrow)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_u32(dst)), "l"(p0buf + (size_t)(k0 + r) * SC + j * 8) : "memory");
      else
        *reinterpret_cast<int4*>(dst) = make_int4(0, 0, 0, 0);
    }
    asm volatile("cp.async.wait_all;\n" ::: "memory");
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if (tid == 0) {
      const uint32_t idesc = umma_idesc(128, SC);
      const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
#pragma unroll
      for (int c = 0; c < SC / 64; c++)
#pragma unroll
        for (int s = 0; s < 4; s++)
          umma_bf16(tmem, umma_desc_sw128(a0 + c * ST * 128 + s * 32), umma_desc_sw128(b0 + c * SC * 128 + s * 32), idesc, (c | s) ? 1u : 0u);
      umma_commit(smem_u32(&mbar));
    }
    __syncwarp();
    mbar_wait(smem_u32(&mbar), parity);
    parity ^= 1u;
    tc_fence_after();
    const uint32_t tlane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    float v[H];
#pragma unroll
    for (int c0 = 0; c0 < H; c0 += 32) tmem_ld32(tlane + half * H + c0, v + c0);
    const bool live = lrow < nrow;
    if (live) {  // + (W1b . g0)[group]
      const float4* gt = reinterpret_cast<const float4*>(gterm + (size_t)sSeg[lrow] * SC + half * H);
#pragma unroll
      for (int j = 0; j < H / 4; j++) {
        float4 t4 = __ldg(gt + j);
        v[4 * j] += t4.x, v[4 * j + 1] += t4.y, v[4 * j + 2] += t4.z, v[4 * j + 3] += t4.w;
      }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < H; i++) s += v[i];
    redA[half][lrow] = s;
    __syncthreads();
    const float mean = (redA[0][lrow] + redA[1][lrow]) / (float)SC;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < H; i++) {
      float t = v[i] - mean;
      q = fmaf(t, t, q);
    }
    redB[half][lrow] = q;
    __syncthreads();
    const float rstd = rsqrtf((redB[0][lrow] + redB[1][lrow]) / (float)SC + d.eps);
    float* trow = sTile + lrow * SPITCH + half * H;
#pragma unroll
    for (int i = 0; i < H; i++) {
      v[i] = act_f((v[i] - mean) * rstd * sG[half * H + i] + sBt[half * H + i], d.act);
      trow[i] = v[i];
    }
    tc_fence_before();
    __syncthreads();
    // coalesced fp32 rows to the caller's buffer (original point order): one warp per row, 4 floats per lane
    for (int r = warp; r < nrow; r += 8) {
      const float* tr = sTile + r * SPITCH;
      float* orow = out_point + (size_t)sRow[r] * ldo;
      const int c = (tid & 31) * 4;
      if ((ldo & 3) == 0) {
        *reinterpret_cast<float4*>(orow + c) = make_float4(tr[c], tr[c + 1], tr[c + 2], tr[c + 3]);
      } else {
        orow[c] = tr[c], orow[c + 1] = tr[c + 1], orow[c + 2] = tr[c + 2], orow[c + 3] = tr[c + 3];
      }
    }
    tile_segmax_ord(sTile, sSeg, nrow, gord);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_slot, SC);
}

}  // namespace

// Host entry (called from sir.cu).  Returns SSTB_ERR_UNSUPPORTED for shapes outside the tensor path.
int sstb_sir_layer_bf16(sstb200_ctx* c, const sstb200_sir_layer* L, const float* in_feats, const float* f_cluster, const long long* inv,
                        const int32_t* order, int N, int G, __nv_bfloat16* p0buf, uint32_t* gord, float* gterm, float* out_point, int ldo,
                        float* out_group) {
  const int cin = L->in_channels;
  if (L->num_vfe != 2 || L->feat_channels[0] != SC || L->feat_channels[1] != SC || L->num_rel != 3 || L->rel_dims[0] != 16 ||
      L->rel_dims[1] != 32 || L->rel_dims[2] != cin || L->rel_in < 1 || L->rel_in > 4 || cin < 3 || cin > 192)
    return sstb_fail(c, SSTB_ERR_UNSUPPORTED,
                     "SIRLayer bf16 path needs feat_channels [128,128], rel-MLP [16,32,cin], cin <= 192 (got cin=%d, C=[%d,%d], rel=%d)", cin,
                     L->feat_channels[0], L->feat_channels[1], L->num_rel);
  SirDev d;
  d.cin = cin;
  d.rel_in = L->rel_in;
  d.act = L->act;
  d.eps = L->norm_eps;
  d.rel_dist_scaler = L->rel_dist_scaler;
  for (int i = 0; i < 3; i++) {
    d.nz[i] = L->xyz_normalizer[i];
    d.rw[i] = L->rel_w[i];
    d.rg[i] = L->rel_ln_w[i];
    d.rb[i] = L->rel_ln_b[i];
    CHECK_ARG(c, d.rw[i] && d.rg[i] && d.rb[i]);
  }
  d.w0 = L->vfe_w[0], d.g0 = L->vfe_ln_w[0], d.b0 = L->vfe_ln_b[0];
  d.w1 = L->vfe_w[1], d.g1 = L->vfe_ln_w[1], d.b1 = L->vfe_ln_b[1];
  CHECK_ARG(c, d.w0 && d.g0 && d.b0 && d.w1 && d.g1 && d.b1);
  cudaStream_t st = c->stream;
  const int KP = (cin + 63) / 64 * 64;
  const int pitch_in = (cin | 1) > SPITCH ? (cin | 1) : SPITCH;
  const size_t smemA = 1024 + (size_t)SC * KP * 2 + (size_t)KP * 128 + (size_t)ST * KP * 2 + (size_t)ST * 128 + (size_t)ST * pitch_in * 4;
  const size_t smemB = 1024 + (size_t)SC * SC * 2 + (size_t)ST * SC * 2 + (size_t)ST * SPITCH * 4;
  if (smemA > 218 * 1024) return sstb_fail(c, SSTB_ERR_UNSUPPORTED, "SIRLayer bf16: tile of cin=%d does not fit shared memory", cin);
  const int ntiles = (N + ST - 1) / ST;
  const int grid = ntiles < c->num_sms ? ntiles : c->num_sms;
  const size_t gn = (size_t)G * SC;
  const int Cg = 2 * SC;
  CUDA_TRY(c, cudaMemsetAsync(gord, 0, gn * 4, st));
#define SIR_A(KPV)                                                                                                          \
  do {                                                                                                                      \
    static bool attr = false;                                                                                               \
    if (!attr) {                                                                                                            \
      CUDA_TRY(c, cudaFuncSetAttribute(sir_a_kernel<KPV, 16, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 218 * 1024)); \
      attr = true;                                                                                                          \
    }                                                                                                                       \
    launch_pdl(sir_a_kernel<KPV, 16, 32>, dim3(grid), dim3(256), smemA, st, d, in_feats, f_cluster, inv, order, N, p0buf, gord, pitch_in); \
  } while (0)
  if (KP == 64) SIR_A(64);
  else if (KP == 128) SIR_A(128);
  else SIR_A(192);
#undef SIR_A
  sstb_sir_segmax_finalize(st, gord, G, SC, out_group, Cg, 0);
  // gterm = g0 . W1b^T  (G rows: SIMT fp32)
  sstb_gemm_rows_ex(st, out_group, Cg, L->vfe_w[1] + SC, 2 * SC, nullptr, nullptr, 0, nullptr, gterm, SC, G, nullptr, SC, SC, 0, nullptr,
                    nullptr, 0, 0, 0, 0);
  CUDA_TRY(c, cudaMemsetAsync(gord, 0, gn * 4, st));
  {
    static bool attr = false;
    if (!attr) {
      CUDA_TRY(c, cudaFuncSetAttribute(sir_b_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemB));
      attr = true;
    }
    launch_pdl(sir_b_kernel, dim3(grid), dim3(256), smemB, st, d, (const __nv_bfloat16*)p0buf, (const float*)gterm, inv, order, N, out_point, ldo, gord);
  }
  sstb_sir_segmax_finalize(st, gord, G, SC, out_group, Cg, SC);
  LAUNCH_CHECK(c);
  return SSTB_OK;
}

// Fused post-attention chain of one SRA encoder layer, warp-specialised (d_model = 128, dim_ff = 256, post-norm LayerNorm, GELU):
//
//     x1 = LayerNorm1(x + att . Wo^T + bo)          GEMM1  [128 x 128] x [128 x 128]
//     h  = GELU(x1 . W1^T + b1)                     GEMM2  [128 x 128] x [128 x 256]   four N = 64 chunks
//     y  = LayerNorm2(x1 + h . W2^T + b2)           GEMM3  [128 x 256] x [256 x 128]   four K = 64 chunks
//     q|k|v (next layer) = (y + pos | y) . Wqkv^T   GEMM4  [128 x 128] x [128 x 384]   three N = 128 chunks (optional tail)
//
// replaces mmdet3d/models/sst/sst_basic_block_v2.py:104-126 (+ the in-projection of the following layer's
// nn.MultiheadAttention, :70) for one 128-token tile per iteration of a persistent CTA.
//
// Roles (18 warps):  warp 17 = TMA producer (one lane): every operand tile - att, the residual x, all weights - arrives by
// cp.async.bulk.tensor into 128B-swizzled shared memory; weights + att stream through a ring of three 32 KB slots guarded by
// full/empty mbarriers.  warp 16 = MMA issuer (one lane): tcgen05.mma with TMEM accumulators, tcgen05.commit -> mbarriers.
// warps 0-15 = epilogue: TMEM -> registers (thread per row, 4 warps per TMEM lane quadrant each owning a column quarter),
// bias / residual / LayerNorm / GELU, operands of the next GEMM written straight into the swizzled K-major layout, outputs
// staged in the same swizzle and stored with TMA.  GEMM2 -> GELU -> GEMM3 is pipelined in 64-column chunks (double-buffered
// accumulator chunks in TMEM, double-buffered hidden chunks in shared memory), so the tensor pipe works on chunk c+1 while
// the epilogue warps run GELU on chunk c, and GEMM1 of the next tile is issued while the q|k|v epilogue of this one runs.
// All rows are in flat token order: every tile of att / x / y / q|k|v is a plain 2-D box.
//
// TMEM columns: [0,128) acc1 / acc3, [128,256) two acc2 chunks (later k), [256,384) x1 fp32 (later v), [384,512) q.
// Shared memory: ring 3 x 32 KB | A 32 KB (x1 operand, y staging lo, q/v staging) | B 32 KB (hidden ring, y staging hi,
// k staging) | C 64 KB (x fp32 tile, LN statistics exchange, then the (y+pos | y) operands) | mbarriers.
#include <stdarg.h>
#include <cuda_fp16.h>
#include "sra.cuh"
#include "tma.cuh"

namespace {

constexpr int TM = 128, D = 128;
constexpr int SLOT = 32768;
constexpr int W_MMA = 16, W_TMA = 17;
constexpr int NTHR = 18 * 32;
constexpr int NEPI = 512;
constexpr int OFF_RING = 0, OFF_A = 3 * SLOT, OFF_B = 4 * SLOT, OFF_C = 5 * SLOT, OFF_BAR = 7 * SLOT;
constexpr int SMEM_BYTES = 7 * SLOT + 512;

enum {
  B_FULL = 0,    // [3] ring slot filled (TMA transaction bytes)
  B_EMPTY = 3,   // [3] ring slot consumed (tcgen05.commit)
  B_XFULL = 6,   // residual tile landed in C
  B_CFREE = 7,   // C may be overwritten by the next residual tile
  B_ACC1 = 8,    // GEMM1 retired
  B_X1 = 9,      // x1 operand (A) + fp32 copy (TMEM) written             [16 warp arrivals]
  B_ACC2F = 10,  // [2] GEMM2 chunk retired
  B_ACC2E = 12,  // [2] acc2 chunk read back                              [16]
  B_HIDF = 14,   // [2] hidden chunk written                              [16]
  B_HIDE = 16,   // [2] hidden chunk consumed by GEMM3
  B_ACC3 = 18,   // GEMM3 retired
  B_YFULL = 19,  // LN2 done: acc3 / x1 read, (y+pos | y) operands written [16]
  B_QKVF = 20,   // [3] q / k / v chunk retired
  B_QKVE = 23,   // [3] q / k / v chunk read back                         [16]
  NBAR = 26
};

struct Chain2Maps {
  CUtensorMap att, x, y, qkv, wo, w1, w2, wqkv;
};

struct Chain2Args {
  const float *bo, *b1, *b2, *g1, *be1, *g2, *be2;
  float eps;
  int M_cap;
  const int32_t* M_dev;
  int has_tail;                  // GEMM4: q|k|v of the next layer
  const float* bqkv;             // [384]
  const int32_t* next_pos_code;  // [tokens]
  const float* pos_tab;          // [ndim][maxw][L]
  int posL, pos_maxw, pos_ndim;
};

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n" ::"r"(taddr),
      "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
      "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
      "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
      "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])),
      "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])), "r"(__float_as_uint(v[18])), "r"(__float_as_uint(v[19])),
      "r"(__float_as_uint(v[20])), "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])), "r"(__float_as_uint(v[23])),
      "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])), "r"(__float_as_uint(v[27])),
      "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])), "r"(__float_as_uint(v[30])), "r"(__float_as_uint(v[31]))
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
}

// tanh-form GELU on the MUFU.TANH unit (see csrc/sra_chain.cu: deviation from the erf form is below the bf16 rounding
// applied right after)
__device__ __forceinline__ float gelu_t(float x) {
  float u = 0.7978845608028654f * fmaf(0.044715f * x, x * x, x);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));
  return 0.5f * x * (1.0f + t);
}

__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// 4 (K = 64) or 8 (K = 128) tcgen05.mma k-steps on K-major SWIZZLE_128B operands; chunk pitch = bytes between 64-wide K chunks
__device__ __forceinline__ void mma_steps(uint32_t d_tmem, uint32_t a0, uint32_t a_chunk, uint32_t b0, uint32_t b_chunk, int kchunks,
                                          uint32_t idesc, bool accum) {
  for (int c = 0; c < kchunks; c++)
#pragma unroll
    for (int s = 0; s < 4; s++) {
      umma_bf16(d_tmem, umma_desc_sw128(a0 + c * a_chunk + s * 32), umma_desc_sw128(b0 + c * b_chunk + s * 32), idesc,
                (accum || c || s) ? 1u : 0u);
    }
}

__global__ void __launch_bounds__(NTHR, 1) sra_chain2_kernel(const __grid_constant__ Chain2Maps maps, const Chain2Args g) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_BAR + NBAR * 8);
#define BAR(i) (sbase + OFF_BAR + (uint32_t)(i) * 8u)

  if (tid == 0) {
    if (sbase & 1023u) __trap();   // SWIZZLE_128B operands / TMA boxes need the 1024-byte alignment
    for (int i = 0; i < NBAR; i++) {
      const bool warps16 = (i == B_X1) || (i == B_ACC2E) || (i == B_ACC2E + 1) || (i == B_HIDF) || (i == B_HIDF + 1) || (i == B_YFULL) ||
                           (i >= B_QKVE && i < B_QKVE + 3);
      mbar_init(BAR(i), warps16 ? 16u : 1u);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == W_TMA && lane == 0) {
    tma_prefetch_desc(&maps.att);
    tma_prefetch_desc(&maps.x);
    tma_prefetch_desc(&maps.wo);
    tma_prefetch_desc(&maps.w1);
    tma_prefetch_desc(&maps.w2);
    tma_prefetch_desc(&maps.y);
    if (g.has_tail) {
      tma_prefetch_desc(&maps.wqkv);
      tma_prefetch_desc(&maps.qkv);
    }
  }
  if (warp == W_MMA) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();     // everything above overlaps the tail of the producer kernel
  pdl_launch();
  const int M = g.M_dev ? *g.M_dev : g.M_cap;
  const int n_tiles = (M + TM - 1) / TM;
  const bool tail = g.has_tail != 0;
  const int NI = tail ? 9 : 6;   // ring items per tile: att, Wo, W1[0:128], W1[128:256], W2[:,0:128], W2[:,128:256], Wq, Wk, Wv

  if (warp == W_TMA) {
    // ================================================= TMA producer =================================================
    if (lane == 0) {
      int it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const int row0 = tile * TM;
        const int k0 = it * NI;
        auto load_item = [&](int j, const CUtensorMap* m, int ca, int ra, int cb, int rb) {
          const int kk = k0 + j, s = kk % 3;
          mbar_wait(BAR(B_EMPTY + s), (uint32_t)(((kk / 3) + 1) & 1));
          mbar_expect_tx(BAR(B_FULL + s), SLOT);
          tma_load_2d(sbase + OFF_RING + s * SLOT, m, ca, ra, BAR(B_FULL + s));
          tma_load_2d(sbase + OFF_RING + s * SLOT + 16384, m, cb, rb, BAR(B_FULL + s));
        };
        load_item(0, &maps.att, 0, row0, 64, row0);
        load_item(1, &maps.wo, 0, 0, 64, 0);
        mbar_wait(BAR(B_CFREE), (uint32_t)((it + 1) & 1));   // previous tile's users of C are done
        mbar_expect_tx(BAR(B_XFULL), 4 * 16384);
#pragma unroll
        for (int q = 0; q < 4; q++) tma_load_2d(sbase + OFF_C + q * 16384, &maps.x, q * 32, row0, BAR(B_XFULL));
        load_item(2, &maps.w1, 0, 0, 64, 0);
        load_item(3, &maps.w1, 0, 128, 64, 128);
        load_item(4, &maps.w2, 0, 0, 64, 0);
        load_item(5, &maps.w2, 128, 0, 192, 0);
        if (tail) {
          load_item(6, &maps.wqkv, 0, 0, 64, 0);
          load_item(7, &maps.wqkv, 0, 128, 64, 128);
          load_item(8, &maps.wqkv, 0, 256, 64, 256);
        }
      }
    }
    __syncwarp();
  } else if (warp == W_MMA) {
    // ================================================== MMA issuer ==================================================
    if (lane == 0) {
      const uint32_t idesc128 = umma_idesc(TM, 128), idesc64 = umma_idesc(TM, 64);
      int it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const int k0 = it * NI;
        const uint32_t par = (uint32_t)(it & 1);
        auto slot = [&](int j) { return sbase + OFF_RING + (uint32_t)((k0 + j) % 3) * SLOT; };
        auto wait_full = [&](int j) {
          const int kk = k0 + j;
          mbar_wait(BAR(B_FULL + kk % 3), (uint32_t)((kk / 3) & 1));
        };
        auto release = [&](int j) { umma_commit(BAR(B_EMPTY + (k0 + j) % 3)); };
        // ---- GEMM1: acc1 = att . Wo^T
        wait_full(0);
        wait_full(1);
        tc_fence_after();
        mma_steps(tmem, slot(0), 16384, slot(1), 16384, 2, idesc128, false);
        umma_commit(BAR(B_ACC1));
        release(0);
        release(1);
        // ---- GEMM2 (N chunks of 64) interleaved with GEMM3 (K chunks of 64)
        mbar_wait(BAR(B_X1), par);
        tc_fence_after();
        auto g2 = [&](int c) {
          if (c == 0) wait_full(2);
          if (c == 2) wait_full(3);
          mbar_wait(BAR(B_ACC2E + (c & 1)), (uint32_t)(((c >> 1) + 1) & 1));
          tc_fence_after();
          mma_steps(tmem + 128 + (c & 1) * 64, sbase + OFF_A, 16384, slot(2 + (c >> 1)) + (c & 1) * 8192, 16384, 2, idesc64, false);
          umma_commit(BAR(B_ACC2F + (c & 1)));
          if (c == 1) release(2);
          if (c == 3) release(3);
        };
        auto g3 = [&](int c) {
          if (c == 0) wait_full(4);
          if (c == 2) wait_full(5);
          mbar_wait(BAR(B_HIDF + (c & 1)), (uint32_t)((c >> 1) & 1));
          tc_fence_after();
          mma_steps(tmem, sbase + OFF_B + (c & 1) * 16384, 0, slot(4 + (c >> 1)) + (c & 1) * 16384, 0, 1, idesc128, c > 0);
          umma_commit(BAR(B_HIDE + (c & 1)));
          if (c == 1) release(4);
          if (c == 3) {
            release(5);
            umma_commit(BAR(B_ACC3));
          }
        };
        g2(0);
        g2(1);
        g3(0);
        g2(2);
        g3(1);
        g2(3);
        g3(2);
        g3(3);
        // ---- LN2 done: acc3 / x1 have been read, (y+pos | y) operands are in C
        mbar_wait(BAR(B_YFULL), par);
        tc_fence_after();
        if (!tail) mbar_arrive(BAR(B_CFREE));   // residual consumed and the LN2 statistics exchange (which lives in C) is over
        if (tail) {
#pragma unroll 1
          for (int nt = 0; nt < 3; nt++) {
            wait_full(6 + nt);
            mbar_wait(BAR(B_QKVE + nt), par ^ 1u);
            tc_fence_after();
            const uint32_t dcol = nt == 0 ? 384u : (nt == 1 ? 128u : 256u);
            mma_steps(tmem + dcol, sbase + OFF_C + (nt < 2 ? 0 : 32768), 16384, slot(6 + nt), 16384, 2, idesc128, false);
            umma_commit(BAR(B_QKVF + nt));
            release(6 + nt);
          }
          umma_commit(BAR(B_CFREE));
        }
      }
    }
    __syncwarp();
  } else {
    // ================================================ epilogue warps ================================================
    const int qd = warp & 3, cq = warp >> 2;
    const int lrow = qd * 32 + lane;                       // tile row == TMEM lane
    const uint32_t tlane = tmem + ((uint32_t)(qd * 32) << 16);
    const int c0 = cq * 32;                                // this thread's 32 columns of a 128-wide row
    const uint32_t sw = (uint32_t)(lrow & 7);
    uint8_t* const rowA = smem + OFF_A + lrow * 128;       // + box * 16384 + ((piece ^ sw) << 4)
    uint8_t* const rowC = smem + OFF_C + lrow * 128;
    // LN statistics exchange: each thread parks (sum, sumsq) in the first 8 bytes of its own (dead) residual span
    float2* const stat_own = reinterpret_cast<float2*>(rowC + cq * 16384 + (sw << 4));
    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const int row0 = tile * TM;
      const uint32_t par = (uint32_t)(it & 1);
      if (tid == 0) tma_store_wait_read<0>();   // the previous tile's stores have drained A / B
      named_bar_sync(1, NEPI);
      float t[32];
      // ---------------- epilogue 1: x1 = LN1(x + acc1 + bo) -> fp32 in TMEM[256..384), bf16 operand in A ----------------
      mbar_wait(BAR(B_XFULL), par);
      mbar_wait(BAR(B_ACC1), par);
      tc_fence_after();
      {
        float v[32];
        tmem_ld32(tlane + c0, v);
        float sum = 0.f, sq = 0.f;
#pragma unroll
        for (int q = 0; q < 8; q++) {
          const float4 r4 = *reinterpret_cast<const float4*>(rowC + cq * 16384 + (((uint32_t)q ^ sw) << 4));
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(g.bo + c0) + q);
          const float a0 = v[4 * q] + b4.x + r4.x, a1 = v[4 * q + 1] + b4.y + r4.y, a2 = v[4 * q + 2] + b4.z + r4.z,
                      a3 = v[4 * q + 3] + b4.w + r4.w;
          t[4 * q] = a0;
          t[4 * q + 1] = a1;
          t[4 * q + 2] = a2;
          t[4 * q + 3] = a3;
          sum += (a0 + a1) + (a2 + a3);
          sq += (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3);
        }
        *stat_own = make_float2(sum, sq);
      }
      named_bar_sync(2 + qd, 128);
      {
        const float2 s0 = *reinterpret_cast<const float2*>(rowC + 0 * 16384 + (sw << 4));
        const float2 s1 = *reinterpret_cast<const float2*>(rowC + 1 * 16384 + (sw << 4));
        const float2 s2 = *reinterpret_cast<const float2*>(rowC + 2 * 16384 + (sw << 4));
        const float2 s3 = *reinterpret_cast<const float2*>(rowC + 3 * 16384 + (sw << 4));
        const float sum = (s0.x + s1.x) + (s2.x + s3.x), sq = (s0.y + s1.y) + (s2.y + s3.y);
        const float mean = sum * (1.0f / D);
        const float rstd = rsqrtf(fmaxf(sq * (1.0f / D) - mean * mean, 0.f) + g.eps);
#pragma unroll
        for (int q = 0; q < 8; q++) {
          const float4 g4 = __ldg(reinterpret_cast<const float4*>(g.g1 + c0) + q), e4 = __ldg(reinterpret_cast<const float4*>(g.be1 + c0) + q);
          t[4 * q] = (t[4 * q] - mean) * rstd * g4.x + e4.x;
          t[4 * q + 1] = (t[4 * q + 1] - mean) * rstd * g4.y + e4.y;
          t[4 * q + 2] = (t[4 * q + 2] - mean) * rstd * g4.z + e4.z;
          t[4 * q + 3] = (t[4 * q + 3] - mean) * rstd * g4.w + e4.w;
        }
        tmem_st32(tlane + 256 + c0, t);
        const int kc = cq >> 1, j0 = (cq & 1) * 4;
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const float* s = &t[q * 8];
          *reinterpret_cast<int4*>(rowA + kc * 16384 + (((uint32_t)(j0 + q) ^ sw) << 4)) =
              make_int4((int)pack_bf16(s[0], s[1]), (int)pack_bf16(s[2], s[3]), (int)pack_bf16(s[4], s[5]), (int)pack_bf16(s[6], s[7]));
        }
      }
      fence_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(BAR(B_X1));

      // ---------------- epilogue 2: hidden chunk c = GELU(acc2 chunk + b1) -> bf16 K-chunk of GEMM3's A operand in B ----------------
#pragma unroll 1
      for (int c = 0; c < 4; c++) {
        const int b = c & 1;
        mbar_wait(BAR(B_ACC2F + b), (uint32_t)((c >> 1) & 1));
        tc_fence_after();
        float v[16];
        tmem_ld16(tlane + 128 + b * 64 + cq * 16, v);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(BAR(B_ACC2E + b));
        const float4* bp = reinterpret_cast<const float4*>(g.b1 + c * 64 + cq * 16);
        uint32_t pk[8];
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          const float4 b4 = __ldg(bp + (i >> 2));
          pk[i >> 1] = pack_bf16(gelu_t(v[i] + b4.x), gelu_t(v[i + 1] + b4.y));
          pk[(i >> 1) + 1] = pack_bf16(gelu_t(v[i + 2] + b4.z), gelu_t(v[i + 3] + b4.w));
        }
        mbar_wait(BAR(B_HIDE + b), (uint32_t)(((c >> 1) + 1) & 1));   // GEMM3 has consumed the chunk that lived here
        uint8_t* hb = smem + OFF_B + b * 16384 + lrow * 128;
        *reinterpret_cast<int4*>(hb + (((uint32_t)(cq * 2) ^ sw) << 4)) = make_int4((int)pk[0], (int)pk[1], (int)pk[2], (int)pk[3]);
        *reinterpret_cast<int4*>(hb + (((uint32_t)(cq * 2 + 1) ^ sw) << 4)) = make_int4((int)pk[4], (int)pk[5], (int)pk[6], (int)pk[7]);
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(BAR(B_HIDF + b));
      }

      // ---------------- epilogue 3: y = LN2(x1 + acc3 + b2) -> fp32 staging (TMA store) + operands of GEMM4 ----------------
      mbar_wait(BAR(B_ACC3), par);
      tc_fence_after();
      {
        float v[32], r[32];
        tmem_ld32(tlane + c0, v);
        tmem_ld32(tlane + 256 + c0, r);
        float sum = 0.f, sq = 0.f;
#pragma unroll
        for (int q = 0; q < 8; q++) {
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(g.b2 + c0) + q);
          const float a0 = v[4 * q] + b4.x + r[4 * q], a1 = v[4 * q + 1] + b4.y + r[4 * q + 1];
          const float a2 = v[4 * q + 2] + b4.z + r[4 * q + 2], a3 = v[4 * q + 3] + b4.w + r[4 * q + 3];
          t[4 * q] = a0;
          t[4 * q + 1] = a1;
          t[4 * q + 2] = a2;
          t[4 * q + 3] = a3;
          sum += (a0 + a1) + (a2 + a3);
          sq += (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3);
        }
        *stat_own = make_float2(sum, sq);
      }
      named_bar_sync(2 + qd, 128);
      {
        const float2 s0 = *reinterpret_cast<const float2*>(rowC + 0 * 16384 + (sw << 4));
        const float2 s1 = *reinterpret_cast<const float2*>(rowC + 1 * 16384 + (sw << 4));
        const float2 s2 = *reinterpret_cast<const float2*>(rowC + 2 * 16384 + (sw << 4));
        const float2 s3 = *reinterpret_cast<const float2*>(rowC + 3 * 16384 + (sw << 4));
        const float sum = (s0.x + s1.x) + (s2.x + s3.x), sq = (s0.y + s1.y) + (s2.y + s3.y);
        const float mean = sum * (1.0f / D);
        const float rstd = rsqrtf(fmaxf(sq * (1.0f / D) - mean * mean, 0.f) + g.eps);
        named_bar_sync(2 + qd, 128);   // every statistic of this quadrant has been read: C may now receive the operands
#pragma unroll
        for (int q = 0; q < 8; q++) {
          const float4 g4 = __ldg(reinterpret_cast<const float4*>(g.g2 + c0) + q), e4 = __ldg(reinterpret_cast<const float4*>(g.be2 + c0) + q);
          float4 o;
          o.x = (t[4 * q] - mean) * rstd * g4.x + e4.x;
          o.y = (t[4 * q + 1] - mean) * rstd * g4.y + e4.y;
          o.z = (t[4 * q + 2] - mean) * rstd * g4.z + e4.z;
          o.w = (t[4 * q + 3] - mean) * rstd * g4.w + e4.w;
          *reinterpret_cast<float4*>(rowA + cq * 16384 + (((uint32_t)q ^ sw) << 4)) = o;   // y staging spans A | B (4 boxes of 32 columns)
          t[4 * q] = o.x;
          t[4 * q + 1] = o.y;
          t[4 * q + 2] = o.z;
          t[4 * q + 3] = o.w;
        }
      }
      if (tail) {
        const int tok = row0 + lrow;
        float pe[32];
#pragma unroll
        for (int i = 0; i < 32; i++) pe[i] = 0.f;
        const int axis = c0 / g.posL;   // posL % 32 == 0 (host check): the 32-column span lies inside one axis
        if (tok < M && axis < g.pos_ndim) {
          const int cv = (g.next_pos_code[tok] >> (8 * axis)) & 255;
          const float4* tp = reinterpret_cast<const float4*>(g.pos_tab + ((size_t)axis * g.pos_maxw + cv) * g.posL + (c0 - axis * g.posL));
#pragma unroll
          for (int q = 0; q < 8; q++) {
            const float4 p4 = __ldg(tp + q);
            pe[4 * q] = p4.x;
            pe[4 * q + 1] = p4.y;
            pe[4 * q + 2] = p4.z;
            pe[4 * q + 3] = p4.w;
          }
        }
        const int kc = cq >> 1, j0 = (cq & 1) * 4;
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const float* sy = &t[q * 8];
          const float* sp = &pe[q * 8];
          uint8_t* dst = rowC + kc * 16384 + (((uint32_t)(j0 + q) ^ sw) << 4);
          *reinterpret_cast<int4*>(dst) = make_int4((int)pack_bf16(sy[0] + sp[0], sy[1] + sp[1]), (int)pack_bf16(sy[2] + sp[2], sy[3] + sp[3]),
                                                    (int)pack_bf16(sy[4] + sp[4], sy[5] + sp[5]), (int)pack_bf16(sy[6] + sp[6], sy[7] + sp[7]));
          *reinterpret_cast<int4*>(dst + 32768) = make_int4((int)pack_bf16(sy[0], sy[1]), (int)pack_bf16(sy[2], sy[3]),
                                                            (int)pack_bf16(sy[4], sy[5]), (int)pack_bf16(sy[6], sy[7]));
        }
      }
      fence_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(BAR(B_YFULL));
      named_bar_sync(1, NEPI);
      if (tid == 0) {
        tma_store_2d(&maps.y, 0, row0, sbase + OFF_A);
        tma_store_2d(&maps.y, 32, row0, sbase + OFF_A + 16384);
        tma_store_commit();
        tma_store_2d(&maps.y, 64, row0, sbase + OFF_B);
        tma_store_2d(&maps.y, 96, row0, sbase + OFF_B + 16384);
        tma_store_commit();
      }

      // ---------------- epilogue 4: q | k | v chunk + bias -> fp16 staging (q, v in A; k in B) -> TMA store ----------------
      if (tail) {
#pragma unroll 1
        for (int nt = 0; nt < 3; nt++) {
          mbar_wait(BAR(B_QKVF + nt), par);
          tc_fence_after();
          float v[32];
          tmem_ld32(tlane + (nt == 0 ? 384 : (nt == 1 ? 128 : 256)) + c0, v);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(BAR(B_QKVE + nt));
          const float4* bp = reinterpret_cast<const float4*>(g.bqkv + nt * 128 + c0);
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 b4 = __ldg(bp + (i >> 2));
            pk[i >> 1] = pack_h2(v[i] + b4.x, v[i + 1] + b4.y);
            pk[(i >> 1) + 1] = pack_h2(v[i + 2] + b4.z, v[i + 3] + b4.w);
          }
          if (tid == 0) tma_store_wait_read<1>();   // the store that last read this staging buffer has drained
          named_bar_sync(1, NEPI);
          uint8_t* st = smem + (nt == 1 ? OFF_B : OFF_A) + (cq >> 1) * 16384 + lrow * 128;
          const int j0 = (cq & 1) * 4;
#pragma unroll
          for (int q = 0; q < 4; q++)
            *reinterpret_cast<int4*>(st + (((uint32_t)(j0 + q) ^ sw) << 4)) =
                make_int4((int)pk[4 * q], (int)pk[4 * q + 1], (int)pk[4 * q + 2], (int)pk[4 * q + 3]);
          fence_async_smem();
          named_bar_sync(1, NEPI);
          if (tid == 0) {
            const uint32_t sa = sbase + (nt == 1 ? OFF_B : OFF_A);
            tma_store_2d(&maps.qkv, nt * 128, row0, sa);
            tma_store_2d(&maps.qkv, nt * 128 + 64, row0, sa + 16384);
            tma_store_commit();
          }
        }
      }
    }
    if (tid == 0) tma_store_wait_all<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) tmem_dealloc(tmem, 512);
#undef BAR
}

}  // namespace

int sstb_sra_chain2_bf16(sstb200_ctx* c, const sstb200_sra_layer* L, const __nv_bfloat16* att, const float* x, float* y, int n_cap,
                         const int32_t* n_dev, const sstb200_sra_layer* next, const sstb200_sra_plan* next_plan, void* next_qkv) {
  Chain2Maps maps;
  Chain2Args g;
  memset(&g, 0, sizeof(g));
  memset(&maps, 0, sizeof(maps));
  const bool tail = next && next_plan && next_qkv;
  if (tail && (next_plan->pos_L % 32 != 0 || !next_plan->pos_code || !next_plan->pos_table))
    return sstb_fail(c, SSTB_ERR_UNSUPPORTED, "fused QKV tail needs pos_L %% 32 == 0 and the next plan's pos_code / pos_table");
  int rc = 0;
  rc |= tmap_2d_sw128(&maps.att, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, att, 128, (uint64_t)n_cap, 256, 64, 128);
  rc |= tmap_2d_sw128(&maps.x, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, x, 128, (uint64_t)n_cap, 512, 32, 128);
  rc |= tmap_2d_sw128(&maps.y, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, y, 128, (uint64_t)n_cap, 512, 32, 128);
  rc |= tmap_2d_sw128(&maps.wo, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, L->out_proj_w_bf16, 128, 128, 256, 64, 128);
  rc |= tmap_2d_sw128(&maps.w1, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, L->lin1_w_bf16, 128, 256, 256, 64, 128);
  rc |= tmap_2d_sw128(&maps.w2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, L->lin2_w_bf16, 256, 128, 512, 64, 128);
  if (tail) {
    rc |= tmap_2d_sw128(&maps.wqkv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, next->in_proj_w_bf16, 128, 384, 256, 64, 128);
    rc |= tmap_2d_sw128(&maps.qkv, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, next_qkv, 384, (uint64_t)n_cap, 768, 64, 128);
    g.has_tail = 1;
    g.bqkv = next->in_proj_b;
    g.next_pos_code = next_plan->pos_code;
    g.pos_tab = next_plan->pos_table;
    g.posL = next_plan->pos_L;
    g.pos_maxw = next_plan->pos_maxw;
    g.pos_ndim = next_plan->pos_ndim;
  }
  if (rc) return sstb_fail(c, SSTB_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d): operands must be 16-byte aligned", rc);
  g.bo = L->out_proj_b;
  g.b1 = L->lin1_b;
  g.b2 = L->lin2_b;
  g.g1 = L->norm1_w;
  g.be1 = L->norm1_b;
  g.g2 = L->norm2_w;
  g.be2 = L->norm2_b;
  g.eps = L->norm_eps;
  g.M_cap = n_cap;
  g.M_dev = n_dev;
  static SmemAttr sa;
  CUDA_TRY(c, ensure_smem(c, sa, sra_chain2_kernel, (size_t)SMEM_BYTES));
  const int tiles_cap = (n_cap + TM - 1) / TM;
  const int grid = c->num_sms < tiles_cap ? c->num_sms : tiles_cap;
  CUDA_TRY(c, launch_pdl(sra_chain2_kernel, dim3(grid), dim3(NTHR), (size_t)SMEM_BYTES, c->stream, maps, g));
  return SSTB_OK;
}

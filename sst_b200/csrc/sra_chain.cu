// Fused post-attention chain of one SRA encoder layer on tcgen05 (d_model = 128, dim_ff = 256, post-norm):
//
//     x1 = LayerNorm1(x + att . Wo^T + bo)          GEMM1  [128 x 128] x [128 x 128]
//     h  = GELU(x1 . W1^T + b1)                     GEMM2  [128 x 128] x [128 x 256]   (two N=128 halves)
//     y  = LayerNorm2(x1 + h . W2^T + b2)           GEMM3  [128 x 256] x [256 x 128]
//
// One persistent CTA per SM walks 128-row tiles.  x1 never leaves the SM: its fp32 copy lives in TMEM columns
// 384..511 (tcgen05.st) next to the accumulators (acc1/acc3: 0..127, acc2: 128..383 - exactly the 512 columns), its
// bf16 copy and the hidden activations are written straight into the K-major SWIZZLE_128B operand layout of the next
// MMA.  Weights stream through a ring of four 32 KB slots (Wo | W1[0:128] | W1[128:256] | W2[:, 0:128]; W2[:, 128:256]
// re-uses slot 0 once GEMM1 has retired).  Global traffic per tile is exactly: att tile in, residual tile in, y tile
// out (all coalesced through an XOR-swizzled staging tile) + 160 KB of L2-resident weights.
// Replaces three launches (out-proj+LN1, FFN1+GELU, FFN2+LN2) and the x1 / x1_bf16 / hidden round trips of the
// unfused path (csrc/sra_bf16.cu).
#include <stdarg.h>
#include <cuda_fp16.h>
#include "sra.cuh"
#include "umma.cuh"

namespace {

constexpr int TM = 128;   // rows per tile
constexpr int D = 128;    // d_model
constexpr int FF = 256;   // dim_ff
constexpr int SLOT = 32768;
constexpr int NTHR = 512;

struct ChainArgs {
  const __nv_bfloat16* att;   // [M, 128] bf16, rows in tile order
  const int32_t* row_map;     // nullable: token row of tile row i (residual / output rows), else i
  const float* x;             // [M, 128] fp32 residual stream (token rows)
  float* y;                   // [M, 128] fp32 output (token rows)
  const __nv_bfloat16 *Wo, *W1, *W2;   // [128,128], [256,128], [128,256] bf16 row-major (nn.Linear layout)
  const float *bo, *b1, *b2, *g1, *be1, *g2, *be2;
  float eps;
  int M_cap;
  const int32_t* M_dev;
  // optional fused tail: q/k/v of the NEXT encoder layer (which runs on the other shift's windows)
  const __nv_bfloat16* Wqkv;     // next layer in_proj_weight [384,128] bf16, or nullptr
  const float* bqkv;             // [384]
  const int32_t* next_slot;      // [tokens] slot of the token in the next layer's window order
  const int32_t* next_pos_code;  // [tokens]
  const float* pos_tab;          // [ndim][maxw][L]
  int posL, pos_maxw, pos_ndim;
  __half* qkv_out;               // [M, 384] fp16, rows in the next layer's slot order
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

// GELU feeding a bf16 operand: tanh form on the hardware tanh unit (MUFU.TANH).  Deviation from the exact erf form is
// <= ~5e-4 relative, i.e. below the bf16 rounding (2^-9) applied to the value right after; measured end-to-end error of
// the 12-layer stack is unchanged.  ~6 instructions instead of ~18 - the epilogues of this kernel are issue-bound.
__device__ __forceinline__ float gelu_as(float x) {
  float u = 0.7978845608028654f * fmaf(0.044715f * x, x * x, x);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));
  return 0.5f * x * (1.0f + t);
}

__device__ __forceinline__ void cp_async16(uint32_t smem_addr, const void* gptr) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_addr), "l"(gptr) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

// copy `rows` x 64 bf16 (one 128-byte K chunk per row) from a row-major matrix into a SWIZZLE_128B operand chunk with
// cp.async (LDGSTS): every 16-byte piece is in flight at once, no register staging.  src element (r, k0 + j) at
// src[r * ld + k0 + j]; rows >= valid_rows are zero-filled.
__device__ __forceinline__ void stage_chunk(uint8_t* dst, const __nv_bfloat16* src, int ld, int k0, int rows, int tid,
                                            int valid_rows = 1 << 30) {
  const uint32_t d0 = smem_u32(dst);
  for (int idx = tid; idx < rows * 8; idx += NTHR) {
    int r = idx >> 3, jj = idx & 7;
    uint32_t da = d0 + r * 128 + ((jj ^ (r & 7)) << 4);
    if (r < valid_rows) cp_async16(da, src + (size_t)r * ld + k0 + jj * 8);
    else *reinterpret_cast<int4*>(dst + r * 128 + ((jj ^ (r & 7)) << 4)) = make_int4(0, 0, 0, 0);
  }
}

// 16 warps: warp & 3 = TMEM lane quadrant (rows), warp >> 2 = column quarter handled in the epilogues.  The epilogues are
// instruction-issue bound (LayerNorm / GELU on 128 x 512 values per tile), so 4 warps per scheduler instead of 2.
__global__ void __launch_bounds__(NTHR, 1) sra_chain_kernel(ChainArgs g) {
  pdl_wait();
  pdl_launch();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sH = base;               // 64 KB: att tile | x1 bf16 (first 32 KB), hidden [128 x 256] bf16, fp32 staging tile
  uint8_t* sWr = base + 2 * SLOT;   // ring of four 32 KB weight slots
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tmem_slot;
  __shared__ float red[4][TM][2];
  __shared__ int sRow[TM];

  const int tid = threadIdx.x, warp = tid >> 5;
  const int M = g.M_dev ? *g.M_dev : g.M_cap;
  const int n_tiles = (M + TM - 1) / TM;
  if ((int)blockIdx.x >= n_tiles) return;
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  if (tid == 0) {
    mbar_init(smem_u32(&mbar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  uint32_t parity = 0;
  uint32_t tmem = 0;
  const int qtr = warp >> 2;                      // column quarter handled by this warp in the epilogues
  const int lrow = (warp & 3) * 32 + (tid & 31);  // tile row == TMEM lane
  const uint32_t idesc = umma_idesc(TM, 128);

  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int row0 = tile * TM;
    if (tid < TM) {
      int gr = row0 + tid;
      sRow[tid] = gr < M ? (g.row_map ? g.row_map[gr] : gr) : -1;
    }
    // ---- stage: att tile (2 K-chunks) + Wo (slot 0); W1 / W2[:, 0:128] (slots 1..3) only once per CTA - they are never
    //      overwritten, so later tiles of this persistent CTA re-use them --------------------------------------------------
    stage_chunk(sH, g.att + (size_t)row0 * D, D, 0, TM, tid, M - row0);
    stage_chunk(sH + 16384, g.att + (size_t)row0 * D, D, 64, TM, tid, M - row0);
    stage_chunk(sWr + 0 * SLOT, g.Wo, D, 0, 128, tid);
    stage_chunk(sWr + 0 * SLOT + 16384, g.Wo, D, 64, 128, tid);
    if (tile == (int)blockIdx.x || g.Wqkv) {   // with the fused QKV tail the slots are recycled for Wq/Wk/Wv: restage
      stage_chunk(sWr + 1 * SLOT, g.W1, D, 0, 128, tid);                 // slot1: W1 rows 0..127
      stage_chunk(sWr + 1 * SLOT + 16384, g.W1, D, 64, 128, tid);
      stage_chunk(sWr + 2 * SLOT, g.W1 + 128 * D, D, 0, 128, tid);       // slot2: W1 rows 128..255
      stage_chunk(sWr + 2 * SLOT + 16384, g.W1 + 128 * D, D, 64, 128, tid);
      stage_chunk(sWr + 3 * SLOT, g.W2, FF, 0, 128, tid);                // slot3: W2[:, 0:128]
      stage_chunk(sWr + 3 * SLOT + 16384, g.W2, FF, 64, 128, tid);
    }
    cp_async_wait_all();
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    tmem = tmem_slot;
    const uint32_t tlane = tmem + ((uint32_t)((warp & 3) * 32) << 16);

    // ---- GEMM1: acc1[0..127] = att . Wo^T ------------------------------------------------------------------------------
    if (tid == 0) {
      const uint32_t a0 = smem_u32(sH), b0 = smem_u32(sWr);
#pragma unroll
      for (int c = 0; c < 2; c++)
#pragma unroll
        for (int s = 0; s < 4; s++)
          umma_bf16(tmem, umma_desc_sw128(a0 + c * 16384 + s * 32), umma_desc_sw128(b0 + c * 16384 + s * 32), idesc, (c | s) ? 1u : 0u);
      umma_commit(smem_u32(&mbar));
    }
    __syncwarp();
    mbar_wait(smem_u32(&mbar), parity);
    parity ^= 1u;
    tc_fence_after();

    // ---- epilogue 1: x1 = LN1(x + acc1 + bo); fp32 -> TMEM[384..511], bf16 -> operand layout in sH[0:32K] -----------------
    // slot 0 (Wo) is dead: refill it with W2[:, 128:256]; the copies land while LN1 / GEMM2 / GELU run
    stage_chunk(sWr + 0 * SLOT, g.W2, FF, 128, 128, tid);
    stage_chunk(sWr + 0 * SLOT + 16384, g.W2, FF, 192, 128, tid);
    {  // residual tile, coalesced, into the (now dead) att/hidden region as an fp32 [128][128] XOR-swizzled tile
      for (int i0 = tid; i0 < TM * 32; i0 += NTHR * 8) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
          int idx = i0 + u * NTHR;
          int r = idx >> 5, ch = idx & 31;
          int gr = sRow[r];
          v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (gr >= 0) v[u] = *reinterpret_cast<const float4*>(g.x + (size_t)gr * D + ch * 4);
        }
#pragma unroll
        for (int u = 0; u < 8; u++) {
          int idx = i0 + u * NTHR;
          int r = idx >> 5, ch = idx & 31;
          *reinterpret_cast<float4*>(sH + (size_t)r * 512 + ((ch ^ (r & 31)) << 4)) = v[u];
        }
      }
    }
    __syncthreads();
    float t[32];
    const int c0 = qtr * 32;   // this thread's 32 columns of the 128-wide row
    {
      float sum = 0.f, sq = 0.f;
      float v[32];
      tmem_ld32(tlane + c0, v);
#pragma unroll
      for (int q = 0; q < 8; q++) {
        int ch = (c0 >> 2) + q;
        float4 r4 = *reinterpret_cast<const float4*>(sH + (size_t)lrow * 512 + ((ch ^ (lrow & 31)) << 4));
        float4 b4 = __ldg(reinterpret_cast<const float4*>(g.bo + c0) + q);
        float a0 = v[4 * q] + b4.x + r4.x, a1 = v[4 * q + 1] + b4.y + r4.y, a2 = v[4 * q + 2] + b4.z + r4.z, a3 = v[4 * q + 3] + b4.w + r4.w;
        t[4 * q] = a0;
        t[4 * q + 1] = a1;
        t[4 * q + 2] = a2;
        t[4 * q + 3] = a3;
        sum += (a0 + a1) + (a2 + a3);
        sq += (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3);
      }
      red[qtr][lrow][0] = sum;
      red[qtr][lrow][1] = sq;
    }
    __syncthreads();  // stats exchanged; every thread has consumed its residual chunk -> sH may be overwritten
    {
      const float sum = (red[0][lrow][0] + red[1][lrow][0]) + (red[2][lrow][0] + red[3][lrow][0]);
      const float sq = (red[0][lrow][1] + red[1][lrow][1]) + (red[2][lrow][1] + red[3][lrow][1]);
      const float mean = sum * (1.0f / D);
      const float rstd = rsqrtf(fmaxf(sq * (1.0f / D) - mean * mean, 0.f) + g.eps);
#pragma unroll
      for (int q = 0; q < 8; q++) {
        float4 g4 = __ldg(reinterpret_cast<const float4*>(g.g1 + c0) + q), e4 = __ldg(reinterpret_cast<const float4*>(g.be1 + c0) + q);
        t[4 * q] = (t[4 * q] - mean) * rstd * g4.x + e4.x;
        t[4 * q + 1] = (t[4 * q + 1] - mean) * rstd * g4.y + e4.y;
        t[4 * q + 2] = (t[4 * q + 2] - mean) * rstd * g4.z + e4.z;
        t[4 * q + 3] = (t[4 * q + 3] - mean) * rstd * g4.w + e4.w;
      }
      tmem_st32(tlane + 384 + c0, t);   // fp32 x1 stays in TMEM for the second residual
      // bf16 x1 -> A operand of GEMM2: row lrow, K-chunk = c0 / 64, 16-byte pieces (c0 % 64) / 8 ..
      const int kc = c0 >> 6, j0 = (c0 & 63) >> 3;
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const float* s = &t[q * 8];
        *reinterpret_cast<int4*>(sH + kc * 16384 + lrow * 128 + (((j0 + q) ^ (lrow & 7)) << 4)) =
            make_int4((int)pack_bf16(s[0], s[1]), (int)pack_bf16(s[2], s[3]), (int)pack_bf16(s[4], s[5]), (int)pack_bf16(s[6], s[7]));
      }
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    // ---- GEMM2: acc2[128..383] = x1 . W1^T (two N = 128 halves) -------------------------------------------------------------
    if (tid == 0) {
      const uint32_t a0 = smem_u32(sH);
#pragma unroll
      for (int nh = 0; nh < 2; nh++) {
        const uint32_t b0 = smem_u32(sWr + (1 + nh) * SLOT);
#pragma unroll
        for (int c = 0; c < 2; c++)
#pragma unroll
          for (int s = 0; s < 4; s++)
            umma_bf16(tmem + 128 + nh * 128, umma_desc_sw128(a0 + c * 16384 + s * 32), umma_desc_sw128(b0 + c * 16384 + s * 32), idesc,
                      (c | s) ? 1u : 0u);
      }
      umma_commit(smem_u32(&mbar));
    }
    __syncwarp();
    mbar_wait(smem_u32(&mbar), parity);
    parity ^= 1u;
    tc_fence_after();

    if (g.Wqkv) {  // W1 is dead: prefetch the next layer's Wq / Wk into slots 1 / 2 (lands during GELU, GEMM3, LN2)
      stage_chunk(sWr + 1 * SLOT, g.Wqkv, D, 0, 128, tid);
      stage_chunk(sWr + 1 * SLOT + 16384, g.Wqkv, D, 64, 128, tid);
      stage_chunk(sWr + 2 * SLOT, g.Wqkv + 128 * D, D, 0, 128, tid);
      stage_chunk(sWr + 2 * SLOT + 16384, g.Wqkv + 128 * D, D, 64, 128, tid);
    }
    // ---- epilogue 2: hidden = GELU(acc2 + b1) -> bf16 operand layout [128 x 256] in sH (4 K-chunks) -----------------------
#pragma unroll 1
    for (int cc = 0; cc < 2; cc++) {
      const int h0 = qtr * 64 + cc * 32;   // hidden column
      float v[32];
      tmem_ld32(tlane + 128 + h0, v);
      const float4* bp = reinterpret_cast<const float4*>(g.b1 + h0);
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        float4 b4 = __ldg(bp + (i >> 2));
        pk[i >> 1] = pack_bf16(gelu_as(v[i] + b4.x), gelu_as(v[i + 1] + b4.y));
        pk[(i >> 1) + 1] = pack_bf16(gelu_as(v[i + 2] + b4.z), gelu_as(v[i + 3] + b4.w));
      }
      const int kc = h0 >> 6;             // K-chunk of GEMM3's A operand
      const int j0 = (h0 & 63) >> 3;      // first 16-byte piece inside the chunk row
#pragma unroll
      for (int q = 0; q < 4; q++)
        *reinterpret_cast<int4*>(sH + kc * 16384 + lrow * 128 + (((j0 + q) ^ (lrow & 7)) << 4)) =
            make_int4((int)pk[4 * q], (int)pk[4 * q + 1], (int)pk[4 * q + 2], (int)pk[4 * q + 3]);
    }
    cp_async_wait_all();   // W2[:, 128:256] refill of slot 0
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    // ---- GEMM3: acc3[0..127] = hidden . W2^T  (K = 256: slot3 holds k 0..127, slot0 holds k 128..255) -------------------------
    if (tid == 0) {
      const uint32_t a0 = smem_u32(sH);
#pragma unroll
      for (int c = 0; c < 4; c++) {
        const uint32_t b0 = smem_u32(sWr + (c < 2 ? 3 : 0) * SLOT) + (c & 1) * 16384;
#pragma unroll
        for (int s = 0; s < 4; s++)
          umma_bf16(tmem, umma_desc_sw128(a0 + c * 16384 + s * 32), umma_desc_sw128(b0 + s * 32), idesc, (c | s) ? 1u : 0u);
      }
      umma_commit(smem_u32(&mbar));
    }
    __syncwarp();
    mbar_wait(smem_u32(&mbar), parity);
    parity ^= 1u;
    tc_fence_after();

    // ---- epilogue 3: y = LN2(x1 + acc3 + b2) -> fp32 staging tile -> coalesced rows -------------------------------------------
    {
      float sum = 0.f, sq = 0.f;
      float v[32], r[32];
      tmem_ld32(tlane + c0, v);
      tmem_ld32(tlane + 384 + c0, r);
#pragma unroll
      for (int q = 0; q < 8; q++) {
        float4 b4 = __ldg(reinterpret_cast<const float4*>(g.b2 + c0) + q);
        float a0 = v[4 * q] + b4.x + r[4 * q], a1 = v[4 * q + 1] + b4.y + r[4 * q + 1];
        float a2 = v[4 * q + 2] + b4.z + r[4 * q + 2], a3 = v[4 * q + 3] + b4.w + r[4 * q + 3];
        t[4 * q] = a0;
        t[4 * q + 1] = a1;
        t[4 * q + 2] = a2;
        t[4 * q + 3] = a3;
        sum += (a0 + a1) + (a2 + a3);
        sq += (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3);
      }
      red[qtr][lrow][0] = sum;
      red[qtr][lrow][1] = sq;
    }
    tc_fence_before();
    __syncthreads();
    {
      const float sum = (red[0][lrow][0] + red[1][lrow][0]) + (red[2][lrow][0] + red[3][lrow][0]);
      const float sq = (red[0][lrow][1] + red[1][lrow][1]) + (red[2][lrow][1] + red[3][lrow][1]);
      const float mean = sum * (1.0f / D);
      const float rstd = rsqrtf(fmaxf(sq * (1.0f / D) - mean * mean, 0.f) + g.eps);
#pragma unroll
      for (int q = 0; q < 8; q++) {
        float4 g4 = __ldg(reinterpret_cast<const float4*>(g.g2 + c0) + q), e4 = __ldg(reinterpret_cast<const float4*>(g.be2 + c0) + q);
        float4 o;
        o.x = (t[4 * q] - mean) * rstd * g4.x + e4.x;
        o.y = (t[4 * q + 1] - mean) * rstd * g4.y + e4.y;
        o.z = (t[4 * q + 2] - mean) * rstd * g4.z + e4.z;
        o.w = (t[4 * q + 3] - mean) * rstd * g4.w + e4.w;
        int ch = (c0 >> 2) + q;
        *reinterpret_cast<float4*>(sH + (size_t)lrow * 512 + ((ch ^ (lrow & 31)) << 4)) = o;  // hidden is dead (GEMM3 retired)
        t[4 * q] = o.x;
        t[4 * q + 1] = o.y;
        t[4 * q + 2] = o.z;
        t[4 * q + 3] = o.w;
      }
    }
    __syncthreads();
    for (int idx = tid; idx < TM * 32; idx += NTHR) {
      int r = idx >> 5, ch = idx & 31;
      int gr = sRow[r];
      if (gr >= 0)
        *reinterpret_cast<float4*>(g.y + (size_t)gr * D + ch * 4) =
            *reinterpret_cast<const float4*>(sH + (size_t)r * 512 + ((ch ^ (r & 31)) << 4));
    }
    if (g.Wqkv) {
      // ---- fused tail: q,k = (y + pos_next) . Wq^T / Wk^T, v = y . Wv^T for the next layer, scattered into its slot order ----
      stage_chunk(sWr + 3 * SLOT, g.Wqkv + 256 * D, D, 0, 128, tid);        // Wv -> slot 3 (W2[:, 0:128] is dead)
      stage_chunk(sWr + 3 * SLOT + 16384, g.Wqkv + 256 * D, D, 64, 128, tid);
      __syncthreads();  // y staging tile fully stored -> sH becomes the two A operands
      {
        const int tok = sRow[lrow];
        float pe[32];
#pragma unroll
        for (int i = 0; i < 32; i++) pe[i] = 0.f;
        const int axis = c0 / g.posL;  // posL % 32 == 0 (checked on the host): the 32-column span lies inside one axis
        if (tok >= 0 && axis < g.pos_ndim) {
          const int cv = (g.next_pos_code[tok] >> (8 * axis)) & 255;
          const float4* tp = reinterpret_cast<const float4*>(g.pos_tab + ((size_t)axis * g.pos_maxw + cv) * g.posL + (c0 - axis * g.posL));
#pragma unroll
          for (int q = 0; q < 8; q++) {
            float4 p4 = __ldg(tp + q);
            pe[4 * q] = p4.x;
            pe[4 * q + 1] = p4.y;
            pe[4 * q + 2] = p4.z;
            pe[4 * q + 3] = p4.w;
          }
        }
        const int kc = c0 >> 6, j0 = (c0 & 63) >> 3;
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const float* sy = &t[q * 8];
          const float* sp = &pe[q * 8];
          const uint32_t off = kc * 16384 + lrow * 128 + (((j0 + q) ^ (lrow & 7)) << 4);
          *reinterpret_cast<int4*>(sH + off) = make_int4((int)pack_bf16(sy[0] + sp[0], sy[1] + sp[1]), (int)pack_bf16(sy[2] + sp[2], sy[3] + sp[3]),
                                                         (int)pack_bf16(sy[4] + sp[4], sy[5] + sp[5]), (int)pack_bf16(sy[6] + sp[6], sy[7] + sp[7]));
          *reinterpret_cast<int4*>(sH + SLOT + off) = make_int4((int)pack_bf16(sy[0], sy[1]), (int)pack_bf16(sy[2], sy[3]),
                                                                (int)pack_bf16(sy[4], sy[5]), (int)pack_bf16(sy[6], sy[7]));
        }
      }
      cp_async_wait_all();
      fence_async_smem();
      tc_fence_before();
      __syncthreads();
      tc_fence_after();
      if (tid == 0) {
        const uint32_t ap = smem_u32(sH), ax = smem_u32(sH + SLOT);
#pragma unroll
        for (int nt = 0; nt < 3; nt++) {
          const uint32_t a0 = nt < 2 ? ap : ax, b0 = smem_u32(sWr + (1 + nt) * SLOT);
#pragma unroll
          for (int c = 0; c < 2; c++)
#pragma unroll
            for (int s2 = 0; s2 < 4; s2++)
              umma_bf16(tmem + nt * 128, umma_desc_sw128(a0 + c * 16384 + s2 * 32), umma_desc_sw128(b0 + c * 16384 + s2 * 32), idesc,
                        (c | s2) ? 1u : 0u);
        }
        umma_commit(smem_u32(&mbar));
      }
      __syncwarp();
      mbar_wait(smem_u32(&mbar), parity);
      parity ^= 1u;
      tc_fence_after();
      // epilogue 4: + bias -> fp16 -> staging tiles (q -> sH[0:32K], k -> sH[32K:64K], v -> ring slot 0) -> rows in next-slot order
#pragma unroll 1
      for (int nt = 0; nt < 3; nt++) {
        float v[32];
        tmem_ld32(tlane + nt * 128 + c0, v);
        const float4* bp = reinterpret_cast<const float4*>(g.bqkv + nt * 128 + c0);
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          float4 b4 = __ldg(bp + (i >> 2));
          __half2 h0 = __floats2half2_rn(v[i] + b4.x, v[i + 1] + b4.y), h1 = __floats2half2_rn(v[i + 2] + b4.z, v[i + 3] + b4.w);
          pk[i >> 1] = *reinterpret_cast<uint32_t*>(&h0);
          pk[(i >> 1) + 1] = *reinterpret_cast<uint32_t*>(&h1);
        }
        uint8_t* st = nt < 2 ? sH + nt * SLOT : sWr;
#pragma unroll
        for (int q = 0; q < 4; q++) {
          int ch = (c0 >> 3) + q;  // 16-byte chunk (8 halfs) of the 128-wide row, XOR-swizzled by row
          *reinterpret_cast<int4*>(st + (size_t)lrow * 256 + ((ch ^ (lrow & 15)) << 4)) =
              make_int4((int)pk[4 * q], (int)pk[4 * q + 1], (int)pk[4 * q + 2], (int)pk[4 * q + 3]);
        }
      }
      tc_fence_before();
      __syncthreads();
      for (int idx = tid; idx < TM * 48; idx += NTHR) {
        int r = idx / 48, cc = idx % 48;
        int nt = cc >> 4, ch = cc & 15;
        int tok = sRow[r];
        if (tok >= 0) {
          const uint8_t* st = nt < 2 ? sH + nt * SLOT : sWr;
          *reinterpret_cast<int4*>(g.qkv_out + (size_t)(g.next_slot ? g.next_slot[tok] : tok) * 384 + nt * 128 + ch * 8) =
              *reinterpret_cast<const int4*>(st + (size_t)r * 256 + ((ch ^ (r & 15)) << 4));
        }
      }
    }
    tc_fence_before();
    __syncthreads();  // staging tile / TMEM fully consumed before the next tile re-stages
  }
  if (warp == 0) tmem_dealloc(tmem, 512);
}

}  // namespace

int sstb_sra_chain_bf16(sstb200_ctx* c, const sstb200_sra_layer* L, const __nv_bfloat16* att, const int32_t* row_map, const float* x,
                        float* y, int n_cap, const int32_t* n_dev, const sstb200_sra_layer* next, const sstb200_sra_plan* next_plan,
                        void* next_qkv) {
  ChainArgs g;
  memset(&g, 0, sizeof(g));
  if (next && next_plan && next_qkv) {
    if (next_plan->pos_L % 32 != 0 || !next_plan->pos_code)
      return sstb_fail(c, SSTB_ERR_UNSUPPORTED, "fused QKV tail needs pos_L %% 32 == 0 and the next plan's pos_code");
    g.Wqkv = (const __nv_bfloat16*)next->in_proj_w_bf16;
    g.bqkv = next->in_proj_b;
    g.next_slot = nullptr;   // q|k|v rows stay in flat token order (the attention kernel gathers its windows)
    g.next_pos_code = next_plan->pos_code;
    g.pos_tab = next_plan->pos_table;
    g.posL = next_plan->pos_L;
    g.pos_maxw = next_plan->pos_maxw;
    g.pos_ndim = next_plan->pos_ndim;
    g.qkv_out = (__half*)next_qkv;
  }
  g.att = att;
  g.row_map = row_map;
  g.x = x;
  g.y = y;
  g.Wo = (const __nv_bfloat16*)L->out_proj_w_bf16;
  g.W1 = (const __nv_bfloat16*)L->lin1_w_bf16;
  g.W2 = (const __nv_bfloat16*)L->lin2_w_bf16;
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
  size_t smem = 6 * (size_t)SLOT + 1024;
  static SmemAttr sa;
  CUDA_TRY(c, ensure_smem(c, sa, sra_chain_kernel, smem));
  int tiles_cap = (n_cap + TM - 1) / TM;
  int grid = c->num_sms < tiles_cap ? c->num_sms : tiles_cap;
  CUDA_TRY(c, launch_pdl(sra_chain_kernel, dim3(grid), dim3(NTHR), smem, c->stream, g));
  return SSTB_OK;
}

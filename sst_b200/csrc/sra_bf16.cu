// 16-bit tensor-core path of the SRA encoder layer (precision 'bf16' of the API; operands are IEEE fp16 - same tensor
// throughput, 8x smaller operand rounding than bf16, and the reference's own mixed-precision setting is fp16,
// configs/sst_refactor/sst_waymoD5_1x_3class_8heads_v2.py:82): tcgen05.mma (UMMA) GEMMs with TMEM accumulators.
//
// Every dense op of the layer (QKV projection, attention out-projection, FFN1, FFN2) is one launch of the same
// kernel template:  C[128-row tile, NT] = A[128, K] . W[NT, K]^T  with the whole K extent staged once in shared
// memory (K <= 256 on this path, so there is no K pipeline to manage), accumulated in TMEM by ONE thread issuing
// K/16 tcgen05.mma instructions, and a fused epilogue read back with tcgen05.ld (one thread per output row):
//     EPI_BF16      + bias, optional GELU, store bf16                       (QKV, FFN1)
//     EPI_RES_LN    + bias + fp32 residual, LayerNorm over the row (two TMEM passes), store fp32 (+ bf16 copy)
// A-operand prologues convert on the fly: fp32 rows (+ positional embedding from the per-axis table) -> bf16, or
// plain bf16 rows; both write the canonical K-major SWIZZLE_128B layout the UMMA descriptors expect.
// fp32 stays the type of the residual stream, softmax, LayerNorm and all accumulation.
#include <stdarg.h>
#include <cuda_fp16.h>
#include "sra.cuh"
#include "sra_attn.cuh"
#include "umma.cuh"

namespace {

// exact-form GELU (x * Phi(x)) with erf from Abramowitz-Stegun 7.1.26 (|err| <= 1.5e-7): ~12 instructions instead of
// erff's ~30, same accuracy class as fp32 erff for an output that is rounded to bf16 anyway.
__device__ __forceinline__ float gelu_fast(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));  // MUFU.RCP (__frcp_rn is a ~25-instruction subroutine)
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float e = 1.0f - p * t * __expf(-z * z);   // erf(|x|/sqrt2)
  return 0.5f * x * (1.0f + copysignf(e, x));
}

__device__ __forceinline__ uint32_t pack_f16(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

enum { PRO_BF16 = 0, PRO_F32 = 1 };
enum { EPI_BF16 = 0, EPI_BF16_GELU = 1, EPI_RES_LN = 2, EPI_F16 = 3 };

struct GemmArgs {
  const void* A;        // [M, lda] bf16 or fp32; tile rows are consecutive rows of A
  int lda;
  const __half* W;  // [N_total, K] bf16
  const float* bias;    // [N_total]
  int M_cap;
  const int32_t* M_dev;
  // prologue
  const float* pos_tab;
  const int32_t* pos_code;
  int posL, pos_maxw, pos_ndim, pos_ntiles;  // add pos for n-tile < pos_ntiles
  int ny;                                    // number of n tiles
  // epilogue
  const int32_t* out_row_map;  // nullable: output / residual row of tile row i is out_row_map[i] (scatter), else i
  __half* out_h16;  // [M, ldo]
  int ldo;
  const float* res;     // [M, NT] fp32 residual (EPI_RES_LN)
  const float *gamma, *beta;
  float eps;
  float* out_f32;       // [M, NT]
};

constexpr int TILE_M = 128;

// Persistent over (row tile, n tile) items.  Per item: stage W and A (16-byte chunks, 8 loads in flight per thread) in
// the K-major SWIZZLE_128B layout -> one thread issues the MMAs -> commit/mbarrier -> epilogue.  The epilogue goes through
// shared memory (the operand buffers are free once the MMA has completed) so that every global access is a coalesced row
// segment: thread-per-row TMEM reads meet warp-per-row global traffic in an XOR-swizzled staging tile.
template <int K, int NT, int PRO, int EPI>
__global__ void __launch_bounds__(256) umma_gemm_kernel(GemmArgs g) {
  pdl_wait();
  pdl_launch();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = base;                           // K/64 chunks x 128 rows x 128 B
  uint8_t* sW = sA + (size_t)TILE_M * K * 2;    // K/64 chunks x NT rows x 128 B
  uint8_t* sE = base;                           // epilogue staging (aliases the operands)
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tmem_slot;
  __shared__ float red[2][TILE_M][2];
  __shared__ int sRow[TILE_M];

  const int tid = threadIdx.x, warp = tid >> 5;
  const int M = g.M_dev ? *g.M_dev : g.M_cap;
  const int n_items = ((M + TILE_M - 1) / TILE_M) * g.ny;
  if ((int)blockIdx.x >= n_items) return;  // uniform per CTA: nothing allocated yet

  if (warp == 0) tmem_alloc(&tmem_slot, NT);
  if (tid == 0) {
    mbar_init(smem_u32(&mbar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  constexpr int CH = K / 8;  // 16-byte chunks per operand row
  constexpr int NTH = 256, UNR = 8;
  uint32_t parity = 0;
  uint32_t tmem = 0;
  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int ytile = item % g.ny;
    const int row0 = (item / g.ny) * TILE_M;
    const int n0 = ytile * NT;
    if (tid < TILE_M) {
      int gr = row0 + tid;
      sRow[tid] = gr < M ? (g.out_row_map ? g.out_row_map[gr] : gr) : -1;
    }
    // ---- stage W tile (rows n0..n0+NT of W[., K]) ----------------------------------------------------------------
    {
      const __half* wsrc = g.W + (size_t)n0 * K;
      for (int i0 = tid; i0 < NT * CH; i0 += NTH * UNR) {
        int4 v[UNR];
#pragma unroll
        for (int u = 0; u < UNR; u++) {
          int idx = i0 + u * NTH;
          if (idx < NT * CH) v[u] = __ldg(reinterpret_cast<const int4*>(wsrc) + idx);
        }
#pragma unroll
        for (int u = 0; u < UNR; u++) {
          int idx = i0 + u * NTH;
          if (idx < NT * CH) {
            int r = idx / CH, j = idx % CH;
            int c = j >> 3, jj = j & 7;
            *reinterpret_cast<int4*>(sW + (size_t)c * NT * 128 + r * 128 + ((jj ^ (r & 7)) << 4)) = v[u];
          }
        }
      }
    }
    // ---- stage A tile -------------------------------------------------------------------------------------------------
    const bool add_pos = (PRO == PRO_F32) && g.pos_tab != nullptr && ytile < g.pos_ntiles;
    if (PRO == PRO_BF16) {
      for (int i0 = tid; i0 < TILE_M * CH; i0 += NTH * UNR) {
        int4 v[UNR];
#pragma unroll
        for (int u = 0; u < UNR; u++) {
          int idx = i0 + u * NTH;
          int r = idx / CH, j = idx % CH;
          v[u] = make_int4(0, 0, 0, 0);
          if (idx < TILE_M * CH && row0 + r < M)
            v[u] = *reinterpret_cast<const int4*>((const __half*)g.A + (size_t)(row0 + r) * g.lda + j * 8);
        }
#pragma unroll
        for (int u = 0; u < UNR; u++) {
          int idx = i0 + u * NTH;
          if (idx < TILE_M * CH) {
            int r = idx / CH, j = idx % CH;
            int c = j >> 3, jj = j & 7;
            *reinterpret_cast<int4*>(sA + (size_t)c * TILE_M * 128 + r * 128 + ((jj ^ (r & 7)) << 4)) = v[u];
          }
        }
      }
    } else {
      constexpr int UF = 4;
      for (int i0 = tid; i0 < TILE_M * CH; i0 += NTH * UF) {
        float4 f0[UF], f1[UF];
        int code[UF];
#pragma unroll
        for (int u = 0; u < UF; u++) {
          int idx = i0 + u * NTH;
          int r = idx / CH, j = idx % CH;
          f0[u] = f1[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          code[u] = 0;
          if (idx < TILE_M * CH && row0 + r < M) {
            const float* ap = (const float*)g.A + (size_t)(row0 + r) * g.lda + j * 8;
            f0[u] = *reinterpret_cast<const float4*>(ap);
            f1[u] = *reinterpret_cast<const float4*>(ap + 4);
            if (add_pos) code[u] = g.pos_code[row0 + r];
          }
        }
#pragma unroll
        for (int u = 0; u < UF; u++) {
          int idx = i0 + u * NTH;
          if (idx < TILE_M * CH) {
            int r = idx / CH, j = idx % CH;
            float f[8] = {f0[u].x, f0[u].y, f0[u].z, f0[u].w, f1[u].x, f1[u].y, f1[u].z, f1[u].w};
            if (add_pos && row0 + r < M) {
              const int k0 = j * 8;
              const int axis = k0 / g.posL;  // posL % 8 == 0 (checked on the host): the chunk lies inside one axis
              if (axis < g.pos_ndim) {
                const int cv = (code[u] >> (8 * axis)) & 255;
                const float4* tp = reinterpret_cast<const float4*>(g.pos_tab + ((size_t)axis * g.pos_maxw + cv) * g.posL + (k0 - axis * g.posL));
                float4 p0 = __ldg(tp), p1 = __ldg(tp + 1);
                f[0] += p0.x; f[1] += p0.y; f[2] += p0.z; f[3] += p0.w;
                f[4] += p1.x; f[5] += p1.y; f[6] += p1.z; f[7] += p1.w;
              }
            }
            int4 v;
            v.x = (int)pack_f16(f[0], f[1]);
            v.y = (int)pack_f16(f[2], f[3]);
            v.z = (int)pack_f16(f[4], f[5]);
            v.w = (int)pack_f16(f[6], f[7]);
            int c = j >> 3, jj = j & 7;
            *reinterpret_cast<int4*>(sA + (size_t)c * TILE_M * 128 + r * 128 + ((jj ^ (r & 7)) << 4)) = v;
          }
        }
      }
    }
    // generic-proxy smem writes -> visible to the tensor core (async proxy); TMEM address visible to all
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    tmem = tmem_slot;

    // ---- MMA issue: one thread ------------------------------------------------------------------------------------------
    if (tid == 0) {
      const uint32_t idesc = umma_idesc_f16(TILE_M, NT);
      const uint32_t a0 = smem_u32(sA), w0 = smem_u32(sW);
#pragma unroll
      for (int c = 0; c < K / 64; c++) {
#pragma unroll
        for (int s = 0; s < 4; s++) {
          uint64_t ad = umma_desc_sw128(a0 + c * TILE_M * 128 + s * 32);
          uint64_t bd = umma_desc_sw128(w0 + c * NT * 128 + s * 32);
          umma_f16(tmem, ad, bd, idesc, (c | s) ? 1u : 0u);
        }
      }
      umma_commit(smem_u32(&mbar));  // implicit tcgen05.fence::before_thread_sync
    }
    __syncwarp();
    mbar_wait(smem_u32(&mbar), parity);
    parity ^= 1u;
    tc_fence_after();

    // ---- epilogue ----------------------------------------------------------------------------------------------------------
    const int half = warp >> 2;                      // warps 0-3: columns [0,NT/2), warps 4-7: [NT/2,NT)
    const int lrow = (warp & 3) * 32 + (tid & 31);   // TMEM lane == row inside the tile
    const uint32_t tlane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    constexpr int CB = NT / 2;
    const int cbeg = half * CB;
    if (EPI == EPI_BF16 || EPI == EPI_BF16_GELU || EPI == EPI_F16) {
      // thread-per-row: bias (+GELU), pack, into the bf16 staging tile [128][NT] (16-byte chunks XOR-swizzled by row)
      constexpr int ECH = NT / 8;
#pragma unroll 1
      for (int c0 = cbeg; c0 < cbeg + CB; c0 += 32) {
        float v[32];
        tmem_ld32(tlane + c0, v);
        uint32_t pk[16];
        const float4* bp = reinterpret_cast<const float4*>(g.bias + n0 + c0);
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          float4 b4 = __ldg(bp + (i >> 2));
          float a0 = v[i] + b4.x, a1 = v[i + 1] + b4.y, a2 = v[i + 2] + b4.z, a3 = v[i + 3] + b4.w;
          if (EPI == EPI_BF16_GELU) {
            a0 = gelu_fast(a0);
            a1 = gelu_fast(a1);
            a2 = gelu_fast(a2);
            a3 = gelu_fast(a3);
          }
          pk[i >> 1] = pack_f16(a0, a1);
          pk[(i >> 1) + 1] = pack_f16(a2, a3);
        }
#pragma unroll
        for (int q = 0; q < 4; q++) {
          int ch = (c0 >> 3) + q;
          *reinterpret_cast<int4*>(sE + (size_t)lrow * NT * 2 + ((ch ^ (lrow & (ECH - 1))) << 4)) =
              make_int4((int)pk[4 * q], (int)pk[4 * q + 1], (int)pk[4 * q + 2], (int)pk[4 * q + 3]);
        }
      }
      tc_fence_before();
      __syncthreads();
      for (int idx = tid; idx < TILE_M * ECH; idx += NTH) {
        int r = idx / ECH, ch = idx % ECH;
        int gr = sRow[r];
        if (gr >= 0)
          *reinterpret_cast<int4*>(g.out_h16 + (size_t)gr * g.ldo + n0 + ch * 8) =
              *reinterpret_cast<const int4*>(sE + (size_t)r * NT * 2 + ((ch ^ (r & (ECH - 1))) << 4));
      }
    } else {  // EPI_RES_LN : NT == row width, fp32 staging tile [128][NT] (16-byte chunks XOR-swizzled by row)
      constexpr int ECH = NT / 4;  // float4 chunks per row
      static_assert(ECH == 32, "LayerNorm epilogue is written for 128-wide rows");
      // 1. residual tile, coalesced
      for (int i0 = tid; i0 < TILE_M * ECH; i0 += NTH * UNR) {
        float4 v[UNR];
#pragma unroll
        for (int u = 0; u < UNR; u++) {
          int idx = i0 + u * NTH;
          int r = idx / ECH, ch = idx % ECH;
          int gr = sRow[r];
          v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (gr >= 0) v[u] = *reinterpret_cast<const float4*>(g.res + (size_t)gr * NT + ch * 4);
        }
#pragma unroll
        for (int u = 0; u < UNR; u++) {
          int idx = i0 + u * NTH;
          int r = idx / ECH, ch = idx % ECH;
          *reinterpret_cast<float4*>(sE + (size_t)r * NT * 4 + ((ch ^ (r & 31)) << 4)) = v[u];
        }
      }
      __syncthreads();
      // 2. thread-per-row: t = acc + bias + residual (kept in the staging tile), row statistics
      float sum = 0.f, sq = 0.f;
#pragma unroll 1
      for (int c0 = cbeg; c0 < cbeg + CB; c0 += 32) {
        float v[32];
        tmem_ld32(tlane + c0, v);
#pragma unroll
        for (int q = 0; q < 8; q++) {
          int ch = (c0 >> 2) + q;
          float4* sp = reinterpret_cast<float4*>(sE + (size_t)lrow * NT * 4 + ((ch ^ (lrow & 31)) << 4));
          float4 r4 = *sp;
          float4 t, b4 = __ldg(reinterpret_cast<const float4*>(g.bias + c0) + q);
          t.x = v[4 * q] + b4.x + r4.x;
          t.y = v[4 * q + 1] + b4.y + r4.y;
          t.z = v[4 * q + 2] + b4.z + r4.z;
          t.w = v[4 * q + 3] + b4.w + r4.w;
          *sp = t;
          sum += (t.x + t.y) + (t.z + t.w);
          sq += (t.x * t.x + t.y * t.y) + (t.z * t.z + t.w * t.w);
        }
      }
      red[half][lrow][0] = sum;
      red[half][lrow][1] = sq;
      tc_fence_before();
      __syncthreads();
      sum = red[0][lrow][0] + red[1][lrow][0];
      sq = red[0][lrow][1] + red[1][lrow][1];
      const float mean = sum * (1.0f / NT);
      const float var = fmaxf(sq * (1.0f / NT) - mean * mean, 0.f);
      const float rstd = rsqrtf(var + g.eps);
#pragma unroll 1
      for (int c0 = cbeg; c0 < cbeg + CB; c0 += 4) {
        int ch = c0 >> 2;
        float4* sp = reinterpret_cast<float4*>(sE + (size_t)lrow * NT * 4 + ((ch ^ (lrow & 31)) << 4));
        float4 t = *sp;
        const float4 g4 = __ldg(reinterpret_cast<const float4*>(g.gamma + c0)), b4 = __ldg(reinterpret_cast<const float4*>(g.beta + c0));
        t.x = (t.x - mean) * rstd * g4.x + b4.x;
        t.y = (t.y - mean) * rstd * g4.y + b4.y;
        t.z = (t.z - mean) * rstd * g4.z + b4.z;
        t.w = (t.w - mean) * rstd * g4.w + b4.w;
        *sp = t;
      }
      __syncthreads();
      // 3. coalesced stores: fp32 rows (+ bf16 copy for the next GEMM's A operand)
      for (int idx = tid; idx < TILE_M * ECH; idx += NTH) {
        int r = idx / ECH, ch = idx % ECH;
        int gr = sRow[r];
        if (gr >= 0)
          *reinterpret_cast<float4*>(g.out_f32 + (size_t)gr * NT + ch * 4) =
              *reinterpret_cast<const float4*>(sE + (size_t)r * NT * 4 + ((ch ^ (r & 31)) << 4));
      }
      if (g.out_h16) {
        for (int idx = tid; idx < TILE_M * (NT / 8); idx += NTH) {
          int r = idx / (NT / 8), c8 = idx % (NT / 8);
          int gr = sRow[r];
          if (gr >= 0) {
            float4 a = *reinterpret_cast<const float4*>(sE + (size_t)r * NT * 4 + (((2 * c8) ^ (r & 31)) << 4));
            float4 b = *reinterpret_cast<const float4*>(sE + (size_t)r * NT * 4 + (((2 * c8 + 1) ^ (r & 31)) << 4));
            *reinterpret_cast<int4*>(g.out_h16 + (size_t)gr * g.ldo + c8 * 8) =
                make_int4((int)pack_f16(a.x, a.y), (int)pack_f16(a.z, a.w), (int)pack_f16(b.x, b.y), (int)pack_f16(b.z, b.w));
          }
        }
      }
    }
    // TMEM fully read and the staging tile stored before the next item overwrites operands / accumulators
    tc_fence_before();
    __syncthreads();
  }  // item loop
  if (warp == 0) tmem_dealloc(tmem, NT);
}

template <int K, int NT, int PRO, int EPI>
int launch_umma(sstb200_ctx* c, const GemmArgs& g, int n_tiles_y) {
  size_t ops = (size_t)TILE_M * K * 2 + (size_t)NT * K * 2;
  size_t stage = (EPI == EPI_RES_LN) ? (size_t)TILE_M * NT * 4 : (size_t)TILE_M * NT * 2;
  size_t smem = (ops > stage ? ops : stage) + 1024;
  auto kern = umma_gemm_kernel<K, NT, PRO, EPI>;
  static SmemAttr sa;
  CUDA_TRY(c, ensure_smem(c, sa, kern, smem));
  GemmArgs ga = g;
  ga.ny = n_tiles_y;
  int items_cap = ((g.M_cap + TILE_M - 1) / TILE_M) * n_tiles_y;
  int per_sm = (int)((220 * 1024) / (smem + 6144));
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 512 / NT) per_sm = 512 / NT;   // TMEM columns
  int grid = c->num_sms * per_sm;
  if (grid > items_cap) grid = items_cap;
  launch_pdl(kern, dim3(grid), dim3(256), (size_t)(smem), c->stream, ga);
  CUDA_TRY(c, cudaGetLastError());
  return SSTB_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
int sstb_sra_layer_bf16(sstb200_ctx* c, const sstb200_sra_layer* L, const sstb200_sra_plan* P, const float* x, float* y,
                        int n_cap, const int32_t* n_dev) {
  const int d = L->d_model, ff = L->dim_ff;
  if (d != 128 || ff != 256 || !L->post_norm || L->norm1_mean || L->act != 2)
    return sstb_fail(c, SSTB_ERR_UNSUPPORTED,
                     "bf16 tensor-core path is built for d_model=128, dim_ff=256, post-norm LayerNorm, gelu (got d=%d ff=%d)", d, ff);
  if (!L->in_proj_w_f16 || !L->out_proj_w_f16 || !L->lin1_w_f16 || !L->lin2_w_f16)
    return sstb_fail(c, SSTB_ERR_ARG, "the tensor-core path needs the *_w_f16 weight copies");
  if (P->pos_table && P->pos_L % 8 != 0)
    return sstb_fail(c, SSTB_ERR_UNSUPPORTED, "bf16 path needs the per-axis positional length (%d) to be a multiple of 8", P->pos_L);
  __half* qkv = arena_alloc<__half>(c, (size_t)n_cap * 3 * d);
  __half* att = arena_alloc<__half>(c, (size_t)n_cap * d);
  __half* x1b = arena_alloc<__half>(c, (size_t)n_cap * d);
  __half* hid = arena_alloc<__half>(c, (size_t)n_cap * ff);
  float* x1 = arena_alloc<float>(c, (size_t)n_cap * d);
  if (!qkv || !att || !x1b || !hid || !x1) return sstb_fail(c, SSTB_ERR_WORKSPACE, "sra bf16 layer: arena too small");
  int rc;
  static int dbg_skip = -1;  // SSTB200_DEBUG_SKIP bitmask (timing experiments only): 1 QKV, 2 attention, 4 chain
  if (dbg_skip < 0) {
    const char* e = getenv("SSTB200_DEBUG_SKIP");
    dbg_skip = e ? atoi(e) : 0;
  }
  // tensor-core attention path: plain scaled-dot-product, 8 heads x 16, windows <= 144 tokens
  const bool tc_attn = !L->tau && L->nhead == 8 && P->max_window_tokens > 0 && P->max_window_tokens <= ATT_MAXT &&
                       P->num_windows_dev && P->win_batch;
  GemmArgs g;
  memset(&g, 0, sizeof(g));
  g.M_cap = n_cap;
  g.M_dev = n_dev;
  // 1. QKV projection: q,k from bf16(x + pos), v from bf16(x)
  g.A = x;
  g.lda = d;
  g.W = (const __half*)L->in_proj_w_f16;
  g.bias = L->in_proj_b;
  g.pos_tab = P->pos_table;
  g.pos_code = P->pos_code;
  g.posL = P->pos_L;
  g.pos_maxw = P->pos_maxw;
  g.pos_ndim = P->pos_ndim;
  g.pos_ntiles = 2;
  g.out_h16 = qkv;
  g.ldo = 3 * d;
  g.out_row_map = nullptr;  // q|k|v, att and the residual stream all live in flat token order; attention gathers its windows
  // q/k/v for the tensor-core attention are written as fp16 (softmax logits need the mantissa: with bf16 q,k the logit
  // error dominates the layer's error budget); the SIMT fallback reads bf16.
  rc = (dbg_skip & 1) ? 0 : (tc_attn ? launch_umma<128, 128, PRO_F32, EPI_F16>(c, g, 3) : launch_umma<128, 128, PRO_F32, EPI_BF16>(c, g, 3));
  if (rc) return rc;
  // 2. ragged window attention (fp32 math on bf16 q/k/v)
  if (dbg_skip & 2)
    rc = 0;
  else if (tc_attn)
    rc = sstb_win_attn_batch(c, qkv, P->num_windows_dev, P->win_offsets, P->win_batch, P->tok_perm, att);
  else  // cosine attention / unbounded windows: SIMT kernel on the bf16 operands
    rc = sstb_win_attn<__half, __half>(c, qkv, d, L->nhead, n_cap, n_dev, P->win_offsets, P->tok_perm, P->tok_win,
                                                     L->tau, L->tau_n, L->tau_min, att);
  if (rc) return rc;
  // 3-5 fused: out-projection + LN1 + FFN1 + GELU + FFN2 + LN2 in one persistent tcgen05 kernel (csrc/sra_chain.cu)
  {
    static int use_chain = -1;
    if (use_chain < 0) {
      const char* e = getenv("SSTB200_CHAIN");
      use_chain = e ? atoi(e) : 1;   // 0: unfused GEMM launches (bring-up reference), else the warp-specialised TMA chain
    }
    if (dbg_skip & 4) return SSTB_OK;
    if (use_chain) return sstb_sra_chain2(c, L, att, x, y, n_cap, n_dev);
  }
  // 3. out-projection + residual + LayerNorm1 (unfused reference path, SSTB200_CHAIN=0)
  memset(&g, 0, sizeof(g));
  g.M_cap = n_cap;
  g.M_dev = n_dev;
  g.A = att;
  g.lda = d;
  g.W = (const __half*)L->out_proj_w_f16;
  g.bias = L->out_proj_b;
  g.res = x;
  g.gamma = L->norm1_w;
  g.beta = L->norm1_b;
  g.eps = L->norm_eps;
  g.out_f32 = x1;
  g.out_h16 = x1b;
  g.ldo = d;
  rc = launch_umma<128, 128, PRO_BF16, EPI_RES_LN>(c, g, 1);
  if (rc) return rc;
  // 4. FFN1 + GELU
  memset(&g, 0, sizeof(g));
  g.M_cap = n_cap;
  g.M_dev = n_dev;
  g.A = x1b;
  g.lda = d;
  g.W = (const __half*)L->lin1_w_f16;
  g.bias = L->lin1_b;
  g.out_h16 = hid;
  g.ldo = ff;
  rc = launch_umma<128, 128, PRO_BF16, EPI_BF16_GELU>(c, g, 2);
  if (rc) return rc;
  // 5. FFN2 + residual + LayerNorm2
  memset(&g, 0, sizeof(g));
  g.M_cap = n_cap;
  g.M_dev = n_dev;
  g.A = hid;
  g.lda = ff;
  g.W = (const __half*)L->lin2_w_f16;
  g.bias = L->lin2_b;
  g.res = x1;
  g.gamma = L->norm2_w;
  g.beta = L->norm2_b;
  g.eps = L->norm_eps;
  g.out_f32 = y;
  rc = launch_umma<256, 128, PRO_BF16, EPI_RES_LN>(c, g, 1);
  return rc;
}


// ------------------------------------------------------------------------------------------------
// whole encoder stack, bf16 tensor-core path: QKV(0) -> [attention(l) -> chain(l) (+ QKV(l+1) fused)] x L
// 2 launches per layer; q/k/v, attention output and the residual stream never take a detour.
// ------------------------------------------------------------------------------------------------
static bool layer_supported_tc(const sstb200_sra_layer* L, const sstb200_sra_plan* P) {
  return L->d_model == 128 && L->dim_ff == 256 && L->post_norm && !L->norm1_mean && L->act == 2 && !L->tau && L->nhead == 8 &&
         L->in_proj_w_f16 && L->out_proj_w_f16 && L->lin1_w_f16 && L->lin2_w_f16 && P->max_window_tokens > 0 &&
         P->max_window_tokens <= ATT_MAXT && P->num_windows_dev && P->win_batch && P->pos_table && P->pos_L % 32 == 0;
}

int sstb_sra_stack_bf16(sstb200_ctx* c, const sstb200_sra_layer* layers, int num_layers, const sstb200_sra_plan* plans, const float* x,
                        float* y, float* scratch, int n_cap, const int32_t* n_dev) {
  (void)scratch;
  for (int l = 0; l < num_layers; l++)
    if (!layer_supported_tc(&layers[l], &plans[l & 1])) return SSTB_ERR_UNSUPPORTED;  // caller falls back to per-layer calls
  const int d = 128;
  __half* qkv = arena_alloc<__half>(c, (size_t)n_cap * 3 * d);
  __half* att = arena_alloc<__half>(c, (size_t)n_cap * d);
  if (!qkv || !att) return sstb_fail(c, SSTB_ERR_WORKSPACE, "sra stack: arena too small");
  // QKV of layer 0 (stand-alone GEMM, fp32 x + pos -> fp16 rows in slot order of shift 0)
  {
    const sstb200_sra_layer* L = &layers[0];
    const sstb200_sra_plan* P = &plans[0];
    GemmArgs g;
    memset(&g, 0, sizeof(g));
    g.M_cap = n_cap;
    g.M_dev = n_dev;
    g.A = x;
    g.lda = d;
    g.W = (const __half*)L->in_proj_w_f16;
    g.bias = L->in_proj_b;
    g.pos_tab = P->pos_table;
    g.pos_code = P->pos_code;
    g.posL = P->pos_L;
    g.pos_maxw = P->pos_maxw;
    g.pos_ndim = P->pos_ndim;
    g.pos_ntiles = 2;
    g.out_h16 = qkv;
    g.ldo = 3 * d;
    int rc = launch_umma<128, 128, PRO_F32, EPI_F16>(c, g, 3);
    if (rc) return rc;
  }
  static int skip = -1;   // SSTB200_STACK_SKIP bitmask (timing experiments only): 2 attention, 4 chain
  if (skip < 0) skip = getenv("SSTB200_STACK_SKIP") ? atoi(getenv("SSTB200_STACK_SKIP")) : 0;
  const float* xin = x;
  for (int l = 0; l < num_layers; l++) {
    const sstb200_sra_plan* P = &plans[l & 1];
    int rc = (skip & 2) ? 0 : sstb_win_attn_batch(c, qkv, P->num_windows_dev, P->win_offsets, P->win_batch, P->tok_perm, att);
    if (rc) return rc;
    const bool has_next = l + 1 < num_layers;
    if (skip & 4) continue;
    // the chain reads the residual rows of a tile before it writes the same rows of y: in-place (xin == y) is safe
    rc = sstb_sra_chain2(c, &layers[l], att, xin, y, n_cap, n_dev, has_next ? &layers[l + 1] : nullptr,
                                has_next ? &plans[(l + 1) & 1] : nullptr, has_next ? qkv : nullptr);
    if (rc) return rc;
    xin = y;
  }
  return SSTB_OK;
}

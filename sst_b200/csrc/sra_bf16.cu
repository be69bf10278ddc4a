// bf16 tensor-core path of the SRA encoder layer: tcgen05.mma (UMMA) GEMMs with TMEM accumulators.
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
#include "sra.cuh"
#include "sra_attn.cuh"

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// UMMA shared-memory descriptor, K-major, SWIZZLE_128B: rows of 128 B (64 bf16), 8-row atoms of 1024 B.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);   // start address >> 4
  d |= (uint64_t)1 << 16;                   // leading byte offset (unused for swizzled K-major) = 1
  d |= (uint64_t)(1024 >> 4) << 32;         // stride byte offset: 8-row group pitch = 1024 B
  d |= (uint64_t)1 << 46;                   // descriptor version 1 (sm_100)
  d |= (uint64_t)2 << 61;                   // layout type SWIZZLE_128B
  return d;
}
// instruction descriptor: D=f32, A=B=bf16, both K-major, dense
__device__ __forceinline__ uint32_t umma_idesc(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t mbar_saddr) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(mbar_saddr) : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t saddr, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(saddr), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t saddr, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred P1;\n\tWAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\tbra WAIT_LOOP;\n\tDONE:\n\t}\n" ::"r"(saddr),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; i++) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

enum { PRO_BF16 = 0, PRO_F32 = 1 };
enum { EPI_BF16 = 0, EPI_BF16_GELU = 1, EPI_RES_LN = 2 };

struct GemmArgs {
  const void* A;        // [M, lda] bf16 or fp32
  int lda;
  const __nv_bfloat16* W;  // [N_total, K] bf16
  const float* bias;    // [N_total]
  int M_cap;
  const int32_t* M_dev;
  // prologue
  const float* pos_tab;
  const int32_t* pos_code;
  int posL, pos_maxw, pos_ndim, pos_ntiles;  // add pos for blockIdx.y < pos_ntiles
  // epilogue
  __nv_bfloat16* out_bf16;  // [M, ldo]
  int ldo;
  const float* res;     // [M, NT] fp32 residual (EPI_RES_LN)
  const float *gamma, *beta;
  float eps;
  float* out_f32;       // [M, NT]
};

constexpr int TILE_M = 128;

template <int K, int NT, int PRO, int EPI>
__global__ void __launch_bounds__(256) umma_gemm_kernel(GemmArgs g) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-B aligned operand tiles (SWIZZLE_128B atoms)
  uint8_t* base = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = base;                           // K/64 chunks x 128 rows x 128 B
  uint8_t* sW = sA + (size_t)TILE_M * K * 2;    // K/64 chunks x NT rows x 128 B
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tmem_slot;
  __shared__ float red[2][TILE_M][2];

  const int tid = threadIdx.x, warp = tid >> 5;
  const int M = g.M_dev ? *g.M_dev : g.M_cap;
  const int row0 = blockIdx.x * TILE_M;
  if (row0 >= M) return;  // uniform per CTA: nothing allocated yet
  const int n0 = blockIdx.y * NT;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_slot)), "r"((uint32_t)NT)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  if (tid == 0) {
    mbar_init(smem_u32(&mbar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }

  // ---- stage W tile (rows n0..n0+NT of W[., K]) and A tile: 16-byte chunks, 8 loads in flight per thread -------------
  constexpr int CH = K / 8;  // 16-byte chunks per row
  constexpr int NTH = 256, UNR = 8;
  {
    const __nv_bfloat16* wsrc = g.W + (size_t)n0 * K;
    for (int i0 = tid; i0 < NT * CH; i0 += NTH * UNR) {
      int4 v[UNR];
#pragma unroll
      for (int u = 0; u < UNR; u++) {
        int idx = i0 + u * NTH;
        if (idx < NT * CH) v[u] = __ldg(reinterpret_cast<const int4*>(wsrc) + idx);  // rows are K*2 bytes = CH chunks: contiguous
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
  const bool add_pos = (PRO == PRO_F32) && g.pos_tab != nullptr && (int)blockIdx.y < g.pos_ntiles;
  if (PRO == PRO_BF16) {
    for (int i0 = tid; i0 < TILE_M * CH; i0 += NTH * UNR) {
      int4 v[UNR];
#pragma unroll
      for (int u = 0; u < UNR; u++) {
        int idx = i0 + u * NTH;
        int r = idx / CH, j = idx % CH;
        v[u] = make_int4(0, 0, 0, 0);
        if (idx < TILE_M * CH && row0 + r < M)
          v[u] = *reinterpret_cast<const int4*>((const __nv_bfloat16*)g.A + (size_t)(row0 + r) * g.lda + j * 8);
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
#pragma unroll
            for (int e = 0; e < 8; e++) {
              int k = j * 8 + e;
              int axis = k / g.posL;
              if (axis < g.pos_ndim) {
                int cv = (code[u] >> (8 * axis)) & 255;
                f[e] += __ldg(&g.pos_tab[((size_t)axis * g.pos_maxw + cv) * g.posL + (k - axis * g.posL)]);
              }
            }
          }
          int4 v;
          v.x = (int)pack_bf16(f[0], f[1]);
          v.y = (int)pack_bf16(f[2], f[3]);
          v.z = (int)pack_bf16(f[4], f[5]);
          v.w = (int)pack_bf16(f[6], f[7]);
          int c = j >> 3, jj = j & 7;
          *reinterpret_cast<int4*>(sA + (size_t)c * TILE_M * 128 + r * 128 + ((jj ^ (r & 7)) << 4)) = v;
        }
      }
    }
  }
  // generic-proxy smem writes -> visible to the tensor core (async proxy); TMEM address visible to all
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = tmem_slot;

  // ---- MMA issue: one thread ----------------------------------------------------------------------------------------
  if (tid == 0) {
    const uint32_t idesc = umma_idesc(TILE_M, NT);
    const uint32_t a0 = smem_u32(sA), w0 = smem_u32(sW);
#pragma unroll
    for (int c = 0; c < K / 64; c++) {
#pragma unroll
      for (int s = 0; s < 4; s++) {
        uint64_t ad = umma_desc_sw128(a0 + c * TILE_M * 128 + s * 32);
        uint64_t bd = umma_desc_sw128(w0 + c * NT * 128 + s * 32);
        umma_bf16(tmem, ad, bd, idesc, (c | s) ? 1u : 0u);
      }
    }
    umma_commit(smem_u32(&mbar));  // implicit tcgen05.fence::before_thread_sync
  }
  __syncwarp();
  mbar_wait(smem_u32(&mbar), 0);
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");

  // ---- epilogue: thread t owns output row row0+t == TMEM lane t -------------------------------------------------------
  const int half = warp >> 2;                  // warps 0-3: columns [0,NT/2), warps 4-7: [NT/2,NT)
  const int lrow = (warp & 3) * 32 + (tid & 31);  // TMEM lane == row inside the tile
  const int grow = row0 + lrow;
  const bool live = grow < M;
  const uint32_t tlane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  constexpr int CB = NT / 2;
  const int cbeg = half * CB;
  if (EPI == EPI_BF16 || EPI == EPI_BF16_GELU) {
#pragma unroll 1
    for (int c0 = cbeg; c0 < cbeg + CB; c0 += 32) {
      float v[32];
      tmem_ld32(tlane + c0, v);
      if (live) {
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float a = v[i] + g.bias[n0 + c0 + i], b = v[i + 1] + g.bias[n0 + c0 + i + 1];
          if (EPI == EPI_BF16_GELU) {
            a = gelu_erf(a);
            b = gelu_erf(b);
          }
          pk[i >> 1] = pack_bf16(a, b);
        }
        int4* dst = reinterpret_cast<int4*>(g.out_bf16 + (size_t)grow * g.ldo + n0 + c0);
#pragma unroll
        for (int q = 0; q < 4; q++) dst[q] = make_int4((int)pk[4 * q], (int)pk[4 * q + 1], (int)pk[4 * q + 2], (int)pk[4 * q + 3]);
      }
    }
  } else {  // EPI_RES_LN : NT == row width
    float sum = 0.f, sq = 0.f;
#pragma unroll 1
    for (int c0 = cbeg; c0 < cbeg + CB; c0 += 32) {
      float v[32];
      tmem_ld32(tlane + c0, v);
      if (live) {
        const float4* rp = reinterpret_cast<const float4*>(g.res + (size_t)grow * NT + c0);
#pragma unroll
        for (int q = 0; q < 8; q++) {
          float4 r4 = rp[q];
          float t0 = v[4 * q] + g.bias[c0 + 4 * q] + r4.x, t1 = v[4 * q + 1] + g.bias[c0 + 4 * q + 1] + r4.y;
          float t2 = v[4 * q + 2] + g.bias[c0 + 4 * q + 2] + r4.z, t3 = v[4 * q + 3] + g.bias[c0 + 4 * q + 3] + r4.w;
          sum += (t0 + t1) + (t2 + t3);
          sq += (t0 * t0 + t1 * t1) + (t2 * t2 + t3 * t3);
        }
      }
    }
    red[half][lrow][0] = sum;
    red[half][lrow][1] = sq;
    __syncthreads();
    sum = red[0][lrow][0] + red[1][lrow][0];
    sq = red[0][lrow][1] + red[1][lrow][1];
    const float mean = sum * (1.0f / NT);
    const float var = fmaxf(sq * (1.0f / NT) - mean * mean, 0.f);
    const float rstd = rsqrtf(var + g.eps);
#pragma unroll 1
    for (int c0 = cbeg; c0 < cbeg + CB; c0 += 32) {
      float v[32];
      tmem_ld32(tlane + c0, v);
      if (live) {
        const float4* rp = reinterpret_cast<const float4*>(g.res + (size_t)grow * NT + c0);
        float4* op = reinterpret_cast<float4*>(g.out_f32 + (size_t)grow * NT + c0);
        uint32_t pk[16];
#pragma unroll
        for (int q = 0; q < 8; q++) {
          float4 r4 = rp[q];
          float o[4];
          float t[4] = {v[4 * q] + r4.x, v[4 * q + 1] + r4.y, v[4 * q + 2] + r4.z, v[4 * q + 3] + r4.w};
#pragma unroll
          for (int e = 0; e < 4; e++) {
            int cc = c0 + 4 * q + e;
            o[e] = (t[e] + g.bias[cc] - mean) * rstd * g.gamma[cc] + g.beta[cc];
          }
          op[q] = make_float4(o[0], o[1], o[2], o[3]);
          pk[2 * q] = pack_bf16(o[0], o[1]);
          pk[2 * q + 1] = pack_bf16(o[2], o[3]);
        }
        if (g.out_bf16) {
          int4* dst = reinterpret_cast<int4*>(g.out_bf16 + (size_t)grow * g.ldo + c0);
#pragma unroll
          for (int q = 0; q < 4; q++) dst[q] = make_int4((int)pk[4 * q], (int)pk[4 * q + 1], (int)pk[4 * q + 2], (int)pk[4 * q + 3]);
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"((uint32_t)NT) : "memory");
  }
}

template <int K, int NT, int PRO, int EPI>
int launch_umma(sstb200_ctx* c, const GemmArgs& g, int n_tiles_y) {
  size_t smem = (size_t)TILE_M * K * 2 + (size_t)NT * K * 2 + 1024;
  auto kern = umma_gemm_kernel<K, NT, PRO, EPI>;
  static bool attr_set = false;
  if (!attr_set) {
    CUDA_TRY(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  dim3 grid((g.M_cap + TILE_M - 1) / TILE_M, n_tiles_y);
  kern<<<grid, 256, smem, c->stream>>>(g);
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
  if (!L->in_proj_w_bf16 || !L->out_proj_w_bf16 || !L->lin1_w_bf16 || !L->lin2_w_bf16)
    return sstb_fail(c, SSTB_ERR_ARG, "bf16 path needs the *_w_bf16 weight copies");
  if (P->pos_table && P->pos_L % 8 != 0 && P->pos_L * P->pos_ndim != d)  // (no alignment requirement; sanity only)
    return sstb_fail(c, SSTB_ERR_ARG, "bad positional table");
  __nv_bfloat16* qkv = arena_alloc<__nv_bfloat16>(c, (size_t)n_cap * 3 * d);
  __nv_bfloat16* att = arena_alloc<__nv_bfloat16>(c, (size_t)n_cap * d);
  __nv_bfloat16* x1b = arena_alloc<__nv_bfloat16>(c, (size_t)n_cap * d);
  __nv_bfloat16* hid = arena_alloc<__nv_bfloat16>(c, (size_t)n_cap * ff);
  float* x1 = arena_alloc<float>(c, (size_t)n_cap * d);
  if (!qkv || !att || !x1b || !hid || !x1) return sstb_fail(c, SSTB_ERR_WORKSPACE, "sra bf16 layer: arena too small");
  int rc;
  GemmArgs g;
  memset(&g, 0, sizeof(g));
  g.M_cap = n_cap;
  g.M_dev = n_dev;
  // 1. QKV projection: q,k from bf16(x + pos), v from bf16(x)
  g.A = x;
  g.lda = d;
  g.W = (const __nv_bfloat16*)L->in_proj_w_bf16;
  g.bias = L->in_proj_b;
  g.pos_tab = P->pos_table;
  g.pos_code = P->pos_code;
  g.posL = P->pos_L;
  g.pos_maxw = P->pos_maxw;
  g.pos_ndim = P->pos_ndim;
  g.pos_ntiles = 2;
  g.out_bf16 = qkv;
  g.ldo = 3 * d;
  rc = launch_umma<128, 128, PRO_F32, EPI_BF16>(c, g, 3);
  if (rc) return rc;
  // 2. ragged window attention (fp32 math on bf16 q/k/v)
  if (!L->tau && L->nhead == 8 && P->max_window_tokens > 0 && P->max_window_tokens <= ATT_MAXT && P->num_windows_dev)
    rc = sstb_win_attn_mma(c, qkv, L->nhead, P->num_windows_dev, P->win_offsets, P->tok_perm, att);
  else  // cosine attention / unbounded windows: SIMT kernel on the bf16 operands
    rc = sstb_win_attn<__nv_bfloat16, __nv_bfloat16>(c, qkv, d, L->nhead, n_cap, n_dev, P->win_offsets, P->tok_perm, P->tok_win,
                                                     L->tau, L->tau_n, L->tau_min, att);
  if (rc) return rc;
  // 3. out-projection + residual + LayerNorm1
  memset(&g, 0, sizeof(g));
  g.M_cap = n_cap;
  g.M_dev = n_dev;
  g.A = att;
  g.lda = d;
  g.W = (const __nv_bfloat16*)L->out_proj_w_bf16;
  g.bias = L->out_proj_b;
  g.res = x;
  g.gamma = L->norm1_w;
  g.beta = L->norm1_b;
  g.eps = L->norm_eps;
  g.out_f32 = x1;
  g.out_bf16 = x1b;
  g.ldo = d;
  rc = launch_umma<128, 128, PRO_BF16, EPI_RES_LN>(c, g, 1);
  if (rc) return rc;
  // 4. FFN1 + GELU
  memset(&g, 0, sizeof(g));
  g.M_cap = n_cap;
  g.M_dev = n_dev;
  g.A = x1b;
  g.lda = d;
  g.W = (const __nv_bfloat16*)L->lin1_w_bf16;
  g.bias = L->lin1_b;
  g.out_bf16 = hid;
  g.ldo = ff;
  rc = launch_umma<128, 128, PRO_BF16, EPI_BF16_GELU>(c, g, 2);
  if (rc) return rc;
  // 5. FFN2 + residual + LayerNorm2
  memset(&g, 0, sizeof(g));
  g.M_cap = n_cap;
  g.M_dev = n_dev;
  g.A = hid;
  g.lda = ff;
  g.W = (const __nv_bfloat16*)L->lin2_w_bf16;
  g.bias = L->lin2_b;
  g.res = x1;
  g.gamma = L->norm2_w;
  g.beta = L->norm2_b;
  g.eps = L->norm_eps;
  g.out_f32 = y;
  rc = launch_umma<256, 128, PRO_BF16, EPI_RES_LN>(c, g, 1);
  return rc;
}

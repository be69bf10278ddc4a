// SIRLayer forward (S1-S3): FSD's segmented point-group MLP + max-pool
// (mmdet3d/models/voxel_encoders/voxel_encoder.py:617-764, build_mlp ops/sst/sst_ops.py:334-361).
//
//   rel   = MLP_rel(f_cluster / rel_dist_scaler)                 3 -> 16 -> 32 -> Cin   (Linear_nobias + LN + act)
//   x0    = [xyz / xyz_normalizer || feats] * rel                [N, Cin]
//   p0    = act(LN(W0 x0))        g0 = segmax(p0)                [N, C0], [G, C0]
//   p1    = act(LN(W1a p0 + (W1b g0)[group]))   g1 = segmax(p1)  ([p0 || g0[group]] never materialised)
//   out   = p1 (+ shortcut when shapes match), group feats [g0 || g1]
//
// The reference runs this as ~25 ATen/torch_scatter launches per block with [N,256] concat tensors and float
// atomics.  Here: one warp-per-point kernel for the relation MLP + gating, FFMA row GEMMs with fused LN/act,
// and a sorted segmented max (CSR order, in-tile reduce, order-preserving atomicMax only at tile edges).
#include <stdarg.h>
#include "index.cuh"
#include "sra.cuh"

// ---- relation MLP + gating: one warp per point -----------------------------------------------------
struct RelDev {
  int cin, n_rel, dims[4], rel_in, act;
  float eps, inv_norm[3], rel_dist_scaler;
  const float *w[4], *g[4], *b[4];
};

__device__ __forceinline__ float act_fn(float x, int act) { return act == 2 ? gelu_erf(x) : fmaxf(x, 0.f); }

#define REL_MAXC 8  // channels per lane (Cin <= 256)

__global__ void __launch_bounds__(256) sir_rel_gate_kernel(RelDev r, const float* __restrict__ feats /*[N,cin]*/,
                                                           const float* __restrict__ f_cluster /*[N,3]*/, int N,
                                                           float* __restrict__ x0 /*[N,cin]*/) {
  pdl_wait();
  pdl_launch();
  extern __shared__ float sw[];  // transposed weights: layer l stored [in_l][out_l]
  float* wl[4];
  {
    float* p = sw;
    int in = r.rel_in;
    for (int l = 0; l < r.n_rel; l++) {
      wl[l] = p;
      for (int i = threadIdx.x; i < in * r.dims[l]; i += blockDim.x) {
        int k = i / r.dims[l], c = i % r.dims[l];
        p[i] = r.w[l][(size_t)c * in + k];
      }
      p += in * r.dims[l];
      in = r.dims[l];
    }
  }
  __syncthreads();
  int ln = lane_id();
  int warps = (gridDim.x * blockDim.x) >> 5;
  for (int p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; p < N; p += warps) {
    // input: f_cluster / scaler (3 values, lanes 0..2)
    float cur[REL_MAXC];
#pragma unroll
    for (int j = 0; j < REL_MAXC; j++) cur[j] = 0.f;
    if (ln < r.rel_in) cur[0] = f_cluster[(size_t)p * r.rel_in + ln] / r.rel_dist_scaler;
    int in = r.rel_in;
    for (int l = 0; l < r.n_rel; l++) {
      int out = r.dims[l];
      float acc[REL_MAXC];
#pragma unroll
      for (int j = 0; j < REL_MAXC; j++) acc[j] = 0.f;
#pragma unroll
      for (int jk = 0; jk < REL_MAXC; jk++) {
        if (32 * jk < in) {
          int kmax = min(32, in - 32 * jk);
          for (int kk = 0; kk < kmax; kk++) {
            float xk = __shfl_sync(0xffffffffu, cur[jk], kk);
            const float* wr = wl[l] + (size_t)(32 * jk + kk) * out;
#pragma unroll
            for (int j = 0; j < REL_MAXC; j++) {
              int c = ln + 32 * j;
              if (c < out) acc[j] = fmaf(wr[c], xk, acc[j]);
            }
          }
        }
      }
      // LayerNorm over `out` channels + activation
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < REL_MAXC; j++)
        if (ln + 32 * j < out) s += acc[j];
      float mean = warp_sum(s) / (float)out;
      float q = 0.f;
#pragma unroll
      for (int j = 0; j < REL_MAXC; j++)
        if (ln + 32 * j < out) {
          float t = acc[j] - mean;
          q += t * t;
        }
      float rstd = rsqrtf(warp_sum(q) / (float)out + r.eps);
#pragma unroll
      for (int j = 0; j < REL_MAXC; j++) {
        int c = ln + 32 * j;
        cur[j] = c < out ? act_fn((acc[j] - mean) * rstd * r.g[l][c] + r.b[l][c], r.act) : 0.f;
      }
      in = out;
    }
    // gate: x0 = [xyz / normalizer || rest] * rel
#pragma unroll
    for (int j = 0; j < REL_MAXC; j++) {
      int c = ln + 32 * j;
      if (c < r.cin) {
        float v = feats[(size_t)p * r.cin + c];
        if (c < 3) v = v / (c == 0 ? r.inv_norm[0] : (c == 1 ? r.inv_norm[1] : r.inv_norm[2]));
        x0[(size_t)p * r.cin + c] = v * cur[j];
      }
    }
  }
}

// ---- sorted segmented max: rows visited in CSR order, one thread per channel -----------------------------
__global__ void fill_u32_kernel(uint32_t* p, size_t n, uint32_t v) {
  pdl_wait();
  pdl_launch();
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
#define SEG_ROWS 64
__global__ void __launch_bounds__(128) segmax_sorted_kernel(const float* __restrict__ src, int C, const int32_t* __restrict__ order,
                                                            const long long* __restrict__ seg_of_row, int N,
                                                            uint32_t* __restrict__ out_ord /*[G, C] order-preserving uint*/) {
  pdl_wait();
  pdl_launch();
  __shared__ int s_row[SEG_ROWS];
  __shared__ int s_seg[SEG_ROWS];
  int k0 = blockIdx.x * SEG_ROWS;
  int nrow = min(SEG_ROWS, N - k0);
  for (int i = threadIdx.x; i < nrow; i += blockDim.x) {
    int row = order[k0 + i];
    s_row[i] = row;
    s_seg[i] = (int)seg_of_row[row];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float m = -INFINITY;
    int seg = s_seg[0];
    for (int i = 0; i < nrow; i++) {
      int sg = s_seg[i];
      if (sg != seg) {
        atomicMax(&out_ord[(size_t)seg * C + c], f2ord(m));
        seg = sg;
        m = -INFINITY;
      }
      m = fmaxf(m, src[(size_t)s_row[i] * C + c]);
    }
    atomicMax(&out_ord[(size_t)seg * C + c], f2ord(m));
  }
}
__global__ void segmax_finalize_kernel(const uint32_t* __restrict__ ord, int G, int C, float* __restrict__ out, int ldo, int col0) {
  pdl_wait();
  pdl_launch();
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)G * C) return;
  int g = (int)(i / C), c = (int)(i % C);
  uint32_t u = ord[i];
  out[(size_t)g * ldo + col0 + c] = (u == 0u) ? 0.f : ord2f(u);  // empty segment -> 0 like torch_scatter
}

void sstb_sir_segmax_finalize(cudaStream_t st, const uint32_t* ord, int G, int C, float* out, int ldo, int col0) {
  size_t gn = (size_t)G * C;
  launch_pdl(segmax_finalize_kernel, dim3((unsigned)((gn + 255) / 256)), dim3(256), (size_t)(0), st, ord, G, C, out, ldo, col0);
}

// with_shortcut (voxel_encoder.py:753-759): point_feats += features[:, 3:] when the shapes agree
__global__ void shortcut_kernel(float* __restrict__ out, const float* __restrict__ in_feats, int N, int C, int cin) {
  pdl_wait();
  pdl_launch();
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)N * C) return;
  int p = (int)(i / C), ch = (int)(i % C);
  out[i] += in_feats[(size_t)p * cin + 3 + ch];
}

// Group CSR (offsets [G+1] int32, order [N] int32): points grouped by `inv`; shared by every block of a SIR backbone.
static int group_csr(sstb200_ctx* c, const long long* inv, int N, int G, int32_t* offsets_out, int32_t* order_out) {
  cudaStream_t st = c->stream;
  int32_t* count = arena_alloc<int32_t>(c, (size_t)G + 2);
  if (!count) return sstb_fail(c, SSTB_ERR_WORKSPACE, "sir: arena");
  CUDA_TRY(c, cudaMemsetAsync(count, 0, ((size_t)G + 2) * 4, st));
  const bool small = G <= CSR_SMALL_MAX;
  if (small)
    launch_pdl(count_small_kernel<long long>, dim3((N + CSR_SMALL_ITEMS - 1) / CSR_SMALL_ITEMS), dim3(1024), (size_t)G * 4, st, inv, N, G, count, count + G + 1);
  else
    launch_pdl(count_index_kernel, dim3((N + 255) / 256), dim3(256), (size_t)(0), st, inv, N, G, count, count + G + 1);
  Csr r;
  r.offsets = (uint32_t*)offsets_out;
  r.order = order_out;
  int rc = csr_build<long long>(c, r, inv, N, count, G, nullptr, nullptr, small ? G : 0);
  if (rc) return rc;
  LAUNCH_CHECK(c);
  return SSTB_OK;
}

extern "C" int sstb200_group_csr(sstb200_ctx* c, const int64_t* inv, int N, int G, int32_t* offsets, int32_t* order) {
  CHECK_ARG(c, c && N >= 0 && G >= 0);
  if (N == 0 || G == 0) return SSTB_OK;
  CHECK_ARG(c, inv && offsets && order);
  arena_reset(c);
  int rc = arena_reserve(c, csr_bytes(N, G) + al256((size_t)G * 4 + 8) + 65536);
  if (rc) return rc;
  return group_csr(c, (const long long*)inv, N, G, offsets, order);
}

int sstb_sir_layer_bf16(sstb200_ctx* c, const sstb200_sir_layer* L, const float* in_feats, const float* f_cluster, const long long* inv,
                        const int32_t* order, int N, int G, __nv_bfloat16* p0buf, uint32_t* gord, float* gterm, float* out_point, int ldo,
                        float* out_group, int in_ld, int gap_at, int gap);

extern "C" int sstb200_sir_layer_forward_ex(sstb200_ctx* c, const sstb200_sir_layer* L, const float* in_feats, int in_ld, int in_gap_at,
                                            int in_gap, const float* f_cluster, const int64_t* inv, int N, int G,
                                            const int32_t* csr_offsets, const int32_t* csr_order, int precision, float* out_point,
                                            int out_point_ld, float* out_group) {
  CHECK_ARG(c, c && L && N >= 0 && G >= 0);
  if (N == 0 || G == 0) return SSTB_OK;
  CHECK_ARG(c, in_feats && f_cluster && inv && out_point && out_group);
  CHECK_ARG(c, (csr_offsets == nullptr) == (csr_order == nullptr));
  CHECK_ARG(c, precision == SSTB200_PREC_FP32 || precision == SSTB200_PREC_BF16);
  CHECK_ARG(c, L->num_vfe >= 1 && L->num_vfe <= 2 && L->num_rel >= 0 && L->num_rel <= 4 && (L->act == 1 || L->act == 2));
  if (!L->mode_max) return sstb_fail(c, SSTB_ERR_UNSUPPORTED, "SIRLayer: only mode='max' is built");
  const int cin = L->in_channels, C0 = L->feat_channels[0], C1 = L->num_vfe > 1 ? L->feat_channels[1] : 0;
  const int Clast = C1 ? C1 : C0;
  if (out_point_ld <= 0) out_point_ld = Clast;
  CHECK_ARG(c, out_point_ld >= Clast);
  if (cin > 32 * REL_MAXC || C0 > 256 || C1 > 256 || C0 % 32 || C1 % 32)
    return sstb_fail(c, SSTB_ERR_UNSUPPORTED, "SIR dims (cin=%d, C0=%d, C1=%d) not supported", cin, C0, C1);
  if (L->num_rel > 0 && L->rel_dims[L->num_rel - 1] != cin) return sstb_fail(c, SSTB_ERR_ARG, "last rel-MLP layer must produce in_channels");
  for (int l = 0; l < L->num_rel; l++)
    if (L->rel_dims[l] > 32 * REL_MAXC) return sstb_fail(c, SSTB_ERR_UNSUPPORTED, "rel-MLP width %d", L->rel_dims[l]);
  const int Cmax = C0 > C1 ? C0 : C1;
  const bool bf16 = precision == SSTB200_PREC_BF16;
  if (in_ld <= 0) in_ld = cin, in_gap_at = cin, in_gap = 0;
  CHECK_ARG(c, in_gap >= 0 && in_gap_at >= 0 && in_gap_at <= cin && in_ld >= cin + in_gap);
  if (!bf16 && (out_point_ld != Clast || in_ld != cin || in_gap != 0))
    return sstb_fail(c, SSTB_ERR_UNSUPPORTED, "SIRLayer fp32 path works on dense point features (ld == C, no gap)");
  arena_reset(c);
  int rc = arena_reserve(c, csr_bytes(N, G) + al256((size_t)G * 4 + 8) + al256((size_t)N * cin * 4) + 2 * al256((size_t)N * Cmax * 4) +
                                2 * al256((size_t)G * Cmax * 4) + al256(((size_t)N + 4) * 4) + 65536);
  if (rc) return rc;
  cudaStream_t st = c->stream;
  const int32_t* order = csr_order;
  if (!order) {
    int32_t* off = arena_alloc<int32_t>(c, (size_t)G + 2);
    int32_t* ord = arena_alloc<int32_t>(c, (size_t)N + 1);
    if (!off || !ord) return sstb_fail(c, SSTB_ERR_WORKSPACE, "sir: arena");
    rc = group_csr(c, (const long long*)inv, N, G, off, ord);
    if (rc) return rc;
    order = ord;
  }
  uint32_t* gord = arena_alloc<uint32_t>(c, (size_t)G * Cmax);
  float* gterm = arena_alloc<float>(c, (size_t)G * Cmax);
  if (!gord || !gterm) return sstb_fail(c, SSTB_ERR_WORKSPACE, "sir: arena");
  if (bf16) {
    __nv_bfloat16* p0buf = arena_alloc<__nv_bfloat16>(c, ((size_t)N + 128) * 128);
    if (!p0buf) return sstb_fail(c, SSTB_ERR_WORKSPACE, "sir: arena");
    rc = sstb_sir_layer_bf16(c, L, in_feats, f_cluster, (const long long*)inv, order, N, G, p0buf, gord, gterm, out_point, out_point_ld, out_group, in_ld, in_gap_at,
                             in_gap);
    if (rc) return rc;
    if (L->with_shortcut && cin - 3 == Clast) {
      if (out_point_ld != Clast || in_ld != cin) return sstb_fail(c, SSTB_ERR_UNSUPPORTED, "shortcut with strided point features");
      launch_pdl(shortcut_kernel, dim3((unsigned)(((size_t)N * Clast + 255) / 256)), dim3(256), (size_t)(0), st, out_point, in_feats, N, Clast, cin);
    }
    LAUNCH_CHECK(c);
    return SSTB_OK;
  }
  float* x0 = arena_alloc<float>(c, (size_t)N * cin);
  float* t = arena_alloc<float>(c, (size_t)N * Cmax);
  float* p0 = arena_alloc<float>(c, (size_t)N * Cmax);
  if (!x0 || !t || !p0) return sstb_fail(c, SSTB_ERR_WORKSPACE, "sir: arena");

  // 1. relation MLP + gating
  const float* xin = in_feats;
  if (L->num_rel > 0) {
    RelDev rd;
    rd.cin = cin;
    rd.n_rel = L->num_rel;
    rd.rel_in = L->rel_in;
    rd.act = L->act;
    rd.eps = L->norm_eps;
    rd.rel_dist_scaler = L->rel_dist_scaler;
    size_t smem = 0;
    int in = L->rel_in;
    for (int l = 0; l < 4; l++) {
      rd.dims[l] = l < L->num_rel ? L->rel_dims[l] : 0;
      rd.w[l] = L->rel_w[l];
      rd.g[l] = L->rel_ln_w[l];
      rd.b[l] = L->rel_ln_b[l];
      if (l < L->num_rel) {
        CHECK_ARG(c, rd.w[l] && rd.g[l] && rd.b[l]);
        smem += (size_t)in * rd.dims[l] * 4;
        in = rd.dims[l];
      }
    }
    for (int i = 0; i < 3; i++) rd.inv_norm[i] = L->xyz_normalizer[i];
    if (smem > 200 * 1024) return sstb_fail(c, SSTB_ERR_UNSUPPORTED, "rel-MLP weights do not fit shared memory");
    static SmemAttr sa;
    CUDA_TRY(c, ensure_smem(c, sa, sir_rel_gate_kernel, (size_t)200 * 1024));
    launch_pdl(sir_rel_gate_kernel, dim3(c->num_sms * 4), dim3(256), (size_t)(smem), st, rd, in_feats, f_cluster, N, x0);
    xin = x0;
  } else {
    return sstb_fail(c, SSTB_ERR_UNSUPPORTED, "SIRLayer without rel-MLP is not built");
  }
  // 2. layer 0: p0 = act(LN(W0 x0)); g0 = segmax(p0)
  CHECK_ARG(c, L->vfe_w[0] && L->vfe_ln_w[0] && L->vfe_ln_b[0]);
  sstb_gemm_rows_ex(st, xin, cin, L->vfe_w[0], cin, nullptr, nullptr, 0, nullptr, t, C0, N, nullptr, C0, cin, 0, nullptr, nullptr, 0, 0, 0, 0);
  float* p0out = (C1 == 0) ? out_point : p0;
  sstb_add_norm_act(st, t, nullptr, L->vfe_ln_w[0], L->vfe_ln_b[0], nullptr, nullptr, L->norm_eps, p0out, N, nullptr, C0, L->act);
  const int Cg = C0 + C1;
  size_t gn = (size_t)G * C0;
  launch_pdl(fill_u32_kernel, dim3((unsigned)((gn + 255) / 256)), dim3(256), (size_t)(0), st, gord, gn, 0u);
  launch_pdl(segmax_sorted_kernel, dim3((N + SEG_ROWS - 1) / SEG_ROWS), dim3(128), (size_t)(0), st, p0out, C0, order, (const long long*)inv, N, gord);
  launch_pdl(segmax_finalize_kernel, dim3((unsigned)((gn + 255) / 256)), dim3(256), (size_t)(0), st, gord, G, C0, out_group, Cg, 0);
  if (C1) {
    CHECK_ARG(c, L->vfe_w[1] && L->vfe_ln_w[1] && L->vfe_ln_b[1]);
    // 3. layer 1 on [p0 || g0[group]]:  W1a p0 + (W1b g0)[group]
    sstb_gemm_rows_ex(st, out_group, Cg, L->vfe_w[1] + C0, 2 * C0, nullptr, nullptr, 0, nullptr, gterm, C1, G, nullptr, C1, C0, 0, nullptr,
                      nullptr, 0, 0, 0, 0);
    sstb_gemm_rows_ex(st, p0, C0, L->vfe_w[1], 2 * C0, nullptr, gterm, C1, (const long long*)inv, t, C1, N, nullptr, C1, C0, 0, nullptr,
                      nullptr, 0, 0, 0, 0);
    sstb_add_norm_act(st, t, nullptr, L->vfe_ln_w[1], L->vfe_ln_b[1], nullptr, nullptr, L->norm_eps, out_point, N, nullptr, C1, L->act);
    gn = (size_t)G * C1;
    launch_pdl(fill_u32_kernel, dim3((unsigned)((gn + 255) / 256)), dim3(256), (size_t)(0), st, gord, gn, 0u);
    launch_pdl(segmax_sorted_kernel, dim3((N + SEG_ROWS - 1) / SEG_ROWS), dim3(128), (size_t)(0), st, out_point, C1, order, (const long long*)inv, N, gord);
    launch_pdl(segmax_finalize_kernel, dim3((unsigned)((gn + 255) / 256)), dim3(256), (size_t)(0), st, gord, G, C1, out_group, Cg, C0);
  }
  {
    int Cl = C1 ? C1 : C0;
    if (L->with_shortcut && cin - 3 == Cl)
      launch_pdl(shortcut_kernel, dim3((unsigned)(((size_t)N * Cl + 255) / 256)), dim3(256), (size_t)(0), st, out_point, in_feats, N, Cl, cin);
  }
  LAUNCH_CHECK(c);
  return SSTB_OK;
}

extern "C" int sstb200_sir_layer_forward(sstb200_ctx* c, const sstb200_sir_layer* L, const float* in_feats, const float* f_cluster,
                                         const int64_t* inv, int N, int G, float* out_point, float* out_group) {
  return sstb200_sir_layer_forward_ex(c, L, in_feats, 0, 0, 0, f_cluster, inv, N, G, nullptr, nullptr, SSTB200_PREC_FP32, out_point, 0, out_group);
}

// Ragged window attention (SIMT, fp32 math), templated on the q/k/v and output element types.
// One thread per (token slot, head): online softmax over the keys of the token's window.  qkv [n, 3d] in flat
// token order; the window CSR gives the key set, so no padding, no mask, no per-level batches.
//   cosine mode (models/sst/cosine_msa.py:123-185): q,k L2-normalised, logits / clamp(tau, tau_min).
#pragma once
#include "common.cuh"

template <typename T, int N>
__device__ __forceinline__ void load_vec(const T* p, float* o);
template <>
__device__ __forceinline__ void load_vec<float, 8>(const float* p, float* o) {
  float4 a = *(const float4*)p, b = *(const float4*)(p + 4);
  o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
}
template <>
__device__ __forceinline__ void load_vec<__nv_bfloat16, 8>(const __nv_bfloat16* p, float* o) {
  int4 v = *(const int4*)p;
  const __nv_bfloat162* h = (const __nv_bfloat162*)&v;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    float2 f = __bfloat1622float2(h[i]);
    o[2 * i] = f.x;
    o[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ void store_vec8(float* p, const float* v) {
  *(float4*)p = make_float4(v[0], v[1], v[2], v[3]);
  *(float4*)(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void store_vec8(__nv_bfloat16* p, const float* v) {
  __nv_bfloat162 h[4];
#pragma unroll
  for (int i = 0; i < 4; i++) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  *(int4*)p = *(const int4*)h;
}

template <typename TI, typename TO, int DH>
__global__ void __launch_bounds__(256) win_attn_kernel(const TI* __restrict__ qkv, int d, int nhead, int n,
                                                       const int32_t* __restrict__ n_dev,
                                                       const int32_t* __restrict__ win_offsets,
                                                       const int32_t* __restrict__ tok_perm,
                                                       const int32_t* __restrict__ tok_win, float scale,
                                                       const float* __restrict__ tau, int tau_n, float tau_min,
                                                       TO* __restrict__ out) {
  if (n_dev) n = *n_dev;
  long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int slot = (int)(g / nhead), h = (int)(g % nhead);
  if (slot >= n) return;
  int tok = tok_perm[slot];
  int w = tok_win[tok];
  int kb = win_offsets[w], ke = win_offsets[w + 1];
  const TI* qp = qkv + (size_t)tok * 3 * d + h * DH;
  float q[DH];
#pragma unroll
  for (int i = 0; i < DH; i += 8) load_vec<TI, 8>(qp + i, q + i);
  bool cosine = tau != nullptr;
  float s_mul = scale;
  if (cosine) {
    float nq = 0.f;
#pragma unroll
    for (int i = 0; i < DH; i++) nq = fmaf(q[i], q[i], nq);
    float t = fmaxf(tau_n > 1 ? tau[h] : tau[0], tau_min);
    s_mul = 1.0f / (fmaxf(sqrtf(nq), 1e-12f) * t);  // F.normalize eps = 1e-12
  }
  float m = -INFINITY, l = 0.f, acc[DH];
#pragma unroll
  for (int i = 0; i < DH; i++) acc[i] = 0.f;
  for (int j = kb; j < ke; j++) {
    int kt = tok_perm[j];
    const TI* kp = qkv + (size_t)kt * 3 * d + d + h * DH;
    const TI* vp = kp + d;
    float kk[DH], vv[DH];
#pragma unroll
    for (int i = 0; i < DH; i += 8) {
      load_vec<TI, 8>(kp + i, kk + i);
      load_vec<TI, 8>(vp + i, vv + i);
    }
    float s = 0.f, nk = 0.f;
#pragma unroll
    for (int i = 0; i < DH; i++) {
      s = fmaf(q[i], kk[i], s);
      if (cosine) nk = fmaf(kk[i], kk[i], nk);
    }
    s *= s_mul;
    if (cosine) s /= fmaxf(sqrtf(nk), 1e-12f);
    float mn = fmaxf(m, s);
    float corr = __expf(m - mn);  // m = -inf on the first key -> 0
    float p = __expf(s - mn);
    l = l * corr + p;
#pragma unroll
    for (int i = 0; i < DH; i++) acc[i] = fmaf(acc[i], corr, p * vv[i]);
    m = mn;
  }
  float inv = 1.0f / l;
#pragma unroll
  for (int i = 0; i < DH; i++) acc[i] *= inv;
  TO* op = out + (size_t)tok * d + h * DH;
#pragma unroll
  for (int i = 0; i < DH; i += 8) store_vec8(op + i, acc + i);
}

template <typename TI, typename TO>
int sstb_win_attn(sstb200_ctx* c, const TI* qkv, int d, int nhead, int n_cap, const int32_t* n_dev, const int32_t* win_offsets,
                  const int32_t* tok_perm, const int32_t* tok_win, const float* tau, int tau_n, float tau_min, TO* out) {
  int dh = d / nhead;
  long long items = (long long)n_cap * nhead;
  unsigned grid = (unsigned)((items + 255) / 256);
  float scale = 1.0f / sqrtf((float)dh);
  if (grid == 0) return SSTB_OK;
#define LAUNCH_ATT(DH)                                                                                                  \
  win_attn_kernel<TI, TO, DH><<<grid, 256, 0, c->stream>>>(qkv, d, nhead, n_cap, n_dev, win_offsets, tok_perm, tok_win, \
                                                           scale, tau, tau_n, tau_min, out)
  if (dh == 16) LAUNCH_ATT(16);
  else if (dh == 8) LAUNCH_ATT(8);
  else if (dh == 32) LAUNCH_ATT(32);
  else if (dh == 64) LAUNCH_ATT(64);
  else return sstb_fail(c, SSTB_ERR_UNSUPPORTED, "head dim %d not supported (8/16/32/64)", dh);
#undef LAUNCH_ATT
  return SSTB_OK;
}

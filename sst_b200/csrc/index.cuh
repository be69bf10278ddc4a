// Bitmap-rank index ("unique rows, sorted" without a sort), CSR grouping and segmented reduce.
// Shared by voxel.cu / window.cu / vfe.cu / sir.cu.  See DESIGN.md "voxel index".
#pragma once
#include "common.cuh"

// ------------------------------------------------------------------------------------------------
struct Extents {
  int ndim;
  long long lo[4];
  long long ext[4];
};

struct KeyIndex {
  long long* keys;        // [P]  linear key or -1
  uint32_t* bitmap;       // [nwords]
  uint32_t* word_prefix;  // [nwords+1]
  uint32_t* total;        // distinct keys (device)
  int32_t* flags;         // [0] has_invalid
  ScanTemps st;
  size_t nwords;
  long long T;
};

static size_t key_index_bytes(size_t P, long long T) {
  size_t nwords = (size_t)((T + 31) / 32);
  size_t nblk = scan_num_blocks(nwords) + 2;
  return al256(P * 8) + al256(nwords * 4) + al256((nwords + 1) * 4) + 4 * al256(nblk * 4) + 1024;
}

static int key_index_alloc(sstb200_ctx* c, KeyIndex& k, size_t P, long long T) {
  k.T = T;
  k.nwords = (size_t)((T + 31) / 32);
  size_t nblk = scan_num_blocks(k.nwords) + 2;
  k.keys = arena_alloc<long long>(c, P ? P : 1);
  k.bitmap = arena_alloc<uint32_t>(c, k.nwords);
  k.word_prefix = arena_alloc<uint32_t>(c, k.nwords + 1);
  k.st.block_sums = arena_alloc<uint32_t>(c, nblk);
  k.st.block_prefix = arena_alloc<uint32_t>(c, nblk);
  k.st.ticket = arena_alloc<uint32_t>(c, 64);
  k.total = k.st.ticket + 1;
  k.flags = (int32_t*)(k.st.ticket + 2);
  if (!k.keys || !k.bitmap || !k.word_prefix || !k.st.block_sums || !k.st.block_prefix || !k.st.ticket)
    return sstb_fail(c, SSTB_ERR_WORKSPACE, "key index: arena too small");
  CUDA_TRY(c, cudaMemsetAsync(k.bitmap, 0, k.nwords * 4, c->stream));
  CUDA_TRY(c, cudaMemsetAsync(k.st.ticket, 0, 64 * 4, c->stream));
  return SSTB_OK;
}

static void key_index_scan(sstb200_ctx* c, KeyIndex& k) {
  launch_exclusive_scan(c->stream, LoadPopc{k.bitmap}, k.nwords, nullptr, k.st, k.word_prefix, k.total, true);
}

// ---- mark kernels --------------------------------------------------------------------------------
template <typename TI>
__global__ void mark_rows_kernel(const TI* __restrict__ rows, int n, Extents e, bool negative_is_invalid,
                                 long long* __restrict__ keys, uint32_t* __restrict__ bitmap,
                                 int32_t* __restrict__ flags, const int32_t* __restrict__ n_dev = nullptr) {
  pdl_wait();
  pdl_launch();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (n_dev) n = *n_dev;
  if (i >= n) return;
  long long key = 0;
  bool bad = false;
#pragma unroll 4
  for (int d = 0; d < e.ndim; d++) {
    long long v = (long long)rows[(size_t)i * e.ndim + d];
    if (negative_is_invalid && v < 0) bad = true;
    long long r = v - e.lo[d];
    if (r < 0 || r >= e.ext[d]) bad = true;
    key = key * e.ext[d] + r;
  }
  if (bad) {
    keys[i] = -1;
    flags[0] = 1;
    return;
  }
  keys[i] = key;
  atomicOr(&bitmap[key >> 5], 1u << (key & 31));
}

// ---- emit unique rows (decode keys of set bits) ----------------------------------------------------
template <typename TO>
__global__ void emit_rows_kernel(const uint32_t* __restrict__ bitmap, const uint32_t* __restrict__ word_prefix,
                                 size_t nwords, Extents e, int shift_if_no_invalid, const int32_t* __restrict__ flags,
                                 TO* __restrict__ out_rows, const uint32_t* __restrict__ total, int32_t* __restrict__ num_out) {
  pdl_wait();
  pdl_launch();
  int shift = (shift_if_no_invalid && flags[0] == 0) ? 1 : 0;
  size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w == 0) {
    int t = (int)(*total) - shift;
    *num_out = t < 0 ? 0 : t;
  }
  for (; w < nwords; w += (size_t)gridDim.x * blockDim.x) {
    uint32_t bits = bitmap[w];
    if (!bits) continue;
    long long v = (long long)word_prefix[w] - shift;
    while (bits) {
      int b = __ffs(bits) - 1;
      bits &= bits - 1;
      if (v >= 0) {
        long long key = (long long)w * 32 + b;
        TO c[4];
#pragma unroll 4
        for (int d = e.ndim - 1; d >= 0; d--) {
          long long q = key / e.ext[d];
          c[d] = (TO)(key - q * e.ext[d] + e.lo[d]);
          key = q;
        }
        for (int d = 0; d < e.ndim; d++) out_rows[(size_t)v * e.ndim + d] = c[d];
      }
      v++;
    }
  }
}

// ---- map + count -------------------------------------------------------------------------------------
template <typename TM>
__global__ void map_count_kernel(const long long* __restrict__ keys, int n, const uint32_t* __restrict__ bitmap,
                                 const uint32_t* __restrict__ word_prefix, int shift_if_no_invalid,
                                 const int32_t* __restrict__ flags, TM* __restrict__ map, int32_t* __restrict__ count,
                                 const int32_t* __restrict__ n_dev = nullptr) {
  pdl_wait();
  pdl_launch();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (n_dev) n = *n_dev;
  if (i >= n) return;
  int shift = (shift_if_no_invalid && flags[0] == 0) ? 1 : 0;
  long long key = keys[i];
  long long v = -1;
  if (key >= 0) {
    size_t w = (size_t)(key >> 5);
    v = (long long)word_prefix[w] + __popc(bitmap[w] & ((1u << (key & 31)) - 1u)) - shift;
  }
  map[i] = (TM)v;
  if (v >= 0 && count) atomicAdd(&count[v], 1);
}

// ---- CSR fill (counting sort, order inside a segment unspecified) -----------------------------------------
template <typename TM>
__global__ void csr_fill_kernel(const TM* __restrict__ map, int n, const uint32_t* __restrict__ offsets,
                                int32_t* __restrict__ cursor, int32_t* __restrict__ order,
                                const int32_t* __restrict__ n_dev = nullptr) {
  pdl_wait();
  pdl_launch();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (n_dev) n = *n_dev;
  if (i >= n) return;
  long long v = (long long)map[i];
  if (v < 0) return;
  int pos = atomicAdd(&cursor[v], 1);
  order[offsets[v] + pos] = i;
}

// ---- segmented reduce: GROUP lanes cooperate on one segment, lanes stride the channels --------------------
template <int GROUP>
__global__ void __launch_bounds__(256) segment_reduce_kernel(const float* __restrict__ src, int C,
                                                             const uint32_t* __restrict__ offsets,
                                                             const int32_t* __restrict__ order, int nseg_host,
                                                             const int32_t* __restrict__ nseg_dev, int mode,
                                                             float empty_value, float* __restrict__ out,
                                                             long long* __restrict__ argmax, int n_rows) {
  pdl_wait();
  pdl_launch();
  int nseg = nseg_dev ? *nseg_dev : nseg_host;
  int groups_per_block = blockDim.x / GROUP;
  int g = threadIdx.x / GROUP, l = threadIdx.x % GROUP;
  for (int s = blockIdx.x * groups_per_block + g; s < nseg; s += gridDim.x * groups_per_block) {
    uint32_t b = offsets[s], e = offsets[s + 1];
    for (int c = l; c < C; c += GROUP) {
      if (mode == SSTB200_REDUCE_MAX) {
        float m = -INFINITY;
        long long am = n_rows;
        for (uint32_t k = b; k < e; k++) {
          int p = order[k];
          float v = src[(size_t)p * C + c];
          if (v > m || (v == m && p < am)) {
            m = v;
            am = p;
          }
        }
        if (b == e) m = empty_value;
        out[(size_t)s * C + c] = m;
        if (argmax) argmax[(size_t)s * C + c] = am;
      } else {
        // fp64 accumulation: exact for these magnitudes, so the result does not depend on CSR order
        double acc = 0.0;
        for (uint32_t k = b; k < e; k++) acc += (double)src[(size_t)order[k] * C + c];
        float r = (float)acc;
        if (mode == SSTB200_REDUCE_MEAN && e > b) r = (float)acc / (float)(e - b);
        out[(size_t)s * C + c] = r;
      }
    }
  }
}

static void launch_segment_reduce(sstb200_ctx* c, const float* src, int C, const uint32_t* offsets,
                                  const int32_t* order, int nseg_cap, const int32_t* nseg_dev, int mode,
                                  float empty_value, float* out, long long* argmax, int n_rows) {
  int grid = c->num_sms * 8;
  if (C >= 32)
    launch_pdl(segment_reduce_kernel<32>, dim3(grid), dim3(256), (size_t)(0), c->stream, src, C, offsets, order, nseg_cap, nseg_dev, mode, empty_value, out, argmax, n_rows);
  else if (C > 4)
    launch_pdl(segment_reduce_kernel<8>, dim3(grid), dim3(256), (size_t)(0), c->stream, src, C, offsets, order, nseg_cap, nseg_dev, mode, empty_value, out, argmax, n_rows);
  else
    launch_pdl(segment_reduce_kernel<4>, dim3(grid), dim3(256), (size_t)(0), c->stream, src, C, offsets, order, nseg_cap, nseg_dev, mode, empty_value, out, argmax, n_rows);
}

// CSR over a map -> (offsets[nseg+1], order[n]) ; count must already hold per-segment counts.
struct Csr {
  uint32_t* offsets;
  int32_t* order;
  int32_t* cursor;
  ScanTemps st;
  uint32_t* total;
};
static size_t csr_bytes(size_t n, size_t nseg_cap) {
  size_t nblk = scan_num_blocks(nseg_cap) + 2;
  return al256((nseg_cap + 2) * 4) * 2 + al256(n * 4 + 4) + 3 * al256(nblk * 4) + 2048;
}
template <typename TM>
static int csr_build(sstb200_ctx* c, Csr& r, const TM* map, int n, const int32_t* count, size_t nseg_cap,
                     const int32_t* nseg_dev, const int32_t* n_dev = nullptr) {
  size_t nblk = scan_num_blocks(nseg_cap) + 2;
  r.offsets = arena_alloc<uint32_t>(c, nseg_cap + 2);
  r.cursor = arena_alloc<int32_t>(c, nseg_cap + 2);
  r.order = arena_alloc<int32_t>(c, n + 1);
  r.st.block_sums = arena_alloc<uint32_t>(c, nblk);
  r.st.block_prefix = arena_alloc<uint32_t>(c, nblk);
  r.st.ticket = arena_alloc<uint32_t>(c, 64);
  if (!r.offsets || !r.cursor || !r.order || !r.st.block_sums || !r.st.block_prefix || !r.st.ticket)
    return sstb_fail(c, SSTB_ERR_WORKSPACE, "csr: arena too small");
  r.total = r.st.ticket + 1;
  CUDA_TRY(c, cudaMemsetAsync(r.st.ticket, 0, 64 * 4, c->stream));
  CUDA_TRY(c, cudaMemsetAsync(r.cursor, 0, (nseg_cap + 2) * 4, c->stream));
  launch_exclusive_scan(c->stream, LoadU32{(const uint32_t*)count}, nseg_cap, nseg_dev, r.st, r.offsets, r.total, true);
  if (n > 0) launch_pdl(csr_fill_kernel<TM>, dim3((n + 255) / 256), dim3(256), (size_t)(0), c->stream, map, n, r.offsets, r.cursor, r.order, n_dev);
  LAUNCH_CHECK(c);
  return SSTB_OK;
}


static int read_back_i32(sstb200_ctx* c, const int32_t* dev, int32_t* host) {
  CUDA_TRY(c, cudaMemcpyAsync(c->pinned_i32, dev, 4, cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  *host = c->pinned_i32[0];
  return SSTB_OK;
}

static int make_extents(sstb200_ctx* c, Extents& e, int ndim, const long long* lo, const long long* hi, long long* T) {
  e.ndim = ndim;
  long long t = 1;
  for (int d = 0; d < 4; d++) {
    e.lo[d] = 0;
    e.ext[d] = 1;
  }
  for (int d = 0; d < ndim; d++) {
    if (hi[d] < lo[d]) return sstb_fail(c, SSTB_ERR_ARG, "extent %d: hi < lo", d);
    e.lo[d] = lo[d];
    e.ext[d] = hi[d] - lo[d] + 1;
    if (t > ((long long)1 << 40) / e.ext[d]) return sstb_fail(c, SSTB_ERR_UNSUPPORTED, "bounding grid too large");
    t *= e.ext[d];
  }
  if (t > ((long long)1 << 34))
    return sstb_fail(c, SSTB_ERR_UNSUPPORTED,
                     "bounding grid of %lld cells exceeds the bitmap-rank limit (2^34); sort fallback not built yet", t);
  *T = t;
  return SSTB_OK;
}


// One warp per segment: rank(i) = #{j in segment : idx_j < idx_i}.  Segments on this path are windows
// (<= a few hundred tokens); cost O(n^2/32) per warp.
static __global__ void __launch_bounds__(256) stable_rank_kernel(const uint32_t* __restrict__ offsets, const int32_t* __restrict__ order,
                                                          const int32_t* __restrict__ nseg_dev, int32_t* __restrict__ sorted_order,
                                                          long long* __restrict__ rank_out_i64, int32_t* __restrict__ rank_out_i32,
                                                          int32_t* __restrict__ slot_out) {
  pdl_wait();
  pdl_launch();
  int nseg = *nseg_dev;
  int warps = (gridDim.x * blockDim.x) >> 5;
  int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  for (int s = w; s < nseg; s += warps) {
    uint32_t b = offsets[s], e = offsets[s + 1];
    uint32_t n = e - b;
    for (uint32_t i = lane_id(); i < n; i += 32) {
      int me = order[b + i];
      int r = 0;
      for (uint32_t j = 0; j < n; j++) r += (order[b + j] < me);
      if (sorted_order) sorted_order[b + r] = me;
      if (rank_out_i64) rank_out_i64[me] = r;
      if (rank_out_i32) rank_out_i32[me] = r;
      if (slot_out) slot_out[me] = (int32_t)(b + r);
    }
  }
}


static __global__ void count_index_kernel(const long long* __restrict__ idx, int n, int nseg, int32_t* __restrict__ count,
                                   int32_t* __restrict__ err) {
  pdl_wait();
  pdl_launch();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  long long v = idx[i];
  if (v < 0 || v >= nseg) {
    *err = 1;
    return;
  }
  atomicAdd(&count[v], 1);
}


// K8: sample reductions -- per-time-step percentiles (np.percentile, linear interpolation),
// cumulative sums, point / cumulative effects, the 10x2 summary table and the tail-area p-value.
//
// Replaces: compile_posterior_inferences      /root/reference/causalimpact/inferences.py:52-249
//           summarize_posterior_inferences    inferences.py:285-369
//           compute_p_value                   inferences.py:372-408
//
// Per-column percentiles: up to 2048 draws a WARP per column runs a bit-plane radix SELECT for the four order
// statistics np.percentile needs (k_col_quantiles_sel); beyond that one CTA sorts the column with
// cub::BlockRadixSort.  Either way a CTA stages 8 adjacent columns of the [nsamp][W] sample matrix through
// shared memory so every global read is a full 32-byte sector.  The summary kernel sorts the nsamp row sums
// of one series per CTA.
// Masked post-period steps (NaN in y_post; causalimpact/data.py:168-174) are dropped from every
// post-period statistic exactly as `post_data[mask]` / `[:, mask]` do in the reference.
#include <cub/block/block_radix_sort.cuh>
#include "ci_abi_common.cuh"

namespace ci {

constexpr int RB = 256;  // threads per sorting CTA

struct QPos {
  int lo0, lo1;   // order statistics bracketing the lower percentile
  int hi0, hi1;   // ... and the upper percentile
  double flo, fhi;  // interpolation fractions
};

// np.percentile(a, [100 alpha/2, 100 - 100 alpha/2]) virtual indices (inferences.py:34-49).
static QPos quantile_positions(int n, double alpha) {
  QPos q;
  const double plo = alpha * 100.0 / 2.0, phi = 100.0 - alpha * 100.0 / 2.0;
  const double vlo = (n - 1) * (plo / 100.0), vhi = (n - 1) * (phi / 100.0);
  q.lo0 = (int)floor(vlo);
  q.hi0 = (int)floor(vhi);
  if (q.lo0 > n - 1) q.lo0 = n - 1;
  if (q.hi0 > n - 1) q.hi0 = n - 1;
  q.lo1 = q.lo0 + 1 < n ? q.lo0 + 1 : n - 1;
  q.hi1 = q.hi0 + 1 < n ? q.hi0 + 1 : n - 1;
  q.flo = vlo - q.lo0;
  q.fhi = vhi - q.hi0;
  return q;
}

__device__ __forceinline__ float lerp_np(float a, float b, double t) {
  // numpy's _lerp: a + (b - a) t, evaluated from the far end when t >= 0.5
  const double da = a, db = b, d = db - da;
  return (float)(t >= 0.5 ? db - d * (1.0 - t) : da + d * t);
}

// Percentiles of every column of mat[n][i][w], i < nsamp, w < W  ->  out[n][0][w], out[n][1][w].
template <int ITEMS, int CT>
__global__ void __launch_bounds__(RB)
k_col_quantiles(int W, int nsamp, QPos q, const float* __restrict__ mat, float* __restrict__ out) {
  using Sort = cub::BlockRadixSort<float, RB, ITEMS>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  typename Sort::TempStorage& temp = *reinterpret_cast<typename Sort::TempStorage*>(smem_raw);
  float* tile = reinterpret_cast<float*>(smem_raw + ((sizeof(typename Sort::TempStorage) + 15) / 16) * 16);
  __shared__ float pick[4];
  const int n = blockIdx.y;
  const int w0 = blockIdx.x * CT;
  const int wn = min(CT, W - w0);
  const float* __restrict__ m = mat + (long)n * nsamp * W;
  for (int idx = threadIdx.x; idx < nsamp * CT; idx += RB) {
    const int row = idx / CT, c = idx % CT;
    tile[idx] = c < wn ? m[(long)row * W + w0 + c] : 0.f;
  }
  __syncthreads();
  for (int c = 0; c < wn; ++c) {
    float keys[ITEMS];
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
      const int row = threadIdx.x * ITEMS + j;
      keys[j] = row < nsamp ? tile[row * CT + c] : INFINITY;
    }
    Sort(temp).Sort(keys);
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
      const int rank = threadIdx.x * ITEMS + j;
      if (rank == q.lo0) pick[0] = keys[j];
      if (rank == q.lo1) pick[1] = keys[j];
      if (rank == q.hi0) pick[2] = keys[j];
      if (rank == q.hi1) pick[3] = keys[j];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      out[((long)n * 2 + 0) * W + w0 + c] = lerp_np(pick[0], pick[1], q.flo);
      out[((long)n * 2 + 1) * W + w0 + c] = lerp_np(pick[2], pick[3], q.fhi);
    }
    __syncthreads();
  }
}

// ---- selection instead of sorting (nsamp <= 2048, the default niter = 1000) ---------------------------
// np.percentile needs FOUR order statistics per column, not the sorted column.  One WARP per column: every
// lane holds 32 NW keys, transposes each group of 32 into bit planes (plane b = bit b of its 32 keys) and the
// warp runs an MSB-first radix select per rank: per bit one AND, one popc and one warp reduction, no data
// movement.  ~3 k instructions per column against ~60 k for the CTA-wide radix sort.
constexpr int SEL_CT = 8;     // columns (= warps) per CTA
constexpr int SEL_PAD = SEL_CT + 1;

__device__ __forceinline__ uint32_t sortable_bits(float f) {
  const uint32_t u = __float_as_uint(f);
  return u ^ ((u >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}
__device__ __forceinline__ float from_sortable(uint32_t s) {
  return __uint_as_float(s ^ ((s >> 31) ? 0x80000000u : 0xFFFFFFFFu));
}

// In-register 32 x 32 bit-matrix transpose (Hacker's Delight 7-3): afterwards a[31 - b] holds, at bit
// 31 - j, bit b of the original a[j].
__device__ __forceinline__ void transpose32(uint32_t (&a)[32]) {
  uint32_t m = 0x0000FFFFu;
#pragma unroll
  for (int j = 16; j != 0; j >>= 1) {
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      if ((k & j) == 0) {
        const uint32_t t = (a[k] ^ (a[k + j] >> j)) & m;
        a[k] ^= t;
        a[k + j] ^= (t << j);
      }
    }
    m ^= (m << (j >> 1));
  }
}

// MSB-first radix select of NR order statistics AT ONCE: per bit one AND + popc per plane word and ONE warp reduction
// (redux.sync, a single instruction for integers) per rank; the NR chains are independent, so their dependent
// popc -> redux -> compare -> mask steps interleave (the four ranks np.percentile needs were four sequential 32-step
// chains of five shuffles each before: 4.84 ms for 2000 series x 1000 draws x 395 steps).
template <int NW, int NR>
__device__ __forceinline__ void warp_select_n(const uint32_t (&plane)[NW][32], const int (&rank)[NR], uint32_t (&key)[NR]) {
  uint32_t mask[NR][NW];
  int remaining[NR];
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    key[r] = 0u;
    remaining[r] = rank[r];
#pragma unroll
    for (int w = 0; w < NW; ++w) mask[r][w] = 0xFFFFFFFFu;
  }
#pragma unroll
  for (int b = 31; b >= 0; --b) {
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      uint32_t z[NW];
      unsigned cz = 0;
#pragma unroll
      for (int w = 0; w < NW; ++w) {
        z[w] = mask[r][w] & ~plane[w][31 - b];   // candidates with bit b clear
        cz += __popc(z[w]);
      }
      cz = __reduce_add_sync(0xffffffffu, cz);
      const bool zero = remaining[r] < (int)cz;   // the wanted key has bit b clear
      if (!zero) {
        remaining[r] -= (int)cz;
        key[r] |= (1u << b);
      }
#pragma unroll
      for (int w = 0; w < NW; ++w) mask[r][w] = zero ? z[w] : (mask[r][w] & plane[w][31 - b]);
    }
  }
}

template <int NW>
__global__ void __launch_bounds__(SEL_CT * 32)
k_col_quantiles_sel(int W, int nsamp, QPos q, const float* __restrict__ mat, float* __restrict__ out) {
  extern __shared__ __align__(16) float tile[];     // [nsamp][SEL_PAD]
  const int n = blockIdx.y;
  const int w0 = blockIdx.x * SEL_CT;
  const int wn = min(SEL_CT, W - w0);
  const float* __restrict__ m = mat + (long)n * nsamp * W;
  for (int idx = threadIdx.x; idx < nsamp * SEL_CT; idx += SEL_CT * 32) {
    const int row = idx / SEL_CT, c = idx % SEL_CT;
    tile[row * SEL_PAD + c] = c < wn ? m[(long)row * W + w0 + c] : 0.f;
  }
  __syncthreads();
  const int c = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (c >= wn) return;
  uint32_t plane[NW][32];
#pragma unroll
  for (int w = 0; w < NW; ++w) {
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const int row = (w * 32 + j) * 32 + lane;
      plane[w][j] = row < nsamp ? sortable_bits(tile[row * SEL_PAD + c]) : 0xFFFFFFFFu;   // padding sorts last
    }
    transpose32(plane[w]);
  }
  const int ranks[4] = {q.lo0, q.lo1, q.hi0, q.hi1};
  uint32_t keys[4];
  warp_select_n<NW, 4>(plane, ranks, keys);
  const float lo0 = from_sortable(keys[0]), lo1 = from_sortable(keys[1]);
  const float hi0 = from_sortable(keys[2]), hi1 = from_sortable(keys[3]);
  if (lane == 0) {
    out[((long)n * 2 + 0) * W + w0 + c] = lerp_np(lo0, lo1, q.flo);
    out[((long)n * 2 + 1) * W + w0 + c] = lerp_np(hi0, hi1, q.fhi);
  }
}

template <int NW>
static int launch_col_quantiles_sel(int N, int W, int nsamp, const QPos& q, const float* mat, float* out,
                                    cudaStream_t st) {
  const size_t smem = (size_t)nsamp * SEL_PAD * sizeof(float);
  auto kern = k_col_quantiles_sel<NW>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(CI_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  }
  kern<<<dim3((W + SEL_CT - 1) / SEL_CT, N), SEL_CT * 32, smem, st>>>(W, nsamp, q, mat, out);
  return CI_OK;
}

template <int ITEMS, int CT>
static int launch_col_quantiles_t(int N, int W, int nsamp, const QPos& q, const float* mat, float* out,
                                  cudaStream_t st) {
  using Sort = cub::BlockRadixSort<float, RB, ITEMS>;
  const size_t smem = ((sizeof(typename Sort::TempStorage) + 15) / 16) * 16 + (size_t)nsamp * CT * sizeof(float);
  auto kern = k_col_quantiles<ITEMS, CT>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(CI_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  }
  kern<<<dim3((W + CT - 1) / CT, N), RB, smem, st>>>(W, nsamp, q, mat, out);
  return CI_OK;
}

static int launch_col_quantiles(int N, int W, int nsamp, double alpha, const float* mat, float* out,
                                cudaStream_t st) {
  if (W <= 0) return CI_OK;
  const QPos q = quantile_positions(nsamp, alpha);
  if (nsamp <= 1024) return launch_col_quantiles_sel<1>(N, W, nsamp, q, mat, out, st);
  if (nsamp <= 2048) return launch_col_quantiles_sel<2>(N, W, nsamp, q, mat, out, st);
  if (nsamp <= RB * 4) return launch_col_quantiles_t<4, 8>(N, W, nsamp, q, mat, out, st);
  if (nsamp <= RB * 8) return launch_col_quantiles_t<8, 8>(N, W, nsamp, q, mat, out, st);
  if (nsamp <= RB * 16) return launch_col_quantiles_t<16, 4>(N, W, nsamp, q, mat, out, st);
  if (nsamp <= RB * 32) return launch_col_quantiles_t<32, 4>(N, W, nsamp, q, mat, out, st);
  return fail(CI_ERR_UNSUPPORTED, "niter=%d exceeds the %d draws one sorting CTA holds", nsamp, RB * 32);
}

// Compacted running sums over the unmasked post steps, one (series, sample) row per thread:
// cs[n][i][c] = cumsum(sims)[c],  ce[n][i][c] = cumsum(y_post - sims)[c]   (inferences.py:159,190)
// The 128 rows of a CTA are one contiguous run of 128 * Tf floats in `sims`, `cs` and `ce`: the run is copied through
// shared memory with coalesced accesses (row pitch Tf | 1 keeps a thread's walk over its own row conflict-free), the
// running sums replace the samples in place.  (Row-per-thread global accesses took 0.97 ms for 2000 series x 1000
// samples x 30 steps -- a quarter of ci_reduce_inferences.)  Tf > CUMSUM_MAX_TF keeps the direct form.
#define CUMSUM_MAX_TF 96
__global__ void __launch_bounds__(128)
k_cumsum_post_tiled(int N, int Tf, int nsamp, const float* __restrict__ y_post, const float* __restrict__ sims,
                    float* __restrict__ cs, float* __restrict__ ce) {
  extern __shared__ float sm[];                                  // [2][128][Tf | 1]
  const int TP = Tf | 1;
  float* sa = sm;
  float* se = sm + 128 * TP;
  const long rows = (long)N * nsamp;
  const long row0 = (long)blockIdx.x * 128;
  const int nrow = (int)((rows - row0) < 128 ? (rows - row0) : 128);
  const long base = row0 * Tf;
  const int tot = nrow * Tf;
  for (int e = threadIdx.x; e < tot; e += 128) sa[(e / Tf) * TP + e % Tf] = sims[base + e];
  __syncthreads();
  if ((int)threadIdx.x < nrow) {
    const long idx = row0 + threadIdx.x;
    const float* __restrict__ yp = y_post + (idx / nsamp) * Tf;
    float* ra = sa + threadIdx.x * TP;
    float* re = se + threadIdx.x * TP;
    float a = 0.f, ee = 0.f;
    int c = 0;
    for (int t = 0; t < Tf; ++t) {
      const float yv = __ldg(yp + t);
      if (yv == yv) {
        const float sv = ra[t];                                  // c <= t: the slot written below is already consumed
        a += sv;
        ee += yv - sv;
        ra[c] = a;
        re[c] = ee;
        ++c;
      }
    }
    for (; c < Tf; ++c) {
      ra[c] = 0.f;
      re[c] = 0.f;
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < tot; e += 128) {
    const int o = (e / Tf) * TP + e % Tf;
    cs[base + e] = sa[o];
    ce[base + e] = se[o];
  }
}

__global__ void k_cumsum_post(int N, int Tf, int nsamp, const float* __restrict__ y_post,
                              const float* __restrict__ sims, float* __restrict__ cs, float* __restrict__ ce) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)N * nsamp) return;
  const int n = (int)(idx / nsamp);
  const float* __restrict__ yp = y_post + (long)n * Tf;
  const float* __restrict__ s = sims + idx * Tf;
  float a = 0.f, e = 0.f;
  int c = 0;
  for (int t = 0; t < Tf; ++t) {
    const float yv = yp[t];
    if (yv == yv) {
      const float sv = s[t];
      a += sv;
      e += yv - sv;
      cs[idx * Tf + c] = a;
      ce[idx * Tf + c] = e;
      ++c;
    }
  }
  for (; c < Tf; ++c) {
    cs[idx * Tf + c] = 0.f;
    ce[idx * Tf + c] = 0.f;
  }
}

// The 16 columns of inferences.py:229-246 for one series per CTA.
__global__ void k_assemble(int T, int Tf, const float* __restrict__ y_pre, const float* __restrict__ y_post,
                           const float* __restrict__ pre_mean, const float* __restrict__ post_mean,
                           const float* __restrict__ pre_q, const float* __restrict__ post_q,
                           const float* __restrict__ cum_q, const float* __restrict__ ceff_q,
                           float* __restrict__ inf) {
  const int n = blockIdx.x;
  const int W = T + Tf + 1;
  float* __restrict__ o = inf + (long)n * 16 * W;
  const float* yp = y_post + (long)n * Tf;
  for (int i = threadIdx.x; i < 16 * W; i += blockDim.x) o[i] = 0.f;
  __syncthreads();
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    const float y = y_pre[(long)n * T + t], mu = pre_mean[(long)n * T + t];
    const float lo = pre_q[((long)n * 2 + 0) * T + t], hi = pre_q[((long)n * 2 + 1) * T + t];
    o[0 * W + t] = mu; o[1 * W + t] = lo; o[2 * W + t] = hi;
    o[10 * W + t] = y - mu; o[11 * W + t] = y - hi; o[12 * W + t] = y - lo;
  }
  if (threadIdx.x == 0) {
    // serial pass over the (short) post period: compaction + cumulative sums
    float cy = 0.f, cm = 0.f, cem = 0.f;
    int c = 0;
    for (int t = 0; t < Tf; ++t) {
      const float y = yp[t];
      if (!(y == y)) continue;
      const float mu = post_mean[(long)n * Tf + t];
      const float lo = post_q[((long)n * 2 + 0) * Tf + t], hi = post_q[((long)n * 2 + 1) * Tf + t];
      o[0 * W + T + c] = mu; o[1 * W + T + c] = lo; o[2 * W + T + c] = hi;
      o[3 * W + c] = mu; o[4 * W + c] = lo; o[5 * W + c] = hi;
      o[10 * W + T + c] = y - mu; o[11 * W + T + c] = y - hi; o[12 * W + T + c] = y - lo;
      cy += y; cm += mu; cem += y - mu;
      o[6 * W + c + 1] = cy;
      o[7 * W + c + 1] = cm;
      o[8 * W + c + 1] = cum_q[((long)n * 2 + 0) * Tf + c];
      o[9 * W + c + 1] = cum_q[((long)n * 2 + 1) * Tf + c];
      o[13 * W + c + 1] = cem;
      o[14 * W + c + 1] = ceff_q[((long)n * 2 + 0) * Tf + c];
      o[15 * W + c + 1] = ceff_q[((long)n * 2 + 1) * Tf + c];
      ++c;
    }
  }
}

// Summary table + p-value, one series per CTA: row sums over the unmasked steps, one sort.
template <int ITEMS>
__global__ void __launch_bounds__(RB)
k_summary(int Tf, int nsamp, QPos q, const float* __restrict__ y_post, const float* __restrict__ post_mean,
          const float* __restrict__ sims, float* __restrict__ summary) {
  using Sort = cub::BlockRadixSort<float, RB, ITEMS>;
  __shared__ typename Sort::TempStorage temp;
  __shared__ float pick[4];
  __shared__ int cnt[2];
  __shared__ float ysum_s, msum_s;
  __shared__ int nobs_s;
  const int n = blockIdx.x;
  const float* __restrict__ yp = y_post + (long)n * Tf;
  if (threadIdx.x == 0) {
    double ys = 0.0, ms = 0.0;
    int c = 0;
    for (int t = 0; t < Tf; ++t) {
      const float y = yp[t];
      if (y == y) { ys += y; ms += post_mean[(long)n * Tf + t]; ++c; }
    }
    ysum_s = (float)ys; msum_s = (float)ms; nobs_s = c;
    cnt[0] = 0; cnt[1] = 0;
  }
  __syncthreads();
  const float ysum = ysum_s;
  float keys[ITEMS];
  int gt = 0, lt = 0;
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    const int i = threadIdx.x * ITEMS + j;
    if (i < nsamp) {
      const float* __restrict__ s = sims + ((long)n * nsamp + i) * Tf;
      double acc = 0.0;
      for (int t = 0; t < Tf; ++t)
        if (yp[t] == yp[t]) acc += s[t];
      keys[j] = (float)acc;
      gt += keys[j] > ysum;
      lt += keys[j] < ysum;
    } else {
      keys[j] = INFINITY;
    }
  }
  atomicAdd(&cnt[0], gt);
  atomicAdd(&cnt[1], lt);
  Sort(temp).Sort(keys);
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    const int rank = threadIdx.x * ITEMS + j;
    if (rank == q.lo0) pick[0] = keys[j];
    if (rank == q.lo1) pick[1] = keys[j];
    if (rank == q.hi0) pick[2] = keys[j];
    if (rank == q.hi1) pick[3] = keys[j];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const float inv = 1.0f / (float)nobs_s;
    const float sum_y = ysum, mean_y = ysum * inv;
    const float sum_p = msum_s, mean_p = msum_s * inv;
    const float slo = lerp_np(pick[0], pick[1], q.flo), shi = lerp_np(pick[2], pick[3], q.fhi);
    const float mlo = lerp_np(pick[0] * inv, pick[1] * inv, q.flo), mhi = lerp_np(pick[2] * inv, pick[3] * inv, q.fhi);
    float* o = summary + (long)n * 21;
    o[0] = mean_y;                    o[1] = sum_y;
    o[2] = mean_p;                    o[3] = sum_p;
    o[4] = mlo;                       o[5] = slo;
    o[6] = mhi;                       o[7] = shi;
    o[8] = mean_y - mean_p;           o[9] = sum_y - sum_p;
    o[10] = mean_y - mhi;             o[11] = sum_y - shi;
    o[12] = mean_y - mlo;             o[13] = sum_y - slo;
    o[14] = (mean_y - mean_p) / mean_p;  o[15] = (sum_y - sum_p) / sum_p;
    o[16] = (mean_y - mhi) / mean_p;     o[17] = (sum_y - shi) / sum_p;
    o[18] = (mean_y - mlo) / mean_p;     o[19] = (sum_y - slo) / sum_p;
    o[20] = (float)min(cnt[0], cnt[1]) / (float)(nsamp + 1);
  }
}

struct RedWs {
  float* pre_q;   // [N][2][T]
  float* post_q;  // [N][2][Tf]
  float* cum_q;   // [N][2][Tf]
  float* ceff_q;  // [N][2][Tf]
  float* cs;      // [N][nsamp][Tf]
  float* ce;      // [N][nsamp][Tf]
};
static size_t red_ws_floats(const ci_layout* L, int nsamp, RedWs* ws, Carver* c) {
  const size_t N = L->N, T = L->T, Tf = L->Tf > 0 ? L->Tf : 1;
  if (ws && c) {
    ws->pre_q = c->floats(N * 2 * T);
    ws->post_q = c->floats(N * 2 * Tf);
    ws->cum_q = c->floats(N * 2 * Tf);
    ws->ceff_q = c->floats(N * 2 * Tf);
    ws->cs = c->floats(N * nsamp * Tf);
    ws->ce = c->floats(N * nsamp * Tf);
  }
  return align_up(N * 2 * T * 4, 256) + 3 * align_up(N * 2 * Tf * 4, 256) + 2 * align_up(N * nsamp * Tf * 4, 256);
}

}  // namespace ci

using namespace ci;

extern "C" {

size_t ci_reduce_workspace_bytes(const ci_layout* L, int32_t nsamp) {
  if (check_layout(L) || nsamp < 1) return 0;
  return red_ws_floats(L, nsamp, nullptr, nullptr);
}

int ci_reduce_inferences(const ci_layout* L, const float* d_y_pre, const float* d_y_post, const float* d_pre_sims,
                         const float* d_pre_mean, const float* d_post_sims, const float* d_post_mean, int32_t nsamp,
                         float alpha, float* d_inferences, void* d_workspace, size_t workspace_bytes, void* stream) {
  if (int e = check_layout(L)) return e;
  CI_CHECK_ARG(d_y_pre && d_y_post && d_pre_sims && d_pre_mean && d_post_sims && d_post_mean && d_inferences &&
                   d_workspace, "NULL pointer");
  CI_CHECK_ARG(nsamp >= 1 && L->Tf >= 1, "need at least one sample and one post-period step");
  CI_CHECK_ARG(alpha >= 0.f && alpha <= 1.f, "alpha must lie in [0, 1]");
  if (workspace_bytes < ci_reduce_workspace_bytes(L, nsamp))
    return fail(CI_ERR_WORKSPACE_TOO_SMALL, "workspace needs %zu bytes", ci_reduce_workspace_bytes(L, nsamp));
  cudaStream_t st = (cudaStream_t)stream;
  Carver c(d_workspace);
  RedWs ws;
  red_ws_floats(L, nsamp, &ws, &c);
  const long rows = (long)L->N * nsamp;
  if (L->Tf <= CUMSUM_MAX_TF) {
    const size_t smem = (size_t)2 * 128 * (L->Tf | 1) * sizeof(float);
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(k_cumsum_post_tiled, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return fail(CI_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    }
    k_cumsum_post_tiled<<<(unsigned)((rows + 127) / 128), 128, smem, st>>>(L->N, L->Tf, nsamp, d_y_post, d_post_sims, ws.cs, ws.ce);
  } else {
    k_cumsum_post<<<(unsigned)((rows + 127) / 128), 128, 0, st>>>(L->N, L->Tf, nsamp, d_y_post, d_post_sims, ws.cs, ws.ce);
  }
  if (int e = launch_col_quantiles(L->N, L->T, nsamp, alpha, d_pre_sims, ws.pre_q, st)) return e;
  if (int e = launch_col_quantiles(L->N, L->Tf, nsamp, alpha, d_post_sims, ws.post_q, st)) return e;
  if (int e = launch_col_quantiles(L->N, L->Tf, nsamp, alpha, ws.cs, ws.cum_q, st)) return e;
  if (int e = launch_col_quantiles(L->N, L->Tf, nsamp, alpha, ws.ce, ws.ceff_q, st)) return e;
  k_assemble<<<L->N, 128, 0, st>>>(L->T, L->Tf, d_y_pre, d_y_post, d_pre_mean, d_post_mean, ws.pre_q, ws.post_q,
                                   ws.cum_q, ws.ceff_q, d_inferences);
  CI_CHECK_LAUNCH();
  return CI_OK;
}

int ci_reduce_summary(const ci_layout* L, const float* d_y_post, const float* d_post_mean, const float* d_post_sims,
                      int32_t nsamp, float alpha, float* d_summary, void* d_workspace, size_t workspace_bytes,
                      void* stream) {
  (void)d_workspace; (void)workspace_bytes;
  if (int e = check_layout(L)) return e;
  CI_CHECK_ARG(d_y_post && d_post_mean && d_post_sims && d_summary, "NULL pointer");
  CI_CHECK_ARG(nsamp >= 1 && L->Tf >= 1, "need at least one sample and one post-period step");
  CI_CHECK_ARG(alpha >= 0.f && alpha <= 1.f, "alpha must lie in [0, 1]");
  cudaStream_t st = (cudaStream_t)stream;
  const QPos q = quantile_positions(nsamp, alpha);
  if (nsamp <= RB * 4) k_summary<4><<<L->N, RB, 0, st>>>(L->Tf, nsamp, q, d_y_post, d_post_mean, d_post_sims, d_summary);
  else if (nsamp <= RB * 8) k_summary<8><<<L->N, RB, 0, st>>>(L->Tf, nsamp, q, d_y_post, d_post_mean, d_post_sims, d_summary);
  else if (nsamp <= RB * 16) k_summary<16><<<L->N, RB, 0, st>>>(L->Tf, nsamp, q, d_y_post, d_post_mean, d_post_sims, d_summary);
  else if (nsamp <= RB * 32) k_summary<32><<<L->N, RB, 0, st>>>(L->Tf, nsamp, q, d_y_post, d_post_mean, d_post_sims, d_summary);
  else return fail(CI_ERR_UNSUPPORTED, "niter=%d exceeds the %d draws one sorting CTA holds", nsamp, RB * 32);
  CI_CHECK_LAUNCH();
  return CI_OK;
}

}  // extern "C"

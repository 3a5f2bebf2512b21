// Persistent-sampling importance weights, ESS/USS, resampling, trimming, evidence reductions.
// Reference: pocomc/particles.py:215-231, pocomc/tools.py:10-186, pocomc/sampler.py:702-713,
// 739-805, 907-913.  All f64 streaming kernels -- HBM-bound, 16 B per history element per probe.
#include "common.cuh"
#include <algorithm>

namespace pmc {

constexpr int RED_THREADS = 256;

// online (max, sum e, sum e^2) accumulator, e = exp(v - max)
struct Lse3 {
  double m, s1, s2;
  __device__ __forceinline__ void init() { m = -INFINITY; s1 = 0.0; s2 = 0.0; }
  __device__ __forceinline__ void push(double v) {
    if (v == -INFINITY) return;
    if (v <= m) { const double e = exp(v - m); s1 += e; s2 += e * e; }
    else {
      const double c = exp(m - v);  // m == -inf -> 0
      s1 = s1 * c + 1.0; s2 = s2 * c * c + 1.0; m = v;
    }
  }
  __device__ __forceinline__ void merge(const Lse3& o) {
    if (o.m == -INFINITY) return;
    if (m == -INFINITY) { *this = o; return; }
    if (o.m <= m) { const double c = exp(o.m - m); s1 += o.s1 * c; s2 += o.s2 * c * c; }
    else { const double c = exp(m - o.m); s1 = s1 * c + o.s1; s2 = s2 * c * c + o.s2; m = o.m; }
  }
};

__device__ __forceinline__ Lse3 warp_merge(Lse3 a) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    Lse3 b;
    b.m = __shfl_xor_sync(FULL, a.m, o); b.s1 = __shfl_xor_sync(FULL, a.s1, o); b.s2 = __shfl_xor_sync(FULL, a.s2, o);
    // merge in a lane-symmetric way so every lane ends with the same value
    Lse3 lo = (a.m >= b.m) ? a : b, hi = (a.m >= b.m) ? b : a;
    lo.merge(hi);
    a = lo;
  }
  return a;
}

__device__ __forceinline__ Lse3 block_merge(Lse3 a, double* sh /*3*8*/) {
  a = warp_merge(a);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { sh[warp * 3] = a.m; sh[warp * 3 + 1] = a.s1; sh[warp * 3 + 2] = a.s2; }
  __syncthreads();
  Lse3 t; t.init();
  if (threadIdx.x == 0) {
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { Lse3 o{sh[w * 3], sh[w * 3 + 1], sh[w * 3 + 2]}; t.merge(o); }
  }
  return t;  // valid in thread 0
}

// ---- persistent-sampling denominator -------------------------------------------------------
__global__ void __launch_bounds__(256)
ps_append_kernel(const double* __restrict__ logl, double* __restrict__ den, const double* __restrict__ beta,
                 const double* __restrict__ logz, int t_new, int t_total, long long n) {
  const long long total = (long long)t_total * n;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int t = (int)(i / n);
    const double l = logl[i];
    double acc;
    int from;
    if (t < t_new) { acc = den[i]; from = t_new; }
    else { acc = l * beta[0] - logz[0]; from = 1; }       // particles.py:222, i = 0 term
    for (int k = from; k < t_total; ++k) acc = logaddexp(acc, l * beta[k] - logz[k]);
    den[i] = acc;
  }
}

__global__ void __launch_bounds__(RED_THREADS)
ps_reduce_kernel(const double* __restrict__ logl, const double* __restrict__ den, double beta_f, double log_t,
                 long long m, double* __restrict__ scratch) {
  __shared__ double sh[24];
  Lse3 a; a.init();
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += stride)
    a.push(logl[i] * beta_f - (den[i] - log_t));           // particles.py:220-224
  a = block_merge(a, sh);
  if (threadIdx.x == 0) { scratch[blockIdx.x * 3] = a.m; scratch[blockIdx.x * 3 + 1] = a.s1; scratch[blockIdx.x * 3 + 2] = a.s2; }
}

__global__ void ps_combine_kernel(const double* __restrict__ scratch, int n_blocks, double* __restrict__ out4) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    Lse3 t; t.init();
    for (int b = 0; b < n_blocks; ++b) { Lse3 o{scratch[b * 3], scratch[b * 3 + 1], scratch[b * 3 + 2]}; t.merge(o); }
    out4[0] = t.m; out4[1] = t.s1; out4[2] = t.s2; out4[3] = 0.0;
  }
}

__global__ void __launch_bounds__(RED_THREADS)
ps_uss_kernel(const double* __restrict__ logl, const double* __restrict__ den, double beta_f, double log_t,
              long long m, double k, const double* __restrict__ stats, double* __restrict__ scratch) {
  __shared__ double sh[8];
  const double mx = stats[0], s1 = stats[1];
  double acc = 0.0;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += stride) {
    const double w = exp(logl[i] * beta_f - (den[i] - log_t) - mx) / s1;
    acc += 1.0 - pow(1.0 - w, k);                          // tools.py:93
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += sh[w];
    scratch[blockIdx.x] = s;
  }
}

__global__ void sum_combine_kernel(const double* __restrict__ scratch, int n_blocks, double* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double s = 0.0;
    for (int b = 0; b < n_blocks; ++b) s += scratch[b];
    *out = s;
  }
}

__global__ void __launch_bounds__(256)
ps_weights_kernel(const double* __restrict__ logl, const double* __restrict__ den, double beta_f, double log_t,
                  long long m, const double* __restrict__ stats, double* __restrict__ w, double* __restrict__ logw) {
  const double mx = stats[0], s1 = stats[1];
  const double lse = mx + log(s1);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += stride) {
    const double lw = logl[i] * beta_f - (den[i] - log_t);
    if (w) w[i] = exp(lw - mx) / s1;                       // sampler.py:780-781
    if (logw) logw[i] = lw - lse;                          // particles.py:228-229
  }
}

// ---- resampling ------------------------------------------------------------------------------
// np.cumsum order: one thread, strictly sequential f64 adds (loads batched for latency).
__global__ void cumsum_kernel(const double* __restrict__ w, double* __restrict__ cdf, long long m) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double acc = 0.0;
  long long i = 0;
  for (; i + 8 <= m; i += 8) {
    double v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = w[i + k];
#pragma unroll
    for (int k = 0; k < 8; ++k) { acc += v[k]; cdf[i + k] = acc; }
  }
  for (; i < m; ++i) { acc += w[i]; cdf[i] = acc; }
}

__global__ void __launch_bounds__(256)
multinomial_kernel(const double* __restrict__ cdf, const double* __restrict__ r, long long* __restrict__ idx,
                   long long m, long long n_out) {
  const double total = cdf[m - 1];
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_out; i += stride) {
    const double v = r[i];
    long long lo = 0, hi = m;   // first index with cdf[idx]/total > v   (searchsorted side='right')
    while (lo < hi) {
      const long long mid = (lo + hi) >> 1;
      if (cdf[mid] / total <= v) lo = mid + 1; else hi = mid;
    }
    idx[i] = lo;
  }
}

__global__ void __launch_bounds__(256)
systematic_kernel(const double* __restrict__ cdf, double u0, long long* __restrict__ idx, long long m, long long n_out) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_out; i += stride) {
    const double pos = (u0 + (double)i) / (double)n_out;   // tools.py:175
    long long lo = 0, hi = m;   // the walk stops at the first j with pos <= cumsum[j]
    while (lo < hi) {
      const long long mid = (lo + hi) >> 1;
      if (cdf[mid] < pos) lo = mid + 1; else hi = mid;
    }
    idx[i] = lo < m ? lo : m - 1;
  }
}

__global__ void __launch_bounds__(256)
gather_rows_kernel(const double* __restrict__ src, const long long* __restrict__ idx, double* __restrict__ dst,
                   long long n_out, int d) {
  const long long total = n_out * d;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const long long r = i / d;
    dst[i] = src[idx[r] * d + (i - r * d)];
  }
}

// ---- trimming --------------------------------------------------------------------------------
constexpr int TRIM_CHUNK = 4096;

__global__ void __launch_bounds__(256)
trim_chunk_sums_kernel(const double* __restrict__ ws, long long m, double* __restrict__ c1, double* __restrict__ c2) {
  __shared__ double sh[16];
  const long long base = (long long)blockIdx.x * TRIM_CHUNK;
  double a = 0.0, b = 0.0;
  for (int k = threadIdx.x; k < TRIM_CHUNK; k += blockDim.x) {
    const long long i = base + k;
    if (i < m) { const double v = ws[i]; a += v; b += v * v; }
  }
  a = warp_sum(a); b = warp_sum(b);
  if ((threadIdx.x & 31) == 0) { sh[threadIdx.x >> 5] = a; sh[8 + (threadIdx.x >> 5)] = b; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0, q = 0.0;
    for (int w = 0; w < 8; ++w) { s += sh[w]; q += sh[8 + w]; }
    c1[blockIdx.x] = s; c2[blockIdx.x] = q;
  }
}

// in-place exclusive suffix: c[k] <- sum_{j>k} c[j]; total in c[n_chunks] slot
__global__ void trim_suffix_kernel(double* __restrict__ c1, double* __restrict__ c2, int n_chunks) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double a = 0.0, b = 0.0;
  for (int k = n_chunks - 1; k >= 0; --k) {
    const double x = c1[k], y = c2[k];
    c1[k] = a; c2[k] = b;
    a += x; b += y;
  }
  c1[n_chunks] = a; c2[n_chunks] = b;
}

// one warp per percentile grid point; grid point g passes if ESS(w >= thr_g)/ESS(all) >= ess_frac
__global__ void __launch_bounds__(256)
trim_grid_kernel(const double* __restrict__ ws, long long m, const double* __restrict__ c1,
                 const double* __restrict__ c2, int n_chunks, double ess_frac, int bins,
                 double* __restrict__ thr_out, double* __restrict__ keep_out, int* __restrict__ pass_out) {
  const int lane = threadIdx.x & 31;
  const int g = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (g >= bins) return;
  // np.linspace(0, 99, bins)[g]; np.percentile(..., method='linear')
  const double step = 99.0 / (double)(bins - 1);
  const double p = (g == bins - 1) ? 99.0 : (double)g * step;
  const double vidx = (double)(m - 1) * (p / 100.0);
  long long lo = (long long)floor(vidx);
  if (lo > m - 1) lo = m - 1;
  const long long nx = lo + 1 < m ? lo + 1 : m - 1;
  const double t = vidx - (double)lo;
  const double a = ws[lo], b = ws[nx], diff = b - a;
  double thr = a + diff * t;
  if (t >= 0.5) thr = b - diff * (1.0 - t);
  long long l = 0, h = m;      // first sorted index with ws[idx] >= thr
  while (l < h) { const long long mid = (l + h) >> 1; if (ws[mid] < thr) l = mid + 1; else h = mid; }
  const long long first = l;
  const int chunk = (int)(first / TRIM_CHUNK);
  const long long end = min((long long)(chunk + 1) * TRIM_CHUNK, m);
  double s = 0.0, q = 0.0;
  for (long long i = first + lane; i < end; i += 32) { const double v = ws[i]; s += v; q += v * v; }
  s = warp_sum(s) + c1[chunk]; q = warp_sum(q) + c2[chunk];
  if (lane == 0) {
    const double ess_total = 1.0 / c2[n_chunks];            // weights are normalised: sum w = 1
    const double wn = s;                                     // trimmed weights renormalise by s
    const double ess_trim = (wn * wn) / q;
    thr_out[g] = thr; keep_out[g] = s;
    pass_out[g] = (first < m) && (ess_trim / ess_total >= ess_frac);
  }
}

__global__ void trim_pick_kernel(const double* __restrict__ thr, const double* __restrict__ keep,
                                 const int* __restrict__ pass, int bins, double* __restrict__ out3) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  int g = bins - 1;
  while (g > 0 && !pass[g]) --g;    // tools.py:39-51 walks down from the top; grid point 0 keeps everything
  out3[0] = thr[g]; out3[1] = keep[g]; out3[2] = (double)g;
}

// ---- evidence ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(RED_THREADS)
lse_partial_kernel(const double* __restrict__ v, long long n, double* __restrict__ scratch) {
  __shared__ double sh[24];
  Lse3 a; a.init();
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) a.push(v[i]);
  a = block_merge(a, sh);
  if (threadIdx.x == 0) { scratch[blockIdx.x * 3] = a.m; scratch[blockIdx.x * 3 + 1] = a.s1; scratch[blockIdx.x * 3 + 2] = a.s2; }
}

__global__ void lse_final_kernel(const double* __restrict__ scratch, int n_blocks, long long n, double* __restrict__ out2) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    Lse3 t; t.init();
    for (int b = 0; b < n_blocks; ++b) { Lse3 o{scratch[b * 3], scratch[b * 3 + 1], scratch[b * 3 + 2]}; t.merge(o); }
    out2[0] = t.m + log(t.s1) - log((double)n);            // sampler.py:910
    out2[1] = t.m;
  }
}

__global__ void __launch_bounds__(RED_THREADS)
lse_bootstrap_kernel(const double* __restrict__ logw, const long long* __restrict__ idx, long long n,
                     long long n_boot, double* __restrict__ out) {
  __shared__ double sh[24];
  for (long long b = blockIdx.x; b < n_boot; b += gridDim.x) {
    Lse3 a; a.init();
    const long long* row = idx + b * n;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) a.push(logw[row[i]]);
    a = block_merge(a, sh);
    if (threadIdx.x == 0) out[b] = a.m + log(a.s1) - log((double)n);   // sampler.py:913
    __syncthreads();
  }
}


// same statistic with the resampling indices drawn ON the device: row b uses idx_i = floor(u_i * n) with u_i from a
// Philox4x32-10 counter keyed by (seed, b, i / 4) -- no [n_boot, n] index matrix exists anywhere (sampler.py:913 draws
// it row by row with np.random.choice; section 8 f2)
__device__ __forceinline__ uint4 philox4x32(uint4 c, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0);
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return c;
}
__global__ void __launch_bounds__(RED_THREADS)
lse_bootstrap_rng_kernel(const double* __restrict__ logw, long long n, long long n_boot, unsigned long long seed,
                         double* __restrict__ out) {
  __shared__ double sh[24];
  for (long long b = blockIdx.x; b < n_boot; b += gridDim.x) {
    Lse3 a; a.init();
    for (long long i4 = threadIdx.x; i4 * 4 < n; i4 += blockDim.x) {
      const uint4 r = philox4x32(make_uint4((uint32_t)i4, (uint32_t)(i4 >> 32), (uint32_t)b, (uint32_t)(b >> 32) ^ 0x626f6f74u),
                                 (uint32_t)seed, (uint32_t)(seed >> 32));
      const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (i4 * 4 + q < n) {
          long long j = (long long)(((unsigned long long)rr[q] * (unsigned long long)n) >> 32);   // floor(u n), u = r / 2^32
          a.push(logw[j < n ? j : n - 1]);
        }
      }
    }
    a = block_merge(a, sh);
    if (threadIdx.x == 0) out[b] = a.m + log(a.s1) - log((double)n);
    __syncthreads();
  }
}


// ---- plain weight statistics (tools.py:56-93 on an explicit weight vector) -------------------
__global__ void __launch_bounds__(RED_THREADS)
wstats_partial_kernel(const double* __restrict__ w, long long m, double* __restrict__ scratch) {
  __shared__ double sh[16];
  double a = 0.0, b = 0.0;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += stride) { const double v = w[i]; a += v; b += v * v; }
  a = warp_sum(a); b = warp_sum(b);
  if ((threadIdx.x & 31) == 0) { sh[threadIdx.x >> 5] = a; sh[8 + (threadIdx.x >> 5)] = b; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0, q = 0.0;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) { s += sh[k]; q += sh[8 + k]; }
    scratch[2 * blockIdx.x] = s; scratch[2 * blockIdx.x + 1] = q;
  }
}
__global__ void wstats_final_kernel(const double* __restrict__ scratch, int n_blocks, double* __restrict__ out3) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double s = 0.0, q = 0.0;
    for (int b = 0; b < n_blocks; ++b) { s += scratch[2 * b]; q += scratch[2 * b + 1]; }
    out3[0] = s; out3[1] = q; out3[2] = 0.0;
  }
}
__global__ void __launch_bounds__(RED_THREADS)
wstats_uss_kernel(const double* __restrict__ w, long long m, double k, const double* __restrict__ stats,
                  double* __restrict__ scratch) {
  __shared__ double sh[8];
  const double tot = stats[0];
  double acc = 0.0;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += stride)
    acc += 1.0 - pow(1.0 - w[i] / tot, k);
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int j = 0; j < (int)(blockDim.x >> 5); ++j) s += sh[j];
    scratch[blockIdx.x] = s;
  }
}

static inline int red_blocks(long long m) {
  long long b = (m + RED_THREADS * 8 - 1) / (RED_THREADS * 8);
  const long long cap = (long long)sm_count() * 4;
  if (b > cap) b = cap;
  return (int)(b < 1 ? 1 : b);
}

}  // namespace pmc

using namespace pmc;

extern "C" int pmc_ps_append(const double* logl, double* den, const double* beta, const double* logz, int32_t t_new,
                             int32_t t_total, int64_t n, pmc_stream_t stream) {
  PMC_REQUIRE(logl && den && beta && logz, "pmc_ps_append: null pointer");
  PMC_REQUIRE(t_new >= 0 && t_new <= t_total && t_total >= 1, "pmc_ps_append: bad iteration range");
  if (n == 0 || t_new == t_total) return 0;
  ps_append_kernel<<<grid_for((long long)t_total * n, 256, 8), 256, 0, as_stream(stream)>>>(logl, den, beta, logz, t_new, t_total, n);
  PMC_LAUNCH_CHECK();
  return 0;
}

extern "C" int64_t pmc_ps_scratch_size(int64_t m) { return 3 * (int64_t)red_blocks(m) + 8; }

extern "C" int pmc_ps_reduce(const double* logl, const double* den, double beta_f, int32_t t_total, int64_t n,
                             int64_t uss_k, double* scratch, double* out4, pmc_stream_t stream) {
  PMC_REQUIRE(logl && den && scratch && out4 && t_total >= 1 && n >= 1, "pmc_ps_reduce: bad arguments");
  const long long m = (long long)t_total * n;
  const int nb = red_blocks(m);
  const double log_t = log((double)t_total);
  cudaStream_t st = as_stream(stream);
  ps_reduce_kernel<<<nb, RED_THREADS, 0, st>>>(logl, den, beta_f, log_t, m, scratch);
  ps_combine_kernel<<<1, 32, 0, st>>>(scratch, nb, out4);
  if (uss_k > 0) {
    ps_uss_kernel<<<nb, RED_THREADS, 0, st>>>(logl, den, beta_f, log_t, m, (double)uss_k, out4, scratch);
    sum_combine_kernel<<<1, 32, 0, st>>>(scratch, nb, out4 + 3);
  }
  PMC_LAUNCH_CHECK();
  return 0;
}

extern "C" int pmc_ps_weights(const double* logl, const double* den, double beta_f, int32_t t_total, int64_t n,
                              const double* stats4, double* w, double* logw, pmc_stream_t stream) {
  PMC_REQUIRE(logl && den && stats4 && (w || logw), "pmc_ps_weights: bad arguments");
  const long long m = (long long)t_total * n;
  ps_weights_kernel<<<grid_for(m, 256, 8), 256, 0, as_stream(stream)>>>(logl, den, beta_f, log((double)t_total), m, stats4, w, logw);
  PMC_LAUNCH_CHECK();
  return 0;
}


extern "C" int pmc_weight_stats(const double* w, int64_t m, int64_t uss_k, double* scratch, double* out3,
                                pmc_stream_t stream) {
  PMC_REQUIRE(w && scratch && out3 && m >= 1, "pmc_weight_stats: bad arguments");
  const int nb = red_blocks(m);
  cudaStream_t st = as_stream(stream);
  wstats_partial_kernel<<<nb, RED_THREADS, 0, st>>>(w, m, scratch);
  wstats_final_kernel<<<1, 32, 0, st>>>(scratch, nb, out3);
  if (uss_k > 0) {
    wstats_uss_kernel<<<nb, RED_THREADS, 0, st>>>(w, m, (double)uss_k, out3, scratch);
    sum_combine_kernel<<<1, 32, 0, st>>>(scratch, nb, out3 + 2);
  }
  PMC_LAUNCH_CHECK();
  return 0;
}

extern "C" int pmc_cumsum_f64(const double* w, double* cdf, int64_t m, pmc_stream_t stream) {
  PMC_REQUIRE(w && cdf && m >= 1, "pmc_cumsum_f64: bad arguments");
  cumsum_kernel<<<1, 32, 0, as_stream(stream)>>>(w, cdf, m);
  PMC_LAUNCH_CHECK();
  return 0;
}

extern "C" int pmc_resample_multinomial(const double* cdf, const double* r, int64_t* idx, int64_t m, int64_t n_out,
                                        pmc_stream_t stream) {
  PMC_REQUIRE(cdf && r && idx && m >= 1, "pmc_resample_multinomial: bad arguments");
  if (n_out == 0) return 0;
  multinomial_kernel<<<grid_for(n_out, 256, 8), 256, 0, as_stream(stream)>>>(cdf, r, (long long*)idx, m, n_out);
  PMC_LAUNCH_CHECK();
  return 0;
}

extern "C" int pmc_resample_systematic(const double* cdf, double u0, int64_t* idx, int64_t m, int64_t n_out,
                                       pmc_stream_t stream) {
  PMC_REQUIRE(cdf && idx && m >= 1, "pmc_resample_systematic: bad arguments");
  if (n_out == 0) return 0;
  systematic_kernel<<<grid_for(n_out, 256, 8), 256, 0, as_stream(stream)>>>(cdf, u0, (long long*)idx, m, n_out);
  PMC_LAUNCH_CHECK();
  return 0;
}

extern "C" int pmc_gather_rows_f64(const double* src, const int64_t* idx, double* dst, int64_t n_out, int32_t d,
                                   pmc_stream_t stream) {
  PMC_REQUIRE(src && idx && dst && d >= 1, "pmc_gather_rows_f64: bad arguments");
  if (n_out == 0) return 0;
  gather_rows_kernel<<<grid_for(n_out * d, 256, 8), 256, 0, as_stream(stream)>>>(src, (const long long*)idx, dst, n_out, d);
  PMC_LAUNCH_CHECK();
  return 0;
}

extern "C" int64_t pmc_trim_scratch_size(int64_t m) {
  const int64_t chunks = (m + TRIM_CHUNK - 1) / TRIM_CHUNK;
  return 2 * (chunks + 1) + 3 * 65536;   // chunk sums + per-grid-point (thr, keep, pass) for bins <= 65536
}

extern "C" int pmc_trim_threshold(const double* ws, int64_t m, double ess_frac, int32_t bins, double* scratch,
                                  double* out3, pmc_stream_t stream) {
  PMC_REQUIRE(ws && scratch && out3 && m >= 1, "pmc_trim_threshold: bad arguments");
  PMC_REQUIRE(bins >= 2 && bins <= 65536, "pmc_trim_threshold: bins out of range");
  const int chunks = (int)((m + TRIM_CHUNK - 1) / TRIM_CHUNK);
  double* c1 = scratch;
  double* c2 = scratch + (chunks + 1);
  double* thr = scratch + 2 * (chunks + 1);
  double* keep = thr + 65536;
  int* pass = reinterpret_cast<int*>(keep + 65536);
  cudaStream_t st = as_stream(stream);
  trim_chunk_sums_kernel<<<chunks, 256, 0, st>>>(ws, m, c1, c2);
  trim_suffix_kernel<<<1, 32, 0, st>>>(c1, c2, chunks);
  trim_grid_kernel<<<(bins * 32 + 255) / 256, 256, 0, st>>>(ws, m, c1, c2, chunks, ess_frac, bins, thr, keep, pass);
  trim_pick_kernel<<<1, 32, 0, st>>>(thr, keep, pass, bins, out3);
  PMC_LAUNCH_CHECK();
  return 0;
}

extern "C" int pmc_lse(const double* logw, int64_t n, double* scratch, double* out2, pmc_stream_t stream) {
  PMC_REQUIRE(logw && scratch && out2 && n >= 1, "pmc_lse: bad arguments");
  const int nb = red_blocks(n);
  lse_partial_kernel<<<nb, RED_THREADS, 0, as_stream(stream)>>>(logw, n, scratch);
  lse_final_kernel<<<1, 32, 0, as_stream(stream)>>>(scratch, nb, n, out2);
  PMC_LAUNCH_CHECK();
  return 0;
}

extern "C" int pmc_lse_bootstrap(const double* logw, const int64_t* idx, int64_t n, int64_t n_boot, double* out,
                                 pmc_stream_t stream) {
  PMC_REQUIRE(logw && idx && out && n >= 1, "pmc_lse_bootstrap: bad arguments");
  if (n_boot == 0) return 0;
  const int blocks = (int)std::min<long long>(n_boot, (long long)sm_count() * 8);
  lse_bootstrap_kernel<<<blocks, RED_THREADS, 0, as_stream(stream)>>>(logw, (const long long*)idx, n, n_boot, out);
  PMC_LAUNCH_CHECK();
  return 0;
}

extern "C" int pmc_lse_bootstrap_rng(const double* logw, int64_t n, int64_t n_boot, uint64_t seed, double* out,
                                     pmc_stream_t stream) {
  PMC_REQUIRE(logw && out && n >= 1 && n < (1ll << 32), "pmc_lse_bootstrap_rng: bad arguments");
  if (n_boot == 0) return 0;
  const int blocks = (int)std::min<long long>(n_boot, (long long)sm_count() * 8);
  lse_bootstrap_rng_kernel<<<blocks, RED_THREADS, 0, as_stream(stream)>>>(logw, n, n_boot, (unsigned long long)seed, out);
  PMC_LAUNCH_CHECK();
  return 0;
}

// Proposal geometry on the device: the O(n D^2) parts of pocomc/geometry.py:31-59 (np.average / np.cov
// with aweights) and of the multivariate Student-t EM fit pocomc/student.py:5-85 (initial moments, Mahalanobis distances,
// the sums behind the degrees-of-freedom score, the weighted scatter update) for clouds of up to 10^6 x 200 particles.
// Everything is f64 FMA work (the reference is f64 and the bar is 1e-12) with FIXED-ORDER two-stage reductions: a result
// does not depend on the number of SMs, so every rank of a sharded run fits the same geometry.  The D x D algebra
// (inverse, eigenvalue bound, digamma / bisection on scalars) stays on the host like the reference's.
#include "common.cuh"
#include <algorithm>

namespace pmc {

constexpr int GEO_ROWS = 2048;       // rows per chunk of the column sums
constexpr int SYRK_ROWS = 8192;      // rows per chunk of the scatter matrix
constexpr int SYRK_TILE = 64;        // output tile (64 x 64 per CTA, 4 x 4 per thread)
constexpr int SYRK_KB = 16;          // rows staged per step

// partial[c][0..d) = sum_i w_i x_i[j], partial[c][d] = sum w, partial[c][d+1] = sum w^2, partial[c][d+2] = max_i |x_i - center|^2
// over the rows of chunk c (w == nullptr: w_i = 1; center == nullptr: the max is not computed)
__global__ void __launch_bounds__(256)
geo_colsum_kernel(const double* __restrict__ x, const double* __restrict__ w, const double* __restrict__ center, long long n, int d,
                  double* __restrict__ partial) {
  __shared__ double red[256];
  const long long r0 = (long long)blockIdx.x * GEO_ROWS, r1 = min(n, r0 + GEO_ROWS);
  double* out = partial + (size_t)blockIdx.x * (d + 3);
  for (int j = threadIdx.x; j < d; j += blockDim.x) {
    double s = 0.0;
    for (long long r = r0; r < r1; ++r) s = fma(w ? w[r] : 1.0, x[(size_t)r * d + j], s);
    out[j] = s;
  }
  // scalars: threads stride the rows, then a fixed-shape tree
  double sw = 0.0, sw2 = 0.0, mx = 0.0;
  for (long long r = r0 + threadIdx.x; r < r1; r += blockDim.x) {
    const double wi = w ? w[r] : 1.0;
    sw += wi; sw2 += wi * wi;
    if (center) {
      double q = 0.0;
      for (int j = 0; j < d; ++j) { const double t = x[(size_t)r * d + j] - center[j]; q = fma(t, t, q); }
      mx = fmax(mx, q);
    }
  }
  for (int pass = 0; pass < 3; ++pass) {
    red[threadIdx.x] = pass == 0 ? sw : (pass == 1 ? sw2 : mx);
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
      if ((int)threadIdx.x < s) red[threadIdx.x] = pass == 2 ? fmax(red[threadIdx.x], red[threadIdx.x + s]) : red[threadIdx.x] + red[threadIdx.x + s];
      __syncthreads();
    }
    if (threadIdx.x == 0) out[d + pass] = red[0];
    __syncthreads();
  }
}

// out[j] = sum over chunks (in chunk order) of partial[c][j]; the max entry takes the maximum
__global__ void geo_reduce_kernel(const double* __restrict__ partial, int n_chunks, int width, int max_col, double* __restrict__ out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= width) return;
  double s = 0.0;
  for (int c = 0; c < n_chunks; ++c) {
    const double v = partial[(size_t)c * width + j];
    s = (j == max_col) ? fmax(s, v) : s + v;
  }
  out[j] = s;
}

// scatter[ti][tj] partial of chunk c: sum_i w_i (x_i - center)[a] (x_i - center)[b] for a in tile ti, b in tile tj (ti <= tj)
__global__ void __launch_bounds__(256)
geo_syrk_kernel(const double* __restrict__ x, const double* __restrict__ w, const double* __restrict__ center, long long n, int d,
                int nt, double* __restrict__ partial) {
  __shared__ double As[SYRK_KB][SYRK_TILE + 1], Bs[SYRK_KB][SYRK_TILE + 1];
  // upper-triangular tile index -> (ti, tj)
  int t = blockIdx.x, ti = 0;
  while (t >= nt - ti) { t -= nt - ti; ++ti; }
  const int tj = ti + t;
  const int a0 = ti * SYRK_TILE, b0 = tj * SYRK_TILE;
  const long long r0 = (long long)blockIdx.y * SYRK_ROWS, r1 = min(n, r0 + SYRK_ROWS);
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  double acc[4][4] = {};
  for (long long rb = r0; rb < r1; rb += SYRK_KB) {
    for (int e = threadIdx.x; e < SYRK_KB * SYRK_TILE; e += blockDim.x) {
      const int k = e / SYRK_TILE, c = e - k * SYRK_TILE;
      const long long r = rb + k;
      double va = 0.0, vb = 0.0;
      if (r < r1) {
        const double wi = w ? w[r] : 1.0;
        if (a0 + c < d) va = wi * (x[(size_t)r * d + a0 + c] - center[a0 + c]);
        if (b0 + c < d) vb = x[(size_t)r * d + b0 + c] - center[b0 + c];
      }
      As[k][c] = va; Bs[k][c] = vb;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SYRK_KB; ++k) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = As[k][ty + 16 * i]; b[i] = Bs[k][tx + 16 * i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  double* out = partial + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * (SYRK_TILE * SYRK_TILE);
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) out[(ty + 16 * i) * SYRK_TILE + tx + 16 * j] = acc[i][j];
}

// C[a][b] = C[b][a] = sum over chunks (in order) of the tile partials
__global__ void geo_syrk_reduce_kernel(const double* __restrict__ partial, int n_chunks, int n_tiles, int nt, int d, double* __restrict__ C) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_tiles * SYRK_TILE * SYRK_TILE) return;
  int t = e / (SYRK_TILE * SYRK_TILE), ti = 0;
  const int within = e - t * (SYRK_TILE * SYRK_TILE);
  const int tile_id = t;
  while (t >= nt - ti) { t -= nt - ti; ++ti; }
  const int tj = ti + t;
  const int a = ti * SYRK_TILE + within / SYRK_TILE, b = tj * SYRK_TILE + within % SYRK_TILE;
  if (a >= d || b >= d) return;
  double s = 0.0;
  for (int c = 0; c < n_chunks; ++c) s += partial[((size_t)c * n_tiles + tile_id) * (SYRK_TILE * SYRK_TILE) + within];
  if (ti != tj || a <= b) { C[(size_t)a * d + b] = s; C[(size_t)b * d + a] = s; }
}

// delta_i = (x_i - center)^T P (x_i - center), P symmetric [d, d]; 64 rows per CTA, P streamed once per CTA
__global__ void __launch_bounds__(256)
geo_mahalanobis_kernel(const double* __restrict__ x, const double* __restrict__ center, const double* __restrict__ P, long long n, int d,
                       double* __restrict__ delta) {
  extern __shared__ double sm[];            // diff [64][d + 1]
  __shared__ double rowsum[64][17];
  const int ld = d + 1;
  const long long r0 = (long long)blockIdx.x * 64;
  for (int e = threadIdx.x; e < 64 * d; e += blockDim.x) {
    const int r = e / d, j = e - r * d;
    sm[r * ld + j] = (r0 + r < n) ? x[(size_t)(r0 + r) * d + j] - center[j] : 0.0;
  }
  __syncthreads();
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;      // rows ty + 16 i, columns j0 + tx + 16 j
  double q[4] = {0.0, 0.0, 0.0, 0.0};
  for (int j0 = 0; j0 < d; j0 += 64) {
    double acc[4][4] = {};
    for (int k = 0; k < d; ++k) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = sm[(ty + 16 * i) * ld + k];
#pragma unroll
      for (int j = 0; j < 4; ++j) { const int c = j0 + tx + 16 * j; b[j] = c < d ? P[(size_t)k * d + c] : 0.0; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) { const int c = j0 + tx + 16 * j; if (c < d) q[i] = fma(acc[i][j], sm[(ty + 16 * i) * ld + c], q[i]); }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) rowsum[ty + 16 * i][tx] = q[i];
  __syncthreads();
  if (threadIdx.x < 64 && r0 + threadIdx.x < n) {
    double s = 0.0;
    for (int c = 0; c < 16; ++c) s += rowsum[threadIdx.x][c];
    delta[r0 + threadIdx.x] = s;
  }
}

// student.py:42-51,56: w_i = (nu + dim) / (nu + delta_i); partial[c] = (sum log w, sum w); optionally stores w
__global__ void __launch_bounds__(256)
geo_tweights_kernel(const double* __restrict__ delta, long long n, double nu, double dim, double* __restrict__ w_out, double* __restrict__ partial) {
  __shared__ double red[2][256];
  const long long r0 = (long long)blockIdx.x * GEO_ROWS, r1 = min(n, r0 + GEO_ROWS);
  double sl = 0.0, sw = 0.0;
  for (long long r = r0 + threadIdx.x; r < r1; r += blockDim.x) {
    const double wi = (nu + dim) / (nu + delta[r]);
    sl += log(wi); sw += wi;
    if (w_out) w_out[r] = wi;
  }
  red[0][threadIdx.x] = sl; red[1][threadIdx.x] = sw;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) { red[0][threadIdx.x] += red[0][threadIdx.x + s]; red[1][threadIdx.x] += red[1][threadIdx.x + s]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { partial[2 * blockIdx.x] = red[0][0]; partial[2 * blockIdx.x + 1] = red[1][0]; }
}

static inline int geo_chunks(long long n, int rows) { return (int)((n + rows - 1) / rows); }
static inline int syrk_nt(int d) { return (d + SYRK_TILE - 1) / SYRK_TILE; }

}  // namespace pmc

using namespace pmc;

/* doubles of scratch for the calls below on an [n, d] cloud */
extern "C" int64_t pmc_geometry_scratch_size(int64_t n, int32_t d) {
  const int nt = syrk_nt(d);
  const int64_t a = (int64_t)geo_chunks(n, GEO_ROWS) * (d + 3);
  const int64_t b = (int64_t)geo_chunks(n, SYRK_ROWS) * (nt * (nt + 1) / 2) * SYRK_TILE * SYRK_TILE;
  return std::max<int64_t>(std::max(a, b), 16);
}

/* out[0..d) = sum_i w_i x_i, out[d] = sum w, out[d+1] = sum w^2, out[d+2] = max_i |x_i - center|^2 (0 without a center) */
extern "C" int pmc_weighted_colsums(const double* x, const double* w, const double* center, int64_t n, int32_t d, double* scratch,
                                    double* out, pmc_stream_t stream) {
  PMC_REQUIRE(x && scratch && out && n > 0 && d >= 1, "pmc_weighted_colsums: bad arguments");
  const int nc = geo_chunks(n, GEO_ROWS);
  geo_colsum_kernel<<<nc, 256, 0, as_stream(stream)>>>(x, w, center, n, d, scratch);
  PMC_LAUNCH_CHECK();
  geo_reduce_kernel<<<(d + 3 + 127) / 128, 128, 0, as_stream(stream)>>>(scratch, nc, d + 3, d + 2, out);
  PMC_LAUNCH_CHECK();
  return 0;
}

/* C [d, d] = sum_i w_i (x_i - center)(x_i - center)^T (w == NULL: w_i = 1) */
extern "C" int pmc_weighted_scatter(const double* x, const double* w, const double* center, int64_t n, int32_t d, double* scratch,
                                    double* C, pmc_stream_t stream) {
  PMC_REQUIRE(x && center && scratch && C && n > 0 && d >= 1, "pmc_weighted_scatter: bad arguments");
  const int nt = syrk_nt(d), n_tiles = nt * (nt + 1) / 2, nc = geo_chunks(n, SYRK_ROWS);
  geo_syrk_kernel<<<dim3(n_tiles, nc), 256, 0, as_stream(stream)>>>(x, w, center, n, d, nt, scratch);
  PMC_LAUNCH_CHECK();
  const int total = n_tiles * SYRK_TILE * SYRK_TILE;
  geo_syrk_reduce_kernel<<<(total + 255) / 256, 256, 0, as_stream(stream)>>>(scratch, nc, n_tiles, nt, d, C);
  PMC_LAUNCH_CHECK();
  return 0;
}

/* delta [n] = (x_i - center)^T P (x_i - center) */
extern "C" int pmc_mahalanobis(const double* x, const double* center, const double* P, int64_t n, int32_t d, double* delta,
                               pmc_stream_t stream) {
  PMC_REQUIRE(x && center && P && delta && n > 0 && d >= 1, "pmc_mahalanobis: bad arguments");
  const size_t smem = (size_t)64 * (d + 1) * sizeof(double);
  PMC_REQUIRE(smem <= 200 * 1024, "pmc_mahalanobis: dimension too large for the shared-memory row tile");
  PMC_TRY(cudaFuncSetAttribute(geo_mahalanobis_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  geo_mahalanobis_kernel<<<(unsigned)((n + 63) / 64), 256, smem, as_stream(stream)>>>(x, center, P, n, d, delta);
  PMC_LAUNCH_CHECK();
  return 0;
}

/* out2 = (sum_i log w_i, sum_i w_i) with w_i = (nu + dim) / (nu + delta_i); w_out (may be NULL) receives the weights */
extern "C" int pmc_student_weights(const double* delta, int64_t n, double nu, double dim, double* w_out, double* scratch, double* out2,
                                   pmc_stream_t stream) {
  PMC_REQUIRE(delta && scratch && out2 && n > 0, "pmc_student_weights: bad arguments");
  const int nc = geo_chunks(n, GEO_ROWS);
  geo_tweights_kernel<<<nc, 256, 0, as_stream(stream)>>>(delta, n, nu, dim, w_out, scratch);
  PMC_LAUNCH_CHECK();
  geo_reduce_kernel<<<1, 128, 0, as_stream(stream)>>>(scratch, nc, 2, -1, out2);
  PMC_LAUNCH_CHECK();
  return 0;
}

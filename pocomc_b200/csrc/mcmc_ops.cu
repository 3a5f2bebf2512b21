// Vectorised MCMC step kernels: proposal, reparameterisation, Metropolis accept/update,
// scalar adaptation, device RNG.  Reference: pocomc/mcmc.py (all four kernels) and
// pocomc/scaler.py.  SMC state is f64 like the reference's numpy arrays; these kernels are
// HBM-bound: one warp per particle row, lanes stride the D columns so every
// global access is a contiguous row segment.
#include "common.cuh"
#include <string.h>
#include <stdlib.h>

namespace pmc {

constexpr int ROWS_PER_BLOCK = 256;  // accept kernel: fixed row->block map => deterministic partials

// ------------------------------------------------------------------------------------------
// proposals
// ------------------------------------------------------------------------------------------
// DT > 0: the dimension as a compile-time constant (the loops unroll completely and the index arithmetic folds: the
// kernel is instruction-issue bound, ncu r2ax); DT = 0: any d.  Same arithmetic either way.
template <typename PosT, int DT>
__global__ void __launch_bounds__(128)
tpcn_propose_kernel(const PosT* __restrict__ pos, const double* __restrict__ ctl,
                    const double* __restrict__ inv_t, const double* __restrict__ chol_t, double nu,
                    const double* __restrict__ g, const double* __restrict__ z,
                    double* __restrict__ prop64, float* __restrict__ prop32,
                    double* __restrict__ m_cur, double* __restrict__ m_prop, long long n, int d_rt) {
  const int d = DT > 0 ? DT : d_rt;
  pdl_enter();
  extern __shared__ double sh[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* diff = sh + (size_t)warp * 3 * d;
  double* zz = diff + d;
  double* dp = zz + d;
  const double sigma = ctl[PMC_CTL_SIGMA];
  const double* mu = ctl + PMC_CTL_MU;
  const double keep = sqrt(1.0 - sigma * sigma);  // (1 - sigma**2)**0.5   mcmc.py:85
  const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < n; row += warps) {
    for (int j = lane; j < d; j += 32) {
      diff[j] = (double)pos[row * d + j] - mu[j];
      zz[j] = z[row * d + j];
    }
    __syncwarp();
    // both mat-vecs that only need the current state run in one j-loop (two independent fma chains, same
    // left-to-right order per chain as before): Sigma^-1 (theta - mu) and L z
    double part = 0.0;
    for (int i = lane; i < d; i += 32) {
      double v = 0.0, lz = 0.0;
#pragma unroll 8
      for (int j = 0; j < d; ++j) {
        v = fma(inv_t[(size_t)j * d + i], diff[j], v);
        lz = fma(chol_t[(size_t)j * d + i], zz[j], lz);
      }
      part = fma(diff[i], v, part);
      dp[i] = lz;                                          // this lane's own slot: read back below by the same lane
    }
    const double m = warp_sum(part);
    const double s = 1.0 / (g[row] * (2.0 / (nu + m)));   // 1 / gamma(a, scale = 2/(nu+m))   mcmc.py:80
    const double amp = sigma * sqrt(s);
    for (int i = lane; i < d; i += 32) {
      const double lz = dp[i];
      const double pr = (mu[i] + keep * diff[i]) + amp * lz;
      prop64[row * d + i] = pr;
      if (prop32) prop32[row * d + i] = (float)pr;
      dp[i] = pr - mu[i];
    }
    __syncwarp();
    part = 0.0;
    for (int i = lane; i < d; i += 32) {
      double v = 0.0;
#pragma unroll 8
      for (int j = 0; j < d; ++j) v = fma(inv_t[(size_t)j * d + i], dp[j], v);
      part = fma(dp[i], v, part);
    }
    const double mp = warp_sum(part);
    if (lane == 0) { m_cur[row] = m; m_prop[row] = mp; }
    __syncwarp();
  }
}

// The same proposal for wide problems (D >= 64): a block owns 32 rows, so that every element of the two D x D matrices is
// read four times per 32 rows instead of once per row (at D = 200 the warp-per-row kernel pulls 960 KB through L1/L2 per
// particle: 120 GB per step at 125 000 particles).  Thread (rg, c): 8 rows x two columns per 128-column pass; every output
// keeps the row kernel's arithmetic -- an ascending-j fma chain per element, the lane-strided partial sums and the xor
// tree of the Mahalanobis reductions -- so both kernels return bit-identical results.
constexpr int PROP_ROWS = 32;
template <typename PosT>
__global__ void __launch_bounds__(256)
tpcn_propose_tiled_kernel(const PosT* __restrict__ pos, const double* __restrict__ ctl,
                          const double* __restrict__ inv_t, const double* __restrict__ chol_t, double nu,
                          const double* __restrict__ g, const double* __restrict__ z,
                          double* __restrict__ prop64, float* __restrict__ prop32,
                          double* __restrict__ m_cur, double* __restrict__ m_prop, long long n, int d) {
  extern __shared__ double sh[];
  pdl_enter();
  const int ld = d + 1;
  double* diff = sh;                          // [32][ld]  theta - mu
  double* zz = diff + PROP_ROWS * ld;         // [32][ld]  z, later the proposal's offset from mu
  double* vv = zz + PROP_ROWS * ld;           // [32][ld]  Sigma^-1 (.)
  double* lz = vv + PROP_ROWS * ld;           // [32][ld]  L z
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rg = threadIdx.x >> 6, c = threadIdx.x & 63;        // 4 row groups of 8 rows, 64 columns per pass
  const double sigma = ctl[PMC_CTL_SIGMA];
  const double* mu = ctl + PMC_CTL_MU;
  const double keep = sqrt(1.0 - sigma * sigma);
  for (long long r0 = (long long)blockIdx.x * PROP_ROWS; r0 < n; r0 += (long long)gridDim.x * PROP_ROWS) {
    for (int e = threadIdx.x; e < PROP_ROWS * d; e += blockDim.x) {
      const int r = e / d, j = e - r * d;
      const bool ok = r0 + r < n;
      diff[r * ld + j] = ok ? (double)pos[(r0 + r) * d + j] - mu[j] : 0.0;
      zz[r * ld + j] = ok ? z[(r0 + r) * d + j] : 0.0;
    }
    __syncthreads();
    for (int i = c; i < d; i += 128) {                        // columns i and i + 64: 8 rows x 2 columns per thread
      const int i2 = i + 64;
      const bool two = i2 < d;
      double v[8], l[8], v2[8], l2[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) { v[q] = 0.0; l[q] = 0.0; v2[q] = 0.0; l2[q] = 0.0; }
      for (int j = 0; j < d; ++j) {
        const double a = inv_t[(size_t)j * d + i], b = chol_t[(size_t)j * d + i];
        const double a2 = two ? inv_t[(size_t)j * d + i2] : 0.0, b2 = two ? chol_t[(size_t)j * d + i2] : 0.0;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const double dq = diff[(rg * 8 + q) * ld + j], zq = zz[(rg * 8 + q) * ld + j];
          v[q] = fma(a, dq, v[q]);
          l[q] = fma(b, zq, l[q]);
          v2[q] = fma(a2, dq, v2[q]);
          l2[q] = fma(b2, zq, l2[q]);
        }
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        vv[(rg * 8 + q) * ld + i] = v[q]; lz[(rg * 8 + q) * ld + i] = l[q];
        if (two) { vv[(rg * 8 + q) * ld + i2] = v2[q]; lz[(rg * 8 + q) * ld + i2] = l2[q]; }
      }
    }
    __syncthreads();
    // m = (theta - mu)^T Sigma^-1 (theta - mu): one warp per row, lanes stride the columns, xor tree (as the row kernel)
    for (int r = warp; r < PROP_ROWS; r += 8) {
      double part = 0.0;
      for (int i = lane; i < d; i += 32) part = fma(diff[r * ld + i], vv[r * ld + i], part);
      const double m = warp_sum(part);
      const long long row = r0 + r;
      if (row < n) {
        const double s = 1.0 / (g[row] * (2.0 / (nu + m)));
        const double amp = sigma * sqrt(s);
        if (lane == 0) m_cur[row] = m;
        for (int i = lane; i < d; i += 32) {
          const double pr = (mu[i] + keep * diff[r * ld + i]) + amp * lz[r * ld + i];
          prop64[row * d + i] = pr;
          if (prop32) prop32[row * d + i] = (float)pr;
          zz[r * ld + i] = pr - mu[i];
        }
      } else {
        for (int i = lane; i < d; i += 32) zz[r * ld + i] = 0.0;
      }
    }
    __syncthreads();
    for (int i = c; i < d; i += 128) {
      const int i2 = i + 64;
      const bool two = i2 < d;
      double v[8], v2[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) { v[q] = 0.0; v2[q] = 0.0; }
      for (int j = 0; j < d; ++j) {
        const double a = inv_t[(size_t)j * d + i], a2 = two ? inv_t[(size_t)j * d + i2] : 0.0;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const double dq = zz[(rg * 8 + q) * ld + j];
          v[q] = fma(a, dq, v[q]);
          v2[q] = fma(a2, dq, v2[q]);
        }
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        vv[(rg * 8 + q) * ld + i] = v[q];
        if (two) vv[(rg * 8 + q) * ld + i2] = v2[q];
      }
    }
    __syncthreads();
    for (int r = warp; r < PROP_ROWS; r += 8) {
      double part = 0.0;
      for (int i = lane; i < d; i += 32) part = fma(zz[r * ld + i], vv[r * ld + i], part);
      const double mp = warp_sum(part);
      if (lane == 0 && r0 + r < n) m_prop[r0 + r] = mp;
    }
    __syncthreads();
  }
}

template <typename PosT>
__global__ void __launch_bounds__(128)
rwm_propose_kernel(const PosT* __restrict__ pos, const double* __restrict__ ctl,
                   const double* __restrict__ chol_t, const double* __restrict__ z,
                   double* __restrict__ prop64, float* __restrict__ prop32, long long n, int d) {
  extern __shared__ double sh[];
  pdl_enter();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* zz = sh + (size_t)warp * d;
  const double sigma = ctl[PMC_CTL_SIGMA];
  const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < n; row += warps) {
    for (int j = lane; j < d; j += 32) zz[j] = z[row * d + j];
    __syncwarp();
    for (int i = lane; i < d; i += 32) {
      double lz = 0.0;
      for (int j = 0; j < d; ++j) lz = fma(chol_t[(size_t)j * d + i], zz[j], lz);
      const double pr = (double)pos[row * d + i] + sigma * lz;   // mcmc.py:253
      prop64[row * d + i] = pr;
      if (prop32) prop32[row * d + i] = (float)pr;
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------
// reparameterisation (scaler.py)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void scaler_inv_dim(int kind, int logit, double v, double lo, double hi,
                                               double& x, double& J) {
  if (kind == 0) { x = v; J = 0.0; }
  else if (kind == 1) { x = exp(v) + lo; J = v; }                    // scaler.py:329-346
  else if (kind == 2) { x = hi - exp(v); J = v; }                    // scaler.py:362-378
  else {
    const double span = hi - lo;
    double p;
    if (logit) {                                                     // scaler.py:418-421
      p = exp(-logaddexp(0.0, -v));
      J = log(span) + log(p) + log(1.0 - p);
    } else {                                                         // scaler.py:422-425
      p = (erf(v / 1.4142135623730951) + 1.0) / 2.0;
      J = log(span) + (-(v * v) / 2.0) - 0.91893853320467267;
    }
    x = p * span + lo;
  }
}

__device__ __forceinline__ double scaler_fwd_dim(int kind, int logit, double x, double lo, double hi) {
  if (kind == 0) return x;
  if (kind == 1) return log(x - lo);
  if (kind == 2) return log(hi - x);
  const double p = (x - lo) / (hi - lo);
  return logit ? log(p / (1.0 - p)) : 1.4142135623730951 * erfinv(2.0 * p - 1.0);
}

__device__ __forceinline__ double wrap_bc(double x, int bc, double lo, double hi) {
  if (!isfinite(x)) return x;  // the reference's while-loops would never terminate here
  int guard = 0;
  if (bc & 1) {                // periodic, scaler.py:109-131
    while (x > hi && guard++ < 100000) x = lo + x - hi;
    while (x < lo && guard++ < 100000) x = hi + x - lo;
  }
  if (bc & 2) {                // reflective, scaler.py:133-157
    while (x > hi && guard++ < 100000) x = hi - x + hi;
    while (x < lo && guard++ < 100000) x = lo + lo - x;
  }
  return x;
}

// log-density of one factor of a product prior: kind 0 = norm(loc, scale), 1 = uniform(loc, loc + scale)
// (scipy.stats logpdf formulas; prior.py:36-44 sums the factors)
__device__ __forceinline__ double prior_term(int kind, double loc, double scale, double v) {
  if (kind == 0) {
    const double t = (v - loc) / scale;
    return -0.5 * t * t - 0.91893853320467267 - log(scale);
  }
  return (v >= loc && v <= loc + scale) ? -log(scale) : -INFINITY;
}

// PRIOR: the product prior of the proposed x' (logprior_kernel below) is evaluated in the same pass, on the x' values
// this kernel has just produced -- one launch less per MCMC step, identical numbers.
template <typename UT, bool PRIOR>
__global__ void __launch_bounds__(256)
scaler_inverse_kernel(const UT* __restrict__ u_in, pmc_scaler sc, double* __restrict__ u_out,
                      double* __restrict__ x_out, double* __restrict__ logdetj,
                      uint8_t* __restrict__ finite, const int* __restrict__ pkind, const double* __restrict__ ploc,
                      const double* __restrict__ pscale, double* __restrict__ logp, long long n, int d) {
  pdl_enter();
  const int lane = threadIdx.x & 31;
  const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < n; row += warps) {
    double jsum = 0.0, ps = 0.0;
    int ok = 1;
    for (int j = lane; j < d; j += 32) {
      double uu = (double)u_in[row * d + j];
      const int kind = sc.kind[j];
      const double lo = sc.low[j], hi = sc.high[j];
      double v = sc.scale ? (sc.mu[j] + sc.sigma[j] * uu) : uu;
      double xx, J;
      scaler_inv_dim(kind, sc.logit, v, lo, hi, xx, J);
      if (sc.bc) {  // mcmc.py:94-97: wrap x, re-forward to u, re-inverse
        xx = wrap_bc(xx, sc.bc[j], lo, hi);
        v = scaler_fwd_dim(kind, sc.logit, xx, lo, hi);
        uu = sc.scale ? (v - sc.mu[j]) / sc.sigma[j] : v;
        v = sc.scale ? (sc.mu[j] + sc.sigma[j] * uu) : uu;
        scaler_inv_dim(kind, sc.logit, v, lo, hi, xx, J);
      }
      u_out[row * d + j] = uu;
      x_out[row * d + j] = xx;
      jsum += J;
      ok &= isfinite(xx) ? 1 : 0;
      if (PRIOR) ps += prior_term(pkind[j], ploc[j], pscale[j], xx);
    }
    jsum = warp_sum(jsum);
    ok = warp_and(ok);
    if (PRIOR) ps = warp_sum(ps);
    if (lane == 0) {
      const double ld = (sc.scale ? sc.log_sigma_sum : 0.0) + jsum;
      logdetj[row] = ld;
      int fin = ok && isfinite(ld);
      if (PRIOR) {                       // logprior_kernel's rule: -inf on rows already out, rows with a non-finite prior go out
        logp[row] = fin ? ps : -INFINITY;
        if (fin && !isfinite(ps)) fin = 0;
      }
      finite[row] = (uint8_t)fin;
    }
  }
}

__global__ void __launch_bounds__(256)
scaler_forward_kernel(const double* __restrict__ x, pmc_scaler sc, double* __restrict__ u, long long total, int d) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int j = (int)(i % d);
    const double v = scaler_fwd_dim(sc.kind[j], sc.logit, x[i], sc.low[j], sc.high[j]);
    u[i] = sc.scale ? (v - sc.mu[j]) / sc.sigma[j] : v;
  }
}


__global__ void __launch_bounds__(256)
apply_bc_kernel(double* __restrict__ x, const int* __restrict__ bc, const double* __restrict__ low,
                const double* __restrict__ high, long long total, int d) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int j = (int)(i % d);
    if (bc[j]) x[i] = wrap_bc(x[i], bc[j], low[j], high[j]);
  }
}

// ------------------------------------------------------------------------------------------
// Metropolis accept + masked update + block partial sums
// ------------------------------------------------------------------------------------------
constexpr int FIN_TILE_FLOATS = 8192, FIN_TILE_COLS = 256;
__device__ __forceinline__ void finalize_body(int kind, double* __restrict__ ctl, const double* __restrict__ partials, int n_blocks,
                                              const float* __restrict__ pos32, int mean_mode, int n_steps, int n_max, long long n, int d,
                                              double* tot, float* tile);

// FUSED: the launch also runs the step's scalar adaptation (finalize_body) in the block that finishes last, and does
// nothing at all once the controller's stop flag is set -- so a host that queues several steps without reading the
// controller back cannot run past the reference's stopping point (mcmc.py:170-180).
// MODE 2 (particle-sharded runs, one process per GPU): the last block pushes this rank's block partials straight into
// every peer's exchange buffer over NVLink (plain stores into peer memory opened through CUDA IPC), publishes an epoch
// flag, waits for the flags of all peers and then adapts the controller from the rank-ordered partials of ALL ranks --
// accept, exchange and adaptation are one launch and no collective library call sits between two MCMC steps.
struct CommDev {
  double* buf[PMC_COMM_MAX_RANKS];                  // exchange buffer of every rank (own entry: local pointer)
  unsigned long long* flag[PMC_COMM_MAX_RANKS];     // epoch flags of every rank: flag[p][src] = last epoch src published to p
  unsigned long long* epoch;                        // this rank's epoch counter (device memory)
  int* error;                                       // set to 1 when a peer did not show up in time
  long long capacity;                               // doubles per parity half of a buffer
  int block_off[PMC_COMM_MAX_RANKS + 1];            // first global block of every rank
  int rank, world;
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

template <int MODE>
__global__ void __launch_bounds__(256)
mh_accept_kernel(int kind, double beta, double nu, float* __restrict__ pos32, double* __restrict__ u,
                 double* __restrict__ x, double* __restrict__ logdetj, double* __restrict__ logl,
                 double* __restrict__ logp, float* __restrict__ ldjf, const double* __restrict__ prop64,
                 const double* __restrict__ u_p, const double* __restrict__ x_p,
                 const double* __restrict__ logdetj_p, const double* __restrict__ logl_p,
                 const double* __restrict__ logp_p, const float* __restrict__ ldjf_p,
                 const double* __restrict__ m_cur, const double* __restrict__ m_prop,
                 const double* __restrict__ r, const uint8_t* __restrict__ finite,
                 double* __restrict__ alpha_out, double* __restrict__ partials, long long n, int d,
                 double* __restrict__ ctl, unsigned int* __restrict__ ticket, int mean_mode, int n_steps, int n_max,
                 const CommDev* __restrict__ comm, long long n_global) {
  constexpr bool FUSED = MODE != 0;
  extern __shared__ double sh[];  // [8 warps][d] theta sums + [8][4] scalars
  __shared__ __align__(16) float fin_tile[FUSED ? FIN_TILE_FLOATS + FIN_TILE_COLS : 1];   // finalize: staging of the block partials / of theta
  __shared__ unsigned int is_last;
  __shared__ unsigned long long epoch_sh;
  pdl_enter();
  if (FUSED && ctl[PMC_CTL_STOP] != 0.0) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* th = sh + (size_t)warp * d;
  double* sc = sh + (size_t)8 * d + warp * 4;
  const bool flow = (kind == PMC_KIND_TPCN_FLOW || kind == PMC_KIND_RWM_FLOW);
  const bool tp = (kind == PMC_KIND_TPCN_FLOW || kind == PMC_KIND_TPCN);
  const bool want_theta = (kind == PMC_KIND_TPCN_FLOW);
  const long long base = (long long)blockIdx.x * ROWS_PER_BLOCK + warp * 32;
  const long long row = base + lane;
  const bool valid = row < n;
  double alpha = 0.0, track = 0.0;
  int acc = 0, fin = 0;
  if (valid) {
    // mcmc.py:130-134 (left-to-right f64 sum; f32 flow log-dets promoted)
    double t = logl_p[row] * beta - logl[row] * beta + logp_p[row] - logp[row] + logdetj_p[row] - logdetj[row];
    if (flow) t = t + (double)ldjf_p[row] - (double)ldjf[row];
    if (tp) {
      const double A = -(d + nu) / 2 * log(1 + m_prop[row] / nu);
      const double B = -(d + nu) / 2 * log(1 + m_cur[row] / nu);
      t = t - A + B;
    }
    alpha = fmin(1.0, exp(t));
    if (isnan(alpha) || isnan(t)) alpha = 0.0;
    acc = r[row] < alpha;
    fin = finite ? finite[row] : 0;
    if (alpha_out) alpha_out[row] = alpha;
    double nl = logl[row], np_ = logp[row], nj = logdetj[row];
    if (acc) {
      nl = logl_p[row]; np_ = logp_p[row]; nj = logdetj_p[row];
      logl[row] = nl; logp[row] = np_; logdetj[row] = nj;
      if (flow) ldjf[row] = ldjf_p[row];
    }
    track = tp ? (nl + np_) : (nl + np_ + nj);   // mcmc.py:170 vs :327
  }
  const unsigned ballot = __ballot_sync(FULL, acc);
  const double s_alpha = warp_sum(alpha), s_track = warp_sum(track);
  const int s_fin = __popc(__ballot_sync(FULL, fin));
  // masked row copy + per-column theta sums.  A lane owns column j and walks the warp's 32 rows in ascending order
  // (the order fixes the f64 sum bit for bit); rows go in batches of 16 whose loads are all issued before the first
  // use -- one dependent round trip per batch instead of per row (the loop was 32 serial L2 latencies, 30 us at
  // 10 000 particles).
  for (int j = lane; j < d; j += 32) {
    double tsum = 0.0;
#pragma unroll 1
    for (int r0 = 0; r0 < 32; r0 += 16) {
      float pv[16];
      double uu[16], xx[16];
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const long long rw = base + r0 + q;
        const bool a = (ballot >> (r0 + q)) & 1u;         // accepted rows are valid rows
        const long long o = rw * d + j;
        pv[q] = 0.f; uu[q] = 0.0; xx[q] = 0.0;
        if (a) {
          uu[q] = u_p[o];
          xx[q] = x_p[o];
          if (pos32) pv[q] = (float)prop64[o];            // theta[mask] = theta_prime[mask] rounds to f32, mcmc.py:141
        } else if (want_theta && rw < n) {
          pv[q] = pos32[o];
        }
      }
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const long long rw = base + r0 + q;
        const long long o = rw * d + j;
        if ((ballot >> (r0 + q)) & 1u) {
          u[o] = uu[q];
          x[o] = xx[q];
          if (pos32) pos32[o] = pv[q];
        }
        if (want_theta && rw < n) tsum += (double)pv[q];
      }
    }
    th[j] = tsum;
  }
  if (lane == 0) { sc[0] = s_alpha; sc[1] = s_track; sc[2] = (double)__popc(ballot); sc[3] = (double)s_fin; }
  __syncthreads();
  double* out = partials + (size_t)blockIdx.x * (d + 4);
  for (int j = threadIdx.x; j < d + 4; j += blockDim.x) {
    double s = 0.0;
    if (j < 4) { for (int w = 0; w < 8; ++w) s += sh[(size_t)8 * d + w * 4 + j]; }
    else if (want_theta) { for (int w = 0; w < 8; ++w) s += sh[(size_t)w * d + (j - 4)]; }
    out[j] = s;
  }
  if (FUSED) {
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    if (MODE == 1) {
      finalize_body(kind, ctl, partials, (int)gridDim.x, pos32, mean_mode, n_steps, n_max, n, d, sh, fin_tile);
    } else {
      const int world = comm->world, rank = comm->rank, w = d + 4;
      if (threadIdx.x == 0) epoch_sh = *comm->epoch + 1ull;
      __syncthreads();
      const unsigned long long e = epoch_sh;
      const size_t half = (size_t)(e & 1ull) * (size_t)comm->capacity;           // double buffer by epoch parity
      const long long mine = (long long)gridDim.x * w;
      for (int p = 0; p < world; ++p) {
        double* dst = comm->buf[p] + half + (size_t)comm->block_off[rank] * w;
        for (long long j = threadIdx.x; j < mine; j += blockDim.x) dst[j] = partials[j];
      }
      __threadfence_system();
      __syncthreads();
      if ((int)threadIdx.x < world) {
        st_release_sys(comm->flag[threadIdx.x] + rank, e);
        const unsigned long long* f = comm->flag[rank] + threadIdx.x;
        const unsigned long long t0 = global_ns();
        while (ld_acquire_sys(f) < e) {
          if (global_ns() - t0 > 20000000000ull) { *comm->error = 1; break; }   // 20 s: a peer is gone; the host raises
        }
      }
      __syncthreads();
      finalize_body(kind, ctl, comm->buf[rank] + half, comm->block_off[world], nullptr, 0, n_steps, n_max, n_global, d, sh, fin_tile);
      if (threadIdx.x == 0) { *comm->epoch = e; if (*comm->error) ctl[PMC_CTL_STOP] = 1.0; }
    }
    if (threadIdx.x == 0) *ticket = 0u;
  }
}

// sigma / mu adaptation and the plateau rule from the block partials (mcmc.py:152-180); one 256-thread block.
// tot: [d + 6] doubles of shared memory, tile: [FIN_TILE_FLOATS + FIN_TILE_COLS] floats of shared memory (8-byte aligned).
__device__ __forceinline__ void finalize_body(int kind, double* __restrict__ ctl, const double* __restrict__ partials, int n_blocks,
                                              const float* __restrict__ pos32, int mean_mode, int n_steps, int n_max, long long n, int d,
                                              double* tot, float* tile) {
  constexpr int TILE_FLOATS = FIN_TILE_FLOATS, TILE_COLS = FIN_TILE_COLS;
  const bool tpf = (kind == PMC_KIND_TPCN_FLOW);
  const bool seq_mean = tpf && mean_mode != 0;
  const int w = d + 4;
  // Column sums of the block partials, added in block order (fixed: the result does not depend on the launch).  The
  // partials of a chunk of blocks are staged through shared memory by ALL threads with coalesced, independent loads
  // (one round trip to L2 per chunk); the thread that owns a column then adds its entries in block order.  Only the
  // d + 4 column owners used to be busy, each with one dependent round trip per 8 blocks.
  double* stage = reinterpret_cast<double*>(tile);
  constexpr int STAGE_DOUBLES = (FIN_TILE_FLOATS + FIN_TILE_COLS) / 2;
  const int cb = STAGE_DOUBLES / w;                            // blocks per chunk (0: a row of partials does not fit)
  for (int j = threadIdx.x; j < w; j += blockDim.x) tot[j] = 0.0;
  // the two pow() of the adaptation that do not depend on the sums are evaluated meanwhile by otherwise idle lanes
  double* pre = tot + w;                                       // [0] = 2.38 / sqrt(d), [1] = (step + 1)^-0.75 weight
  if (threadIdx.x == 64) pre[0] = 2.38 / pow((double)d, 0.5);
  if (threadIdx.x == 96) pre[1] = 1.0 / pow(ctl[PMC_CTL_STEP] + 1.0 + 1.0, 0.75);
  if (cb >= 1) {
    for (int b0 = 0; b0 < n_blocks; b0 += cb) {
      const int nb = min(cb, n_blocks - b0);
      __syncthreads();
      const double* src = partials + (size_t)b0 * w;
      for (int e = threadIdx.x; e < nb * w; e += blockDim.x) stage[e] = src[e];
      __syncthreads();
      for (int j = threadIdx.x; j < w; j += blockDim.x) {
        if ((j >= 4 && seq_mean) || !(j < 4 || tpf)) continue;
        double s = tot[j];
        for (int b = 0; b < nb; ++b) s += stage[b * w + j];
        tot[j] = s;
      }
    }
  } else {
    for (int j = threadIdx.x; j < w; j += blockDim.x) {
      if ((j >= 4 && seq_mean) || !(j < 4 || tpf)) continue;
      double s = 0.0;
      for (int b0 = 0; b0 < n_blocks; b0 += 8) {              // 8 loads in flight, added in block order
        double v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = (b0 + q < n_blocks) ? partials[(size_t)(b0 + q) * w + j] : 0.0;
#pragma unroll
        for (int q = 0; q < 8; ++q) if (b0 + q < n_blocks) s += v[q];
      }
      tot[j] = s;
    }
  }
  __syncthreads();                                             // the staging area is reused by the sequential mean below
  if (tpf && !seq_mean)
    for (int j = 4 + threadIdx.x; j < w; j += blockDim.x) tot[j] /= (double)n;
  if (seq_mean) {
    // np.mean(theta, axis=0) on the f32 theta array (mcmc.py:156): per column a SEQUENTIAL f32 accumulation over the
    // rows in ascending order and an f32 divide.  The order is kept exactly; what changes is how the rows reach the
    // adder: whole [rows x cols] tiles are staged through shared memory with coalesced, independent loads (a thread
    // walking its column straight from global memory paid one L2 round trip per few rows: milliseconds per step at
    // 10 000 particles).
    for (int c0 = 0; c0 < d; c0 += TILE_COLS) {
      const int cols = min(TILE_COLS, d - c0), ld = cols + 1;
      const int rows_per_tile = min(256, TILE_FLOATS / cols);
      float a = 0.0f;
      for (long long r0 = 0; r0 < n; r0 += rows_per_tile) {
        const int rows = (int)min((long long)rows_per_tile, n - r0);
        for (int e = threadIdx.x; e < rows * cols; e += blockDim.x) {
          const int rr = e / cols, cc = e - rr * cols;
          tile[rr * ld + cc] = pos32[(size_t)(r0 + rr) * d + c0 + cc];
        }
        __syncthreads();
        if ((int)threadIdx.x < cols)
          for (int rr = 0; rr < rows; ++rr) a += tile[rr * ld + threadIdx.x];
        __syncthreads();
      }
      if ((int)threadIdx.x < cols) tot[4 + c0 + threadIdx.x] = (double)(a / (float)n);
    }
  }
  __syncthreads();
  const double step = ctl[PMC_CTL_STEP] + 1.0;   // i (1-based) of the step just finished
  if (tpf) {
    for (int j = threadIdx.x; j < d; j += blockDim.x) {
      const double mu = ctl[PMC_CTL_MU + j];
      ctl[PMC_CTL_MU + j] = mu + 1.0 / (step + 1.0) * (tot[4 + j] - mu);   // mcmc.py:156
    }
  }
  if (threadIdx.x == 0) {
    const double mean_alpha = tot[0] / (double)n;
    const double track = tot[1] / (double)n;
    double sigma = ctl[PMC_CTL_SIGMA];
    const double cap = pre[0];                                 // 2.38 / pow(d, 0.5)
    if (kind == PMC_KIND_TPCN_FLOW || kind == PMC_KIND_TPCN)
      sigma = fabs(fmin(sigma + pre[1] * (mean_alpha - 0.234), fmin(cap, 0.99)));  // mcmc.py:152, pre[1] = 1 / pow(step + 1, 0.75)
    else if (kind == PMC_KIND_RWM_FLOW)
      sigma = sigma + 1.0 / (step + 1.0) * (mean_alpha - 0.234);                                         // mcmc.py:314
    else
      sigma = fabs(sigma + 1.0 / (step + 1.0) * (mean_alpha - 0.234));                                   // mcmc.py:627
    double best = ctl[PMC_CTL_BEST], cnt = ctl[PMC_CTL_CNT], stop = 0.0;
    if (track > best) { cnt = 0.0; best = track; }
    else {
      cnt += 1.0;
      double ratio = cap / sigma;
      if (kind == PMC_KIND_RWM_FLOW) ratio = fmin(1.0, ratio);                                           // mcmc.py:333
      if (cnt >= n_steps * pow(ratio, 2.0)) stop = 1.0;
    }
    if (step >= (double)n_max) stop = 1.0;
    ctl[PMC_CTL_SIGMA] = sigma;
    ctl[PMC_CTL_STEP] = step;
    ctl[PMC_CTL_BEST] = best;
    ctl[PMC_CTL_CNT] = cnt;
    ctl[PMC_CTL_STOP] = stop;
    ctl[PMC_CTL_ACCEPT] = mean_alpha;
    ctl[PMC_CTL_CALLS] += tot[3];
    ctl[PMC_CTL_TRACK] = track;
    ctl[PMC_CTL_NACC] = tot[2];
  }
}

__global__ void __launch_bounds__(256)
mcmc_finalize_kernel(int kind, double* __restrict__ ctl, const double* __restrict__ partials, int n_blocks,
                     const float* __restrict__ pos32, int mean_mode, int n_steps, int n_max, long long n, int d) {
  extern __shared__ double tot[];  // [d + 6]
  __shared__ __align__(16) float tile[FIN_TILE_FLOATS + FIN_TILE_COLS];   // staging: block partials, then [rows][cols + 1] of theta for the sequential mean
  pdl_enter();
  finalize_body(kind, ctl, partials, n_blocks, pos32, mean_mode, n_steps, n_max, n, d, tot, tile);
}

// ------------------------------------------------------------------------------------------
// Philox4x32-10 counter RNG (throughput mode; parity mode uploads numpy's legacy stream instead)
// ------------------------------------------------------------------------------------------
struct Philox {
  uint32_t k0, k1;
  __device__ __forceinline__ uint4 operator()(uint4 c) const {
    uint32_t a = k0, b = k1;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
      const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
      const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
      c = make_uint4(hi1 ^ c.y ^ a, lo1, hi0 ^ c.w ^ b, lo0);
      a += 0x9E3779B9u; b += 0xBB67AE85u;
    }
    return c;
  }
};
__device__ __forceinline__ double u53(uint32_t hi, uint32_t lo) {  // (0,1)
  const unsigned long long v = (((unsigned long long)hi << 32) | lo) >> 11;
  return ((double)v + 0.5) * (1.0 / 9007199254740992.0);
}
__device__ __forceinline__ void box_muller(uint4 c, double& n0, double& n1) {
  const double u1 = u53(c.x, c.y), u2 = u53(c.z, c.w);
  const double rad = sqrt(-2.0 * log(u1));
  double sn, cs;
  sincospi(2.0 * u2, &sn, &cs);
  n0 = rad * cs; n1 = rad * sn;
}

// Thread map: the first n * ceil(d / 2) threads draw one pair of normals each, the next n threads one uniform + one gamma
// each -- every warp runs ONE of the two code paths (a per-row interleave made every warp pay for both: the gamma path
// is as long as the normal path and 1 lane in d/2 + 1 used it).  Counters are keyed by (particle id, step, slot), so the
// numbers do not depend on the map.
template <typename IndexT>
__global__ void __launch_bounds__(256)
rng_fill_kernel(uint64_t seed, uint64_t step, const double* __restrict__ ctl, long long offset, double shape, double* __restrict__ g,
                double* __restrict__ z, double* __restrict__ r, long long n, int d) {
  pdl_enter();
  if (ctl) step = (uint64_t)ctl[PMC_CTL_STEP] + 1;      // the step about to run, read where the previous step left it
  const Philox ph{(uint32_t)seed, (uint32_t)(seed >> 32)};
  const IndexT half = (IndexT)((d + 1) / 2);
  const IndexT n_normal = (IndexT)n * half, total = n_normal + (IndexT)n;
  const IndexT stride = (IndexT)gridDim.x * blockDim.x;
  for (IndexT i = (IndexT)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    if (i < n_normal) {  // two normals
      const IndexT row = i / half;
      const int slot = (int)(i - row * half);
      const unsigned long long pid = (unsigned long long)((long long)row + offset);
      double a, b;
      box_muller(ph(make_uint4((uint32_t)pid, (uint32_t)(pid >> 32), (uint32_t)step, (uint32_t)(step >> 32) ^ ((uint32_t)slot << 8))), a, b);
      z[(long long)row * d + 2 * slot] = a;
      if (2 * slot + 1 < d) z[(long long)row * d + 2 * slot + 1] = b;
    } else {            // uniform + gamma (Marsaglia-Tsang, shape >= 1)
      const long long row = (long long)(i - n_normal);
      const unsigned long long pid = (unsigned long long)(row + offset);
      const uint32_t tag = 0x80000000u;
      uint4 c = ph(make_uint4((uint32_t)pid, (uint32_t)(pid >> 32), (uint32_t)step, ((uint32_t)(step >> 32)) ^ tag));
      if (r) r[row] = u53(c.x, c.y);
      if (g && shape > 0.0) {
        const double a = shape < 1.0 ? shape + 1.0 : shape;
        const double dd = a - 1.0 / 3.0, cc = 1.0 / sqrt(9.0 * dd);
        double out = dd;
        for (uint32_t att = 1; att < 4096; ++att) {
          uint4 c1 = ph(make_uint4((uint32_t)pid, (uint32_t)(pid >> 32), (uint32_t)step, ((uint32_t)(step >> 32)) ^ tag ^ (att << 8)));
          uint4 c2 = ph(make_uint4((uint32_t)pid, (uint32_t)(pid >> 32), (uint32_t)step, ((uint32_t)(step >> 32)) ^ tag ^ (att << 8) ^ 1u));
          double xn, unused;
          box_muller(c1, xn, unused);
          const double uu = u53(c2.x, c2.y);
          double v = 1.0 + cc * xn;
          if (v <= 0.0) continue;
          v = v * v * v;
          if (log(uu) < 0.5 * xn * xn + dd - dd * v + dd * log(v)) { out = dd * v; break; }
        }
        if (shape < 1.0) out *= pow(u53(c.z, c.w), 1.0 / shape);
        g[row] = out;
      }
    }
  }
}

static inline int launch_rng_fill(uint64_t seed, uint64_t step, const double* ctl, long long offset, double shape, double* g, double* z,
                                  double* r, long long n, int d, cudaStream_t stream) {
  const long long total = n * ((d + 1) / 2 + 1);
  const int blocks = grid_for(total, 256, 8);
  if (total < (1ll << 31) - (long long)blocks * 256)
    PMC_TRY(launch_chain(rng_fill_kernel<uint32_t>, dim3(blocks), dim3(256), 0, stream, seed, step, ctl, offset, shape, g, z, r, n, d));
  else
    PMC_TRY(launch_chain(rng_fill_kernel<long long>, dim3(blocks), dim3(256), 0, stream, seed, step, ctl, offset, shape, g, z, r, n, d));
  return 0;
}

// ------------------------------------------------------------------------------------------
// synthetic likelihoods / product priors on device (bench + opt-in fast path)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
loglike_kernel(int which, const double* __restrict__ x, const uint8_t* __restrict__ finite,
               const double* __restrict__ mat_t, double p0, double p1, double* __restrict__ logl,
               long long n, int d) {
  extern __shared__ double sh[];
  pdl_enter();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* xr = sh + (size_t)warp * d;
  const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < n; row += warps) {
    if (finite && !finite[row]) { if (lane == 0) logl[row] = -INFINITY; continue; }
    for (int j = lane; j < d; j += 32) xr[j] = x[row * d + j];
    __syncwarp();
    double part = 0.0, res;
    if (which == PMC_LIKE_GAUSS) {          // -0.5 x^T P x + p0
      for (int i = lane; i < d; i += 32) {
        double v = 0.0;
        for (int j = 0; j < d; ++j) v = fma(mat_t[(size_t)j * d + i], xr[j], v);
        part = fma(xr[i], v, part);
      }
      res = -0.5 * warp_sum(part) + p0;
    } else if (which == PMC_LIKE_ROSENBROCK) {   // README.md:53-55
      for (int i = 2 * lane; i + 1 < d; i += 64) {
        const double a = xr[i] * xr[i] - xr[i + 1], b = xr[i] - 1.0;
        part += 10.0 * a * a + b * b;
      }
      res = -warp_sum(part);
    } else if (which == PMC_LIKE_MIXTURE) {      // logaddexp(N(x;+c,s^2 I), N(x;-c,s^2 I)) - log 2
      double qa = 0.0, qb = 0.0;
      for (int i = lane; i < d; i += 32) { qa += (xr[i] - p0) * (xr[i] - p0); qb += (xr[i] + p0) * (xr[i] + p0); }
      qa = warp_sum(qa); qb = warp_sum(qb);
      const double norm = -0.5 * d * log(2.0 * 3.14159265358979323846 * p1 * p1);
      res = logaddexp(norm - 0.5 * qa / (p1 * p1), norm - 0.5 * qb / (p1 * p1)) - 0.69314718055994530942;
    } else {                                     // Neal funnel: x0 ~ N(0,p0^2), x_i ~ N(0, e^{x0})
      const double x0 = xr[0];
      for (int i = lane + 1; i < d; i += 32) part += xr[i] * xr[i];
      part = warp_sum(part);
      res = -0.5 * x0 * x0 / (p0 * p0) - 0.5 * log(2.0 * 3.14159265358979323846 * p0 * p0)
            - 0.5 * part * exp(-x0) - 0.5 * (d - 1) * (log(2.0 * 3.14159265358979323846) + x0);
    }
    if (lane == 0) logl[row] = res;
    __syncwarp();
  }
}

__global__ void __launch_bounds__(256)
logprior_kernel(const double* __restrict__ x, uint8_t* __restrict__ finite, const int* __restrict__ kind,
                const double* __restrict__ loc, const double* __restrict__ scale, double* __restrict__ logp,
                long long n, int d) {
  pdl_enter();
  const int lane = threadIdx.x & 31;
  const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < n; row += warps) {
    if (finite && !finite[row]) { if (lane == 0) logp[row] = -INFINITY; continue; }
    double s = 0.0;
    for (int j = lane; j < d; j += 32) {
      s += prior_term(kind[j], loc[j], scale[j], x[row * d + j]);
    }
    s = warp_sum(s);
    if (lane == 0) {
      logp[row] = s;
      if (finite && !isfinite(s)) finite[row] = 0;
    }
  }
}

}  // namespace pmc

using namespace pmc;

extern "C" int pmc_tpcn_propose(int32_t pos_is_f32, const void* pos, const double* ctl, const double* inv_cov_t,
                                const double* chol_t, double nu, const double* g, const double* z, double* prop64,
                                float* prop32, double* m_cur, double* m_prop, int64_t n, int32_t d,
                                pmc_stream_t stream) {
  PMC_REQUIRE(pos && ctl && inv_cov_t && chol_t && g && z && prop64 && m_cur && m_prop, "pmc_tpcn_propose: null pointer");
  PMC_REQUIRE(d >= 1 && d <= 1024, "pmc_tpcn_propose: unsupported dimension");
  if (n == 0) return 0;
  const size_t tiled_smem = (size_t)4 * PROP_ROWS * (d + 1) * sizeof(double);
  const char* force_rows = getenv("PMC_TPCN_ROW_KERNEL");          // tests: compare the two kernels bit for bit
  static const int tiled_min_d = [] { const char* e = getenv("PMC_TPCN_TILED_MIN_D"); return e && e[0] ? atoi(e) : 64; }();
  if (d >= tiled_min_d && tiled_smem <= 220 * 1024 && !(force_rows && force_rows[0] == '1')) {
    const int tblocks = grid_for(n, PROP_ROWS, 1);
    if (pos_is_f32) {
      PMC_TRY(cudaFuncSetAttribute(tpcn_propose_tiled_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tiled_smem));
      PMC_TRY(launch_chain(tpcn_propose_tiled_kernel<float>, dim3(tblocks), dim3(256), tiled_smem, as_stream(stream), (const float*)pos, ctl, inv_cov_t, chol_t, nu, g, z, prop64, prop32, m_cur, m_prop, n, d));
    } else {
      PMC_TRY(cudaFuncSetAttribute(tpcn_propose_tiled_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tiled_smem));
      PMC_TRY(launch_chain(tpcn_propose_tiled_kernel<double>, dim3(tblocks), dim3(256), tiled_smem, as_stream(stream), (const double*)pos, ctl, inv_cov_t, chol_t, nu, g, z, prop64, prop32, m_cur, m_prop, n, d));
    }
    PMC_LAUNCH_CHECK();
    return 0;
  }
  const size_t smem = (size_t)4 * 3 * d * sizeof(double);
  const int blocks = grid_for(n, 4, 16);
  if (pos_is_f32 && d == 32)
    PMC_TRY(launch_chain(tpcn_propose_kernel<float, 32>, dim3(blocks), dim3(128), smem, as_stream(stream), (const float*)pos, ctl, inv_cov_t, chol_t, nu, g, z, prop64, prop32, m_cur, m_prop, n, d));
  else if (pos_is_f32)
    PMC_TRY(launch_chain(tpcn_propose_kernel<float, 0>, dim3(blocks), dim3(128), smem, as_stream(stream), (const float*)pos, ctl, inv_cov_t, chol_t, nu, g, z, prop64, prop32, m_cur, m_prop, n, d));
  else
    PMC_TRY(launch_chain(tpcn_propose_kernel<double, 0>, dim3(blocks), dim3(128), smem, as_stream(stream), (const double*)pos, ctl, inv_cov_t, chol_t, nu, g, z, prop64, prop32, m_cur, m_prop, n, d));
  PMC_LAUNCH_CHECK();
  return 0;
}

extern "C" int pmc_rwm_propose(int32_t pos_is_f32, const void* pos, const double* ctl, const double* chol_t,
                               const double* z, double* prop64, float* prop32, int64_t n, int32_t d,
                               pmc_stream_t stream) {
  PMC_REQUIRE(pos && ctl && chol_t && z && prop64, "pmc_rwm_propose: null pointer");
  PMC_REQUIRE(d >= 1 && d <= 1024, "pmc_rwm_propose: unsupported dimension");
  if (n == 0) return 0;
  const size_t smem = (size_t)4 * d * sizeof(double);
  const int blocks = grid_for(n, 4, 16);
  if (pos_is_f32)
    PMC_TRY(launch_chain(rwm_propose_kernel<float>, dim3(blocks), dim3(128), smem, as_stream(stream), (const float*)pos, ctl, chol_t, z, prop64, prop32, n, d));
  else
    PMC_TRY(launch_chain(rwm_propose_kernel<double>, dim3(blocks), dim3(128), smem, as_stream(stream), (const double*)pos, ctl, chol_t, z, prop64, prop32, n, d));
  PMC_LAUNCH_CHECK();
  return 0;
}

extern "C" int pmc_scaler_inverse(int32_t u_is_f32, const void* u_in, const pmc_scaler* sc, double* u_out, double* x,
                                  double* logdetj, uint8_t* finite, int64_t n, int32_t d, pmc_stream_t stream) {
  if (n == 0) return 0;   // empty batch: torch hands out null data pointers for 0-element tensors
  PMC_REQUIRE(u_in && sc && u_out && x && logdetj && finite, "pmc_scaler_inverse: null pointer");
  PMC_REQUIRE(sc->kind && sc->low && sc->high && (!sc->scale || (sc->mu && sc->sigma)), "pmc_scaler_inverse: incomplete scaler");
  const int blocks = grid_for(n, 8, 8);
  if (u_is_f32)
    PMC_TRY(launch_chain(scaler_inverse_kernel<float, false>, dim3(blocks), dim3(256), 0, as_stream(stream), (const float*)u_in, *sc, u_out, x, logdetj, finite, nullptr, nullptr, nullptr, nullptr, n, d));
  else
    PMC_TRY(launch_chain(scaler_inverse_kernel<double, false>, dim3(blocks), dim3(256), 0, as_stream(stream), (const double*)u_in, *sc, u_out, x, logdetj, finite, nullptr, nullptr, nullptr, nullptr, n, d));
  PMC_LAUNCH_CHECK();
  return 0;
}

extern "C" int pmc_scaler_inverse_prior(int32_t u_is_f32, const void* u_in, const pmc_scaler* sc, const int32_t* prior_kind,
                                        const double* prior_loc, const double* prior_scale, double* u_out, double* x,
                                        double* logdetj, uint8_t* finite, double* logp, int64_t n, int32_t d,
                                        pmc_stream_t stream) {
  if (n == 0) return 0;
  PMC_REQUIRE(u_in && sc && u_out && x && logdetj && finite && logp && prior_kind && prior_loc && prior_scale,
              "pmc_scaler_inverse_prior: null pointer");
  PMC_REQUIRE(sc->kind && sc->low && sc->high && (!sc->scale || (sc->mu && sc->sigma)), "pmc_scaler_inverse_prior: incomplete scaler");
  const int blocks = grid_for(n, 8, 8);
  if (u_is_f32)
    PMC_TRY(launch_chain(scaler_inverse_kernel<float, true>, dim3(blocks), dim3(256), 0, as_stream(stream), (const float*)u_in, *sc, u_out, x, logdetj, finite, prior_kind, prior_loc, prior_scale, logp, n, d));
  else
    PMC_TRY(launch_chain(scaler_inverse_kernel<double, true>, dim3(blocks), dim3(256), 0, as_stream(stream), (const double*)u_in, *sc, u_out, x, logdetj, finite, prior_kind, prior_loc, prior_scale, logp, n, d));
  PMC_LAUNCH_CHECK();
  return 0;
}

extern "C" int pmc_scaler_forward(const double* x, const pmc_scaler* sc, double* u, int64_t n, int32_t d,
                                  pmc_stream_t stream) {
  if (n == 0) return 0;
  PMC_REQUIRE(x && sc && u, "pmc_scaler_forward: null pointer");
  const int blocks = grid_for(n * d, 256, 8);
  scaler_forward_kernel<<<blocks, 256, 0, as_stream(stream)>>>(x, *sc, u, n * (long long)d, d);
  PMC_LAUNCH_CHECK();
  return 0;
}


extern "C" int pmc_apply_bc(double* x, const int32_t* bc, const double* low, const double* high, int64_t n, int32_t d,
                            pmc_stream_t stream) {
  if (n == 0) return 0;
  PMC_REQUIRE(x && bc && low && high, "pmc_apply_bc: null pointer");
  apply_bc_kernel<<<grid_for(n * d, 256, 8), 256, 0, as_stream(stream)>>>(x, bc, low, high, n * (long long)d, d);
  PMC_LAUNCH_CHECK();
  return 0;
}

static inline int64_t mh_blocks(int64_t n) { return (n + ROWS_PER_BLOCK - 1) / ROWS_PER_BLOCK; }

extern "C" int64_t pmc_mh_partials_size(int64_t n, int32_t d) { return mh_blocks(n) * (int64_t)(d + 4); }

extern "C" int pmc_mh_accept_update(int32_t kind, double beta, double nu, float* pos32, double* u, double* x,
                                    double* logdetj, double* logl, double* logp, float* logdetj_flow,
                                    const double* prop64, const double* u_p, const double* x_p,
                                    const double* logdetj_p, const double* logl_p, const double* logp_p,
                                    const float* logdetj_flow_p, const double* m_cur, const double* m_prop,
                                    const double* r, const uint8_t* finite, double* alpha_out, double* partials,
                                    int64_t n, int32_t d, pmc_stream_t stream) {
  PMC_REQUIRE(kind >= 0 && kind <= 3, "pmc_mh_accept_update: bad kind");
  PMC_REQUIRE(u && x && logdetj && logl && logp && u_p && x_p && logdetj_p && logl_p && logp_p && r && partials,
              "pmc_mh_accept_update: null pointer");
  const bool flow = (kind == PMC_KIND_TPCN_FLOW || kind == PMC_KIND_RWM_FLOW);
  const bool tp = (kind == PMC_KIND_TPCN_FLOW || kind == PMC_KIND_TPCN);
  PMC_REQUIRE(!flow || (pos32 && prop64 && logdetj_flow && logdetj_flow_p), "pmc_mh_accept_update: flow kinds need theta + flow log-dets");
  PMC_REQUIRE(!tp || (m_cur && m_prop), "pmc_mh_accept_update: tpCN kinds need the Mahalanobis distances");
  if (n == 0) return 0;
  const size_t smem = ((size_t)8 * d + 32) * sizeof(double);
  PMC_TRY(launch_chain(mh_accept_kernel<0>, dim3((unsigned)mh_blocks(n)), dim3(256), smem, as_stream(stream), kind, beta, nu, flow ? pos32 : nullptr, u, x, logdetj, logl, logp, logdetj_flow, prop64, u_p, x_p, logdetj_p,
      logl_p, logp_p, logdetj_flow_p, m_cur, m_prop, r, finite, alpha_out, partials, n, d, nullptr, nullptr, 0, 0, 0, nullptr, 0));
  PMC_LAUNCH_CHECK();
  return 0;
}

extern "C" int pmc_mh_accept_finalize(int32_t kind, double beta, double nu, float* pos32, double* u, double* x,
                                      double* logdetj, double* logl, double* logp, float* logdetj_flow,
                                      const double* prop64, const double* u_p, const double* x_p,
                                      const double* logdetj_p, const double* logl_p, const double* logp_p,
                                      const float* logdetj_flow_p, const double* m_cur, const double* m_prop,
                                      const double* r, const uint8_t* finite, double* alpha_out, double* partials,
                                      double* ctl, uint32_t* ticket, int32_t mean_mode, int32_t n_steps, int32_t n_max,
                                      int64_t n, int32_t d, pmc_stream_t stream) {
  PMC_REQUIRE(kind >= 0 && kind <= 3, "pmc_mh_accept_finalize: bad kind");
  PMC_REQUIRE(u && x && logdetj && logl && logp && u_p && x_p && logdetj_p && logl_p && logp_p && r && partials && ctl && ticket,
              "pmc_mh_accept_finalize: null pointer");
  const bool flow = (kind == PMC_KIND_TPCN_FLOW || kind == PMC_KIND_RWM_FLOW);
  const bool tp = (kind == PMC_KIND_TPCN_FLOW || kind == PMC_KIND_TPCN);
  PMC_REQUIRE(!flow || (pos32 && prop64 && logdetj_flow && logdetj_flow_p), "pmc_mh_accept_finalize: flow kinds need theta + flow log-dets");
  PMC_REQUIRE(!tp || (m_cur && m_prop), "pmc_mh_accept_finalize: tpCN kinds need the Mahalanobis distances");
  PMC_REQUIRE(n > 0, "pmc_mh_accept_finalize: empty batch");
  const size_t smem = ((size_t)8 * d + 32) * sizeof(double);
  PMC_TRY(launch_chain(mh_accept_kernel<1>, dim3((unsigned)mh_blocks(n)), dim3(256), smem, as_stream(stream), kind, beta, nu, flow ? pos32 : nullptr, u, x, logdetj, logl, logp, logdetj_flow, prop64, u_p, x_p, logdetj_p,
      logl_p, logp_p, logdetj_flow_p, m_cur, m_prop, r, finite, alpha_out, partials, n, d, ctl, ticket, mean_mode, n_steps, n_max,
      nullptr, 0));
  PMC_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------
// peer-memory exchange context (one per process = per GPU)
// ------------------------------------------------------------------------------------------
struct pmc_comm {
  int rank, world, device;
  long long capacity;                  // doubles per parity half
  double* buf;                         // [2 * capacity] doubles, then [world] flags, epoch, error -- ONE cudaMalloc (one IPC handle)
  void* peer_base[PMC_COMM_MAX_RANKS]; // opened peer allocations (nullptr for self / not connected)
  CommDev host;                        // host image of the device descriptor
  CommDev* dev;
  bool connected;
};

static size_t comm_bytes(long long capacity, int world) { return (size_t)2 * capacity * sizeof(double) + ((size_t)world + 2) * sizeof(unsigned long long); }

extern "C" int pmc_comm_create(int32_t rank, int32_t world, int64_t capacity_doubles, void** comm_out, unsigned char* handle64) {
  PMC_REQUIRE(comm_out && handle64 && world >= 1 && world <= PMC_COMM_MAX_RANKS && rank >= 0 && rank < world && capacity_doubles > 0,
              "pmc_comm_create: bad arguments");
  static_assert(sizeof(cudaIpcMemHandle_t) == PMC_COMM_HANDLE_BYTES, "IPC handle size");
  pmc_comm* c = new pmc_comm();
  c->rank = rank; c->world = world; c->capacity = capacity_doubles; c->connected = false; c->dev = nullptr; c->buf = nullptr;
  for (int p = 0; p < PMC_COMM_MAX_RANKS; ++p) c->peer_base[p] = nullptr;
  PMC_TRY(cudaGetDevice(&c->device));
  const size_t bytes = comm_bytes(capacity_doubles, world);
  PMC_TRY(cudaMalloc(&c->buf, bytes));
  PMC_TRY(cudaMemset(c->buf, 0, bytes));
  PMC_TRY(cudaMalloc(&c->dev, sizeof(CommDev)));
  PMC_TRY(cudaDeviceSynchronize());
  cudaIpcMemHandle_t h;
  PMC_TRY(cudaIpcGetMemHandle(&h, c->buf));
  memcpy(handle64, &h, sizeof(h));
  *comm_out = c;
  return 0;
}

extern "C" int pmc_comm_connect(void* comm, const unsigned char* handles, const int32_t* block_off) {
  pmc_comm* c = static_cast<pmc_comm*>(comm);
  PMC_REQUIRE(c && handles && block_off && !c->connected, "pmc_comm_connect: bad arguments");
  const size_t flags_off = (size_t)2 * c->capacity * sizeof(double);
  for (int p = 0; p < c->world; ++p) {
    unsigned char* base;
    if (p == c->rank) base = reinterpret_cast<unsigned char*>(c->buf);
    else {
      cudaIpcMemHandle_t h;
      memcpy(&h, handles + (size_t)p * PMC_COMM_HANDLE_BYTES, sizeof(h));
      void* ptr = nullptr;
      PMC_TRY(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
      c->peer_base[p] = ptr;
      base = static_cast<unsigned char*>(ptr);
    }
    c->host.buf[p] = reinterpret_cast<double*>(base);
    c->host.flag[p] = reinterpret_cast<unsigned long long*>(base + flags_off);
  }
  unsigned char* mine = reinterpret_cast<unsigned char*>(c->buf);
  c->host.epoch = reinterpret_cast<unsigned long long*>(mine + flags_off) + c->world;
  c->host.error = reinterpret_cast<int*>(reinterpret_cast<unsigned long long*>(mine + flags_off) + c->world + 1);
  c->host.capacity = c->capacity;
  c->host.rank = c->rank; c->host.world = c->world;
  for (int p = 0; p <= c->world; ++p) c->host.block_off[p] = block_off[p];
  PMC_REQUIRE((long long)block_off[c->world] >= 0, "pmc_comm_connect: bad block offsets");
  PMC_TRY(cudaMemcpy(c->dev, &c->host, sizeof(CommDev), cudaMemcpyHostToDevice));
  c->connected = true;
  return 0;
}

extern "C" int pmc_comm_set_blocks(void* comm, const int32_t* block_off, pmc_stream_t stream) {
  pmc_comm* c = static_cast<pmc_comm*>(comm);
  PMC_REQUIRE(c && c->connected && block_off, "pmc_comm_set_blocks: not connected");
  for (int p = 0; p <= c->world; ++p) c->host.block_off[p] = block_off[p];
  PMC_TRY(cudaMemcpyAsync(c->dev, &c->host, sizeof(CommDev), cudaMemcpyHostToDevice, as_stream(stream)));
  PMC_TRY(cudaStreamSynchronize(as_stream(stream)));
  return 0;
}

extern "C" int pmc_comm_error(void* comm) {
  pmc_comm* c = static_cast<pmc_comm*>(comm);
  if (!c || !c->connected) return -1;
  int e = 0;
  if (cudaMemcpy(&e, c->host.error, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  return e;
}

extern "C" int pmc_comm_destroy(void* comm) {
  pmc_comm* c = static_cast<pmc_comm*>(comm);
  if (!c) return 0;
  cudaDeviceSynchronize();
  for (int p = 0; p < c->world; ++p) if (c->peer_base[p]) cudaIpcCloseMemHandle(c->peer_base[p]);
  if (c->dev) cudaFree(c->dev);
  if (c->buf) cudaFree(c->buf);
  delete c;
  return 0;
}

extern "C" int pmc_mh_accept_finalize_p2p(int32_t kind, double beta, double nu, float* pos32, double* u, double* x,
                                          double* logdetj, double* logl, double* logp, float* logdetj_flow,
                                          const double* prop64, const double* u_p, const double* x_p,
                                          const double* logdetj_p, const double* logl_p, const double* logp_p,
                                          const float* logdetj_flow_p, const double* m_cur, const double* m_prop,
                                          const double* r, const uint8_t* finite, double* alpha_out, double* partials,
                                          double* ctl, uint32_t* ticket, int32_t n_steps, int32_t n_max,
                                          int64_t n, int32_t d, void* comm, int64_t n_global, pmc_stream_t stream) {
  pmc_comm* c = static_cast<pmc_comm*>(comm);
  PMC_REQUIRE(kind >= 0 && kind <= 3, "pmc_mh_accept_finalize_p2p: bad kind");
  PMC_REQUIRE(u && x && logdetj && logl && logp && u_p && x_p && logdetj_p && logl_p && logp_p && r && partials && ctl && ticket,
              "pmc_mh_accept_finalize_p2p: null pointer");
  PMC_REQUIRE(c && c->connected, "pmc_mh_accept_finalize_p2p: exchange context not connected (pmc_comm_create / pmc_comm_connect)");
  const bool flow = (kind == PMC_KIND_TPCN_FLOW || kind == PMC_KIND_RWM_FLOW);
  const bool tp = (kind == PMC_KIND_TPCN_FLOW || kind == PMC_KIND_TPCN);
  PMC_REQUIRE(!flow || (pos32 && prop64 && logdetj_flow && logdetj_flow_p), "pmc_mh_accept_finalize_p2p: flow kinds need theta + flow log-dets");
  PMC_REQUIRE(!tp || (m_cur && m_prop), "pmc_mh_accept_finalize_p2p: tpCN kinds need the Mahalanobis distances");
  PMC_REQUIRE(n > 0 && n_global >= n, "pmc_mh_accept_finalize_p2p: empty batch");
  const int* bo = c->host.block_off;
  PMC_REQUIRE(bo[c->rank + 1] - bo[c->rank] == (int)mh_blocks(n), "pmc_mh_accept_finalize_p2p: block table does not match this rank's batch");
  PMC_REQUIRE((long long)bo[c->world] * (d + 4) <= c->capacity, "pmc_mh_accept_finalize_p2p: exchange buffer too small");
  const size_t smem = ((size_t)8 * d + 32) * sizeof(double);
  PMC_TRY(launch_chain(mh_accept_kernel<2>, dim3((unsigned)mh_blocks(n)), dim3(256), smem, as_stream(stream), kind, beta, nu, flow ? pos32 : nullptr, u, x, logdetj, logl, logp, logdetj_flow, prop64, u_p, x_p, logdetj_p,
      logl_p, logp_p, logdetj_flow_p, m_cur, m_prop, r, finite, alpha_out, partials, n, d, ctl, ticket, 0, n_steps, n_max,
      c->dev, n_global));
  PMC_LAUNCH_CHECK();
  return 0;
}

extern "C" int pmc_mcmc_finalize(int32_t kind, double* ctl, const double* partials, int64_t n_blocks,
                                 const float* pos32, int32_t mean_mode, int32_t n_steps, int32_t n_max, int64_t n,
                                 int32_t d, pmc_stream_t stream) {
  PMC_REQUIRE(ctl && partials && n > 0, "pmc_mcmc_finalize: bad arguments");
  PMC_REQUIRE(!(kind == PMC_KIND_TPCN_FLOW && mean_mode == 1) || pos32, "pmc_mcmc_finalize: mean_mode 1 needs theta");
  PMC_TRY(launch_chain(mcmc_finalize_kernel, dim3(1), dim3(256), (size_t)(d + 6) * sizeof(double), as_stream(stream), kind, ctl, partials, (int)(n_blocks > 0 ? n_blocks : mh_blocks(n)), pos32, mean_mode, n_steps, n_max, n, d));
  PMC_LAUNCH_CHECK();
  return 0;
}

extern "C" int pmc_rng_fill(uint64_t seed, uint64_t step, int64_t particle_offset, double gamma_shape, double* g,
                            double* z, double* r, int64_t n, int32_t d, pmc_stream_t stream) {
  PMC_REQUIRE(z && n >= 0 && d >= 1, "pmc_rng_fill: bad arguments");
  if (n == 0) return 0;
  if (launch_rng_fill(seed, step, nullptr, particle_offset, gamma_shape, g, z, r, n, d, as_stream(stream))) return 1;
  PMC_LAUNCH_CHECK();
  return 0;
}

extern "C" int pmc_rng_fill_ctl(uint64_t seed, const double* ctl, int64_t particle_offset, double gamma_shape, double* g,
                                double* z, double* r, int64_t n, int32_t d, pmc_stream_t stream) {
  PMC_REQUIRE(z && ctl && n >= 0 && d >= 1, "pmc_rng_fill_ctl: bad arguments");
  if (n == 0) return 0;
  if (launch_rng_fill(seed, 0, ctl, particle_offset, gamma_shape, g, z, r, n, d, as_stream(stream))) return 1;
  PMC_LAUNCH_CHECK();
  return 0;
}

extern "C" int pmc_loglike(int32_t which, const double* x, const uint8_t* finite, const double* mat_t, double p0,
                           double p1, double* logl, int64_t n, int32_t d, pmc_stream_t stream) {
  PMC_REQUIRE(x && logl && which >= 0 && which <= 3, "pmc_loglike: bad arguments");
  PMC_REQUIRE(which != PMC_LIKE_GAUSS || mat_t, "pmc_loglike: gauss needs the precision matrix");
  if (n == 0) return 0;
  const int blocks = grid_for(n, 4, 16);
  PMC_TRY(launch_chain(loglike_kernel, dim3(blocks), dim3(128), (size_t)4 * d * sizeof(double), as_stream(stream), which, x, finite, mat_t, p0, p1, logl, n, d));
  PMC_LAUNCH_CHECK();
  return 0;
}

extern "C" int pmc_logprior(const double* x, uint8_t* finite, const int32_t* kind, const double* loc,
                            const double* scale, double* logp, int64_t n, int32_t d, pmc_stream_t stream) {
  PMC_REQUIRE(x && kind && loc && scale && logp, "pmc_logprior: null pointer");
  if (n == 0) return 0;
  const int blocks = grid_for(n, 8, 8);
  PMC_TRY(launch_chain(logprior_kernel, dim3(blocks), dim3(256), 0, as_stream(stream), x, finite, kind, loc, scale, logp, n, d));
  PMC_LAUNCH_CHECK();
  return 0;
}

// Layer-wise training step of Flow.fit (pocomc/flow.py:301-319) for EVERY flow the package builds: zuko NSF (the reference's
// default preset, rational-quadratic spline heads, 23 parameters per feature) and MAF of any width (H = 512 / 1024 at the
// BASELINE configs 3-4), i.e. the shapes the fused kernels of flow_train.cu do not cover.
//
// One call = weighted negative log-likelihood of one mini-batch + its gradient with respect to the flat parameter blob,
// as a fixed sequence of launches on the caller's stream (captured into the optimiser step's CUDA graph by the host):
//   gather batch rows, mask the weights (zuko MaskedLinear: mask * weight) ->
//   per transform: 4 masked linear layers (fp32 GEMM, bias / residual / ReLU in the epilogue) -> univariate head
//   (affine or spline) -> base log-density, loss partials, d loss / d z ->
//   per transform backwards: head gradient (forward-mode duals through the spline: no hand-derived formula to get wrong),
//   4 input-gradient GEMMs (ReLU gate and residual pass-through in the epilogue), 4 weight-gradient GEMMs (mask in the
//   epilogue, written straight into the gradient blob), bias gradients.
// Everything is deterministic (no atomics); arithmetic is fp32 like the reference's autograd path.
#include "common.cuh"
#include <algorithm>

namespace pmc {

constexpr int LW_TM = 64, LW_TN = 64, LW_TK = 16;
enum { LW_EPI_FWD = 0, LW_EPI_BWD = 1, LW_EPI_WGRAD = 2 };

struct LwGemm {
  const float* A; const float* B; float* C;
  int M, N, K, lda, ldb, ldc;
  int epi;
  const float* bias;      // FWD: [N]
  const float* res;       // FWD / BWD: added before the activation / gate, [M, ldr]
  const float* gate;      // BWD: multiply by (gate > 0), [M, ldg];  WGRAD: multiply by the mask [M, ldg]
  int ldr, ldg, relu;
};

// C[M,N] = epilogue(op(A)[M,K] . op(B)[K,N]);  TA: A[m][k] stored at A[k*lda + m], else A[m*lda + k];
// TB: B[k][n] stored at B[n*ldb + k], else B[k*ldb + n]
template <bool TA, bool TB>
__device__ __forceinline__ void lw_gemm_body(const LwGemm& g, float (&As)[LW_TK][LW_TM + 4], float (&Bs)[LW_TK][LW_TN + 4]) {
  const int m0 = blockIdx.y * LW_TM, n0 = blockIdx.x * LW_TN;
  if (m0 >= g.M || n0 >= g.N) return;
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  float acc[4][4] = {};
  // 16-byte loads when every row of the operand starts on a 16-byte boundary and the tile is interior
  const bool va = (g.lda & 3) == 0 && ((size_t)g.A & 15) == 0, vb = (g.ldb & 3) == 0 && ((size_t)g.B & 15) == 0;
  // each thread moves four A and four B elements per k-tile; the contiguous direction of an operand is k (row-major
  // [rows][k]) or the tile's 64 rows / columns.  (ar, ak) / (br, bk): tile coordinates of this thread's first element
  const int ar = TA ? (threadIdx.x & 15) * 4 : threadIdx.x >> 2, ak = TA ? threadIdx.x >> 4 : (threadIdx.x & 3) * 4;
  const int br = TB ? threadIdx.x >> 2 : (threadIdx.x & 15) * 4, bk = TB ? (threadIdx.x & 3) * 4 : threadIdx.x >> 4;
  float ra[4], rb[4];
  auto fetch = [&](const int k0) {
    const bool full_k = k0 + LW_TK <= g.K;
    if (TA) {                                           // A[k*lda + m]
      if (va && full_k && m0 + LW_TM <= g.M) {
        const float4 v = *reinterpret_cast<const float4*>(g.A + (size_t)(k0 + ak) * g.lda + m0 + ar);
        ra[0] = v.x; ra[1] = v.y; ra[2] = v.z; ra[3] = v.w;
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) ra[i] = (m0 + ar + i < g.M && k0 + ak < g.K) ? g.A[(size_t)(k0 + ak) * g.lda + m0 + ar + i] : 0.f;
      }
    } else {                                            // A[m*lda + k]
      if (va && full_k && m0 + ar < g.M) {
        const float4 v = *reinterpret_cast<const float4*>(g.A + (size_t)(m0 + ar) * g.lda + k0 + ak);
        ra[0] = v.x; ra[1] = v.y; ra[2] = v.z; ra[3] = v.w;
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) ra[i] = (m0 + ar < g.M && k0 + ak + i < g.K) ? g.A[(size_t)(m0 + ar) * g.lda + k0 + ak + i] : 0.f;
      }
    }
    if (TB) {                                           // B[n*ldb + k]
      if (vb && full_k && n0 + br < g.N) {
        const float4 v = *reinterpret_cast<const float4*>(g.B + (size_t)(n0 + br) * g.ldb + k0 + bk);
        rb[0] = v.x; rb[1] = v.y; rb[2] = v.z; rb[3] = v.w;
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) rb[i] = (n0 + br < g.N && k0 + bk + i < g.K) ? g.B[(size_t)(n0 + br) * g.ldb + k0 + bk + i] : 0.f;
      }
    } else {                                            // B[k*ldb + n]
      if (vb && full_k && n0 + LW_TN <= g.N) {
        const float4 v = *reinterpret_cast<const float4*>(g.B + (size_t)(k0 + bk) * g.ldb + n0 + br);
        rb[0] = v.x; rb[1] = v.y; rb[2] = v.z; rb[3] = v.w;
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) rb[i] = (n0 + br + i < g.N && k0 + bk < g.K) ? g.B[(size_t)(k0 + bk) * g.ldb + n0 + br + i] : 0.f;
      }
    }
  };
  auto stash = [&]() {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (TA) As[ak][ar + i] = ra[i]; else As[ak + i][ar] = ra[i];
      if (TB) Bs[bk + i][br] = rb[i]; else Bs[bk][br + i] = rb[i];
    }
  };
  fetch(0);
  for (int k0 = 0; k0 < g.K; k0 += LW_TK) {
    stash();
    __syncthreads();
    if (k0 + LW_TK < g.K) fetch(k0 + LW_TK);          // the next tile's global loads fly while this tile is multiplied
#pragma unroll
    for (int k = 0; k < LW_TK; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = As[k][ty + 16 * i]; b[i] = Bs[k][tx + 16 * i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty + 16 * i;
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx + 16 * j;
      if (n >= g.N) continue;
      float v = acc[i][j];
      if (g.epi == LW_EPI_FWD) {
        v += g.bias[n];
        if (g.res) v += g.res[(size_t)m * g.ldr + n];
        if (g.relu) v = fmaxf(v, 0.f);
      } else if (g.epi == LW_EPI_BWD) {
        if (g.res) v += g.res[(size_t)m * g.ldr + n];
        if (g.gate) v = g.gate[(size_t)m * g.ldg + n] > 0.f ? v : 0.f;
      } else {
        v *= g.gate[(size_t)m * g.ldg + n];
      }
      g.C[(size_t)m * g.ldc + n] = v;
    }
  }
}
template <bool TA, bool TB>
__global__ void __launch_bounds__(256) lw_gemm_kernel(const LwGemm g) {
  __shared__ float As[LW_TK][LW_TM + 4], Bs[LW_TK][LW_TN + 4];
  lw_gemm_body<TA, TB>(g, As, Bs);
}

__global__ void lw_mask_kernel(const float* __restrict__ raw, const float* __restrict__ mask, float* __restrict__ out, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = raw[i] * mask[i];
}

// batch gather + per-row loss coefficient: xb[r] = x[idx[cursor][r]];  coef[r] = -1000 w m / sum(w m) (weighted, flow.py:307-310)
// or -m (flow.py:305); one block
__global__ void __launch_bounds__(256)
lw_gather_kernel(const float* __restrict__ x, const float* __restrict__ w, const long long* __restrict__ idx_all, const float* __restrict__ mask_all,
                 const long long* __restrict__ cursor, int B, int D, float* __restrict__ xb, float* __restrict__ coef, float* __restrict__ ladj) {
  __shared__ float red[256];
  const long long* idx = idx_all + (size_t)(*cursor) * B;
  const float* msk = mask_all + (size_t)(*cursor) * B;
  float s = 0.f;
  for (int r = threadIdx.x; r < B; r += 256) s += w ? w[idx[r]] * msk[r] : 0.f;
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) { if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o]; __syncthreads(); }
  const float tot = red[0];
  for (int r = threadIdx.x; r < B; r += 256) {
    coef[r] = w ? -1000.0f * (w[idx[r]] * msk[r]) / tot : -msk[r];
    ladj[r] = 0.f;
  }
  for (int e = threadIdx.x; e < B * D; e += 256) { const int r = e / D, j = e - r * D; xb[e] = x[(size_t)idx[r] * D + j]; }
}

// ---- univariate heads -------------------------------------------------------------------------------------------------
constexpr float LW_LOG_SLOPE = -6.90775527898213705205f;     // log(1e-3)

// forward-mode dual number over NV independent variables
template <int NV>
struct Dual {
  float v, d[NV];
  __device__ static Dual constant(float c) { Dual r; r.v = c; for (int i = 0; i < NV; ++i) r.d[i] = 0.f; return r; }
  __device__ static Dual variable(float c, int k) { Dual r = constant(c); r.d[k] = 1.f; return r; }
};
template <int NV> __device__ __forceinline__ Dual<NV> operator+(const Dual<NV>& a, const Dual<NV>& b) { Dual<NV> r; r.v = a.v + b.v; for (int i = 0; i < NV; ++i) r.d[i] = a.d[i] + b.d[i]; return r; }
template <int NV> __device__ __forceinline__ Dual<NV> operator-(const Dual<NV>& a, const Dual<NV>& b) { Dual<NV> r; r.v = a.v - b.v; for (int i = 0; i < NV; ++i) r.d[i] = a.d[i] - b.d[i]; return r; }
template <int NV> __device__ __forceinline__ Dual<NV> operator*(const Dual<NV>& a, const Dual<NV>& b) { Dual<NV> r; r.v = a.v * b.v; for (int i = 0; i < NV; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i]; return r; }
template <int NV> __device__ __forceinline__ Dual<NV> operator/(const Dual<NV>& a, const Dual<NV>& b) {
  Dual<NV> r; const float inv = 1.0f / b.v; r.v = a.v * inv;
  for (int i = 0; i < NV; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) * inv;
  return r;
}
template <int NV> __device__ __forceinline__ Dual<NV> dlog(const Dual<NV>& a) { Dual<NV> r; r.v = logf(a.v); const float inv = 1.0f / a.v; for (int i = 0; i < NV; ++i) r.d[i] = a.d[i] * inv; return r; }

__device__ __forceinline__ float softclip(float a, float c) { return a / (1.0f + fabsf(a / c)); }
__device__ __forceinline__ float softclip_grad(float a, float c) { const float t = 1.0f + fabsf(a / c); return 1.0f / (t * t); }

// knots of one feature from its 23 raw parameters (bins = 8, bound = 5; zuko MonotonicRQSTransform)
struct RqsKnots {
  float W[8], Hh[8], Dv[9], hx[9], hy[9];
};
__device__ __forceinline__ void rqs_knots(const float* __restrict__ p, RqsKnots& k) {
  float w[8], h[8], mw = -1e30f, mh = -1e30f;
#pragma unroll
  for (int i = 0; i < 8; ++i) { w[i] = softclip(p[i], 0.5f * LW_LOG_SLOPE); h[i] = softclip(p[8 + i], 0.5f * LW_LOG_SLOPE); mw = fmaxf(mw, w[i]); mh = fmaxf(mh, h[i]); }
  float sw = 0.f, sh = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) { k.W[i] = expf(w[i] - mw); k.Hh[i] = expf(h[i] - mh); sw += k.W[i]; sh += k.Hh[i]; }
  float cw = 0.f, ch = 0.f;
  k.hx[0] = -5.f; k.hy[0] = -5.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    k.W[i] /= sw; k.Hh[i] /= sh;
    cw += k.W[i]; ch += k.Hh[i];
    k.hx[i + 1] = 5.f * (2.f * cw - 1.f);
    k.hy[i + 1] = 5.f * (2.f * ch - 1.f);
  }
  k.Dv[0] = 1.f; k.Dv[8] = 1.f;
#pragma unroll
  for (int i = 0; i < 7; ++i) k.Dv[i + 1] = expf(softclip(p[16 + i], LW_LOG_SLOPE));
}
// torch.searchsorted(hx, x) - 1: the bin with hx[k] < x <= hx[k+1]; outside [0, 8) the transform is the identity
__device__ __forceinline__ int rqs_bin(const RqsKnots& k, float x) {
  int c = 0;
#pragma unroll
  for (int i = 0; i < 9; ++i) c += (k.hx[i] < x) ? 1 : 0;
  return c - 1;
}

// y, ladj of the spline as functions of (x, x0, x1, y0, y1, d0, d1), any scalar type
template <typename S>
__device__ __forceinline__ void rqs_eval(const S& x, const S& x0, const S& x1, const S& y0, const S& y1, const S& d0, const S& d1, const S& one,
                                         const S& two, S& y, S& ladj_num, S& den) {
  const S dx = x1 - x0, dy = y1 - y0;
  const S s = dy / dx;
  const S z = (x - x0) / dx;
  const S omz = one - z;
  den = s + (d0 + d1 - two * s) * z * omz;
  y = y0 + dy * (s * z * z + d0 * z * omz) / den;
  ladj_num = s * s * (two * s * z * omz + d0 * omz * omz + d1 * z * z);
}

// forward: y [B,D], ladj[r] += sum_j ladj;  one thread per (row, feature) for the map, rows reduced by one thread each
template <bool RQS>
__global__ void __launch_bounds__(128)
lw_head_fwd_kernel(const float* __restrict__ v, const float* __restrict__ phi, int B, int D, int total, float* __restrict__ y, float* __restrict__ lrow) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= B * D) return;
  const float x = v[e];
  const float* p = phi + (size_t)e * total;
  float out, l;
  if (!RQS) {
    const float ls = softclip(p[1], LW_LOG_SLOPE);
    out = fmaf(x, expf(ls), p[0]);
    l = ls;
  } else {
    RqsKnots k;
    rqs_knots(p, k);
    const int b = rqs_bin(k, x);
    if (b < 0 || b >= 8) { out = x; l = 0.f; }
    else {
      float yy, num, den;
      rqs_eval<float>(x, k.hx[b], k.hx[b + 1], k.hy[b], k.hy[b + 1], k.Dv[b], k.Dv[b + 1], 1.f, 2.f, yy, num, den);
      out = yy;
      l = logf(num / (den * den));
    }
  }
  y[e] = out;
  lrow[e] = l;
}
// ladj[r] += sum_j lrow[r][j] in feature order
__global__ void lw_rowsum_kernel(const float* __restrict__ lrow, int B, int D, float* __restrict__ ladj) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= B) return;
  float s = 0.f;
  for (int j = 0; j < D; ++j) s += lrow[(size_t)r * D + j];
  ladj[r] += s;
}

// backward of the head: gy [B,D] = d loss / d y, coef[r] = d loss / d ladj[r]  ->  dphi [B, D*total], gv [B,D] (direct part)
template <bool RQS>
__global__ void __launch_bounds__(128)
lw_head_bwd_kernel(const float* __restrict__ v, const float* __restrict__ phi, const float* __restrict__ gy, const float* __restrict__ coef,
                   int B, int D, int total, float* __restrict__ dphi, float* __restrict__ gv) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= B * D) return;
  const int r = e / D;
  const float x = v[e], g_y = gy[e], g_l = coef[r];
  const float* p = phi + (size_t)e * total;
  float* dp = dphi + (size_t)e * total;
  if (!RQS) {
    const float ls = softclip(p[1], LW_LOG_SLOPE), es = expf(ls);
    dp[0] = g_y;
    dp[1] = (g_y * x * es + g_l) * softclip_grad(p[1], LW_LOG_SLOPE);
    gv[e] = g_y * es;
    return;
  }
  RqsKnots k;
  rqs_knots(p, k);
  const int b = rqs_bin(k, x);
#pragma unroll
  for (int i = 0; i < 23; ++i) dp[i] = 0.f;
  if (b < 0 || b >= 8) { gv[e] = g_y; return; }
  using D7 = Dual<7>;
  D7 yy, num, den;
  rqs_eval<D7>(D7::variable(x, 0), D7::variable(k.hx[b], 1), D7::variable(k.hx[b + 1], 2), D7::variable(k.hy[b], 3), D7::variable(k.hy[b + 1], 4),
               D7::variable(k.Dv[b], 5), D7::variable(k.Dv[b + 1], 6), D7::constant(1.f), D7::constant(2.f), yy, num, den);
  const D7 ladj = dlog(num) - D7::constant(2.f) * dlog(den);
  float gq[7];
#pragma unroll
  for (int i = 0; i < 7; ++i) gq[i] = g_y * yy.d[i] + g_l * ladj.d[i];
  gv[e] = gq[0];
  // knots -> softmax probabilities: hx_k = 5 (2 sum_{m<k} W_m - 1)  =>  d hx_k / d W_m = 10 [m < k]
  float gW[8], gH[8], sW = 0.f, sH = 0.f;
#pragma unroll
  for (int m = 0; m < 8; ++m) {
    gW[m] = 10.f * ((m < b ? gq[1] : 0.f) + (m < b + 1 ? gq[2] : 0.f));
    gH[m] = 10.f * ((m < b ? gq[3] : 0.f) + (m < b + 1 ? gq[4] : 0.f));
    sW += k.W[m] * gW[m]; sH += k.Hh[m] * gH[m];
  }
#pragma unroll
  for (int m = 0; m < 8; ++m) {
    dp[m] = k.W[m] * (gW[m] - sW) * softclip_grad(p[m], 0.5f * LW_LOG_SLOPE);
    dp[8 + m] = k.Hh[m] * (gH[m] - sH) * softclip_grad(p[8 + m], 0.5f * LW_LOG_SLOPE);
  }
  // derivatives: Dv[i] = exp(softclip(p[16 + i - 1])) for i = 1..7 (the two boundary ones are the constant 1)
  if (b >= 1) dp[16 + b - 1] += gq[5] * k.Dv[b] * softclip_grad(p[16 + b - 1], LW_LOG_SLOPE);
  if (b + 1 <= 7) dp[16 + b] += gq[6] * k.Dv[b + 1] * softclip_grad(p[16 + b], LW_LOG_SLOPE);
}

// base density + loss: lp = sum_j (-z^2/2 - log(2 pi)/2) + ladj; partial[blk] = sum_r coef[r] lp[r] (f64); gz = coef (-z)
__global__ void __launch_bounds__(128)
lw_loss_kernel(const float* __restrict__ z, const float* __restrict__ ladj, const float* __restrict__ coef, int B, int D, float* __restrict__ gz,
               double* __restrict__ partial) {
  __shared__ double red[128];
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  double contrib = 0.0;
  if (r < B) {
    float s = 0.f;
    const float c = coef[r];
    for (int j = 0; j < D; ++j) {
      const float zz = z[(size_t)r * D + j];
      s += -0.5f * zz * zz - 0.91893853320467274178f;
      gz[(size_t)r * D + j] = -c * zz;
    }
    contrib = (double)c * (double)(s + ladj[r]);
  }
  red[threadIdx.x] = contrib;
  __syncthreads();
  for (int o = 64; o > 0; o >>= 1) { if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o]; __syncthreads(); }
  if (threadIdx.x == 0) partial[blockIdx.x] = red[0];
}

// the four weight-gradient GEMMs of one transform in one launch (blockIdx.z = layer), and their four bias gradients
// (out[n] = sum_m Y[m][n]: 32 columns x 8 row groups per block, four independent partial sums per thread, fixed order)
struct LwGemm4 { LwGemm g[4]; };
__global__ void __launch_bounds__(256) lw_wgrad4_kernel(const LwGemm4 q) {
  __shared__ float As[LW_TK][LW_TM + 4], Bs[LW_TK][LW_TN + 4];
  lw_gemm_body<true, false>(q.g[blockIdx.z], As, Bs);
}
struct LwColsum4 { const float* Y[4]; float* out[4]; int N[4]; int M; };
__global__ void __launch_bounds__(256) lw_colsum4_kernel(const LwColsum4 q) {
  __shared__ float red[8][33];
  const int l = blockIdx.y;
  const float* __restrict__ Y = q.Y[l];
  const int N = q.N[l], M = q.M, ld = N;
  const int c = threadIdx.x & 31, rg = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + c;
  if (blockIdx.x * 32 >= N) return;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if (n < N) {
    int m = rg;
    for (; m + 24 < M; m += 32) {
      s0 += Y[(size_t)m * ld + n]; s1 += Y[(size_t)(m + 8) * ld + n]; s2 += Y[(size_t)(m + 16) * ld + n]; s3 += Y[(size_t)(m + 24) * ld + n];
    }
    for (; m < M; m += 8) s0 += Y[(size_t)m * ld + n];
  }
  red[rg][c] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (rg == 0 && n < N) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += red[i][c];
    q.out[l][n] = s;
  }
}

static void lw_launch_gemm(bool ta, bool tb, const LwGemm& g, cudaStream_t st) {
  dim3 grid((g.N + LW_TN - 1) / LW_TN, (g.M + LW_TM - 1) / LW_TM);
  if (!ta && tb) lw_gemm_kernel<false, true><<<grid, 256, 0, st>>>(g);
  else if (!ta && !tb) lw_gemm_kernel<false, false><<<grid, 256, 0, st>>>(g);
  else lw_gemm_kernel<true, false><<<grid, 256, 0, st>>>(g);
}

}  // namespace pmc

using namespace pmc;

/* floats of scratch for a (padded) batch of B rows: masked weights | xb | coef | ladj | lrow | per transform (v, h0, h1, h2, phi) |
 * z | g (D) x 2 | dphi | dh x 3 */
extern "C" int64_t pmc_flow_train_lw_scratch_size(int32_t D, int32_t H, int32_t T, int32_t total, int64_t numel, int64_t B) {
  const int64_t P = (int64_t)D * total;
  return numel + B * D + 2 * B + B * D + (int64_t)T * B * (D + 3 * (int64_t)H + P) + B * D + 2 * B * D + B * P + 3 * B * H + 64;
}
extern "C" int32_t pmc_flow_train_lw_partials(int64_t B) { return (int32_t)((B + 127) / 128); }

/* One mini-batch of Flow.fit on the layer-wise kernels.  raw / mask / grad: flat parameter blob (module order: per transform
 * W0 [H,D], b0, W1 [H,H], b1, W2, b2, W3 [D*total,H], b3), its MADE mask laid out alike (1 for biases), its gradient.
 * x [*, D] f32 training matrix, w (NULL: unweighted), idx_all / mask_all [*, B] batch tables, cursor: device int64 selecting the
 * batch.  kind 0 = affine heads (total = 2), 1 = spline heads (total = 23, bins = 8).  partials: pmc_flow_train_lw_partials(B)
 * doubles whose sum is the batch loss.  train = 0: loss only.                                                              */
extern "C" int pmc_flow_train_step_lw(const float* raw, const float* mask, int32_t D, int32_t H, int32_t T, int32_t kind, int64_t numel,
                                      const float* x, const float* w, const int64_t* idx_all, const float* mask_all, const int64_t* cursor,
                                      int64_t B64, float* scratch, double* partials, float* grad, int32_t train, pmc_stream_t stream) {
  PMC_REQUIRE(raw && mask && x && idx_all && mask_all && cursor && scratch && partials && (grad || !train), "pmc_flow_train_step_lw: null pointer");
  PMC_REQUIRE(D >= 2 && H >= 1 && T >= 1 && (kind == 0 || kind == 1) && B64 > 0 && B64 <= 65536, "pmc_flow_train_step_lw: bad shape");
  const int B = (int)B64, total = kind == 0 ? 2 : 23, P = D * total;
  const int64_t tstride = (int64_t)H * D + H + 2 * ((int64_t)H * H + H) + (int64_t)P * H + P;
  PMC_REQUIRE(numel == tstride * T, "pmc_flow_train_step_lw: blob size does not match (D, H, T, kind)");
  cudaStream_t st = as_stream(stream);
  float* f = scratch;
  auto take = [&](size_t n) { float* p_ = f; f += (n + 3) & ~(size_t)3; return p_; };     // every buffer starts on a 16-byte boundary
  float* Wm = take((size_t)numel);
  float* coef = take(B);
  float* ladj = take(B);
  float* lrow = take((size_t)B * D);
  float* acts = take((size_t)T * B * (D + 3 * (size_t)H + P));
  float* z = take((size_t)B * D);
  float* gA = take((size_t)B * D);
  float* gB = take((size_t)B * D);
  float* dphi = take((size_t)B * P);
  float* dh[3] = {take((size_t)B * H), take((size_t)B * H), take((size_t)B * H)};     // d loss / d pre-activation of layers 2, 1, 0
  const size_t per_t = (size_t)B * (D + 3 * (size_t)H + P);
  auto act = [&](int t, int which) -> float* {      // 0: v (input), 1..3: h0..h2, 4: phi
    float* base = acts + (size_t)t * per_t;
    if (which == 0) return base;
    if (which <= 3) return base + (size_t)B * D + (size_t)(which - 1) * B * H;
    return base + (size_t)B * D + (size_t)3 * B * H;
  };
  const int64_t oW[4] = {0, (int64_t)H * D + H, (int64_t)H * D + H + (int64_t)H * H + H, (int64_t)H * D + H + 2 * ((int64_t)H * H + H)};
  const int64_t oB[4] = {(int64_t)H * D, oW[1] + (int64_t)H * H, oW[2] + (int64_t)H * H, oW[3] + (int64_t)P * H};
  const int Kin[4] = {D, H, H, H}, Nout[4] = {H, H, H, P};

  lw_mask_kernel<<<grid_for(numel, 256, 8), 256, 0, st>>>(raw, mask, Wm, numel);
  lw_gather_kernel<<<1, 256, 0, st>>>(x, w, reinterpret_cast<const long long*>(idx_all), mask_all, reinterpret_cast<const long long*>(cursor), B, D,
                                      act(0, 0), coef, ladj);
  PMC_LAUNCH_CHECK();
  const int head_blocks = (B * D + 127) / 128;
  // ---------------- forward ----------------
  for (int t = 0; t < T; ++t) {
    const float* Wt = Wm + (size_t)t * tstride;
    for (int l = 0; l < 4; ++l) {
      LwGemm g{};
      g.A = act(t, l); g.lda = Kin[l];
      g.B = Wt + oW[l]; g.ldb = Kin[l];
      g.C = act(t, l + 1); g.ldc = Nout[l];
      g.M = B; g.N = Nout[l]; g.K = Kin[l];
      g.epi = LW_EPI_FWD; g.bias = Wt + oB[l];
      g.res = (l == 1 || l == 2) ? act(t, l) : nullptr; g.ldr = H;
      g.relu = l < 3 ? 1 : 0;
      lw_launch_gemm(false, true, g, st);
    }
    float* out = (t + 1 < T) ? act(t + 1, 0) : z;
    if (kind == 0) lw_head_fwd_kernel<false><<<head_blocks, 128, 0, st>>>(act(t, 0), act(t, 4), B, D, total, out, lrow);
    else lw_head_fwd_kernel<true><<<head_blocks, 128, 0, st>>>(act(t, 0), act(t, 4), B, D, total, out, lrow);
    lw_rowsum_kernel<<<(B + 127) / 128, 128, 0, st>>>(lrow, B, D, ladj);
  }
  lw_loss_kernel<<<(B + 127) / 128, 128, 0, st>>>(z, ladj, coef, B, D, gA, partials);
  PMC_LAUNCH_CHECK();
  if (!train) return 0;
  // ---------------- backward ----------------
  float* gy = gA;          // d loss / d (output of transform t)
  float* gnext = gB;
  for (int t = T - 1; t >= 0; --t) {
    float* Gt = grad + (size_t)t * tstride;
    const float* Wt = Wm + (size_t)t * tstride;
    const float* Mt = mask + (size_t)t * tstride;
    // head: dphi and the direct part of d loss / d v (into gnext)
    if (kind == 0) lw_head_bwd_kernel<false><<<head_blocks, 128, 0, st>>>(act(t, 0), act(t, 4), gy, coef, B, D, total, dphi, gnext);
    else lw_head_bwd_kernel<true><<<head_blocks, 128, 0, st>>>(act(t, 0), act(t, 4), gy, coef, B, D, total, dphi, gnext);
    // input-gradient chain: d input_l = dpre_l . Wm_l (+ dpre_l for the residual layers 1, 2), gated by the ReLU of the layer
    // below; at l = 0 the input is v: add the head's direct part, no gate.  dpre_3 = dphi, dpre_2..0 = dh[0..2]
    const float* dpre_of[4] = {dh[2], dh[1], dh[0], dphi};
    for (int l = 3; l >= 0; --l) {
      const float* dpre = dpre_of[l];
      LwGemm gd{};
      gd.A = dpre; gd.lda = Nout[l];
      gd.B = Wt + oW[l]; gd.ldb = Kin[l];
      gd.M = B; gd.N = Kin[l]; gd.K = Nout[l];
      gd.epi = LW_EPI_BWD;
      if (l == 0) { gd.C = gnext; gd.ldc = D; gd.res = gnext; gd.ldr = D; gd.gate = nullptr; }
      else {
        gd.C = const_cast<float*>(dpre_of[l - 1]); gd.ldc = H;
        gd.res = (l == 1 || l == 2) ? dpre : nullptr; gd.ldr = H;
        gd.gate = act(t, l); gd.ldg = H;               // h_{l-1} > 0  <=>  its pre-activation passed the ReLU
      }
      lw_launch_gemm(false, false, gd, st);
    }
    // the four weight gradients dW_l [Nout, Kin] = dpre_l^T . input_l (masked, straight into the gradient blob) in ONE
    // launch, the four bias gradients (column sums of dpre_l) in another
    LwGemm4 gw4{};
    LwColsum4 cs4{};
    int max_m = 0, max_n = 0, max_c = 0;
    for (int l = 0; l < 4; ++l) {
      LwGemm& gw = gw4.g[l];
      gw.A = dpre_of[l]; gw.lda = Nout[l];
      gw.B = act(t, l); gw.ldb = Kin[l];
      gw.C = Gt + oW[l]; gw.ldc = Kin[l];
      gw.M = Nout[l]; gw.N = Kin[l]; gw.K = B;
      gw.epi = LW_EPI_WGRAD; gw.gate = Mt + oW[l]; gw.ldg = Kin[l];
      max_m = std::max(max_m, (Nout[l] + LW_TM - 1) / LW_TM);
      max_n = std::max(max_n, (Kin[l] + LW_TN - 1) / LW_TN);
      cs4.Y[l] = dpre_of[l]; cs4.out[l] = Gt + oB[l]; cs4.N[l] = Nout[l];
      max_c = std::max(max_c, (Nout[l] + 31) / 32);
    }
    cs4.M = B;
    lw_wgrad4_kernel<<<dim3(max_n, max_m, 4), 256, 0, st>>>(gw4);
    lw_colsum4_kernel<<<dim3(max_c, 4), 256, 0, st>>>(cs4);
    std::swap(gy, gnext);
  }
  PMC_LAUNCH_CHECK();
  return 0;
}

// Univariate heads of the autoregressive flows (zuko 1.1: MonotonicAffineTransform of MAF, MonotonicRQSTransform of NSF;
// SURVEY App. A), shared by the fp32-FMA sweep (flow_sweep.cu) and the tcgen05 block-triangular sweep (flow_tri.cu).
#pragma once
#include "common.cuh"

namespace pmc {

constexpr float LOG_SLOPE = -6.90775527898213705205f;  // log(1e-3)

__device__ __forceinline__ float softclip(float a, float ls) { return a / (1.0f + fabsf(a / ls)); }

// ---- univariate transforms -----------------------------------------------------------------
// zuko MonotonicAffineTransform: y = x*exp(ls) + shift, ls soft-clipped.
struct Affine {
  static constexpr int TOTAL = 2, TP = 4;
  __device__ static __forceinline__ float apply(const float* phi, float v, bool inverse, float& ladj) {
    const float ls = softclip(phi[1], LOG_SLOPE);
    ladj = ls;
    const float sc = expf(ls);
    return inverse ? (v - phi[0]) / sc : fmaf(v, sc, phi[0]);
  }
};

// zuko MonotonicRQSTransform, bins = 8, bound = 5.
struct Rqs {
  static constexpr int BINS = 8, TOTAL = 23, TP = 24;
  __device__ static __forceinline__ void knots(const float* a, float* out /*BINS+1*/) {
    float c[BINS], mx = -INFINITY, sum = 0.f;
#pragma unroll
    for (int i = 0; i < BINS; ++i) { c[i] = softclip(a[i], 0.5f * LOG_SLOPE); mx = fmaxf(mx, c[i]); }
#pragma unroll
    for (int i = 0; i < BINS; ++i) { c[i] = expf(c[i] - mx); sum += c[i]; }
    double acc = 0.0;  // torch's CPU cumsum accumulates float inputs in double
    out[0] = -5.0f;
#pragma unroll
    for (int i = 0; i < BINS; ++i) {
      acc += (double)(c[i] / sum);
      out[i + 1] = 5.0f * (2.0f * (float)acc - 1.0f);
    }
  }
  __device__ static __forceinline__ float apply(const float* phi, float v, bool inverse, float& ladj) {
    float hx[BINS + 1], hy[BINS + 1], dv[BINS + 1];
    knots(phi, hx);
    knots(phi + BINS, hy);
    dv[0] = 1.0f; dv[BINS] = 1.0f;
#pragma unroll
    for (int i = 0; i < BINS - 1; ++i) dv[i + 1] = expf(softclip(phi[2 * BINS + i], LOG_SLOPE));
    int cnt = 0;  // searchsorted(left) = #knots < v
#pragma unroll
    for (int i = 0; i <= BINS; ++i) cnt += ((inverse ? hy[i] : hx[i]) < v) ? 1 : 0;
    const int k = cnt - 1;
    const bool in = (k >= 0) && (k < BINS);
    const int kk = ((k % BINS) + BINS) % BINS;
    float x0 = 0, x1 = 0, y0 = 0, y1 = 0, d0 = 0, d1 = 0;
#pragma unroll
    for (int i = 0; i < BINS; ++i)
      if (i == kk) { x0 = hx[i]; x1 = hx[i + 1]; y0 = hy[i]; y1 = hy[i + 1]; d0 = dv[i]; d1 = dv[i + 1]; }
    const float s = (y1 - y0) / (x1 - x0);
    const float t2 = d0 + d1 - 2.0f * s;
    float x = v, res = v;
    if (inverse) {
      const float y_ = in ? (v - y0) : 0.0f;
      const float a = (y1 - y0) * (s - d0) + y_ * t2;
      const float b = (y1 - y0) * d0 - y_ * t2;
      const float c = -s * y_;
      const float z = 2.0f * c / (-b - sqrtf(b * b - 4.0f * a * c));
      x = in ? (x0 + z * (x1 - x0)) : v;
      res = x;
    }
    const float z = in ? (x - x0) / (x1 - x0) : 0.0f;
    const float den = s + t2 * z * (1.0f - z);
    const float jac = s * s * (2.0f * s * z * (1.0f - z) + d0 * (1.0f - z) * (1.0f - z) + d1 * z * z) / (den * den);
    ladj = in ? logf(jac) : 0.0f;
    if (!inverse) res = in ? (y0 + (y1 - y0) * (s * z * z + d0 * z * (1.0f - z)) / den) : v;
    return res;
  }
};

// The same spline with the instruction count cut for the block-triangular sweep, where ONE thread runs it on the
// critical path of every order position (csrc/flow_tri.cu): the soft clips divide by a constant (one reciprocal
// multiply + one fast division each), the softmax normalises with one reciprocal per knot vector, the cumulative sums
// stay in fp32, and only the two slopes of the selected bin are exponentiated.  Same formulae as Rqs::apply; results
// agree to a few ulp of the knot positions (the spline flows' parity bar is 5e-4).
struct RqsLean {
  static constexpr int BINS = 8;
  __device__ static __forceinline__ float clip(const float a, const float inv_ls) { return __fdividef(a, 1.0f + fabsf(a * inv_ls)); }
  __device__ static __forceinline__ void knots(const float* a, float* out /*BINS+1*/) {
    constexpr float INV = 1.0f / (0.5f * LOG_SLOPE);
    float c[BINS], mx = -INFINITY, sum = 0.f;
#pragma unroll
    for (int i = 0; i < BINS; ++i) { c[i] = clip(a[i], INV); mx = fmaxf(mx, c[i]); }
#pragma unroll
    for (int i = 0; i < BINS; ++i) { c[i] = __expf(c[i] - mx); sum += c[i]; }
    const float scale = __fdividef(10.0f, sum);
    float acc = 0.f;
    out[0] = -5.0f;
#pragma unroll
    for (int i = 0; i < BINS; ++i) {
      acc += c[i];
      out[i + 1] = fmaf(acc, scale, -5.0f);
    }
    out[BINS] = 5.0f;
  }
  template <bool INVERSE>
  __device__ static __forceinline__ float apply(const float* phi, const float v, float& ladj) {
    constexpr float INV_LS = 1.0f / LOG_SLOPE;
    float hx[BINS + 1], hy[BINS + 1];
    knots(phi, hx);
    knots(phi + BINS, hy);
    int cnt = 0;  // searchsorted(left) = #knots < v
#pragma unroll
    for (int i = 0; i <= BINS; ++i) cnt += ((INVERSE ? hy[i] : hx[i]) < v) ? 1 : 0;
    const int k = cnt - 1;
    const bool in = (k >= 0) && (k < BINS);
    const int kk = in ? k : (k < 0 ? BINS - 1 : 0);        // ((k % BINS) + BINS) % BINS for k in {-1, BINS}
    float x0 = 0, x1 = 0, y0 = 0, y1 = 0, r0 = 0, r1 = 0;
#pragma unroll
    for (int i = 0; i < BINS; ++i)
      if (i == kk) {
        x0 = hx[i]; x1 = hx[i + 1]; y0 = hy[i]; y1 = hy[i + 1];
        r0 = i == 0 ? 0.f : phi[2 * BINS + i - 1];
        r1 = i == BINS - 1 ? 0.f : phi[2 * BINS + i];
      }
    const float d0 = kk == 0 ? 1.0f : __expf(clip(r0, INV_LS));
    const float d1 = kk == BINS - 1 ? 1.0f : __expf(clip(r1, INV_LS));
    const float iw = __fdividef(1.0f, x1 - x0);
    const float s = (y1 - y0) * iw;
    const float t2 = d0 + d1 - 2.0f * s;
    float x = v, res = v;
    if (INVERSE) {
      const float y_ = in ? (v - y0) : 0.0f;
      const float a = (y1 - y0) * (s - d0) + y_ * t2;
      const float b = (y1 - y0) * d0 - y_ * t2;
      const float c = -s * y_;
      const float z = __fdividef(2.0f * c, -b - sqrtf(b * b - 4.0f * a * c));
      x = in ? (x0 + z * (x1 - x0)) : v;
      res = x;
    }
    const float z = in ? (x - x0) * iw : 0.0f;
    const float den = s + t2 * z * (1.0f - z);
    const float iden = __fdividef(1.0f, den);
    const float jac = s * s * (2.0f * s * z * (1.0f - z) + d0 * (1.0f - z) * (1.0f - z) + d1 * z * z) * iden * iden;
    ladj = in ? logf(jac) : 0.0f;                          // the accurate log: 192 of these add up in one log-determinant
    if (!INVERSE) res = in ? (y0 + (y1 - y0) * (s * z * z + d0 * z * (1.0f - z)) * iden) : v;
    return res;
  }
};

}  // namespace pmc

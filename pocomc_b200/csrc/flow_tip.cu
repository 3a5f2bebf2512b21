// Bulk/tip sweep (EXPERIMENTAL, opt-in: config.sweep_variant = "tip"): Flow.forward / Flow.inverse of zuko MAF.
//
// Why: the default stream kernel (flow_sweep.cu) walks T * D * (L+1) dependent hops per particle and every hop is a
// full dot product + cross-lane reduction + shared-memory round trip: 430 ns per hop at 10 000 particles, 11 % of the
// fp32 FMA peak, the same time for 2 560 particles as for 10 000 (DESIGN.md sections 4-5).  The dependency is much
// thinner than the arithmetic.  At order position k the only values that are NEW are x_k and the hidden units of
// degree k + 1 (one group of <= 8 units per layer); everything else an output or hidden unit of this position reads
// was finished one position earlier.  So every dot product is split (made_layout.build_stream_tip):
//
//   bulk : the part over inputs finished one position earlier.  All bulks of position k (layers 0..L-1 of group
//          k + 1 and the output of feature k + 1) depend on nothing computed in position k: they are issued first,
//          back to back, with independent accumulators -- throughput work that hides the tip chain of other warps.
//   tip  : x_k (one FMA per unit) or the <= 8 fresh units of the previous layer.  The fresh values never touch
//          shared memory on the dependent path: a lane owns one unit of the group, the group is exchanged with four
//          shuffles, and the lane adds its tip, bias and residual to the bulk sum it already holds.
//
// The dependent chain per position shrinks from four full dot products to: out tip (<= 8 FMA) -> exp -> 1 FMA ->
// 4 shuffles + <= 8 FMA (x L - 1).  Same warp / lane mapping as the stream kernel with 4 lanes per particle
// (lane = q * 8 + p), same TMA-fed ring, same activation arrays [unit][8 particles] for the bulk reads.
//
// Status (round 1): layout and decomposition pinned against the oracle on the CPU (tests/sweep_emul.py:
// sweep_stream_tip), kernel parity-green on B200 (tests/test_zz_gpu_experimental.py::test_bulk_tip_sweep_matches_oracle), but
// SLOWER than the stream kernel in its first form: 388 us vs 331 us per 10 000-particle maf6 / 32-D inverse.  The
// dependent chain is shorter, the instruction stream is not (bulk dots + a reduce-scatter per chunk + shuffle
// exchanges), and no ncu capture exists yet -- the default stays the stream kernel.
#include "common.cuh"
#include "tc_common.cuh"
#include <algorithm>
#include <stdlib.h>

namespace pmc {
namespace tip {

using namespace tc;

// meta header slots -- keep in sync with made_layout.py / flow_sweep.cu
enum { M_D = 0, M_H, M_L, M_T, M_KIND, M_TOTAL, M_TP, M_NG, M_TSTRIDE, M_HP, M_MAXCH,
       M_OFF_GSTART, M_OFF_NCHUNK, M_OFF_SLOT, M_OFF_W0, M_OFF_WH, M_OFF_WO, M_OFF_B0, M_OFF_BH,
       M_OFF_BO, M_RAW_TSTRIDE, M_BINS, M_VERSION, M_NCHUNKS, M_SLOT_FLOATS, M_OFF_CHUNKS };

constexpr int NS = 3;               // ring depth (made_layout.STREAM_STAGES)
constexpr int MAX_THREADS = 640;    // 1 producer + 19 consumer warps
constexpr int PW = 8;               // particles per warp (4 lanes per particle)
constexpr int MAXL = 4;             // hidden layers held in registers as bulk sums
constexpr float LOG_SLOPE = -6.90775527898213705205f;  // log(1e-3), zuko MonotonicAffineTransform

// producer-side wait with back-off (a spinning lone lane steals issue slots from the consumers on its scheduler)
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  while (true) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (done) break;
    __nanosleep(400);
  }
}

// acc[j] = sum over this lane's rows s = q, q + 4, ... of slab[s][j] * act[s][p]; rows is a multiple of 16
__device__ __forceinline__ void dot4_partial(const float4* __restrict__ wp, const float* __restrict__ ap, int rows, float (&acc)[4]) {
  acc[0] = 0.f; acc[1] = 0.f; acc[2] = 0.f; acc[3] = 0.f;
#pragma unroll 2
  for (int s = 0; s < rows; s += 16) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 w = wp[4 * j];
      const float x = ap[32 * j];
      acc[0] = fmaf(w.x, x, acc[0]); acc[1] = fmaf(w.y, x, acc[1]);
      acc[2] = fmaf(w.z, x, acc[2]); acc[3] = fmaf(w.w, x, acc[3]);
    }
    wp += 16; ap += 128;
  }
}

// sum the 4 slices of a particle; afterwards lane q holds the complete sum of unit q
__device__ __forceinline__ float reduce_scatter4(const float (&acc)[4], int lane) {
  const bool hi = lane & 16;
  const float k0 = (hi ? acc[2] : acc[0]) + __shfl_xor_sync(FULL, hi ? acc[0] : acc[2], 16);
  const float k1 = (hi ? acc[3] : acc[1]) + __shfl_xor_sync(FULL, hi ? acc[1] : acc[3], 16);
  const bool mid = lane & 8;
  return (mid ? k1 : k0) + __shfl_xor_sync(FULL, mid ? k0 : k1, 8);
}

template <int MAXCH>
__global__ void __launch_bounds__(MAX_THREADS, 1)
made_sweep_tip_kernel(const float* __restrict__ stream, const int* __restrict__ meta, int meta_len,
                      const float* __restrict__ in, float* __restrict__ out, float* __restrict__ ladj_out,
                      long long n, int inverse, int ppc) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  int* sm = reinterpret_cast<int*>(smem_raw);
  for (int i = threadIdx.x; i < meta_len; i += blockDim.x) sm[i] = meta[i];
  __syncthreads();
  const int D = sm[M_D], H = sm[M_H], L = sm[M_L], T = sm[M_T], ng = sm[M_NG];
  const int Dp = (D + 15) & ~15, Hp = (H + 15) & ~15;
  const int tstride = sm[M_TSTRIDE], nchunks = sm[M_NCHUNKS], slot_floats = sm[M_SLOT_FLOATS];
  const int* gstart = sm + sm[M_OFF_GSTART];
  const int* nchunk = sm + sm[M_OFF_NCHUNK];
  const int* chunks = sm + sm[M_OFF_CHUNKS];
  size_t off = ((size_t)meta_len * 4 + 15) & ~(size_t)15;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + off);
  uint64_t* empty = full + NS;
  off = (off + 2 * NS * 8 + 127) & ~(size_t)127;
  float* ring = reinterpret_cast<float*>(smem_raw + off);
  off += (size_t)NS * slot_floats * 4;
  float* acts = reinterpret_cast<float*>(smem_raw + off);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long cta_row0 = (long long)blockIdx.x * ppc;
  const int cta_rows = (int)min((long long)ppc, n - cta_row0);
  const int active = (cta_rows + PW - 1) / PW;
  if (threadIdx.x == 0) {
    for (int i = 0; i < NS; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, active); }
    mbar_fence_init();
  }
  __syncthreads();

  if (warp == 0) {  // ---- producer
    if (lane == 0) {
      int it = 0;
      for (int tt = 0; tt < T; ++tt) {
        const int t = inverse ? (T - 1 - tt) : tt;
        const float* src = stream + (size_t)t * tstride;
        for (int c = 0; c < nchunks; ++c, ++it) {
          const int slot = it % NS;
          if (it >= NS) mbar_wait_backoff(empty + slot, ((it / NS) - 1) & 1);
          const uint32_t bytes = (uint32_t)chunks[4 * c + 3] * 4u;
          mbar_expect_tx(full + slot, bytes);
          bulk_g2s(ring + (size_t)slot * slot_floats, src + chunks[4 * c + 2], bytes, full + slot);
        }
      }
    }
    return;
  }
  const int cw = warp - 1;
  if (cw >= active) return;

  // ---- consumers: lane = q * 8 + p
  const int p = lane & 7, q = lane >> 3;
  const int per_warp = (D + Dp + L * Hp) * PW;
  float* cur = acts + (size_t)cw * per_warp;   // [D][PW] running vector (feature order)
  float* xs = cur + D * PW;                    // [Dp][PW] data-side values by order position
  float* act = xs + Dp * PW;                   // [L][Hp][PW] hidden activations by sorted unit
  const long long row0 = cta_row0 + (long long)cw * PW;
  const int rows = (int)min((long long)PW, n - row0);
  for (int i = lane; i < (Dp + L * Hp) * PW; i += 32) xs[i] = 0.0f;
  for (int i = lane; i < PW * D; i += 32) {
    const int r = i / D, c = i - r * D;
    cur[c * PW + r] = (r < rows) ? in[row0 * D + i] : 0.0f;
  }
  __syncwarp();
  const float* act_last = act + (size_t)(L - 1) * Hp * PW + lane;
  float ladj = 0.0f;
  int it = 0;
  for (int tt = 0; tt < T; ++tt) {
    const int t = inverse ? (T - 1 - tt) : tt;
    const bool rev = (t & 1);
    float bout0 = 0.f, bout1 = 0.f;              // bulk part of (shift, scale_raw) of the current position
    float fresh[4 * MAXCH];                      // last-layer activations of the group born one position ago
#pragma unroll
    for (int j = 0; j < 4 * MAXCH; ++j) fresh[j] = 0.f;
    for (int c = 0; c < nchunks; ++c, ++it) {
      const int slot = it % NS;
      mbar_wait(full + slot, (it / NS) & 1);
      const float4* w = reinterpret_cast<const float4*>(ring + (size_t)slot * slot_floats);
      const int k0 = chunks[4 * c], k1 = chunks[4 * c + 1];
      for (int k = k0; k < k1; ++k) {
        const int feat = rev ? (D - 1 - k) : k;
        // ---- tip head: (shift, scale_raw) = bulk + out tip + bias
        const int nchp = (k >= 1) ? nchunk[k - 1] : 0;
        float phi0 = bout0, phi1 = bout1;
#pragma unroll
        for (int cc = 0; cc < MAXCH; ++cc) {
          if (cc < nchp) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4 t4 = w[4 * cc + j];
              phi0 = fmaf(fresh[4 * cc + j], t4.x, phi0);
              phi1 = fmaf(fresh[4 * cc + j], t4.y, phi1);
            }
          }
        }
        {
          const float4 b4 = w[4 * nchp];
          phi0 += b4.x; phi1 += b4.y;
        }
        w += 4 * nchp + 1;
        const float v = cur[feat * PW + p];
        const float sden = 1.0f + fabsf(phi1 / LOG_SLOPE);
        const float ls = phi1 / sden;                                   // soft-clipped log-scale
        const float sc = expf(ls);
        const float res = inverse ? (v - phi0) / sc : fmaf(v, sc, phi0);
        ladj = inverse ? (ladj - ls) : (ladj + ls);
        const float xk = inverse ? res : v;
        const int g = k + 1;
        const bool has_group = g <= ng;
        // ---- bulk phase: nothing below depends on xk / res until the tips
        const int nch = has_group ? nchunk[k] : 0;
        const int ek16 = (gstart[k] + 15) & ~15;                       // sorted units of degree <= k, padded
        const int k16 = (k + 15) & ~15;                                // inputs of order < k, padded
        float bulk[MAXL][MAXCH];
#pragma unroll
        for (int l_ = 0; l_ < MAXL; ++l_)
#pragma unroll
          for (int cc = 0; cc < MAXCH; ++cc) bulk[l_][cc] = 0.f;
        if (has_group) {
#pragma unroll
          for (int l_ = 0; l_ < MAXL; ++l_) {
            if (l_ < L) {
              const int nrows = (l_ == 0) ? k16 : ek16;
              const float* src = (l_ == 0) ? (xs + lane) : (act + (size_t)(l_ - 1) * Hp * PW + lane);
#pragma unroll
              for (int cc = 0; cc < MAXCH; ++cc) {
                if (cc < nch) {
                  float acc[4];
                  dot4_partial(w + q, src, nrows, acc);
                  w += nrows;
                  bulk[l_][cc] = reduce_scatter4(acc, lane);
                }
              }
            }
          }
        }
        float nb0 = 0.f, nb1 = 0.f;
        if (k + 1 < D) {                                               // output bulk of the next position
          float acc[4];
          dot4_partial(w + q, act_last, ek16, acc);
          w += ek16;
          nb0 = acc[0] + __shfl_xor_sync(FULL, acc[0], 8);
          nb1 = acc[1] + __shfl_xor_sync(FULL, acc[1], 8);
          nb0 += __shfl_xor_sync(FULL, nb0, 16);
          nb1 += __shfl_xor_sync(FULL, nb1, 16);
        }
        // every lane has read cur[feat] / the activation arrays of this position: publish x_k
        __syncwarp();
        if (q == 0) {
          xs[k * PW + p] = xk;
          cur[feat * PW + p] = res;
        }
        // ---- tips of group g: lane q owns unit 4 cc + q of the group
        if (has_group) {
          const int gs = gstart[k], gsz = gstart[k + 1] - gs;
          float mine[MAXCH], prev[4 * MAXCH];
#pragma unroll
          for (int j = 0; j < 4 * MAXCH; ++j) prev[j] = 0.f;
#pragma unroll
          for (int cc = 0; cc < MAXCH; ++cc) mine[cc] = 0.f;
#pragma unroll
          for (int l_ = 0; l_ < MAXL; ++l_) {
            if (l_ < L) {
              float* dst = act + (size_t)l_ * Hp * PW;
              const int stride = (l_ == 0) ? 1 : (1 + nch);            // float4 per (chunk, lane-unit)
              float nw[MAXCH];
#pragma unroll
              for (int cc = 0; cc < MAXCH; ++cc) {
                nw[cc] = 0.f;
                if (cc < nch) {
                  const float4* base = w + (size_t)(4 * cc + q) * stride;
                  const float4 head = base[0];
                  float pre = bulk[l_][cc] + head.x;
                  if (l_ == 0) {
                    pre = fmaf(head.y, xk, pre);
                  } else {
#pragma unroll
                    for (int c2 = 0; c2 < MAXCH; ++c2) {
                      if (c2 < nch) {
                        const float4 t4 = base[1 + c2];
                        pre = fmaf(t4.x, prev[4 * c2 + 0], pre); pre = fmaf(t4.y, prev[4 * c2 + 1], pre);
                        pre = fmaf(t4.z, prev[4 * c2 + 2], pre); pre = fmaf(t4.w, prev[4 * c2 + 3], pre);
                      }
                    }
                    pre += mine[cc];                                    // residual hidden layers
                  }
                  nw[cc] = fmaxf(pre, 0.f);
                  if (4 * cc + q < gsz) dst[(gs + 4 * cc + q) * PW + p] = nw[cc];
                }
              }
              w += (size_t)4 * nch * stride;
              // exchange the group: every lane of the particle gets all 4 nch fresh values of this layer
#pragma unroll
              for (int cc = 0; cc < MAXCH; ++cc) {
                mine[cc] = nw[cc];
#pragma unroll
                for (int j = 0; j < 4; ++j) prev[4 * cc + j] = __shfl_sync(FULL, nw[cc], 8 * j + p);
              }
            }
          }
#pragma unroll
          for (int j = 0; j < 4 * MAXCH; ++j) fresh[j] = prev[j];
        }
        bout0 = nb0; bout1 = nb1;
        __syncwarp();   // x_k and the group's activations are visible to the bulk reads of the next position
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty + slot);
    }
  }
  for (int i = lane; i < PW * D; i += 32) {
    const int r = i / D, c = i - r * D;
    if (r < rows) out[row0 * D + i] = cur[c * PW + r];
  }
  if (q == 0 && p < rows) ladj_out[row0 + p] = ladj;
}

// ---- v2: the same schedule with PPL particles per lane (register blocking) -----------------------------------
// Arithmetic of the shared-memory pipe for the 4-lanes-per-particle mapping: one lane-row of a dot product costs an
// LDS.128 of weights -- 4 wavefronts, each delivering only 16 distinct bytes because the 8 lanes of a slice read the
// same float4 -- plus one wavefront of activations, for 4 FFMA warp instructions: 128 FMA per 5 wavefronts.  At one
// wavefront per cycle per SM that caps the FMA pipe at 20 %; the stream kernel sits at 11 % with ~75 % of the
// wavefront slots busy (58 M wavefronts in 387 us, profiles/r1c).  The sweep is shared-memory-bandwidth bound, not
// latency bound -- which is why halving its instruction count (flow_block.cu) or cutting its dependency chain (v1
// above) bought nothing.  With PPL particles per lane the same weight load feeds PPL x 4 FFMAs: 8 wavefronts per
// 512 FMA at PPL = 2, 8 per 1024 at PPL = 4.  That blocking was tried in the stream kernel and lost to latency
// (half / a quarter of the warps, each a serial chain of full dot products); with the bulk/tip split the chain is
// short, so here it can pay.  NOT yet run on a GPU: selected only by PMC_TIP_PPL = 2 | 4.
template <int PPL>
__device__ __forceinline__ void dot4_partial_ppl(const float4* __restrict__ wp, const float* __restrict__ ap, int rows,
                                                 float (&acc)[PPL][4]) {
#pragma unroll
  for (int e = 0; e < PPL; ++e) { acc[e][0] = 0.f; acc[e][1] = 0.f; acc[e][2] = 0.f; acc[e][3] = 0.f; }
#pragma unroll 2
  for (int s = 0; s < rows; s += 16) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 w = wp[4 * j];
      float x[PPL];
      if (PPL == 1) {
        x[0] = ap[32 * j];
      } else if (PPL == 2) {
        const float2 t2 = *reinterpret_cast<const float2*>(ap + 64 * j);
        x[0] = t2.x; x[PPL - 1] = t2.y;
      } else {
        const float4 t4 = *reinterpret_cast<const float4*>(ap + 128 * j);
        x[0] = t4.x; x[1 % PPL] = t4.y; x[2 % PPL] = t4.z; x[3 % PPL] = t4.w;
      }
#pragma unroll
      for (int e = 0; e < PPL; ++e) {
        acc[e][0] = fmaf(w.x, x[e], acc[e][0]); acc[e][1] = fmaf(w.y, x[e], acc[e][1]);
        acc[e][2] = fmaf(w.z, x[e], acc[e][2]); acc[e][3] = fmaf(w.w, x[e], acc[e][3]);
      }
    }
    wp += 16; ap += 128 * PPL;
  }
}

// consumer warps per CTA for PPL particles per lane: fewer, fatter warps leave room for the blocked accumulators
__host__ __device__ constexpr int ppl_threads(int ppl) { return ppl == 1 ? MAX_THREADS : ppl == 2 ? 352 : 192; }

template <int MAXCH, int PPL>
__global__ void __launch_bounds__(ppl_threads(PPL), 1)
made_sweep_tip_ppl_kernel(const float* __restrict__ stream, const int* __restrict__ meta, int meta_len,
                          const float* __restrict__ in, float* __restrict__ out, float* __restrict__ ladj_out,
                          long long n, int inverse, int ppc) {
  static_assert(PPL == 1 || PPL == 2 || PPL == 4, "particles per lane");
  constexpr int PWV = PW * PPL;                // particles per warp
  extern __shared__ __align__(128) unsigned char smem_raw[];
  int* sm = reinterpret_cast<int*>(smem_raw);
  for (int i = threadIdx.x; i < meta_len; i += blockDim.x) sm[i] = meta[i];
  __syncthreads();
  const int D = sm[M_D], H = sm[M_H], L = sm[M_L], T = sm[M_T], ng = sm[M_NG];
  const int Dp = (D + 15) & ~15, Hp = (H + 15) & ~15;
  const int tstride = sm[M_TSTRIDE], nchunks = sm[M_NCHUNKS], slot_floats = sm[M_SLOT_FLOATS];
  const int* gstart = sm + sm[M_OFF_GSTART];
  const int* nchunk = sm + sm[M_OFF_NCHUNK];
  const int* chunks = sm + sm[M_OFF_CHUNKS];
  size_t off = ((size_t)meta_len * 4 + 15) & ~(size_t)15;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + off);
  uint64_t* empty = full + NS;
  off = (off + 2 * NS * 8 + 127) & ~(size_t)127;
  float* ring = reinterpret_cast<float*>(smem_raw + off);
  off += (size_t)NS * slot_floats * 4;
  float* acts = reinterpret_cast<float*>(smem_raw + off);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long cta_row0 = (long long)blockIdx.x * ppc;
  const int cta_rows = (int)min((long long)ppc, n - cta_row0);
  const int active = (cta_rows + PWV - 1) / PWV;
  if (threadIdx.x == 0) {
    for (int i = 0; i < NS; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, active); }
    mbar_fence_init();
  }
  __syncthreads();

  if (warp == 0) {  // ---- producer
    if (lane == 0) {
      int it = 0;
      for (int tt = 0; tt < T; ++tt) {
        const int t = inverse ? (T - 1 - tt) : tt;
        const float* src = stream + (size_t)t * tstride;
        for (int c = 0; c < nchunks; ++c, ++it) {
          const int slot = it % NS;
          if (it >= NS) mbar_wait_backoff(empty + slot, ((it / NS) - 1) & 1);
          const uint32_t bytes = (uint32_t)chunks[4 * c + 3] * 4u;
          mbar_expect_tx(full + slot, bytes);
          bulk_g2s(ring + (size_t)slot * slot_floats, src + chunks[4 * c + 2], bytes, full + slot);
        }
      }
    }
    return;
  }
  const int cw = warp - 1;
  if (cw >= active) return;

  // ---- consumers: lane = q * 8 + p carries particles p * PPL + e, e < PPL, of the warp's PWV
  const int p = lane & 7, q = lane >> 3;
  const int pe = p * PPL;
  const int per_warp = (D + Dp + L * Hp) * PWV;
  float* cur = acts + (size_t)cw * per_warp;   // [D][PWV]
  float* xs = cur + D * PWV;                   // [Dp][PWV]
  float* act = xs + Dp * PWV;                  // [L][Hp][PWV]
  const long long row0 = cta_row0 + (long long)cw * PWV;
  const int rows = (int)min((long long)PWV, n - row0);
  for (int i = lane; i < (Dp + L * Hp) * PWV; i += 32) xs[i] = 0.0f;
  for (int i = lane; i < PWV * D; i += 32) {
    const int r = i / D, c = i - r * D;
    cur[c * PWV + r] = (r < rows) ? in[row0 * D + i] : 0.0f;
  }
  __syncwarp();
  const int lane_off = PPL * lane;             // row s = q + 4 j of particle pe + e sits at s * PWV + pe + e = lane_off + 32 PPL j + e
  const float* act_last = act + (size_t)(L - 1) * Hp * PWV + lane_off;
  float ladj[PPL];
#pragma unroll
  for (int e = 0; e < PPL; ++e) ladj[e] = 0.f;
  int it = 0;
  for (int tt = 0; tt < T; ++tt) {
    const int t = inverse ? (T - 1 - tt) : tt;
    const bool rev = (t & 1);
    float bout0[PPL], bout1[PPL], fresh[4 * MAXCH][PPL];
#pragma unroll
    for (int e = 0; e < PPL; ++e) {
      bout0[e] = 0.f; bout1[e] = 0.f;
#pragma unroll
      for (int j = 0; j < 4 * MAXCH; ++j) fresh[j][e] = 0.f;
    }
    for (int c = 0; c < nchunks; ++c, ++it) {
      const int slot = it % NS;
      mbar_wait(full + slot, (it / NS) & 1);
      const float4* w = reinterpret_cast<const float4*>(ring + (size_t)slot * slot_floats);
      const int k0 = chunks[4 * c], k1 = chunks[4 * c + 1];
      for (int k = k0; k < k1; ++k) {
        const int feat = rev ? (D - 1 - k) : k;
        // ---- tip head
        const int nchp = (k >= 1) ? nchunk[k - 1] : 0;
        float phi0[PPL], phi1[PPL];
#pragma unroll
        for (int e = 0; e < PPL; ++e) { phi0[e] = bout0[e]; phi1[e] = bout1[e]; }
#pragma unroll
        for (int cc = 0; cc < MAXCH; ++cc) {
          if (cc < nchp) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4 t4 = w[4 * cc + j];
#pragma unroll
              for (int e = 0; e < PPL; ++e) {
                phi0[e] = fmaf(fresh[4 * cc + j][e], t4.x, phi0[e]);
                phi1[e] = fmaf(fresh[4 * cc + j][e], t4.y, phi1[e]);
              }
            }
          }
        }
        const float4 b4 = w[4 * nchp];
        w += 4 * nchp + 1;
        float xk[PPL], res[PPL];
#pragma unroll
        for (int e = 0; e < PPL; ++e) {
          const float s0 = phi0[e] + b4.x, s1 = phi1[e] + b4.y;
          const float v = cur[feat * PWV + pe + e];
          const float ls = s1 / (1.0f + fabsf(s1 / LOG_SLOPE));
          const float sc = expf(ls);
          res[e] = inverse ? (v - s0) / sc : fmaf(v, sc, s0);
          ladj[e] = inverse ? (ladj[e] - ls) : (ladj[e] + ls);
          xk[e] = inverse ? res[e] : v;
        }
        const int g = k + 1;
        const bool has_group = g <= ng;
        // ---- bulk phase
        const int nch = has_group ? nchunk[k] : 0;
        const int ek16 = (gstart[k] + 15) & ~15;
        const int k16 = (k + 15) & ~15;
        float bulk[MAXL][MAXCH][PPL];
#pragma unroll
        for (int l_ = 0; l_ < MAXL; ++l_)
#pragma unroll
          for (int cc = 0; cc < MAXCH; ++cc)
#pragma unroll
            for (int e = 0; e < PPL; ++e) bulk[l_][cc][e] = 0.f;
        if (has_group) {
#pragma unroll
          for (int l_ = 0; l_ < MAXL; ++l_) {
            if (l_ < L) {
              const int nrows = (l_ == 0) ? k16 : ek16;
              const float* src = (l_ == 0) ? (xs + lane_off) : (act + (size_t)(l_ - 1) * Hp * PWV + lane_off);
#pragma unroll
              for (int cc = 0; cc < MAXCH; ++cc) {
                if (cc < nch) {
                  float acc[PPL][4];
                  dot4_partial_ppl<PPL>(w + q, src, nrows, acc);
                  w += nrows;
#pragma unroll
                  for (int e = 0; e < PPL; ++e) bulk[l_][cc][e] = reduce_scatter4(acc[e], lane);
                }
              }
            }
          }
        }
        float nb0[PPL], nb1[PPL];
#pragma unroll
        for (int e = 0; e < PPL; ++e) { nb0[e] = 0.f; nb1[e] = 0.f; }
        if (k + 1 < D) {
          float acc[PPL][4];
          dot4_partial_ppl<PPL>(w + q, act_last, ek16, acc);
          w += ek16;
#pragma unroll
          for (int e = 0; e < PPL; ++e) {
            float a0 = acc[e][0] + __shfl_xor_sync(FULL, acc[e][0], 8);
            float a1 = acc[e][1] + __shfl_xor_sync(FULL, acc[e][1], 8);
            a0 += __shfl_xor_sync(FULL, a0, 16);
            a1 += __shfl_xor_sync(FULL, a1, 16);
            nb0[e] = a0; nb1[e] = a1;
          }
        }
        __syncwarp();
        if (q == 0) {
#pragma unroll
          for (int e = 0; e < PPL; ++e) {
            xs[k * PWV + pe + e] = xk[e];
            cur[feat * PWV + pe + e] = res[e];
          }
        }
        // ---- tips
        if (has_group) {
          const int gs = gstart[k], gsz = gstart[k + 1] - gs;
          float mine[MAXCH][PPL], prev[4 * MAXCH][PPL];
#pragma unroll
          for (int e = 0; e < PPL; ++e) {
#pragma unroll
            for (int j = 0; j < 4 * MAXCH; ++j) prev[j][e] = 0.f;
#pragma unroll
            for (int cc = 0; cc < MAXCH; ++cc) mine[cc][e] = 0.f;
          }
#pragma unroll
          for (int l_ = 0; l_ < MAXL; ++l_) {
            if (l_ < L) {
              float* dst = act + (size_t)l_ * Hp * PWV;
              const int stride = (l_ == 0) ? 1 : (1 + nch);
              float nw[MAXCH][PPL];
#pragma unroll
              for (int cc = 0; cc < MAXCH; ++cc) {
#pragma unroll
                for (int e = 0; e < PPL; ++e) nw[cc][e] = 0.f;
                if (cc < nch) {
                  const float4* base = w + (size_t)(4 * cc + q) * stride;
                  const float4 head = base[0];
                  float pre[PPL];
#pragma unroll
                  for (int e = 0; e < PPL; ++e) pre[e] = bulk[l_][cc][e] + head.x;
                  if (l_ == 0) {
#pragma unroll
                    for (int e = 0; e < PPL; ++e) pre[e] = fmaf(head.y, xk[e], pre[e]);
                  } else {
#pragma unroll
                    for (int c2 = 0; c2 < MAXCH; ++c2) {
                      if (c2 < nch) {
                        const float4 t4 = base[1 + c2];
#pragma unroll
                        for (int e = 0; e < PPL; ++e) {
                          pre[e] = fmaf(t4.x, prev[4 * c2 + 0][e], pre[e]); pre[e] = fmaf(t4.y, prev[4 * c2 + 1][e], pre[e]);
                          pre[e] = fmaf(t4.z, prev[4 * c2 + 2][e], pre[e]); pre[e] = fmaf(t4.w, prev[4 * c2 + 3][e], pre[e]);
                        }
                      }
                    }
#pragma unroll
                    for (int e = 0; e < PPL; ++e) pre[e] += mine[cc][e];
                  }
#pragma unroll
                  for (int e = 0; e < PPL; ++e) {
                    nw[cc][e] = fmaxf(pre[e], 0.f);
                    if (4 * cc + q < gsz) dst[(gs + 4 * cc + q) * PWV + pe + e] = nw[cc][e];
                  }
                }
              }
              w += (size_t)4 * nch * stride;
#pragma unroll
              for (int cc = 0; cc < MAXCH; ++cc)
#pragma unroll
                for (int e = 0; e < PPL; ++e) {
                  mine[cc][e] = nw[cc][e];
#pragma unroll
                  for (int j = 0; j < 4; ++j) prev[4 * cc + j][e] = __shfl_sync(FULL, nw[cc][e], 8 * j + p);
                }
            }
          }
#pragma unroll
          for (int j = 0; j < 4 * MAXCH; ++j)
#pragma unroll
            for (int e = 0; e < PPL; ++e) fresh[j][e] = prev[j][e];
        }
#pragma unroll
        for (int e = 0; e < PPL; ++e) { bout0[e] = nb0[e]; bout1[e] = nb1[e]; }
        __syncwarp();
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty + slot);
    }
  }
  for (int i = lane; i < PWV * D; i += 32) {
    const int r = i / D, c = i - r * D;
    if (r < rows) out[row0 * D + i] = cur[c * PWV + r];
  }
  if (q == 0) {
#pragma unroll
    for (int e = 0; e < PPL; ++e)
      if (pe + e < rows) ladj_out[row0 + pe + e] = ladj[e];
  }
}

}  // namespace tip

int launch_tip(const float* stream, const int* meta, int meta_len, const int* hmeta, const float* in, float* out,
               float* ladj, long long n, int inverse, cudaStream_t st) {
  using namespace tip;
  const int D = hmeta[M_D], H = hmeta[M_H], L = hmeta[M_L];
  PMC_REQUIRE(hmeta[M_KIND] == 0 && hmeta[M_TOTAL] == 2, "pmc_flow_sweep: the bulk/tip sweep is built for affine transforms");
  PMC_REQUIRE(L >= 1 && L <= MAXL && hmeta[M_MAXCH] >= 1 && hmeta[M_MAXCH] <= 2, "pmc_flow_sweep: flow shape not built for the bulk/tip sweep");
  const size_t fixed = (((size_t)meta_len * 4 + 15) & ~(size_t)15) + 2 * NS * 8 + 256 + (size_t)NS * hmeta[M_SLOT_FLOATS] * 4;
  const size_t per_particle = (size_t)(D + ((D + 15) & ~15) + L * ((H + 15) & ~15)) * 4;
  const size_t budget = 227 * 1024;
  PMC_REQUIRE(fixed + PW * per_particle <= budget, "pmc_flow_sweep: flow too large for the bulk/tip sweep kernel");
  const int sms = sm_count();
  const long long max_smem = (long long)((budget - fixed) / per_particle);
  const long long per_sm = (n + sms - 1) / sms;
  int ppl = 1;                                                    // particles per lane: 1 = the validated kernel
  if (const char* e = getenv("PMC_TIP_PPL")) ppl = atoi(e);
  PMC_REQUIRE(ppl == 1 || ppl == 2 || ppl == 4, "pmc_flow_sweep: PMC_TIP_PPL must be 1, 2 or 4");
  const int pw = PW * ppl;
  PMC_REQUIRE(max_smem >= pw, "pmc_flow_sweep: flow too large for this many particles per lane");
  long long cap = std::min<long long>(max_smem, (long long)(ppl_threads(ppl) / 32 - 1) * pw) / pw * pw;
  const long long waves = (per_sm + cap - 1) / cap;
  long long ppc = (n + waves * sms - 1) / (waves * sms);
  ppc = std::min(cap, (ppc + pw - 1) / pw * pw);
  const long long grid = (n + ppc - 1) / ppc;
  const int threads = 32 * (1 + (int)(ppc / pw));
  const size_t smem = fixed + (size_t)ppc * per_particle;
  const int mc = hmeta[M_MAXCH];
#define PMC_TIP_LAUNCH(KERN)                                                                          \
  do {                                                                                                 \
    PMC_TRY(cudaFuncSetAttribute(KERN, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));       \
    KERN<<<(unsigned)grid, threads, smem, st>>>(stream, meta, meta_len, in, out, ladj, n, inverse, (int)ppc); \
  } while (0)
  if (ppl == 1) {
    if (mc == 1) PMC_TIP_LAUNCH(made_sweep_tip_kernel<1>);
    else PMC_TIP_LAUNCH(made_sweep_tip_kernel<2>);
  } else if (ppl == 2) {
    if (mc == 1) PMC_TIP_LAUNCH((made_sweep_tip_ppl_kernel<1, 2>));
    else PMC_TIP_LAUNCH((made_sweep_tip_ppl_kernel<2, 2>));
  } else {
    if (mc == 1) PMC_TIP_LAUNCH((made_sweep_tip_ppl_kernel<1, 4>));
    else PMC_TIP_LAUNCH((made_sweep_tip_ppl_kernel<2, 4>));
  }
#undef PMC_TIP_LAUNCH
  PMC_LAUNCH_CHECK();
  return 0;
}

}  // namespace pmc

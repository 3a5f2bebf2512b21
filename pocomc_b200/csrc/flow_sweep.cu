// Degree-ordered MADE sweep: Flow.forward / Flow.inverse for zuko MAF (affine) and NSF (RQS).
//
// Reference path: pocomc/flow.py:99-132 -> zuko transform.call_and_ladj / .inv.call_and_ladj
// (1 hyper-network pass forward, D+1 passes inverse).  Here one sweep in order of autoregressive
// degree computes every hidden unit and every output exactly once; the packed slab
// layout is built by pocomc_b200/made_layout.py.
//
// Mapping: a warp owns PW = 32/LPP particles; lane = q*PW + p (q = slice of the reduction, p =
// particle).  Activations live in shared memory as [unit][PW] so lane l touches word 32*j + l
// (conflict free); the weights of one 4-unit chunk are one 16-byte read-only load shared by the
// PW lanes of a slice.  fp32 FMA throughout -- the reference flow is fp32 (tools.py:292).
#include "common.cuh"
#include "flow_heads.cuh"
#include <algorithm>
#include <stdlib.h>

namespace pmc {

// meta header slots -- keep in sync with made_layout.py
enum { M_D = 0, M_H, M_L, M_T, M_KIND, M_TOTAL, M_TP, M_NG, M_TSTRIDE, M_HP, M_MAXCH,
       M_OFF_GSTART, M_OFF_NCHUNK, M_OFF_SLOT, M_OFF_W0, M_OFF_WH, M_OFF_WO, M_OFF_B0, M_OFF_BH,
       M_OFF_BO, M_RAW_TSTRIDE, M_BINS, M_VERSION, M_NCHUNKS, M_SLOT_FLOATS, M_OFF_CHUNKS };

// univariate heads (zuko MonotonicAffineTransform / MonotonicRQSTransform): flow_heads.cuh

// ---- partial dot products --------------------------------------------------------------------
// acc[0..3] = sum_{s in slice q} slab[s][col..col+3] * act[s][p], then butterfly over the LPP slices.
template <int LPP>
__device__ __forceinline__ void dot4(const float* __restrict__ slab, int wd, int col, int nrows,
                                     const float* act, int p, int q, float (&acc)[4]) {
  constexpr int PW = 32 / LPP;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, b0 = 0.f, b1 = 0.f, b2 = 0.f, b3 = 0.f;
  const float* wp = slab + (size_t)q * wd + col;
  const float* ap = act + q * PW + p;
  const size_t wstep = (size_t)LPP * wd;
  int s = q;
  for (; s + LPP < nrows; s += 2 * LPP) {
    const float4 w0 = __ldg(reinterpret_cast<const float4*>(wp));
    const float4 w1 = __ldg(reinterpret_cast<const float4*>(wp + wstep));
    const float x0 = ap[0], x1 = ap[32];
    a0 = fmaf(w0.x, x0, a0); a1 = fmaf(w0.y, x0, a1); a2 = fmaf(w0.z, x0, a2); a3 = fmaf(w0.w, x0, a3);
    b0 = fmaf(w1.x, x1, b0); b1 = fmaf(w1.y, x1, b1); b2 = fmaf(w1.z, x1, b2); b3 = fmaf(w1.w, x1, b3);
    wp += 2 * wstep; ap += 64;
  }
  if (s < nrows) {
    const float4 w0 = __ldg(reinterpret_cast<const float4*>(wp));
    const float x0 = ap[0];
    a0 = fmaf(w0.x, x0, a0); a1 = fmaf(w0.y, x0, a1); a2 = fmaf(w0.z, x0, a2); a3 = fmaf(w0.w, x0, a3);
  }
  acc[0] = a0 + b0; acc[1] = a1 + b1; acc[2] = a2 + b2; acc[3] = a3 + b3;
#pragma unroll
  for (int o = PW; o < 32; o <<= 1) {
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i] += __shfl_xor_sync(FULL, acc[i], o);
  }
}

template <class UNI, int LPP>
__global__ void __launch_bounds__(128)
made_sweep_kernel(const float* __restrict__ packed, const int* __restrict__ meta, int meta_len,
                  const float* __restrict__ in, float* __restrict__ out, float* __restrict__ ladj_out,
                  long long n, int inverse) {
  constexpr int PW = 32 / LPP;
  constexpr int TP = UNI::TP;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int* sm = reinterpret_cast<int*>(smem_raw);
  for (int i = threadIdx.x; i < meta_len; i += blockDim.x) sm[i] = meta[i];
  __syncthreads();
  const int D = sm[M_D], H = sm[M_H], L = sm[M_L], T = sm[M_T], ng = sm[M_NG];
  const int tstride = sm[M_TSTRIDE];
  const int* gstart = sm + sm[M_OFF_GSTART];
  const int* nchunk = sm + sm[M_OFF_NCHUNK];
  const int* slot = sm + sm[M_OFF_SLOT];
  const int* off_w0 = sm + sm[M_OFF_W0];
  const int* off_wh = sm + sm[M_OFF_WH];
  const int* off_wo = sm + sm[M_OFF_WO];
  const int* off_bh = sm + sm[M_OFF_BH];
  const int off_b0 = sm[M_OFF_B0], off_bo = sm[M_OFF_BO];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int p = lane % PW, q = lane / PW;
  const int per_warp = (2 * D + L * H) * PW;
  float* base = reinterpret_cast<float*>(smem_raw + (((size_t)meta_len * 4 + 15) & ~(size_t)15)) + (size_t)warp * per_warp;
  float* cur = base;            // [D][PW] running vector (feature order)
  float* xs = base + D * PW;    // [D][PW] data-side values by ORDER position (MLP inputs)
  float* act = xs + D * PW;     // [L][H][PW]

  const long long n_tiles = (n + PW - 1) / PW;
  const long long warps_total = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long tile = (long long)blockIdx.x * (blockDim.x >> 5) + warp; tile < n_tiles; tile += warps_total) {
    const long long row0 = tile * PW;
    const int rows = (int)min((long long)PW, n - row0);
    // stage the PW input rows (contiguous in global) transposed into cur[feat][p]
    for (int i = lane; i < PW * D; i += 32) {
      const int r = i / D, c = i - r * D;
      cur[c * PW + r] = (r < rows) ? in[row0 * D + i] : 0.0f;
    }
    __syncwarp();
    float ladj = 0.0f;
    for (int tt = 0; tt < T; ++tt) {
      const int t = inverse ? (T - 1 - tt) : tt;
      const float* P = packed + (size_t)t * tstride;
      const bool rev = (t & 1);
      for (int k = 0; k < D; ++k) {
        const int feat = rev ? (D - 1 - k) : k;
        const int Ek = gstart[k];  // sorted units with degree <= k
        float phi[TP];
        const float* slab = P + off_wo[k];
        const float* bo = P + off_bo + k * TP;
#pragma unroll
        for (int c = 0; c < TP / 4; ++c) {
          float acc[4];
          dot4<LPP>(slab, TP, 4 * c, Ek, act + (size_t)(L - 1) * H * PW, p, q, acc);
          const float4 b = __ldg(reinterpret_cast<const float4*>(bo + 4 * c));
          phi[4 * c + 0] = acc[0] + b.x; phi[4 * c + 1] = acc[1] + b.y;
          phi[4 * c + 2] = acc[2] + b.z; phi[4 * c + 3] = acc[3] + b.w;
        }
        const float v = cur[feat * PW + p];
        float l;
        const float res = UNI::apply(phi, v, inverse != 0, l);
        ladj = inverse ? (ladj - l) : (ladj + l);
        __syncwarp();  // every slice has read cur[feat] before it is overwritten
        if (q == 0) {
          xs[k * PW + p] = inverse ? res : v;
          cur[feat * PW + p] = res;
        }
        __syncwarp();
        const int g = k + 1;
        if (g > ng) continue;
        const int gs = gstart[g - 1], ge = gstart[g];
        if (ge == gs) continue;
        const int nch = nchunk[g - 1], wd = 4 * nch, sl = slot[g - 1];
        {  // input layer: units of degree g see the orders 0..g-1
          const float* w = P + off_w0[g - 1];
          for (int c = 0; c < nch; ++c) {
            float acc[4];
            dot4<LPP>(w, wd, 4 * c, g, xs, p, q, acc);
            const float4 b = __ldg(reinterpret_cast<const float4*>(P + off_b0 + sl + 4 * c));
            const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int u = gs + 4 * c + i;
              if (u < ge && q == (i % LPP)) act[u * PW + p] = fmaxf(acc[i] + bb[i], 0.0f);
            }
          }
        }
        __syncwarp();
        for (int l_ = 1; l_ < L; ++l_) {  // residual hidden layers: h + W h, then ReLU
          const float* w = P + off_wh[(l_ - 1) * ng + (g - 1)];
          const float* src = act + (size_t)(l_ - 1) * H * PW;
          float* dst = act + (size_t)l_ * H * PW;
          const int ob = off_bh[l_ - 1] + sl;
          for (int c = 0; c < nch; ++c) {
            float acc[4];
            dot4<LPP>(w, wd, 4 * c, ge, src, p, q, acc);
            const float4 b = __ldg(reinterpret_cast<const float4*>(P + ob + 4 * c));
            const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int u = gs + 4 * c + i;
              if (u < ge && q == (i % LPP)) dst[u * PW + p] = fmaxf(src[u * PW + p] + (acc[i] + bb[i]), 0.0f);
            }
          }
          __syncwarp();
        }
      }
    }
    // write back
    for (int i = lane; i < PW * D; i += 32) {
      const int r = i / D, c = i - r * D;
      if (r < rows) out[row0 * D + i] = cur[c * PW + r];
    }
    if (q == 0 && p < rows) ladj_out[row0 + p] = ladj;
    __syncwarp();
  }
}

// ---- v2: TMA-streamed weights ------------------------------------------------------------------
// The v1 kernel above reads every weight with a dependent global load, so each of the T*D*(L+1)
// sequential hops of a particle pays L1/L2 latency (ncu: 6% issue-active, 1.9 ms at ANY n).  Here the
// weights arrive in consumption order (made_layout.build_stream) through a shared-memory ring filled
// by one producer lane with cp.async.bulk (TMA 1-D bulk copy) + mbarrier complete_tx; the consumer
// warps (PW particles x LPP reduction slices each) only ever touch shared memory inside a hop.
constexpr int STREAM_STAGES = 3;   // ring depth; keep in sync with made_layout.STREAM_STAGES
constexpr int STREAM_MAX_THREADS = 640;   // 1 producer + 19 consumer warps: leaves 102 registers per thread

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// producer-side wait: back off between polls so the lone producer lane does not eat its scheduler's issue
// slots (ncu r1: 17 % of all issued instructions were this spin loop, all on one of the four schedulers)
__device__ __forceinline__ void mbar_wait_backoff(unsigned long long* bar, unsigned parity) {
  unsigned done = 0;
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
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// Partial dot products of one 4-unit slab for the PPL particles of a lane:
// acc[e][j] = sum over this lane's rows s = q, q+LPP, ... of slab[s][j] * act[s][p][e].
// `rows` is a multiple of 16 (the stream pads every slab with zero rows and the activation arrays
// are zero-initialised / finite), so the loop has no tail.
template <int LPP, int PPL>
__device__ __forceinline__ void dot4_partial(const float4* __restrict__ wp, const float* __restrict__ ap, int rows,
                                             float (&acc)[PPL][4]) {
  constexpr int STEP = 16 / LPP;          // lane-rows per 16 slab rows (4 / 2 / 1 for LPP = 4 / 8 / 16)
#pragma unroll
  for (int e = 0; e < PPL; ++e) { acc[e][0] = 0.f; acc[e][1] = 0.f; acc[e][2] = 0.f; acc[e][3] = 0.f; }
#pragma unroll 2
  for (int s = 0; s < rows; s += 16) {
#pragma unroll
    for (int j = 0; j < STEP; ++j) {
      const float4 w = wp[j * LPP];
      float x[PPL];
      if (PPL == 2) { const float2 t = *reinterpret_cast<const float2*>(ap + j * 32 * PPL); x[0] = t.x; x[PPL - 1] = t.y; }
      else x[0] = ap[j * 32];
#pragma unroll
      for (int e = 0; e < PPL; ++e) {
        acc[e][0] = fmaf(w.x, x[e], acc[e][0]); acc[e][1] = fmaf(w.y, x[e], acc[e][1]);
        acc[e][2] = fmaf(w.z, x[e], acc[e][2]); acc[e][3] = fmaf(w.w, x[e], acc[e][3]);
      }
    }
    wp += 16; ap += 16 * (32 / LPP) * PPL;
  }
}

// Sum the LPP slices of a particle.  Reduce-scatter: after the call every lane holds the complete sum
// of ONE unit, index ((lane>>4)&1)*2 + ((lane>>3)&1) (7 shuffles+adds instead of a 24-instruction
// butterfly over all four).
template <int LPP>
__device__ __forceinline__ float reduce_scatter4(const float (&acc)[4], int lane) {
  const bool hi = lane & 16;
  const float k0 = (hi ? acc[2] : acc[0]) + __shfl_xor_sync(FULL, hi ? acc[0] : acc[2], 16);
  const float k1 = (hi ? acc[3] : acc[1]) + __shfl_xor_sync(FULL, hi ? acc[1] : acc[3], 16);
  const bool mid = lane & 8;
  float r = (mid ? k1 : k0) + __shfl_xor_sync(FULL, mid ? k0 : k1, 8);
  if (LPP >= 8) r += __shfl_xor_sync(FULL, r, 4);
  if (LPP >= 16) r += __shfl_xor_sync(FULL, r, 2);
  return r;
}
template <int LPP>
__device__ __forceinline__ float butterfly(float v) {
#pragma unroll
  for (int o = 32 / LPP; o < 32; o <<= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}

// Warp = PWL lane-particles x LPP reduction slices, lane = q * PWL + p, each lane carrying PPL
// particles (register blocking: one weight load feeds PPL particles and the two dependency chains
// interleave).  A warp therefore owns PW = PWL * PPL particles; particle index in the warp = p*PPL + e.
template <class UNI, int LPP, int PPL>
__global__ void __launch_bounds__(STREAM_MAX_THREADS, 1)
made_sweep_stream_kernel(const float* __restrict__ stream, const int* __restrict__ meta, int meta_len,
                         const float* __restrict__ in, float* __restrict__ out, float* __restrict__ ladj_out,
                         long long n, int inverse, int ppc) {
  constexpr int PWL = 32 / LPP;
  constexpr int PW = PWL * PPL;
  constexpr int TP = UNI::TP;
  constexpr int NS = STREAM_STAGES;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  int* sm = reinterpret_cast<int*>(smem_raw);
  for (int i = threadIdx.x; i < meta_len; i += blockDim.x) sm[i] = meta[i];
  __syncthreads();
  const int D = sm[M_D], H = sm[M_H], L = sm[M_L], T = sm[M_T], ng = sm[M_NG];
  const int Dp = (D + 15) & ~15, Hp = (H + 15) & ~15;      // padded row counts of the activation arrays
  const int tstride = sm[M_TSTRIDE], nchunks = sm[M_NCHUNKS], slot_floats = sm[M_SLOT_FLOATS];
  const int* gstart = sm + sm[M_OFF_GSTART];
  const int* nchunk = sm + sm[M_OFF_NCHUNK];
  const int* chunks = sm + sm[M_OFF_CHUNKS];
  size_t off = ((size_t)meta_len * 4 + 15) & ~(size_t)15;
  unsigned long long* full = reinterpret_cast<unsigned long long*>(smem_raw + off);
  unsigned long long* empty = full + NS;
  off = (off + 2 * NS * 8 + 127) & ~(size_t)127;
  float* ring = reinterpret_cast<float*>(smem_raw + off);
  off += (size_t)NS * slot_floats * 4;
  float* acts = reinterpret_cast<float*>(smem_raw + off);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long cta_row0 = (long long)blockIdx.x * ppc;
  const int cta_rows = (int)min((long long)ppc, n - cta_row0);
  const int active = (cta_rows + PW - 1) / PW;          // consumer warps with particles
  if (threadIdx.x == 0) {
    for (int i = 0; i < NS; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, active); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == 0) {  // ---- producer: one lane streams T * nchunks bulk copies through the ring
    if (lane == 0) {
      int it = 0;
      for (int tt = 0; tt < T; ++tt) {
        const int t = inverse ? (T - 1 - tt) : tt;
        const float* src = stream + (size_t)t * tstride;
        for (int c = 0; c < nchunks; ++c, ++it) {
          const int slot = it % NS;
          if (it >= NS) mbar_wait_backoff(empty + slot, ((it / NS) - 1) & 1);
          const unsigned bytes = (unsigned)chunks[4 * c + 3] * 4u;
          mbar_expect_tx(full + slot, bytes);
          bulk_g2s(ring + (size_t)slot * slot_floats, src + chunks[4 * c + 2], bytes, full + slot);
        }
      }
    }
    return;
  }
  const int cw = warp - 1;
  if (cw >= active) return;

  // ---- consumers
  const int p = lane % PWL, q = lane / PWL;
  const int per_warp = (D + Dp + L * Hp) * PW;
  float* cur = acts + (size_t)cw * per_warp;  // [D][PW] running vector (feature order)
  float* xs = cur + D * PW;                   // [Dp][PW] data-side values by ORDER position (MLP inputs)
  float* act = xs + Dp * PW;                  // [L][Hp][PW]
  const long long row0 = cta_row0 + (long long)cw * PW;
  const int rows = (int)min((long long)PW, n - row0);
  for (int i = lane; i < (Dp + L * Hp) * PW; i += 32) xs[i] = 0.0f;
  for (int i = lane; i < PW * D; i += 32) {
    const int r = i / D, c = i - r * D;
    cur[c * PW + r] = (r < rows) ? in[row0 * D + i] : 0.0f;
  }
  __syncwarp();
  const int unit_i = ((lane >> 4) & 1) * 2 + ((lane >> 3) & 1);           // unit this lane owns after a reduce-scatter
  const bool writer = (LPP == 4) ? true : (LPP == 8) ? ((lane & 4) == 0) : ((lane & 6) == 0);  // one lane per (unit, particle)
  const int lane_off = lane * PPL;                                          // == (q*PWL + p) * PPL
  const int pe = p * PPL;                                                   // first particle of this lane in the warp tile
  const float* act_last = act + (size_t)(L - 1) * Hp * PW + lane_off;
  float ladj[PPL];
#pragma unroll
  for (int e = 0; e < PPL; ++e) ladj[e] = 0.0f;
  int it = 0;
  for (int tt = 0; tt < T; ++tt) {
    const int t = inverse ? (T - 1 - tt) : tt;
    const bool rev = (t & 1);
    for (int c = 0; c < nchunks; ++c, ++it) {
      const int slot = it % NS;
      mbar_wait(full + slot, (it / NS) & 1);
      const float4* w = reinterpret_cast<const float4*>(ring + (size_t)slot * slot_floats);
      const int k0 = chunks[4 * c], k1 = chunks[4 * c + 1];
      for (int k = k0; k < k1; ++k) {
        const int feat = rev ? (D - 1 - k) : k;
        const int Ek16 = (gstart[k] + 15) & ~15;   // sorted units with degree <= k, padded
        float phi[PPL][TP];
#pragma unroll
        for (int cc = 0; cc < TP / 4; ++cc) {
          float acc[PPL][4];
          dot4_partial<LPP, PPL>(w + q, act_last, Ek16, acc);
          w += Ek16;
#pragma unroll
          for (int e = 0; e < PPL; ++e)
#pragma unroll
            for (int j = 0; j < 4; ++j)
              phi[e][4 * cc + j] = (4 * cc + j < UNI::TOTAL) ? butterfly<LPP>(acc[e][j]) : 0.0f;
        }
#pragma unroll
        for (int cc = 0; cc < TP / 4; ++cc) {
          const float4 b = w[cc];
#pragma unroll
          for (int e = 0; e < PPL; ++e) {
            phi[e][4 * cc + 0] += b.x; phi[e][4 * cc + 1] += b.y; phi[e][4 * cc + 2] += b.z; phi[e][4 * cc + 3] += b.w;
          }
        }
        w += TP / 4;
        float v[PPL], res[PPL];
#pragma unroll
        for (int e = 0; e < PPL; ++e) {
          v[e] = cur[feat * PW + pe + e];
          float l;
          res[e] = UNI::apply(phi[e], v[e], inverse != 0, l);
          ladj[e] = inverse ? (ladj[e] - l) : (ladj[e] + l);
        }
        __syncwarp();  // every slice has read cur[feat] before it is overwritten
        if (q == 0) {
#pragma unroll
          for (int e = 0; e < PPL; ++e) {
            xs[k * PW + pe + e] = inverse ? res[e] : v[e];
            cur[feat * PW + pe + e] = res[e];
          }
        }
        __syncwarp();
        const int g = k + 1;
        if (g > ng) continue;
        const int gs = gstart[g - 1], ge = gstart[g];
        if (ge == gs) continue;
        const int nch = nchunk[g - 1];
        const int g16 = (g + 15) & ~15, Eg16 = (ge + 15) & ~15;
        const float* src = xs;
        float* dst = act;
        for (int l_ = 0; l_ < L; ++l_) {
          const int nrows = (l_ == 0) ? g16 : Eg16;
          const float* bias = reinterpret_cast<const float*>(w + (size_t)nch * nrows);
          int u = gs + unit_i;
          for (int cc = 0; cc < nch; ++cc, u += 4) {
            float acc[PPL][4];
            dot4_partial<LPP, PPL>(w + q, src + lane_off, nrows, acc);
            w += nrows;
            const float bv = bias[4 * cc + unit_i];
            float r[PPL];
#pragma unroll
            for (int e = 0; e < PPL; ++e) r[e] = reduce_scatter4<LPP>(acc[e], lane) + bv;
            if (writer && u < ge) {
#pragma unroll
              for (int e = 0; e < PPL; ++e) {
                if (l_ > 0) r[e] += src[u * PW + pe + e];          // residual hidden layers
                dst[u * PW + pe + e] = fmaxf(r[e], 0.0f);
              }
            }
          }
          w += nch;
          src = dst;
          dst += Hp * PW;
          __syncwarp();
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty + slot);
    }
  }
  for (int i = lane; i < PW * D; i += 32) {
    const int r = i / D, c = i - r * D;
    if (r < rows) out[row0 * D + i] = cur[c * PW + r];
  }
  if (q == 0) {
#pragma unroll
    for (int e = 0; e < PPL; ++e)
      if (pe + e < rows) ladj_out[row0 + pe + e] = ladj[e];
  }
}

__global__ void pack_kernel(const float* __restrict__ raw, const int* __restrict__ gather,
                            float* __restrict__ packed, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int g = gather[i];
    packed[i] = g >= 0 ? raw[g] : 0.0f;
  }
}

__global__ void base_logprob_kernel(const float* __restrict__ z, const float* __restrict__ ladj,
                                    float* __restrict__ lp, long long n, int d) {
  // warp per row: sum_d (-0.5 z^2 - 0.5 log 2pi) + ladj   (zuko DiagNormal.log_prob + ladj)
  const int lane = threadIdx.x & 31;
  const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n; r += warps) {
    float s = 0.f;
    for (int j = lane; j < d; j += 32) {
      const float v = z[r * d + j];
      s += -0.5f * v * v - 0.91893853320467274178f;
    }
    s = warp_sum(s);
    if (lane == 0) lp[r] = s + ladj[r];
  }
}

template <class UNI>
static int launch_sweep(const float* packed, const int* meta, int meta_len, const int* hmeta,
                        const float* in, float* out, float* ladj, long long n, int inverse, cudaStream_t st) {
  const int D = hmeta[M_D], H = hmeta[M_H], L = hmeta[M_L];
  const size_t meta_bytes = ((size_t)meta_len * 4 + 15) & ~(size_t)15;
  const int sms = sm_count();
  const int warps_per_block = 4;
  // pick lanes-per-particle: smallest LPP whose tile fits shared memory, then raise it while the
  // launch would leave SMs without enough resident warps (latency-bound regime at small N).
  int lpp = 1;
  auto smem_for = [&](int l) { return meta_bytes + (size_t)warps_per_block * (2 * D + L * H) * (32 / l) * 4; };
  while (lpp < 32 && smem_for(lpp) > 100 * 1024) lpp <<= 1;
  while (lpp < 8 && (n + (32 / lpp) - 1) / (32 / lpp) < (long long)sms * 16) lpp <<= 1;
  const size_t smem = smem_for(lpp);
  if (smem > 227 * 1024) { set_error("flow too large for the sweep kernel: %zu B shared memory", smem); return 3; }
  const long long tiles = (n + (32 / lpp) - 1) / (32 / lpp);
  long long blocks = (tiles + warps_per_block - 1) / warps_per_block;
  const long long cap = (long long)sms * std::max(1, (int)((200 * 1024) / smem));
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
#define PMC_SWEEP_CASE(LPPV)                                                                          \
  case LPPV: {                                                                                        \
    auto kern = made_sweep_kernel<UNI, LPPV>;                                                         \
    PMC_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));      \
    kern<<<(unsigned)blocks, warps_per_block * 32, smem, st>>>(packed, meta, meta_len, in, out, ladj, n, inverse); \
  } break;
  switch (lpp) {
    PMC_SWEEP_CASE(1) PMC_SWEEP_CASE(2) PMC_SWEEP_CASE(4) PMC_SWEEP_CASE(8) PMC_SWEEP_CASE(16) PMC_SWEEP_CASE(32)
  }
#undef PMC_SWEEP_CASE
  PMC_LAUNCH_CHECK();
  return 0;
}


template <class UNI>
static int launch_stream(const float* stream, const int* meta, int meta_len, const int* hmeta, const float* in,
                         float* out, float* ladj, long long n, int inverse, cudaStream_t st) {
  const int D = hmeta[M_D], H = hmeta[M_H], L = hmeta[M_L];
  const size_t fixed = (((size_t)meta_len * 4 + 15) & ~(size_t)15) + 2 * STREAM_STAGES * 8 + 256 +
                       (size_t)STREAM_STAGES * hmeta[M_SLOT_FLOATS] * 4;
  const size_t per_particle = (size_t)(D + ((D + 15) & ~15) + L * ((H + 15) & ~15)) * 4;
  const size_t budget = 227 * 1024;
  PMC_REQUIRE(fixed + 8 * per_particle <= budget, "pmc_flow_sweep: flow too large for the stream kernel");
  const int sms = sm_count();
  const long long max_smem = (long long)((budget - fixed) / per_particle);
  // lanes per particle / particles per lane: enough particles per SM -> 8 slices x 2 particles per lane
  // (register blocking); few particles -> more slices per particle to shorten the per-hop latency
  const long long per_sm = (n + sms - 1) / sms;
  // measured on B200 (tests/sweep_bench.py, D=32 maf6, 10 000 particles): LPP 4 / 8 / 16 = 344 / 378 / 666 us,
  // 2 particles per lane = no gain (fewer warps cancel the instruction savings)
  int lpp = 8, ppl = 1;
  if (per_sm <= 8) lpp = 16;
  if (per_sm >= 48) lpp = 4;
  if (const char* e = getenv("PMC_SWEEP_LPP")) lpp = atoi(e);   // tuning overrides
  if (const char* e = getenv("PMC_SWEEP_PPL")) ppl = atoi(e);
  if (UNI::TP > 4) ppl = 1;
  const int pw = (32 / lpp) * ppl;
  long long cap = std::min<long long>(max_smem, (long long)(STREAM_MAX_THREADS / 32 - 1) * pw) / pw * pw;
  if (cap < pw) { set_error("pmc_flow_sweep: shared memory too small for one tile"); return 3; }
  long long waves = (per_sm + cap - 1) / cap;
  long long ppc = (n + waves * sms - 1) / (waves * sms);
  ppc = std::min(cap, (ppc + pw - 1) / pw * pw);
  const long long grid = (n + ppc - 1) / ppc;
  const int threads = 32 * (1 + (int)(ppc / pw));
  const size_t smem = fixed + (size_t)ppc * per_particle;
#define PMC_STREAM_CASE(LPPV, PPLV)                                                                  \
  if (lpp == LPPV && ppl == PPLV) {                                                                  \
    auto kern = made_sweep_stream_kernel<UNI, LPPV, PPLV>;                                           \
    PMC_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));     \
    kern<<<(unsigned)grid, threads, smem, st>>>(stream, meta, meta_len, in, out, ladj, n, inverse, (int)ppc); \
    PMC_LAUNCH_CHECK();                                                                              \
    return 0;                                                                                        \
  }
  PMC_STREAM_CASE(4, 1) PMC_STREAM_CASE(8, 1) PMC_STREAM_CASE(16, 1)
  if constexpr (UNI::TP <= 4) { PMC_STREAM_CASE(4, 2) PMC_STREAM_CASE(8, 2) }
#undef PMC_STREAM_CASE
  set_error("pmc_flow_sweep: unsupported lanes-per-particle / particles-per-lane (%d, %d)", lpp, ppl);
  return 2;
}

}  // namespace pmc

using namespace pmc;

extern "C" int pmc_flow_pack(const float* raw, const int32_t* gather, float* packed, int64_t n, pmc_stream_t stream) {
  PMC_REQUIRE(raw && gather && packed && n > 0, "pmc_flow_pack: bad arguments");
  const int blocks = grid_for(n, 256, 8);
  pack_kernel<<<blocks, 256, 0, as_stream(stream)>>>(raw, gather, packed, n);
  PMC_LAUNCH_CHECK();
  return 0;
}

extern "C" int pmc_flow_sweep(const float* packed, const int32_t* meta, const int32_t* meta_host, int32_t meta_len,
                              const float* in, float* out, float* ladj, int64_t n, int32_t inverse,
                              pmc_stream_t stream) {
  PMC_REQUIRE(packed && meta && meta_host && in && out && ladj, "pmc_flow_sweep: null pointer");
  PMC_REQUIRE(meta_len >= 32, "pmc_flow_sweep: meta too short");
  if (n == 0) return 0;
  const int* hm = meta_host;
  PMC_REQUIRE(hm[M_D] >= 2 && hm[M_H] >= 1 && hm[M_L] >= 1 && hm[M_T] >= 1, "pmc_flow_sweep: bad meta header");
  if (hm[M_VERSION] == 2) {
    PMC_REQUIRE(hm[M_KIND] == 0 || (hm[M_BINS] == 8 && hm[M_TOTAL] == 23), "pmc_flow_sweep: only bins=8 splines are built");
    if (hm[M_KIND] == 0) return launch_stream<Affine>(packed, meta, meta_len, hm, in, out, ladj, n, inverse, as_stream(stream));
    return launch_stream<Rqs>(packed, meta, meta_len, hm, in, out, ladj, n, inverse, as_stream(stream));
  }
  if (hm[M_KIND] == 0) return launch_sweep<Affine>(packed, meta, meta_len, hm, in, out, ladj, n, inverse, as_stream(stream));
  PMC_REQUIRE(hm[M_BINS] == 8 && hm[M_TOTAL] == 23, "pmc_flow_sweep: only bins=8 splines are built");
  return launch_sweep<Rqs>(packed, meta, meta_len, hm, in, out, ladj, n, inverse, as_stream(stream));
}

extern "C" int pmc_flow_base_logprob(const float* z, const float* ladj, float* logprob, int64_t n, int32_t d,
                                     pmc_stream_t stream) {
  PMC_REQUIRE(z && ladj && logprob, "pmc_flow_base_logprob: null pointer");
  if (n == 0) return 0;
  const int blocks = grid_for(n, 8, 8);
  base_logprob_kernel<<<blocks, 256, 0, as_stream(stream)>>>(z, ladj, logprob, n, d);
  PMC_LAUNCH_CHECK();
  return 0;
}

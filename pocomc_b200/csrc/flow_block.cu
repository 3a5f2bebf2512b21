// Blocked degree-ordered MADE sweep: Flow.forward / Flow.inverse for zuko MAF (affine transforms).
//
// Reference path: pocomc/flow.py:99-132 -> zuko transform.call_and_ladj / .inv.call_and_ladj; the inverse
// (pocomc/mcmc.py:88) is ~93 % of every preconditioned MCMC step of the reference (D+1 hyper-network passes).
//
// The sweep of flow_sweep.cu is a nonlinear forward substitution with T*D*(L+1) dependent hops per particle,
// every hop a dot product over ALL units of lower degree.  Here the degrees are split into blocks
// (made_layout.build_block): what a block needs from earlier blocks is final before the block starts and is
// evaluated as dense [16 units x K] x [K x 8 particles] products on the warp tensor path (mma.sync m16n8k8
// TF32, 3-pass hi/lo split = fp32 fidelity, weights pre-split by pmc_flow_tc_pack) with no dependency chain;
// only the triangular part inside a block is still hop by hop, with dot products over <= one block of units
// evaluated one (unit, particle) pair per lane -- no cross-lane reduction.  The kernel is an interpreter
// over the layout's op program; weights arrive in consumption order through a shared-memory ring filled by a
// producer lane with cp.async.bulk + mbarrier complete_tx (same scheme as the stream kernel).
//
// tcgen05 is not used for the dense part on purpose: its minimum tile is 64/128 particles per CTA while a
// 10 000-particle launch leaves 68 particles per SM, and every block would pay a TMEM round trip.
//
// Warp = 8 particles.  Activations in shared memory as [particle][unit] with a row stride == 4 (mod 8)
// words, which makes the MMA B-fragment loads, the accumulator stores and the 128-bit row reads of the
// dot products all bank-conflict free.
#include "common.cuh"
#include <algorithm>
#include <stdlib.h>

namespace pmc {
namespace blk {

// meta header slots -- keep in sync with made_layout.py
enum { M_D = 0, M_H, M_L, M_T, M_KIND, M_TOTAL, M_TP, M_NG, M_TSTRIDE, M_VERSION = 22, M_NCHUNKS = 23, M_SLOT_FLOATS = 24,
       M_OFF_CHUNKS = 25, M_OFF_PROG = 26, M_NOPS = 27, M_HPB = 28, M_SX = 29, M_SO = 30 };
enum { OP_MMA = 0, OP_STEP = 1 };
enum { F_FIRST = 1, F_LAST = 2, F_NEWCHUNK = 4, F_BLOCKFIRST = 8, F_NEXTOUT = 16 };

constexpr int NS = 4;              // ring depth; keep in sync with made_layout.BLOCK_STAGES
constexpr int MAX_WARPS = 10;      // consumer warps per CTA
constexpr int PW = 8;              // particles per warp
constexpr float LOG_SLOPE = -6.90775527898213705205f;  // log(1e-3)

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
__device__ __forceinline__ bool mbar_try(unsigned long long* bar, unsigned parity) {
  unsigned done;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return done != 0;
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  while (!mbar_try(bar, parity)) {}
}
__device__ __forceinline__ void mbar_wait_backoff(unsigned long long* bar, unsigned parity) {
  while (!mbar_try(bar, parity)) __nanosleep(1500);
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const float4& a, unsigned b0, unsigned b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(__float_as_uint(a.x)), "r"(__float_as_uint(a.y)), "r"(__float_as_uint(a.z)), "r"(__float_as_uint(a.w)),
                 "r"(b0), "r"(b1));
}
// x = hi + lo exactly; hi has its 13 low mantissa bits cleared (what the tensor core keeps of an fp32 operand)
__device__ __forceinline__ void split_tf32(float x, unsigned& hi, unsigned& lo) {
  hi = __float_as_uint(x) & 0xffffe000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}
__device__ __forceinline__ float softclip(float a, float ls) { return a / (1.0f + fabsf(a / ls)); }
// four independent accumulation chains per dot product (one per float4 component)
__device__ __forceinline__ void fma4(float4& a, const float4& w, const float4& x) {
  a.x = fmaf(w.x, x.x, a.x); a.y = fmaf(w.y, x.y, a.y); a.z = fmaf(w.z, x.z, a.z); a.w = fmaf(w.w, x.w, a.w);
}
__device__ __forceinline__ float sum4(const float4& a) { return (a.x + a.y) + (a.z + a.w); }
__device__ __forceinline__ float dotv(const float4& w, float v0, float v1, float v2, float v3) {
  return fmaf(w.x, v0, w.y * v1) + fmaf(w.z, v2, w.w * v3);
}

// Degree group of one step through the L hidden layers (+ its share of the next output).  Lane = unit slot
// hi_ (0..3; second pass: 4..7 when TWO) x particle hp.  Rows that were final before the step come from shared
// memory ("old" dots, 4 accumulation chains each); values born in this step travel by shuffles only.
template <bool TWO>
__device__ __forceinline__ float step_layers(const float4*& w4, float* xs, float* act, int L, int sx, int sh, int hp, int hi_,
                                             int u0, int pbj, int l0_d0, int l0_r, int lh, float xk, bool next_out) {
  constexpr int S = TWO ? 8 : 4, P = TWO ? 2 : 1;
  float h0, h1 = 0.f;
  {  // layer 0: orders of the block before k from xs, order k (xk) from the register
    float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
    const float4* xq = reinterpret_cast<const float4*>(xs + hp * sx + l0_d0);
    const float4* wq = w4 + hi_;
    const float4* const we = wq + (l0_r >> 2) * S;
    for (; wq != we; wq += S, ++xq) {
      const float4 x = *xq;
      fma4(a0, wq[0], x);
      if (TWO) fma4(a1, wq[4], x);
    }
    const float* wn = reinterpret_cast<const float*>(we - hi_);
    float* dst = act + hp * sh + u0 + hi_;
    h0 = fmaxf(dst[0] + sum4(a0) + wn[hi_] * xk, 0.0f);
    dst[0] = h0;
    if (TWO) { h1 = fmaxf(dst[4] + sum4(a1) + wn[4 + hi_] * xk, 0.0f); dst[4] = h1; }
    w4 = we - hi_ + P;
  }
  const float* ap = act + hp * sh + pbj;
  float* al = act + PW * sh + hp * sh + u0 + hi_;
  const int nold = (lh >> 2) * S;
  for (int l_ = 1; l_ < L; ++l_) {
    const float v0 = __shfl_sync(FULL, h0, hp), v1 = __shfl_sync(FULL, h0, 8 + hp);
    const float v2 = __shfl_sync(FULL, h0, 16 + hp), v3 = __shfl_sync(FULL, h0, 24 + hp);
    float v4 = 0.f, v5 = 0.f, v6 = 0.f, v7 = 0.f;
    if (TWO) {
      v4 = __shfl_sync(FULL, h1, hp); v5 = __shfl_sync(FULL, h1, 8 + hp);
      v6 = __shfl_sync(FULL, h1, 16 + hp); v7 = __shfl_sync(FULL, h1, 24 + hp);
    }
    float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
    const float4* xq = reinterpret_cast<const float4*>(ap);
    const float4* wq = w4 + hi_;
    const float4* const we = wq + nold;
    for (; wq != we; wq += S, ++xq) {
      const float4 x = *xq;
      fma4(a0, wq[0], x);
      if (TWO) fma4(a1, wq[4], x);
    }
    const float4* wn = we - hi_;                                 // [S dst][S src]
    float p0 = al[0] + h0 + sum4(a0) + dotv(wn[hi_ * P], v0, v1, v2, v3);           // h0: residual of the own unit
    if (TWO) {
      p0 += dotv(wn[hi_ * P + 1], v4, v5, v6, v7);
      const float p1 = al[4] + h1 + sum4(a1) + dotv(wn[(4 + hi_) * P], v0, v1, v2, v3) + dotv(wn[(4 + hi_) * P + 1], v4, v5, v6, v7);
      h1 = fmaxf(p1, 0.0f);
      al[4] = h1;
    }
    h0 = fmaxf(p0, 0.0f);
    al[0] = h0;
    w4 = wn + S * P;
    ap += PW * sh; al += PW * sh;
  }
  float onew = 0.f;
  if (next_out) {                                                // the group's share of output k+1
    const float v0 = __shfl_sync(FULL, h0, hp), v1 = __shfl_sync(FULL, h0, 8 + hp);
    const float v2 = __shfl_sync(FULL, h0, 16 + hp), v3 = __shfl_sync(FULL, h0, 24 + hp);
    onew = dotv(w4[(hi_ & 1) * P], v0, v1, v2, v3);
    if (TWO) {
      const float v4 = __shfl_sync(FULL, h1, hp), v5 = __shfl_sync(FULL, h1, 8 + hp);
      const float v6 = __shfl_sync(FULL, h1, 16 + hp), v7 = __shfl_sync(FULL, h1, 24 + hp);
      onew += dotv(w4[(hi_ & 1) * P + 1], v4, v5, v6, v7);
    }
    w4 += 2 * P;
  }
  return onew;
}

__global__ void __launch_bounds__(32 * (1 + MAX_WARPS), 1)
made_sweep_block_kernel(const float* __restrict__ stream, const int* __restrict__ meta, int meta_len,
                        const float* __restrict__ in, float* __restrict__ out, float* __restrict__ ladj_out,
                        long long n, int inverse, int ppc) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  int* sm = reinterpret_cast<int*>(smem_raw);
  for (int i = threadIdx.x; i < meta_len; i += blockDim.x) sm[i] = meta[i];
  __syncthreads();
  const int D = sm[M_D], L = sm[M_L], T = sm[M_T];
  const int tstride = sm[M_TSTRIDE], nchunks = sm[M_NCHUNKS], slot_floats = sm[M_SLOT_FLOATS];
  const int nops = sm[M_NOPS], sx = sm[M_SX], so = sm[M_SO], sh = sm[M_HPB] + 4;
  const int* chunks = sm + sm[M_OFF_CHUNKS];
  const int4* prog = reinterpret_cast<const int4*>(sm + sm[M_OFF_PROG]);
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
  const int per_warp = PW * (D + sx + L * sh + so);
  float* cur = acts + (size_t)cw * per_warp;  // [D][8]        running vector, feature order
  float* xs = cur + D * PW;                   // [8][sx]       data-side values by ORDER position (layer-0 inputs)
  float* act = xs + PW * sx;                  // [L][8][sh]    hidden activations, block-padded unit index
  float* ph = act + (size_t)L * PW * sh;      // [8][so]       (shift, log-scale) of the block's order positions
  const long long row0 = cta_row0 + (long long)cw * PW;
  const int rows = (int)min((long long)PW, n - row0);
  for (int i = lane; i < PW * (sx + L * sh + so); i += 32) xs[i] = 0.0f;
  for (int i = lane; i < PW * D; i += 32) {
    const int r = i / D, c = i - r * D;
    cur[c * PW + r] = (r < rows) ? in[row0 * D + i] : 0.0f;
  }
  __syncwarp();
  const int fr = lane >> 2, fc = lane & 3;      // MMA fragment row / column of this lane
  const int hi_ = lane >> 3, hp = lane & 7;     // dot-product mapping: unit slot / particle
  float hh[3][4], hl[3][4], lh[3][4];           // per tile: a_hi*b_hi, a_hi*b_lo, a_lo*b_hi (3xTF32), carried across op pieces
#pragma unroll
  for (int tl = 0; tl < 3; ++tl)
#pragma unroll
    for (int i = 0; i < 4; ++i) { hh[tl][i] = 0.f; hl[tl][i] = 0.f; lh[tl][i] = 0.f; }
  float ladj = 0.0f, onew = 0.0f;
  int it = -1, slot = 0;
  const float* w = ring;
  float* const act_last = act + (size_t)(L - 1) * PW * sh;
  int4 n0 = prog[0], n1 = prog[1];
  for (int tt = 0; tt < T; ++tt) {
    const int t = inverse ? (T - 1 - tt) : tt;
    const bool rev = (t & 1);
    for (int op = 0; op < nops; ++op) {
      const int4 o0 = n0, o1 = n1;                               // type a b c | d e f flags
      {                                                          // next descriptor is in flight while this op runs
        const int nx = (op + 1 == nops) ? 0 : op + 1;
        n0 = prog[2 * nx]; n1 = prog[2 * nx + 1];
      }
      if (o1.w & F_NEWCHUNK) {
        __syncwarp();                                            // every lane is done reading the slot being released
        if (it >= 0 && lane == 0) mbar_arrive(empty + slot);
        ++it;
        slot = it % NS;
        mbar_wait(full + slot, (it / NS) & 1);
        w = ring + (size_t)slot * slot_floats;
      }
      if (o0.x == OP_STEP) {
        const int k = o0.y & 0xffff, c = o0.y >> 16, out_old = o0.z & 0xffff, pbj = o0.z >> 16;
        const int u0 = o0.w & 0xffff, cnt = o0.w >> 16, l0_d0 = o1.x & 0xffff, l0_r = o1.x >> 16, lh = o1.y;
        const float4* w4 = reinterpret_cast<const float4*>(w);
        if (o1.w & F_BLOCKFIRST) onew = 0.f;
        // ---- output k.  Lane = (param, K-half) x particle.  ph holds bias + everything before the block (phase A),
        //      onew the share of degree group k (carried in registers from the previous step), the rest is read here.
        float phi = onew + ph[hp * so + c + (hi_ & 1)];
        if (out_old) {
          const int par = hi_ & 1, q = hi_ >> 1;
          const float4* xq = reinterpret_cast<const float4*>(act_last + hp * sh + pbj) + q;
          const float4* wq = w4 + 2 * q + par;
          const float4* const wend = w4 + (out_old >> 1);
          float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
          for (; wq < wend; wq += 4, xq += 2) fma4(a, *wq, *xq);
          float sacc = sum4(a);
          sacc += __shfl_xor_sync(FULL, sacc, 16);
          phi += sacc;
          w4 = wend;
        }
        const float ls_raw = __shfl_down_sync(FULL, phi, 8);     // lanes 8..15 hold the log-scale of particle lane-8
        // ---- univariate transform (zuko MonotonicAffineTransform: y = x * exp(ls) + shift, ls soft-clipped)
        float xval = 0.f;
        if (lane < PW) {
          const int feat = rev ? (D - 1 - k) : k;
          const float ls = softclip(ls_raw, LOG_SLOPE);
          const float v = cur[feat * PW + lane];
          const float sc = expf(ls);
          const float res = inverse ? (v - phi) / sc : fmaf(v, sc, phi);
          ladj = inverse ? (ladj - ls) : (ladj + ls);
          xval = inverse ? res : v;
          xs[lane * sx + k] = xval;
          cur[feat * PW + lane] = res;
        }
        onew = 0.f;
        if (cnt) {
          const float xk = __shfl_sync(FULL, xval, hp);
          const bool nxt = (o1.w & F_NEXTOUT) != 0;
          onew = (cnt > 4) ? step_layers<true>(w4, xs, act, L, sx, sh, hp, hi_, u0, pbj, l0_d0, l0_r, lh, xk, nxt)
                           : step_layers<false>(w4, xs, act, L, sx, sh, hp, hi_, u0, pbj, l0_d0, l0_r, lh, xk, nxt);
        }
        w = reinterpret_cast<const float*>(w4);
        __syncwarp();
      } else {  // OP_MMA: up to 3 consecutive 16-unit tiles x 8 particles, 3xTF32
        const int sid = o0.y, nt = o1.z;
        const float* src;
        int ss;
        if (sid == 0) { src = xs; ss = sx; } else { src = act + (size_t)(sid - 1) * PW * sh; ss = sh; }
        src += fr * ss + o0.z + fc;                              // B fragment: particle fr, k = fc (+4)
        const float4* w4 = reinterpret_cast<const float4*>(w) + lane;
        if (o1.w & F_FIRST) {
#pragma unroll
          for (int tl = 0; tl < 3; ++tl)
            if (tl < nt) {
              const float4 b = w4[32 * tl];
              hh[tl][0] = b.x; hh[tl][1] = b.y; hh[tl][2] = b.z; hh[tl][3] = b.w;
#pragma unroll
              for (int i = 0; i < 4; ++i) { hl[tl][i] = 0.f; lh[tl][i] = 0.f; }
            }
          w4 += 32 * nt;
        }
        const int nks = o0.w;
        for (int ks = 0; ks < nks; ++ks) {
          const float x0 = src[8 * ks], x1 = src[8 * ks + 4];
          unsigned bh0, bl0, bh1, bl1;
          split_tf32(x0, bh0, bl0);
          split_tf32(x1, bh1, bl1);
#pragma unroll
          for (int tl = 0; tl < 3; ++tl)
            if (tl < nt) {
              const float4 ah = w4[64 * tl], al = w4[64 * tl + 32];
              mma_tf32(hh[tl], ah, bh0, bh1);
              mma_tf32(hl[tl], ah, bl0, bl1);
              mma_tf32(lh[tl], al, bh0, bh1);
            }
          w4 += 64 * nt;
        }
        w = reinterpret_cast<const float*>(w4 - lane);
        if (o1.w & F_LAST) {
          const int did = o1.x;
          float* dst;
          int ds;
          if (did <= L) { dst = act + (size_t)(did - 1) * PW * sh; ds = sh; } else { dst = ph; ds = so; }
          dst += (2 * fc) * ds + o1.y + fr;                      // C fragment: units fr / fr+8, particles 2fc / 2fc+1
#pragma unroll
          for (int tl = 0; tl < 3; ++tl)
            if (tl < nt) {
              dst[16 * tl] = hh[tl][0] + (hl[tl][0] + lh[tl][0]);
              dst[16 * tl + ds] = hh[tl][1] + (hl[tl][1] + lh[tl][1]);
              dst[16 * tl + 8] = hh[tl][2] + (hl[tl][2] + lh[tl][2]);
              dst[16 * tl + ds + 8] = hh[tl][3] + (hl[tl][3] + lh[tl][3]);
            }
          __syncwarp();
        }
      }
    }
  }
  if (it >= 0 && lane == 0) mbar_arrive(empty + slot);
  for (int i = lane; i < PW * D; i += 32) {
    const int r = i / D, c = i - r * D;
    if (r < rows) out[row0 * D + i] = cur[c * PW + r];
  }
  if (lane < rows) ladj_out[row0 + lane] = ladj;
}

}  // namespace blk

int launch_block(const float* stream, const int* meta, int meta_len, const int* hmeta, const float* in, float* out,
                 float* ladj, long long n, int inverse, cudaStream_t st) {
  using namespace blk;
  const int D = hmeta[M_D], L = hmeta[M_L];
  PMC_REQUIRE(hmeta[M_KIND] == 0 && hmeta[M_TOTAL] == 2, "pmc_flow_sweep: the blocked sweep is built for affine transforms");
  PMC_REQUIRE(hmeta[M_OFF_PROG] % 4 == 0 && hmeta[M_OFF_PROG] + 8 * hmeta[M_NOPS] <= meta_len, "pmc_flow_sweep: bad op program");
  const size_t fixed = (((size_t)meta_len * 4 + 15) & ~(size_t)15) + 2 * NS * 8 + 256 + (size_t)NS * hmeta[M_SLOT_FLOATS] * 4;
  const size_t per_warp = (size_t)PW * (D + hmeta[M_SX] + L * (hmeta[M_HPB] + 4) + hmeta[M_SO]) * 4;
  const size_t budget = 227 * 1024;
  PMC_REQUIRE(fixed + per_warp <= budget, "pmc_flow_sweep: flow too large for the blocked sweep kernel");
  const int sms = sm_count();
  int max_warps = (int)std::min<size_t>(MAX_WARPS, (budget - fixed) / per_warp);
  if (const char* e = getenv("PMC_BLOCK_WARPS")) max_warps = std::max(1, std::min(max_warps, atoi(e)));
  const long long cap = (long long)max_warps * PW;
  const long long per_sm = (n + sms - 1) / sms;
  const long long waves = (per_sm + cap - 1) / cap;
  long long ppc = (n + waves * sms - 1) / (waves * sms);
  ppc = std::min(cap, (ppc + PW - 1) / PW * PW);
  const long long grid = (n + ppc - 1) / ppc;
  const int threads = 32 * (1 + (int)(ppc / PW));
  const size_t smem = fixed + (size_t)(ppc / PW) * per_warp;
  PMC_TRY(cudaFuncSetAttribute(made_sweep_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  made_sweep_block_kernel<<<(unsigned)grid, threads, smem, st>>>(stream, meta, meta_len, in, out, ladj, n, inverse, (int)ppc);
  PMC_LAUNCH_CHECK();
  return 0;
}

}  // namespace pmc

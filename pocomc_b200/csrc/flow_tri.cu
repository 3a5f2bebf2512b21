// Windowed block-triangular MADE sweep on the 5th-generation tensor cores: Flow.inverse (and forward) of zuko MAF flows
// of any width (D = 8 .. 200 at the preset hidden widths 32 .. 1024).
//
// Reference path: pocomc/flow.py:116-132 -> zuko transform.inv.call_and_ladj, the call that dominates every
// preconditioned MCMC step (pocomc/mcmc.py:88,256): D+1 dense hyper-network passes per transform.  Sorted by
// autoregressive degree the masks are block lower-triangular, so ONE forward substitution computes every hidden unit
// and every output once.  This kernel runs that substitution with the dense part on tcgen05:
//
//   * a CTA owns 128 particles = the 128 TMEM lanes; order positions are cut into blocks of 4, blocks are grouped into
//     WINDOWS (pocomc_b200/tri_layout.py): TENSOR MEMORY holds the running pre-activations of the hidden units of a
//     window (three layers) and of its outputs, one fp32 column each, <= 512 columns;
//   * inside a window the schedule is right-looking: when a block is finished, its activations -- an A tile [128 x K]
//     per layer, TF32 hi / lo images in shared memory -- update the accumulators of ALL later columns of the window,
//     one group of tcgen05.mma (A and B from shared memory, K-major no-swizzle) per layer;
//   * a finished block also appends its activations (plain fp32, same K-major image) to a per-CTA scratch area in
//     global memory; when a window ends the next window's accumulators are INITIALISED left-looking: [A chunk of the
//     scratch area | B chunk of weights] pairs stream through the same shared-memory ring, the (otherwise idle)
//     particle threads split each A chunk into its TF32 hi / lo images in place, the issuer multiplies;
//   * fp32 fidelity by the 3-pass split a_hi b_hi + a_lo b_hi + a_hi b_lo (parity bar 5e-5), passes = 1 for plain TF32;
//   * what stays inside a block -- the dependencies between its own 4 degree groups -- is fp32 FMA work with one
//     thread per particle: the block's accumulators are pulled out of TMEM into registers (they BECOME the activation
//     registers), the in-block weights arrive as warp-uniform LDS.128 broadcasts from a slab the producer streamed in;
//   * spline flows (zuko NSF, the reference's default presets; template parameter RQS) differ in the univariate head only: an
//     order position owns 24 output columns (8 widths, 8 heights, 7 slopes, one zero column) instead of 2, pulled out of
//     TMEM position by position; their in-block part is 24 packed-FMA dot products per source unit, then the monotonic
//     rational-quadratic spline of flow_heads.cuh (the code the fp32-FMA sweep runs); a window holds two blocks (256
//     output columns, the widest tcgen05.mma);
//   * weights stream from L2 through two shared-memory rings (update slabs / init chunks, in-block slabs) with 1-D bulk
//     copies and mbarrier transaction counts; producer lanes, the MMA issuer and the 128 particle threads are coupled
//     by mbarriers only.
#include "common.cuh"
#include "tc_common.cuh"
#include "flow_heads.cuh"
#include <algorithm>
#include <stdlib.h>

namespace pmc {

using namespace tc;

// header / tables of tri_layout.build_tri -- keep in sync
enum { TRI_VER = 0, TRI_D, TRI_H, TRI_L, TRI_T, TRI_GSIZE, TRI_NB, TRI_NW, TRI_TSTRIDE, TRI_KCHUNK, TRI_SLOT_BYTES, TRI_DSLOT_BYTES,
       TRI_TILE_BYTES, TRI_NSTAGES, TRI_KH_TOTAL, TRI_KX_TOTAL, TRI_WS_FLOATS, TRI_OFF_BLOCKS, TRI_OFF_WINDOWS, TRI_SMEM_BYTES,
       TRI_NCOLS, TRI_CHUNK_OFF, TRI_KIND, TRI_KCHUNK_OUT, TRI_HEADER };
enum { TB_K0 = 0, TB_NST, TB_NR, TB_W, TB_KP, TB_WIN, TB_WC, TB_OC, TB_DOFF, TB_DN, TB_KS, TB_UPD_N, TB_UPD_DCOL, TB_OUT_N, TB_OUT_DCOL,
       TB_FLAGS, TB_FIELDS };
enum { TW_B0 = 0, TW_NB, TW_WP, TW_OP, TW_COL_OUT, TW_KH, TW_KX, TW_PAD, TW_FIELDS };
enum { TBF_LAST_IN_WIN = 1, TBF_LAST = 2 };

constexpr int TRI_LAYOUT_VERSION = 302;
constexpr int TRI_MAX_STAGES = 8;          // ring depth: as many slots as fit, decided by tri_layout.build_tri
constexpr int TRI_MAX_SLOTS = 12;          // + the slots a window initialisation borrows from the (then idle) A-tile area
constexpr int TRI_MAX_BLOCKS = 64;
constexpr int TRI_MAX_WINDOWS = 32;
constexpr int TRI_P_RQS = 24;             // tensor-memory columns per order position of a spline flow: 23 parameters + a zero column
constexpr int TRI_THREADS = 224;           // warps 0-3 particles, 4 ring producer, 5 MMA issuer, 6 in-block slab producer
constexpr int G = 4;                       // order positions per block
constexpr float TRI_LOG_SLOPE = -6.90775527898213705205f;

struct TriParams {
  const float* packed;
  const int* tables;       // device copy of the block / window tables (tri_layout.build_tri meta from TRI_HEADER on)
  const float* in;
  float* out;
  float* ladj;
  float* ws;               // scratch: gridDim.x * ws_floats
  long long n;
  long long ws_floats;
  int D, T, NB, NW, tstride, chunk_off, passes, stages, extra, kc, kc_out, kh_total, kx_total;   // kc / kc_out: K extent of an init chunk (hidden layers / outputs)
  uint32_t slot_bytes, dslot_bytes, tile_bytes;
};

template <int NR>
struct TriShape {
  static constexpr int E = NR - 4;                       // extra units per group behind the four regular ones (0, 1, 2)
  static constexpr int W = 4 * G + E * G;                // slots of a block: 16, 20, 24
  static constexpr int KP = (W + 7) / 8 * 8;             // K extent of the block's A tiles: 16, 24, 24
  static constexpr int NRV = (NR == 4 ? 1 : 2);          // float4 per bias vector
  static constexpr int Q1 = (2 * NR + 3) / 4;            // float4 per x pair of the layer-1 weights
  static constexpr int XV = (E == 0 ? 0 : (E == 1 ? 2 : 3));   // float4 of extra-source weights per source group
  static constexpr int GRPV = NR + XV;                   // float4 per source group of the layer-2/3 weights
  static constexpr int OGV = 2 + (E ? 1 : 0);            // float4 per source group of the output weights
};

__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ float2 ffma2(const float2 a, const float2 b, const float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float get4(const float4& v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w)); }
// k-th float2 / float of an array of float4 (compile-time k)
template <int N>
__device__ __forceinline__ float2 pair_of(const float4 (&v)[N], int k) {
  return (k & 1) ? make_float2(v[k >> 1].z, v[k >> 1].w) : make_float2(v[k >> 1].x, v[k >> 1].y);
}
template <int N>
__device__ __forceinline__ float elem_of(const float4 (&v)[N], int k) { return get4(v[k >> 2], k & 3); }
template <int N>
__device__ __forceinline__ void load_v(const float4* __restrict__ p, float4 (&v)[N]) {
#pragma unroll
  for (int i = 0; i < N; ++i) v[i] = p[i];
}

__device__ __forceinline__ void tmem_ld4_async(uint32_t taddr, uint32_t (&r)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld8_async(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_fence4(uint32_t (&r)[4]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]) :: "memory");
}
__device__ __forceinline__ void tmem_ld_fence8(uint32_t (&r)[8]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7])
               :: "memory");
}

// the registers of one block and one hidden layer: first the accumulators pulled out of tensor memory, then -- unit by
// unit -- the activations.  Regular units as source pairs (packed FMAs); the extra unit(s) of group j in e[j] (E = 1: .x).
struct TriActs {
  float2 r[2 * G];
  float2 e[G];
};

__device__ __forceinline__ void acts_zero(TriActs& a) {
#pragma unroll
  for (int i = 0; i < 2 * G; ++i) a.r[i] = make_float2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < G; ++i) a.e[i] = make_float2(0.f, 0.f);
}

struct TriRaw {            // raw TMEM images of one layer's block segment
  uint32_t r[4 * G];
  uint32_t x[2 * G];
};
template <int NR>
__device__ __forceinline__ void raw_load(const uint32_t taddr, TriRaw& t) {
  tmem_ld16_async(taddr, t.r);
  if constexpr (NR == 5) {
    uint32_t(&x4)[4] = reinterpret_cast<uint32_t(&)[4]>(t.x);
    tmem_ld4_async(taddr + 16, x4);
  }
  if constexpr (NR == 6) tmem_ld8_async(taddr + 16, t.x);
}
template <int NR>
__device__ __forceinline__ void raw_take(TriRaw& t, TriActs& a) {
  tmem_ld_fence16(t.r);
  if constexpr (NR == 5) {
    uint32_t(&x4)[4] = reinterpret_cast<uint32_t(&)[4]>(t.x);
    tmem_ld_fence4(x4);
  }
  if constexpr (NR == 6) tmem_ld_fence8(t.x);
#pragma unroll
  for (int i = 0; i < 2 * G; ++i) a.r[i] = make_float2(__uint_as_float(t.r[2 * i]), __uint_as_float(t.r[2 * i + 1]));
#pragma unroll
  for (int i = 0; i < G; ++i) {
    if constexpr (NR == 4) a.e[i] = make_float2(0.f, 0.f);
    if constexpr (NR == 5) a.e[i] = make_float2(__uint_as_float(t.x[i]), 0.f);
    if constexpr (NR == 6) a.e[i] = make_float2(__uint_as_float(t.x[2 * i]), __uint_as_float(t.x[2 * i + 1]));
  }
}

// accumulator / activation of unit s of in-block group J
template <int J>
__device__ __forceinline__ float unit_of(const TriActs& a, const int s) {
  return s == 0 ? a.r[2 * J].x : (s == 1 ? a.r[2 * J].y : (s == 2 ? a.r[2 * J + 1].x : (s == 3 ? a.r[2 * J + 1].y : (s == 4 ? a.e[J].x : a.e[J].y))));
}
template <int NR, int J>
__device__ __forceinline__ void relu_into(TriActs& dst, const float2 (&acc)[NR]) {
  dst.r[2 * J] = make_float2(fmaxf(acc[0].x + acc[0].y, 0.f), fmaxf(acc[1].x + acc[1].y, 0.f));
  dst.r[2 * J + 1] = make_float2(fmaxf(acc[2].x + acc[2].y, 0.f), fmaxf(acc[3].x + acc[3].y, 0.f));
  if constexpr (NR == 5) dst.e[J] = make_float2(fmaxf(acc[4].x + acc[4].y, 0.f), 0.f);
  if constexpr (NR == 6) dst.e[J] = make_float2(fmaxf(acc[4].x + acc[4].y, 0.f), fmaxf(acc[5].x + acc[5].y, 0.f));
}

// one residual hidden layer (2 or 3) of in-block group J: dst[own] = relu(src[own] + acc + bias + sum over groups 0..J)
template <int NR, int J>
__device__ __forceinline__ void tri_hidden(const float4* __restrict__ q, int& off, const TriActs& src, TriActs& dst) {
  using S = TriShape<NR>;
  float4 bias[S::NRV];
  load_v(q + off, bias);
  off += S::NRV;
  float2 acc[NR];
#pragma unroll
  for (int s = 0; s < NR; ++s) acc[s] = make_float2(unit_of<J>(dst, s) + elem_of(bias, s), unit_of<J>(src, s));
#pragma unroll
  for (int c = 0; c <= J; ++c) {
    float4 w[S::GRPV];
    load_v(q + off, w);
    off += S::GRPV;
#pragma unroll
    for (int s = 0; s < NR; ++s) {
      acc[s] = ffma2(pair_of(w, s), src.r[2 * c], acc[s]);
      acc[s] = ffma2(pair_of(w, NR + s), src.r[2 * c + 1], acc[s]);
    }
    if constexpr (NR == 5) {
      const float4 wx[2] = {w[NR], w[NR + 1]};
#pragma unroll
      for (int s = 0; s < NR; ++s) acc[s].x = fmaf(elem_of(wx, s), src.e[c].x, acc[s].x);
    }
    if constexpr (NR == 6) {
      const float4 wx[3] = {w[NR], w[NR + 1], w[NR + 2]};
#pragma unroll
      for (int s = 0; s < NR; ++s) acc[s] = ffma2(pair_of(wx, s), src.e[c], acc[s]);
    }
  }
  relu_into<NR, J>(dst, acc);
}

// spline head of order position J: the 24 parameter columns of the position (accumulators in tensor memory: everything
// earlier blocks contribute) + bias + the contribution of the block's own earlier groups, then the spline itself.
// In-block slab: 6 float4 of bias, then per source unit (4 regular units of a group, then its extras) 6 float4 = the
// unit's weight into parameters 0..23.
template <int NR, bool INV, int J>
__device__ __forceinline__ float rqs_head(const float4* __restrict__ q, int& off, const TriActs& a3, const uint32_t ocol,
                                          const bool have_acc, const float v, float& lj) {
  constexpr int E = TriShape<NR>::E;
  float2 ph[12];
  if (have_acc) {
    uint32_t r0[8], r1[8], r2[8];
    tmem_ld8_async(ocol + TRI_P_RQS * J, r0);
    tmem_ld8_async(ocol + TRI_P_RQS * J + 8, r1);
    tmem_ld8_async(ocol + TRI_P_RQS * J + 16, r2);
    tmem_ld_fence8(r0);
    tmem_ld_fence8(r1);
    tmem_ld_fence8(r2);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      ph[i] = make_float2(__uint_as_float(r0[2 * i]), __uint_as_float(r0[2 * i + 1]));
      ph[4 + i] = make_float2(__uint_as_float(r1[2 * i]), __uint_as_float(r1[2 * i + 1]));
      ph[8 + i] = make_float2(__uint_as_float(r2[2 * i]), __uint_as_float(r2[2 * i + 1]));
    }
  } else {
#pragma unroll
    for (int i = 0; i < 12; ++i) ph[i] = make_float2(0.f, 0.f);
  }
  {
    float4 b[6];
    load_v(q + off, b);
    off += 6;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      ph[2 * i].x += b[i].x; ph[2 * i].y += b[i].y;
      ph[2 * i + 1].x += b[i].z; ph[2 * i + 1].y += b[i].w;
    }
  }
  auto add_unit = [&](const float a) {
    float4 w[6];
    load_v(q + off, w);
    off += 6;
    const float2 aa = make_float2(a, a);
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      ph[2 * i] = ffma2(make_float2(w[i].x, w[i].y), aa, ph[2 * i]);
      ph[2 * i + 1] = ffma2(make_float2(w[i].z, w[i].w), aa, ph[2 * i + 1]);
    }
  };
#pragma unroll
  for (int c = 0; c < J; ++c) {
    add_unit(a3.r[2 * c].x);
    add_unit(a3.r[2 * c].y);
    add_unit(a3.r[2 * c + 1].x);
    add_unit(a3.r[2 * c + 1].y);
    if constexpr (E >= 1) add_unit(a3.e[c].x);
    if constexpr (E == 2) add_unit(a3.e[c].y);
  }
  float phi[24];
#pragma unroll
  for (int i = 0; i < 12; ++i) { phi[2 * i] = ph[i].x; phi[2 * i + 1] = ph[i].y; }
#ifdef PMC_TRI_RQS_REFERENCE_HEAD
  return Rqs::apply(phi, v, INV, lj);            // diagnostic build: the reference's operation order (tests/nsf_truth.py)
#else
  return RqsLean::apply<INV>(phi, v, lj);
#endif
}

// order position J of a block: output -> affine map / spline -> the degree group's three hidden layers
template <int NR, bool INV, bool RQS, int J>
__device__ __forceinline__ void tri_stage(const float4* __restrict__ q, int& off, TriActs& a1, TriActs& a2, TriActs& a3,
                                          const uint32_t (&o)[2 * G], const uint32_t ocol, const bool have_acc, float2 (&xbp)[G / 2],
                                          const float (&y)[G], float& ladj, float* __restrict__ out_row, const bool valid,
                                          const int kstep, const int feat0) {
  using S = TriShape<NR>;
  if constexpr (RQS) {
    float lj;
    const float res = rqs_head<NR, INV, J>(q, off, a3, ocol, have_acc, y[J], lj);
    const float xk = INV ? res : y[J];
    ladj = INV ? ladj - lj : ladj + lj;
    if (valid) out_row[feat0 + J * kstep] = res;
    if (J & 1) xbp[J >> 1].y = xk; else xbp[J >> 1].x = xk;
  } else {
    const float4 ob = q[off];
    off += 1;
    float2 shf = make_float2(__uint_as_float(o[2 * J]) + ob.x, 0.f), srw = make_float2(__uint_as_float(o[2 * J + 1]) + ob.y, 0.f);
#pragma unroll
    for (int c = 0; c < J; ++c) {
      float4 w[S::OGV];
      load_v(q + off, w);
      off += S::OGV;
      shf = ffma2(make_float2(w[0].x, w[0].y), a3.r[2 * c], shf);
      srw = ffma2(make_float2(w[0].z, w[0].w), a3.r[2 * c], srw);
      shf = ffma2(make_float2(w[1].x, w[1].y), a3.r[2 * c + 1], shf);
      srw = ffma2(make_float2(w[1].z, w[1].w), a3.r[2 * c + 1], srw);
      if constexpr (NR == 5) {
        shf.x = fmaf(w[2].x, a3.e[c].x, shf.x);
        srw.x = fmaf(w[2].y, a3.e[c].x, srw.x);
      }
      if constexpr (NR == 6) {
        shf = ffma2(make_float2(w[2].x, w[2].y), a3.e[c], shf);
        srw = ffma2(make_float2(w[2].z, w[2].w), a3.e[c], srw);
      }
    }
    const float shift = shf.x + shf.y, sraw = srw.x + srw.y;
    const float ls = sraw / (1.0f + fabsf(sraw / TRI_LOG_SLOPE));
    float xk, res;
    if (INV) { xk = (y[J] - shift) * expf(-ls); res = xk; ladj -= ls; }
    else { xk = y[J]; res = fmaf(xk, expf(ls), shift); ladj += ls; }
    if (valid) out_row[feat0 + J * kstep] = res;
    if (J & 1) xbp[J >> 1].y = xk; else xbp[J >> 1].x = xk;
  }
  {   // hidden layer 1: inputs x of order positions k0 .. k0 + J (pairs; the unborn one of the last pair is 0)
    float4 bias[S::NRV];
    load_v(q + off, bias);
    off += S::NRV;
    float2 acc[NR];
#pragma unroll
    for (int s = 0; s < NR; ++s) acc[s] = make_float2(unit_of<J>(a1, s) + elem_of(bias, s), 0.f);
#pragma unroll
    for (int c = 0; c <= (J >> 1); ++c) {
      float4 w[S::Q1];
      load_v(q + off, w);
      off += S::Q1;
#pragma unroll
      for (int s = 0; s < NR; ++s) acc[s] = ffma2(pair_of(w, s), xbp[c], acc[s]);
    }
    relu_into<NR, J>(a1, acc);
  }
  tri_hidden<NR, J>(q, off, a1, a2);
  tri_hidden<NR, J>(q, off, a2, a3);
}

// order positions J .. G-1 of a block (compile-time recursion: every register index stays static)
template <int NR, bool INV, bool RQS, int J>
__device__ __forceinline__ void tri_stages(const float4* __restrict__ q, int& off, const int nst, TriActs& a1, TriActs& a2,
                                           TriActs& a3, const uint32_t (&o)[2 * G], const uint32_t ocol, const bool have_acc,
                                           float2 (&xbp)[G / 2], const float (&y)[G], float& ladj,
                                           float* __restrict__ out_row, const bool valid, const int kstep, const int feat0) {
  if (J >= nst) return;
  tri_stage<NR, INV, RQS, J>(q, off, a1, a2, a3, o, ocol, have_acc, xbp, y, ladj, out_row, valid, kstep, feat0);
  if constexpr (J + 1 < G) tri_stages<NR, INV, RQS, J + 1>(q, off, nst, a1, a2, a3, o, ocol, have_acc, xbp, y, ladj, out_row, valid, kstep, feat0);
}

struct TriShared {
  uint64_t bfull[TRI_MAX_SLOTS], bempty[TRI_MAX_SLOTS], asplit[TRI_MAX_SLOTS], dfull[2], dempty[2], a_ready, acc_ready, s_ready;
  uint32_t tmem_slot;
  int blocks[TRI_MAX_BLOCKS][TB_FIELDS];
  int wins[TRI_MAX_WINDOWS][TW_FIELDS];
  alignas(16) uint32_t issue[TRI_MAX_BLOCKS][4][8];   // per block and update op: what the MMA issuer needs, precomputed (kernel prologue)
};

// A tile of one layer: [128 rows x KP] as 16-byte chunks of 4 consecutive k, chunk (r, c) at c * 2048 + r * 16.
// `lo` == nullptr: plain fp32 image only (the scratch area in global memory).
template <int NR, bool SPLIT>
__device__ __forceinline__ void store_tile(unsigned char* hi, unsigned char* lo, const int r, const TriActs& a) {
  auto put = [&](const int c, const float4 v) {
    if constexpr (SPLIT) {
      float4 h, l;
      split_tf32(v.x, h.x, l.x); split_tf32(v.y, h.y, l.y); split_tf32(v.z, h.z, l.z); split_tf32(v.w, h.w, l.w);
      *reinterpret_cast<float4*>(hi + c * 2048 + r * 16) = h;
      *reinterpret_cast<float4*>(lo + c * 2048 + r * 16) = l;
    } else {
      *reinterpret_cast<float4*>(hi + c * 2048 + r * 16) = v;
    }
  };
#pragma unroll
  for (int c = 0; c < G; ++c) put(c, make_float4(a.r[2 * c].x, a.r[2 * c].y, a.r[2 * c + 1].x, a.r[2 * c + 1].y));
  if constexpr (NR == 5) {       // slots 16..19 = the extra unit of groups 0..3, K padded to 24
    put(G, make_float4(a.e[0].x, a.e[1].x, a.e[2].x, a.e[3].x));
    put(G + 1, make_float4(0.f, 0.f, 0.f, 0.f));
  }
  if constexpr (NR == 6) {       // slots 16 + 2j, 17 + 2j = the two extra units of group j
    put(G, make_float4(a.e[0].x, a.e[0].y, a.e[1].x, a.e[1].y));
    put(G + 1, make_float4(a.e[2].x, a.e[2].y, a.e[3].x, a.e[3].y));
  }
}

// one block of the substitution on the 128 particle threads
template <int NR, bool INV, bool RQS>
__device__ __forceinline__ void run_block(const TriParams& p, const int bi, const int t, const uint32_t lane_base, unsigned char* smem,
                                          const float4* __restrict__ slab, TriShared& sh, uint32_t& n_groups, const int row_in_tile,
                                          float* out_row, const bool valid, float& ladj, const float (&y)[G], float* ws) {
  const int* B = sh.blocks[bi];
  const int k0 = B[TB_K0], nst = B[TB_NST], flags = B[TB_FLAGS], win = B[TB_WIN];
  const bool rev = (t & 1);
  const int feat0 = rev ? (p.D - 1 - k0) : k0;
  const int kstep = rev ? -1 : 1;
  TriActs a1, a2, a3;
  uint32_t o[2 * G];
  const uint32_t ocol = lane_base + (uint32_t)(sh.wins[win][TW_COL_OUT] + B[TB_OC]);     // this block's output columns
  if (bi == 0) {
    acts_zero(a1); acts_zero(a2); acts_zero(a3);
#pragma unroll
    for (int i = 0; i < 2 * G; ++i) o[i] = 0u;
  } else {
    // accumulators of this block are complete once the previous block's update (or this window's initialisation) has landed
    mbar_wait(&sh.acc_ready, (n_groups - 1) & 1);
    tc_fence_after();
    const int Wp = sh.wins[win][TW_WP], wc = B[TB_WC];
    TriRaw t1, t2, t3;
    raw_load<NR>(lane_base + wc, t1);
    raw_load<NR>(lane_base + Wp + wc, t2);
    raw_load<NR>(lane_base + 2 * Wp + wc, t3);
    if constexpr (!RQS) tmem_ld8_async(ocol, o);                   // spline heads pull their 24 columns position by position
    raw_take<NR>(t1, a1);
    raw_take<NR>(t2, a2);
    raw_take<NR>(t3, a3);
    if constexpr (!RQS) tmem_ld_fence8(o);
    else {
#pragma unroll
      for (int i = 0; i < 2 * G; ++i) o[i] = 0u;
    }
  }
  float2 xbp[G / 2];
#pragma unroll
  for (int i = 0; i < G / 2; ++i) xbp[i] = make_float2(0.f, 0.f);
  int off = 0;
  tri_stages<NR, INV, RQS, 0>(slab, off, nst, a1, a2, a3, o, ocol, bi != 0, xbp, y, ladj, out_row, valid, kstep, feat0);
  if (!(flags & TBF_LAST)) {
    const float4 xv = make_float4(xbp[0].x, xbp[0].y, xbp[1].x, xbp[1].y), z = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.NW > 1 && win + 1 < p.NW) {
      // append to the scratch area (plain fp32; later windows read it back chunk by chunk)
      unsigned char* g = reinterpret_cast<unsigned char*>(ws);
      *reinterpret_cast<float4*>(g + (size_t)(2 * bi) * 2048 + row_in_tile * 16) = xv;
      *reinterpret_cast<float4*>(g + (size_t)(2 * bi + 1) * 2048 + row_in_tile * 16) = z;
      unsigned char* h = g + (size_t)p.kx_total * 512 + (size_t)(B[TB_KS] >> 2) * 2048;
      store_tile<NR, false>(h, nullptr, row_in_tile, a1);
      store_tile<NR, false>(h + (size_t)p.kh_total * 512, nullptr, row_in_tile, a2);
      store_tile<NR, false>(h + (size_t)p.kh_total * 1024, nullptr, row_in_tile, a3);
    }
    if (!(flags & TBF_LAST_IN_WIN)) {
      // hand the block's activations to the tensor core (the update group that read the tiles last has completed:
      // its commit is what released this block's accumulators)
      unsigned char* tiles = smem;
      store_tile<NR, true>(tiles, tiles + p.tile_bytes, row_in_tile, a1);
      store_tile<NR, true>(tiles + 2 * p.tile_bytes, tiles + 3 * p.tile_bytes, row_in_tile, a2);
      store_tile<NR, true>(tiles + 4 * p.tile_bytes, tiles + 5 * p.tile_bytes, row_in_tile, a3);
      unsigned char* xt = tiles + 6 * p.tile_bytes;
      float4 h, l;
      split_tf32(xv.x, h.x, l.x); split_tf32(xv.y, h.y, l.y); split_tf32(xv.z, h.z, l.z); split_tf32(xv.w, h.w, l.w);
      *reinterpret_cast<float4*>(xt + row_in_tile * 16) = h;
      *reinterpret_cast<float4*>(xt + 4096 + row_in_tile * 16) = l;
      *reinterpret_cast<float4*>(xt + 2048 + row_in_tile * 16) = z;              // K = 8: order positions 4..7 do not exist
      *reinterpret_cast<float4*>(xt + 4096 + 2048 + row_in_tile * 16) = z;
    }
    // generic-proxy stores -> async proxy (tcgen05.mma operand reads, bulk copies of the scratch area): the global half of the
    // fence only when this block wrote scratch
    if (p.NW > 1 && win + 1 < p.NW) fence_proxy_async_all(); else fence_proxy_async();
    tc_fence_before();
    mbar_arrive(&sh.a_ready);
    if (flags & TBF_LAST_IN_WIN) mbar_arrive(&sh.s_ready);
    ++n_groups;
  }
}

// The operand ring as every role walks it.  Update slabs cycle through the `stages` slots behind the tiles; the chunks of a
// window initialisation cycle through stages + extra slots -- the extra ones lie in the A-tile area, which nothing reads
// between the last update of a window and the first block of the next.  A slot's mbarriers complete once per use whichever
// sequence used it, so every role keeps ONE parity bit per slot for the barrier it waits on.
struct Ring {
  uint32_t su = 0, si = 0;        // next slot of the update / initialisation sequence
  uint32_t par = 0;               // per slot: parity of the next completion this role waits for
  uint32_t used = 0;              // producer: slots filled at least once (their `empty` barrier has something to wait for)
  __device__ __forceinline__ uint32_t next_update(const uint32_t stages) { const uint32_t s = su; su = (su + 1 == stages) ? 0 : su + 1; return s; }
  __device__ __forceinline__ uint32_t next_init(const uint32_t slots) { const uint32_t s = si; si = (si + 1 == slots) ? 0 : si + 1; return s; }
  __device__ __forceinline__ void wait(uint64_t* bars, const uint32_t s) { mbar_wait(bars + s, (par >> s) & 1u); par ^= 1u << s; }
  __device__ __forceinline__ void pass(const uint32_t s) { par ^= 1u << s; }                  // a use this role does not wait for
  __device__ __forceinline__ void acquire(uint64_t* empty, const uint32_t s) {               // producer: the slot's previous use is over
    if ((used >> s) & 1u) wait(empty, s);
    used |= 1u << s;
  }
};

template <bool INV, bool RQS>
__global__ void __launch_bounds__(TRI_THREADS, 1)
made_sweep_tri_kernel(const __grid_constant__ TriParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ TriShared sh;
  // shared-memory map: [3 layers x (hi, lo) A tiles][x tile hi 4 KB][x tile lo 4 KB][ring][in-block slab ring]
  unsigned char* ring = smem + 6 * p.tile_bytes + 8192;
  unsigned char* dring = ring + (size_t)p.stages * p.slot_bytes;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // programmatic dependent launch: the next kernel of the MCMC step may be scheduled now; this kernel's own set-up
  // (tables -- written once when the layout is built --, descriptor words, mbarriers, the TMEM allocation) overlaps the
  // tail of the previous kernel, and pdl_wait() below comes before the first read of anything a kernel produces
  pdl_launch_dependents();
  for (int i = threadIdx.x; i < p.NB * TB_FIELDS; i += blockDim.x) (&sh.blocks[0][0])[i] = p.tables[i];
  for (int i = threadIdx.x; i < p.NW * TW_FIELDS; i += blockDim.x) (&sh.wins[0][0])[i] = p.tables[p.NB * TB_FIELDS + i];
  __syncthreads();
  for (int c = threadIdx.x; c < p.NB * 4; c += blockDim.x) {
    // the MMA issuer is ONE thread whose instruction stream sits on the critical path of every block: hand it
    // ready-made descriptor words.  K-major no-swizzle descriptors: [0,14) address >> 4, [16,30) LBO >> 4,
    // [32,46) SBO >> 4 = 8, bit 46 version; A tiles: LBO = 2048 (128 rows x 16 B), B slabs: LBO = N x 16.
    const int bi = c >> 2, op = c & 3;                  // op 0: layer 1 (A = x tile), 1, 2: layers 2, 3, 3: outputs
    const int* B = sh.blocks[bi];
    const int* Wn = sh.wins[B[TB_WIN]];
    const int N = op < 3 ? B[TB_UPD_N] : B[TB_OUT_N];
    const int nks = op == 0 ? 1 : B[TB_KP] / 8;
    const int dcol = op < 3 ? op * Wn[TW_WP] + B[TB_UPD_DCOL] : Wn[TW_COL_OUT] + B[TB_OUT_DCOL];
    const uint32_t tiles_addr = smem_u32(smem);
    const uint32_t a_base = op == 0 ? tiles_addr + 6 * p.tile_bytes : tiles_addr + (uint32_t)(2 * (op - 1)) * p.tile_bytes;
    const uint32_t a_lo = a_base + (op == 0 ? 4096u : p.tile_bytes);
    uint32_t* w = sh.issue[bi][op];
    w[0] = ((a_base & 0x3FFFF) >> 4) | ((2048u >> 4) << 16);
    w[1] = ((a_lo & 0x3FFFF) >> 4) | ((2048u >> 4) << 16);
    w[2] = (((uint32_t)N * 16u) >> 4) << 16;
    w[3] = (uint32_t)(nks * N * 32) >> 4;
    w[4] = ((uint32_t)N * 32u) >> 4;
    w[5] = idesc_tf32(128, N);
    w[6] = (uint32_t)dcol;
    w[7] = (uint32_t)nks | ((uint32_t)(bi == 0 ? 1 : 0) << 8);
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < TRI_MAX_SLOTS; ++i) { mbar_init(sh.bfull + i, 1); mbar_init(sh.bempty + i, 1); mbar_init(sh.asplit + i, 128); }
    for (int i = 0; i < 2; ++i) { mbar_init(sh.dfull + i, 1); mbar_init(sh.dempty + i, 128); }
    mbar_init(&sh.a_ready, 128);
    mbar_init(&sh.acc_ready, 1);
    mbar_init(&sh.s_ready, 128);
    mbar_fence_init();
  }
  if (warp == 4) tmem_alloc<512>(&sh.tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();
  const uint32_t tm = sh.tmem_slot;
  const long long n_tiles = (p.n + 127) / 128;
  float* ws = p.ws + (size_t)blockIdx.x * (size_t)p.ws_floats;
  const uint32_t stages = (uint32_t)p.stages, init_slots = (uint32_t)(p.stages + p.extra);
  auto slot_ptr = [&](const uint32_t sl) -> unsigned char* { return sl < stages ? ring + (size_t)sl * p.slot_bytes : smem + (size_t)(sl - stages) * p.slot_bytes; };

  if (warp == 4) {
    // ---------------- producer 1: update slabs (B operands) and window-initialisation chunks (A from scratch + B) ----------------
    if (lane == 0) {
      Ring rg;
      uint32_t n_tr = 0;
      for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int tt = 0; tt < p.T; ++tt) {
          const int t = INV ? p.T - 1 - tt : tt;
          const float* src = p.packed + (size_t)t * p.tstride + p.chunk_off;
          for (int bi = 0; bi < p.NB; ++bi) {
            const int* B = sh.blocks[bi];
            const int flags = B[TB_FLAGS];
            if (!(flags & TBF_LAST_IN_WIN)) {
              for (int op = 0; op < 4; ++op) {
                const uint32_t N = (uint32_t)(op < 3 ? B[TB_UPD_N] : B[TB_OUT_N]), K = (uint32_t)(op == 0 ? 8 : B[TB_KP]);
                const uint32_t bytes = N * K * 8u;
                const uint32_t sl = rg.next_update(stages);
                rg.acquire(sh.bempty, sl);
                mbar_expect_tx(sh.bfull + sl, bytes);
                bulk_g2s(slot_ptr(sl), src, bytes, sh.bfull + sl);
                src += N * K * 2u;
              }
            } else if (!(flags & TBF_LAST)) {
              const int* Wn = sh.wins[B[TB_WIN] + 1];
              mbar_wait(&sh.s_ready, n_tr & 1);          // every block of the finished windows is in the scratch area
              ++n_tr;
              fence_proxy_async_all();
              for (int op = 0; op < 4; ++op) {
                const int ktot = op == 0 ? Wn[TW_KX] : Wn[TW_KH];
                const uint32_t N = (uint32_t)(op < 3 ? Wn[TW_WP] : Wn[TW_OP]);
                const float* a_src = ws + (op == 0 ? (size_t)0 : ((size_t)p.kx_total + (size_t)(op - 1) * p.kh_total) * 128);
                const int kc = op < 3 ? p.kc : p.kc_out;
                for (int k = 0; k < ktot; k += kc) {
                  const uint32_t ke = (uint32_t)min(kc, ktot - k);
                  const uint32_t a_bytes = ke * 512u, b_bytes = N * ke * 8u;
                  const uint32_t sl = rg.next_init(init_slots);
                  rg.acquire(sh.bempty, sl);
                  unsigned char* dst = slot_ptr(sl);
                  mbar_expect_tx(sh.bfull + sl, a_bytes + b_bytes);
                  bulk_g2s(dst, a_src + (size_t)k * 128, a_bytes, sh.bfull + sl);
                  bulk_g2s(dst + 2 * a_bytes, src, b_bytes, sh.bfull + sl);
                  src += N * ke * 2u;
                }
              }
            }
          }
        }
      }
    }
  } else if (warp == 6) {
    // ---------------- producer 2: in-block (FFMA) weight slabs ----------------
    if (lane == 0) {
      uint32_t it = 0;
      for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int tt = 0; tt < p.T; ++tt) {
          const int t = INV ? p.T - 1 - tt : tt;
          const float* base = p.packed + (size_t)t * p.tstride;
          for (int bi = 0; bi < p.NB; ++bi) {
            const uint32_t slot = it & 1;
            if (it >= 2) mbar_wait(sh.dempty + slot, ((it >> 1) - 1) & 1);
            const uint32_t bytes = (uint32_t)sh.blocks[bi][TB_DN] * 4u;
            mbar_expect_tx(sh.dfull + slot, bytes);
            bulk_g2s(dring + (size_t)slot * p.dslot_bytes, base + sh.blocks[bi][TB_DOFF], bytes, sh.dfull + slot);
            ++it;
          }
        }
      }
    }
  } else if (warp == 5) {
    // ---------------- issuer: one thread drives the tensor core ----------------
    if (lane == 0) {
      Ring rg;
      uint32_t n_blk = 0, split_phase = 0;              // split_phase: one parity bit per ring slot (asplit completes only on init chunks)
      const uint32_t ring16 = (smem_u32(ring) & 0x3FFFF) >> 4, slot16 = p.slot_bytes >> 4, tiles16 = (smem_u32(smem) & 0x3FFFF) >> 4;
      const uint32_t desc_top = (128u >> 4) | (1u << 14);              // high word: SBO = 128 bytes, descriptor version 1
      const bool split = p.passes > 1;
      auto mma_group = [&](const uint32_t d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, const uint32_t b_lo_off, const uint32_t b_step,
                           const uint32_t idesc, const uint32_t nks, uint32_t accum) {
        for (uint32_t ks = 0; ks < nks && p.passes > 0; ++ks) {
          const uint64_t dah = ((uint64_t)desc_top << 32) | a_hi, dbh = ((uint64_t)desc_top << 32) | b_hi;
          mma_tf32_ss(d, dah, dbh, idesc, accum);
          accum = 1u;
          if (split) {
            const uint64_t dal = ((uint64_t)desc_top << 32) | a_lo, dbl = ((uint64_t)desc_top << 32) | (b_hi + b_lo_off);
            mma_tf32_ss(d, dal, dbh, idesc, 1u);
            mma_tf32_ss(d, dah, dbl, idesc, 1u);
          }
          a_hi += 256u; a_lo += 256u; b_hi += b_step;
        }
      };
      for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int tt = 0; tt < p.T; ++tt) {
          for (int bi = 0; bi + 1 < p.NB; ++bi) {
            const int* B = sh.blocks[bi];
            mbar_wait(&sh.a_ready, n_blk & 1);
            ++n_blk;
            tc_fence_after();
            if (!(B[TB_FLAGS] & TBF_LAST_IN_WIN)) {
              for (int op = 0; op < 4; ++op) {
                const uint4 r0 = *reinterpret_cast<const uint4*>(sh.issue[bi][op]);
                const uint4 r1 = *reinterpret_cast<const uint4*>(sh.issue[bi][op] + 4);
                const uint32_t sl = rg.next_update(stages);
                rg.wait(sh.bfull, sl);
                tc_fence_after();
                mma_group(tm + r1.z, r0.x, r0.y, r0.z + ring16 + sl * slot16, r0.w, r1.x, r1.y, r1.w & 0xffu, (r1.w & 0x100u) ? 0u : 1u);
                mma_commit(sh.bempty + sl);
              }
            } else {
              const int* Wn = sh.wins[B[TB_WIN] + 1];
              for (int op = 0; op < 4; ++op) {
                const int ktot = op == 0 ? Wn[TW_KX] : Wn[TW_KH];
                const uint32_t N = (uint32_t)(op < 3 ? Wn[TW_WP] : Wn[TW_OP]);
                const uint32_t d = tm + (uint32_t)(op < 3 ? op * Wn[TW_WP] : Wn[TW_COL_OUT]);
                const uint32_t idesc = idesc_tf32(128, (int)N);
                const int kc = op < 3 ? p.kc : p.kc_out;
                for (int k = 0; k < ktot; k += kc) {
                  const uint32_t ke = (uint32_t)min(kc, ktot - k);
                  const uint32_t sl = rg.next_init(init_slots);
                  rg.pass(sl);                                   // the chunk's arrival is observed by the splitting threads
                  const uint32_t base16 = sl < stages ? ring16 + sl * slot16 : tiles16 + (sl - stages) * slot16;
                  mbar_wait(sh.asplit + sl, (split_phase >> sl) & 1u);
                  split_phase ^= 1u << sl;
                  tc_fence_after();
                  const uint32_t a_hi = base16 | ((2048u >> 4) << 16);
                  mma_group(d, a_hi, a_hi + ke * 32u, (base16 + ke * 64u) | (N << 16), (N * ke * 4u) >> 4, N * 2u, idesc, ke >> 3, k == 0 ? 0u : 1u);
                  mma_commit(sh.bempty + sl);
                }
              }
            }
            mma_commit(&sh.acc_ready);
          }
        }
      }
    }
  } else if (warp < 4) {
    // ---------------- substitution: thread = particle row = TMEM lane ----------------
    const int row_in_tile = threadIdx.x;
    const uint32_t lane_base = tm + ((uint32_t)(warp * 32) << 16);
    uint32_t n_groups = 0, dit = 0;
    Ring rg;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const long long row = tile * 128 + row_in_tile;
      const bool valid = row < p.n;
      // the particle's working vector lives in its row of `out` (read and rewritten in place by this thread only);
      // rows past the end of the batch run the same instruction stream on zeros (the TMEM loads are warp-collective)
      // with every global access predicated off
      float* out_row = p.out + (valid ? row : 0) * p.D;
      if (valid && p.in != p.out) {
        const float* in_row = p.in + row * p.D;
        if ((p.D & 3) == 0) {
          for (int c = 0; c < p.D; c += 4) *reinterpret_cast<float4*>(out_row + c) = *reinterpret_cast<const float4*>(in_row + c);
        } else {
          for (int c = 0; c < p.D; ++c) out_row[c] = in_row[c];
        }
      }
      float ladj = 0.f;
      float y[G];
      auto load_y = [&](const int t, const int bi, float (&dst)[G]) {
        const int k0 = sh.blocks[bi][TB_K0], nst = sh.blocks[bi][TB_NST];
        const bool rev = (t & 1);
        const int feat0 = rev ? (p.D - 1 - k0) : k0, kstep = rev ? -1 : 1;
#pragma unroll
        for (int j = 0; j < G; ++j) dst[j] = (valid && j < nst) ? out_row[feat0 + j * kstep] : 0.f;
      };
      load_y(INV ? p.T - 1 : 0, 0, y);
      for (int tt = 0; tt < p.T; ++tt) {
        const int t = INV ? p.T - 1 - tt : tt;
        for (int bi = 0; bi < p.NB; ++bi) {
          const uint32_t slot = dit & 1;
          mbar_wait(sh.dfull + slot, (dit >> 1) & 1);
          const float4* dslab = reinterpret_cast<const float4*>(dring + (size_t)slot * p.dslot_bytes);
          const int nr = sh.blocks[bi][TB_NR], flags = sh.blocks[bi][TB_FLAGS];
          if (nr == 4) run_block<4, INV, RQS>(p, bi, t, lane_base, smem, dslab, sh, n_groups, row_in_tile, out_row, valid, ladj, y, ws);
          else if (nr == 5) run_block<5, INV, RQS>(p, bi, t, lane_base, smem, dslab, sh, n_groups, row_in_tile, out_row, valid, ladj, y, ws);
          else run_block<6, INV, RQS>(p, bi, t, lane_base, smem, dslab, sh, n_groups, row_in_tile, out_row, valid, ladj, y, ws);
          mbar_arrive(sh.dempty + slot);
          ++dit;
          // the next block's inputs (this thread's own earlier stores; L2 latency hides behind the tensor-core work)
          if (bi + 1 < p.NB) load_y(t, bi + 1, y);
          else if (tt + 1 < p.T) load_y(INV ? t - 1 : t + 1, 0, y);
          if (!(flags & TBF_LAST_IN_WIN)) {
            for (int op = 0; op < 4; ++op) rg.pass(rg.next_update(stages));      // the four update slabs of this block go by untouched
          } else if (!(flags & TBF_LAST)) {
            // window initialisation: split this row of every A chunk into its TF32 hi / lo images
            // (publishing 2 - 4 chunks behind one proxy fence was measured and is no faster: profiles/r2bf_split_batch.log)
            const int* Wn = sh.wins[sh.blocks[bi][TB_WIN] + 1];
            for (int op = 0; op < 4; ++op) {
              const int ktot = op == 0 ? Wn[TW_KX] : Wn[TW_KH];
              const int kc = op < 3 ? p.kc : p.kc_out;
              for (int k = 0; k < ktot; k += kc) {
                const int ke = min(kc, ktot - k);
                const uint32_t sl = rg.next_init(init_slots);
                rg.wait(sh.bfull, sl);
                float4* a = reinterpret_cast<float4*>(slot_ptr(sl)) + row_in_tile;
                for (int c = 0; c < (ke >> 2); ++c) {
                  const float4 v = a[c * 128];
                  float4 h, l;
                  split_tf32(v.x, h.x, l.x); split_tf32(v.y, h.y, l.y); split_tf32(v.z, h.z, l.z); split_tf32(v.w, h.w, l.w);
                  a[c * 128] = h;
                  a[(ke >> 2) * 128 + c * 128] = l;
                }
                fence_proxy_async();
                mbar_arrive(sh.asplit + sl);
              }
            }
          }
        }
      }
      if (valid) p.ladj[row] = ladj;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc<512>(tm);
}

static int tri_check(const int32_t* m, int32_t meta_len) {
  PMC_REQUIRE(m && meta_len >= TRI_HEADER && m[TRI_VER] == TRI_LAYOUT_VERSION, "pmc_flow_sweep_tri: not a block-triangular layout table (tri_layout.build_tri)");
  PMC_REQUIRE(m[TRI_KIND] == 0 || m[TRI_KIND] == 1, "pmc_flow_sweep_tri: unknown univariate head");
  PMC_REQUIRE(m[TRI_L] == 3 && m[TRI_GSIZE] == G, "pmc_flow_sweep_tri: built for 3 hidden layers and blocks of 4 order positions");
  PMC_REQUIRE(m[TRI_NB] >= 2 && m[TRI_NB] <= TRI_MAX_BLOCKS && m[TRI_NW] >= 1 && m[TRI_NW] <= TRI_MAX_WINDOWS, "pmc_flow_sweep_tri: table sizes out of range");
  PMC_REQUIRE(m[TRI_OFF_BLOCKS] == TRI_HEADER && m[TRI_OFF_WINDOWS] == TRI_HEADER + m[TRI_NB] * TB_FIELDS &&
              meta_len >= m[TRI_OFF_WINDOWS] + m[TRI_NW] * TW_FIELDS, "pmc_flow_sweep_tri: bad table offsets");
  return 0;
}

}  // namespace pmc

using namespace pmc;

extern "C" int64_t pmc_flow_sweep_tri_workspace(const int32_t* meta_host, int32_t meta_len, int64_t n) {
  if (tri_check(meta_host, meta_len) != 0) return -1;
  const long long tiles = (n + 127) / 128;
  return (int64_t)meta_host[TRI_WS_FLOATS] * std::min<long long>(std::max<long long>(tiles, 1), sm_count());
}

extern "C" int pmc_flow_sweep_tri(const float* packed, const int32_t* meta_host, const int32_t* meta_dev, int32_t meta_len, const float* in,
                                  float* out, float* ladj, int64_t n, int32_t inverse, int32_t passes, float* workspace,
                                  int64_t workspace_floats, pmc_stream_t stream) {
  PMC_REQUIRE(packed && meta_host && meta_dev && in && out && ladj, "pmc_flow_sweep_tri: null pointer");
  if (int rc = tri_check(meta_host, meta_len)) return rc;
  PMC_REQUIRE(passes == 1 || passes == 3, "pmc_flow_sweep_tri: passes must be 1 (TF32) or 3 (3xTF32, fp32 fidelity)");
  if (n == 0) return 0;
  const int* m = meta_host;
  TriParams q;
  q.packed = packed; q.in = in; q.out = out; q.ladj = ladj; q.n = n; q.ws = workspace;
  q.D = m[TRI_D]; q.T = m[TRI_T]; q.NB = m[TRI_NB]; q.NW = m[TRI_NW];
  q.tstride = m[TRI_TSTRIDE]; q.chunk_off = m[TRI_CHUNK_OFF]; q.passes = passes;
  q.stages = m[TRI_NSTAGES]; q.kc = m[TRI_KCHUNK]; q.kc_out = m[TRI_KCHUNK_OUT]; q.kh_total = m[TRI_KH_TOTAL]; q.kx_total = m[TRI_KX_TOTAL];
  q.ws_floats = m[TRI_WS_FLOATS];
  {
    const char* e = getenv("PMC_TRI_STAGES");
    if (e && atoi(e) >= 2 && atoi(e) < q.stages) q.stages = atoi(e);
    const char* nm = getenv("PMC_TRI_NOMMA");      // timing experiment only: results are wrong
    if (nm && nm[0] == '1') q.passes = 0;
  }
  q.slot_bytes = (uint32_t)m[TRI_SLOT_BYTES]; q.dslot_bytes = (uint32_t)m[TRI_DSLOT_BYTES]; q.tile_bytes = (uint32_t)m[TRI_TILE_BYTES];
  PMC_REQUIRE(m[TRI_NCOLS] <= 512, "pmc_flow_sweep_tri: accumulators exceed tensor memory");
  PMC_REQUIRE(q.kc_out >= 8 && q.kc_out % 8 == 0, "pmc_flow_sweep_tri: bad K chunking of the outputs");
  PMC_REQUIRE(q.kc >= 8 && q.kc % 8 == 0 && q.kh_total % 8 == 0 && q.kx_total == 8 * q.NB, "pmc_flow_sweep_tri: bad K chunking");
  q.tables = meta_dev + TRI_HEADER;
  const int* mb = m + m[TRI_OFF_BLOCKS];
  const int* mw = m + m[TRI_OFF_WINDOWS];
  for (int b = 0; b < q.NB; ++b) {
    const int* B = mb + b * TB_FIELDS;
    PMC_REQUIRE(B[TB_NR] >= 4 && B[TB_NR] <= 6 && B[TB_NST] >= 1 && B[TB_NST] <= G, "pmc_flow_sweep_tri: bad block shape");
    PMC_REQUIRE(B[TB_W] == 4 * B[TB_NR] && B[TB_KP] == (B[TB_W] + 7) / 8 * 8, "pmc_flow_sweep_tri: block width does not match its group size");
    PMC_REQUIRE((uint32_t)B[TB_DN] * 4u <= q.dslot_bytes && B[TB_DN] % 4 == 0 && B[TB_DOFF] % 4 == 0, "pmc_flow_sweep_tri: bad in-block slab");
    PMC_REQUIRE(B[TB_WIN] >= 0 && B[TB_WIN] < q.NW && B[TB_KS] % 8 == 0, "pmc_flow_sweep_tri: bad window index");
    PMC_REQUIRE(((B[TB_FLAGS] & TBF_LAST) != 0) == (b == q.NB - 1), "pmc_flow_sweep_tri: bad block flags");
    if (!(B[TB_FLAGS] & TBF_LAST_IN_WIN)) {
      PMC_REQUIRE(B[TB_UPD_N] % 16 == 0 && B[TB_UPD_N] >= 16 && B[TB_UPD_N] <= 256 && B[TB_OUT_N] % 16 == 0 && B[TB_OUT_N] >= 16 &&
                  B[TB_OUT_N] <= 256, "pmc_flow_sweep_tri: bad update width");
      PMC_REQUIRE((uint32_t)(B[TB_KP] * std::max(B[TB_UPD_N], B[TB_OUT_N]) * 8) <= q.slot_bytes, "pmc_flow_sweep_tri: update slab exceeds a ring slot");
    }
  }
  for (int w = 0; w < q.NW; ++w) {
    const int* Wn = mw + w * TW_FIELDS;
    PMC_REQUIRE(Wn[TW_WP] % 16 == 0 && Wn[TW_WP] >= 16 && Wn[TW_WP] <= 256 && Wn[TW_OP] % 16 == 0 && Wn[TW_OP] >= 16 &&
                3 * Wn[TW_WP] + Wn[TW_OP] <= 512 && Wn[TW_COL_OUT] == 3 * Wn[TW_WP], "pmc_flow_sweep_tri: bad window shape");
    PMC_REQUIRE(Wn[TW_KH] % 8 == 0 && Wn[TW_KX] % 8 == 0 && Wn[TW_KH] <= q.kh_total && Wn[TW_KX] <= q.kx_total, "pmc_flow_sweep_tri: bad window K extents");
    if (w > 0) PMC_REQUIRE((uint32_t)q.kc * (1024u + (uint32_t)Wn[TW_WP] * 8u) <= q.slot_bytes && (uint32_t)q.kc_out * (1024u + (uint32_t)Wn[TW_OP] * 8u) <= q.slot_bytes,
                           "pmc_flow_sweep_tri: init chunk exceeds a ring slot");
  }
  PMC_REQUIRE(q.stages >= 2 && q.stages <= TRI_MAX_STAGES, "pmc_flow_sweep_tri: bad ring depth");
  // a window initialisation borrows the A-tile area (3 layers x hi / lo + the x tile: idle between the last update of a
  // window and the first block of the next) as extra ring slots: more bytes in flight for the phase bound by the latency
  // of its operand stream
  q.extra = q.NW > 1 ? (int)std::min<size_t>((size_t)(TRI_MAX_SLOTS - q.stages), ((size_t)6 * q.tile_bytes + 8192) / q.slot_bytes) : 0;
  {
    const char* ex = getenv("PMC_TRI_EXTRA");
    if (ex && atoi(ex) >= 0 && atoi(ex) < q.extra) q.extra = atoi(ex);
  }
  const size_t smem = (size_t)6 * q.tile_bytes + 8192 + (size_t)q.stages * q.slot_bytes + 2 * (size_t)q.dslot_bytes;
  PMC_REQUIRE(smem + sizeof(TriShared) <= 227 * 1024, "pmc_flow_sweep_tri: shared memory budget exceeded");
  PMC_REQUIRE(q.slot_bytes % 1024 == 0 && q.dslot_bytes % 1024 == 0 && q.tile_bytes % 1024 == 0, "pmc_flow_sweep_tri: unaligned slot sizes");
  const long long tiles = (n + 127) / 128;
  const int grid = (int)std::min<long long>(tiles, sm_count());
  if (q.NW > 1) {
    PMC_REQUIRE(workspace != nullptr && workspace_floats >= (int64_t)q.ws_floats * grid && q.ws_floats == ((long long)q.kx_total + 3LL * q.kh_total) * 128,
                "pmc_flow_sweep_tri: workspace too small (pmc_flow_sweep_tri_workspace)");
  }
  cudaStream_t st = as_stream(stream);
  const bool rqs = m[TRI_KIND] != 0;
  auto kern = inverse ? (rqs ? made_sweep_tri_kernel<true, true> : made_sweep_tri_kernel<true, false>)
                      : (rqs ? made_sweep_tri_kernel<false, true> : made_sweep_tri_kernel<false, false>);
  PMC_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  PMC_TRY(launch_chain(kern, dim3(grid), dim3(TRI_THREADS), smem, st, q));
  PMC_LAUNCH_CHECK();
  return 0;
}

// Block-triangular MADE sweep on the 5th-generation tensor cores: Flow.inverse (and forward) of zuko MAF flows.
//
// Reference path: pocomc/flow.py:116-132 -> zuko transform.inv.call_and_ladj, the call that dominates every
// preconditioned MCMC step (pocomc/mcmc.py:88,256): D+1 dense hyper-network passes per transform.  Sorted by
// autoregressive degree the masks are block lower-triangular, so ONE forward substitution computes every hidden unit
// and every output once.  This kernel runs that substitution with the dense part on tcgen05:
//
//   * a CTA owns 128 particles = the 128 TMEM lanes; order positions are cut into blocks of 8 (made_layout.build_tri);
//   * TENSOR MEMORY holds the running pre-activations of every hidden unit of the three hidden layers and of every
//     output (one fp32 column each, <= 512 columns);
//   * when a block is finished, its activations -- an A tile [128 x K] per layer, TF32 hi / lo images in shared
//     memory -- update the accumulators of ALL later units: one group of tcgen05.mma (A and B from shared memory,
//     K-major no-swizzle) per layer, "urgent" columns (the next block) first and committed on their own mbarrier so the
//     next block can start while the updates of the blocks behind it are still in flight (right-looking schedule);
//     fp32 fidelity by the 3-pass split a_hi b_hi + a_lo b_hi + a_hi b_lo (parity bar 5e-5), passes = 1 for plain TF32;
//   * what stays inside a block -- the dependencies between its own 8 degree groups -- is fp32 FMA work with one
//     thread per particle: the block's accumulators are pulled out of TMEM into registers (they BECOME the activation
//     registers), the in-block weights arrive as warp-uniform LDS.128 broadcasts from a slab the producer streamed in;
//   * weights stream from L2 through two shared-memory rings (update slabs, in-block slabs) with 1-D bulk copies and
//     mbarrier transaction counts; producer warps, the MMA issuer and the 128 substitution threads are coupled by
//     mbarriers only.
#include "common.cuh"
#include "tc_common.cuh"
#include <algorithm>
#include <stdlib.h>

namespace pmc {

using namespace tc;

// header of made_layout.build_tri -- keep in sync
enum { TRI_D = 0, TRI_H, TRI_L, TRI_T, TRI_NB, TRI_HC, TRI_COL_OUT, TRI_NCOLS, TRI_TSTRIDE, TRI_NCHUNKS, TRI_SLOT_BYTES,
       TRI_DSLOT_BYTES, TRI_TILE_BYTES, TRI_OFF_BLOCKS, TRI_OFF_CHUNKS, TRI_VER, TRI_NSTAGES, TRI_GSIZE, TRI_HEADER };
enum { TB_K0 = 0, TB_NST, TB_U, TB_W, TB_HC, TB_DOFF, TB_DN, TB_C0, TB_NURG, TB_NCH, TB_FIELDS };
enum { TCK_ASRC = 0, TCK_KS0, TCK_NKS, TCK_N, TCK_DCOL, TCK_FIRST, TCK_OFF, TCK_FLAGS, TCK_FIELDS };

constexpr int TRI_MAX_STAGES = 16;         // update-slab ring depth: as many slots as fit, decided by made_layout.build_tri
constexpr int TRI_MAX_BLOCKS = 12;
constexpr int TRI_MAX_CHUNKS = 96;
constexpr int TRI_THREADS = 224;           // warps 0-3 substitution, 4 update-slab producer, 5 MMA issuer, 6 in-block slab producer
#define TRI_WAIT(bar, par) do { if (p.spin) mbar_spin(bar, par); else mbar_wait(bar, par); } while (0)
constexpr float TRI_LOG_SLOPE = -6.90775527898213705205f;

struct TriParams {
  const float* packed;
  const int* tables;       // device copy of the block / chunk tables (made_layout.build_tri meta from TRI_OFF_BLOCKS on)
  const float* in;
  float* out;
  float* ladj;
  long long n;
  int D, L, T, NB, Hc, col_out, tstride, n_chunks, passes, inverse, stages, spin;
  uint32_t slot_bytes, dslot_bytes, tile_bytes;
};

template <int NR, int G = 4>
struct TriShape {
  static constexpr int E = NR - 4;                       // extra units per group behind the four regular ones (0 or 1)
  static constexpr int W = 4 * G + E * G;                // slots of a block (G = 4: 16 or 20; G = 8: 32 or 40)
  static constexpr int KP = (W + 7) / 8 * 8;             // K extent of the block's A tiles
  static constexpr int NRV = (NR == 4 ? 1 : 2);          // float4 per bias vector
  static constexpr int Q1 = (2 * NR + 3) / 4;            // float4 per x pair of the layer-1 weights
  static constexpr int GRPV = NR + (E ? 2 : 0);          // float4 per source group of the layer-2/3 weights
  static constexpr int OGV = 2 + (E ? 1 : 0);            // float4 per source group of the output weights
};

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
// unit -- the activations.  Regular units as source pairs (packed FMAs), the extra unit of every group separately.
template <int NR, int G>
struct TriActs {
  float2 r[2 * G];
  float x[G];
};

template <int NR, int G>
__device__ __forceinline__ void acts_zero(TriActs<NR, G>& a) {
#pragma unroll
  for (int i = 0; i < 2 * G; ++i) a.r[i] = make_float2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < G; ++i) a.x[i] = 0.f;
}

template <int G>
struct TriRaw {            // raw TMEM images of one layer's block segment
  uint32_t r[4 * G];
  uint32_t x[G];
};
template <int NR, int G>
__device__ __forceinline__ void raw_load(const uint32_t taddr, TriRaw<G>& t) {
  if constexpr (G == 4) {
    tmem_ld16_async(taddr, t.r);
    if constexpr (NR == 5) tmem_ld4_async(taddr + 16, t.x);
  } else {
    tmem_ld32_async(taddr, t.r);
    if constexpr (NR == 5) tmem_ld8_async(taddr + 32, t.x);
  }
}
template <int NR, int G>
__device__ __forceinline__ void raw_take(TriRaw<G>& t, TriActs<NR, G>& a) {
  if constexpr (G == 4) {
    tmem_ld_fence16(t.r);
    if constexpr (NR == 5) tmem_ld_fence4(t.x);
  } else {
    tmem_ld_fence32(t.r);
    if constexpr (NR == 5) tmem_ld_fence8(t.x);
  }
#pragma unroll
  for (int i = 0; i < 2 * G; ++i) a.r[i] = make_float2(__uint_as_float(t.r[2 * i]), __uint_as_float(t.r[2 * i + 1]));
#pragma unroll
  for (int i = 0; i < G; ++i) a.x[i] = (NR == 5) ? __uint_as_float(t.x[i]) : 0.f;
}

// one residual hidden layer (2 or 3) of in-block group J: dst[own] = relu(src[own] + acc + bias + sum over groups 0..J)
template <int NR, int G, int J>
__device__ __forceinline__ void tri_hidden(const float4* __restrict__ q, int& off, const TriActs<NR, G>& src, TriActs<NR, G>& dst) {
  using S = TriShape<NR, G>;
  float4 bias[S::NRV];
  load_v(q + off, bias);
  off += S::NRV;
  float2 acc[NR];
#pragma unroll
  for (int s = 0; s < 4; ++s)
    acc[s] = make_float2(((s & 1) ? dst.r[2 * J + (s >> 1)].y : dst.r[2 * J + (s >> 1)].x) + elem_of(bias, s),
                         (s & 1) ? src.r[2 * J + (s >> 1)].y : src.r[2 * J + (s >> 1)].x);
  if constexpr (NR == 5) acc[4] = make_float2(dst.x[J] + elem_of(bias, 4), src.x[J]);
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
      for (int s = 0; s < NR; ++s) acc[s].x = fmaf(elem_of(wx, s), src.x[c], acc[s].x);
    }
  }
  dst.r[2 * J] = make_float2(fmaxf(acc[0].x + acc[0].y, 0.f), fmaxf(acc[1].x + acc[1].y, 0.f));
  dst.r[2 * J + 1] = make_float2(fmaxf(acc[2].x + acc[2].y, 0.f), fmaxf(acc[3].x + acc[3].y, 0.f));
  if constexpr (NR == 5) dst.x[J] = fmaxf(acc[4].x + acc[4].y, 0.f);
}

// order position J of a block: output -> affine map -> the degree group's three hidden layers
template <int NR, int G, bool INV, int J>
__device__ __forceinline__ void tri_stage(const float4* __restrict__ q, int& off, TriActs<NR, G>& a1, TriActs<NR, G>& a2, TriActs<NR, G>& a3,
                                          const uint32_t (&o)[2 * G], float2 (&xbp)[G / 2], const float (&y)[G], float& ladj,
                                          float* __restrict__ out_row, const bool valid, const int kstep, const int feat0) {
  using S = TriShape<NR, G>;
  {
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
        shf.x = fmaf(w[2].x, a3.x[c], shf.x);
        srw.x = fmaf(w[2].y, a3.x[c], srw.x);
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
    for (int s = 0; s < 4; ++s) acc[s] = make_float2(((s & 1) ? a1.r[2 * J + (s >> 1)].y : a1.r[2 * J + (s >> 1)].x) + elem_of(bias, s), 0.f);
    if constexpr (NR == 5) acc[4] = make_float2(a1.x[J] + elem_of(bias, 4), 0.f);
#pragma unroll
    for (int c = 0; c <= (J >> 1); ++c) {
      float4 w[S::Q1];
      load_v(q + off, w);
      off += S::Q1;
#pragma unroll
      for (int s = 0; s < NR; ++s) acc[s] = ffma2(pair_of(w, s), xbp[c], acc[s]);
    }
    a1.r[2 * J] = make_float2(fmaxf(acc[0].x + acc[0].y, 0.f), fmaxf(acc[1].x + acc[1].y, 0.f));
    a1.r[2 * J + 1] = make_float2(fmaxf(acc[2].x + acc[2].y, 0.f), fmaxf(acc[3].x + acc[3].y, 0.f));
    if constexpr (NR == 5) a1.x[J] = fmaxf(acc[4].x + acc[4].y, 0.f);
  }
  tri_hidden<NR, G, J>(q, off, a1, a2);
  tri_hidden<NR, G, J>(q, off, a2, a3);
}

// order positions J .. G-1 of a block (compile-time recursion: every register index stays static)
template <int NR, int G, bool INV, int J>
__device__ __forceinline__ void tri_stages(const float4* __restrict__ q, int& off, const int nst, TriActs<NR, G>& a1, TriActs<NR, G>& a2,
                                           TriActs<NR, G>& a3, const uint32_t (&o)[2 * G], float2 (&xbp)[G / 2], const float (&y)[G], float& ladj,
                                           float* __restrict__ out_row, const bool valid, const int kstep, const int feat0) {
  if (J >= nst) return;
  tri_stage<NR, G, INV, J>(q, off, a1, a2, a3, o, xbp, y, ladj, out_row, valid, kstep, feat0);
  if constexpr (J + 1 < G) tri_stages<NR, G, INV, J + 1>(q, off, nst, a1, a2, a3, o, xbp, y, ladj, out_row, valid, kstep, feat0);
}

struct TriShared {
  uint64_t bfull[TRI_MAX_STAGES], bempty[TRI_MAX_STAGES], dfull[2], dempty[2], a_ready, urgent_done, rest_done;
  uint32_t tmem_slot;
  int blocks[TRI_MAX_BLOCKS][TB_FIELDS];
  int chunks[TRI_MAX_CHUNKS][TCK_FIELDS];
  alignas(16) uint32_t issue[TRI_MAX_CHUNKS][8];   // per chunk: everything the MMA issuer needs, precomputed (see the kernel prologue)
};

// A tile of one layer: [128 rows x KP] as 16-byte chunks of 4 consecutive k, chunk (r, c) at c * 2048 + r * 16;
// hi / lo TF32 images
template <int NR, int G>
__device__ __forceinline__ void store_tile(unsigned char* hi, unsigned char* lo, const int r, const TriActs<NR, G>& a) {
#pragma unroll
  for (int c = 0; c < G; ++c) {
    float4 h, l;
    split_tf32(a.r[2 * c].x, h.x, l.x);
    split_tf32(a.r[2 * c].y, h.y, l.y);
    split_tf32(a.r[2 * c + 1].x, h.z, l.z);
    split_tf32(a.r[2 * c + 1].y, h.w, l.w);
    *reinterpret_cast<float4*>(hi + c * 2048 + r * 16) = h;
    *reinterpret_cast<float4*>(lo + c * 2048 + r * 16) = l;
  }
  if constexpr (NR == 5) {
#pragma unroll
    for (int c = 0; c < G / 4; ++c) {
      float4 h, l;
      split_tf32(a.x[4 * c], h.x, l.x); split_tf32(a.x[4 * c + 1], h.y, l.y); split_tf32(a.x[4 * c + 2], h.z, l.z); split_tf32(a.x[4 * c + 3], h.w, l.w);
      *reinterpret_cast<float4*>(hi + (G + c) * 2048 + r * 16) = h;
      *reinterpret_cast<float4*>(lo + (G + c) * 2048 + r * 16) = l;
    }
    if constexpr (TriShape<NR, G>::KP > TriShape<NR, G>::W) {      // K is padded to a multiple of 8 (G = 4: slots 20..23)
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4*>(hi + (G + G / 4) * 2048 + r * 16) = z;
      *reinterpret_cast<float4*>(lo + (G + G / 4) * 2048 + r * 16) = z;
    }
  }
}

// one block of the substitution on the 128 particle threads
template <int NR, int G, bool INV>
__device__ __forceinline__ void run_block(const TriParams& p, const int bi, const int t, const uint32_t lane_base, unsigned char* smem,
                                          const float4* __restrict__ slab, TriShared& sh, uint32_t& n_updates, const int row_in_tile,
                                          float* out_row, const bool valid, float& ladj, const float (&y)[G]) {
  const int* B = sh.blocks[bi];
  const int k0 = B[TB_K0], nst = B[TB_NST], hc = B[TB_HC];
  const bool rev = (t & 1);
  const int feat0 = rev ? (p.D - 1 - k0) : k0;
  const int kstep = rev ? -1 : 1;
  TriActs<NR, G> a1, a2, a3;
  uint32_t o[2 * G];
  if (bi == 0) {
    acts_zero(a1); acts_zero(a2); acts_zero(a3);
#pragma unroll
    for (int i = 0; i < 2 * G; ++i) o[i] = 0u;
  } else {
    // accumulators of this block are complete once the previous block's urgent updates have landed
    TRI_WAIT(&sh.urgent_done, (n_updates - 1) & 1);
    tc_fence_after();
    TriRaw<G> t1, t2, t3;
    raw_load<NR, G>(lane_base + hc, t1);
    raw_load<NR, G>(lane_base + p.Hc + hc, t2);
    raw_load<NR, G>(lane_base + 2 * p.Hc + hc, t3);
    if constexpr (G == 4) tmem_ld8_async(lane_base + p.col_out + 2 * k0, o); else tmem_ld16_async(lane_base + p.col_out + 2 * k0, o);
    raw_take<NR, G>(t1, a1);
    raw_take<NR, G>(t2, a2);
    raw_take<NR, G>(t3, a3);
    if constexpr (G == 4) tmem_ld_fence8(o); else tmem_ld_fence16(o);
  }
  float2 xbp[G / 2];
#pragma unroll
  for (int i = 0; i < G / 2; ++i) xbp[i] = make_float2(0.f, 0.f);
  int off = 0;
  tri_stages<NR, G, INV, 0>(slab, off, nst, a1, a2, a3, o, xbp, y, ladj, out_row, valid, kstep, feat0);
  if (bi + 1 < p.NB) {
    // hand the block's activations to the tensor core: the previous update group must have finished reading the tiles
    if (n_updates > 0) TRI_WAIT(&sh.rest_done, (n_updates - 1) & 1);
    unsigned char* tiles = smem;
    store_tile<NR, G>(tiles, tiles + p.tile_bytes, row_in_tile, a1);
    store_tile<NR, G>(tiles + 2 * p.tile_bytes, tiles + 3 * p.tile_bytes, row_in_tile, a2);
    store_tile<NR, G>(tiles + 4 * p.tile_bytes, tiles + 5 * p.tile_bytes, row_in_tile, a3);
    unsigned char* xt = tiles + 6 * p.tile_bytes;
    {
      float4 h, l;
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      split_tf32(xbp[0].x, h.x, l.x); split_tf32(xbp[0].y, h.y, l.y); split_tf32(xbp[1].x, h.z, l.z); split_tf32(xbp[1].y, h.w, l.w);
      *reinterpret_cast<float4*>(xt + row_in_tile * 16) = h;
      *reinterpret_cast<float4*>(xt + 4096 + row_in_tile * 16) = l;
      if constexpr (G == 4) {                                                  // K = 8: order positions 4..7 do not exist
        *reinterpret_cast<float4*>(xt + 2048 + row_in_tile * 16) = z;
        *reinterpret_cast<float4*>(xt + 4096 + 2048 + row_in_tile * 16) = z;
      } else {
        split_tf32(xbp[2].x, h.x, l.x); split_tf32(xbp[2].y, h.y, l.y); split_tf32(xbp[3].x, h.z, l.z); split_tf32(xbp[3].y, h.w, l.w);
        *reinterpret_cast<float4*>(xt + 2048 + row_in_tile * 16) = h;
        *reinterpret_cast<float4*>(xt + 4096 + 2048 + row_in_tile * 16) = l;
      }
    }
    fence_proxy_async();
    tc_fence_before();
    mbar_arrive(&sh.a_ready);
    ++n_updates;
  }
}

template <bool INV, int G>
__global__ void __launch_bounds__(TRI_THREADS, 1)
made_sweep_tri_kernel(const __grid_constant__ TriParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ TriShared sh;
  // shared-memory map: [3 layers x (hi, lo) A tiles][x tile hi 4 KB][x tile lo 4 KB][update-slab ring][in-block slab ring]
  unsigned char* ring = smem + 6 * p.tile_bytes + 8192;
  unsigned char* dring = ring + (size_t)p.stages * p.slot_bytes;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < p.NB * TB_FIELDS; i += blockDim.x) (&sh.blocks[0][0])[i] = p.tables[i];
  for (int i = threadIdx.x; i < p.n_chunks * TCK_FIELDS; i += blockDim.x) (&sh.chunks[0][0])[i] = p.tables[p.NB * TB_FIELDS + i];
  __syncthreads();
  for (int c = threadIdx.x; c < p.n_chunks; c += blockDim.x) {
    // the MMA issuer is ONE thread whose instruction stream sits on the critical path of every block: hand it
    // ready-made descriptor words.  K-major no-swizzle descriptors: [0,14) address >> 4, [16,30) LBO >> 4,
    // [32,46) SBO >> 4 = 8, bit 46 version; A tiles: LBO = 2048 (128 rows x 16 B), B slabs: LBO = N x 16.
    const int* ck = sh.chunks[c];
    const int N = ck[TCK_N], nks = ck[TCK_NKS], asrc = ck[TCK_ASRC];
    const uint32_t tiles_addr = smem_u32(smem);
    const uint32_t a_base = (asrc == 0 ? tiles_addr + 6 * p.tile_bytes : tiles_addr + (uint32_t)(2 * (asrc - 1)) * p.tile_bytes) +
                            (uint32_t)ck[TCK_KS0] * 4096u;
    const uint32_t a_lo = a_base + (asrc == 0 ? 4096u : p.tile_bytes);
    uint32_t* w = sh.issue[c];
    w[0] = ((a_base & 0x3FFFF) >> 4) | ((2048u >> 4) << 16);
    w[1] = ((a_lo & 0x3FFFF) >> 4) | ((2048u >> 4) << 16);
    w[2] = (((uint32_t)N * 16u) >> 4) << 16;
    w[3] = (uint32_t)(nks * N * 32) >> 4;
    w[4] = ((uint32_t)N * 32u) >> 4;
    w[5] = idesc_tf32(128, N);
    w[6] = (uint32_t)ck[TCK_DCOL];
    w[7] = (uint32_t)nks | ((uint32_t)(ck[TCK_FIRST] ? 1 : 0) << 8) | ((uint32_t)ck[TCK_FLAGS] << 16);
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < TRI_MAX_STAGES; ++i) { mbar_init(sh.bfull + i, 1); mbar_init(sh.bempty + i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(sh.dfull + i, 1); mbar_init(sh.dempty + i, 128); }
    mbar_init(&sh.a_ready, 128);
    mbar_init(&sh.urgent_done, 1);
    mbar_init(&sh.rest_done, 1);
    mbar_fence_init();
  }
  if (warp == 4) tmem_alloc<512>(&sh.tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = sh.tmem_slot;
  const long long n_tiles = (p.n + 127) / 128;

  if (warp == 4) {
    // ---------------- producer 1: update slabs (B operands of the tensor-core updates) ----------------
    if (lane == 0) {
      uint32_t slot = 0, round = 0;
      for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int tt = 0; tt < p.T; ++tt) {
          const int t = INV ? p.T - 1 - tt : tt;
          const float* base = p.packed + (size_t)t * p.tstride;
          for (int c = 0; c < p.n_chunks; ++c) {
            if (round > 0) mbar_wait(sh.bempty + slot, (round - 1) & 1);
            const uint32_t bytes = (uint32_t)(sh.chunks[c][TCK_NKS] * sh.chunks[c][TCK_N] * 64);
            mbar_expect_tx(sh.bfull + slot, bytes);
            bulk_g2s(ring + (size_t)slot * p.slot_bytes, base + sh.chunks[c][TCK_OFF], bytes, sh.bfull + slot);
            if (++slot == (uint32_t)p.stages) { slot = 0; ++round; }
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
      uint32_t slot = 0, round = 0, n_upd = 0;
      const uint32_t ring16 = (smem_u32(ring) & 0x3FFFF) >> 4, slot16 = p.slot_bytes >> 4;
      const uint32_t desc_top = (128u >> 4) | (1u << 14);              // high word: SBO = 128 bytes, descriptor version 1
      const bool split = p.passes > 1;
      for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int tt = 0; tt < p.T; ++tt) {
          for (int bi = 0; bi + 1 < p.NB; ++bi) {
            const int c0 = sh.blocks[bi][TB_C0], c1 = c0 + sh.blocks[bi][TB_NCH];
            TRI_WAIT(&sh.a_ready, n_upd & 1);
            ++n_upd;
            tc_fence_after();
            for (int c = c0; c < c1; ++c) {
              const uint4 r0 = *reinterpret_cast<const uint4*>(sh.issue[c]);
              const uint4 r1 = *reinterpret_cast<const uint4*>(sh.issue[c] + 4);
              const uint32_t nks = r1.w & 0xffu, flags = r1.w >> 16;
              uint32_t accum = (r1.w & 0x100u) ? 0u : 1u;
              uint32_t a_hi = r0.x, a_lo = r0.y, b_hi = r0.z + ring16 + slot * slot16;
              const uint32_t d = tm + r1.z;
              TRI_WAIT(sh.bfull + slot, round & 1);
              tc_fence_after();
              for (uint32_t ks = 0; ks < nks && p.passes > 0; ++ks) {
                const uint64_t dah = ((uint64_t)desc_top << 32) | a_hi, dbh = ((uint64_t)desc_top << 32) | b_hi;
                mma_tf32_ss(d, dah, dbh, r1.y, accum);
                accum = 1u;
                if (split) {
                  const uint64_t dal = ((uint64_t)desc_top << 32) | a_lo, dbl = ((uint64_t)desc_top << 32) | (b_hi + r0.w);
                  mma_tf32_ss(d, dal, dbh, r1.y, 1u);
                  mma_tf32_ss(d, dah, dbl, r1.y, 1u);
                }
                a_hi += 256u; a_lo += 256u; b_hi += r1.x;
              }
              mma_commit(sh.bempty + slot);
              if (flags & 1) mma_commit(&sh.urgent_done);
              if (flags & 2) mma_commit(&sh.rest_done);
              if (++slot == (uint32_t)p.stages) { slot = 0; ++round; }
            }
          }
        }
      }
    }
  } else if (warp < 4) {
    // ---------------- substitution: thread = particle row = TMEM lane ----------------
    const int row_in_tile = threadIdx.x;
    const uint32_t lane_base = tm + ((uint32_t)(warp * 32) << 16);
    uint32_t n_updates = 0, dit = 0;
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
          if (sh.blocks[bi][TB_U] <= 4) run_block<4, G, INV>(p, bi, t, lane_base, smem, dslab, sh, n_updates, row_in_tile, out_row, valid, ladj, y);
          else run_block<5, G, INV>(p, bi, t, lane_base, smem, dslab, sh, n_updates, row_in_tile, out_row, valid, ladj, y);
          mbar_arrive(sh.dempty + slot);
          ++dit;
          // the next block's inputs (this thread's own earlier stores; L2 latency hides behind the tensor-core update)
          if (bi + 1 < p.NB) load_y(t, bi + 1, y);
          else if (tt + 1 < p.T) load_y(INV ? t - 1 : t + 1, 0, y);
        }
      }
      if (valid) p.ladj[row] = ladj;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc<512>(tm);
}

}  // namespace pmc

using namespace pmc;

extern "C" int pmc_flow_sweep_tri(const float* packed, const int32_t* meta_host, const int32_t* meta_dev, int32_t meta_len, const float* in,
                                  float* out, float* ladj, int64_t n, int32_t inverse, int32_t passes, pmc_stream_t stream) {
  PMC_REQUIRE(packed && meta_host && meta_dev && in && out && ladj, "pmc_flow_sweep_tri: null pointer");
  PMC_REQUIRE(meta_len >= TRI_HEADER && meta_host[TRI_VER] == 203, "pmc_flow_sweep_tri: not a block-triangular layout table");
  PMC_REQUIRE(passes == 1 || passes == 3, "pmc_flow_sweep_tri: passes must be 1 (TF32) or 3 (3xTF32, fp32 fidelity)");
  if (n == 0) return 0;
  const int* m = meta_host;
  TriParams q;
  q.packed = packed; q.in = in; q.out = out; q.ladj = ladj; q.n = n;
  q.D = m[TRI_D]; q.L = m[TRI_L]; q.T = m[TRI_T]; q.NB = m[TRI_NB]; q.Hc = m[TRI_HC]; q.col_out = m[TRI_COL_OUT];
  q.tstride = m[TRI_TSTRIDE]; q.n_chunks = m[TRI_NCHUNKS]; q.passes = passes; q.inverse = inverse;
  q.stages = m[TRI_NSTAGES];
  {
    const char* e = getenv("PMC_TRI_STAGES");
    if (e && atoi(e) >= 2 && atoi(e) < q.stages) q.stages = atoi(e);
    const char* sp = getenv("PMC_TRI_SPIN");
    q.spin = (sp && sp[0] == '1') ? 1 : 0;
    const char* nm = getenv("PMC_TRI_NOMMA");      // timing experiment only: results are wrong
    if (nm && nm[0] == '1') q.passes = 0;
  }
  q.slot_bytes = (uint32_t)m[TRI_SLOT_BYTES]; q.dslot_bytes = (uint32_t)m[TRI_DSLOT_BYTES]; q.tile_bytes = (uint32_t)m[TRI_TILE_BYTES];
  PMC_REQUIRE(q.L == 3, "pmc_flow_sweep_tri: built for 3 hidden layers");
  const int G = m[TRI_GSIZE];
  PMC_REQUIRE(G == 4 || G == 8, "pmc_flow_sweep_tri: blocks of 4 or 8 order positions");
  PMC_REQUIRE(q.NB >= 2 && q.NB <= TRI_MAX_BLOCKS && q.n_chunks >= 1 && q.n_chunks <= TRI_MAX_CHUNKS, "pmc_flow_sweep_tri: table sizes out of range");
  PMC_REQUIRE(m[TRI_NCOLS] <= 512, "pmc_flow_sweep_tri: accumulators exceed tensor memory");
  PMC_REQUIRE(m[TRI_OFF_BLOCKS] == TRI_HEADER && m[TRI_OFF_CHUNKS] == TRI_HEADER + q.NB * TB_FIELDS &&
              meta_len >= m[TRI_OFF_CHUNKS] + q.n_chunks * TCK_FIELDS, "pmc_flow_sweep_tri: bad table offsets");
  q.tables = meta_dev + TRI_HEADER;
  const int* mb = m + m[TRI_OFF_BLOCKS];
  for (int b = 0; b < q.NB; ++b) {
    const int* B = mb + b * TB_FIELDS;
    PMC_REQUIRE(B[TB_U] <= 5 && B[TB_NST] >= 1 && B[TB_NST] <= G, "pmc_flow_sweep_tri: bad block shape");
    PMC_REQUIRE(B[TB_W] == (B[TB_U] <= 4 ? 4 * G : 5 * G), "pmc_flow_sweep_tri: block width does not match its group size");
    PMC_REQUIRE((uint32_t)B[TB_DN] * 4u <= q.dslot_bytes && B[TB_DN] % 4 == 0 && B[TB_DOFF] % 4 == 0, "pmc_flow_sweep_tri: bad in-block slab");
  }
  const int* mc = m + m[TRI_OFF_CHUNKS];
  for (int c = 0; c < q.n_chunks; ++c) {
    const int* C = mc + c * TCK_FIELDS;
    PMC_REQUIRE(C[TCK_N] % 16 == 0 && C[TCK_N] >= 16 && C[TCK_N] <= 256, "pmc_flow_sweep_tri: bad update width");
    PMC_REQUIRE((uint32_t)(C[TCK_NKS] * C[TCK_N] * 64) <= q.slot_bytes && C[TCK_OFF] % 4 == 0, "pmc_flow_sweep_tri: bad update slab");
  }
  PMC_REQUIRE(q.stages >= 2 && q.stages <= TRI_MAX_STAGES, "pmc_flow_sweep_tri: bad ring depth");
  const size_t smem = (size_t)6 * q.tile_bytes + 8192 + (size_t)q.stages * q.slot_bytes + 2 * (size_t)q.dslot_bytes;
  PMC_REQUIRE(smem + sizeof(TriShared) <= 227 * 1024, "pmc_flow_sweep_tri: shared memory budget exceeded");
  PMC_REQUIRE(q.slot_bytes % 1024 == 0 && q.dslot_bytes % 1024 == 0 && q.tile_bytes % 1024 == 0, "pmc_flow_sweep_tri: unaligned slot sizes");
  const long long tiles = (n + 127) / 128;
  const int grid = (int)std::min<long long>(tiles, sm_count());
  cudaStream_t st = as_stream(stream);
#define PMC_TRI_LAUNCH(INVV, GV)                                                                    \
  {                                                                                                  \
    auto kern = made_sweep_tri_kernel<INVV, GV>;                                                     \
    PMC_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));     \
    kern<<<grid, TRI_THREADS, smem, st>>>(q);                                                        \
  }
  if (inverse && G == 4) PMC_TRI_LAUNCH(true, 4)
  else if (inverse) PMC_TRI_LAUNCH(true, 8)
  else if (G == 4) PMC_TRI_LAUNCH(false, 4)
  else PMC_TRI_LAUNCH(false, 8)
#undef PMC_TRI_LAUNCH
  PMC_LAUNCH_CHECK();
  return 0;
}

// Dense MADE forward on the 5th-generation tensor cores: Flow.forward / Flow.log_prob for zuko MAF.
//
// Reference path: pocomc/flow.py:99-114,134-147 -> zuko transform.call_and_ladj: per transform ONE pass of
// the masked MLP hyper-network (Linear-ReLU, L-1 residual Linear blocks, Linear) followed by the
// elementwise monotonic affine map.  That pass is a chain of genuine dense contractions
// [128 particles x K] x [K x N], so it runs on tcgen05:
//
//   * a CTA owns a tile of 128 particles = the 128 TMEM lanes; one thread per particle row does the
//     epilogues (bias, residual, ReLU, affine map, log-det accumulation);
//   * ACTIVATIONS NEVER LEAVE TENSOR MEMORY: the epilogue writes the next layer's A operand straight
//     back into TMEM with tcgen05.st and the MMA reads A from TMEM (tcgen05.mma [d], [a], b-desc);
//   * weights (mask already folded in, pre-split into TF32 hi/lo images by pmc_flow_tc_pack) stream
//     from L2 through a shared-memory ring with 1-D bulk copies (cp.async.bulk + mbarrier tx counts),
//     issued by a producer thread; an issuer thread feeds the tensor core; both are decoupled from the
//     128 epilogue threads by mbarriers only (no __syncthreads in the steady state);
//   * fp32 fidelity (the reference flow is fp32; parity bar 2e-5): every product is evaluated as
//     a_hi*b_hi + a_lo*b_hi + a_hi*b_lo with hi = value rounded to TF32 (the tensor core itself truncates,
//     measured in tests/tc_probe.cu: a truncated hi biases every product) and lo = value - hi (exact), fp32 accumulation in TMEM.
//     passes = 1 issues only the hi*hi term (plain TF32, for throughput experiments).
//
//   * biases ride on the tensor core too: the last weight chunk of every layer carries one extra k-step
//     whose k = 0 / k = 1 rows are bias_hi / bias_lo, multiplied by a constant A block (1, 1, 0, ...) that
//     sits in 8 TMEM columns -- the epilogue never touches global memory for parameters.
//
// TMEM column map (512 columns x 128 lanes x 32 bit):
//   [0,128) accumulator | [128,256) h hi | [256,384) h lo | [384,432) x hi | [432,480) x lo | [480,488) ones
#include "common.cuh"
#include "tc_common.cuh"
#include <algorithm>

namespace pmc {

using namespace tc;

// meta header of the tensor-core layout -- keep in sync with made_layout.build_tc
enum { TC_D = 0, TC_H, TC_L, TC_T, TC_KIND, TC_KX, TC_NOUT, TC_TSTRIDE, TC_BIAS_OFF, TC_NCHUNKS, TC_SLOT_BYTES, TC_VERSION, TC_LEN };

constexpr int TC_MAX_STAGES = 6;             // ring depth (slots of <= 32 KB): as many as fit beside the bias and x tables
constexpr int TC_KCHUNK = 32;                // k extent of one weight chunk = columns per epilogue step
constexpr int TC_XLD = 49;                   // row stride of the x table (Kx <= 48)
// tensor memory: TWO accumulators (layer g writes ACC[g & 1] while the epilogue still reads the other one) and the A
// operand of the next layer as TF32 hi / lo images; the transform's input x lives in the first Kx columns of the A images
// while layer 0 runs (nothing else needs them then) and in a shared-memory table for the output epilogue
constexpr uint32_t COL_ACC = 0, COL_H_HI = 256, COL_H_LO = 384;
constexpr float TC_LOG_SLOPE = -6.90775527898213705205f;  // log(1e-3), zuko MonotonicAffineTransform

struct TcParams {
  const float* packed;    // per transform: weight chunks (hi image, lo image) ..., then biases
  const float* in;        // [n, D]
  float* out;             // [n, D]
  float* ladj;            // [n]
  long long n;
  int D, H, L, T, Kx, Nout, tstride, bias_off, passes, stages;
  uint32_t slot_bytes;
};

// the chunk sequence of one transform, identical for producer and issuer:
// layer l = 0..L, K_l = (l == 0 ? Kx : H), N_l = (l == L ? Nout : H), chunks of <= 32 k
struct ChunkIter {
  int l, i, c0, K, N, kc;       // i: position in the layer's issue order, c0: first k of the chunk at that position
  __device__ __forceinline__ void start(const TcParams& p) { l = 0; i = 0; set(p); }
  __device__ __forceinline__ void set(const TcParams& p) {
    K = (l == 0) ? p.Kx : p.H;
    N = (l == p.L) ? p.Nout : p.H;
    c0 = TC_KCHUNK * ((l == 0) ? i : ((p.H == 128) ? ((i & 1) * 2 + (i >> 1)) : i));
    kc = min(TC_KCHUNK, K - c0);
  }
  __device__ __forceinline__ int n_chunks() const { return (K + TC_KCHUNK - 1) / TC_KCHUNK; }
  __device__ __forceinline__ bool last_of_layer() const { return i + 1 >= n_chunks(); }
  __device__ __forceinline__ uint32_t bytes() const { return (uint32_t)(2 * kc * N * 4); }      // hi image + lo image
  // byte offset of this chunk inside its layer's block of the packed image (chunks are stored in k order)
  __device__ __forceinline__ uint32_t offset_in_layer() const { return (uint32_t)(2 * c0 * N * 4); }
  __device__ __forceinline__ uint32_t layer_bytes() const { return (uint32_t)(2 * K * N * 4); }
  // returns false once the transform is exhausted
  __device__ __forceinline__ bool next(const TcParams& p) {
    if (++i >= n_chunks()) { ++l; i = 0; if (l > p.L) return false; }
    set(p);
    return true;
  }
};

__host__ __device__ constexpr int tc_parts(int h) { return h >= 64 ? 2 : 1; }
__host__ __device__ constexpr int tc_threads(int h) { return (4 * tc_parts(h) + 2) * 32; }
// (ChunkIter walks the k-chunks of a hidden / output layer in the order in which the epilogue parts finish them: at H = 128 both
// parts deliver their first chunk, then their second -- 0, 2, 1, 3)

template <int H>
__global__ void __launch_bounds__(tc_threads(H), 1)
made_forward_tc_kernel(const TcParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // a_chunk[j]: columns 32 j .. 32 j + 31 of the next hidden / output layer's A operand are in tensor memory; x_chunk[j]: the
  // same for layer 0 (the transform's input).  The issuer starts k-chunk j on it while the epilogue threads are still producing
  // the chunks behind it from the other accumulator.
  constexpr int NPART = tc_parts(H), CP = H / NPART, EPI_WARPS = 4 * NPART, EPI_THREADS = 128 * NPART;
  __shared__ uint64_t full[TC_MAX_STAGES], empty[TC_MAX_STAGES], a_chunk[4], x_chunk[2], acc_full;
  __shared__ uint32_t tmem_slot;
  __shared__ float ladj_s[NPART][128];
  unsigned char* ring = smem_raw;
  const int bias_per_t = p.L * p.H + p.Nout;
  float* bias_s = reinterpret_cast<float*>(smem_raw + (size_t)p.stages * p.slot_bytes);      // [T][L*H + Nout]
  float* xs = bias_s + (size_t)p.T * bias_per_t;                                             // [128][TC_XLD]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < TC_MAX_STAGES; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, 1); }
    for (int i = 0; i < 4; ++i) mbar_init(a_chunk + i, 128);
    for (int i = 0; i < 2; ++i) mbar_init(x_chunk + i, EPI_THREADS);
    mbar_init(&acc_full, 1);
    mbar_fence_init();
  }
  for (int i = threadIdx.x; i < p.T * bias_per_t; i += blockDim.x) {
    const int t = i / bias_per_t, j = i - t * bias_per_t;
    bias_s[i] = p.packed[(size_t)t * p.tstride + p.bias_off + j];
  }
  if (warp == EPI_WARPS) tmem_alloc<512>(&tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_slot;
  const long long n_tiles = (p.n + 127) / 128;
  const uint32_t stages = (uint32_t)p.stages;

  if (warp == EPI_WARPS) {
    // ---------------- producer: stream the weight chunks of every transform, once per tile ----------------
    if (lane == 0) {
      uint32_t slot = 0, round = 0;
      for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int t = 0; t < p.T; ++t) {
          const unsigned char* layer_src = reinterpret_cast<const unsigned char*>(p.packed + (size_t)t * p.tstride);
          ChunkIter ci;
          ci.start(p);
          bool more = true;
          while (more) {
            if (round > 0) mbar_wait(empty + slot, (round - 1) & 1);
            const uint32_t bytes = ci.bytes();
            mbar_expect_tx(full + slot, bytes);
            bulk_g2s(ring + (size_t)slot * p.slot_bytes, layer_src + ci.offset_in_layer(), bytes, full + slot);
            if (++slot == stages) { slot = 0; ++round; }
            const int l_before = ci.l;
            const uint32_t lb = ci.layer_bytes();
            more = ci.next(p);
            if (!more || ci.l != l_before) layer_src += lb;
          }
        }
      }
    }
  } else if (warp == EPI_WARPS + 1) {
    // ---------------- issuer: one thread drives the tensor core ----------------
    if (lane == 0) {
      uint32_t slot = 0, round = 0, ph_a = 0, ph_x = 0, g = 0;      // ph_*: one parity bit per chunk barrier; g: layers issued so far
      const uint32_t ring_addr = smem_u32(ring);
      for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int t = 0; t < p.T; ++t) {
          ChunkIter ci;
          ci.start(p);
          bool more = true;
          while (more) {
            const uint32_t j = (uint32_t)ci.c0 / TC_KCHUNK;
            if (ci.l == 0) { mbar_wait(x_chunk + j, (ph_x >> j) & 1u); ph_x ^= 1u << j; }     // this k-chunk of the A operand is in tensor memory
            else { mbar_wait(a_chunk + j, (ph_a >> j) & 1u); ph_a ^= 1u << j; }
            mbar_wait(full + slot, round & 1);
            tc_fence_after();
            const uint32_t acc = tm + COL_ACC + (g & 1u) * 128u;
            const uint32_t a_hi = tm + COL_H_HI + ci.c0, a_lo = tm + COL_H_LO + ci.c0;
            const uint32_t id = idesc_tf32(128, ci.N);
            const uint32_t kstride = (uint32_t)ci.N * 16u;            // bytes between consecutive 4-k chunk columns
            const uint32_t b_hi = ring_addr + slot * p.slot_bytes, b_lo = b_hi + (uint32_t)(ci.kc * ci.N * 4);
            for (int ks = 0; ks < ci.kc / 8; ++ks) {
              const uint64_t dh = smem_desc(b_hi + ks * 2 * kstride, kstride, 128);
              const uint32_t first = (ci.i == 0 && ks == 0) ? 0u : 1u;
              mma_tf32_ts(acc, a_hi + ks * 8, dh, id, first);
              if (p.passes > 1) {
                const uint64_t dl = smem_desc(b_lo + ks * 2 * kstride, kstride, 128);
                mma_tf32_ts(acc, a_lo + ks * 8, dh, id, 1u);
                mma_tf32_ts(acc, a_hi + ks * 8, dl, id, 1u);
              }
            }
            mma_commit(empty + slot);          // slot reusable once these MMAs have read it
            if (++slot == stages) { slot = 0; ++round; }
            const int l_before = ci.l;
            more = ci.next(p);
            if (!more || ci.l != l_before) { mma_commit(&acc_full); ++g; }   // layer complete -> epilogue
          }
        }
      }
    }
  } else {
    // ---------------- epilogue: thread = (particle row = TMEM lane, one part of the columns) ----------------
    const int part = warp >> 2;                                 // warps 4 q .. 4 q + 3: columns [q CP, (q + 1) CP)
    const int row_in_tile = (warp & 3) * 32 + lane;             // 0..127
    const uint32_t lane_base = tm + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t ph_acc = 0, g = 0;
    float hreg[CP];                                             // this row's hidden activations of this part (residual input)
    float* xrow = xs + (size_t)row_in_tile * TC_XLD;
    const bool vec4 = (p.D % 4) == 0;
    const int x_chunks = (p.Kx + TC_KCHUNK - 1) / TC_KCHUNK;    // k-chunks of layer 0 (1 or 2)
    auto publish = [&](uint64_t* bar) {                         // the A columns behind `bar` are written: hand them to the issuer
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(bar);
    };
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const long long row = tile * 128 + row_in_tile;
      const bool valid = row < p.n;
      // stage the input row as the layer-0 A operand (hi / lo) and in the x table: 4-column groups dealt over the parts
      for (int c = 4 * part; c < p.Kx; c += 4 * NPART) {
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (valid) {
          if (vec4 && c < p.D) {
            const float4 q = *reinterpret_cast<const float4*>(p.in + row * p.D + c);
            v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) if (c + j < p.D) v[j] = p.in[row * p.D + c + j];
          }
        }
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) { float a, b; split_tf32(v[j], a, b); hi[j] = __float_as_uint(a); lo[j] = __float_as_uint(b); xrow[c + j] = v[j]; }
        tmem_st4(lane_base + COL_H_HI + c, hi);
        tmem_st4(lane_base + COL_H_LO + c, lo);
      }
      for (int j = 0; j < x_chunks; ++j) publish(x_chunk + j);
      float ladj = 0.f;
      for (int t = 0; t < p.T; ++t) {
        const float* bt = bias_s + (size_t)t * bias_per_t;
        for (int l = 0; l < p.L; ++l) {
          mbar_wait(&acc_full, ph_acc);
          ph_acc ^= 1;
          tc_fence_after();
          const uint32_t acc = lane_base + COL_ACC + (g & 1u) * 128u + part * CP;
          ++g;
          const float* bl = bt + l * H + part * CP;
          // this part's accumulator columns, 32 at a time: bias + residual + ReLU + hi/lo split, back to TMEM as a k-chunk of the
          // next layer's A operand, which the issuer multiplies into the OTHER accumulator while the next 32 are produced (their
          // TMEM load is already in flight)
          uint32_t buf[2][32];
          tmem_ld32_async(acc, buf[0]);
#pragma unroll
          for (int c = 0; c < CP; c += 32) {
            uint32_t (&cur)[32] = buf[(c >> 5) & 1];
            tmem_ld_fence32(cur);
            if (c + 32 < CP) tmem_ld32_async(acc + c + 32, buf[((c >> 5) + 1) & 1]);
#pragma unroll
            for (int q16 = 0; q16 < 2; ++q16) {
              uint32_t hi[16], lo[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                float a = __uint_as_float(cur[16 * q16 + j]) + bl[c + 16 * q16 + j];      // W h + b
                if (l > 0) a += hreg[c + 16 * q16 + j];                                    // residual block
                a = fmaxf(a, 0.f);
                hreg[c + 16 * q16 + j] = a;
                float x0, x1;
                split_tf32(a, x0, x1);
                hi[j] = __float_as_uint(x0); lo[j] = __float_as_uint(x1);
              }
              tmem_st16(lane_base + COL_H_HI + part * CP + c + 16 * q16, hi);
              tmem_st16(lane_base + COL_H_LO + part * CP + c + 16 * q16, lo);
            }
            publish(a_chunk + ((part * CP + c) >> 5));
          }
        }
        // output layer: phi[d] = (shift, scale_raw) -> y = x exp(ls) + shift, ladj += ls; 32-column steps dealt over the parts
        mbar_wait(&acc_full, ph_acc);
        ph_acc ^= 1;
        tc_fence_after();
        const uint32_t acc = lane_base + COL_ACC + (g & 1u) * 128u;
        ++g;
        const float* bo = bt + p.L * H;
        const bool last = (t == p.T - 1);
        for (int c = 32 * part; c < p.Nout; c += 32 * NPART) {   // 32 accumulator columns = 16 features
          uint32_t v[32];
          const int d0 = c >> 1;
          tmem_ld32_async(acc + c, v);
          tmem_ld_fence32(v);
          float y[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float r = 0.f;
            if (d0 + j < p.D) {
              const float x = xrow[d0 + j];
              const float shift = __uint_as_float(v[2 * j]) + bo[c + 2 * j];
              const float sraw = __uint_as_float(v[2 * j + 1]) + bo[c + 2 * j + 1];
              const float ls = sraw / (1.0f + fabsf(sraw / TC_LOG_SLOPE));
              ladj += ls;
              r = fmaf(x, expf(ls), shift);
            }
            y[j] = r;
          }
          if (last) {
            if (valid) {
              if (vec4) {
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4)
                  if (d0 + 4 * q4 < p.D)
                    *reinterpret_cast<float4*>(p.out + row * p.D + d0 + 4 * q4) = make_float4(y[4 * q4], y[4 * q4 + 1], y[4 * q4 + 2], y[4 * q4 + 3]);
              } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) if (d0 + j < p.D) p.out[row * p.D + d0 + j] = y[j];
              }
            }
          } else {
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              float x0, x1;
              split_tf32(y[j], x0, x1);
              hi[j] = __float_as_uint(x0); lo[j] = __float_as_uint(x1);
              if (d0 + j < TC_XLD - 1) xrow[d0 + j] = y[j];
            }
            tmem_st16(lane_base + COL_H_HI + d0, hi);             // next transform's layer-0 A operand
            tmem_st16(lane_base + COL_H_LO + d0, lo);
          }
        }
        if (!last) for (int j = 0; j < x_chunks; ++j) publish(x_chunk + j);
      }
      // the row's log-determinant: every part holds the terms of the features it mapped
      if (NPART > 1) {
        ladj_s[part][row_in_tile] = ladj;
        asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
        if (part == 0) {
#pragma unroll
          for (int q = 1; q < NPART; ++q) ladj += ladj_s[q][row_in_tile];
        }
      }
      if (part == 0 && valid) p.ladj[row] = ladj;
      if (NPART > 1) asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == EPI_WARPS) tmem_dealloc<512>(tm);
}

// hi/lo aware pack: gather[i] >= 0 -> hi(raw[g]) = nearest TF32; gather[i] <= -2 -> lo(raw[-g-2]) = nearest TF32 of the rest; -1 -> 0;
// entries flagged with bit 30 (biases) are copied unsplit
__global__ void pack_tc_kernel(const float* __restrict__ raw, const int* __restrict__ gather, float* __restrict__ packed, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int g = gather[i];
    float v = 0.f;
    if (g >= 0) {
      if (g & (1 << 30)) v = raw[g & ~(1 << 30)];
      else v = tf32_round(raw[g]);
    } else if (g <= -2) {
      const float x = raw[-g - 2];
      v = tf32_round(x - tf32_round(x));
    }
    packed[i] = v;
  }
}

}  // namespace pmc

using namespace pmc;

extern "C" int pmc_flow_tc_pack(const float* raw, const int32_t* gather, float* packed, int64_t n, pmc_stream_t stream) {
  PMC_REQUIRE(raw && gather && packed && n > 0, "pmc_flow_tc_pack: bad arguments");
  const int blocks = grid_for(n, 256, 8);
  pack_tc_kernel<<<blocks, 256, 0, as_stream(stream)>>>(raw, gather, packed, n);
  PMC_LAUNCH_CHECK();
  return 0;
}

extern "C" int pmc_flow_forward_tc(const float* packed, const int32_t* meta_host, int32_t meta_len, const float* in,
                                   float* out, float* ladj, int64_t n, int32_t passes, pmc_stream_t stream) {
  PMC_REQUIRE(packed && meta_host && in && out && ladj, "pmc_flow_forward_tc: null pointer");
  PMC_REQUIRE(meta_len >= TC_LEN && meta_host[TC_VERSION] == 101, "pmc_flow_forward_tc: not a tensor-core layout table");
  PMC_REQUIRE(passes == 1 || passes == 3, "pmc_flow_forward_tc: passes must be 1 (TF32) or 3 (3xTF32, fp32 fidelity)");
  if (n == 0) return 0;
  const int* m = meta_host;
  TcParams p;
  p.packed = packed; p.in = in; p.out = out; p.ladj = ladj; p.n = n;
  p.D = m[TC_D]; p.H = m[TC_H]; p.L = m[TC_L]; p.T = m[TC_T]; p.Kx = m[TC_KX]; p.Nout = m[TC_NOUT];
  p.tstride = m[TC_TSTRIDE]; p.bias_off = m[TC_BIAS_OFF]; p.passes = passes; p.slot_bytes = (uint32_t)m[TC_SLOT_BYTES];
  PMC_REQUIRE(m[TC_KIND] == 0, "pmc_flow_forward_tc: only affine (MAF) transforms are built for the tensor-core path");
  PMC_REQUIRE(p.D >= 2 && p.Kx % 8 == 0 && p.Kx >= p.D && p.Kx <= 48, "pmc_flow_forward_tc: n_dim out of range (2..48)");
  PMC_REQUIRE(p.Nout % 16 == 0 && p.Nout >= 2 * p.D && p.Nout <= 128, "pmc_flow_forward_tc: output width out of range");
  PMC_REQUIRE(p.L >= 1 && p.T >= 1, "pmc_flow_forward_tc: bad layer / transform count");
  PMC_REQUIRE(p.slot_bytes % 1024 == 0 && p.slot_bytes >= (uint32_t)(2 * 4 * std::min(TC_KCHUNK, p.H) * p.H),
              "pmc_flow_forward_tc: bad slot size");
  PMC_REQUIRE(p.tstride == p.bias_off + p.L * p.H + p.Nout, "pmc_flow_forward_tc: bias table missing from the packed image");
  const size_t tables = ((size_t)p.T * (p.L * p.H + p.Nout) + (size_t)128 * TC_XLD) * sizeof(float);
  p.stages = (int)std::min<size_t>(TC_MAX_STAGES, ((size_t)224 * 1024 - tables) / p.slot_bytes);
  PMC_REQUIRE(p.stages >= 2, "pmc_flow_forward_tc: weight ring does not fit shared memory");
  const size_t smem = (size_t)p.stages * p.slot_bytes + tables;
  const long long tiles = (n + 127) / 128;
  const int grid = (int)std::min<long long>(tiles, sm_count());
  cudaStream_t st = as_stream(stream);
#define PMC_TC_CASE(HV)                                                                              \
  case HV: {                                                                                         \
    auto kern = made_forward_tc_kernel<HV>;                                                          \
    PMC_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));     \
    kern<<<grid, tc_threads(HV), smem, st>>>(p);                                                                \
  } break;
  switch (p.H) {
    PMC_TC_CASE(32) PMC_TC_CASE(64) PMC_TC_CASE(128)
    default:
      set_error("pmc_flow_forward_tc: hidden width %d not built (32, 64, 128)", p.H);
      return 2;
  }
#undef PMC_TC_CASE
  PMC_LAUNCH_CHECK();
  return 0;
}

// Fused training step of Flow.fit (pocomc/flow.py:301-319) for zuko MAF: for one mini-batch computes the
// weighted negative log-likelihood and its gradient with respect to every flow parameter.
//
// The reference builds this with torch autograd over ~70 small ops per transform; the batch is at most
// 512 rows (sampler.py:289), so the step is launch- and latency-bound, not throughput-bound.  Here:
//
//   flow_train_fb_kernel<ROWS> : one CTA per ROWS = 8 / 16 / 32 batch rows runs the WHOLE forward chain (T transforms x (L+1)
//       masked linear layers + affine map) and the WHOLE input-gradient chain back, fp32 FMA.  Weight
//       images (mask folded in, transposed as each GEMM wants them; built by pmc_flow_pack from the
//       flat blob) stream through a double-buffered shared-memory slot with 1-D bulk copies
//       (cp.async.bulk + mbarrier), one image per layer, prefetched a layer ahead.  Activations live in
//       shared memory transposed ([feature][row]) so the inner product loop is one broadcast LDS.128
//       plus conflict-free LDS.32s per k.  Activations / pre-activation gradients are also written to a
//       global scratch for the weight-gradient pass.
//   flow_train_wgrad_kernel<CH> : grouped GEMM dW = dpre^T . input over the whole batch (CH-row chunks fetched one
//       ahead), one 32x32 tile of one layer per CTA, scattered straight into the flat gradient blob (no atomics,
//       deterministic).
//
// Arithmetic matches the autograd path operation for operation (fp32, same association up to the
// order of the k-loop); parity is tested against the oracle's gradients.
#include "common.cuh"
#include "tc_common.cuh"
#include <algorithm>

namespace pmc {

using namespace tc;

enum { TR_D = 0, TR_DP, TR_H, TR_L, TR_T, TR_NO, TR_TSTRIDE, TR_BIAS_OFF, TR_RAW_TSTRIDE, TR_MAP_TSTRIDE, TR_NTILES, TR_VERSION, TR_LEN };

constexpr int TR_PAD = 32;        // the padded batch is a multiple of this many rows
constexpr int TR_PART = 8;        // rows per loss partial (the finest CTA tile)
constexpr float TR_LOG_SLOPE_ABS = 6.90775527898213705205f;   // |log(1e-3)|, zuko MonotonicAffineTransform
constexpr float TR_HALF_LOG_2PI = 0.91893853320467274178f;

struct TrainParams {
  const float* packed;
  const float* xdata;        // [rows][D] training matrix
  const float* wdata;        // [rows] sample weights (weighted) or nullptr
  const long long* idx;      // [n_batches][Bp] row indices of every batch of the epoch
  const float* mask;         // [n_batches][Bp] 1 for real rows, 0 for padding
  const long long* cursor;   // device scalar: which batch
  float* X;                  // [T+1][Bp][Dp] transform inputs (X[T] = latent)
  float* S;                  // [T][Bp][Dp] raw log-scales
  float* Hs;                 // [T][L][Bp][H] hidden activations
  float* Gh;                 // [T][L][Bp][H] gradients w.r.t. hidden pre-activations
  float* Go;                 // [T][Bp][No] gradients w.r.t. the output layer (shift | scale_raw)
  double* loss_partials;     // [grid] per-CTA loss sums
  float* logprob;            // [Bp] per-row log-probability (may be nullptr)
  int Bp, D, Dp, H, L, T, No, tstride, bias_off;
  int weighted, backward;
  int cpb;                   // CTAs per batch (Bp / ROWS): a forward-only launch may cover several consecutive batches
  int ns;                    // weight ring depth (2..8 slots)
  int wslot;                 // floats per ring slot: the largest image, or less -- images then stream in k-chunks
  int splitk;                // experimental (PMC_TRAIN_SPLITK=1): hidden-width GEMMs split K over the column groups too
};

// acc[i][j] += sum_k At[k][4 rg + i] * W[k][c0 + CSTR j]   (k over the rows of one streamed chunk of the image, N floats per row)
template <int TN, int CSTR, int LDA>
__device__ __forceinline__ void gemm_tile(const float* __restrict__ At, int K, const float* __restrict__ W, int N, int rg, int c0,
                                          float (&acc)[4][8]) {
  const float* a = At + 4 * rg;
  const float* w = W + c0;
#pragma unroll 4
  for (int k = 0; k < K; ++k) {
    const float4 av = *reinterpret_cast<const float4*>(a);
    float wv[TN];
#pragma unroll
    for (int j = 0; j < TN; ++j) wv[j] = w[CSTR * j];
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      acc[0][j] = fmaf(av.x, wv[j], acc[0][j]);
      acc[1][j] = fmaf(av.y, wv[j], acc[1][j]);
      acc[2][j] = fmaf(av.z, wv[j], acc[2][j]);
      acc[3][j] = fmaf(av.w, wv[j], acc[3][j]);
    }
    a += LDA;
    w += N;
  }
}
template <int CSTR, int LDA>
__device__ __forceinline__ void gemm_any(int tn, const float* At, int K, const float* W, int N, int rg, int c0, float (&acc)[4][8]) {
  if (tn == 8) gemm_tile<8, CSTR, LDA>(At, K, W, N, rg, c0, acc);
  else if (tn == 4) gemm_tile<4, CSTR, LDA>(At, K, W, N, rg, c0, acc);
  else if (tn == 2) gemm_tile<2, CSTR, LDA>(At, K, W, N, rg, c0, acc);
  else gemm_tile<1, CSTR, LDA>(At, K, W, N, rg, c0, acc);
}

// ROWS batch rows per CTA.  The 8 warps form RG = ROWS/4 row groups (4 rows each, one broadcast LDS.128 per k) times
// CG = 8/RG column groups: a hidden layer's H output columns are dealt to the column groups in 32-wide tiles (tile index
// cg + CG j), so a smaller ROWS spreads one batch over more SMs at the same work per FFMA.  The two narrow GEMMs of a
// transform (output layer, N = 2 Dp; input gradient, N = Dp) keep shift and scale of a feature in one thread: there the
// column groups split K and the cg == 0 ("lead") warps sum the partials and run the affine maps.
template <int ROWS>
__global__ void __launch_bounds__(256, 1) flow_train_fb_kernel(const TrainParams p) {
  constexpr int TR_ROWS = ROWS, TR_LDA = ROWS + 4, RG = ROWS / 4, CG = 8 / RG, CSW = 32 * CG;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ uint64_t full[8];
  __shared__ float crow[TR_ROWS];
  __shared__ double lossrow[TR_ROWS];
  __shared__ float red[8];
  const int D = p.D, Dp = p.Dp, H = p.H, L = p.L, T = p.T, No = p.No, Bp = p.Bp;
  const int wmax = p.wslot;
  const int brows = max(H, No);
  const int NS = p.ns;
  float* wring = reinterpret_cast<float*>(smem_raw);
  float* bufA = wring + (size_t)NS * wmax;
  float* bufB = bufA + brows * TR_LDA;
  float* xT = bufB + brows * TR_LDA;
  float* gT = xT + Dp * TR_LDA;
  uint32_t* relu_bits = reinterpret_cast<uint32_t*>(gT + Dp * TR_LDA);   // [T*L][256]: (h > 0) of this thread's 4 x 8 outputs
  float* redbuf = reinterpret_cast<float*>(relu_bits + (size_t)T * L * 256);   // [(CG-1) ROWS][No] split-K partial sums

  const int tid = threadIdx.x, warp = tid >> 5, tx = tid & 31;
  const int ty = warp % RG, cg = warp / RG;                     // row group (rows 4 ty .. 4 ty + 3), column group
  const int cw0 = tx + 32 * cg;                                // first column of this thread in a hidden-width GEMM
  const bool lead = (cg == 0);                                 // warps that run the narrow GEMMs and the affine maps
  const int row0 = (blockIdx.x % p.cpb) * TR_ROWS;             // row inside the CTA's batch
  const int n_img = (p.backward ? 2 : 1) * T * (L + 1);
  const int fwd_img = T * (L + 1);
  const int bwd_base = (D + 1) * H + (L - 1) * (H + 1) * H + (H + 1) * No;

  // image s of the step: float offset, K rows of N floats, bias row behind the last weight row (forward images)
  auto img = [&](int s, int& off, int& K, int& N, int& bias) {
    if (s < fwd_img) {
      const int t = s / (L + 1), l = s - t * (L + 1);
      off = t * p.tstride + (l == 0 ? 0 : (D + 1) * H + (l - 1) * (H + 1) * H);
      K = (l == 0) ? D : H; N = (l < L) ? H : No; bias = 1;
    } else {
      const int q = s - fwd_img;
      const int tb = T - 1 - q / (L + 1), j = q % (L + 1);
      off = tb * p.tstride + bwd_base + (j == 0 ? 0 : No * H + (j - 1) * H * H);
      K = (j == 0) ? No : H; N = (j < L) ? H : Dp; bias = 0;
    }
  };
  // an image larger than a ring slot streams in chunks of kc rows (the bias rides with the last chunk)
  auto rows_per_chunk = [&](int K, int N, int bias) { return min(K, (wmax - (bias ? N : 0)) / N); };
  int pi = 0, pk = 0, pcount = 0;                              // producer cursor (thread 0): image, first row, chunks issued
  auto issue_next = [&]() {
    if (pi >= n_img) return;
    int off, K, N, bias;
    img(pi, off, K, N, bias);
    const int kc = rows_per_chunk(K, N, bias);
    const int rows = min(kc, K - pk);
    const bool last = pk + rows >= K;
    const uint32_t bytes = (uint32_t)(rows * N + ((last && bias) ? N : 0)) * 4u;
    const int slot = pcount % NS;
    mbar_expect_tx(full + slot, bytes);
    bulk_g2s(wring + (size_t)slot * wmax, p.packed + off + (size_t)pk * N, bytes, full + slot);
    ++pcount;
    if (last) { ++pi; pk = 0; } else pk += rows;
  };
  if (tid == 0) {
    for (int i = 0; i < NS; ++i) mbar_init(full + i, 1);
    mbar_fence_init();
    for (int i = 0; i < NS; ++i) issue_next();
  }
  int s = 0;                                                   // chunks consumed so far (ring position)
  float acc[4][8];
  // acc = in[32 rows x K] . image[K x 32 tn], the image arriving chunk by chunk; on return the LAST chunk is still in
  // its slot (bias at wl[rows_last * N]) and must be released with release() after the epilogue has read it
  const float* wl = nullptr;
  int rows_last = 0;
  auto release = [&]() {
    __syncthreads();                                           // every thread has read the slot (and, for callers, more)
    if (tid == 0) { fence_proxy_async(); issue_next(); }
    ++s;
  };
  // wide (N = H): every warp owns the columns cw0 + CSW j of its 4 rows.
  // narrow (output layer N = 2 Dp, input gradient N = Dp): shift and scale of a feature must meet in one thread, so
  // the columns stay whole (tx + 32 j) and the K rows of every chunk are dealt over the column groups instead
  // (split-K); the partial sums meet in `red` and the lead
  // warps (cg == 0) add them in group order, so the result does not depend on timing.
  const bool wide_splitk = (CG > 1) && p.splitk;
  auto stream_gemm = [&](bool wide, int tn, const float* in, int K, int N, int bias, float* red) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    const int kc = rows_per_chunk(K, N, bias);
    for (int k0 = 0; k0 < K; k0 += kc) {
      const int rows = min(kc, K - k0);
      mbar_wait(full + (s % NS), (s / NS) & 1);
      wl = wring + (size_t)(s % NS) * wmax;
      if (wide && !wide_splitk) {
        gemm_any<CSW, TR_LDA>(tn, in + k0 * TR_LDA, rows, wl, N, ty, cw0, acc);
      } else if (wide) {
        // every warp: ALL column tiles of its 4 rows (a 4 x 4 register tile per lane and k: one broadcast LDS.128 of
        // activations per 16 FFMA instead of per 4) over its quarter of the k rows
        const int lo = cg * rows / CG, hi = (cg + 1) * rows / CG;
        gemm_any<32, TR_LDA>(tn * CG, in + (k0 + lo) * TR_LDA, hi - lo, wl + lo * N, N, ty, tx, acc);
      } else {
        const int lo = cg * rows / CG, hi = (cg + 1) * rows / CG;
        gemm_any<32, TR_LDA>(tn, in + (k0 + lo) * TR_LDA, hi - lo, wl + lo * N, N, ty, tx, acc);
      }
      rows_last = rows;
      if (k0 + rows < K) release();
    }
    if (wide && wide_splitk) {
      // partial sums of every warp -> red[cg][column][row]; then warp (rg, cg) sums ITS column tiles cg + CG j over the
      // CG partials in group order (deterministic) and carries on as the owner of those columns, like the plain path
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (j >= tn * CG) continue;
        *reinterpret_cast<float4*>(red + ((size_t)(cg * N + tx + 32 * j) * TR_ROWS + 4 * ty)) =
            make_float4(acc[0][j], acc[1][j], acc[2][j], acc[3][j]);
      }
      __syncthreads();
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (j >= tn) continue;
        const int c = cw0 + CSW * j;
        float4 sum = *reinterpret_cast<const float4*>(red + ((size_t)c * TR_ROWS + 4 * ty));
        for (int c2 = 1; c2 < CG; ++c2) {
          const float4 v = *reinterpret_cast<const float4*>(red + ((size_t)(c2 * N + c) * TR_ROWS + 4 * ty));
          sum.x += v.x; sum.y += v.y; sum.z += v.z; sum.w += v.w;
        }
        acc[0][j] = sum.x; acc[1][j] = sum.y; acc[2][j] = sum.z; acc[3][j] = sum.w;
      }
    }
    if (!wide && CG > 1) {
      if (!lead) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (j >= tn) continue;
#pragma unroll
          for (int i = 0; i < 4; ++i) red[((cg - 1) * TR_ROWS + 4 * ty + i) * N + tx + 32 * j] = acc[i][j];
        }
      }
      __syncthreads();
      if (lead) {
        for (int c = 1; c < CG; ++c)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (j >= tn) continue;
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[i][j] += red[((c - 1) * TR_ROWS + 4 * ty + i) * N + tx + 32 * j];
          }
      }
    }
  };

  // ---- batch rows, loss coefficients c_r (flow.py:305-310) ----
  const long long cur = p.cursor[0] + blockIdx.x / p.cpb;
  const long long* bidx = p.idx + cur * Bp;
  const float* bmask = p.mask + cur * Bp;
  float sumw = 1.0f;
  if (p.weighted) {
    float s = 0.f;
    for (int r = tid; r < Bp; r += 256) s += p.wdata[bidx[r]] * bmask[r];
    s = warp_sum(s);
    if (tx == 0) red[warp] = s;
    __syncthreads();
    sumw = 0.f;
    for (int i = 0; i < 8; ++i) sumw += red[i];
  }
  if (tid < TR_ROWS) {
    const int r = row0 + tid;
    const float m = bmask[r];
    crow[tid] = p.weighted ? (p.wdata[bidx[r]] * m * 1000.0f / sumw) : m;
  }
  for (int i = tid; i < TR_ROWS * Dp; i += 256) {
    const int r = i / Dp, d = i - r * Dp;
    const float v = (d < D) ? p.xdata[bidx[row0 + r] * D + d] : 0.f;
    xT[d * TR_LDA + r] = v;
    if (p.backward) p.X[((size_t)0 * Bp + row0 + r) * Dp + d] = v;
  }
  __syncthreads();

  const int tnH = H / 32, tnO = No / 32, tnD = Dp / 32;
  const int tnW = tnH / CG;                                    // 32-wide column tiles of a hidden layer per thread
  float ladj_p[4] = {0.f, 0.f, 0.f, 0.f};
  float znew[4][2] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};
  // =========================== forward ===========================
  for (int t = 0; t < T; ++t) {
    float* in = xT;
    float* out = bufA;
    for (int l = 0; l < L; ++l) {
      stream_gemm(true, tnW, in, (l == 0) ? D : H, H, 1, redbuf);
      const float* bias = wl + rows_last * H;                 // bias rides behind the last weight rows in the same bulk copy
      float* hs = p.Hs + ((size_t)(t * L + l) * Bp + row0 + 4 * ty) * H;
      uint32_t bits = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (j >= tnW) continue;
        const int c = cw0 + CSW * j;
        const float b = bias[c];
        float h[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float v = acc[i][j] + b;
          if (l > 0) v += in[c * TR_LDA + 4 * ty + i];        // residual block
          h[i] = fmaxf(v, 0.f);
          bits |= (h[i] > 0.f ? 1u : 0u) << (8 * i + j);
          if (p.backward) hs[(size_t)i * H + c] = h[i];
        }
        *reinterpret_cast<float4*>(out + c * TR_LDA + 4 * ty) = make_float4(h[0], h[1], h[2], h[3]);
      }
      relu_bits[(t * L + l) * 256 + tid] = bits;
      release();
      in = out;
      out = (out == bufA) ? bufB : bufA;
    }
    // output layer + affine map (lead warps)
    stream_gemm(false, tnO, in, H, No, 1, redbuf);
    if (lead) {
      const float* bo = wl + rows_last * No;
#pragma unroll
      for (int jj = 0; jj < 2; ++jj) {
        if (jj >= tnD) continue;
        const int d = tx + 32 * jj;
        const float bs = bo[d], br = bo[Dp + d];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float z = 0.f;
          if (d < D) {
            const float shift = acc[i][jj] + bs, sraw = ((tnD == 1) ? acc[i][jj + 1] : acc[i][(jj + 2) & 3]) + br;
            const float ls = sraw / (1.0f + fabsf(sraw) / TR_LOG_SLOPE_ABS);
            z = fmaf(xT[d * TR_LDA + 4 * ty + i], expf(ls), shift);
            ladj_p[i] += ls;
            if (p.backward) p.S[((size_t)t * Bp + row0 + 4 * ty + i) * Dp + d] = sraw;
          }
          znew[i][jj] = z;
        }
      }
    }
    release();                                                 // every thread has read xT and the weight slot
    if (lead) {
#pragma unroll
      for (int jj = 0; jj < 2; ++jj) {
        if (jj >= tnD) continue;
        const int d = tx + 32 * jj;
        *reinterpret_cast<float4*>(xT + d * TR_LDA + 4 * ty) = make_float4(znew[0][jj], znew[1][jj], znew[2][jj], znew[3][jj]);
        if (p.backward)
#pragma unroll
          for (int i = 0; i < 4; ++i) p.X[((size_t)(t + 1) * Bp + row0 + 4 * ty + i) * Dp + d] = znew[i][jj];
      }
    }
    __syncthreads();
  }
  // ---- log-probability and loss: one partial sum per TR_PART rows, in row order ----
  if (lead) {
    float sq[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int jj = 0; jj < 2; ++jj) {
      if (jj >= tnD) continue;
#pragma unroll
      for (int i = 0; i < 4; ++i) sq[i] = fmaf(znew[i][jj], znew[i][jj], sq[i]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float sl = warp_sum(ladj_p[i]), ss = warp_sum(sq[i]);
      if (tx == 0) {
        const float lp = -0.5f * ss - (float)D * TR_HALF_LOG_2PI + sl;
        const int r = 4 * ty + i;
        lossrow[r] = (crow[r] != 0.f) ? -(double)crow[r] * (double)lp : 0.0;
        if (p.logprob) p.logprob[(size_t)(blockIdx.x / p.cpb) * Bp + row0 + r] = lp;
      }
    }
  }
  __syncthreads();
  if (tid < TR_ROWS / TR_PART) {
    double sum = 0.0;
    for (int r = 0; r < TR_PART; ++r) sum += lossrow[tid * TR_PART + r];
    p.loss_partials[(size_t)blockIdx.x * (TR_ROWS / TR_PART) + tid] = sum;
  }
  if (!p.backward) return;
  // =========================== backward ===========================
  // d loss / d z_T = c_r z ; d loss / d ladj = -c_r
  if (lead) {
#pragma unroll
    for (int jj = 0; jj < 2; ++jj) {
      if (jj >= tnD) continue;
      const int d = tx + 32 * jj;
#pragma unroll
      for (int i = 0; i < 4; ++i) gT[d * TR_LDA + 4 * ty + i] = crow[4 * ty + i] * znew[i][jj];
    }
  }
  __syncthreads();
  // transform inputs / raw log-scales of the affine backward are prefetched one transform ahead so the
  // global-memory latency never sits on the layer chain
  float xn[4][2], sn[4][2];
  auto prefetch_xs = [&](int t) {
    if (!lead) return;
#pragma unroll
    for (int jj = 0; jj < 2; ++jj) {
      if (jj >= tnD) continue;
      const int d = tx + 32 * jj;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const size_t o = ((size_t)t * Bp + row0 + 4 * ty + i) * Dp + d;
        xn[i][jj] = (d < D) ? p.X[o] : 0.f;
        sn[i][jj] = (d < D) ? p.S[o] : 0.f;
      }
    }
  };
#pragma unroll
  for (int jj = 0; jj < 2; ++jj)
#pragma unroll
    for (int i = 0; i < 4; ++i) { xn[i][jj] = 0.f; sn[i][jj] = 0.f; }
  prefetch_xs(T - 1);
  for (int t = T - 1; t >= 0; --t) {
    float gxd[4][2], xc[4][2], sc[4][2];
#pragma unroll
    for (int jj = 0; jj < 2; ++jj)
#pragma unroll
      for (int i = 0; i < 4; ++i) { xc[i][jj] = xn[i][jj]; sc[i][jj] = sn[i][jj]; gxd[i][jj] = 0.f; }
    if (t > 0) prefetch_xs(t - 1);
    // affine map backward -> gradient of the output layer (shift | scale_raw) as the next GEMM's input
    if (lead) {
#pragma unroll
      for (int jj = 0; jj < 2; ++jj) {
        if (jj >= tnD) continue;
        const int d = tx + 32 * jj;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = 4 * ty + i;
          float gshift = 0.f, gsraw = 0.f, gx = 0.f;
          if (d < D) {
            const float gz = gT[d * TR_LDA + r];
            const float x = xc[i][jj], sraw = sc[i][jj];
            const float den = 1.0f + fabsf(sraw) / TR_LOG_SLOPE_ABS;
            const float e = expf(sraw / den);
            gshift = gz;
            gsraw = (gz * x * e - crow[r]) / (den * den);
            gx = gz * e;
          }
          gxd[i][jj] = gx;
          bufA[d * TR_LDA + r] = gshift;
          bufA[(Dp + d) * TR_LDA + r] = gsraw;
          float* go = p.Go + ((size_t)t * Bp + row0 + r) * No;
          go[d] = gshift;
          go[Dp + d] = gsraw;
        }
      }
    }
    __syncthreads();
    float* in = bufA;
    float* out = bufB;
    for (int j = 0; j < L; ++j) {                              // images B_o, B_{L-1}, ..., B_1
      const int lh = L - 1 - j;                                // hidden layer whose pre-activation gradient comes out
      stream_gemm(true, tnW, in, j == 0 ? No : H, H, 0, redbuf);
      const uint32_t bits = relu_bits[(t * L + lh) * 256 + tid];
      float* gh = p.Gh + ((size_t)(t * L + lh) * Bp + row0 + 4 * ty) * H;
#pragma unroll
      for (int jn = 0; jn < 8; ++jn) {
        if (jn >= tnW) continue;
        const int c = cw0 + CSW * jn;
        float g[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float v = acc[i][jn];
          if (j > 0) v += in[c * TR_LDA + 4 * ty + i];         // residual path
          g[i] = ((bits >> (8 * i + jn)) & 1u) ? v : 0.f;       // ReLU
          gh[(size_t)i * H + c] = g[i];
        }
        *reinterpret_cast<float4*>(out + c * TR_LDA + 4 * ty) = make_float4(g[0], g[1], g[2], g[3]);
      }
      release();
      float* tmp = in; in = out; out = tmp;
    }
    // image B_0: gradient w.r.t. the transform input through the hyper-network + the direct path (lead warps)
    stream_gemm(false, tnD, in, H, Dp, 0, redbuf);
    if (lead) {
#pragma unroll
      for (int jj = 0; jj < 2; ++jj) {
        if (jj >= tnD) continue;
        const int d = tx + 32 * jj;
        *reinterpret_cast<float4*>(gT + d * TR_LDA + 4 * ty) =
            make_float4(gxd[0][jj] + acc[0][jj], gxd[1][jj] + acc[1][jj], gxd[2][jj] + acc[2][jj], gxd[3][jj] + acc[3][jj]);
      }
    }
    release();
  }
}

// dW_l[n][k] = sum_r dpre_l[r][n] * input_l[r][k], db_l[n] = sum_r dpre_l[r][n]; one 32 x 32 tile per CTA.
// The batch is walked in chunks of CH rows whose loads are issued one chunk ahead: the kernel is bound by the L2
// latency of those loads (the arithmetic of a chunk is ~0.5 us), so CH = 128 leaves 4 exposed round trips for a
// 512-row batch where CH = 32 had 16.
template <int CH>
__global__ void __launch_bounds__(256) flow_train_wgrad_kernel(const TrainParams p, const int* __restrict__ tiles,
                                                               const int* __restrict__ wmap, int map_tstride,
                                                               float* __restrict__ grad) {
  constexpr int Q = CH / 8;                                    // elements of each operand per thread per chunk
  __shared__ float dp[CH][33], in[CH][33];
  const int t = tiles[4 * blockIdx.x], l = tiles[4 * blockIdx.x + 1], n0 = tiles[4 * blockIdx.x + 2], k0 = tiles[4 * blockIdx.x + 3];
  const int D = p.D, H = p.H, L = p.L, No = p.No, Bp = p.Bp, Dp = p.Dp;
  const int n_img = (l < L) ? H : No, k_true = (l == 0) ? D : H;
  const float* dpre = (l < L) ? p.Gh + (size_t)(t * L + l) * Bp * H : p.Go + (size_t)t * Bp * No;
  const float* inp = (l == 0) ? p.X + (size_t)t * Bp * Dp : p.Hs + (size_t)(t * L + (l - 1)) * Bp * H;
  const int ldk = (l == 0) ? Dp : H;
  int moff = 0;
  for (int q = 0; q < l; ++q) moff += H * (q == 0 ? D : H) + H;
  const int* map = wmap + (size_t)t * map_tstride + moff;
  const int* bmap = map + n_img * k_true;
  const int tid = threadIdx.x, a = tid >> 4, b = tid & 15;
  float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}}, bacc[2] = {0.f, 0.f};
  float pd[Q], pi[Q];                                          // next chunk, fetched while the current one is reduced
  auto fetch = [&](int r0) {
#pragma unroll
    for (int q = 0; q < Q; ++q) {
      const int e = tid + 256 * q, rr = e >> 5, c = e & 31;
      const bool ok = r0 + rr < Bp;
      pd[q] = ok ? dpre[(size_t)(r0 + rr) * n_img + n0 + c] : 0.f;
      pi[q] = (ok && k0 + c < k_true) ? inp[(size_t)(r0 + rr) * ldk + k0 + c] : 0.f;
    }
  };
  fetch(0);
  for (int r0 = 0; r0 < Bp; r0 += CH) {
#pragma unroll
    for (int q = 0; q < Q; ++q) {
      const int e = tid + 256 * q, rr = e >> 5, c = e & 31;
      dp[rr][c] = pd[q];
      in[rr][c] = pi[q];
    }
    __syncthreads();
    if (r0 + CH < Bp) fetch(r0 + CH);
    const int nr = min(CH, Bp - r0);                           // a multiple of 32
#pragma unroll 8
    for (int rr = 0; rr < nr; ++rr) {
      const float d0 = dp[rr][2 * a], d1 = dp[rr][2 * a + 1], i0 = in[rr][2 * b], i1 = in[rr][2 * b + 1];
      acc[0][0] = fmaf(d0, i0, acc[0][0]); acc[0][1] = fmaf(d0, i1, acc[0][1]);
      acc[1][0] = fmaf(d1, i0, acc[1][0]); acc[1][1] = fmaf(d1, i1, acc[1][1]);
      bacc[0] += d0; bacc[1] += d1;
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int n = n0 + 2 * a + i;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int k = k0 + 2 * b + j;
      if (k < k_true) {
        const int m = map[n * k_true + k];
        if (m >= 0) grad[m] = acc[i][j];
      }
    }
    if (k0 == 0 && b == 0) {
      const int m = bmap[n];
      if (m >= 0) grad[m] = bacc[i];
    }
  }
}

static size_t train_smem(int rows, int splitk, int D, int Dp, int H, int No, int T, int L, int& ns, int& wslot) {
  const size_t TR_LDA = (size_t)rows + 4;
  const size_t wmax = (size_t)std::max(std::max((D + 1) * H, (H + 1) * H), std::max((H + 1) * No, H * Dp));
  const size_t brows = (size_t)std::max(H, No);
  const size_t act = (2 * brows * TR_LDA + 2 * (size_t)Dp * TR_LDA) * 4 + (size_t)T * L * 256 * 4 +
                     (splitk && rows < 32 ? (size_t)32 * brows * 4      // split-K partial sums: [CG][max(H, No)][rows]
                                          : (size_t)(32 - rows) * No * 4);   // narrow GEMMs only: [CG - 1][rows][No]
  const size_t budget = 200 * 1024;
  if (act + 2 * wmax * 4 <= budget) {
    // small networks: whole images, and a deeper ring lets the bulk copies run several layers ahead of the chain
    ns = (int)std::min<size_t>(8, (budget - act) / (wmax * 4));
    wslot = (int)wmax;
  } else {
    // wide networks (H = 256): images stream through 3 slots in k-chunks of whole rows
    ns = 3;
    const size_t widest = (size_t)std::max(H, std::max(No, Dp));
    wslot = (int)(((budget - act) / 3 / 4) / widest * widest);      // a multiple of every row length (H, No, Dp divide widest)
  }
  return (size_t)ns * wslot * 4 + act;
}

}  // namespace pmc

using namespace pmc;

// floats of scratch for a padded batch of Bp rows: X | S | Hs | Gh | Go
extern "C" int64_t pmc_flow_train_scratch_size(const int32_t* meta_host, int64_t bp) {
  const int64_t Dp = meta_host[TR_DP], H = meta_host[TR_H], L = meta_host[TR_L], T = meta_host[TR_T], No = meta_host[TR_NO];
  return (T + 1) * bp * Dp + T * bp * Dp + 2 * T * L * bp * H + T * bp * No;
}

static int train_launch(const float* packed, const int32_t* meta_host, int32_t meta_len, const float* xdata, const float* wdata,
                        const int64_t* idx, const float* mask, const int64_t* cursor, int64_t bp, int64_t n_batches,
                        float* scratch, double* loss_partials, float* logprob, const int32_t* tiles, const int32_t* wmap,
                        float* grad, int32_t backward, pmc_stream_t stream) {
  const int* m = meta_host;
  TrainParams p;
  p.packed = packed; p.xdata = xdata; p.wdata = wdata;
  p.idx = reinterpret_cast<const long long*>(idx); p.mask = mask; p.cursor = reinterpret_cast<const long long*>(cursor);
  p.Bp = (int)bp; p.D = m[TR_D]; p.Dp = m[TR_DP]; p.H = m[TR_H]; p.L = m[TR_L]; p.T = m[TR_T]; p.No = m[TR_NO];
  p.tstride = m[TR_TSTRIDE]; p.bias_off = m[TR_BIAS_OFF];
  p.weighted = wdata ? 1 : 0; p.backward = backward ? 1 : 0;
  PMC_REQUIRE((p.H == 32 || p.H == 64 || p.H == 128 || p.H == 256) && p.Dp % 32 == 0 && p.Dp <= 64 && p.No == 2 * p.Dp && p.D >= 2 && p.D <= p.Dp && p.L >= 1,
              "pmc_flow_train_step: unsupported flow shape");
  // Rows per CTA.  An optimiser step is ONE batch (<= 512 rows = 16 tiles of 32 rows on 148 SMs) and a chain of
  // T (L+1) dependent layers, so the smallest tile the hidden width can be dealt over (H/32 >= column groups) wins;
  // a validation pass over many batches already fills the machine and keeps 32-row tiles (4x less weight traffic).
  int rows = (p.H >= 128) ? 8 : (p.H == 64) ? 16 : 32;
  if (!backward && n_batches * (bp / 32) >= sm_count()) rows = 32;
  if (const char* e = getenv("PMC_TRAIN_ROWS")) {              // tuning override
    const int r = atoi(e);
    if ((r == 8 || r == 16 || r == 32) && (p.H / 32) % (32 / r) == 0) rows = r;
  }
  p.cpb = (int)(bp / rows);
  const size_t bpz = (size_t)bp;
  p.X = scratch;                                               // forward-only launches never touch the scratch
  p.S = p.X + (size_t)(p.T + 1) * bpz * p.Dp;
  p.Hs = p.S + (size_t)p.T * bpz * p.Dp;
  p.Gh = p.Hs + (size_t)p.T * p.L * bpz * p.H;
  p.Go = p.Gh + (size_t)p.T * p.L * bpz * p.H;
  p.loss_partials = loss_partials; p.logprob = logprob;
  p.splitk = 0;
  if (const char* e = getenv("PMC_TRAIN_SPLITK")) p.splitk = atoi(e) ? 1 : 0;   // experimental, not yet run on a GPU
  const size_t smem = train_smem(rows, p.splitk, p.D, p.Dp, p.H, p.No, p.T, p.L, p.ns, p.wslot);
  PMC_REQUIRE(smem <= 220 * 1024 && p.wslot >= 2 * std::max(p.H, p.No), "pmc_flow_train_step: shared memory budget exceeded");
  cudaStream_t st = as_stream(stream);
  const unsigned grid = (unsigned)(n_batches * p.cpb);
#define PMC_TRAIN_CASE(R)                                                                                           \
  if (rows == R) {                                                                                                  \
    PMC_TRY(cudaFuncSetAttribute(flow_train_fb_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    flow_train_fb_kernel<R><<<grid, 256, smem, st>>>(p);                                                            \
  }
  PMC_TRAIN_CASE(8) PMC_TRAIN_CASE(16) PMC_TRAIN_CASE(32)
#undef PMC_TRAIN_CASE
  PMC_LAUNCH_CHECK();
  if (backward) {
    int ch = (bp >= 128) ? 128 : 32;
    if (const char* e = getenv("PMC_WGRAD_CHUNK")) ch = atoi(e);     // tuning override
    if (ch == 128) flow_train_wgrad_kernel<128><<<(unsigned)m[TR_NTILES], 256, 0, st>>>(p, tiles, wmap, m[TR_MAP_TSTRIDE], grad);
    else if (ch == 64) flow_train_wgrad_kernel<64><<<(unsigned)m[TR_NTILES], 256, 0, st>>>(p, tiles, wmap, m[TR_MAP_TSTRIDE], grad);
    else flow_train_wgrad_kernel<32><<<(unsigned)m[TR_NTILES], 256, 0, st>>>(p, tiles, wmap, m[TR_MAP_TSTRIDE], grad);
    PMC_LAUNCH_CHECK();
  }
  return 0;
}

extern "C" int pmc_flow_train_step(const float* packed, const int32_t* meta_host, int32_t meta_len, const float* xdata,
                                   const float* wdata, const int64_t* idx, const float* mask, const int64_t* cursor, int64_t bp,
                                   float* scratch, double* loss_partials, float* logprob, const int32_t* tiles,
                                   const int32_t* wmap, float* grad, int32_t backward, pmc_stream_t stream) {
  PMC_REQUIRE(packed && meta_host && xdata && idx && mask && cursor && scratch && loss_partials, "pmc_flow_train_step: null pointer");
  PMC_REQUIRE(meta_len >= TR_LEN && meta_host[TR_VERSION] == 200, "pmc_flow_train_step: not a training layout table");
  PMC_REQUIRE(bp > 0 && bp % TR_PAD == 0, "pmc_flow_train_step: padded batch must be a multiple of 32 rows");
  PMC_REQUIRE(!backward || (tiles && wmap && grad), "pmc_flow_train_step: backward needs tiles, wmap and grad");
  return train_launch(packed, meta_host, meta_len, xdata, wdata, idx, mask, cursor, bp, 1, scratch, loss_partials, logprob, tiles,
                      wmap, grad, backward, stream);
}

extern "C" int pmc_flow_eval_batches(const float* packed, const int32_t* meta_host, int32_t meta_len, const float* xdata,
                                     const float* wdata, const int64_t* idx, const float* mask, const int64_t* cursor, int64_t bp,
                                     int64_t n_batches, double* loss_partials, float* logprob, pmc_stream_t stream) {
  PMC_REQUIRE(packed && meta_host && xdata && idx && mask && cursor && loss_partials, "pmc_flow_eval_batches: null pointer");
  PMC_REQUIRE(meta_len >= TR_LEN && meta_host[TR_VERSION] == 200, "pmc_flow_eval_batches: not a training layout table");
  PMC_REQUIRE(bp > 0 && bp % TR_PAD == 0 && n_batches >= 0, "pmc_flow_eval_batches: padded batch must be a multiple of 32 rows");
  if (n_batches == 0) return 0;
  return train_launch(packed, meta_host, meta_len, xdata, wdata, idx, mask, cursor, bp, n_batches, nullptr, loss_partials, logprob,
                      nullptr, nullptr, nullptr, 0, stream);
}

// Standalone probe (test infrastructure, not product): pins the tcgen05 descriptor semantics that
// pocomc_b200/csrc/flow_tc.cu relies on, on real hardware.  Build + run:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/tc_probe tests/tc_probe.cu && /tmp/tc_probe
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include "../pocomc_b200/csrc/tc_common.cuh"

using namespace pmc::tc;

constexpr int M = 128, N = 128, K = 32;

// mode 0: SS, LBO = k-chunk stride, SBO = 8-row-group stride (assumed semantics)
// mode 1: SS, fields swapped
// mode 2: TS (A from TMEM: lane = row, column = k), B as mode 0
// mode 3: TS, B as mode 1
// mode 4: SS timing loop (reps MMAs), mode 5: TS timing loop
__global__ void __launch_bounds__(128) probe(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ Dout,
                                             long long* cycles, int mode, int reps) {
  extern __shared__ __align__(1024) unsigned char smem[];
  float* sA = reinterpret_cast<float*>(smem);                  // [K/4][M][4]
  float* sB = sA + M * K;                                      // [K/4][N][4]
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int t = threadIdx.x, warp = t >> 5;
  if (t == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc<256>(&tmem_slot);
  // stage operands in the canonical no-swizzle K-major image: chunk (r, c) at c*(R*16) + r*16 bytes
  for (int i = t; i < M * K; i += 128) { const int r = i / K, k = i % K; sA[(k / 4) * (M * 4) + r * 4 + (k % 4)] = A[i]; }
  for (int i = t; i < N * K; i += 128) { const int r = i / K, k = i % K; sB[(k / 4) * (N * 4) + r * 4 + (k % 4)] = B[i]; }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_slot;
  const uint32_t acc = tm, atm = tm + 128;                     // accumulator columns [0,128), A-in-TMEM columns [128,160)
  if (mode == 2 || mode == 3 || mode == 5) {                   // thread = row: write own row of A into TMEM
    for (int c = 0; c < K; c += 4) {
      uint32_t r[4];
      for (int j = 0; j < 4; ++j) r[j] = __float_as_uint(A[t * K + c + j]);
      tmem_st4(atm + ((uint32_t)(warp * 32) << 16) + c, r);
    }
    tmem_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  long long t0 = 0, t1 = 0;
  if (t == 0) {
    tc_fence_after();
    const uint32_t id = idesc_tf32(M, N);
    const bool swapped = (mode == 1 || mode == 3);
    const uint32_t a_k = M * 16, b_k = N * 16, mn = 128;
    t0 = clock64();
    for (int rep = 0; rep < reps; ++rep) {
      for (int ks = 0; ks < K / 8; ++ks) {
        const uint32_t a_addr = smem_u32(sA) + ks * 2 * a_k, b_addr = smem_u32(sB) + ks * 2 * b_k;
        const uint64_t ad = swapped ? smem_desc(a_addr, mn, a_k) : smem_desc(a_addr, a_k, mn);
        const uint64_t bd = swapped ? smem_desc(b_addr, mn, b_k) : smem_desc(b_addr, b_k, mn);
        const uint32_t accum = (rep > 0 || ks > 0) ? 1u : 0u;
        if (mode == 2 || mode == 3 || mode == 5) mma_tf32_ts(acc, atm + ks * 8, bd, id, accum);
        else mma_tf32_ss(acc, ad, bd, id, accum);
      }
    }
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  if (t == 0) { t1 = clock64(); cycles[0] = t1 - t0; }
  tc_fence_after();
  for (int c = 0; c < N; c += 16) {
    float v[16];
    tmem_ld16(acc + ((uint32_t)(warp * 32) << 16) + c, v);
    for (int j = 0; j < 16; ++j) Dout[t * N + c + j] = v[j];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<256>(tm);
}

int main(int argc, char** argv) {
  const int sel = argc > 1 ? atoi(argv[1]) : -1;   // -1: everything; 0..3 one correctness mode; 6 rounding; 4/5 timing
  std::vector<float> A(M * K), B(N * K), D(M * N), ref(M * N);
  for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) A[m * K + k] = (float)((m * 3 + k * 7) % 11 - 5);
  for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) B[n * K + k] = (float)((n * 5 + k * 3) % 13 - 6);
  for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
    double s = 0; for (int k = 0; k < K; ++k) s += (double)A[m * K + k] * B[n * K + k];
    ref[m * N + n] = (float)s;
  }
  float *dA, *dB, *dD; long long* dC;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4); cudaMalloc(&dC, 8);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  const int smem = (M + N) * K * 4;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int mode = 0; mode < 4; ++mode) {
    if (sel >= 0 && sel != mode) continue;
    cudaMemset(dD, 0, D.size() * 4);
    probe<<<1, 128, smem>>>(dA, dB, dD, dC, mode, 1);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0; int bad = 0;
    for (int i = 0; i < M * N; ++i) { double d = fabs((double)D[i] - ref[i]); if (d > maxerr) maxerr = d; if (d > 0) ++bad; }
    printf("mode %d: %s maxerr %.3f mismatches %d / %d   D[0][0..3] = %g %g %g %g (ref %g %g %g %g)  D[9][17]=%g (ref %g)\n", mode,
           cudaGetErrorString(e), maxerr, bad, M * N, D[0], D[1], D[2], D[3], ref[0], ref[1], ref[2], ref[3], D[9 * N + 17], ref[9 * N + 17]);
    if (e != cudaSuccess) return 1;
  }
  // rounding probe: A = 1 + 2^-11 + 2^-13 everywhere (mantissa bits just below TF32's 10), B = 1
  if (sel < 0 || sel == 6) {
    const float a = 1.0f + ldexpf(1.0f, -11) + ldexpf(1.0f, -13);
    for (auto& x : A) x = a;
    for (auto& x : B) x = 1.0f;
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    for (int mode = 0; mode <= 2; mode += 2) {
      probe<<<1, 128, smem>>>(dA, dB, dD, dC, mode, 1);
      cudaDeviceSynchronize();
      cudaMemcpy(D.data(), dD, 16, cudaMemcpyDeviceToHost);
      printf("rounding (mode %d): sum_k a*1 over K=%d: got %.9g  truncate-> %.9g  round-nearest-> %.9g  exact-> %.9g\n", mode, K, D[0],
             (double)K, K * (1.0 + ldexp(1.0, -10)), K * (double)a);
    }
  }
  for (int mode = 4; mode <= 5; ++mode) {
    if (sel >= 0 && sel != mode) continue;
    for (int reps : {16, 64, 256}) {
      probe<<<1, 128, smem>>>(dA, dB, dD, dC, mode, reps);
      cudaError_t e = cudaDeviceSynchronize();
      long long c; cudaMemcpy(&c, dC, 8, cudaMemcpyDeviceToHost);
      printf("timing mode %d (%s) reps %d: %lld cycles, %.1f cycles per 128x128x8 MMA (%s)\n", mode, mode == 4 ? "SS" : "TS", reps, c,
             (double)c / (reps * (K / 8)), cudaGetErrorString(e));
    }
  }
  return 0;
}

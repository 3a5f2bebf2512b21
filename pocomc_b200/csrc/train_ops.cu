// Flow.fit optimiser step (pocomc/flow.py:268,314-319): torch.nn.utils.clip_grad_norm_ followed by
// torch.optim.AdamW.step over the flow's single flat parameter blob, fused into two launches.
// Hyper-parameters and the step counter live in device memory so the pair can sit inside a CUDA graph
// whose replay follows learning-rate changes (ReduceLROnPlateau, flow.py:271-277,361) without re-capture.
//
// HBM-bound: per parameter read g, p, m, v and write p, m, v = 28 bytes (+4 for the norm pass).
#include "common.cuh"
#include <algorithm>

namespace pmc {

enum { HY_LR = 0, HY_BETA1, HY_BETA2, HY_EPS, HY_WD, HY_CLIP, HY_LEN };   // keep in sync with flow.py
constexpr int NORM_BLOCKS = 296;                                           // 2 per SM; partials summed in fixed order

// book: optional bookkeeping of the training loop folded into block 0 (saves three tiny launches per optimiser step):
// loss_acc += sum(loss_partials[0..n_loss)) in index order, cursor += 1
struct StepBook {
  const double* loss_partials;
  int n_loss;
  double* loss_acc;
  long long* cursor;
};

__global__ void __launch_bounds__(256) grad_sqnorm_kernel(const float* __restrict__ g, long long n, double* __restrict__ partials,
                                                          long long* __restrict__ step, StepBook book) {
  double acc = 0.0;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const double v = (double)g[i];
    acc += v * v;
  }
  acc = warp_sum(acc);
  __shared__ double ws[8];
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += ws[i];
    partials[blockIdx.x] = s;
    if (blockIdx.x == 0) {
      step[0] += 1;                           // the AdamW step counter t (read by the update kernel)
      if (book.cursor) book.cursor[0] += 1;
    }
  }
  // loss of the step: block 0 sums the partials with all its threads (fixed assignment and tree, so the result does not
  // depend on timing); one thread walking them in index order cost ~3 us of dependent L2 round trips per step
  if (blockIdx.x == 0 && book.loss_acc) {
    double l = 0.0;
    for (int i = threadIdx.x; i < book.n_loss; i += blockDim.x) l += book.loss_partials[i];
    l = warp_sum(l);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = l;
    __syncthreads();
    if (threadIdx.x == 0) {
      double s = 0.0;
      for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += ws[i];
      book.loss_acc[0] += s;
    }
  }
}

__global__ void __launch_bounds__(256) adamw_clip_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                         float* __restrict__ v, long long n, const double* __restrict__ partials,
                                                         int n_partials, const double* __restrict__ hyper,
                                                         const long long* __restrict__ step, float* __restrict__ gnorm_out,
                                                         const int* __restrict__ pos_a, const int* __restrict__ pos_b,
                                                         float* __restrict__ image) {
  __shared__ float s_coef, s_step_size, s_bc2_sqrt;
  __shared__ double ws[8];
  {  // every block needs the total: all threads load (fixed assignment + tree: deterministic), not one thread x 296 loads
    double a = 0.0;
    for (int i = threadIdx.x; i < n_partials; i += blockDim.x) a += partials[i];
    a = warp_sum(a);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = a;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += ws[i];
    const float total = (float)sqrt(s);                          // torch.linalg.vector_norm result is fp32
    float coef = 1.0f;
    if (hyper[HY_CLIP] > 0.0) {
      coef = (float)hyper[HY_CLIP] / (total + 1e-6f);            // clip_grad_norm_: max_norm / (total_norm + 1e-6), clamped to 1
      coef = fminf(coef, 1.0f);
    }
    s_coef = coef;
    if (blockIdx.x == 0 && gnorm_out) gnorm_out[0] = total;
    const double t = (double)step[0];                            // bias corrections: two double pow per block, not per thread
    s_step_size = (float)(hyper[HY_LR] / (1.0 - pow(hyper[HY_BETA1], t)));
    s_bc2_sqrt = (float)sqrt(1.0 - pow(hyper[HY_BETA2], t));
  }
  __syncthreads();
  const float coef = s_coef;
  const double lr = hyper[HY_LR], b1 = hyper[HY_BETA1], b2 = hyper[HY_BETA2];
  const float step_size = s_step_size, bc2_sqrt = s_bc2_sqrt;
  const float eps = (float)hyper[HY_EPS];
  const float decay = (float)(1.0 - lr * hyper[HY_WD]);
  const float w1 = (float)(1.0 - b1), fb2 = (float)b2, w2 = (float)(1.0 - b2);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float gi = g[i] * coef;
    float pi = p[i] * decay;                                     // param.mul_(1 - lr * weight_decay)
    const float mi = m[i] + w1 * (gi - m[i]);                    // exp_avg.lerp_(grad, 1 - beta1)
    const float vi = fmaf(w2 * gi, gi, v[i] * fb2);              // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    pi -= step_size * (mi / denom);                              // param.addcdiv_(exp_avg, denom, value=-step_size)
    p[i] = pi; m[i] = mi; v[i] = vi;
    if (image) {                                                 // keep the kernel-side weight image in step (no separate pack)
      const int a = pos_a[i], b = pos_b[i];
      if (a >= 0) image[a] = pi;
      if (b >= 0) image[b] = pi;
    }
  }
}

}  // namespace pmc

using namespace pmc;

extern "C" int64_t pmc_adamw_scratch_size(void) { return NORM_BLOCKS; }

static int adamw_launch(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, const double* hyper,
                        int64_t* step, double* scratch, float* gnorm_out, StepBook book, const int* pos_a, const int* pos_b,
                        float* image, pmc_stream_t stream) {
  cudaStream_t st = as_stream(stream);
  int blocks = (int)std::min<long long>(NORM_BLOCKS, (n + 255) / 256);
  grad_sqnorm_kernel<<<blocks, 256, 0, st>>>(grad, n, scratch, reinterpret_cast<long long*>(step), book);
  PMC_LAUNCH_CHECK();
  const int ublocks = grid_for(n, 256 * 4, 8);
  adamw_clip_kernel<<<ublocks, 256, 0, st>>>(param, grad, exp_avg, exp_avg_sq, n, scratch, blocks, hyper,
                                              reinterpret_cast<const long long*>(step), gnorm_out, pos_a, pos_b, image);
  PMC_LAUNCH_CHECK();
  return 0;
}

extern "C" int pmc_adamw_clip_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                                   const double* hyper, int64_t* step, double* scratch, float* gnorm_out,
                                   pmc_stream_t stream) {
  PMC_REQUIRE(param && grad && exp_avg && exp_avg_sq && hyper && step && scratch && n > 0, "pmc_adamw_clip_step: bad arguments");
  return adamw_launch(param, grad, exp_avg, exp_avg_sq, n, hyper, step, scratch, gnorm_out, StepBook{nullptr, 0, nullptr, nullptr},
                      nullptr, nullptr, nullptr, stream);
}

extern "C" int pmc_adamw_clip_step_ex(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                                      const double* hyper, int64_t* step, double* scratch, float* gnorm_out,
                                      const double* loss_partials, int32_t n_loss, double* loss_acc, int64_t* cursor,
                                      const int32_t* pos_a, const int32_t* pos_b, float* image, pmc_stream_t stream) {
  PMC_REQUIRE(param && grad && exp_avg && exp_avg_sq && hyper && step && scratch && n > 0, "pmc_adamw_clip_step_ex: bad arguments");
  PMC_REQUIRE(!loss_acc || (loss_partials && n_loss >= 0), "pmc_adamw_clip_step_ex: loss accumulation needs the partial sums");
  PMC_REQUIRE(!image || (pos_a && pos_b), "pmc_adamw_clip_step_ex: the image update needs both position maps");
  return adamw_launch(param, grad, exp_avg, exp_avg_sq, n, hyper, step, scratch, gnorm_out,
                      StepBook{loss_partials, n_loss, loss_acc, reinterpret_cast<long long*>(cursor)}, pos_a, pos_b, image, stream);
}

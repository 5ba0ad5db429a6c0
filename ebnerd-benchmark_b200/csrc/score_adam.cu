// Click score + in-list softmax cross-entropy (reference nrms.py:201-202, 61-62) with its
// gradient, the scorer's sigmoid head (nrms.py:204-205), the Keras-form dense Adam
// (nrms.py:76-77) and the dropout-mask export used by the parity tests.
#include "ebk_common.cuh"

namespace ebk {
namespace {

constexpr int SC_MAXC = 32;  // candidates handled per impression per pass by the CE kernel

// one CTA (128 threads = 4 warps) per impression
// kind 0: categorical cross-entropy from the softmax's logits; kind 1: Keras binary_crossentropy on that output,
// which Keras evaluates as sigmoid cross-entropy of the SAME cached logits, averaged over the C candidates.
__global__ void __launch_bounds__(128) score_ce_kernel(int kind, int C, int D, const float* __restrict__ news,
                                                        const float* __restrict__ user,
                                                        const float* __restrict__ labels, float grad_scale, float loss_scale,
                                                        float* __restrict__ probs, float* __restrict__ loss_sum,
                                                        float* __restrict__ d_news, float* __restrict__ d_user) {
  extern __shared__ float sm[];  // z[C], dz[C]
  float* z = sm;
  float* dz = sm + C;
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* u = user + (long)b * D;
  const float* nb = news + (long)b * C * D;
  for (int c = warp; c < C; c += 4) {
    float acc = 0.0f;
    for (int d = lane; d < D; d += 32) acc = fmaf(nb[(long)c * D + d], u[d], acc);
    acc = warp_sum(acc);
    if (lane == 0) z[c] = acc;
  }
  __syncthreads();
  if (warp == 0) {
    float mx = -INFINITY;
    for (int c = lane; c < C; c += 32) mx = fmaxf(mx, z[c]);
    mx = warp_max(mx);
    float s = 0.0f, ysum = 0.0f, yz = 0.0f;
    for (int c = lane; c < C; c += 32) {
      s += expf(z[c] - mx);
      float y = labels[(long)b * C + c];
      ysum += y;
      yz = fmaf(y, z[c], yz);
    }
    s = warp_sum(s);
    ysum = warp_sum(ysum);
    yz = warp_sum(yz);
    float lse = mx + logf(s);
    float rs = 1.0f / s;
    float bce = 0.0f;
    const float invC = 1.0f / (float)C;
    for (int c = lane; c < C; c += 32) {
      float p = expf(z[c] - mx) * rs;
      probs[(long)b * C + c] = p;
      const float y = labels[(long)b * C + c];
      if (kind == 0) {
        dz[c] = (p * ysum - y) * grad_scale;
      } else {
        // sigmoid_cross_entropy_with_logits: max(z,0) - z*y + log(1 + exp(-|z|)); d/dz = sigmoid(z) - y
        const float zc = z[c];
        bce += fmaxf(zc, 0.0f) - zc * y + log1pf(expf(-fabsf(zc)));
        dz[c] = (1.0f / (1.0f + expf(-zc)) - y) * invC * grad_scale;
      }
    }
    if (kind != 0) bce = warp_sum(bce);
    if (lane == 0) atomicAdd(loss_sum, (kind == 0 ? (lse * ysum - yz) : bce * invC) * loss_scale);
  }
  if (d_news == nullptr) return;
  __syncthreads();
  for (int d = threadIdx.x; d < D; d += 128) {
    float ud = u[d];
    float acc = 0.0f;
    for (int c = 0; c < C; ++c) {
      float g = dz[c];
      d_news[((long)b * C + c) * D + d] = g * ud;
      acc = fmaf(g, nb[(long)c * D + d], acc);
    }
    d_user[(long)b * D + d] = acc;
  }
}

// one warp per (b, c)
__global__ void score_sigmoid_kernel(long BC, int C, int D, const float* __restrict__ news,
                                     const float* __restrict__ user, float* __restrict__ out) {
  long item = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (item >= BC) return;
  long b = item / C;
  const float* nrow = news + item * D;
  const float* u = user + b * D;
  float acc = 0.0f;
  for (int d = lane; d < D; d += 32) acc = fmaf(nrow[d], u[d], acc);
  acc = warp_sum(acc);
  if (lane == 0) out[item] = 1.0f / (1.0f + expf(-acc));
}

// Keras-form Adam, 128-bit vectorised, grid-stride; pure HBM streaming (7 floats moved / param).
__global__ void __launch_bounds__(256) adam_keras_kernel(float4* __restrict__ theta, float4* __restrict__ g,
                                                          float4* __restrict__ m, float4* __restrict__ v,
                                                          size_t n4, float alpha, const float* __restrict__ alpha_dev,
                                                          float omb1, float omb2, float eps, int zero_grad) {
  if (alpha_dev != nullptr) alpha = __ldg(alpha_dev);   // CUDA-graph replay: this step's alpha lives in device memory
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 th = theta[i], gg = g[i], mm = m[i], vv = v[i];
#define UPD(c)                                  \
  mm.c += (gg.c - mm.c) * omb1;                 \
  vv.c += (gg.c * gg.c - vv.c) * omb2;          \
  th.c -= (mm.c * alpha) / (sqrtf(vv.c) + eps);
    UPD(x) UPD(y) UPD(z) UPD(w)
#undef UPD
    theta[i] = th;
    m[i] = mm;
    v[i] = vv;
    if (zero_grad) g[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}
__global__ void adam_keras_tail_kernel(float* theta, float* g, float* m, float* v, size_t beg, size_t n,
                                       float alpha, const float* alpha_dev, float omb1, float omb2, float eps,
                                       int zero_grad) {
  if (alpha_dev != nullptr) alpha = *alpha_dev;
  size_t i = beg + blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  float gg = g[i];
  float mm = m[i] + (gg - m[i]) * omb1;
  float vv = v[i] + (gg * gg - v[i]) * omb2;
  theta[i] -= (mm * alpha) / (sqrtf(vv) + eps);
  m[i] = mm;
  v[i] = vv;
  if (zero_grad) g[i] = 0.0f;
}

__global__ void dropout_mask_kernel(uint64_t seed, uint32_t thr, size_t n, float* out) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = (thr == 0 || dropout_keep(seed, i, thr)) ? 1.0f : 0.0f;
}

}  // namespace
}  // namespace ebk

using namespace ebk;

extern "C" int ebk_score_loss(int32_t kind, int32_t B, int32_t C, int32_t D, const float* news, const float* user,
                              const float* labels, float grad_scale, float loss_scale, float* probs, float* loss_sum,
                              float* d_news, float* d_user, void* stream) {
  if (B <= 0) return EBK_OK;
  EBK_CHECK_ARG(kind == EBK_LOSS_CATEGORICAL_CE || kind == EBK_LOSS_BINARY_CE, "score_loss: unknown loss kind %d", kind);
  EBK_CHECK_ARG(C >= 1 && D >= 1, "score_loss: bad shape C=%d D=%d", C, D);
  EBK_CHECK_ARG(news && user && labels && probs && loss_sum, "score_loss: null pointer");
  EBK_CHECK_ARG((d_news == nullptr) == (d_user == nullptr), "score_loss: d_news/d_user must both be set or both NULL");
  EBK_CHECK_ARG(C <= 4096, "score_loss: C=%d > 4096", C);
  (void)SC_MAXC;
  cudaStream_t st = (cudaStream_t)stream;
  if (prof_on()) prof_begin(T_SCORE, st);
  score_ce_kernel<<<B, 128, 2 * C * sizeof(float), st>>>(kind, C, D, news, user, labels, grad_scale, loss_scale, probs,
                                                         loss_sum, d_news, d_user);
  if (prof_on()) prof_end(T_SCORE, st);
  EBK_LAUNCH_CHECK();
  return EBK_OK;
}

extern "C" int ebk_score_softmax_ce(int32_t B, int32_t C, int32_t D, const float* news, const float* user,
                                    const float* labels, float loss_scale, float* probs, float* loss_sum,
                                    float* d_news, float* d_user, void* stream) {
  return ebk_score_loss(EBK_LOSS_CATEGORICAL_CE, B, C, D, news, user, labels, loss_scale, loss_scale, probs, loss_sum,
                        d_news, d_user, stream);
}

extern "C" int ebk_score_sigmoid(int32_t B, int32_t C, int32_t D, const float* news, const float* user,
                                 float* out, void* stream) {
  if (B <= 0 || C <= 0) return EBK_OK;
  EBK_CHECK_ARG(news && user && out && D >= 1, "score_sigmoid: bad argument");
  long BC = (long)B * C;
  long threads = BC * 32;
  score_sigmoid_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(BC, C, D, news, user,
                                                                                            out);
  EBK_LAUNCH_CHECK();
  return EBK_OK;
}

extern "C" int ebk_adam_keras_step(float* theta, float* g, float* m, float* v, size_t n, float alpha,
                                   double beta1, double beta2, float eps, int zero_grad, void* stream) {
  return ebk_adam_keras_step_p(theta, g, m, v, n, alpha, nullptr, beta1, beta2, eps, zero_grad, stream);
}

extern "C" int ebk_adam_keras_step_p(float* theta, float* g, float* m, float* v, size_t n, float alpha,
                                     const ebk_step_params* step_dev, double beta1, double beta2, float eps,
                                     int zero_grad, void* stream) {
  if (n == 0) return EBK_OK;
  const float* alpha_dev = step_dev ? &step_dev->alpha : nullptr;
  EBK_CHECK_ARG(theta && g && m && v, "adam: null pointer");
  EBK_CHECK_ARG(((uintptr_t)theta % 16 == 0) && ((uintptr_t)g % 16 == 0) && ((uintptr_t)m % 16 == 0) &&
                    ((uintptr_t)v % 16 == 0),
                "adam: buffers must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const float omb1 = (float)(1.0 - beta1), omb2 = (float)(1.0 - beta2);
  size_t n4 = n / 4;
  if (prof_on()) prof_begin(T_ADAM, st);
  if (n4) {
    size_t blocks = (n4 + 255) / 256;
    size_t cap = 148 * 16;
    adam_keras_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, st>>>(
        reinterpret_cast<float4*>(theta), reinterpret_cast<float4*>(g), reinterpret_cast<float4*>(m),
        reinterpret_cast<float4*>(v), n4, alpha, alpha_dev, omb1, omb2, eps, zero_grad);
    EBK_LAUNCH_CHECK();
  }
  if (n4 * 4 < n) {
    adam_keras_tail_kernel<<<1, 32, 0, st>>>(theta, g, m, v, n4 * 4, n, alpha, alpha_dev, omb1, omb2, eps, zero_grad);
    EBK_LAUNCH_CHECK();
  }
  if (prof_on()) prof_end(T_ADAM, st);
  return EBK_OK;
}

extern "C" int ebk_dropout_mask(uint64_t seed, float p, size_t n, float* out, void* stream) {
  if (n == 0) return EBK_OK;
  EBK_CHECK_ARG(out && p >= 0.0f && p < 1.0f, "dropout_mask: bad argument");
  uint32_t thr = p > 0.0f ? dropout_threshold(p) : 0;
  dropout_mask_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(seed, thr, n, out);
  EBK_LAUNCH_CHECK();
  return EBK_OK;
}

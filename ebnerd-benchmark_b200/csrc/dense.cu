// Dense(+ReLU) -> BatchNormalization -> Dropout layer of the NRMSDocVec news encoder
// (reference nrms_docvec.py:118-130) forward and backward.  The contraction runs on the tcgen05 GEMM;
// this file holds the HBM-bound elementwise / column-reduction kernels around it:
//   fwd: a = relu(x W + b);  [mean, var over the rows of THIS call];  y = dropout(gamma (a-mean)/sqrt(var+eps) + beta)
//   bwd: dyd = dropout'(dy); dgamma, dbeta; da (training-mode BN backward); dz = da [a>0]; db; dW = x^T dz (+2 l2 W); dx = dz W^T
// Keras semantics: biased batch variance for both normalisation and the moving average; moving stats
// updated per call: m = m*momentum + batch*(1-momentum).
#include "ebk_common.cuh"

namespace ebk {
namespace {

constexpr int ROWS_PER_BLK = 64;   // rows per partial block
constexpr int RB_X = 32, RB_Y = 8;   // reduction blocks: 32 columns (or float4 column groups) x 8 row slices

// a = act(z + b) in place, 128-bit
__global__ void bias_act_kernel(float4* __restrict__ z, const float4* __restrict__ b, long n4, int U4, int relu) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 v = z[i];
  const float4 bb = b[i % U4];
  v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
  if (relu) {
    v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
  }
  z[i] = v;
}

// partial[blk][0][j] = sum_r a[r,j], partial[blk][1][j] = sum_r a[r,j]^2 over the block's rows
__device__ __forceinline__ float4 f4add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
// Column reductions of this file: block (32 float4 column groups) x (8 row slices) over ROWS_PER_BLK rows, slices
// combined through shared memory in a fixed order (deterministic), one partial row per block; the final kernels
// reduce the partial rows the same way (32 columns x 8 slices).
__global__ void __launch_bounds__(RB_X * RB_Y) col_stats_partial_kernel(const float* __restrict__ a, int N, int U,
                                                                        float* __restrict__ partial) {
  __shared__ float4 sh[2][RB_Y][RB_X];
  const int U4 = U >> 2, c4 = blockIdx.x * RB_X + threadIdx.x;
  const int r0 = blockIdx.y * ROWS_PER_BLK, r1 = min(N, r0 + ROWS_PER_BLK);
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = s;
  if (c4 < U4) {
#pragma unroll 4
    for (int r = r0 + threadIdx.y; r < r1; r += RB_Y) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(a) + (long)r * U4 + c4);
      s = f4add(s, v);
      q.x = fmaf(v.x, v.x, q.x); q.y = fmaf(v.y, v.y, q.y); q.z = fmaf(v.z, v.z, q.z); q.w = fmaf(v.w, v.w, q.w);
    }
  }
  sh[0][threadIdx.y][threadIdx.x] = s;
  sh[1][threadIdx.y][threadIdx.x] = q;
  __syncthreads();
  if (threadIdx.y == 0 && c4 < U4) {
    for (int k = 1; k < RB_Y; ++k) {
      s = f4add(s, sh[0][k][threadIdx.x]);
      q = f4add(q, sh[1][k][threadIdx.x]);
    }
    reinterpret_cast<float4*>(partial + ((long)blockIdx.y * 2 + 0) * U)[c4] = s;
    reinterpret_cast<float4*>(partial + ((long)blockIdx.y * 2 + 1) * U)[c4] = q;
  }
}
// mean / invstd of this call and the Keras moving-average update
__global__ void __launch_bounds__(RB_X * RB_Y) col_stats_final_kernel(const float* __restrict__ partial, int nblk, int N, int U,
                                                                      float eps, float momentum, float* __restrict__ mean,
                                                                      float* __restrict__ invstd, float* __restrict__ mov_mean,
                                                                      float* __restrict__ mov_var) {
  __shared__ double sh[2][RB_Y][RB_X];
  const int j = blockIdx.x * RB_X + threadIdx.x;
  double s = 0.0, q = 0.0;
  if (j < U)
    for (int b = threadIdx.y; b < nblk; b += RB_Y) {
      s += partial[((long)b * 2 + 0) * U + j];
      q += partial[((long)b * 2 + 1) * U + j];
    }
  sh[0][threadIdx.y][threadIdx.x] = s;
  sh[1][threadIdx.y][threadIdx.x] = q;
  __syncthreads();
  if (threadIdx.y != 0 || j >= U) return;
  for (int k = 1; k < RB_Y; ++k) {
    s += sh[0][k][threadIdx.x];
    q += sh[1][k][threadIdx.x];
  }
  const double m = s / N;
  double var = q / N - m * m;
  var = var < 0.0 ? 0.0 : var;
  mean[j] = (float)m;
  invstd[j] = (float)(1.0 / sqrt(var + (double)eps));
  mov_mean[j] = mov_mean[j] * momentum + (float)m * (1.0f - momentum);
  mov_var[j] = mov_var[j] * momentum + (float)var * (1.0f - momentum);
}
__global__ void inference_stats_kernel(int U, float eps, const float* __restrict__ mov_mean,
                                       const float* __restrict__ mov_var, float* __restrict__ mean,
                                       float* __restrict__ invstd) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= U) return;
  mean[j] = mov_mean[j];
  invstd[j] = rsqrtf(mov_var[j] + eps);
}
// y = dropout(gamma * (a - mean) * invstd + beta)
__global__ void bn_apply_kernel(const float4* __restrict__ a, long n4, int U4, const float4* __restrict__ mean,
                                const float4* __restrict__ invstd, const float4* __restrict__ gamma,
                                const float4* __restrict__ beta, Dropout drop, float4* __restrict__ y) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const int c = (int)(i % U4);
  const float4 v = a[i], m = mean[c], is = invstd[c], g = gamma[c], b = beta[c];
  float4 o;
  o.x = fmaf((v.x - m.x) * is.x, g.x, b.x);
  o.y = fmaf((v.y - m.y) * is.y, g.y, b.y);
  o.z = fmaf((v.z - m.z) * is.z, g.z, b.z);
  o.w = fmaf((v.w - m.w) * is.w, g.w, b.w);
  if (drop.on()) {
    const float4 f = drop.factor4_group((uint64_t)i);
    o.x *= f.x; o.y *= f.y; o.z *= f.z; o.w *= f.w;
  }
  y[i] = o;
}

// partial sums of dyd and dyd*xhat per column (dyd = dy * dropout factor)
__global__ void __launch_bounds__(RB_X * RB_Y) bn_bwd_partial_kernel(const float* __restrict__ dy, const float* __restrict__ a,
                                                                     int N, int U, const float* __restrict__ mean,
                                                                     const float* __restrict__ invstd, Dropout drop,
                                                                     float* __restrict__ partial) {
  __shared__ float4 sh[2][RB_Y][RB_X];
  const int U4 = U >> 2, c4 = blockIdx.x * RB_X + threadIdx.x;
  const int r0 = blockIdx.y * ROWS_PER_BLK, r1 = min(N, r0 + ROWS_PER_BLK);
  float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
  if (c4 < U4) {
    const float4 m = __ldg(reinterpret_cast<const float4*>(mean) + c4), is = __ldg(reinterpret_cast<const float4*>(invstd) + c4);
#pragma unroll 4
    for (int r = r0 + threadIdx.y; r < r1; r += RB_Y) {
      const long i4 = (long)r * U4 + c4;
      float4 g = __ldg(reinterpret_cast<const float4*>(dy) + i4);
      const float4 av = __ldg(reinterpret_cast<const float4*>(a) + i4);
      if (drop.on()) {
        const float4 f = drop.factor4_group((uint64_t)i4);   // U % 4 == 0: group i4 = elements 4 i4 .. 4 i4 + 3
        g.x *= f.x; g.y *= f.y; g.z *= f.z; g.w *= f.w;
      }
      s1 = f4add(s1, g);
      s2.x = fmaf(g.x, (av.x - m.x) * is.x, s2.x); s2.y = fmaf(g.y, (av.y - m.y) * is.y, s2.y);
      s2.z = fmaf(g.z, (av.z - m.z) * is.z, s2.z); s2.w = fmaf(g.w, (av.w - m.w) * is.w, s2.w);
    }
  }
  sh[0][threadIdx.y][threadIdx.x] = s1;
  sh[1][threadIdx.y][threadIdx.x] = s2;
  __syncthreads();
  if (threadIdx.y == 0 && c4 < U4) {
    for (int k = 1; k < RB_Y; ++k) {
      s1 = f4add(s1, sh[0][k][threadIdx.x]);
      s2 = f4add(s2, sh[1][k][threadIdx.x]);
    }
    reinterpret_cast<float4*>(partial + ((long)blockIdx.y * 2 + 0) * U)[c4] = s1;
    reinterpret_cast<float4*>(partial + ((long)blockIdx.y * 2 + 1) * U)[c4] = s2;
  }
}
__global__ void __launch_bounds__(RB_X * RB_Y) bn_bwd_final_kernel(const float* __restrict__ partial, int nblk, int U,
                                                                   float* __restrict__ sums, float* __restrict__ dgamma,
                                                                   float* __restrict__ dbeta) {
  __shared__ float sh[2][RB_Y][RB_X];
  const int j = blockIdx.x * RB_X + threadIdx.x;
  float s1 = 0.f, s2 = 0.f;
  if (j < U)
    for (int b = threadIdx.y; b < nblk; b += RB_Y) {
      s1 += partial[((long)b * 2 + 0) * U + j];
      s2 += partial[((long)b * 2 + 1) * U + j];
    }
  sh[0][threadIdx.y][threadIdx.x] = s1;
  sh[1][threadIdx.y][threadIdx.x] = s2;
  __syncthreads();
  if (threadIdx.y != 0 || j >= U) return;
  for (int k = 1; k < RB_Y; ++k) {
    s1 += sh[0][k][threadIdx.x];
    s2 += sh[1][k][threadIdx.x];
  }
  sums[j] = s1;
  sums[U + j] = s2;
  dbeta[j] += s1;
  dgamma[j] += s2;
}
// dz = relu'(a) * invstd * gamma * (dyd - s1/N - xhat * s2/N)      (training-mode BN backward), 4 columns per thread
__global__ void bn_bwd_apply_kernel(const float4* __restrict__ dy, const float4* __restrict__ a, long n4, int N, int U4,
                                    const float4* __restrict__ mean, const float4* __restrict__ invstd,
                                    const float4* __restrict__ gamma, const float4* __restrict__ sums, Dropout drop,
                                    int relu, int round_out, float4* __restrict__ dz) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const int c = (int)(i % U4);
  float4 g = __ldg(dy + i);
  if (drop.on()) {
    const float4 f = drop.factor4_group((uint64_t)i);
    g.x *= f.x; g.y *= f.y; g.z *= f.z; g.w *= f.w;
  }
  const float4 av = __ldg(a + i), m = __ldg(mean + c), is = __ldg(invstd + c), ga = __ldg(gamma + c);
  const float4 s1 = __ldg(sums + c), s2 = __ldg(sums + U4 + c);
  const float invN = 1.0f / (float)N;
  auto one = [&](float gv, float avv, float mv, float isv, float gav, float s1v, float s2v) {
    const float xhat = (avv - mv) * isv;
    float v = isv * gav * (gv - s1v * invN - xhat * s2v * invN);
    if (relu && !(avv > 0.f)) v = 0.f;
    return round_out ? round_tf32_bits(v) : v;
  };
  float4 o;
  o.x = one(g.x, av.x, m.x, is.x, ga.x, s1.x, s2.x);
  o.y = one(g.y, av.y, m.y, is.y, ga.y, s1.y, s2.y);
  o.z = one(g.z, av.z, m.z, is.z, ga.z, s1.z, s2.z);
  o.w = one(g.w, av.w, m.w, is.w, ga.w, s1.w, s2.w);
  dz[i] = o;
}
// no BN: dz = dy * dropout' * relu'(y), 4 elements per thread
__global__ void act_bwd_kernel(const float4* __restrict__ dy, const float4* __restrict__ yv, long n4, Dropout drop, int relu,
                               int round_out, float4* __restrict__ dz) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 g = __ldg(dy + i);
  if (drop.on()) {
    const float4 f = drop.factor4_group((uint64_t)i);
    g.x *= f.x; g.y *= f.y; g.z *= f.z; g.w *= f.w;
  }
  if (relu) {
    const float4 y = __ldg(yv + i);
    if (!(y.x > 0.f)) g.x = 0.f;
    if (!(y.y > 0.f)) g.y = 0.f;
    if (!(y.z > 0.f)) g.z = 0.f;
    if (!(y.w > 0.f)) g.w = 0.f;
  }
  if (round_out) { g.x = round_tf32_bits(g.x); g.y = round_tf32_bits(g.y); g.z = round_tf32_bits(g.z); g.w = round_tf32_bits(g.w); }
  dz[i] = g;
}
__global__ void axpy_kernel(float* __restrict__ y, const float* __restrict__ x, long n, float a) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i < n) y[i] = fmaf(a, x[i], y[i]);
}
__global__ void sumsq_kernel(const float* __restrict__ x, long n, float scale, float* __restrict__ out) {
  float acc = 0.f;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) acc = fmaf(x[i], x[i], acc);
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) atomicAdd(out, acc * scale);
}

struct DenseWs {
  float *a, *dz, *mean, *invstd, *sums, *partial, *w_f, *w_d, *colsum;
  float *xr, *wr;   // all-TMA path (EBK_MATH_TF32): tf32-rounded copies of the layer input [N, K] and of W [K, U]
  size_t bytes;
};
DenseWs dense_layout(const ebk_dense_desc& d, void* base) {
  size_t off = 0;
  auto take = [&](size_t nfloat) {
    float* p = base ? reinterpret_cast<float*>(reinterpret_cast<char*>(base) + off) : nullptr;
    off += align_up(nfloat * sizeof(float), 256);
    return p;
  };
  const size_t N = d.N, U = d.U;
  const size_t nblk = (N + ROWS_PER_BLK - 1) / ROWS_PER_BLK;
  DenseWs w;
  w.a = take(N * U);
  w.dz = take(N * U);
  w.mean = take(U);
  w.invstd = take(U);
  w.sums = take(2 * U);
  w.partial = take(nblk * 2 * U);
  w.w_f = take(gemm_tf32_packed_floats(d.U, d.K, false));
  w.w_d = take(gemm_tf32_packed_floats(d.K, d.U, true));
  w.colsum = take(colsum_partial_floats((int)N, (int)U));
  w.xr = take(N * (size_t)d.K);
  w.wr = take((size_t)d.K * U);
  w.bytes = off;
  return w;
}
// EBK_MATH_TF32 layers run their three contractions on the all-TMA tcgen05 GEMM (gemm_tma_sm100.cu) from rounded copies of
// x and W kept in the workspace (forward -> backward); EBK_DENSE_TMA=0 keeps the older register-staged GEMM (A/B runs).
bool dense_uses_tma(const ebk_dense_desc& d, const DenseWs& ws, const float* x, const float* W) {
  static const bool on = !(getenv("EBK_DENSE_TMA") && atoi(getenv("EBK_DENSE_TMA")) == 0);
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  return on && d.math == EBK_MATH_TF32 && al(x) && al(W) && gemm_tma_eligible(ws.xr, d.K, ws.wr, d.U, d.N, d.U, d.K);
}
int check_dense(const ebk_dense_desc* d) {
  EBK_CHECK_ARG(d != nullptr, "dense: null descriptor");
  EBK_CHECK_ARG(d->N >= 0 && d->K >= 4 && d->U >= 4 && d->K % 4 == 0 && d->U % 4 == 0,
                "dense: need K, U positive multiples of 4 (N=%d K=%d U=%d)", d->N, d->K, d->U);
  EBK_CHECK_ARG(d->dropout >= 0.f && d->dropout < 1.f, "dense: dropout=%f", d->dropout);
  return EBK_OK;
}

}  // namespace

// launchers shared with naml.cu
int bias_act(float* z, const float* b, long n, int U, int relu, cudaStream_t st) {
  if (n == 0) return EBK_OK;
  bias_act_kernel<<<(unsigned)((n / 4 + 255) / 256), 256, 0, st>>>(reinterpret_cast<float4*>(z),
                                                                   reinterpret_cast<const float4*>(b), n / 4, U / 4, relu);
  EBK_LAUNCH_CHECK();
  return EBK_OK;
}
int act_bwd(const float* dy, const float* y, long n, Dropout drop, int relu, int round_out, float* dz, cudaStream_t st) {
  if (n == 0) return EBK_OK;
  EBK_CHECK_ARG(n % 4 == 0 && ((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(dz)) & 15) == 0,
                "act_bwd: n=%ld must be a multiple of 4 and the buffers 16-byte aligned", n);
  act_bwd_kernel<<<(unsigned)((n / 4 + 255) / 256), 256, 0, st>>>(reinterpret_cast<const float4*>(dy),
                                                                  reinterpret_cast<const float4*>(y), n / 4, drop, relu,
                                                                  round_out, reinterpret_cast<float4*>(dz));
  EBK_LAUNCH_CHECK();
  return EBK_OK;
}
}  // namespace ebk

using namespace ebk;

extern "C" size_t ebk_dense_workspace_bytes(const ebk_dense_desc* d) {
  if (check_dense(d) != EBK_OK) return 0;
  return dense_layout(*d, nullptr).bytes;
}

namespace {
// dropout of a Dense+BN layer: host seed, or (CUDA-graph replay) seed1 / seed2 of a device-resident ebk_step_params + add
int dense_dropout(const ebk_dense_desc* d, int training, const ebk_step_params* step_dev, int seed_sel, uint64_t seed_add,
                  Dropout* out) {
  EBK_TRY(check_dense(d));
  EBK_CHECK_ARG(step_dev != nullptr && (seed_sel == 0 || seed_sel == 1), "dense: step_dev must be set and seed_sel 0 or 1");
  *out = make_dropout(training != 0, d->dropout, 0, seed_sel ? &step_dev->seed2 : &step_dev->seed1);
  out->seed_add = seed_add;
  return EBK_OK;
}
int dense_fwd_impl(const ebk_dense_desc* d, const float* x, const float* W, const float* b, const float* gamma,
                   const float* beta, float* mov_mean, float* mov_var, int training, Dropout drop, void* workspace,
                   size_t workspace_bytes, float* y, void* stream);
int dense_bwd_impl(const ebk_dense_desc* d, const float* x, const float* W, const float* gamma, const float* y, int training,
                   Dropout drop, void* workspace, size_t workspace_bytes, const float* dy, float l2_grad_scale, float* dW,
                   float* db, float* dgamma, float* dbeta, float* dx, void* stream);
}  // namespace

extern "C" int ebk_dense_fwd(const ebk_dense_desc* d, const float* x, const float* W, const float* b, const float* gamma,
                             const float* beta, float* mov_mean, float* mov_var, int training, uint64_t seed,
                             void* workspace, size_t workspace_bytes, float* y, void* stream) {
  EBK_TRY(check_dense(d));
  return dense_fwd_impl(d, x, W, b, gamma, beta, mov_mean, mov_var, training, make_dropout(training != 0, d->dropout, seed),
                        workspace, workspace_bytes, y, stream);
}
extern "C" int ebk_dense_fwd_p(const ebk_dense_desc* d, const float* x, const float* W, const float* b, const float* gamma,
                               const float* beta, float* mov_mean, float* mov_var, int training,
                               const ebk_step_params* step_dev, int seed_sel, uint64_t seed_add, void* workspace,
                               size_t workspace_bytes, float* y, void* stream) {
  Dropout drop;
  EBK_TRY(dense_dropout(d, training, step_dev, seed_sel, seed_add, &drop));
  return dense_fwd_impl(d, x, W, b, gamma, beta, mov_mean, mov_var, training, drop, workspace, workspace_bytes, y, stream);
}
extern "C" int ebk_dense_bwd(const ebk_dense_desc* d, const float* x, const float* W, const float* gamma, const float* y,
                             int training, uint64_t seed, void* workspace, size_t workspace_bytes, const float* dy,
                             float l2_grad_scale, float* dW, float* db, float* dgamma, float* dbeta, float* dx,
                             void* stream) {
  EBK_TRY(check_dense(d));
  return dense_bwd_impl(d, x, W, gamma, y, training, make_dropout(training != 0, d->dropout, seed), workspace,
                        workspace_bytes, dy, l2_grad_scale, dW, db, dgamma, dbeta, dx, stream);
}
extern "C" int ebk_dense_bwd_p(const ebk_dense_desc* d, const float* x, const float* W, const float* gamma, const float* y,
                               int training, const ebk_step_params* step_dev, int seed_sel, uint64_t seed_add,
                               void* workspace, size_t workspace_bytes, const float* dy, float l2_grad_scale, float* dW,
                               float* db, float* dgamma, float* dbeta, float* dx, void* stream) {
  Dropout drop;
  EBK_TRY(dense_dropout(d, training, step_dev, seed_sel, seed_add, &drop));
  return dense_bwd_impl(d, x, W, gamma, y, training, drop, workspace, workspace_bytes, dy, l2_grad_scale, dW, db, dgamma,
                        dbeta, dx, stream);
}

namespace {
int dense_fwd_impl(const ebk_dense_desc* d, const float* x, const float* W, const float* b, const float* gamma,
                   const float* beta, float* mov_mean, float* mov_var, int training, Dropout drop, void* workspace,
                   size_t workspace_bytes, float* y, void* stream) {
  EBK_TRY(check_dense(d));
  if (d->N == 0) return EBK_OK;
  EBK_CHECK_ARG(x && W && b && y && workspace, "dense_fwd: null pointer");
  EBK_CHECK_ARG(!d->bn || (gamma && beta && mov_mean && mov_var), "dense_fwd: BatchNorm needs gamma/beta/moving stats");
  DenseWs ws = dense_layout(*d, workspace);
  if (workspace_bytes < ws.bytes) {
    set_error("dense_fwd: workspace %zu < %zu bytes", workspace_bytes, ws.bytes);
    return EBK_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int N = d->N, K = d->K, U = d->U;
  const long n = (long)N * U;
  const Dropout none = make_dropout(false, 0.f, 0);
  const bool tc = d->math != EBK_MATH_FP32;
  const bool x3 = d->math == EBK_MATH_TF32X3;
  GemmOperandA ax{x, K, false, nullptr, 0, none, 0};
  float* a = d->bn ? ws.a : y;  // without BN the activation IS the output
  if (dense_uses_tma(*d, ws, x, W)) {
    EBK_TRY(round_tf32_copy(ws.xr, x, (size_t)N * K, st));
    EBK_TRY(round_tf32_copy(ws.wr, W, (size_t)K * U, st));
    EBK_TRY(gemm_tma(ws.xr, K, false, ws.wr, U, false, a, U, N, U, K, 0.0f, 1.0f, st, -1));
  } else {
    const bool pk = tc && !x3 && gemm_tf32_eligible(ax, W, U, N, U, K);
    if (pk) {
      EBK_TRY(gemm_tf32_pack_b(ws.w_f, nullptr, W, U, false, U, K, st));
      EBK_TRY(gemm_tf32_pack_b(ws.w_d, nullptr, W, U, true, K, U, st));
    }
    EBK_TRY(gemm_dispatch(d->math, ax, pk ? ws.w_f : W, U, false, a, U, N, U, K, 0.0f, st, pk ? GEMM_B_PACKED : GEMM_B_RAW));
  }
  bias_act_kernel<<<(unsigned)((n / 4 + 255) / 256), 256, 0, st>>>(reinterpret_cast<float4*>(a),
                                                                   reinterpret_cast<const float4*>(b), n / 4, U / 4, d->relu);
  EBK_LAUNCH_CHECK();
  if (!d->bn) return EBK_OK;
  if (training) {
    const int nblk = ceil_div(N, ROWS_PER_BLK);
    col_stats_partial_kernel<<<dim3(ceil_div(U / 4, RB_X), nblk), dim3(RB_X, RB_Y), 0, st>>>(a, N, U, ws.partial);
    EBK_LAUNCH_CHECK();
    col_stats_final_kernel<<<ceil_div(U, RB_X), dim3(RB_X, RB_Y), 0, st>>>(ws.partial, nblk, N, U, d->bn_eps, d->bn_momentum,
                                                                           ws.mean, ws.invstd, mov_mean, mov_var);
    EBK_LAUNCH_CHECK();
  } else {
    inference_stats_kernel<<<ceil_div(U, 128), 128, 0, st>>>(U, d->bn_eps, mov_mean, mov_var, ws.mean, ws.invstd);
    EBK_LAUNCH_CHECK();
  }
  bn_apply_kernel<<<(unsigned)((n / 4 + 255) / 256), 256, 0, st>>>(
      reinterpret_cast<const float4*>(a), n / 4, U / 4, reinterpret_cast<const float4*>(ws.mean),
      reinterpret_cast<const float4*>(ws.invstd), reinterpret_cast<const float4*>(gamma),
      reinterpret_cast<const float4*>(beta), drop, reinterpret_cast<float4*>(y));
  EBK_LAUNCH_CHECK();
  return EBK_OK;
}

int dense_bwd_impl(const ebk_dense_desc* d, const float* x, const float* W, const float* gamma, const float* y, int training,
                   Dropout drop, void* workspace, size_t workspace_bytes, const float* dy, float l2_grad_scale, float* dW,
                   float* db, float* dgamma, float* dbeta, float* dx, void* stream) {
  EBK_TRY(check_dense(d));
  if (d->N == 0) return EBK_OK;
  EBK_CHECK_ARG(x && W && dy && dW && db && workspace, "dense_bwd: null pointer");
  EBK_CHECK_ARG(!d->bn || (gamma && dgamma && dbeta), "dense_bwd: BatchNorm needs gamma/dgamma/dbeta");
  EBK_CHECK_ARG(((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(workspace)) & 15) == 0,
                "dense_bwd: dy, gamma and the workspace must be 16-byte aligned");
  EBK_CHECK_ARG(d->bn || y, "dense_bwd: the layer output y is needed when there is no BatchNorm");
  EBK_CHECK_ARG(!d->bn || training, "dense_bwd: BatchNorm backward is defined for training mode");
  DenseWs ws = dense_layout(*d, workspace);
  if (workspace_bytes < ws.bytes) {
    set_error("dense_bwd: workspace %zu < %zu bytes", workspace_bytes, ws.bytes);
    return EBK_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int N = d->N, K = d->K, U = d->U;
  const long n = (long)N * U;
  const Dropout none = make_dropout(false, 0.f, 0);
  const bool tc = d->math != EBK_MATH_FP32;
  const bool x3 = d->math == EBK_MATH_TF32X3;
  const bool rnd = tc && !x3;
  if (d->bn) {
    const int nblk = ceil_div(N, ROWS_PER_BLK);
    bn_bwd_partial_kernel<<<dim3(ceil_div(U / 4, RB_X), nblk), dim3(RB_X, RB_Y), 0, st>>>(dy, ws.a, N, U, ws.mean, ws.invstd,
                                                                                          drop, ws.partial);
    EBK_LAUNCH_CHECK();
    bn_bwd_final_kernel<<<ceil_div(U, RB_X), dim3(RB_X, RB_Y), 0, st>>>(ws.partial, nblk, U, ws.sums, dgamma, dbeta);
    EBK_LAUNCH_CHECK();
    auto f4 = [](const float* p) { return reinterpret_cast<const float4*>(p); };
    bn_bwd_apply_kernel<<<(unsigned)((n / 4 + 255) / 256), 256, 0, st>>>(f4(dy), f4(ws.a), n / 4, N, U / 4, f4(ws.mean),
                                                                         f4(ws.invstd), f4(gamma), f4(ws.sums), drop, d->relu,
                                                                         rnd ? 1 : 0, reinterpret_cast<float4*>(ws.dz));
    EBK_LAUNCH_CHECK();
  } else {
    EBK_TRY(act_bwd(dy, y, n, drop, d->relu, rnd ? 1 : 0, ws.dz, st));
  }
  EBK_TRY(colsum_accum_ws(N, U, ws.dz, U, nullptr, db, ws.colsum, st));
  // dW += x^T dz  (+ 2 l2 W)
  const bool tma = dense_uses_tma(*d, ws, x, W);   // same predicate as the forward call: ws.xr / ws.wr hold its copies
  GemmOperandA axT{x, K, true, nullptr, 0, none, 0};
  if (tma) EBK_TRY(gemm_tma(ws.xr, K, true, ws.dz, U, false, dW, U, K, U, N, 1.0f, 1.0f, st, -1));
  else EBK_TRY(gemm_dispatch(d->math, axT, ws.dz, U, false, dW, U, K, U, N, 1.0f, st, rnd ? GEMM_B_ROUNDED : GEMM_B_RAW));
  if (d->l2 > 0.f && l2_grad_scale != 0.f) {
    const long nw = (long)K * U;
    axpy_kernel<<<(unsigned)((nw + 255) / 256), 256, 0, st>>>(dW, W, nw, 2.0f * d->l2 * l2_grad_scale);
    EBK_LAUNCH_CHECK();
  }
  if (dx && tma) {
    EBK_TRY(gemm_tma(ws.dz, U, false, ws.wr, U, true, dx, K, N, K, U, 0.0f, 1.0f, st, -1));
  } else if (dx) {
    GemmOperandA adz{ws.dz, U, false, nullptr, 0, none, 0};
    GemmOperandA ax{x, K, false, nullptr, 0, none, 0};
    const bool pk = rnd && gemm_tf32_eligible(ax, W, U, N, U, K) && gemm_tf32_eligible(adz, W, U, N, K, U);
    EBK_TRY(gemm_dispatch(d->math, adz, pk ? ws.w_d : W, U, true, dx, K, N, K, U, 0.0f, st, pk ? GEMM_B_PACKED : GEMM_B_RAW));
  }
  return EBK_OK;
}
}  // namespace

extern "C" int ebk_sumsq_accum(const float* x, size_t n, float scale, float* out, void* stream) {
  if (n == 0) return EBK_OK;
  EBK_CHECK_ARG(x && out, "sumsq: null pointer");
  sumsq_kernel<<<148 * 4, 256, 0, (cudaStream_t)stream>>>(x, (long)n, scale, out);
  EBK_LAUNCH_CHECK();
  return EBK_OK;
}

// Embedding-table half of the optimizer step, fused with the Embedding's IndexedSlices gradient
// (reference: keras.layers.Embedding backward + tf.keras.optimizers.Adam, nrms.py:125-134, 76-77).
//
// The dense path costs three streaming passes over table-sized buffers per step: the scatter of the
// R gathered-row gradients into a [V, E] gradient buffer (read-modify-write), the Adam read of that
// buffer, and clearing it.  Keras' Adam is NON-lazy -- m, v and theta of every row move every step -- but
// the gradient itself is row-sparse, so here it never becomes a dense buffer:
//   1. token CSR: count[v], disjoint segments offset[v] (block-local scan + one atomic per block), perm
//      (row numbers r grouped by token);
//   2. one warp per table row v: g = sum over its segment of dX[r, :] * dropout'(r, :) (rows visited in
//      ascending r, so the sum is bit-reproducible), then the Keras-form Adam update of theta/m/v[v, :].
// Rows referenced more than CAP times (padding / very frequent tokens in real data) are pre-reduced with
// atomics into the dense gradient buffer by a separate pass and read from there, so one warp never walks
// a long list.  HBM bytes per step: 24 B/parameter + R*E*4 instead of 32 B/parameter + 3*R*E*4.
#include <stdlib.h>

#include "ebk_common.cuh"

namespace ebk {
namespace {

constexpr int CAP = 32;          // longest segment summed by the row's own warp
constexpr int SCAN_T = 256;      // threads per scan block
constexpr int SCAN_PER_T = 4;    // elements per thread

__global__ void tok_count_kernel(int R, int V, const int32_t* __restrict__ tok, int* __restrict__ count) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  const int t = tok[r];
  if (t >= 0 && t < V) atomicAdd(count + t, 1);
}

// offset[v] = start of a private segment of count[v] slots: block-local exclusive scan + one atomic per block
// (segments of different blocks are disjoint but not ordered by v -- nothing depends on their order)
__global__ void __launch_bounds__(SCAN_T) tok_offset_kernel(int V, const int* __restrict__ count, int* __restrict__ offset,
                                                            int* __restrict__ total) {
  __shared__ int warp_sums[SCAN_T / 32];
  __shared__ int block_base;
  const int v0 = (blockIdx.x * SCAN_T + threadIdx.x) * SCAN_PER_T;
  int c[SCAN_PER_T], s = 0;
#pragma unroll
  for (int i = 0; i < SCAN_PER_T; ++i) {
    c[i] = (v0 + i < V) ? count[v0 + i] : 0;
    s += c[i];
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int n = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += n;
  }
  if (lane == 31) warp_sums[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = lane < SCAN_T / 32 ? warp_sums[lane] : 0, wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int n = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += n;
    }
    if (lane < SCAN_T / 32) warp_sums[lane] = wi - w;   // exclusive prefix of the warp totals
    if (lane == 31) block_base = atomicAdd(total, wi);  // wi of lane 31 = block total
  }
  __syncthreads();
  int run = block_base + warp_sums[warp] + incl - s;
#pragma unroll
  for (int i = 0; i < SCAN_PER_T; ++i) {
    if (v0 + i < V) offset[v0 + i] = run;
    run += c[i];
  }
}

__global__ void tok_fill_kernel(int R, int V, const int32_t* __restrict__ tok, const int* __restrict__ offset,
                                int* __restrict__ cursor, int* __restrict__ perm) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  const int t = tok[r];
  if (t < 0 || t >= V) return;
  perm[offset[t] + atomicAdd(cursor + t, 1)] = r;
}

// rows of tokens with more than CAP occurrences: d_table[tok[r], :] += dX[r, :] * dropout'(r, :)   (one warp per r)
__global__ void heavy_scatter_kernel(int R, int E4, int V, const int32_t* __restrict__ tok, const int* __restrict__ count,
                                     const float4* __restrict__ dX, Dropout drop, float* __restrict__ d_table) {
  const long gw = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (gw >= R) return;
  const int r = (int)gw, t = tok[r];
  if (t < 0 || t >= V || count[t] <= CAP) return;
  for (int c4 = lane; c4 < E4; c4 += 32) {
    float4 g = dX[(long)r * E4 + c4];
    if (drop.on()) {
      const float4 f = drop.factor4_group((uint64_t)r * (uint64_t)E4 + (uint64_t)c4);
      g.x *= f.x; g.y *= f.y; g.z *= f.z; g.w *= f.w;
    }
    float* dst = d_table + ((long)t * E4 + c4) * 4;
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(g.x), "f"(g.y), "f"(g.z), "f"(g.w)
                 : "memory");
  }
}

// One warp per table row: gather-sum the row's gradient, Keras-form Adam update.  The row is processed in
// slices of CH x 32 float4 so that few registers are live and many warps (loads in flight) fit on an SM.
template <int CH>
__global__ void __launch_bounds__(256) embed_adam_kernel(int V, int E4, const int* __restrict__ count,
                                                          const int* __restrict__ offset, const int* __restrict__ perm,
                                                          const float4* __restrict__ dX, Dropout drop,
                                                          float4* __restrict__ d_table, float4* __restrict__ theta,
                                                          float4* __restrict__ m, float4* __restrict__ v, float alpha,
                                                          const float* __restrict__ alpha_dev, float omb1, float omb2,
                                                          float eps) {
  if (alpha_dev != nullptr) alpha = __ldg(alpha_dev);   // CUDA-graph replay: this step's alpha lives in device memory
  const int lane = threadIdx.x & 31;
  const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
  for (long row = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5; row < V; row += nwarps) {
    const int n = count[row];
    const long rbase = row * E4;
    // this lane's entry of the segment and its rank by row number (ascending r => reproducible sum order)
    int mine = 0x7fffffff, rank = 0;
    if (n > 0 && n <= CAP) {
      if (lane < n) mine = perm[offset[row] + lane];
      for (int j = 0; j < n; ++j) rank += (__shfl_sync(0xffffffffu, mine, j) < mine) ? 1 : 0;
    }
    for (int cb = 0; cb * 32 < E4; cb += CH) {
      float4 g[CH], th[CH], mm[CH], vv[CH];
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        const int c4 = lane + 32 * (cb + c);
        g[c] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c4 < E4) {   // state loads first: they do not depend on the gradient
          th[c] = theta[rbase + c4];
          mm[c] = m[rbase + c4];
          vv[c] = v[rbase + c4];
        }
      }
      if (n > CAP) {
#pragma unroll
        for (int c = 0; c < CH; ++c) {
          const int c4 = lane + 32 * (cb + c);
          if (c4 < E4) {
            g[c] = d_table[rbase + c4];
            d_table[rbase + c4] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
      } else {
        for (int k = 0; k < n; ++k) {
          const unsigned who = __ballot_sync(0xffffffffu, lane < n && rank == k);
          const int r = __shfl_sync(0xffffffffu, mine, __ffs(who) - 1);
#pragma unroll
          for (int c = 0; c < CH; ++c) {
            const int c4 = lane + 32 * (cb + c);
            if (c4 < E4) {
              float4 x;
              const float4* src = dX + (long)r * E4 + c4;
              asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                           : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w)
                           : "l"(src));
              if (drop.on()) {
                const float4 f = drop.factor4_group((uint64_t)r * (uint64_t)E4 + (uint64_t)c4);
                x.x *= f.x; x.y *= f.y; x.z *= f.z; x.w *= f.w;
              }
              g[c].x += x.x; g[c].y += x.y; g[c].z += x.z; g[c].w += x.w;
            }
          }
        }
      }
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        const int c4 = lane + 32 * (cb + c);
        if (c4 < E4) {
#define UPD(f)                                          \
  mm[c].f += (g[c].f - mm[c].f) * omb1;                 \
  vv[c].f += (g[c].f * g[c].f - vv[c].f) * omb2;        \
  th[c].f -= (mm[c].f * alpha) / (sqrtf(vv[c].f) + eps);
          UPD(x) UPD(y) UPD(z) UPD(w)
#undef UPD
          theta[rbase + c4] = th[c];
          m[rbase + c4] = mm[c];
          v[rbase + c4] = vv[c];
        }
      }
    }
  }
}

}  // namespace
}  // namespace ebk

using namespace ebk;

namespace ebk {
// count[V] | cursor[V] | total[64] (zeroed each build) | offset[V] | perm[R]
size_t token_csr_bytes(int R, int V) { return ((size_t)3 * V + 64 + (size_t)R) * sizeof(int) + 256; }

int token_csr_build(int R, int V, const int32_t* tok, void* ws, TokenCsr* out, cudaStream_t st) {
  TokenCsr c;
  c.count = reinterpret_cast<int*>(ws);
  c.cursor = c.count + V;
  c.total = c.cursor + V;
  c.offset = c.total + 64;
  c.perm = c.offset + V;
  EBK_CUDA(cudaMemsetAsync(c.count, 0, ((size_t)2 * V + 64) * sizeof(int), st));
  if (R > 0) {
    tok_count_kernel<<<(R + 255) / 256, 256, 0, st>>>(R, V, tok, c.count);
    EBK_LAUNCH_CHECK();
  }
  tok_offset_kernel<<<ceil_div(V, SCAN_T * SCAN_PER_T), SCAN_T, 0, st>>>(V, c.count, c.offset, c.total);
  EBK_LAUNCH_CHECK();
  if (R > 0) {
    tok_fill_kernel<<<(R + 255) / 256, 256, 0, st>>>(R, V, tok, c.offset, c.cursor, c.perm);
    EBK_LAUNCH_CHECK();
  }
  *out = c;
  return EBK_OK;
}
}  // namespace ebk

extern "C" size_t ebk_token_csr_bytes(int32_t R, int32_t V) {
  if (R < 0 || V < 0) return 0;
  return token_csr_bytes(R, V);
}
extern "C" size_t ebk_embed_adam_workspace_bytes(int32_t R, int32_t V) {
  if (R < 0 || V < 0) return 0;
  return token_csr_bytes(R, V);
}

extern "C" int ebk_embed_adam_step(int32_t R, int32_t E, int32_t V, const int32_t* tok, const float* dX, float drop_p,
                                   uint64_t drop_seed, float* theta, float* d_table, float* m, float* v, float alpha,
                                   double beta1, double beta2, float eps, void* workspace, size_t workspace_bytes,
                                   void* stream) {
  return ebk_embed_adam_step_p(R, E, V, tok, dX, drop_p, drop_seed, theta, d_table, m, v, alpha, nullptr, beta1, beta2, eps,
                               workspace, workspace_bytes, stream);
}

extern "C" int ebk_embed_adam_step_p(int32_t R, int32_t E, int32_t V, const int32_t* tok, const float* dX, float drop_p,
                                     uint64_t drop_seed, float* theta, float* d_table, float* m, float* v, float alpha,
                                     const ebk_step_params* step_dev, double beta1, double beta2, float eps,
                                     void* workspace, size_t workspace_bytes, void* stream) {
  EBK_CHECK_ARG(R >= 0 && E >= 4 && E % 4 == 0 && E <= 1024 && V >= 1, "embed_adam: need E %% 4 == 0, E <= 1024 (E=%d)", E);
  EBK_CHECK_ARG((R == 0 || (tok && dX)) && theta && d_table && m && v && workspace, "embed_adam: null pointer");
  EBK_CHECK_ARG(((uintptr_t)theta % 16 == 0) && ((uintptr_t)d_table % 16 == 0) && ((uintptr_t)m % 16 == 0) &&
                    ((uintptr_t)v % 16 == 0) && (R == 0 || (uintptr_t)dX % 16 == 0),
                "embed_adam: buffers must be 16-byte aligned");
  if (workspace_bytes < ebk_embed_adam_workspace_bytes(R, V)) {
    set_error("embed_adam: workspace %zu < %zu bytes", workspace_bytes, ebk_embed_adam_workspace_bytes(R, V));
    return EBK_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const Dropout drop = make_dropout(drop_p > 0.0f, drop_p, drop_seed, step_dev ? &step_dev->seed1 : nullptr);
  const float* alpha_dev = step_dev ? &step_dev->alpha : nullptr;
  const int E4 = E / 4;
  prof_set_group(0);
  if (prof_on()) prof_begin(T_SCATTER, st);
  TokenCsr csr;
  EBK_TRY(token_csr_build(R, V, tok, workspace, &csr, st));
  int *count = csr.count, *offset = csr.offset, *perm = csr.perm;
  if (R > 0) {
    heavy_scatter_kernel<<<(unsigned)(((long)R * 32 + 255) / 256), 256, 0, st>>>(
        R, E4, V, tok, count, reinterpret_cast<const float4*>(dX), drop, d_table);
    EBK_LAUNCH_CHECK();
  }
  if (prof_on()) prof_end(T_SCATTER, st);
  const float omb1 = (float)(1.0 - beta1), omb2 = (float)(1.0 - beta2);
  // 128-thread blocks: measured 0.90 -> 0.86 ms on its own, and more of its warps fit next to a resident CTA of the QKV
  // weight-gradient GEMM that runs beside it on the side stream (step -0.05 ms; EBK_ADAM_BLOCK=256 / 64 for A/B runs)
  static const int env_blk = getenv("EBK_ADAM_BLOCK") ? atoi(getenv("EBK_ADAM_BLOCK")) : 0;
  const int blk = (env_blk == 256 || env_blk == 64) ? env_blk : 128;
  const long warps = V;
  const long blocks = (warps * 32 + blk - 1) / blk;
  static const int env_ch = getenv("EBK_ADAM_CH") ? atoi(getenv("EBK_ADAM_CH")) : 0;      // tuning knobs
  static const int env_cap = getenv("EBK_ADAM_CAP") ? atoi(getenv("EBK_ADAM_CAP")) : 0;
  const long cap = 148L * (env_cap > 0 ? env_cap : 8) * 8 * (256 / blk);
  const unsigned grid = (unsigned)(blocks < cap ? blocks : cap);
  if (prof_on()) prof_begin(T_ADAM, st);
  const int nc = ceil_div(E4, 32);
  // measured on B200 (tools/sweep_adam.sh, E = 768): slices of 2 x 32 float4 -> 0.85 ms, 3 -> 0.97, 6 -> 1.13:
  // fewer live registers = more resident warps = more loads in flight
  (void)nc;
  int ch = env_ch > 0 ? env_ch : 2;
#define RUN(CH_)                                                                                                  \
  embed_adam_kernel<CH_><<<grid, blk, 0, st>>>(V, E4, count, offset, perm, reinterpret_cast<const float4*>(dX), drop, \
                                               reinterpret_cast<float4*>(d_table), reinterpret_cast<float4*>(theta),  \
                                               reinterpret_cast<float4*>(m), reinterpret_cast<float4*>(v), alpha, alpha_dev, \
                                               omb1, omb2, eps)
  if (ch == 1) RUN(1); else if (ch == 2) RUN(2); else if (ch == 3) RUN(3); else RUN(6);
#undef RUN
  if (prof_on()) prof_end(T_ADAM, st);
  EBK_LAUNCH_CHECK();
  return EBK_OK;
}

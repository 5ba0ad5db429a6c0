// AttLayer2 (reference layers.py:55-81) around the GEMM X.W:
//   h = tanh(X W + b); a = h q; e = exp(a) (NO max subtraction); w = e/(sum e + 1e-7); y = sum_t w_t X_t
// X = dropout2(Y0) is recomputed from the saved attention output and the counter-based mask.
// One CTA per sequence; also the column-sum and embedding-row scatter helpers.
#include <stdlib.h>

#include "ebk_common.cuh"

namespace ebk {
namespace {

constexpr int PT = 256;  // threads per CTA (measured: 256 beats 128 -- 4 rows per warp instead of 8 in flight)
constexpr float K_EPS = 1e-7f;  // keras.backend.epsilon()

// hbuf [R, att]: in = X W (pre-activation without bias), out = tanh(. + b)
__global__ void __launch_bounds__(PT) attpool_fwd_kernel(int L, int D, int att, const float* __restrict__ y0,
                                                          Dropout drop, float* __restrict__ hbuf,
                                                          const float* __restrict__ attb,
                                                          const float* __restrict__ attq,
                                                          float* __restrict__ w, float* __restrict__ out,
                                                          int out_ld) {
  __shared__ float a_s[64];
  __shared__ float w_s[64];
  const int n = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = PT / 32;
  const bool vec = (att & 3) == 0 && ((reinterpret_cast<uintptr_t>(hbuf) | reinterpret_cast<uintptr_t>(attb) |
                                      reinterpret_cast<uintptr_t>(attq)) & 15) == 0;
  for (int t = warp; t < L; t += nwarp) {
    float* hrow = hbuf + ((long)n * L + t) * att;
    float acc = 0.0f;
    if (vec) {
      for (int j = lane; j < (att >> 2); j += 32) {
        float4 h = reinterpret_cast<float4*>(hrow)[j];
        const float4 bb = __ldg(reinterpret_cast<const float4*>(attb) + j), qq = __ldg(reinterpret_cast<const float4*>(attq) + j);
        h.x = tanhf(h.x + bb.x); h.y = tanhf(h.y + bb.y); h.z = tanhf(h.z + bb.z); h.w = tanhf(h.w + bb.w);
        reinterpret_cast<float4*>(hrow)[j] = h;
        acc = fmaf(h.x, qq.x, fmaf(h.y, qq.y, fmaf(h.z, qq.z, fmaf(h.w, qq.w, acc))));
      }
    } else {
      for (int j = lane; j < att; j += 32) {
        float h = tanhf(hrow[j] + attb[j]);
        hrow[j] = h;
        acc = fmaf(h, attq[j], acc);
      }
    }
    acc = warp_sum(acc);
    if (lane == 0) a_s[t] = acc;
  }
  __syncthreads();
  if (warp == 0) {
    float s = 0.0f;
    for (int t = lane; t < L; t += 32) {
      float e = expf(a_s[t]);  // layers.py:70-71: plain exp
      w_s[t] = e;
      s += e;
    }
    s = warp_sum(s);
    float r = 1.0f / (s + K_EPS);  // layers.py:75-77
    for (int t = lane; t < L; t += 32) {
      float ww = w_s[t] * r;
      w_s[t] = ww;
      w[(long)n * L + t] = ww;
    }
  }
  __syncthreads();
  for (int d = threadIdx.x; d < D; d += PT) {
    float acc = 0.0f;
    for (int t = 0; t < L; ++t) {
      long r = (long)n * L + t;
      float x = y0[r * D + d];
      if (drop.on()) x *= drop.factor((uint64_t)r * (uint64_t)D + (uint64_t)d);
      acc = fmaf(w_s[t], x, acc);
    }
    out[(long)n * out_ld + d] = acc;
  }
}

__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit_wait_all() {
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// TMA-path forward (y0 already holds dropout(Y0)): every HBM read of the sequence is issued up front -- the [L, D]
// rows of y0 go to shared memory with cp.async while the tanh / q-dot pass streams the [L, att] rows of h -- so the
// pooling pass reads shared memory instead of waiting on 30 dependent global loads per thread.
__global__ void __launch_bounds__(PT) attpool_fwd_fast_kernel(int L, int D, int att, const float* __restrict__ y0,
                                                               float* __restrict__ hbuf, const float* __restrict__ attb,
                                                               const float* __restrict__ attq, float* __restrict__ w,
                                                               float* __restrict__ out, int out_ld) {
  extern __shared__ __align__(16) float sm_f[];   // h rows [L, att] | y rows [L, D]
  __shared__ float a_s[64];
  __shared__ float w_s[64];
  const int n = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = PT / 32;
  const int D4 = D >> 2, A4 = att >> 2;
  float* h_s = sm_f;
  float* y_s = sm_f + (size_t)L * att;
  // every HBM read of this sequence is issued NOW: first the h rows (needed first), then the y rows
  const float4* hsrc = reinterpret_cast<const float4*>(hbuf + (long)n * L * att);
  for (int i = threadIdx.x; i < L * A4; i += PT) cp_async16(reinterpret_cast<float4*>(h_s) + i, hsrc + i);
  asm volatile("cp.async.commit_group;" ::: "memory");
  const float4* ysrc = reinterpret_cast<const float4*>(y0 + (long)n * L * D);
  for (int i = threadIdx.x; i < L * D4; i += PT) cp_async16(reinterpret_cast<float4*>(y_s) + i, ysrc + i);
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 1;" ::: "memory");   // h rows have landed (this thread's copies) ...
  __syncthreads();                                        // ... and everybody else's
  // this lane's slices of b and q stay in registers for every token (att <= 4 * 32 * PV floats, else re-read per token)
  constexpr int PV = 2;
  float4 bb_r[PV], qq_r[PV];
#pragma unroll
  for (int u = 0; u < PV; ++u) {
    const int j = lane + 32 * u;
    bb_r[u] = j < A4 ? __ldg(reinterpret_cast<const float4*>(attb) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
    qq_r[u] = j < A4 ? __ldg(reinterpret_cast<const float4*>(attq) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int t = warp; t < L; t += nwarp) {
    float4* hrow_g = reinterpret_cast<float4*>(hbuf + ((long)n * L + t) * att);
    const float4* hrow = reinterpret_cast<const float4*>(h_s + (size_t)t * att);
    float acc = 0.0f;
#pragma unroll
    for (int u = 0; u < PV; ++u) {
      const int j = lane + 32 * u;
      if (j < A4) {
        float4 h = hrow[j];
        const float4 bb = bb_r[u], qq = qq_r[u];
        h.x = tanhf(h.x + bb.x); h.y = tanhf(h.y + bb.y); h.z = tanhf(h.z + bb.z); h.w = tanhf(h.w + bb.w);
        hrow_g[j] = h;     // tanh values: what the backward pass reads
        acc = fmaf(h.x, qq.x, fmaf(h.y, qq.y, fmaf(h.z, qq.z, fmaf(h.w, qq.w, acc))));
      }
    }
    for (int j = lane + 32 * PV; j < A4; j += 32) {
      float4 h = hrow[j];
      const float4 bb = __ldg(reinterpret_cast<const float4*>(attb) + j), qq = __ldg(reinterpret_cast<const float4*>(attq) + j);
      h.x = tanhf(h.x + bb.x); h.y = tanhf(h.y + bb.y); h.z = tanhf(h.z + bb.z); h.w = tanhf(h.w + bb.w);
      hrow_g[j] = h;
      acc = fmaf(h.x, qq.x, fmaf(h.y, qq.y, fmaf(h.z, qq.z, fmaf(h.w, qq.w, acc))));
    }
    acc = warp_sum(acc);
    if (lane == 0) a_s[t] = acc;
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    float s = 0.0f;
    for (int t = lane; t < L; t += 32) {
      float e = expf(a_s[t]);  // layers.py:70-71: plain exp
      w_s[t] = e;
      s += e;
    }
    s = warp_sum(s);
    float r = 1.0f / (s + K_EPS);  // layers.py:75-77
    for (int t = lane; t < L; t += 32) {
      float ww = w_s[t] * r;
      w_s[t] = ww;
      w[(long)n * L + t] = ww;
    }
  }
  __syncthreads();
  for (int d4 = threadIdx.x; d4 < D4; d4 += PT) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int t = 0; t < L; ++t) {
      const float4 x = reinterpret_cast<const float4*>(y_s)[t * D4 + d4];
      const float ww = w_s[t];
      acc.x = fmaf(ww, x.x, acc.x); acc.y = fmaf(ww, x.y, acc.y); acc.z = fmaf(ww, x.z, acc.z); acc.w = fmaf(ww, x.w, acc.w);
    }
    *reinterpret_cast<float4*>(out + (long)n * out_ld + d4 * 4) = acc;
  }
}

// dy[r,:] = w_r * d_out[n,:]   (the dpre.W^T term is added afterwards by a beta=1 GEMM)
// da[r]   = w_r (dw_r - sum_j w_j dw_j),  dw_r = X_r . d_out[n]
// dpre[r,j] = da_r q_j (1 - h_rj^2)
__global__ void __launch_bounds__(PT) attpool_bwd_kernel(int L, int D, int att, const float* __restrict__ y0,
                                                          Dropout drop, const float* __restrict__ hbuf,
                                                          const float* __restrict__ attq,
                                                          const float* __restrict__ w,
                                                          const float* __restrict__ d_out, int dout_ld,
                                                          float* __restrict__ da, float* __restrict__ dpre,
                                                          float* __restrict__ dy, bool round_dpre) {
  __shared__ float dw_s[64];
  __shared__ float da_s[64];
  const int n = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = PT / 32;
  const float* g = d_out + (long)n * dout_ld;
  for (int t = warp; t < L; t += nwarp) {
    long r = (long)n * L + t;
    float wt = w[r];
    float acc = 0.0f;
    for (int d = lane; d < D; d += 32) {
      float x = y0[r * D + d];
      if (drop.on()) x *= drop.factor((uint64_t)r * (uint64_t)D + (uint64_t)d);
      float gd = g[d];
      acc = fmaf(x, gd, acc);
      dy[r * D + d] = wt * gd;
    }
    acc = warp_sum(acc);
    if (lane == 0) dw_s[t] = acc;
  }
  __syncthreads();
  if (warp == 0) {
    float s = 0.0f;
    for (int t = lane; t < L; t += 32) s = fmaf(w[(long)n * L + t], dw_s[t], s);
    s = warp_sum(s);
    for (int t = lane; t < L; t += 32) {
      float v = w[(long)n * L + t] * (dw_s[t] - s);
      da_s[t] = v;
      da[(long)n * L + t] = v;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < L * att; i += PT) {
    int t = i / att, j = i % att;
    long r = (long)n * L + t;
    float h = hbuf[r * att + j];
    const float v = da_s[t] * attq[j] * (1.0f - h * h);
    dpre[r * att + j] = round_dpre ? round_tf32_bits(v) : v;
  }
}

// TMA-path AttLayer2 backward tail: one CTA per sequence.
//   dw_t = X_t . g;  da_t = w_t (dw_t - sum_j w_j dw_j);  dpre[t, j] = tf32(da_t q_j (1 - h_tj^2))
//   colpart[n, j] = sum_t dpre[t, j];  colpart[n, att + j] = sum_t h_tj da_t     (deterministic partials)
__global__ void __launch_bounds__(PT) attpool_bwd_fused_kernel(int L, int D, int att, const float* __restrict__ y0,
                                                                const float* __restrict__ hbuf,
                                                                const float* __restrict__ attq,
                                                                const float* __restrict__ w,
                                                                const float* __restrict__ d_out, int dout_ld,
                                                                float* __restrict__ da, float* __restrict__ dpre,
                                                                float* __restrict__ colpart, int h_in_smem) {
  extern __shared__ __align__(16) float h_s[];   // [L, att] | [L, D] (when launched with shared memory)
  __shared__ float dw_s[64];
  __shared__ float da_s[64];
  const int n = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = PT / 32;
  const float4* g4 = reinterpret_cast<const float4*>(d_out + (long)n * dout_ld);
  const int D4 = D >> 2;
  float* y_s = h_s + (size_t)L * att;
  if (h_in_smem) {   // every HBM read of this sequence is issued now: the y rows (needed first), then the h rows
    const float4* ysrc = reinterpret_cast<const float4*>(y0 + (long)n * L * D);
    for (int i = threadIdx.x; i < L * D4; i += PT) cp_async16(reinterpret_cast<float4*>(y_s) + i, ysrc + i);
    asm volatile("cp.async.commit_group;" ::: "memory");
    const float4* hsrc = reinterpret_cast<const float4*>(hbuf + (long)n * L * att);
    for (int i = threadIdx.x; i < L * (att >> 2); i += PT) cp_async16(reinterpret_cast<float4*>(h_s) + i, hsrc + i);
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    __syncthreads();
  }
  // this lane's slice of d_out[n, :] stays in registers for every token (D <= 4 * 32 * GV floats, else re-read per token)
  constexpr int GV = 4;
  float4 g_r[GV];
#pragma unroll
  for (int u = 0; u < GV; ++u) g_r[u] = lane + 32 * u < D4 ? __ldg(g4 + lane + 32 * u) : make_float4(0.f, 0.f, 0.f, 0.f);
  for (int t = warp; t < L; t += nwarp) {
    const float4* x4 = h_in_smem ? reinterpret_cast<const float4*>(y_s + (size_t)t * D)
                                 : reinterpret_cast<const float4*>(y0 + ((long)n * L + t) * D);
    float acc = 0.0f;
#pragma unroll
    for (int u = 0; u < GV; ++u) {
      const int d = lane + 32 * u;
      if (d < D4) {
        const float4 x = x4[d], gd = g_r[u];
        acc = fmaf(x.x, gd.x, fmaf(x.y, gd.y, fmaf(x.z, gd.z, fmaf(x.w, gd.w, acc))));
      }
    }
    for (int d = lane + 32 * GV; d < D4; d += 32) {
      const float4 x = x4[d], gd = __ldg(g4 + d);
      acc = fmaf(x.x, gd.x, fmaf(x.y, gd.y, fmaf(x.z, gd.z, fmaf(x.w, gd.w, acc))));
    }
    acc = warp_sum(acc);
    if (lane == 0) dw_s[t] = acc;
  }
  __syncthreads();
  if (warp == 0) {
    float s = 0.0f;
    for (int t = lane; t < L; t += 32) s = fmaf(w[(long)n * L + t], dw_s[t], s);
    s = warp_sum(s);
    for (int t = lane; t < L; t += 32) {
      const float v = w[(long)n * L + t] * (dw_s[t] - s);
      da_s[t] = v;
      da[(long)n * L + t] = v;
    }
  }
  if (h_in_smem) asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  for (int j = threadIdx.x; j < att; j += PT) {
    const float qj = attq[j];
    float sb = 0.0f, sq = 0.0f;
    for (int t = 0; t < L; ++t) {
      const long r = (long)n * L + t;
      const float h = h_in_smem ? h_s[t * att + j] : hbuf[r * att + j];
      const float v = round_tf32_bits(da_s[t] * qj * (1.0f - h * h));
      dpre[r * att + j] = v;
      sb += v;
      sq = fmaf(h, da_s[t], sq);
    }
    colpart[(long)n * 2 * att + j] = sb;
    colpart[(long)n * 2 * att + att + j] = sq;
  }
}

// Two-stage deterministic column reduction: partial[blk, j] then out[j] += sum_blk.
constexpr int CS_ROWS = 32;   // rows per partial block (32: enough blocks to hide the load latency even for the 256-row user encoder)
__global__ void colsum_partial_kernel(int R, int Ncols, const float* __restrict__ X, int ldx,
                                      const float* __restrict__ coef, float* __restrict__ partial) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= Ncols) return;
  int r0 = blockIdx.y * CS_ROWS, r1 = min(R, r0 + CS_ROWS);
  float acc = 0.0f;
  for (int r = r0; r < r1; ++r) {
    float x = X[(long)r * ldx + j];
    acc = coef ? fmaf(coef[r], x, acc) : acc + x;
  }
  partial[(long)blockIdx.y * Ncols + j] = acc;
}
// final stage: block = 32 columns x 8 slices of the partial rows, slices combined in a fixed order (deterministic)
__global__ void __launch_bounds__(256) colsum_final_kernel(int nblk, int Ncols, const float* __restrict__ partial,
                                                           float* __restrict__ out) {
  __shared__ float sh[8][32];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + tx;
  float acc = 0.0f;
  if (j < Ncols)
    for (int b = ty; b < nblk; b += 8) acc += partial[(long)b * Ncols + j];
  sh[ty][tx] = acc;
  __syncthreads();
  if (ty != 0 || j >= Ncols) return;
  for (int k = 1; k < 8; ++k) acc += sh[k][tx];
  out[j] += acc;
}

// the same with the columns split between two outputs (db | dq partials of AttLayer2): out0[j] for j < n0, out1[j - n0] after
__global__ void colsum_final2_kernel(int nblk, int Ncols, int n0, const float* __restrict__ partial, float* __restrict__ out0,
                                     float* __restrict__ out1) {
  __shared__ float sh[8][32];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + tx;
  float acc = 0.0f;
  if (j < Ncols)
    for (int b = ty; b < nblk; b += 8) acc += partial[(long)b * Ncols + j];
  sh[ty][tx] = acc;
  __syncthreads();
  if (ty != 0 || j >= Ncols) return;
  for (int k = 1; k < 8; ++k) acc += sh[k][tx];
  if (j < n0) out0[j] += acc;
  else out1[j - n0] += acc;
}

__global__ void round_tf32_copy_kernel(float4* __restrict__ dst, const float4* __restrict__ src, size_t n4) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 v = src[i];
  v.x = round_tf32_bits(v.x); v.y = round_tf32_bits(v.y); v.z = round_tf32_bits(v.z); v.w = round_tf32_bits(v.w);
  dst[i] = v;
}

__global__ void split_tf32_copy_kernel(float* __restrict__ hi, float* __restrict__ lo, const float* __restrict__ src,
                                       size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float x = src[i];
  const float h = round_tf32_bits(x);
  hi[i] = h;
  lo[i] = round_tf32_bits(x - h);
}

// Materialises the A operand of the QKV projection for the TMA GEMM: gather + dropout + tf32 rounding.
// PEERS: the table is sharded over the ranks of one NVSwitch box and each 16-byte chunk is read from its OWNER's
// HBM over NVLink (peer pointers mapped with CUDA IPC) -- the all-gather of the updated table that data
// parallel would otherwise need every step is fused into the gather that has to happen anyway.
template <bool PEERS>
__global__ void embed_rows_kernel(long n4, int E4, int V, const int32_t* __restrict__ tok,
                                  const float4* __restrict__ src, Dropout drop, float4* __restrict__ xd,
                                  PeerTables peers, long group0) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const int r = (int)(i / E4), c4 = (int)(i - (long)r * E4);
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (tok != nullptr) {
    const int t = __ldg(tok + r);
    if (t >= 0 && t < V) {
      const long chunk = (long)t * E4 + c4;
      const float4* p = src + chunk;
      if (PEERS) {
        const int owner = (int)(((unsigned long long)chunk * 4ull) / peers.shard_floats);
        p = reinterpret_cast<const float4*>(peers.p[owner < peers.world ? owner : peers.world - 1]) + chunk;
      }
      asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                   : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                   : "l"(p));
    }
  } else {
    v = __ldg(src + i);
  }
  if (drop.on()) {
    const float4 f = drop.factor4_group((uint64_t)(group0 + i));   // group0: first group of this row chunk
    v.x *= f.x; v.y *= f.y; v.z *= f.z; v.w *= f.w;
  }
  v.x = round_tf32_bits(v.x); v.y = round_tf32_bits(v.y); v.z = round_tf32_bits(v.z); v.w = round_tf32_bits(v.w);
  xd[i] = v;
}

// The token gather proper: one warp per row, each lane keeps U independent 16-byte loads in flight (one token-id load
// per warp instead of one per float4; ncu on the per-float4 kernel: 67 % of the stall samples on the two dependent loads).
template <bool PEERS>
__global__ void __launch_bounds__(256) embed_rows_warp_kernel(int R, int E4, int V, const int32_t* __restrict__ tok,
                                                              const float4* __restrict__ src, Dropout drop,
                                                              float4* __restrict__ xd, PeerTables peers, long group0) {
  constexpr int U = 6;
  const int lane = threadIdx.x & 31;
  const long r = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= R) return;
  const int t = __ldg(tok + r);
  const bool ok = t >= 0 && t < V;   // ids outside the table read a zero row
  const long rowbase = (long)t * E4, obase = r * E4;
  for (int c0 = 0; c0 < E4; c0 += 32 * U) {
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int c4 = c0 + u * 32 + lane;
      v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ok && c4 < E4) {
        const long chunk = rowbase + c4;
        const float4* p = src + chunk;
        if (PEERS) {
          const int owner = (int)(((unsigned long long)chunk * 4ull) / peers.shard_floats);
          p = reinterpret_cast<const float4*>(peers.p[owner < peers.world ? owner : peers.world - 1]) + chunk;
        }
        asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                     : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w)
                     : "l"(p));
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int c4 = c0 + u * 32 + lane;
      if (c4 < E4) {
        float4 x = v[u];
        if (drop.on()) {
          const float4 f = drop.factor4_group((uint64_t)(group0 + obase + c4));
          x.x *= f.x; x.y *= f.y; x.z *= f.z; x.w *= f.w;
        }
        x.x = round_tf32_bits(x.x); x.y = round_tf32_bits(x.y); x.z = round_tf32_bits(x.z); x.w = round_tf32_bits(x.w);
        xd[obase + c4] = x;
      }
    }
  }
}

// ---- gather through a token CSR: every DISTINCT token's row is read once and written to all of its positions (each
// with its own dropout mask).  Under data parallel 7/8 of the rows are remote and the gather is NVLink-read bound, so
// reading 134 k distinct rows instead of 192 k positions (uniform synthetic ids; far fewer distinct ids in real titles)
// is what moves it.  One warp per token id; tokens with more than CSR_CAP positions and ids outside the table are left
// to embed_rows_rest_kernel (one warp per position), so no warp walks a long list.
constexpr int CSR_CAP = 32;
template <bool PEERS>
__global__ void __launch_bounds__(256) embed_rows_csr_kernel(int V, int E4, const int* __restrict__ count,
                                                             const int* __restrict__ offset, const int* __restrict__ perm,
                                                             const float4* __restrict__ src, Dropout drop,
                                                             float4* __restrict__ xd, PeerTables peers, int rr_ways,
                                                             int rr_rows) {
  constexpr int U = 6;
  const int lane = threadIdx.x & 31;
  const long w = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  // warps walk the token ids round-robin over the owners' row ranges, so that local (HBM) and remote (NVLink) rows are
  // in flight together for the whole kernel (in id order the kernel would run an HBM phase, then an NVLink phase)
  const long t = (w % rr_ways) * rr_rows + w / rr_ways;
  if (w >= (long)rr_ways * rr_rows || t >= V) return;
  const int cnt = __ldg(count + t);
  if (cnt == 0 || cnt > CSR_CAP) return;
  const int off = __ldg(offset + t);
  const int my_pos = lane < cnt ? __ldg(perm + off + lane) : 0;   // cnt <= 32: one position per lane, broadcast below
  const long rowbase = t * E4;
  for (int c0 = 0; c0 < E4; c0 += 32 * U) {
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int c4 = c0 + u * 32 + lane;
      v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c4 < E4) {
        const long chunk = rowbase + c4;
        const float4* p = src + chunk;
        if (PEERS) {
          const int owner = (int)(((unsigned long long)chunk * 4ull) / peers.shard_floats);
          p = reinterpret_cast<const float4*>(peers.p[owner < peers.world ? owner : peers.world - 1]) + chunk;
        }
        asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                     : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w)
                     : "l"(p));
      }
    }
    for (int j = 0; j < cnt; ++j) {
      const long obase = (long)__shfl_sync(0xffffffffu, my_pos, j) * E4;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int c4 = c0 + u * 32 + lane;
        if (c4 < E4) {
          float4 x = v[u];
          if (drop.on()) {
            const float4 f = drop.factor4_group((uint64_t)(obase + c4));
            x.x *= f.x; x.y *= f.y; x.z *= f.z; x.w *= f.w;
          }
          x.x = round_tf32_bits(x.x); x.y = round_tf32_bits(x.y); x.z = round_tf32_bits(x.z); x.w = round_tf32_bits(x.w);
          xd[obase + c4] = x;
        }
      }
    }
  }
}
// the positions embed_rows_csr_kernel leaves: ids outside the table (zero row) and tokens with more than CSR_CAP positions
template <bool PEERS>
__global__ void __launch_bounds__(256) embed_rows_rest_kernel(int R, int E4, int V, const int32_t* __restrict__ tok,
                                                              const int* __restrict__ count, const float4* __restrict__ src,
                                                              Dropout drop, float4* __restrict__ xd, PeerTables peers) {
  const int lane = threadIdx.x & 31;
  const long r = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= R) return;
  const int t = __ldg(tok + r);
  const bool ok = t >= 0 && t < V;
  if (ok && __ldg(count + t) <= CSR_CAP) return;
  const long rowbase = (long)t * E4, obase = r * E4;
  for (int c4 = lane; c4 < E4; c4 += 32) {
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ok) {
      const long chunk = rowbase + c4;
      const float4* p = src + chunk;
      if (PEERS) {
        const int owner = (int)(((unsigned long long)chunk * 4ull) / peers.shard_floats);
        p = reinterpret_cast<const float4*>(peers.p[owner < peers.world ? owner : peers.world - 1]) + chunk;
      }
      x = __ldg(p);   // (allocating load: a frequent row is re-read by many warps)
      if (drop.on()) {
        const float4 f = drop.factor4_group((uint64_t)(obase + c4));
        x.x *= f.x; x.y *= f.y; x.z *= f.z; x.w *= f.w;
      }
      x.x = round_tf32_bits(x.x); x.y = round_tf32_bits(x.y); x.z = round_tf32_bits(x.z); x.w = round_tf32_bits(x.w);
    }
    xd[obase + c4] = x;
  }
}

__global__ void scatter_rows_add_kernel(int R, int E4, int V, const int32_t* __restrict__ tok,
                                        const float4* __restrict__ dX, Dropout drop,
                                        float* __restrict__ d_table) {
  long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= (long)R * E4) return;
  int r = (int)(i / E4), c4 = (int)(i % E4);
  int t = tok[r];
  if (t < 0 || t >= V) return;
  float4 g = dX[i];
  if (drop.on()) {
    float4 f = drop.factor4((uint64_t)i * 4ull);
    g.x *= f.x; g.y *= f.y; g.z *= f.z; g.w *= f.w;
  }
  float* dst = d_table + ((long)t * E4 + c4) * 4;
  // 128-bit vector reduction (sm_90+): one L2 atomic transaction for 4 floats
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(g.x), "f"(g.y), "f"(g.z),
               "f"(g.w)
               : "memory");
}

}  // namespace

int attpool_fwd(int n_seq, int L, int D, int att, const float* y0, Dropout drop, float* hbuf, const float* attb,
                const float* attq, float* w, float* out, cudaStream_t st, int out_ld) {
  if (n_seq <= 0) return EBK_OK;
  EBK_CHECK_ARG(L <= 64, "attpool: L=%d > 64", L);
  const int old = out_ld > 0 ? out_ld : D;
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  const size_t ysm = (size_t)L * (D + att) * sizeof(float);   // h rows + y rows of one sequence
  static const bool fast_on = !(getenv("EBK_ATTPOOL_FAST") && atoi(getenv("EBK_ATTPOOL_FAST")) == 0);
  if (fast_on && !drop.on() && D % 4 == 0 && att % 4 == 0 && old % 4 == 0 && ysm <= 96 * 1024 && al(y0) && al(hbuf) &&
      al(attb) && al(attq) && al(out)) {
    if (ysm > 48 * 1024)
      EBK_CUDA(cudaFuncSetAttribute(attpool_fwd_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    attpool_fwd_fast_kernel<<<n_seq, PT, ysm, st>>>(L, D, att, y0, hbuf, attb, attq, w, out, old);
    EBK_LAUNCH_CHECK();
    return EBK_OK;
  }
  attpool_fwd_kernel<<<n_seq, PT, 0, st>>>(L, D, att, y0, drop, hbuf, attb, attq, w, out, old);
  EBK_LAUNCH_CHECK();
  return EBK_OK;
}

int attpool_bwd(int n_seq, int L, int D, int att, const float* y0, Dropout drop, const float* hbuf,
                const float* attq, const float* w, const float* d_out, float* da, float* dpre, float* dy,
                bool round_dpre, cudaStream_t st, int dout_ld) {
  if (n_seq <= 0) return EBK_OK;
  EBK_CHECK_ARG(L <= 64, "attpool: L=%d > 64", L);
  attpool_bwd_kernel<<<n_seq, PT, 0, st>>>(L, D, att, y0, drop, hbuf, attq, w, d_out, dout_ld > 0 ? dout_ld : D, da, dpre, dy,
                                           round_dpre);
  EBK_LAUNCH_CHECK();
  return EBK_OK;
}

int attpool_bwd_fused(int n_seq, int L, int D, int att, const float* y0, const float* hbuf, const float* attq,
                      const float* w, const float* d_out, float* da, float* dpre, float* colpart, cudaStream_t st,
                      int dout_ld) {
  if (n_seq <= 0) return EBK_OK;
  if (dout_ld <= 0) dout_ld = D;
  EBK_CHECK_ARG(L <= 64 && D % 4 == 0 && dout_ld % 4 == 0 && (reinterpret_cast<uintptr_t>(d_out) & 15) == 0,
                "attpool: L=%d > 64, or D=%d / dout_ld=%d not multiples of 4", L, D, dout_ld);
  const size_t hsm = (size_t)L * (att + D) * sizeof(float);   // h rows + y rows of one sequence
  static const bool fast_on = !(getenv("EBK_ATTPOOL_FAST") && atoi(getenv("EBK_ATTPOOL_FAST")) == 0);
  const bool h_smem = fast_on && att % 4 == 0 && hsm <= 96 * 1024 && (reinterpret_cast<uintptr_t>(hbuf) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(y0) & 15) == 0;
  if (h_smem && hsm > 48 * 1024)
    EBK_CUDA(cudaFuncSetAttribute(attpool_bwd_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
  attpool_bwd_fused_kernel<<<n_seq, PT, h_smem ? hsm : 0, st>>>(L, D, att, y0, hbuf, attq, w, d_out, dout_ld, da, dpre, colpart,
                                                                h_smem ? 1 : 0);
  EBK_LAUNCH_CHECK();
  return EBK_OK;
}

// scratch for colsum partials: a small static device buffer per process would break
// re-entrancy, so the caller passes it through the workspace (see api.cu); this variant
// takes it explicitly.
int colsum_accum_ws(int R, int Ncols, const float* X, int ldx, const float* coef, float* out, float* partial,
                    cudaStream_t st) {
  if (R <= 0 || Ncols <= 0) return EBK_OK;
  int nblk = ceil_div(R, CS_ROWS);
  dim3 grid(ceil_div(Ncols, 128), nblk);
  colsum_partial_kernel<<<grid, 128, 0, st>>>(R, Ncols, X, ldx, coef, partial);
  EBK_LAUNCH_CHECK();
  colsum_final_kernel<<<ceil_div(Ncols, 32), 256, 0, st>>>(nblk, Ncols, partial, out);
  EBK_LAUNCH_CHECK();
  return EBK_OK;
}
// out0[j] += sum_r X[r, j] (j < n0), out1[j - n0] += sum_r X[r, j] (n0 <= j < Ncols): one launch pair for both
int colsum_accum2_ws(int R, int Ncols, int n0, const float* X, int ldx, float* out0, float* out1, float* partial,
                     cudaStream_t st) {
  if (R <= 0 || Ncols <= 0) return EBK_OK;
  int nblk = ceil_div(R, CS_ROWS);
  dim3 grid(ceil_div(Ncols, 128), nblk);
  colsum_partial_kernel<<<grid, 128, 0, st>>>(R, Ncols, X, ldx, nullptr, partial);
  EBK_LAUNCH_CHECK();
  colsum_final2_kernel<<<ceil_div(Ncols, 32), 256, 0, st>>>(nblk, Ncols, n0, partial, out0, out1);
  EBK_LAUNCH_CHECK();
  return EBK_OK;
}
size_t colsum_partial_floats(int R, int Ncols) { return (size_t)ceil_div(R, CS_ROWS) * (size_t)Ncols; }

int round_tf32_copy(float* dst, const float* src, size_t n, cudaStream_t st) {
  if (n == 0) return EBK_OK;
  EBK_CHECK_ARG(n % 4 == 0, "round_tf32_copy: n=%zu must be a multiple of 4", n);
  size_t n4 = n / 4;
  round_tf32_copy_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(reinterpret_cast<float4*>(dst),
                                                                       reinterpret_cast<const float4*>(src), n4);
  EBK_LAUNCH_CHECK();
  return EBK_OK;
}

int split_tf32_copy(float* hi, float* lo, const float* src, size_t n, cudaStream_t st) {
  if (n == 0) return EBK_OK;
  split_tf32_copy_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(hi, lo, src, n);
  EBK_LAUNCH_CHECK();
  return EBK_OK;
}

int embed_rows(int R, int E, int V, const int32_t* tok, const float* table_or_x, Dropout drop, float* xd,
               cudaStream_t st, const PeerTables* peers, long row_offset) {
  if (R <= 0) return EBK_OK;
  EBK_CHECK_ARG(E % 4 == 0, "embed_rows: E=%d must be a multiple of 4", E);
  const long n4 = (long)R * (E / 4);
  static const bool warp_rows = !(getenv("EBK_EMBED_WARP_ROWS") && atoi(getenv("EBK_EMBED_WARP_ROWS")) == 0);
  const bool by_warp = warp_rows && tok != nullptr && E / 4 >= 32;
  if (peers != nullptr && peers->world > 1 && tok != nullptr) {
    EBK_CHECK_ARG(peers->world <= 8 && peers->shard_floats % 4 == 0 && peers->shard_floats > 0, "embed_rows: bad peer table");
    if (by_warp)
      embed_rows_warp_kernel<true><<<(unsigned)((R + 7) / 8), 256, 0, st>>>(
          R, E / 4, V, tok, reinterpret_cast<const float4*>(table_or_x), drop, reinterpret_cast<float4*>(xd), *peers,
          row_offset * (E / 4));
    else
      embed_rows_kernel<true><<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(
          n4, E / 4, V, tok, reinterpret_cast<const float4*>(table_or_x), drop, reinterpret_cast<float4*>(xd), *peers,
          row_offset * (E / 4));
    EBK_LAUNCH_CHECK();
    return EBK_OK;
  }
  PeerTables none;
  none.world = 1;
  none.shard_floats = 0;
  if (by_warp) {
    embed_rows_warp_kernel<false><<<(unsigned)((R + 7) / 8), 256, 0, st>>>(
        R, E / 4, V, tok, reinterpret_cast<const float4*>(table_or_x), drop, reinterpret_cast<float4*>(xd), none,
        row_offset * (E / 4));
    EBK_LAUNCH_CHECK();
    return EBK_OK;
  }
  embed_rows_kernel<false><<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(n4, E / 4, V, tok,
                                                                 reinterpret_cast<const float4*>(table_or_x), drop,
                                                                 reinterpret_cast<float4*>(xd), none, row_offset * (E / 4));
  EBK_LAUNCH_CHECK();
  return EBK_OK;
}

int embed_rows_csr(int R, int E, int V, const int32_t* tok, const float* table, Dropout drop, float* xd, cudaStream_t st,
                   const PeerTables* peers, const TokenCsr& csr) {
  if (R <= 0) return EBK_OK;
  EBK_CHECK_ARG(E % 4 == 0 && tok != nullptr && table != nullptr, "embed_rows_csr: E=%d must be a multiple of 4, tok / table set", E);
  const int E4 = E / 4;
  const float4* src = reinterpret_cast<const float4*>(table);
  float4* dst = reinterpret_cast<float4*>(xd);
  const unsigned gr = (unsigned)((R + 7) / 8);
  if (peers != nullptr && peers->world > 1) {
    EBK_CHECK_ARG(peers->world <= 8 && peers->shard_floats % 4 == 0 && peers->shard_floats > 0, "embed_rows_csr: bad peer table");
    const int ways = peers->world, rows = ceil_div(V, ways);
    const unsigned gt = (unsigned)(((long)ways * rows + 7) / 8);
    embed_rows_csr_kernel<true><<<gt, 256, 0, st>>>(V, E4, csr.count, csr.offset, csr.perm, src, drop, dst, *peers, ways, rows);
    EBK_LAUNCH_CHECK();
    embed_rows_rest_kernel<true><<<gr, 256, 0, st>>>(R, E4, V, tok, csr.count, src, drop, dst, *peers);
    EBK_LAUNCH_CHECK();
    return EBK_OK;
  }
  PeerTables none;
  none.world = 1;
  none.shard_floats = 0;
  embed_rows_csr_kernel<false><<<(unsigned)((V + 7) / 8), 256, 0, st>>>(V, E4, csr.count, csr.offset, csr.perm, src, drop, dst,
                                                                        none, 1, V);
  EBK_LAUNCH_CHECK();
  embed_rows_rest_kernel<false><<<gr, 256, 0, st>>>(R, E4, V, tok, csr.count, src, drop, dst, none);
  EBK_LAUNCH_CHECK();
  return EBK_OK;
}

int scatter_rows_add(int R, int E, int V, const int32_t* tok, const float* dX, Dropout drop, float* d_table,
                     cudaStream_t st) {
  if (R <= 0) return EBK_OK;
  EBK_CHECK_ARG(E % 4 == 0, "scatter: E=%d must be a multiple of 4", E);
  long n = (long)R * (E / 4);
  scatter_rows_add_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(R, E / 4, V, tok,
                                                                       reinterpret_cast<const float4*>(dX), drop,
                                                                       d_table);
  EBK_LAUNCH_CHECK();
  return EBK_OK;
}

}  // namespace ebk

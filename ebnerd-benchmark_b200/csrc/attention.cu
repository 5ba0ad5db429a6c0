// Multi-head self-attention core of layers.SelfAttention (reference layers.py:231-252):
//   S = Q K^T / sqrt(dh);  A = softmax_k(S);  O[k,:] = sum_q A[q,k] V[q,:]   (adjoint_a=True!)
// operating on the packed projection buffer qkv [R, 3D] (R = n_seq*L rows; columns
// Q | K | V, head h at h*dh inside each).  One warp owns one (sequence, head) pair: the
// L<=64 tokens live in that warp's shared-memory slice, lane q owns score row q, so the
// softmax over keys is thread-local (registers / the lane's own smem row) and the
// transposed product only needs a __syncwarp.
#include "ebk_common.cuh"

namespace ebk {
namespace {

constexpr int WARPS = 4;

// per-warp smem floats: Q,K,V (3 * L * DP) + A (L * LP) [+ bwd: dO (L*DP) + dS (L*LP)]
template <int DH>
struct Lay {
  static constexpr int DP = DH + 1;  // odd-ish stride: lane q walks rows without bank conflicts
};

template <int DH>
__global__ void __launch_bounds__(WARPS * 32) attn_fwd_kernel(int n_seq, int L, int nh, int dh,
                                                               const float* __restrict__ qkv,
                                                               float* __restrict__ y) {
  extern __shared__ float smem[];
  constexpr int DP = Lay<DH>::DP;
  const int LP = L | 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int per_warp = 3 * L * DP + L * LP;
  float* Qs = smem + warp * per_warp;
  float* Ks = Qs + L * DP;
  float* Vs = Ks + L * DP;
  float* As = Vs + L * DP;
  const int D = nh * dh;
  const float inv = rsqrtf((float)dh);
  const long total = (long)n_seq * nh;

  for (long item = (long)blockIdx.x * WARPS + warp; item < total; item += (long)gridDim.x * WARPS) {
    const int n = (int)(item / nh), h = (int)(item % nh);
    const float* base = qkv + (long)n * L * 3 * D + h * dh;
    // stage Q,K,V slices [L, dh]
    for (int i = lane; i < L * dh; i += 32) {
      int t = i / dh, d = i % dh;
      const float* row = base + (long)t * 3 * D + d;
      Qs[t * DP + d] = row[0];
      Ks[t * DP + d] = row[D];
      Vs[t * DP + d] = row[2 * D];
    }
    __syncwarp();
    // lane q: score row q
    for (int q = lane; q < L; q += 32) {
      float qr[DH];
#pragma unroll
      for (int d = 0; d < DH; ++d) qr[d] = (d < dh) ? Qs[q * DP + d] : 0.0f;
      float mx = -INFINITY;
      for (int k = 0; k < L; ++k) {
        float s = 0.0f;
#pragma unroll
        for (int d = 0; d < DH; ++d)
          if (d < dh) s = fmaf(qr[d], Ks[k * DP + d], s);
        s *= inv;
        As[q * LP + k] = s;
        mx = fmaxf(mx, s);
      }
      float sum = 0.0f;
      for (int k = 0; k < L; ++k) {
        float e = __expf(As[q * LP + k] - mx);
        As[q * LP + k] = e;
        sum += e;
      }
      float r = 1.0f / sum;
      for (int k = 0; k < L; ++k) As[q * LP + k] *= r;
    }
    __syncwarp();
    // lane k: O[k,:] = sum_q A[q,k] V[q,:]
    for (int k = lane; k < L; k += 32) {
      float o[DH];
#pragma unroll
      for (int d = 0; d < DH; ++d) o[d] = 0.0f;
      for (int q = 0; q < L; ++q) {
        float a = As[q * LP + k];
#pragma unroll
        for (int d = 0; d < DH; ++d)
          if (d < dh) o[d] = fmaf(a, Vs[q * DP + d], o[d]);
      }
      float* out = y + ((long)n * L + k) * D + h * dh;
#pragma unroll
      for (int d = 0; d < DH; ++d)
        if (d < dh) out[d] = o[d];
    }
    __syncwarp();
  }
}

// Backward: given dO (= dy, optionally dropout-masked on read), recompute A and emit
//   dV = A dO ; dA = V dO^T ; dS = A o (dA - rowsum(dA o A)) ; dQ = dS K /sqrt(dh) ; dK = dS^T Q /sqrt(dh)
template <int DH>
__global__ void __launch_bounds__(WARPS * 32) attn_bwd_kernel(int n_seq, int L, int nh, int dh,
                                                               const float* __restrict__ qkv,
                                                               const float* __restrict__ dy, Dropout drop,
                                                               float* __restrict__ dqkv, bool round_out) {
  extern __shared__ float smem[];
  constexpr int DP = Lay<DH>::DP;
  const int LP = L | 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int per_warp = 4 * L * DP + 2 * L * LP;
  float* Qs = smem + warp * per_warp;
  float* Ks = Qs + L * DP;
  float* Vs = Ks + L * DP;
  float* Gs = Vs + L * DP;   // dO
  float* As = Gs + L * DP;   // A
  float* Ds = As + L * LP;   // dS
  const int D = nh * dh;
  const float inv = rsqrtf((float)dh);
  const long total = (long)n_seq * nh;

  for (long item = (long)blockIdx.x * WARPS + warp; item < total; item += (long)gridDim.x * WARPS) {
    const int n = (int)(item / nh), h = (int)(item % nh);
    const float* base = qkv + (long)n * L * 3 * D + h * dh;
    for (int i = lane; i < L * dh; i += 32) {
      int t = i / dh, d = i % dh;
      const float* row = base + (long)t * 3 * D + d;
      Qs[t * DP + d] = row[0];
      Ks[t * DP + d] = row[D];
      Vs[t * DP + d] = row[2 * D];
      long r = (long)n * L + t;
      long c = h * dh + d;
      float g = dy[r * D + c];
      if (drop.on()) g *= drop.factor((uint64_t)r * (uint64_t)D + (uint64_t)c);
      Gs[t * DP + d] = g;
    }
    __syncwarp();
    for (int q = lane; q < L; q += 32) {
      float qr[DH], vr[DH];
#pragma unroll
      for (int d = 0; d < DH; ++d) {
        qr[d] = (d < dh) ? Qs[q * DP + d] : 0.0f;
        vr[d] = (d < dh) ? Vs[q * DP + d] : 0.0f;
      }
      float mx = -INFINITY;
      for (int k = 0; k < L; ++k) {
        float s = 0.0f;
#pragma unroll
        for (int d = 0; d < DH; ++d)
          if (d < dh) s = fmaf(qr[d], Ks[k * DP + d], s);
        s *= inv;
        As[q * LP + k] = s;
        mx = fmaxf(mx, s);
      }
      float sum = 0.0f;
      for (int k = 0; k < L; ++k) {
        float e = __expf(As[q * LP + k] - mx);
        As[q * LP + k] = e;
        sum += e;
      }
      float rs = 1.0f / sum;
      // dA row, rowsum(dA o A), dV row
      float dv[DH];
#pragma unroll
      for (int d = 0; d < DH; ++d) dv[d] = 0.0f;
      float dot = 0.0f;
      for (int k = 0; k < L; ++k) {
        float a = As[q * LP + k] * rs;
        As[q * LP + k] = a;
        float da = 0.0f;
#pragma unroll
        for (int d = 0; d < DH; ++d)
          if (d < dh) {
            float g = Gs[k * DP + d];
            da = fmaf(vr[d], g, da);
            dv[d] = fmaf(a, g, dv[d]);
          }
        Ds[q * LP + k] = da;
        dot = fmaf(da, a, dot);
      }
      float dq[DH];
#pragma unroll
      for (int d = 0; d < DH; ++d) dq[d] = 0.0f;
      for (int k = 0; k < L; ++k) {
        float ds = As[q * LP + k] * (Ds[q * LP + k] - dot);
        Ds[q * LP + k] = ds;
#pragma unroll
        for (int d = 0; d < DH; ++d)
          if (d < dh) dq[d] = fmaf(ds, Ks[k * DP + d], dq[d]);
      }
      float* out = dqkv + ((long)n * L + q) * 3 * D + h * dh;
#pragma unroll
      for (int d = 0; d < DH; ++d)
        if (d < dh) {
          out[d] = round_out ? round_tf32_bits(dq[d] * inv) : dq[d] * inv;
          out[2 * D + d] = round_out ? round_tf32_bits(dv[d]) : dv[d];
        }
    }
    __syncwarp();
    for (int k = lane; k < L; k += 32) {
      float dk[DH];
#pragma unroll
      for (int d = 0; d < DH; ++d) dk[d] = 0.0f;
      for (int q = 0; q < L; ++q) {
        float ds = Ds[q * LP + k];
#pragma unroll
        for (int d = 0; d < DH; ++d)
          if (d < dh) dk[d] = fmaf(ds, Qs[q * DP + d], dk[d]);
      }
      float* out = dqkv + ((long)n * L + k) * 3 * D + D + h * dh;
#pragma unroll
      for (int d = 0; d < DH; ++d)
        if (d < dh) out[d] = round_out ? round_tf32_bits(dk[d] * inv) : dk[d] * inv;
    }
    __syncwarp();
  }
}

template <typename Kern>
int launch_cfg(Kern kern, size_t smem, long total, int* grid) {
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("attention: smem %zu too large: %s", smem, cudaGetErrorString(e));
      return EBK_ERR_CUDA;
    }
  }
  long blocks = (total + WARPS - 1) / WARPS;
  long cap = 148L * 16;
  *grid = (int)(blocks < cap ? blocks : cap);
  return EBK_OK;
}

}  // namespace

int attention_core_fwd(int n_seq, int L, int nh, int dh, const float* qkv, float* y, cudaStream_t st) {
  if (n_seq <= 0) return EBK_OK;
  EBK_CHECK_ARG(L >= 1 && L <= 64 && dh >= 1 && dh <= 32 && nh >= 1, "attention: need 1<=L<=64, 1<=dh<=32 (L=%d dh=%d)", L, dh);
  const int LP = L | 1;
  long total = (long)n_seq * nh;
  int grid;
#define RUN(DH_)                                                                        \
  {                                                                                     \
    size_t smem = (size_t)WARPS * (3 * L * (DH_ + 1) + L * LP) * sizeof(float);         \
    EBK_TRY(launch_cfg(attn_fwd_kernel<DH_>, smem, total, &grid));                      \
    attn_fwd_kernel<DH_><<<grid, WARPS * 32, smem, st>>>(n_seq, L, nh, dh, qkv, y);     \
  }
  if (dh <= 16) RUN(16) else if (dh <= 20) RUN(20) else RUN(32)
#undef RUN
  EBK_LAUNCH_CHECK();
  return EBK_OK;
}

int attention_core_bwd(int n_seq, int L, int nh, int dh, const float* qkv, const float* dy, Dropout drop,
                       float* dqkv, bool round_out, cudaStream_t st) {
  if (n_seq <= 0) return EBK_OK;
  EBK_CHECK_ARG(L >= 1 && L <= 64 && dh >= 1 && dh <= 32 && nh >= 1, "attention: need 1<=L<=64, 1<=dh<=32 (L=%d dh=%d)", L, dh);
  const int LP = L | 1;
  long total = (long)n_seq * nh;
  int grid;
#define RUN(DH_)                                                                               \
  {                                                                                            \
    size_t smem = (size_t)WARPS * (4 * L * (DH_ + 1) + 2 * L * LP) * sizeof(float);            \
    EBK_TRY(launch_cfg(attn_bwd_kernel<DH_>, smem, total, &grid));                             \
    attn_bwd_kernel<DH_><<<grid, WARPS * 32, smem, st>>>(n_seq, L, nh, dh, qkv, dy, drop, dqkv, round_out); \
  }
  if (dh <= 16) RUN(16) else if (dh <= 20) RUN(20) else RUN(32)
#undef RUN
  EBK_LAUNCH_CHECK();
  return EBK_OK;
}

}  // namespace ebk

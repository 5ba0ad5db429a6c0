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

// ---------------------------------------------------------------------------------------------
// Vectorised variants (dh % 4 == 0, the shipped configurations dh = 20 and 16): rows are 16-byte
// aligned in shared memory, K/V/dO rows are read with broadcast LDS.128 (4 FMAs per shared load
// instead of 1), global traffic is 128-bit.  Same math, same thread ownership as above.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float dot4(const float4& a, const float4& b, float acc) {
  acc = fmaf(a.x, b.x, acc);
  acc = fmaf(a.y, b.y, acc);
  acc = fmaf(a.z, b.z, acc);
  return fmaf(a.w, b.w, acc);
}
__device__ __forceinline__ void axpy4(float a, const float4& x, float4& y) {
  y.x = fmaf(a, x.x, y.x);
  y.y = fmaf(a, x.y, y.y);
  y.z = fmaf(a, x.z, y.z);
  y.w = fmaf(a, x.w, y.w);
}

template <int DH>
__global__ void __launch_bounds__(WARPS * 32) attn_fwd_v4_kernel(int n_seq, int L, int nh, const float* __restrict__ qkv,
                                                                  float* __restrict__ y) {
  extern __shared__ __align__(16) float smem[];
  constexpr int D4 = DH / 4;
  const int LP = L | 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int per_warp = (3 * L * DH + L * LP + 3) & ~3;
  float* Qs = smem + warp * per_warp;
  float* Ks = Qs + L * DH;
  float* Vs = Ks + L * DH;
  float* As = Vs + L * DH;
  const int D = nh * DH;
  const float inv = rsqrtf((float)DH);
  const long total = (long)n_seq * nh;

  for (long item = (long)blockIdx.x * WARPS + warp; item < total; item += (long)gridDim.x * WARPS) {
    const int n = (int)(item / nh), h = (int)(item % nh);
    const float* base = qkv + (long)n * L * 3 * D + h * DH;
    for (int i = lane; i < 3 * L * D4; i += 32) {
      const int mtx = i / (L * D4), rem = i - mtx * (L * D4);
      const int t = rem / D4, j = rem - t * D4;
      const float4 v = __ldg(reinterpret_cast<const float4*>(base + (long)t * 3 * D + mtx * D) + j);
      reinterpret_cast<float4*>(Qs + mtx * L * DH + t * DH)[j] = v;
    }
    __syncwarp();
    for (int q = lane; q < L; q += 32) {
      float4 qr[D4];
#pragma unroll
      for (int j = 0; j < D4; ++j) qr[j] = reinterpret_cast<const float4*>(Qs + q * DH)[j];
      float mx = -INFINITY;
      for (int k = 0; k < L; ++k) {
        const float4* kr = reinterpret_cast<const float4*>(Ks + k * DH);
        float s = 0.0f;
#pragma unroll
        for (int j = 0; j < D4; ++j) s = dot4(qr[j], kr[j], s);
        s *= inv;
        As[q * LP + k] = s;
        mx = fmaxf(mx, s);
      }
      float sum = 0.0f;
      for (int k = 0; k < L; ++k) {
        const float e = __expf(As[q * LP + k] - mx);
        As[q * LP + k] = e;
        sum += e;
      }
      const float r = 1.0f / sum;
      for (int k = 0; k < L; ++k) As[q * LP + k] *= r;
    }
    __syncwarp();
    for (int k = lane; k < L; k += 32) {
      float4 o[D4];
#pragma unroll
      for (int j = 0; j < D4; ++j) o[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int q = 0; q < L; ++q) {
        const float a = As[q * LP + k];
        const float4* vr = reinterpret_cast<const float4*>(Vs + q * DH);
#pragma unroll
        for (int j = 0; j < D4; ++j) axpy4(a, vr[j], o[j]);
      }
      float4* out = reinterpret_cast<float4*>(y + ((long)n * L + k) * D + h * DH);
#pragma unroll
      for (int j = 0; j < D4; ++j) out[j] = o[j];
    }
    __syncwarp();
  }
}

// Optional second copy of dQKV in the packed B-operand layout of the weight-gradient GEMM
// (gemm_tf32_sm100.cu: [n-tile][k-step] blocks, SWIZZLE_128B_BASE32B atoms), so that GEMM can fetch
// its B tiles with one bulk copy per stage.
struct PackedOut {
  float* ptr;        // nullptr = disabled
  int BN, groups;    // tile width (multiple of 32), BN/32
  int ksteps;        // ceil(R / 32)
  int block_floats;  // floats per [n-tile][k-step] block
  __device__ __forceinline__ float* at(long r, int col) const {
    const int tn = col / BN, cl = col - tn * BN;
    const int ks = (int)(r >> 5), kr = (int)(r & 31);
    const int c = cl >> 2;
    const unsigned unit32 = (unsigned)(((c & 7) >> 1) ^ (kr & 3));
    const unsigned off = (unsigned)((kr >> 2) * groups + (c >> 3)) * 512u + (unsigned)(kr & 3) * 128u + (unit32 << 5) +
                         (unsigned)((c & 1) << 4);
    return ptr + ((long)tn * ksteps + ks) * block_floats + (off >> 2) + (cl & 3);
  }
};

template <int DH>
__global__ void __launch_bounds__(WARPS * 32) attn_bwd_v4_kernel(int n_seq, int L, int nh, const float* __restrict__ qkv,
                                                                  const float* __restrict__ dy, Dropout drop,
                                                                  float* __restrict__ dqkv, bool round_out,
                                                                  PackedOut pk) {
  extern __shared__ __align__(16) float smem[];
  constexpr int D4 = DH / 4;
  const int LP = L | 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int per_warp = (4 * L * DH + 2 * L * LP + 3) & ~3;
  float* Qs = smem + warp * per_warp;
  float* Ks = Qs + L * DH;
  float* Vs = Ks + L * DH;
  float* Gs = Vs + L * DH;   // dO
  float* As = Gs + L * DH;   // A
  float* Ds = As + L * LP;   // dS
  const int D = nh * DH;
  const float inv = rsqrtf((float)DH);
  const long total = (long)n_seq * nh;

  for (long item = (long)blockIdx.x * WARPS + warp; item < total; item += (long)gridDim.x * WARPS) {
    const int n = (int)(item / nh), h = (int)(item % nh);
    const float* base = qkv + (long)n * L * 3 * D + h * DH;
    for (int i = lane; i < 3 * L * D4; i += 32) {
      const int mtx = i / (L * D4), rem = i - mtx * (L * D4);
      const int t = rem / D4, j = rem - t * D4;
      const float4 v = __ldg(reinterpret_cast<const float4*>(base + (long)t * 3 * D + mtx * D) + j);
      reinterpret_cast<float4*>(Qs + mtx * L * DH + t * DH)[j] = v;
    }
    for (int i = lane; i < L * D4; i += 32) {
      const int t = i / D4, j = i - t * D4;
      const long r = (long)n * L + t;
      const long c = h * DH + j * 4;
      float4 g = __ldg(reinterpret_cast<const float4*>(dy + r * D + c));
      if (drop.on()) {
        const float4 f = drop.factor4((uint64_t)r * (uint64_t)D + (uint64_t)c);
        g.x *= f.x; g.y *= f.y; g.z *= f.z; g.w *= f.w;
      }
      reinterpret_cast<float4*>(Gs + t * DH)[j] = g;
    }
    __syncwarp();
    for (int q = lane; q < L; q += 32) {
      float4 qr[D4], vr[D4];
#pragma unroll
      for (int j = 0; j < D4; ++j) {
        qr[j] = reinterpret_cast<const float4*>(Qs + q * DH)[j];
        vr[j] = reinterpret_cast<const float4*>(Vs + q * DH)[j];
      }
      float mx = -INFINITY;
      for (int k = 0; k < L; ++k) {
        const float4* kr = reinterpret_cast<const float4*>(Ks + k * DH);
        float s = 0.0f;
#pragma unroll
        for (int j = 0; j < D4; ++j) s = dot4(qr[j], kr[j], s);
        s *= inv;
        As[q * LP + k] = s;
        mx = fmaxf(mx, s);
      }
      float sum = 0.0f;
      for (int k = 0; k < L; ++k) {
        const float e = __expf(As[q * LP + k] - mx);
        As[q * LP + k] = e;
        sum += e;
      }
      const float rs = 1.0f / sum;
      float4 dv[D4];
#pragma unroll
      for (int j = 0; j < D4; ++j) dv[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      float dot = 0.0f;
      for (int k = 0; k < L; ++k) {
        const float a = As[q * LP + k] * rs;
        As[q * LP + k] = a;
        const float4* gr = reinterpret_cast<const float4*>(Gs + k * DH);
        float da = 0.0f;
#pragma unroll
        for (int j = 0; j < D4; ++j) {
          const float4 g = gr[j];
          da = dot4(vr[j], g, da);
          axpy4(a, g, dv[j]);
        }
        Ds[q * LP + k] = da;
        dot = fmaf(da, a, dot);
      }
      float4 dq[D4];
#pragma unroll
      for (int j = 0; j < D4; ++j) dq[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int k = 0; k < L; ++k) {
        const float ds = As[q * LP + k] * (Ds[q * LP + k] - dot);
        Ds[q * LP + k] = ds;
        const float4* kr = reinterpret_cast<const float4*>(Ks + k * DH);
#pragma unroll
        for (int j = 0; j < D4; ++j) axpy4(ds, kr[j], dq[j]);
      }
      const long r = (long)n * L + q;
      float* out = dqkv + r * 3 * D + h * DH;
#pragma unroll
      for (int j = 0; j < D4; ++j) {
        float4 a = make_float4(dq[j].x * inv, dq[j].y * inv, dq[j].z * inv, dq[j].w * inv);
        float4 b = dv[j];
        if (round_out) {
          a.x = round_tf32_bits(a.x); a.y = round_tf32_bits(a.y); a.z = round_tf32_bits(a.z); a.w = round_tf32_bits(a.w);
          b.x = round_tf32_bits(b.x); b.y = round_tf32_bits(b.y); b.z = round_tf32_bits(b.z); b.w = round_tf32_bits(b.w);
        }
        reinterpret_cast<float4*>(out)[j] = a;
        reinterpret_cast<float4*>(out + 2 * D)[j] = b;
        if (pk.ptr) {
          *reinterpret_cast<float4*>(pk.at(r, h * DH + j * 4)) = a;
          *reinterpret_cast<float4*>(pk.at(r, 2 * D + h * DH + j * 4)) = b;
        }
      }
    }
    __syncwarp();
    for (int k = lane; k < L; k += 32) {
      float4 dk[D4];
#pragma unroll
      for (int j = 0; j < D4; ++j) dk[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int q = 0; q < L; ++q) {
        const float ds = Ds[q * LP + k];
        const float4* qr = reinterpret_cast<const float4*>(Qs + q * DH);
#pragma unroll
        for (int j = 0; j < D4; ++j) axpy4(ds, qr[j], dk[j]);
      }
      const long r = (long)n * L + k;
      float* out = dqkv + r * 3 * D + D + h * DH;
#pragma unroll
      for (int j = 0; j < D4; ++j) {
        float4 a = make_float4(dk[j].x * inv, dk[j].y * inv, dk[j].z * inv, dk[j].w * inv);
        if (round_out) {
          a.x = round_tf32_bits(a.x); a.y = round_tf32_bits(a.y); a.z = round_tf32_bits(a.z); a.w = round_tf32_bits(a.w);
        }
        reinterpret_cast<float4*>(out)[j] = a;
        if (pk.ptr) *reinterpret_cast<float4*>(pk.at(r, D + h * DH + j * 4)) = a;
      }
    }
    __syncwarp();
  }
}

// zero rows [R, 32*ksteps) of every n-tile block of the last k-step (they multiply zero-filled A rows)
__global__ void zero_packed_tail_kernel(PackedOut pk, long R, int tiles_n) {
  const long rows_pad = (long)pk.ksteps * 32;
  const long ntail = rows_pad - R;
  const long total = ntail * tiles_n * pk.BN;
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int col = (int)(i % ((long)tiles_n * pk.BN));
  const long r = R + i / ((long)tiles_n * pk.BN);
  *pk.at(r, col) = 0.0f;
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
  const bool v4 = (dh == 8 || dh == 16 || dh == 20 || dh == 32) && ((reinterpret_cast<uintptr_t>(qkv) & 15) == 0) &&
                  ((reinterpret_cast<uintptr_t>(y) & 15) == 0);
#define RUNV(DH_)                                                                                   \
  {                                                                                                 \
    size_t smem = (size_t)WARPS * ((3 * L * DH_ + L * LP + 3) & ~3) * sizeof(float);                \
    EBK_TRY(launch_cfg(attn_fwd_v4_kernel<DH_>, smem, total, &grid));                               \
    attn_fwd_v4_kernel<DH_><<<grid, WARPS * 32, smem, st>>>(n_seq, L, nh, qkv, y);                  \
  }
#define RUN(DH_)                                                                        \
  {                                                                                     \
    size_t smem = (size_t)WARPS * (3 * L * (DH_ + 1) + L * LP) * sizeof(float);         \
    EBK_TRY(launch_cfg(attn_fwd_kernel<DH_>, smem, total, &grid));                      \
    attn_fwd_kernel<DH_><<<grid, WARPS * 32, smem, st>>>(n_seq, L, nh, dh, qkv, y);     \
  }
  if (v4 && dh == 20) RUNV(20)
  else if (v4 && dh == 16) RUNV(16)
  else if (v4 && dh == 8) RUNV(8)
  else if (v4 && dh == 32) RUNV(32)
  else if (dh <= 16) RUN(16) else if (dh <= 20) RUN(20) else RUN(32)
#undef RUN
#undef RUNV
  EBK_LAUNCH_CHECK();
  return EBK_OK;
}

int attention_core_bwd(int n_seq, int L, int nh, int dh, const float* qkv, const float* dy, Dropout drop,
                       float* dqkv, bool round_out, cudaStream_t st, float* dqkv_packed, int packed_bn) {
  if (n_seq <= 0) return EBK_OK;
  EBK_CHECK_ARG(L >= 1 && L <= 64 && dh >= 1 && dh <= 32 && nh >= 1, "attention: need 1<=L<=64, 1<=dh<=32 (L=%d dh=%d)", L, dh);
  const int LP = L | 1;
  long total = (long)n_seq * nh;
  int grid;
  const bool v4 = (dh == 8 || dh == 16 || dh == 20 || dh == 32) && ((reinterpret_cast<uintptr_t>(qkv) & 15) == 0) &&
                  ((reinterpret_cast<uintptr_t>(dy) & 15) == 0) && ((reinterpret_cast<uintptr_t>(dqkv) & 15) == 0);
  EBK_CHECK_ARG(dqkv_packed == nullptr || v4, "attention_bwd: packed output needs the vectorised path");
  PackedOut pk{dqkv_packed, packed_bn, packed_bn / 32, 0, 0};
  const long R = (long)n_seq * L;
  if (dqkv_packed) {
    pk.ksteps = (int)((R + 31) / 32);
    pk.block_floats = packed_bn * 32;
    const int tiles_n = (3 * nh * dh + packed_bn - 1) / packed_bn;
    const long ntail = (long)pk.ksteps * 32 - R;
    if (ntail > 0) {
      const long tot = ntail * tiles_n * packed_bn;
      zero_packed_tail_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(pk, R, tiles_n);
      EBK_LAUNCH_CHECK();
    }
  }
#define RUNV(DH_)                                                                                                \
  {                                                                                                              \
    size_t smem = (size_t)WARPS * ((4 * L * DH_ + 2 * L * LP + 3) & ~3) * sizeof(float);                         \
    EBK_TRY(launch_cfg(attn_bwd_v4_kernel<DH_>, smem, total, &grid));                                            \
    attn_bwd_v4_kernel<DH_><<<grid, WARPS * 32, smem, st>>>(n_seq, L, nh, qkv, dy, drop, dqkv, round_out, pk);   \
  }
#define RUN(DH_)                                                                               \
  {                                                                                            \
    size_t smem = (size_t)WARPS * (4 * L * (DH_ + 1) + 2 * L * LP) * sizeof(float);            \
    EBK_TRY(launch_cfg(attn_bwd_kernel<DH_>, smem, total, &grid));                             \
    attn_bwd_kernel<DH_><<<grid, WARPS * 32, smem, st>>>(n_seq, L, nh, dh, qkv, dy, drop, dqkv, round_out); \
  }
  if (v4 && dh == 20) RUNV(20)
  else if (v4 && dh == 16) RUNV(16)
  else if (v4 && dh == 8) RUNV(8)
  else if (v4 && dh == 32) RUNV(32)
  else if (dh <= 16) RUN(16) else if (dh <= 20) RUN(20) else RUN(32)
#undef RUN
#undef RUNV
  EBK_LAUNCH_CHECK();
  return EBK_OK;
}

}  // namespace ebk

// Tensor-core variant of the per-head attention core for TRAINING (EBK_MATH_TF32):
//   S = Q K^T / sqrt(dh);  A = softmax_k(S);  O[k,:] = sum_q A[q,k] V[q,:]   (layers.py:231-252, adjoint_a=True)
// and its backward.  One warp owns one (sequence, head) pair padded to 32 tokens x DP head dims; the
// 32x32xDP products run on warp-level mma.sync.m16n8k8 tf32 (fp32 accumulate) with operands rounded to
// nearest tf32 when they are staged in shared memory; softmax and the dS algebra stay in fp32 registers
// in the accumulator-fragment layout (row reductions = two quad shuffles).  The sequences here are 30x30x20
// per head -- far too small for a 128-row tcgen05 tile, which is why this is the one place the warp-level
// MMA is used.  Inference (3xTF32 / fp32 modes) keeps the exact fp32 kernels of attention.cu.
#include "ebk_common.cuh"

namespace ebk {
namespace {

constexpr int WARPS = 4;
constexpr int LP = 32;  // padded sequence length

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ float rn(float x) { return __uint_as_float(__float_as_uint(x) + 0x1000u); }
__device__ __forceinline__ uint32_t u(float x) { return __float_as_uint(x); }

// strides (floats) chosen so the fragment loads below are bank-conflict free:
//   "row" operands (index = row*stride + col, lanes vary row by g and col by t): stride = 4 (mod 8)
//   "transposed" operands (lanes vary row by t and col by g):                   stride = 8 or 24 (mod 32)
template <int DP> struct Str {
  static constexpr int QS = DP + 4;               // Q, K, dO, and V-as-row operand
  static constexpr int VS = (DP == 32) ? 40 : 24; // V / Q as "transposed" B operand
  static constexpr int PS = 40;                   // 32x32 score-shaped matrices
};

// A-fragment (16x8) of a row-major matrix M[row][col]: rows r0.., cols c0..
__device__ __forceinline__ void lda_row(uint32_t (&a)[4], const float* M, int stride, int r0, int c0, int g, int t) {
  a[0] = u(M[(r0 + g) * stride + c0 + t]);
  a[1] = u(M[(r0 + g + 8) * stride + c0 + t]);
  a[2] = u(M[(r0 + g) * stride + c0 + t + 4]);
  a[3] = u(M[(r0 + g + 8) * stride + c0 + t + 4]);
}
// A-fragment of M^T where M is row-major: element (row, col) = M[col][row]
__device__ __forceinline__ void lda_tr(uint32_t (&a)[4], const float* M, int stride, int r0, int c0, int g, int t) {
  a[0] = u(M[(c0 + t) * stride + r0 + g]);
  a[1] = u(M[(c0 + t) * stride + r0 + g + 8]);
  a[2] = u(M[(c0 + t + 4) * stride + r0 + g]);
  a[3] = u(M[(c0 + t + 4) * stride + r0 + g + 8]);
}
// B-fragment (8x8) with B(k, n) = M[n][k]  (M row-major, "col" operand)
__device__ __forceinline__ void ldb_nk(uint32_t (&b)[2], const float* M, int stride, int k0, int n0, int g, int t) {
  b[0] = u(M[(n0 + g) * stride + k0 + t]);
  b[1] = u(M[(n0 + g) * stride + k0 + t + 4]);
}
// B-fragment with B(k, n) = M[k][n]
__device__ __forceinline__ void ldb_kn(uint32_t (&b)[2], const float* M, int stride, int k0, int n0, int g, int t) {
  b[0] = u(M[(k0 + t) * stride + n0 + g]);
  b[1] = u(M[(k0 + t + 4) * stride + n0 + g]);
}

// stage one [L, dh] slice (global row stride ld) into smem [32][stride], zero padded, tf32-rounded
__device__ __forceinline__ void stage(float* dst, int stride, const float* src, long ld, int L, int dh, int DPv, int lane,
                                      const Dropout* drop, long r0, int c0, int Dfull) {
  const int d4 = DPv >> 2;
  for (int i = lane; i < LP * d4; i += 32) {
    const int t = i / d4, j = i - t * d4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t < L && j * 4 < dh) {
      v = __ldg(reinterpret_cast<const float4*>(src + (long)t * ld) + j);
      if (drop != nullptr && drop->on()) {
        const float4 f = drop->factor4((uint64_t)(r0 + t) * (uint64_t)Dfull + (uint64_t)(c0 + j * 4));
        v.x *= f.x; v.y *= f.y; v.z *= f.z; v.w *= f.w;
      }
    }
    float* p = dst + t * stride + j * 4;
    p[0] = rn(v.x); p[1] = rn(v.y); p[2] = rn(v.z); p[3] = rn(v.w);
  }
}

// Stage NM [L, dh] slices at once: ALL global loads are issued before the first shared store, so a warp
// pays one memory round trip per item instead of one per matrix.  src[m] row stride ld[m]; dst[m] row
// stride st[m]; zero padded to [32][DP]; rounded to tf32.  Only the LAST matrix may carry a dropout mask.
template <int DP, int NM>
__device__ __forceinline__ void stage_all(float* const (&dst)[NM], const int (&st)[NM], const float* const (&src)[NM],
                                          const long (&ld)[NM], int L, int dh, int lane, const Dropout* drop, long r0,
                                          int c0, int Dfull) {
  constexpr int D4 = DP / 4;
  constexpr int ITER = LP * D4 / 32;  // float4 chunks per lane per matrix
  float4 v[NM][ITER];
#pragma unroll
  for (int m = 0; m < NM; ++m)
#pragma unroll
    for (int it = 0; it < ITER; ++it) {
      const int i = lane + it * 32;
      const int t = i / D4, j = i - t * D4;
      v[m][it] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (t < L && j * 4 < dh) v[m][it] = __ldg(reinterpret_cast<const float4*>(src[m] + (long)t * ld[m]) + j);
    }
#pragma unroll
  for (int m = 0; m < NM; ++m)
#pragma unroll
    for (int it = 0; it < ITER; ++it) {
      const int i = lane + it * 32;
      const int t = i / D4, j = i - t * D4;
      float4 x = v[m][it];
      if (m == NM - 1 && drop != nullptr && drop->on() && t < L && j * 4 < dh) {
        const float4 f = drop->factor4((uint64_t)(r0 + t) * (uint64_t)Dfull + (uint64_t)(c0 + j * 4));
        x.x *= f.x; x.y *= f.y; x.z *= f.z; x.w *= f.w;
      }
      *reinterpret_cast<float4*>(dst[m] + t * st[m] + j * 4) = make_float4(rn(x.x), rn(x.y), rn(x.z), rn(x.w));
    }
}

// S (32x32, accumulator fragments acc[mt][nt][4]) = X Y^T over DP, X/Y row-major with stride QS
template <int DP>
__device__ __forceinline__ void gemm_xyT(float (&acc)[2][4][4], const float* X, const float* Y, int g, int t) {
  constexpr int QS = Str<DP>::QS;
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[mt][nt][e] = 0.0f;
#pragma unroll
  for (int ks = 0; ks < DP / 8; ++ks) {
    uint32_t a[2][4], b[4][2];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) lda_row(a[mt], X, QS, mt * 16, ks * 8, g, t);
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) ldb_nk(b[nt], Y, QS, ks * 8, nt * 8, g, t);
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) mma_tf32(acc[mt][nt], a[mt], b[nt]);
  }
}

// in-register row softmax of S*inv over the first L columns (fragment layout); masked columns -> 0
__device__ __forceinline__ void softmax_rows(float (&acc)[2][4][4], float inv, int L, int t) {
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {  // rows g (regs 0,1) and g+8 (regs 2,3)
      float mx = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int col = nt * 8 + 2 * t + e;
          float s = acc[mt][nt][hf * 2 + e] * inv;
          s = col < L ? s : -INFINITY;
          acc[mt][nt][hf * 2 + e] = s;
          mx = fmaxf(mx, s);
        }
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      float sum = 0.0f;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const float ex = __expf(acc[mt][nt][hf * 2 + e] - mx);
          acc[mt][nt][hf * 2 + e] = ex;
          sum += ex;
        }
      sum += __shfl_xor_sync(0xffffffffu, sum, 1);
      sum += __shfl_xor_sync(0xffffffffu, sum, 2);
      const float r = 1.0f / sum;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) acc[mt][nt][hf * 2 + e] *= r;
    }
}

// store a 32x32 accumulator-fragment matrix to smem [32][PS] (tf32-rounded: it feeds later MMAs)
__device__ __forceinline__ void store_frag(float* P, const float (&acc)[2][4][4], int g, int t) {
  constexpr int PS = 40;
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const int col = nt * 8 + 2 * t;
      *reinterpret_cast<float2*>(P + (mt * 16 + g) * PS + col) = make_float2(rn(acc[mt][nt][0]), rn(acc[mt][nt][1]));
      *reinterpret_cast<float2*>(P + (mt * 16 + g + 8) * PS + col) = make_float2(rn(acc[mt][nt][2]), rn(acc[mt][nt][3]));
    }
}

template <int DP>
__global__ void __launch_bounds__(WARPS * 32) attn_fwd_mma_kernel(int n_seq, int L, int nh, int dh,
                                                                  const float* __restrict__ qkv, float* __restrict__ y,
                                                                  Dropout drop, bool round_out) {
  extern __shared__ __align__(16) float smem[];
  constexpr int QS = Str<DP>::QS, VS = Str<DP>::VS, PS = Str<DP>::PS;
  constexpr int QK_FLOATS = (2 * LP * QS > LP * PS) ? 2 * LP * QS : LP * PS;
  constexpr int PER_WARP = QK_FLOATS + LP * VS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  float* Qs = smem + warp * PER_WARP;
  float* Ks = Qs + LP * QS;
  float* Vs = Qs + QK_FLOATS;
  float* Ps = Qs;  // overlays Q/K once the scores are in registers
  const int D = nh * dh;
  const float inv = rsqrtf((float)dh);
  const long total = (long)n_seq * nh;
  for (long item = (long)blockIdx.x * WARPS + warp; item < total; item += (long)gridDim.x * WARPS) {
    const int n = (int)(item / nh), h = (int)(item % nh);
    const float* base = qkv + (long)n * L * 3 * D + h * dh;
    {
      float* const dst[3] = {Qs, Ks, Vs};
      const int st[3] = {QS, QS, VS};
      const float* const src[3] = {base, base + D, base + 2 * D};
      const long ld[3] = {3L * D, 3L * D, 3L * D};
      stage_all<DP, 3>(dst, st, src, ld, L, dh, lane, nullptr, 0, 0, 0);
    }
    __syncwarp();
    float acc[2][4][4];
    gemm_xyT<DP>(acc, Qs, Ks, g, t);
    softmax_rows(acc, inv, L, t);
    __syncwarp();  // everyone is done reading Q/K
    store_frag(Ps, acc, g, t);
    __syncwarp();
    // O[k, d] = sum_q P[q, k] V[q, d]:  A operand = P^T, B operand = V
    float o[2][DP / 8][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < DP / 8; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) o[mt][nt][e] = 0.0f;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t a[2][4], b[DP / 8][2];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) lda_tr(a[mt], Ps, PS, mt * 16, ks * 8, g, t);
#pragma unroll
      for (int nt = 0; nt < DP / 8; ++nt) ldb_kn(b[nt], Vs, VS, ks * 8, nt * 8, g, t);
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < DP / 8; ++nt) mma_tf32(o[mt][nt], a[mt], b[nt]);
    }
    float* out = y + (long)n * L * D + h * dh;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < DP / 8; ++nt) {
        const int col = nt * 8 + 2 * t;
        if (col < dh) {
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            const int r = mt * 16 + g + hf * 8;
            if (r >= L) continue;
            float2 v = make_float2(o[mt][nt][hf * 2], o[mt][nt][hf * 2 + 1]);
            if (drop.on()) {  // the layer after the attention only ever reads dropout(y): store it masked
              const uint64_t idx = (uint64_t)((long)n * L + r) * (uint64_t)D + (uint64_t)(h * dh + col);
              const float4 f = drop.factor4_group(idx >> 2);
              v.x *= (idx & 2ull) ? f.z : f.x;
              v.y *= (idx & 2ull) ? f.w : f.y;
            }
            if (round_out) {
              v.x = round_tf32_bits(v.x);
              v.y = round_tf32_bits(v.y);
            }
            *reinterpret_cast<float2*>(out + (long)r * D + col) = v;
          }
        }
      }
    __syncwarp();
  }
}

// Optional second copy of dQKV in the packed B-operand layout of the weight-gradient GEMM (see attention.cu)
struct PackedOut {
  float* ptr;
  int BN, groups, ksteps, block_floats;
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

// write a [32 x DP] accumulator-fragment matrix (rows = tokens) to dqkv columns [col0, col0+dh) (+ packed copy)
template <int DP>
__device__ __forceinline__ void store_grad(const float (&acc)[2][DP / 8][4], float scale, float* __restrict__ dqkv, long row0,
                                           int ld, int col0, int L, int dh, bool round_out, const PackedOut& pk, int g,
                                           int t) {
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < DP / 8; ++nt) {
      const int col = nt * 8 + 2 * t;
      if (col >= dh) continue;
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const int r = mt * 16 + g + hf * 8;
        if (r >= L) continue;
        float2 v = make_float2(acc[mt][nt][hf * 2] * scale, acc[mt][nt][hf * 2 + 1] * scale);
        if (round_out) {
          v.x = round_tf32_bits(v.x);
          v.y = round_tf32_bits(v.y);
        }
        *reinterpret_cast<float2*>(dqkv + (row0 + r) * ld + col0 + col) = v;
        if (pk.ptr) *reinterpret_cast<float2*>(pk.at(row0 + r, col0 + col)) = v;
      }
    }
}

template <int DP>
__global__ void __launch_bounds__(WARPS * 32) attn_bwd_mma_kernel(int n_seq, int L, int nh, int dh,
                                                                  const float* __restrict__ qkv,
                                                                  const float* __restrict__ dy, Dropout drop,
                                                                  float* __restrict__ dqkv, bool round_out, PackedOut pk) {
  extern __shared__ __align__(16) float smem[];
  constexpr int QS = Str<DP>::QS, VS = Str<DP>::VS, PS = Str<DP>::PS;
  // Q, K, V, dO with stride QS (conflict-free as "row" operands; 2-way conflicts when read as the
  // transposed B operand, accepted to keep 8 warps resident per SM); A and dS as 32x32 matrices (stride PS)
  constexpr int PER_WARP = 4 * LP * QS + 2 * LP * PS;
  (void)VS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  float* Qs = smem + warp * PER_WARP;
  float* Ks = Qs + LP * QS;
  float* Vs = Ks + LP * QS;
  float* Gs = Vs + LP * QS;        // dO
  float* As = Gs + LP * QS;        // A
  float* Ds = As + LP * PS;        // dS
  const int D = nh * dh;
  const float inv = rsqrtf((float)dh);
  const long total = (long)n_seq * nh;
  for (long item = (long)blockIdx.x * WARPS + warp; item < total; item += (long)gridDim.x * WARPS) {
    const int n = (int)(item / nh), h = (int)(item % nh);
    const float* base = qkv + (long)n * L * 3 * D + h * dh;
    const long row0 = (long)n * L;
    {
      float* const dst[4] = {Qs, Ks, Vs, Gs};
      const int st[4] = {QS, QS, QS, QS};
      const float* const src[4] = {base, base + D, base + 2 * D, dy + row0 * D + h * dh};
      const long ld[4] = {3L * D, 3L * D, 3L * D, (long)D};
      stage_all<DP, 4>(dst, st, src, ld, L, dh, lane, &drop, row0, h * dh, D);
    }
    __syncwarp();
    // A = softmax(Q K^T / sqrt(dh))
    float a_acc[2][4][4];
    gemm_xyT<DP>(a_acc, Qs, Ks, g, t);
    softmax_rows(a_acc, inv, L, t);
    // dA = V dO^T   (same fragment layout as A)
    float d_acc[2][4][4];
    gemm_xyT<DP>(d_acc, Vs, Gs, g, t);
    // dS = A o (dA - rowsum(dA o A))
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        float dot = 0.0f;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
          for (int e = 0; e < 2; ++e) dot = fmaf(d_acc[mt][nt][hf * 2 + e], a_acc[mt][nt][hf * 2 + e], dot);
        dot += __shfl_xor_sync(0xffffffffu, dot, 1);
        dot += __shfl_xor_sync(0xffffffffu, dot, 2);
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
          for (int e = 0; e < 2; ++e)
            d_acc[mt][nt][hf * 2 + e] = a_acc[mt][nt][hf * 2 + e] * (d_acc[mt][nt][hf * 2 + e] - dot);
      }
    store_frag(As, a_acc, g, t);
    store_frag(Ds, d_acc, g, t);
    __syncwarp();
    float* out = dqkv;
    {  // dV[q, d] = sum_k A[q, k] dO[k, d]
      float acc[2][DP / 8][4] = {};
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t a[2][4], b[DP / 8][2];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) lda_row(a[mt], As, PS, mt * 16, ks * 8, g, t);
#pragma unroll
        for (int nt = 0; nt < DP / 8; ++nt) ldb_kn(b[nt], Gs, QS, ks * 8, nt * 8, g, t);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int nt = 0; nt < DP / 8; ++nt) mma_tf32(acc[mt][nt], a[mt], b[nt]);
      }
      store_grad<DP>(acc, 1.0f, out, row0, 3 * D, 2 * D + h * dh, L, dh, round_out, pk, g, t);
    }
    {  // dQ[q, d] = sum_k dS[q, k] K[k, d] / sqrt(dh)   (K as B(k, n) = K[k][d]: row stride QS)
      float acc[2][DP / 8][4] = {};
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t a[2][4], b[DP / 8][2];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) lda_row(a[mt], Ds, PS, mt * 16, ks * 8, g, t);
#pragma unroll
        for (int nt = 0; nt < DP / 8; ++nt) ldb_kn(b[nt], Ks, QS, ks * 8, nt * 8, g, t);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int nt = 0; nt < DP / 8; ++nt) mma_tf32(acc[mt][nt], a[mt], b[nt]);
      }
      store_grad<DP>(acc, inv, out, row0, 3 * D, h * dh, L, dh, round_out, pk, g, t);
    }
    {  // dK[k, d] = sum_q dS[q, k] Q[q, d] / sqrt(dh)   (A operand = dS^T)
      float acc[2][DP / 8][4] = {};
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t a[2][4], b[DP / 8][2];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) lda_tr(a[mt], Ds, PS, mt * 16, ks * 8, g, t);
#pragma unroll
        for (int nt = 0; nt < DP / 8; ++nt) ldb_kn(b[nt], Qs, QS, ks * 8, nt * 8, g, t);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int nt = 0; nt < DP / 8; ++nt) mma_tf32(acc[mt][nt], a[mt], b[nt]);
      }
      store_grad<DP>(acc, inv, out, row0, 3 * D, D + h * dh, L, dh, round_out, pk, g, t);
    }
    __syncwarp();
  }
}

__global__ void zero_packed_tail_kernel2(PackedOut pk, long R, int tiles_n) {
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
int cfg(Kern kern, size_t smem, long total, int* grid) {
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    set_error("attention_mma: smem %zu: %s", smem, cudaGetErrorString(e));
    return EBK_ERR_CUDA;
  }
  long blocks = (total + WARPS - 1) / WARPS;
  long cap = 148L * 8;
  *grid = (int)(blocks < cap ? blocks : cap);
  return EBK_OK;
}

}  // namespace

bool attention_mma_supported(int L, int dh, const void* p0, const void* p1, const void* p2) {
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  return L >= 1 && L <= 32 && dh >= 4 && dh <= 32 && (dh % 4 == 0) && al(p0) && al(p1) && al(p2);
}

int attention_core_fwd_mma(int n_seq, int L, int nh, int dh, const float* qkv, float* y, cudaStream_t st,
                           Dropout drop_out, bool round_out) {
  if (n_seq <= 0) return EBK_OK;
  EBK_CHECK_ARG(attention_mma_supported(L, dh, qkv, y, qkv) && (nh * dh) % 4 == 0, "attention_mma: unsupported shape L=%d dh=%d", L, dh);
  const long total = (long)n_seq * nh;
  int grid;
#define RUN(DP_)                                                                                              \
  {                                                                                                           \
    constexpr int QKF = (2 * LP * Str<DP_>::QS > LP * Str<DP_>::PS) ? 2 * LP * Str<DP_>::QS : LP * Str<DP_>::PS; \
    size_t smem = (size_t)WARPS * (QKF + LP * Str<DP_>::VS) * sizeof(float);                                  \
    EBK_TRY(cfg(attn_fwd_mma_kernel<DP_>, smem, total, &grid));                                               \
    attn_fwd_mma_kernel<DP_><<<grid, WARPS * 32, smem, st>>>(n_seq, L, nh, dh, qkv, y, drop_out, round_out);  \
  }
  if (dh <= 16) RUN(16) else if (dh <= 24) RUN(24) else RUN(32)
#undef RUN
  EBK_LAUNCH_CHECK();
  return EBK_OK;
}

int attention_core_bwd_mma(int n_seq, int L, int nh, int dh, const float* qkv, const float* dy, Dropout drop,
                           float* dqkv, bool round_out, cudaStream_t st, float* dqkv_packed, int packed_bn) {
  if (n_seq <= 0) return EBK_OK;
  EBK_CHECK_ARG(attention_mma_supported(L, dh, qkv, dy, dqkv) && (nh * dh) % 4 == 0, "attention_mma: unsupported shape L=%d dh=%d", L, dh);
  const long total = (long)n_seq * nh;
  const long R = (long)n_seq * L;
  PackedOut pk{dqkv_packed, packed_bn, packed_bn / 32, 0, 0};
  if (dqkv_packed) {
    pk.ksteps = (int)((R + 31) / 32);
    pk.block_floats = packed_bn * 32;
    const int tiles_n = (3 * nh * dh + packed_bn - 1) / packed_bn;
    const long ntail = (long)pk.ksteps * 32 - R;
    if (ntail > 0) {
      const long tot = ntail * tiles_n * packed_bn;
      zero_packed_tail_kernel2<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(pk, R, tiles_n);
      EBK_LAUNCH_CHECK();
    }
  }
  int grid;
#define RUN(DP_)                                                                                                  \
  {                                                                                                               \
    size_t smem = (size_t)WARPS * (4 * LP * Str<DP_>::QS + 2 * LP * Str<DP_>::PS) * sizeof(float);              \
    EBK_TRY(cfg(attn_bwd_mma_kernel<DP_>, smem, total, &grid));                                                   \
    attn_bwd_mma_kernel<DP_><<<grid, WARPS * 32, smem, st>>>(n_seq, L, nh, dh, qkv, dy, drop, dqkv, round_out, pk); \
  }
  if (dh <= 16) RUN(16) else if (dh <= 24) RUN(24) else RUN(32)
#undef RUN
  EBK_LAUNCH_CHECK();
  return EBK_OK;
}

}  // namespace ebk

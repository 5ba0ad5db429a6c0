// Warp-level building blocks of the per-(sequence, head) attention products on mma.sync.m16n8k8 tf32
// (layers.py:231-252, adjoint_a=True): shared by the stand-alone attention kernels (attention_pre.cu) and by the
// fused QKV-projection + attention epilogue of the tcgen05 GEMM (gemm_tma_sm100.cu).
//   * 32x32xDH products; the softmax / dS algebra stays in the accumulator registers;
//   * operands are [32][ST] shared-memory tiles (rows = tokens, zero rows past L) holding tf32 values.
// Shared-memory row stride ST: DH (=20) or DH+4, chosen so that the 8 rows x 4 columns touched by one
// fragment load fall into 32 distinct banks.
#pragma once
#include "ebk_common.cuh"

namespace ebk {
namespace att {

constexpr int LP = 32;   // padded sequence length
constexpr int PS = 40;   // stride of the 32x32 score-shaped tile

template <int DH> struct Cfg {
  static constexpr int ST = (DH == 20) ? 20 : DH + 4;
  static constexpr int KF = DH / 8;            // full k-steps over the head dim
  static constexpr bool KH = (DH % 8) != 0;    // plus one half k-step (4 columns)
  static constexpr int NT = (DH + 7) / 8;      // n-tiles over the head dim
  static constexpr int MAT = LP * ST;          // floats per staged matrix
};

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ uint32_t u(float x) { return __float_as_uint(x); }
__device__ __forceinline__ uint32_t ur(float x) { return (__float_as_uint(x) + 0x1000u) & 0xFFFFE000u; }  // tf32 round
// 16-byte async copy; bytes == 0 writes zeros (the source is not read)
__device__ __forceinline__ void cp16(float* dst, const float* src, uint32_t bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// issue the cp.async copies of NM [L, DH] slices (global row strides ld[m]) into dst + m * MAT; every
// 16-byte chunk of the [32][ST] tiles is written (rows >= L and columns >= DH with zeros)
template <int DH, int NM>
__device__ __forceinline__ void stage_async(float* dst, const float* const (&src)[NM], const long (&ld)[NM], int L, int lane) {
  constexpr int ST = Cfg<DH>::ST, CPR = ST / 4, MAT = Cfg<DH>::MAT;
#pragma unroll
  for (int it = 0; it < LP * CPR / 32; ++it) {
    const int i = lane + it * 32;
    const int t = i / CPR, j = i - t * CPR;
    const bool ok = t < L && j * 4 < DH;
#pragma unroll
    for (int m = 0; m < NM; ++m)
      cp16(dst + m * MAT + t * ST + j * 4, ok ? src[m] + (long)t * ld[m] + j * 4 : src[m], ok ? 16u : 0u);
  }
}

// acc (32x32 fragments) = X Y^T over the head dim; X, Y staged row-major with stride ST
template <int DH>
__device__ __forceinline__ void gemm_xyT(float (&acc)[2][4][4], const float* X, const float* Y, int g, int t) {
  constexpr int ST = Cfg<DH>::ST;
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[mt][nt][e] = 0.0f;
#pragma unroll
  for (int ks = 0; ks < Cfg<DH>::KF + (Cfg<DH>::KH ? 1 : 0); ++ks) {
    const bool half = ks >= Cfg<DH>::KF;  // compile-time after unrolling: only columns ks*8 .. ks*8+3 exist
    uint32_t a[2][4], b[4][2];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      const float* p = X + (mt * 16 + g) * ST + ks * 8 + t;
      a[mt][0] = u(p[0]);
      a[mt][1] = u(p[8 * ST]);
      a[mt][2] = half ? 0u : u(p[4]);
      a[mt][3] = half ? 0u : u(p[8 * ST + 4]);
    }
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const float* p = Y + (nt * 8 + g) * ST + ks * 8 + t;
      b[nt][0] = u(p[0]);
      b[nt][1] = half ? 0u : u(p[4]);
    }
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) mma_tf32(acc[mt][nt], a[mt], b[nt]);
  }
}

// out[32 x DH] (fragments o[mt][nt]) = P M, P given as accumulator fragments (rounded here), M staged
// [key][d] with stride ST.  k-slot permutation: slot t <-> key 8ks+2t, slot t+4 <-> key 8ks+2t+1.
template <int DH>
__device__ __forceinline__ void gemm_regP(float (&o)[2][Cfg<DH>::NT][4], const float (&P)[2][4][4], const float* M, int g,
                                          int t) {
  constexpr int ST = Cfg<DH>::ST, NT = Cfg<DH>::NT;
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) o[mt][nt][e] = 0.0f;
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    uint32_t a[2][4], b[NT][2];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      a[mt][0] = ur(P[mt][ks][0]);
      a[mt][1] = ur(P[mt][ks][2]);
      a[mt][2] = ur(P[mt][ks][1]);
      a[mt][3] = ur(P[mt][ks][3]);
    }
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const float* p = M + (ks * 8 + 2 * t) * ST + nt * 8 + g;
      b[nt][0] = u(p[0]);
      b[nt][1] = u(p[ST]);
    }
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) mma_tf32(o[mt][nt], a[mt], b[nt]);
  }
}

// out[32 x DH] = T^T M, T a 32x32 tile in shared memory (stride PS, already tf32), M staged [q][d] (stride ST)
template <int DH>
__device__ __forceinline__ void gemm_smemT(float (&o)[2][Cfg<DH>::NT][4], const float* T, const float* M, int g, int t) {
  constexpr int ST = Cfg<DH>::ST, NT = Cfg<DH>::NT;
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) o[mt][nt][e] = 0.0f;
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    uint32_t a[2][4], b[NT][2];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      const float* p = T + (ks * 8 + t) * PS + mt * 16 + g;   // A(row, col) = T[col][row]
      a[mt][0] = u(p[0]);
      a[mt][1] = u(p[8]);
      a[mt][2] = u(p[4 * PS]);
      a[mt][3] = u(p[4 * PS + 8]);
    }
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const float* p = M + (ks * 8 + t) * ST + nt * 8 + g;
      b[nt][0] = u(p[0]);
      b[nt][1] = u(p[4 * ST]);
    }
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) mma_tf32(o[mt][nt], a[mt], b[nt]);
  }
}

// in-register row softmax of S*inv over the first L columns (fragment layout); masked columns -> 0
__device__ __forceinline__ void softmax_rows(float (&acc)[2][4][4], float inv, int L, int t) {
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
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

// store a 32x32 accumulator-fragment matrix to smem [32][PS], rounded to tf32
__device__ __forceinline__ void store_frag(float* P, const float (&acc)[2][4][4], int g, int t) {
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const int col = nt * 8 + 2 * t;
      *reinterpret_cast<uint2*>(P + (mt * 16 + g) * PS + col) = make_uint2(ur(acc[mt][nt][0]), ur(acc[mt][nt][1]));
      *reinterpret_cast<uint2*>(P + (mt * 16 + g + 8) * PS + col) = make_uint2(ur(acc[mt][nt][2]), ur(acc[mt][nt][3]));
    }
}

// write a [32 x DH] fragment matrix (rows = tokens) to out[(row0 + r) * ld + col0 + c], rounded to tf32
template <int DH>
__device__ __forceinline__ void store_rows(const float (&acc)[2][Cfg<DH>::NT][4], float scale, float* __restrict__ out,
                                           long row0, int ld, int col0, int L, int g, int t) {
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < Cfg<DH>::NT; ++nt) {
      const int col = nt * 8 + 2 * t;
      if (col >= DH) continue;
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const int r = mt * 16 + g + hf * 8;
        if (r >= L) continue;
        const uint2 v = make_uint2(ur(acc[mt][nt][hf * 2] * scale), ur(acc[mt][nt][hf * 2 + 1] * scale));
        *reinterpret_cast<uint2*>(out + (row0 + r) * ld + col0 + col) = v;
      }
    }
}


// ---- half-unit variants: ONE 16-row m-tile `mt` of the 32x32 products, so that two warps can share a
// (sequence, head) unit (fused attention epilogue of gemm_tma_sm100.cu: halves the dependent chain per warp) ----
template <int DH>
__device__ __forceinline__ void gemm_xyT_half(float (&acc)[4][4], const float* X, const float* Y, int mt, int g, int t) {
  constexpr int ST = Cfg<DH>::ST;
#pragma unroll
  for (int nt = 0; nt < 4; ++nt)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[nt][e] = 0.0f;
#pragma unroll
  for (int ks = 0; ks < Cfg<DH>::KF + (Cfg<DH>::KH ? 1 : 0); ++ks) {
    const bool half = ks >= Cfg<DH>::KF;
    uint32_t a[4], b[4][2];
    const float* p = X + (mt * 16 + g) * ST + ks * 8 + t;
    a[0] = u(p[0]);
    a[1] = u(p[8 * ST]);
    a[2] = half ? 0u : u(p[4]);
    a[3] = half ? 0u : u(p[8 * ST + 4]);
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const float* q = Y + (nt * 8 + g) * ST + ks * 8 + t;
      b[nt][0] = u(q[0]);
      b[nt][1] = half ? 0u : u(q[4]);
    }
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) mma_tf32(acc[nt], a, b[nt]);
  }
}

__device__ __forceinline__ void softmax_rows_half(float (&acc)[4][4], float inv, int L, int t) {
#pragma unroll
  for (int hf = 0; hf < 2; ++hf) {
    float mx = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int col = nt * 8 + 2 * t + e;
        float s = acc[nt][hf * 2 + e] * inv;
        s = col < L ? s : -INFINITY;
        acc[nt][hf * 2 + e] = s;
        mx = fmaxf(mx, s);
      }
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
    float sum = 0.0f;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const float ex = __expf(acc[nt][hf * 2 + e] - mx);
        acc[nt][hf * 2 + e] = ex;
        sum += ex;
      }
    sum += __shfl_xor_sync(0xffffffffu, sum, 1);
    sum += __shfl_xor_sync(0xffffffffu, sum, 2);
    const float r = 1.0f / sum;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) acc[nt][hf * 2 + e] *= r;
  }
}

__device__ __forceinline__ void store_frag_half(float* P, const float (&acc)[4][4], int mt, int g, int t) {
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    const int col = nt * 8 + 2 * t;
    *reinterpret_cast<uint2*>(P + (mt * 16 + g) * PS + col) = make_uint2(ur(acc[nt][0]), ur(acc[nt][1]));
    *reinterpret_cast<uint2*>(P + (mt * 16 + g + 8) * PS + col) = make_uint2(ur(acc[nt][2]), ur(acc[nt][3]));
  }
}

// rows mt*16 .. mt*16+15 of out[32 x DH] = T^T M
template <int DH>
__device__ __forceinline__ void gemm_smemT_half(float (&o)[Cfg<DH>::NT][4], const float* T, const float* M, int mt, int g, int t) {
  constexpr int ST = Cfg<DH>::ST, NT = Cfg<DH>::NT;
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int e = 0; e < 4; ++e) o[nt][e] = 0.0f;
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    uint32_t a[4], b[NT][2];
    const float* p = T + (ks * 8 + t) * PS + mt * 16 + g;   // A(row, col) = T[col][row]
    a[0] = u(p[0]);
    a[1] = u(p[8]);
    a[2] = u(p[4 * PS]);
    a[3] = u(p[4 * PS + 8]);
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const float* q = M + (ks * 8 + t) * ST + nt * 8 + g;
      b[nt][0] = u(q[0]);
      b[nt][1] = u(q[4 * ST]);
    }
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) mma_tf32(o[nt], a, b[nt]);
  }
}

}  // namespace att
}  // namespace ebk

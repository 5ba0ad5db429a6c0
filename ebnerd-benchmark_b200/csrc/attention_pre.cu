// Attention core of the all-TMA training path (layers.py:231-252, adjoint_a=True) and its backward, for
// operands that are ALREADY tf32 values: the QKV projection and the AttLayer2 dgrad GEMM round (and, for dY,
// dropout-mask) their outputs in their epilogues, so this file only moves data and multiplies:
//   * one warp per (sequence, head); Q, K, V (and dO) go global -> shared with 16-byte cp.async straight
//     into the fragment-friendly layout (zero-filled padding included, so the buffers can be re-used as
//     scratch); STAGES = 2 keeps the loads of the warp's next item in flight while the current one is
//     computed, STAGES = 1 spends the shared memory on more resident warps instead;
//   * 32x32xDH products on mma.sync.m16n8k8 tf32; the softmax / dS algebra stays in the accumulator
//     registers;
//   * A and dS feed the next products (dV = A dO, dQ = dS K) DIRECTLY from the accumulator registers:
//     the MMA's k index is permuted (slot t <-> key 2t, slot t+4 <-> key 2t+1) so that each thread's
//     accumulator pair is exactly its A-fragment; only the transposed uses (O = A^T V, dK = dS^T Q) go
//     through a 32x40 shared tile.
// Shared-memory row stride ST: DH (=20) or DH+4, chosen so that the 8 rows x 4 columns touched by one
// fragment load fall into 32 distinct banks.
#include <stdlib.h>

#include "ebk_common.cuh"
#include "attention_tiles.cuh"

namespace ebk {
namespace {

constexpr int WARPS = 4;
using namespace att;

template <int DH, int STAGES>
__global__ void __launch_bounds__(WARPS * 32, 5) attn_fwd_pre_kernel(int n_seq, int L, int nh, const float* __restrict__ qkv,
                                                                  float* __restrict__ y, Dropout drop) {
  extern __shared__ __align__(16) float smem[];
  constexpr int MAT = Cfg<DH>::MAT;
  constexpr int PER_WARP = STAGES * 3 * MAT;
  static_assert(2 * MAT >= LP * PS, "the score tile overlays Q and K");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  float* base_s = smem + warp * PER_WARP;
  const int D = nh * DH;
  const float inv = rsqrtf((float)DH);
  const long total = (long)n_seq * nh;
  const long stride = (long)gridDim.x * WARPS;
  long item = (long)blockIdx.x * WARPS + warp;
  auto issue = [&](long it, int buf) {
    const int n = (int)(it / nh), h = (int)(it - (long)n * nh);
    const float* b = qkv + (long)n * L * 3 * D + h * DH;
    const float* const src[3] = {b, b + D, b + 2 * D};
    const long ld[3] = {3L * D, 3L * D, 3L * D};
    stage_async<DH, 3>(base_s + buf * 3 * MAT, src, ld, L, lane);
    cp_commit();
  };
  int buf = 0;
  if (STAGES == 2 && item < total) issue(item, 0);
  for (; item < total; item += stride) {
    if (STAGES == 2) {
      const long next = item + stride;
      if (next < total) {
        issue(next, buf ^ 1);
        cp_wait<1>();
      } else {
        cp_wait<0>();
      }
    } else {
      issue(item, 0);
      cp_wait<0>();
    }
    __syncwarp();
    const int n = (int)(item / nh), h = (int)(item - (long)n * nh);
    float* Qs = base_s + buf * 3 * MAT;
    const float* Ks = Qs + MAT;
    const float* Vs = Ks + MAT;
    float* Ps = Qs;  // overlays Q and K once the scores are in registers
    float acc[2][4][4];
    gemm_xyT<DH>(acc, Qs, Ks, g, t);
    softmax_rows(acc, inv, L, t);
    __syncwarp();  // everyone is done reading Q / K
    store_frag(Ps, acc, g, t);
    __syncwarp();
    float o[2][Cfg<DH>::NT][4];
    gemm_smemT<DH>(o, Ps, Vs, g, t);   // O[k, d] = sum_q P[q, k] V[q, d]
    float* out = y + (long)n * L * D + h * DH;
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
          float2 v = make_float2(o[mt][nt][hf * 2], o[mt][nt][hf * 2 + 1]);
          if (drop.on()) {  // AttLayer2 only ever reads dropout(y): store it masked and scaled
            const uint64_t idx = (uint64_t)((long)n * L + r) * (uint64_t)D + (uint64_t)(h * DH + col);
            const float4 f = drop.factor4_group(idx >> 2);
            v.x *= (idx & 2ull) ? f.z : f.x;
            v.y *= (idx & 2ull) ? f.w : f.y;
          }
          *reinterpret_cast<uint2*>(out + (long)r * D + col) = make_uint2(ur(v.x), ur(v.y));
        }
      }
    __syncwarp();  // everyone is done with this buffer before it is overwritten
    if (STAGES == 2) buf ^= 1;
  }
}

template <int DH, int STAGES, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB) attn_bwd_pre_kernel(int n_seq, int L, int nh, const float* __restrict__ qkv,
                                                                  const float* __restrict__ dy, float* __restrict__ dqkv,
                                                                  int tiled) {
  extern __shared__ __align__(16) float smem[];
  constexpr int MAT = Cfg<DH>::MAT;
  constexpr int PER_WARP = STAGES * 4 * MAT;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  float* base_s = smem + warp * PER_WARP;
  const int D = nh * DH;
  const float inv = rsqrtf((float)DH);
  const long total = (long)n_seq * nh;
  const long stride = (long)gridDim.x * WARPS;
  long item = (long)blockIdx.x * WARPS + warp;
  auto issue = [&](long it, int buf) {
    const int n = (int)(it / nh), h = (int)(it - (long)n * nh);
    // rows of the [n_seq*L, 3D] projection buffer, or the zero-padded [32][ST] tiles saved by the fused forward
    const float* b = tiled ? qkv + (long)it * 3 * MAT : qkv + (long)n * L * 3 * D + h * DH;
    const long qs = tiled ? (long)Cfg<DH>::ST : 3L * D, step = tiled ? (long)MAT : (long)D;
    const float* const src[4] = {b, b + step, b + 2 * step, dy + (long)n * L * D + h * DH};
    const long ld[4] = {qs, qs, qs, (long)D};
    stage_async<DH, 4>(base_s + buf * 4 * MAT, src, ld, L, lane);
    cp_commit();
  };
  int buf = 0;
  if (STAGES == 2 && item < total) issue(item, 0);
  for (; item < total; item += stride) {
    if (STAGES == 2) {
      const long next = item + stride;
      if (next < total) {
        issue(next, buf ^ 1);
        cp_wait<1>();
      } else {
        cp_wait<0>();
      }
    } else {
      issue(item, 0);
      cp_wait<0>();
    }
    __syncwarp();
    const int n = (int)(item / nh), h = (int)(item - (long)n * nh);
    const long row0 = (long)n * L;
    const float* Qs = base_s + buf * 4 * MAT;
    const float* Ks = Qs + MAT;
    float* Vs = base_s + buf * 4 * MAT + 2 * MAT;
    const float* Gs = Vs + MAT;  // dO
    float* Ds = Vs;              // dS overlays V and dO once dV is done
    float a_acc[2][4][4], d_acc[2][4][4];
    float o[2][Cfg<DH>::NT][4];
    gemm_xyT<DH>(a_acc, Qs, Ks, g, t);   // A = softmax(Q K^T / sqrt(dh))
    softmax_rows(a_acc, inv, L, t);
    // dV first: afterwards A is only needed for the dS algebra, so A, dA and an output tile are never all live
    gemm_regP<DH>(o, a_acc, Gs, g, t);   // dV[q, d] = sum_k A[q, k] dO[k, d]
    store_rows<DH>(o, 1.0f, dqkv, row0, 3 * D, 2 * D + h * DH, L, g, t);
    gemm_xyT<DH>(d_acc, Vs, Gs, g, t);   // dA = V dO^T
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
    __syncwarp();                        // everyone is done reading V and dO
    store_frag(Ds, d_acc, g, t);         // only the transposed use (dK) needs dS in shared memory
    gemm_regP<DH>(o, d_acc, Ks, g, t);   // dQ[q, d] = sum_k dS[q, k] K[k, d] / sqrt(dh)
    store_rows<DH>(o, inv, dqkv, row0, 3 * D, h * DH, L, g, t);
    __syncwarp();
    gemm_smemT<DH>(o, Ds, Qs, g, t);     // dK[k, d] = sum_q dS[q, k] Q[q, d] / sqrt(dh)
    store_rows<DH>(o, inv, dqkv, row0, 3 * D, D + h * DH, L, g, t);
    __syncwarp();
    if (STAGES == 2) buf ^= 1;
  }
}

template <typename Kern>
int cfg(Kern kern, size_t smem, long total, int* grid) {
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    set_error("attention_pre: smem %zu: %s", smem, cudaGetErrorString(e));
    return EBK_ERR_CUDA;
  }
  int dev = 0, sms = 0, occ = 0;
  EBK_CUDA(cudaGetDevice(&dev));
  EBK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  EBK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, WARPS * 32, smem));
  const long blocks = (total + WARPS - 1) / WARPS;
  const long cap = (long)sms * (occ > 0 ? occ : 1);  // persistent: every resident warp walks its items
  *grid = (int)(blocks < cap ? blocks : cap);
  return EBK_OK;
}

int att_minb() {   // backward kernel: CTAs per SM the register allocation targets (3: no spills, 4: 128 registers)
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("EBK_ATT_MINB");
    v = (e && e[0] == '4') ? 4 : 3;
  }
  return v;
}

int att_stages() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("EBK_ATT_STAGES");
    v = (e && e[0] == '2') ? 2 : 1;
  }
  return v;
}

}  // namespace

bool attention_pre_supported(int L, int dh, const void* p0, const void* p1, const void* p2) {
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  return L >= 1 && L <= 32 && (dh == 16 || dh == 20 || dh == 24 || dh == 32) && al(p0) && al(p1) && al(p2);
}

#define EBK_ATT_DISPATCH(RUN) \
  if (dh == 16) RUN(16) else if (dh == 20) RUN(20) else if (dh == 24) RUN(24) else RUN(32)

int attention_core_fwd_pre(int n_seq, int L, int nh, int dh, const float* qkv, float* y, Dropout drop_out, cudaStream_t st) {
  if (n_seq <= 0) return EBK_OK;
  EBK_CHECK_ARG(attention_pre_supported(L, dh, qkv, y, qkv), "attention_pre: unsupported shape L=%d dh=%d", L, dh);
  const long total = (long)n_seq * nh;
  int grid;
#define RUN2(DH_, ST_)                                                                            \
  {                                                                                               \
    const size_t smem = (size_t)WARPS * (ST_ * 3 * Cfg<DH_>::MAT) * sizeof(float) + 64; /* n-tile overhang */ \
    EBK_TRY(cfg(attn_fwd_pre_kernel<DH_, ST_>, smem, total, &grid));                              \
    attn_fwd_pre_kernel<DH_, ST_><<<grid, WARPS * 32, smem, st>>>(n_seq, L, nh, qkv, y, drop_out); \
  }
#define RUN(DH_) { if (att_stages() == 2) RUN2(DH_, 2) else RUN2(DH_, 1) }
  EBK_ATT_DISPATCH(RUN)
#undef RUN2
#undef RUN
  EBK_LAUNCH_CHECK();
  return EBK_OK;
}

int attention_core_bwd_pre(int n_seq, int L, int nh, int dh, const float* qkv, const float* dy, float* dqkv,
                           cudaStream_t st, bool tiled) {
  if (n_seq <= 0) return EBK_OK;
  EBK_CHECK_ARG(attention_pre_supported(L, dh, qkv, dy, dqkv), "attention_pre: unsupported shape L=%d dh=%d", L, dh);
  const long total = (long)n_seq * nh;
  int grid;
#define RUN2(DH_, ST_)                                                                            \
  {                                                                                               \
    const size_t smem = (size_t)WARPS * (ST_ * 4 * Cfg<DH_>::MAT) * sizeof(float) + 64; /* n-tile overhang */ \
    if (att_minb() == 4) {                                                                        \
      EBK_TRY(cfg(attn_bwd_pre_kernel<DH_, ST_, 4>, smem, total, &grid));                         \
      attn_bwd_pre_kernel<DH_, ST_, 4><<<grid, WARPS * 32, smem, st>>>(n_seq, L, nh, qkv, dy, dqkv, tiled ? 1 : 0); \
    } else {                                                                                      \
      EBK_TRY(cfg(attn_bwd_pre_kernel<DH_, ST_, 3>, smem, total, &grid));                         \
      attn_bwd_pre_kernel<DH_, ST_, 3><<<grid, WARPS * 32, smem, st>>>(n_seq, L, nh, qkv, dy, dqkv, tiled ? 1 : 0); \
    }                                                                                             \
  }
#define RUN(DH_) { if (att_stages() == 2) RUN2(DH_, 2) else RUN2(DH_, 1) }
  EBK_ATT_DISPATCH(RUN)
#undef RUN2
#undef RUN
  EBK_LAUNCH_CHECK();
  return EBK_OK;
}

}  // namespace ebk

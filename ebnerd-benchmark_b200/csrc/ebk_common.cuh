// Shared helpers for the ebk kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/ebk.h"

namespace ebk {

// ---- error plumbing -------------------------------------------------------------------
void set_error(const char* fmt, ...);

#define EBK_CHECK_ARG(cond, ...)            \
  do {                                      \
    if (!(cond)) {                          \
      ::ebk::set_error(__VA_ARGS__);        \
      return EBK_ERR_INVALID;               \
    }                                       \
  } while (0)

#define EBK_CUDA(call)                                                                     \
  do {                                                                                     \
    cudaError_t e__ = (call);                                                              \
    if (e__ != cudaSuccess) {                                                              \
      ::ebk::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      return EBK_ERR_CUDA;                                                                 \
    }                                                                                      \
  } while (0)

// every kernel launch of the library goes through this macro -> also counts launches
#define EBK_LAUNCH_CHECK()           \
  do {                               \
    ::ebk::note_launch();            \
    EBK_CUDA(cudaGetLastError());    \
  } while (0)

#define EBK_TRY(call)              \
  do {                             \
    int s__ = (call);              \
    if (s__ != EBK_OK) return s__; \
  } while (0)

void note_launch();

// ---- optional per-kernel profiling (CUDA events on the launching stream) ------------------
enum ProfTag {
  T_QKV_FWD = 0, T_ATTN_FWD, T_ATT_GEMM_FWD, T_POOL_FWD, T_POOL_BWD, T_COLSUM, T_ATT_WGRAD, T_ATT_DGRAD,
  T_ATTN_BWD, T_QKV_WGRAD, T_QKV_DGRAD, T_SCATTER, T_SCORE, T_ADAM,
  T_EMBED_PAD, T_CONV_FWD, T_CONV_DZ, T_CONV_WGRAD, T_CONV_DGRAD, T_CATVIEW, T_EMBED_GATHER, T_NUM_TAGS
};
// profiler group of the following launches: 0 = "news." slots, 1 = "user." slots
void prof_set_group(int g);
bool prof_on();
void prof_begin(int tag, cudaStream_t st);
void prof_end(int tag, cudaStream_t st);
// run `call` (an int-returning launcher using stream `st`), timing it when profiling is enabled
#define EBK_PROF(tag, call)                         \
  do {                                              \
    if (::ebk::prof_on()) ::ebk::prof_begin(tag, st); \
    int s__ = (call);                               \
    if (::ebk::prof_on()) ::ebk::prof_end(tag, st); \
    if (s__ != EBK_OK) return s__;                  \
  } while (0)

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- counter-based dropout (mirrors oracle/nrms_oracle.py::dropout_keep_mask) ----------
// group g = idx >> 2 draws two 32-bit words (x, y) from (seed, g); lane j = idx & 3 uses
// x[0:16], x[16:32], y[0:16], y[16:32]; element kept iff bits >= thr, thr = floor(p*65536 + 0.5).
// Two 32-bit avalanche hashes (lowbias32-style) give the 4 x 16 random bits of a group.
__host__ __device__ __forceinline__ void dropout_group_bits(uint64_t seed, uint64_t group, uint32_t& x, uint32_t& y) {
  const uint32_t s0 = (uint32_t)seed, s1 = (uint32_t)(seed >> 32);
  x = (uint32_t)group * 0x9E3779B1u + (uint32_t)(group >> 32) * 0x85EBCA77u + s0;
  x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16;
  y = (x ^ s1) * 0x9E3779B1u;
  y ^= y >> 15; y *= 0x2C1B3C6Du; y ^= y >> 12; y *= 0x297A2D39u; y ^= y >> 15;
}
__host__ __device__ __forceinline__ uint32_t dropout_threshold(float p) {
  return (uint32_t)floorf(p * 65536.0f + 0.5f);
}
// keep flag for a single element
__host__ __device__ __forceinline__ bool dropout_keep(uint64_t seed, uint64_t idx, uint32_t thr) {
  uint32_t x, y;
  dropout_group_bits(seed, idx >> 2, x, y);
  const uint32_t lane = (uint32_t)(idx & 3ull);
  const uint32_t w = (lane & 2u) ? y : x;
  return ((lane & 1u) ? (w >> 16) : (w & 0xFFFFu)) >= thr;
}

struct Dropout {
  uint64_t seed;
  uint32_t thr;   // 0 => disabled
  float scale;    // 1/(1-p)
  // CUDA-graph replay: the seed of THIS step is read from device memory (ebk_step_params), so that a captured
  // launch does not bake a seed in; NULL => `seed`
  const uint64_t* seed_dev;
  uint64_t seed_add;   // added to *seed_dev (one device-resident step seed serves a stack of layers: seed + layer index)
  __host__ __device__ bool on() const { return thr != 0; }
  __device__ __forceinline__ uint64_t cur_seed() const {
#ifdef __CUDA_ARCH__
    return seed_dev != nullptr ? __ldg(reinterpret_cast<const unsigned long long*>(seed_dev)) + seed_add : seed;
#else
    return seed;
#endif
  }
  // scale factor (0 or 1/(1-p)) of element idx
  __device__ __forceinline__ float factor(uint64_t idx) const {
    return dropout_keep(cur_seed(), idx, thr) ? scale : 0.0f;
  }
  // factors of the 4 elements of group g (elements 4g .. 4g+3)
  __device__ __forceinline__ float4 factor4_group(uint64_t g) const {
    uint32_t x, y;
    dropout_group_bits(cur_seed(), g, x, y);
    float4 f;
    f.x = ((x & 0xFFFFu) >= thr) ? scale : 0.0f;
    f.y = ((x >> 16) >= thr) ? scale : 0.0f;
    f.z = ((y & 0xFFFFu) >= thr) ? scale : 0.0f;
    f.w = ((y >> 16) >= thr) ? scale : 0.0f;
    return f;
  }
  // factors of the 4 elements of an aligned group starting at idx (idx % 4 == 0)
  __device__ __forceinline__ float4 factor4(uint64_t idx) const { return factor4_group(idx >> 2); }
};
static inline Dropout make_dropout(bool training, float p, uint64_t seed, const uint64_t* seed_dev = nullptr) {
  Dropout d;
  d.seed = seed;
  d.seed_dev = seed_dev;
  d.seed_add = 0;
  if (training && p > 0.0f) {
    d.thr = dropout_threshold(p);
    d.scale = 1.0f / (1.0f - p);
  } else {
    d.thr = 0;
    d.scale = 1.0f;
  }
  return d;
}

// round-to-nearest to tf32 by bit arithmetic (the tensor core truncates the low 13 mantissa bits)
__host__ __device__ __forceinline__ float round_tf32_bits(float x) {
#ifdef __CUDA_ARCH__
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
#else
  uint32_t u;
  memcpy(&u, &x, 4);
  u = (u + 0x1000u) & 0xFFFFE000u;
  memcpy(&x, &u, 4);
  return x;
#endif
}

// ---- warp helpers ---------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---- internal launchers shared between translation units ---------------------------------
struct GemmOperandA {
  const float* ptr;       // storage [rows, lda]
  int lda;
  bool trans;             // false: A(m,k)=S(m,k); true: A(m,k)=S(k,m)  (S = storage)
  const int32_t* gather;  // optional: storage row s is read from row gather[s]
  int gather_limit;       // rows outside [0, limit) read as zero
  Dropout drop;           // optional dropout on S(s,c) with element index s*drop_ld + c
  int drop_ld;
};

int gemm_f32(const GemmOperandA& A, const float* B, int ldb, bool transB, float* C, int ldc,
             int M, int N, int K, float beta, cudaStream_t st);
// How the B operand of the tcgen05 GEMM is supplied (see gemm_tf32_sm100.cu)
enum GemmBMode {
  GEMM_B_PACKED = 0,   // pre-rounded + pre-arranged by gemm_tf32_pack_b (weights): one bulk copy per stage
  GEMM_B_ROUNDED = 1,  // row-major, values already tf32-rounded (activations): cp.async
  GEMM_B_RAW = 2       // row-major fp32, rounded in flight through registers
};
// x3: error-compensated 3xTF32 (needs B_lo unless b_mode == GEMM_B_RAW)
int gemm_tf32(const GemmOperandA& A, const float* B, int ldb, bool transB, float* C, int ldc,
              int M, int N, int K, float beta, cudaStream_t st, int b_mode = GEMM_B_RAW, bool x3 = false,
              const float* B_lo = nullptr);
bool gemm_tf32_eligible(const GemmOperandA& A, const float* B, int ldb, int M, int N, int K);
size_t gemm_tf32_packed_floats(int N, int K, bool transB);
int gemm_tf32_bn(int N, int K, bool transB);  // tile width the GEMM will use for this B
int gemm_tf32_pack_b(float* dst_hi, float* dst_lo, const float* B, int ldb, bool transB, int N, int K,
                     cudaStream_t st);
int gemm_dispatch(int math, const GemmOperandA& A, const float* B, int ldb, bool transB, float* C,
                  int ldc, int M, int N, int K, float beta, cudaStream_t st, int b_mode = GEMM_B_RAW,
                  const float* B_lo = nullptr);
// All-TMA tcgen05 GEMM on dense operands that already hold tf32-representable values (gemm_tma_sm100.cu):
// C (+)= alpha * opA(A) . opB(B); beta in {0,1}; tall: 256-row tiles (halves the L2 traffic of B)
// Optional fused epilogue (beta == 0 only), applied in this order to v = alpha * acc:
//   v += rowscale[row] * rowvec[(row / L) * rowvec_ld + col]   (if rowscale)
//   v *= dropout factor of element row * drop_ld + col            (if drop.on())
//   v  = tf32(v)                                                   (if round_out: C feeds another tensor-core op)
struct GemmEpilogue {
  const float* rowscale; const float* rowvec; int rowvec_ld; int L;
  Dropout drop; int drop_ld;
  bool round_out;
};
bool gemm_tma_eligible(const float* A, int lda, const float* B, int ldb, int M, int N, int K);
int gemm_tma(const float* A, int lda, bool transA, const float* B, int ldb, bool transB, float* C, int ldc, int M,
             int N, int K, float beta, float alpha, cudaStream_t st, int tall = 0, const GemmEpilogue* epi = nullptr,
             int cluster = 0);
// Fused SelfAttention forward on the all-TMA path (gemm_tma_sm100.cu): QKV projection with the per-head
// softmax(QK^T/sqrt(dh))^T V computed in the GEMM epilogue (accumulator tile = whole sequences x whole heads).
//   wp = permute_round_wqkv(W): [Din, (head, Q|K|V, d)] tf32 copy of Wqkv [Din, (Q|K|V, head, d)]
//   qkv_t (nullable): Q | K | V saved as [n_seq, nh, 3, 32, ST] zero-padded tiles for attention_core_bwd_pre(tiled)
bool qkv_attn_fused_supported(int L, int dh, int Din, const float* xd, const float* wp, const float* y);
size_t qkv_tiles_floats(int n_seq, int nh, int dh);
int permute_round_wqkv(float* wp, const float* W, int Din, int nh, int dh, cudaStream_t st);
int qkv_attn_fused(const float* xd, int Din, const float* wp, int n_seq, int L, int nh, int dh, float* qkv_t, float* y,
                   Dropout drop, cudaStream_t st);
// hi[i] = tf32(src[i]) (round to nearest), lo[i] = src[i] - hi[i]
int split_tf32_copy(float* hi, float* lo, const float* src, size_t n, cudaStream_t st);
// dst[i] = round-to-nearest-tf32(src[i])  (so that the tensor core's truncation is exact)
int round_tf32_copy(float* dst, const float* src, size_t n, cudaStream_t st);

int attention_core_fwd(int n_seq, int L, int nh, int dh, const float* qkv, float* y, cudaStream_t st);
// round_out: store dqkv rounded to tf32 (it is the cp.async-staged B operand of the wgrad GEMM)
// dqkv_packed (optional): second copy of dqkv in the packed B layout of the wgrad GEMM (tile width packed_bn)
int attention_core_bwd(int n_seq, int L, int nh, int dh, const float* qkv, const float* dy,
                       Dropout drop, float* dqkv, bool round_out, cudaStream_t st, float* dqkv_packed = nullptr,
                       int packed_bn = 0);

// warp-level tensor-core (mma.sync tf32) variants for training, L <= 32 (attention_mma.cu)
bool attention_mma_supported(int L, int dh, const void* p0, const void* p1, const void* p2);
// drop_out / round_out: store dropout(y) (mask and 1/(1-p) scale) and/or round it to tf32 -- what the
// TMA GEMM of the following AttLayer2 consumes
int attention_core_fwd_mma(int n_seq, int L, int nh, int dh, const float* qkv, float* y, cudaStream_t st,
                           Dropout drop_out = Dropout{0, 0, 1.0f, nullptr}, bool round_out = false);
int attention_core_bwd_mma(int n_seq, int L, int nh, int dh, const float* qkv, const float* dy, Dropout drop,
                           float* dqkv, bool round_out, cudaStream_t st, float* dqkv_packed = nullptr,
                           int packed_bn = 0);

// TMA-path attention (attention_pre.cu): qkv / dy already hold tf32 values (dy already dropout-masked);
// cp.async double-buffered staging, register-resident A / dS.  fwd stores tf32(dropout(y)); bwd stores tf32(dqkv).
bool attention_pre_supported(int L, int dh, const void* p0, const void* p1, const void* p2);
int attention_core_fwd_pre(int n_seq, int L, int nh, int dh, const float* qkv, float* y, Dropout drop_out, cudaStream_t st);
// tiled: qkv is the [n_seq, nh, 3, 32, ST] tile layout written by qkv_attn_fused instead of [n_seq*L, 3D] rows
int attention_core_bwd_pre(int n_seq, int L, int nh, int dh, const float* qkv, const float* dy, float* dqkv,
                           cudaStream_t st, bool tiled = false);

// AttLayer2 pieces
int attpool_fwd(int n_seq, int L, int D, int att, const float* y0, Dropout drop, float* hbuf /*[R,att] in: pre-act, out: tanh*/,
                const float* attb, const float* attq, float* w /*[R]*/, float* out /*[n_seq,out_ld]*/, cudaStream_t st,
                int out_ld = 0 /*0 => D*/);
int attpool_bwd(int n_seq, int L, int D, int att, const float* y0, Dropout drop, const float* hbuf,
                const float* attq, const float* w, const float* d_out, float* da /*[R]*/,
                float* dpre /*[R,att]*/, float* dy /*[R,D] = w_t*d_out*/, bool round_dpre, cudaStream_t st,
                int dout_ld = 0 /*0 => D*/);
// TMA-path variant: y0 is already dropout(Y0); dy is NOT written (the dpre.W^T GEMM adds w_t*d_out in its
// epilogue); per-sequence column sums of dpre and of h*da go to colpart [n_seq, 2*att] (dattb | dattq partials)
int attpool_bwd_fused(int n_seq, int L, int D, int att, const float* y0, const float* hbuf, const float* attq,
                      const float* w, const float* d_out, float* da, float* dpre, float* colpart, cudaStream_t st,
                      int dout_ld = 0 /*0 => D*/);
// column sums: out[j] += sum_r coef[r] * X[r,j]   (coef may be NULL => 1); deterministic
// two-stage reduction through `partial` (colsum_partial_floats(R, Ncols) floats of scratch)
int colsum_accum_ws(int R, int Ncols, const float* X, int ldx, const float* coef, float* out, float* partial,
                    cudaStream_t st);
int colsum_accum2_ws(int R, int Ncols, int n0, const float* X, int ldx, float* out0, float* out1, float* partial,
                     cudaStream_t st);
size_t colsum_partial_floats(int R, int Ncols);

// z = act(z + b) in place (n = rows*U elements, U % 4 == 0);  dz = dy * dropout' * relu'(y)
int bias_act(float* z, const float* b, long n, int U, int relu, cudaStream_t st);
int act_bwd(const float* dy, const float* y, long n, Dropout drop, int relu, int round_out, float* dz, cudaStream_t st);

// Xd[r, :] = tf32(dropout(table[tok[r], :]))  (ids outside [0, V) -> zero row); tok == NULL: Xd = tf32(x)
// Data parallel with a rank-sharded table: float index f of the table lives in the buffer of rank f / shard_floats
// (every rank maps every peer's full-size table buffer through CUDA IPC; only the owner's slice is current).
struct PeerTables {
  int world;                 // <= 1: single table
  unsigned long long shard_floats;
  const float* p[8];
};
// Token CSR: the positions r of tok[0..R) grouped by token id (count[t], offset[t], perm[offset[t] .. +count[t])); ids
// outside [0, V) are in no group.  Built by embed_adam.cu (the single-GPU optimizer groups the per-row gradients the
// same way); `total` is the scan's running total, `cursor` the fill cursors.
struct TokenCsr {
  int *count, *cursor, *total, *offset, *perm;
};
size_t token_csr_bytes(int R, int V);
int token_csr_build(int R, int V, const int32_t* tok, void* ws, TokenCsr* out, cudaStream_t st);
// the gather of embed_rows with every DISTINCT token's row read once (tokens with more than 32 positions: per position)
int embed_rows_csr(int R, int E, int V, const int32_t* tok, const float* table, Dropout drop, float* xd, cudaStream_t st,
                   const PeerTables* peers, const TokenCsr& csr);
int embed_rows(int R, int E, int V, const int32_t* tok, const float* table_or_x, Dropout drop, float* xd,
               cudaStream_t st, const PeerTables* peers = nullptr, long row_offset = 0 /* rows before this chunk: dropout index */);
// d_table[tok[r], :] += dX[r, :] * dropout(r*E + e)
int scatter_rows_add(int R, int E, int V, const int32_t* tok, const float* dX, Dropout drop,
                     float* d_table, cudaStream_t st);

}  // namespace ebk

// fp32 CUDA-core GEMM with gathered / dropout-masked / transposed A operand.
// Used for EBK_MATH_FP32 (bit-faithful fp32 FMA path for parity debugging) and for the
// small contractions of the path (user encoder, AttLayer2 at small R).  The large
// projections go through gemm_tf32_sm100.cu when EBK_MATH_TF32 is selected.
#include "ebk_common.cuh"

namespace ebk {

namespace {

constexpr int BM = 64, BN = 64, BK = 16, TM = 4, TN = 4;
constexpr int NTHREADS = (BM / TM) * (BN / TN);  // 256

struct AView {
  const float* ptr;
  int lda;
  const int32_t* gather;
  int gather_limit;
  Dropout drop;
  int drop_ld;
  // storage element (s, c); returns 0 outside [0,S)x[0,Cc)
  __device__ __forceinline__ float load(int s, int c, int S, int Cc) const {
    if (s >= S || c >= Cc) return 0.0f;
    long row = s;
    if (gather) {
      int g = gather[s];
      if (g < 0 || g >= gather_limit) return 0.0f;
      row = g;
    }
    float v = ptr[row * (long)lda + c];
    if (drop.on()) v *= drop.factor((uint64_t)s * (uint64_t)drop_ld + (uint64_t)c);
    return v;
  }
};

// C tile = A[M,K] * B[K,N]; split-K over gridDim.z with atomic accumulation.
template <bool TA, bool TB>
__global__ void __launch_bounds__(NTHREADS) gemm_f32_kernel(AView A, const float* __restrict__ B, int ldb,
                                                            float* __restrict__ C, int ldc, int M, int N,
                                                            int K, int k_per_split, float beta, int atomic) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kbeg = blockIdx.z * k_per_split;
  const int kend = min(K, kbeg + k_per_split);
  const int tx = tid % (BN / TN), ty = tid / (BN / TN);

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.0f;

  for (int k0 = kbeg; k0 < kend; k0 += BK) {
    // ---- stage A tile: As[k][m] ----
    if (!TA) {
      // storage (s=m, c=k): consecutive threads walk k (contiguous)
      for (int i = tid; i < BM * BK; i += NTHREADS) {
        int m = i / BK, k = i % BK;
        int gk = k0 + k;
        As[k][m] = (gk < kend) ? A.load(m0 + m, gk, M, K) : 0.0f;
      }
    } else {
      // storage (s=k, c=m): consecutive threads walk m (contiguous)
      for (int i = tid; i < BM * BK; i += NTHREADS) {
        int k = i / BM, m = i % BM;
        int gk = k0 + k;
        As[k][m] = (gk < kend) ? A.load(gk, m0 + m, K, M) : 0.0f;
      }
    }
    // ---- stage B tile: Bs[k][n] ----
    if (!TB) {
      for (int i = tid; i < BN * BK; i += NTHREADS) {
        int k = i / BN, n = i % BN;
        int gk = k0 + k, gn = n0 + n;
        Bs[k][n] = (gk < kend && gn < N) ? B[(long)gk * ldb + gn] : 0.0f;
      }
    } else {
      for (int i = tid; i < BN * BK; i += NTHREADS) {
        int n = i / BK, k = i % BK;
        int gk = k0 + k, gn = n0 + n;
        Bs[k][n] = (gk < kend && gn < N) ? B[(long)gn * ldb + gk] : 0.0f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * TM]);
      float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * TN]);
      float a[TM] = {a4.x, a4.y, a4.z, a4.w};
      float b[TN] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int m = m0 + ty * TM + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int n = n0 + tx * TN + j;
      if (n >= N) continue;
      float* c = C + (long)m * ldc + n;
      if (atomic) {
        atomicAdd(c, acc[i][j]);
      } else {
        *c = (beta != 0.0f ? *c : 0.0f) + acc[i][j];
      }
    }
  }
}

__global__ void zero_matrix_kernel(float* C, int ldc, int M, int N) {
  long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i < (long)M * N) C[(i / N) * (long)ldc + (i % N)] = 0.0f;
}

}  // namespace

int gemm_f32(const GemmOperandA& A, const float* B, int ldb, bool transB, float* C, int ldc, int M,
             int N, int K, float beta, cudaStream_t st) {
  if (M <= 0 || N <= 0) return EBK_OK;
  EBK_CHECK_ARG(K >= 0 && A.ptr && B && C, "gemm_f32: null operand");
  AView av{A.ptr, A.lda, A.gather, A.gather_limit, A.drop, A.drop_ld};
  dim3 grid(ceil_div(N, BN), ceil_div(M, BM), 1);
  // split-K when the tile grid cannot fill 148 SMs and K is long (weight-gradient shapes)
  int splitk = 1;
  long tiles = (long)grid.x * grid.y;
  if (K >= 4096 && tiles < 2 * 148) {
    splitk = (int)min((long)ceil_div(K, 1024), (long)max(1L, (4L * 148) / tiles));
  }
  int k_per_split = ceil_div(ceil_div(K, splitk), BK) * BK;
  splitk = ceil_div(K, k_per_split);
  if (splitk < 1) splitk = 1;
  grid.z = splitk;
  int atomic = splitk > 1;
  if (atomic && beta == 0.0f) {
    long n = (long)M * N;
    zero_matrix_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(C, ldc, M, N);
    EBK_LAUNCH_CHECK();
  }
#define LAUNCH(TA_, TB_) \
  gemm_f32_kernel<TA_, TB_><<<grid, NTHREADS, 0, st>>>(av, B, ldb, C, ldc, M, N, K, k_per_split, beta, atomic)
  if (!A.trans && !transB) LAUNCH(false, false);
  else if (!A.trans && transB) LAUNCH(false, true);
  else if (A.trans && !transB) LAUNCH(true, false);
  else LAUNCH(true, true);
#undef LAUNCH
  EBK_LAUNCH_CHECK();
  return EBK_OK;
}

}  // namespace ebk

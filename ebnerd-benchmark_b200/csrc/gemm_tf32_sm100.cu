// tcgen05 (5th-gen tensor core) TF32 GEMM for sm_100a with a gathered / dropout-masked /
// transposed A operand:   C[M,N] (+)= opA(A)[M,K] . opB(B)[K,N],  fp32 in, fp32 accumulate.
//
// Why kind::tf32: the reference computes in fp32 and the parity gate is 1e-3 relative on
// click scores.  tf32 reads the fp32 words exactly as they sit in the embedding table /
// activations (no conversion pass, no extra HBM traffic) and keeps 10 mantissa bits.
//
// Structure (one CTA per 128 x BN output tile, 2 CTAs resident per SM so that one CTA's
// epilogue overlaps the other's main loop):
//   warps 0-3  producers: 16-byte cp.async (LDGSTS) global -> shared with the UMMA
//              SWIZZLE_128B pattern applied by hand (the A rows are *gathered* by token id,
//              which tiled TMA cannot express); optional in-place dropout on A; then
//              fence.proxy.async + mbarrier arrive.  After the main loop the same warps run
//              the epilogue: tcgen05.ld (TMEM -> registers) -> global store / red.add.
//   warp 4     allocates TMEM and issues tcgen05.mma (one elected lane), releasing stages
//              with tcgen05.commit -> mbarrier.
// Shared-memory operand layouts (both are "rows of 128 bytes, 16-byte chunk c stored at
// c ^ (row & 7)", i.e. Swizzle<3,4,3> on 1024-byte aligned atoms):
//   K-major  operand: row = m (or n), 128 B = 32 consecutive k       (SBO = 1024)
//   MN-major operand: row = k, 128 B = 32 consecutive m (or n); SWIZZLE_128B_BASE32B atoms
//                     [4 k][32 mn] laid out [k-group][mn-group]      (LBO = 512, SBO = groups*512)
#include "ebk_common.cuh"

namespace ebk {
namespace {

constexpr int BM = 128;
constexpr int BK = 32;             // fp32 elements per stage along K = one 128-byte swizzle row
constexpr int UMMA_K = 8;          // tf32
constexpr int MAX_STAGES = 6;
constexpr int PRODUCER_THREADS = 128;
constexpr int THREADS = PRODUCER_THREADS + 32;
constexpr int TMEM_COLS = 256;
constexpr uint32_t SPIN_LIMIT = 1u << 26;  // bounded mbarrier spins: a protocol bug traps instead of hanging the GPU

struct Params {
  const float* A; int lda; const int32_t* a_gather; int a_gather_limit; Dropout a_drop; int a_drop_ld;
  const float* B; int ldb;
  const float* B_lo;  // X3 + pre-split B: the low parts (same layout as B)
  float* C; int ldc;
  int M, N, K;
  int BN;             // tile width, multiple of 16, <= 256
  int stages;
  int ksteps_total;   // ceil(K / BK)
  int ksteps_per_split;
  int out_mode;       // 0 store, 1 load-add-store (beta=1), 2 atomic add (split-K)
  float alpha;        // output scale (carries the dropout 1/(1-p) of a masked A operand)
};

// ---- PTX wrappers -------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > SPIN_LIMIT) __trap();
  }
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// UMMA shared-memory descriptor (cute::UMMA::SmemDescriptor bit layout): start>>4 [0,14),
// LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48), layout SWIZZLE_128B=2 [61,64).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                              uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout_type << 61;  // 2 = SWIZZLE_128B, 1 = SWIZZLE_128B_BASE32B
  return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D=f32 [4,6)=1, A,B=tf32 [7,10)=[10,13)=2,
// a_major bit15, b_major bit16 (1 = MN-major), N>>3 [17,23), M>>4 [24,29).
__device__ __forceinline__ uint32_t make_idesc(bool a_mn, bool b_mn, int n) {
  uint32_t d = 0;
  d |= 1u << 4;
  d |= 2u << 7;
  d |= 2u << 10;
  d |= (a_mn ? 1u : 0u) << 15;
  d |= (b_mn ? 1u : 0u) << 16;
  d |= (uint32_t)(n >> 3) << 17;
  d |= (uint32_t)(BM >> 4) << 24;
  return d;
}

// ---- producer helpers ------------------------------------------------------------------------
// Both operand layouts are "rows of 128 bytes" in shared memory:
//   K-major  (SWIZZLE_128B):         row = mn index, 128 B = 32 consecutive k; 16-byte chunk c of row r
//                                    sits at r*128 + ((c ^ (r&7)) << 4)                 (SBO = 1024)
//   MN-major (SWIZZLE_128B_BASE32B): row = k index, 128 B = 32 consecutive mn.  32-bit MN-major
//                                    operands only exist in this layout (cute Layout_MN_SW128_32B_Atom):
//                                    atoms of [4 k-rows][32 mn], 32-byte units XOR-swizzled by the k-row
//                                    (Swizzle<2,5,2> on the byte address), atoms laid out
//                                    [k-group][mn-group]                       (LBO = 512, SBO = groups*512)
__device__ __forceinline__ uint32_t k_off(int r, int c) { return (uint32_t)r * 128u + (uint32_t)((c ^ (r & 7)) << 4); }
__device__ __forceinline__ uint32_t mn_off(int kr, int c, int groups) {
  const uint32_t unit32 = (uint32_t)(((c & 7) >> 1) ^ (kr & 3));
  return (uint32_t)((kr >> 2) * groups + (c >> 3)) * 512u + (uint32_t)(kr & 3) * 128u + (unit32 << 5) +
         (uint32_t)((c & 1) << 4);
}

// Where chunk `i` of this thread lives.  MN=false: tile = `ext` mn-rows x 32 k.  MN=true: tile = 32
// k-rows x `ext` mn (ext multiple of 32).  Returns false when i is past the tile.
template <bool MN>
__device__ __forceinline__ bool chunk_coord(int i, int ptid, int ext, int& row, int& c) {
  if (!MN) {
    row = (ptid >> 3) + i * (PRODUCER_THREADS / 8);
    c = ptid & 7;
    return row < ext;
  } else {
    const int cpr = ext >> 2;
    const int idx = ptid + i * PRODUCER_THREADS;
    row = idx / cpr;
    c = idx - row * cpr;
    return row < BK;
  }
}
// Global source of that chunk: pointer and number of valid floats (0..4).
template <bool MN>
__device__ __forceinline__ const float* chunk_src(int row, int c, const float* base, int ld, const int32_t* gather,
                                                  int gather_limit, int mn0, int mn_end, int k0, int K, int& nvalid) {
  const int mn = MN ? mn0 + c * 4 : mn0 + row;
  const int k = MN ? k0 + row : k0 + c * 4;
  const int srow = MN ? k : mn;             // storage row index (gathered)
  const int scol = MN ? mn : k;             // storage column
  const int row_end = MN ? K : mn_end, col_end = MN ? mn_end : K;
  nvalid = 0;
  if (srow >= row_end) return base;
  int nv = col_end - scol;
  nv = nv < 0 ? 0 : (nv > 4 ? 4 : nv);
  if (nv == 0) return base;
  long r = srow;
  if (gather) {
    const int g = __ldg(gather + srow);
    if (g < 0 || g >= gather_limit) return base;
    r = g;
  }
  nvalid = nv;
  return base + r * (long)ld + scol;
}
__device__ __forceinline__ float4 ldg_chunk(const float* src, int nvalid) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (nvalid == 4) {
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(src));
  } else if (nvalid > 0) {
    v.x = __ldg(src);
    if (nvalid > 1) v.y = __ldg(src + 1);
    if (nvalid > 2) v.z = __ldg(src + 2);
  }
  return v;
}
// round-to-nearest to tf32: the tensor core TRUNCATES the low 13 mantissa bits of an fp32 word,
// so adding half a tf32 ulp to the bit pattern beforehand makes that truncation a rounding.
__device__ __forceinline__ float rn_tf32(float x) { return __uint_as_float(__float_as_uint(x) + 0x1000u); }
// 3xTF32 split: x = hi + lo with hi exactly a tf32 value (round to nearest) and lo = x - hi exact
// in fp32 (|lo| <= 2^-11 |x|); the products hi*hi + lo*hi + hi*lo recover ~fp32 accuracy.
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
  lo = rn_tf32(x - hi);
}
__device__ __forceinline__ void split4(const float4& v, float4& hi, float4& lo) {
  split_tf32(v.x, hi.x, lo.x); split_tf32(v.y, hi.y, lo.y); split_tf32(v.z, hi.z, lo.z); split_tf32(v.w, hi.w, lo.w);
}

// cp.async path (operand already tf32-rounded in memory, no mask)
template <bool MN>
__device__ __forceinline__ void stage_async(uint32_t sdst, const float* base, int ld, int mn0, int mn_end, int ext,
                                            int k0, int K, int ptid) {
  const int groups = ext >> 5;
#pragma unroll 4
  for (int i = 0; i < 16; ++i) {
    int row, c, nvalid;
    if (!chunk_coord<MN>(i, ptid, ext, row, c)) break;
    const float* src = chunk_src<MN>(row, c, base, ld, nullptr, 0, mn0, mn_end, k0, K, nvalid);
    cp_async16(sdst + (MN ? mn_off(row, c, groups) : k_off(row, c)), src, (uint32_t)nvalid * 4u);
  }
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// B_REG: B is staged through registers with in-flight tf32 rounding (generic callers); otherwise B
// must already be tf32-rounded in memory and is staged with cp.async.
// X3: error-compensated 3xTF32 (hi/lo split of both operands, three MMAs per k-chunk) for the
// inference forward, whose click scores must match the fp32 reference to 1e-3.
template <bool A_MN, bool B_MN, bool B_REG, bool X3>
__global__ void __launch_bounds__(THREADS, X3 ? 1 : 2) gemm_tf32_kernel(const Params p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[MAX_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[MAX_STAGES];
  __shared__ __align__(8) uint64_t accum_bar;
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int S = p.stages;
  const int BN = p.BN;
  const int bn_pad = (BN + 31) & ~31;  // MN-major B rows are staged in 32-wide groups
  const int b_ext = B_MN ? bn_pad : BN;
  constexpr uint32_t NSPLIT = X3 ? 2u : 1u;                               // hi [+ lo] tiles per operand
  const uint32_t a_tile = BM * BK * 4;                                    // 16 KB
  const uint32_t b_tile = ((uint32_t)b_ext * BK * 4 + 1023u) & ~1023u;    // <= 32 KB
  const uint32_t a_bytes = NSPLIT * a_tile;
  const uint32_t stage_bytes = a_bytes + NSPLIT * b_tile;
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t smem_base = smem_u32(smem);

  const int m0 = blockIdx.y * BM;
  const int n0 = blockIdx.x * BN;
  const int ks_begin = blockIdx.z * p.ksteps_per_split;
  int ks_end = ks_begin + p.ksteps_per_split;
  if (ks_end > p.ksteps_total) ks_end = p.ksteps_total;
  const int nk = ks_end - ks_begin;  // >= 1 by construction

  if (tid == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(smem_u32(&full_bar[s]), PRODUCER_THREADS);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    mbar_init(smem_u32(&accum_bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)),
                 "n"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (warp < 4) {
    // ===================== producers =====================
    const int ptid = tid;
    const int LAG = B_REG ? 0 : S - 1;  // cp.async stages kept in flight per thread
    const int a_groups = BM >> 5, b_groups = b_ext >> 5;
    for (int i = 0; i < nk + LAG; ++i) {
      if (i < nk) {
        const int s = i % S, round = i / S;
        const int k0 = (ks_begin + i) * BK;
        // ---- A: global -> registers (gather by token id) BEFORE waiting for the slot ----
        float4 av[8];
        int arow[8], ac[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          int nvalid;
          chunk_coord<A_MN>(q, ptid, BM, arow[q], ac[q]);
          const float* src = chunk_src<A_MN>(arow[q], ac[q], p.A, p.lda, p.a_gather, p.a_gather_limit, m0, p.M, k0,
                                             p.K, nvalid);
          av[q] = ldg_chunk(src, nvalid);
        }
        mbar_wait(smem_u32(&empty_bar[s]), (uint32_t)((round & 1) ^ 1));
        uint8_t* pa = smem + (size_t)s * stage_bytes;
        const uint32_t sb = smem_base + (uint32_t)s * stage_bytes + a_bytes;
        if (!B_REG) {
          stage_async<B_MN>(sb, p.B, p.ldb, n0, p.N, b_ext, k0, p.K, ptid);
          if (X3) stage_async<B_MN>(sb + b_tile, p.B_lo, p.ldb, n0, p.N, b_ext, k0, p.K, ptid);
        }
        // ---- A: dropout mask (zeroing only; 1/(1-p) is applied as alpha in the epilogue),
        //         round to tf32, store with the UMMA swizzle ----
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float4 v = av[q];
          if (p.a_drop.on()) {
            const uint64_t srow = A_MN ? (uint64_t)(k0 + arow[q]) : (uint64_t)(m0 + arow[q]);
            const uint64_t scol = A_MN ? (uint64_t)(m0 + ac[q] * 4) : (uint64_t)(k0 + ac[q] * 4);
            const float4 f = p.a_drop.factor4(srow * (uint64_t)p.a_drop_ld + scol);  // scale forced to 1
            v.x *= f.x; v.y *= f.y; v.z *= f.z; v.w *= f.w;
          }
          const uint32_t off = A_MN ? mn_off(arow[q], ac[q], a_groups) : k_off(arow[q], ac[q]);
          if (X3) {
            float4 hi, lo;
            split4(v, hi, lo);
            *reinterpret_cast<float4*>(pa + off) = hi;
            *reinterpret_cast<float4*>(pa + a_tile + off) = lo;
          } else {
            v.x = rn_tf32(v.x); v.y = rn_tf32(v.y); v.z = rn_tf32(v.z); v.w = rn_tf32(v.w);
            *reinterpret_cast<float4*>(pa + off) = v;
          }
        }
        if (B_REG) {
          uint8_t* pb = pa + a_bytes;
#pragma unroll 1
          for (int h = 0; h < 2; ++h) {
            float4 bv[8];
            int brow[8], bc[8];
            bool bok[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              int nvalid;
              bok[q] = chunk_coord<B_MN>(h * 8 + q, ptid, b_ext, brow[q], bc[q]);
              bv[q] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (bok[q]) {
                const float* src = chunk_src<B_MN>(brow[q], bc[q], p.B, p.ldb, nullptr, 0, n0, p.N, k0, p.K, nvalid);
                bv[q] = ldg_chunk(src, nvalid);
              }
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              if (!bok[q]) continue;
              float4 v = bv[q];
              const uint32_t off = B_MN ? mn_off(brow[q], bc[q], b_groups) : k_off(brow[q], bc[q]);
              if (X3) {
                float4 hi, lo;
                split4(v, hi, lo);
                *reinterpret_cast<float4*>(pb + off) = hi;
                *reinterpret_cast<float4*>(pb + b_tile + off) = lo;
              } else {
                v.x = rn_tf32(v.x); v.y = rn_tf32(v.y); v.z = rn_tf32(v.z); v.w = rn_tf32(v.w);
                *reinterpret_cast<float4*>(pb + off) = v;
              }
            }
          }
        }
      }
      if (!B_REG) cp_async_commit();
      const int j = i - LAG;  // step whose B copies are now guaranteed complete
      if (j >= 0) {
        if (!B_REG) {
          switch (LAG) {
            case 0: cp_async_wait<0>(); break;
            case 1: cp_async_wait<1>(); break;
            case 2: cp_async_wait<2>(); break;
            case 3: cp_async_wait<3>(); break;
            case 4: cp_async_wait<4>(); break;
            default: cp_async_wait<5>(); break;
          }
        }
        fence_proxy_async();  // this thread's st.shared / cp.async writes -> visible to the tensor core
        mbar_arrive(smem_u32(&full_bar[j % S]));
      }
    }
    // ===================== epilogue =====================
    mbar_wait(smem_u32(&accum_bar), 0);
    tc_fence_after();
    const int row = m0 + warp * 32 + lane;
    const uint32_t tbase = tmem + ((uint32_t)(warp * 32) << 16);
    float* crow = p.C + (long)row * p.ldc + n0;
    const bool vec_ok = ((p.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0) && ((n0 & 3) == 0);
    const float alpha = p.alpha;
    for (int c0 = 0; c0 < BN; c0 += 16) {
      float v[16];
      tmem_ld16(tbase + (uint32_t)c0, v);  // warp-collective: executed by all lanes
      if (row < p.M) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int col = c0 + q * 4;
          if (n0 + col >= p.N) break;
          float4 o = make_float4(v[q * 4] * alpha, v[q * 4 + 1] * alpha, v[q * 4 + 2] * alpha, v[q * 4 + 3] * alpha);
          if (vec_ok && n0 + col + 3 < p.N) {
            float4* dst = reinterpret_cast<float4*>(crow + col);
            if (p.out_mode == 0) {
              *dst = o;
            } else if (p.out_mode == 1) {
              float4 c = *dst;
              c.x += o.x; c.y += o.y; c.z += o.z; c.w += o.w;
              *dst = c;
            } else {
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(o.x), "f"(o.y), "f"(o.z),
                           "f"(o.w)
                           : "memory");
            }
          } else {
            const float oe[4] = {o.x, o.y, o.z, o.w};
            for (int e = 0; e < 4; ++e) {
              if (n0 + col + e >= p.N) break;
              float* dst = crow + col + e;
              if (p.out_mode == 0) *dst = oe[e];
              else if (p.out_mode == 1) *dst += oe[e];
              else atomicAdd(dst, oe[e]);
            }
          }
        }
      }
    }
    tc_fence_before();
  } else {
    // ===================== MMA issuer (warp 4) =====================
    if (lane == 0) {
      const uint32_t idesc = make_idesc(A_MN, B_MN, BN);
      const uint32_t a_sbo = A_MN ? (uint32_t)(BM / 32) * 512u : 1024u;
      const uint32_t b_sbo = B_MN ? (uint32_t)(bn_pad / 32) * 512u : 1024u;
      const uint32_t a_lbo = A_MN ? 512u : 16u;
      const uint32_t b_lbo = B_MN ? 512u : 16u;
      const uint32_t a_lt = A_MN ? 1u : 2u, b_lt = B_MN ? 1u : 2u;
      for (int i = 0; i < nk; ++i) {
        const int s = i % S, round = i / S;
        mbar_wait(smem_u32(&full_bar[s]), (uint32_t)(round & 1));
        tc_fence_after();
        const uint32_t sa = smem_base + (uint32_t)s * stage_bytes;
        const uint32_t sb = sa + a_bytes;
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) {
          // K-major: +32 bytes inside the 128-byte swizzle row; MN-major: 8 k = two 4-row k-groups
          const uint32_t a_addr = sa + (A_MN ? (uint32_t)(2 * k) * a_sbo : (uint32_t)k * 32u);
          const uint32_t b_addr = sb + (B_MN ? (uint32_t)(2 * k) * b_sbo : (uint32_t)k * 32u);
          const uint64_t da = make_desc(a_addr, a_lbo, a_sbo, a_lt), db = make_desc(b_addr, b_lbo, b_sbo, b_lt);
          if (X3) {  // small terms first: lo*hi + hi*lo + hi*hi
            umma_tf32(tmem, make_desc(a_addr + a_tile, a_lbo, a_sbo, a_lt), db, idesc, (uint32_t)((i | k) != 0));
            umma_tf32(tmem, da, make_desc(b_addr + b_tile, b_lbo, b_sbo, b_lt), idesc, 1u);
            umma_tf32(tmem, da, db, idesc, 1u);
          } else {
            umma_tf32(tmem, da, db, idesc, (uint32_t)((i | k) != 0));
          }
        }
        umma_commit(smem_u32(&empty_bar[s]));  // frees the stage when these MMAs retire
      }
      umma_commit(smem_u32(&accum_bar));       // accumulator complete
    }
    __syncwarp();
  }
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TMEM_COLS) : "memory");
  }
}

__global__ void zero_matrix_kernel(float* C, int ldc, int M, int N) {
  long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i < (long)M * N) C[(i / N) * (long)ldc + (i % N)] = 0.0f;
}

}  // namespace

int gemm_tf32(const GemmOperandA& A, const float* B, int ldb, bool transB, float* C, int ldc, int M, int N, int K,
              float beta, cudaStream_t st, bool b_rounded, bool x3, const float* B_lo) {
  if (M <= 0 || N <= 0) return EBK_OK;
  EBK_CHECK_ARG(K >= 0 && A.ptr && B && C, "gemm_tf32: null operand");
  // 16-byte cp.async needs 4-float aligned rows; tiny or unaligned problems take the fp32 FMA kernel.
  const bool aligned = (A.lda % 4 == 0) && (ldb % 4 == 0) && ((reinterpret_cast<uintptr_t>(A.ptr) & 15) == 0) &&
                       ((reinterpret_cast<uintptr_t>(B) & 15) == 0) && (!A.drop.on() || A.drop_ld % 4 == 0);
  if (!aligned || K < 8 || (long)M * N * K < (1L << 18)) return gemm_f32(A, B, ldb, transB, C, ldc, M, N, K, beta, st);

  Params p;
  p.A = A.ptr; p.lda = A.lda; p.a_gather = A.gather; p.a_gather_limit = A.gather_limit; p.a_drop = A.drop;
  p.a_drop_ld = A.drop_ld;
  p.alpha = A.drop.on() ? A.drop.scale : 1.0f;  // mask in the operand, scale on the accumulator
  p.a_drop.scale = 1.0f;
  EBK_CHECK_ARG(!(x3 && b_rounded && B_lo == nullptr), "gemm_tf32: 3xTF32 with a pre-split B needs B_lo");
  p.B = B; p.B_lo = B_lo; p.ldb = ldb; p.C = C; p.ldc = ldc; p.M = M; p.N = N; p.K = K;
  const bool a_mn = A.trans, b_mn = !transB;
  const int ntiles_n = ceil_div(N, 256);
  // MN-major B is staged in 32-column swizzle atoms -> keep the UMMA N a whole number of atoms
  const int bn_quant = b_mn ? 32 : 16;
  p.BN = ceil_div(ceil_div(N, ntiles_n), bn_quant) * bn_quant;
  const int bn_pad = (p.BN + 31) & ~31;
  const size_t b_bytes = align_up((size_t)(b_mn ? bn_pad : p.BN) * BK * 4, 1024);
  const size_t stage_bytes = ((size_t)BM * BK * 4 + b_bytes) * (x3 ? 2 : 1);
  const size_t budget = x3 ? 220 * 1024 : 111 * 1024;  // two CTAs per SM (one for the 3-pass variant)
  int stages = (int)((budget - 1024) / stage_bytes);
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  if (stages < 2) stages = 2;
  p.stages = stages;
  p.ksteps_total = ceil_div(K, BK);
  dim3 grid(ceil_div(N, p.BN), ceil_div(M, BM), 1);
  long tiles = (long)grid.x * grid.y;
  int splitk = 1;
  if (tiles < 148 && p.ksteps_total >= 16) {
    splitk = (int)((2L * 148 + tiles - 1) / tiles);
    int maxsplit = p.ksteps_total / 8;
    if (splitk > maxsplit) splitk = maxsplit;
    if (splitk < 1) splitk = 1;
  }
  p.ksteps_per_split = ceil_div(p.ksteps_total, splitk);
  splitk = ceil_div(p.ksteps_total, p.ksteps_per_split);
  grid.z = splitk;
  p.out_mode = splitk > 1 ? 2 : (beta != 0.0f ? 1 : 0);
  if (splitk > 1 && beta == 0.0f) {
    long n = (long)M * N;
    zero_matrix_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(C, ldc, M, N);
    EBK_LAUNCH_CHECK();
  }
  const size_t smem = (size_t)stages * stage_bytes + 1024;
#define LAUNCH4(AMN_, BMN_, BREG_, X3_)                                                                  \
  {                                                                                                      \
    EBK_CUDA(cudaFuncSetAttribute(gemm_tf32_kernel<AMN_, BMN_, BREG_, X3_>,                               \
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));              \
    gemm_tf32_kernel<AMN_, BMN_, BREG_, X3_><<<grid, THREADS, smem, st>>>(p);                             \
  }
#define LAUNCH(AMN_, BMN_)                                   \
  {                                                          \
    if (b_rounded && !x3) LAUNCH4(AMN_, BMN_, false, false)  \
    else if (b_rounded && x3) LAUNCH4(AMN_, BMN_, false, true) \
    else if (!x3) LAUNCH4(AMN_, BMN_, true, false)           \
    else LAUNCH4(AMN_, BMN_, true, true)                     \
  }
  if (!a_mn && !b_mn) LAUNCH(false, false)
  else if (!a_mn && b_mn) LAUNCH(false, true)
  else if (a_mn && !b_mn) LAUNCH(true, false)
  else LAUNCH(true, true)
#undef LAUNCH
#undef LAUNCH4
  EBK_LAUNCH_CHECK();
  return EBK_OK;
}

}  // namespace ebk

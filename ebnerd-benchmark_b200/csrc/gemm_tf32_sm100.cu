// tcgen05 (5th-gen tensor core) TF32 GEMM for sm_100a with a gathered / dropout-masked /
// transposed A operand:   C[M,N] (+)= alpha * opA(A)[M,K] . opB(B)[K,N],  fp32 in, fp32 accumulate.
//
// Why kind::tf32: the reference computes in fp32.  tf32 reads the fp32 words as they sit in the
// embedding table / activations (no converted copy in HBM) and keeps 10 mantissa bits.  The tensor
// core TRUNCATES fp32 operands to tf32, which biases long dot products, so every operand is rounded to
// nearest on its way into shared memory (+0x1000 on the bit pattern).  X3 = error-compensated 3xTF32
// (hi/lo split of both operands, three MMAs per k-chunk) gives ~fp32 accuracy for the inference
// forward, whose click scores must match the fp32 reference to 1e-3.
//
// Kernel structure: PERSISTENT, warp-specialised, one CTA per SM (416 threads):
//   warps 0-7   producers.  A: ld.global.nc 16 B (rows gathered by token id -- tiled TMA cannot
//               express the gather) into registers, PF k-steps ahead; dropout mask (zeroing only, the
//               1/(1-p) scale is applied as alpha in the epilogue); tf32 rounding; st.shared with the
//               UMMA swizzle applied by hand.  B: cp.async 16 B of pre-rounded data (or the register
//               path for generic callers).  Then fence.proxy.async + mbarrier arrive.
//   warp 8      allocates TMEM (2 x 256 columns: double-buffered accumulator) and issues tcgen05.mma
//               (one elected lane); tcgen05.commit releases smem stages / publishes the accumulator.
//   warps 9-12  epilogue: tcgen05.ld (TMEM -> registers) -> alpha -> global store / red.add, overlapped
//               with the main loop of the next tile.
// Work items = (m-tile, n-tile, k-split) enumerated n-fastest so the gathered A rows stay L2-hot.
//
// Shared-memory operand layouts (both "rows of 128 bytes"):
//   K-major  (SWIZZLE_128B):         row = mn index, 128 B = 32 consecutive k; 16-byte chunk c of row r
//                                    sits at r*128 + ((c ^ (r&7)) << 4)                 (SBO = 1024)
//   MN-major (SWIZZLE_128B_BASE32B): row = k index, 128 B = 32 consecutive mn.  32-bit MN-major operands
//                                    only exist in this layout (cute Layout_MN_SW128_32B_Atom): atoms of
//                                    [4 k-rows][32 mn], 32-byte units XOR-swizzled by the k-row
//                                    (Swizzle<2,5,2> on the byte address), atoms laid out
//                                    [k-group][mn-group]                       (LBO = 512, SBO = groups*512)
#include "ebk_common.cuh"

namespace ebk {
namespace {

constexpr int BM = 128;
constexpr int BK = 32;             // fp32 elements per stage along K = one 128-byte swizzle row
constexpr int UMMA_K = 8;          // tf32
constexpr int MAX_STAGES = 8;
constexpr int PROD_WARPS = 16;
constexpr int PROD_THREADS = PROD_WARPS * 32;
constexpr int MMA_WARP = PROD_WARPS;
constexpr int EPI_THREADS = 128;
constexpr int THREADS = PROD_THREADS + 32 + EPI_THREADS;  // 416
constexpr int TMEM_COLS = 512;     // two 256-column accumulator buffers
constexpr int PF = 3;              // A register prefetch depth (k-steps)
constexpr int A_CH = BM * BK / 4 / PROD_THREADS;  // 16-byte A chunks per producer thread per k-step = 4
constexpr int B_CH = 256 * BK / 4 / PROD_THREADS; // max B chunks per thread per k-step = 8
constexpr uint32_t SPIN_LIMIT = 1u << 27;  // bounded mbarrier spins: a protocol bug traps instead of hanging the GPU

struct Params {
  const float* A; int lda; const int32_t* a_gather; int a_gather_limit; Dropout a_drop; int a_drop_ld;
  const float* B; int ldb;
  const float* B_lo;  // X3 + pre-split B: the low parts (same layout as B)
  float* C; int ldc;
  int M, N, K;
  int BN;             // tile width, multiple of 16 (32 for MN-major B), <= 256
  int stages, lag;
  int tiles_m, tiles_n, splitk;
  int ksteps_total;   // ceil(K / BK)
  int ksteps_per_split;
  int out_mode;       // 0 store, 1 load-add-store (beta=1), 2 atomic add (split-K)
  float alpha;        // output scale (carries the dropout 1/(1-p) of a masked A operand)
  uint32_t a_prefetch_bytes;  // K-major A: bytes of each row to prefetch into L2 per item (0 = off)
  long long* dbg;             // optional timeline buffer (CTA 0): [role][step][4] clock64 stamps
};
constexpr int DBG_STEPS = 96;
__device__ __forceinline__ void dbg_stamp(long long* dbg, int role, int step, int slot) {
#ifdef EBK_GEMM_TIMELINE  // tools/gemm_timeline.py; compiled out of the product build
  if (dbg != nullptr && blockIdx.x == 0 && step < DBG_STEPS) dbg[(role * DBG_STEPS + step) * 4 + slot] = clock64();
#endif
}

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
// waiting roles that are idle most of the time (MMA issuer, epilogue) back off so that their spin
// does not steal issue slots from the producer warps
__device__ __forceinline__ void mbar_wait_backoff(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    __nanosleep(32);
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
__device__ __forceinline__ void cp_async_wait_dyn(int n) {
  switch (n) {
    case 0: cp_async_wait<0>(); break;
    case 1: cp_async_wait<1>(); break;
    case 2: cp_async_wait<2>(); break;
    case 3: cp_async_wait<3>(); break;
    case 4: cp_async_wait<4>(); break;
    case 5: cp_async_wait<5>(); break;
    default: cp_async_wait<6>(); break;
  }
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
// LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48), layout type [61,64).
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
    row = (ptid >> 3) + i * (PROD_THREADS / 8);
    c = ptid & 7;
    return row < ext;
  } else {
    const int cpr = ext >> 2;
    const int idx = ptid + i * PROD_THREADS;
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
__device__ __forceinline__ float4 rn4(float4 v) {
  v.x = rn_tf32(v.x); v.y = rn_tf32(v.y); v.z = rn_tf32(v.z); v.w = rn_tf32(v.w);
  return v;
}
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
#pragma unroll
  for (int i = 0; i < B_CH; ++i) {
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

// Walks this CTA's work items (m-tile, n-tile, k-split; n fastest) k-step by k-step.
struct Cursor {
  int item, ks, ks_end;   // current item, current k-step, end k-step of the item
  int m0, n0, tn;
  __device__ __forceinline__ void load(const Params& p) {
    const int per_m = p.tiles_n * p.splitk;
    const int tm = item / per_m, rem = item - tm * per_m;
    tn = rem / p.splitk;
    const int sp = rem - tn * p.splitk;
    m0 = tm * BM;
    n0 = tn * p.BN;
    ks = sp * p.ksteps_per_split;
    ks_end = min(p.ksteps_total, ks + p.ksteps_per_split);
  }
  __device__ __forceinline__ void init(const Params& p, int first, int n_items) {
    item = first;
    ks = ks_end = m0 = n0 = tn = 0;
    if (item < n_items) load(p);
  }
  __device__ __forceinline__ bool valid(int n_items) const { return item < n_items; }
  // returns true when the step just consumed was the last of its item
  __device__ __forceinline__ bool advance(const Params& p, int n_items, int stride) {
    if (++ks < ks_end) return false;
    item += stride;
    if (item < n_items) load(p);
    return true;
  }
};

// BMODE selects how the B operand reaches shared memory:
//   0  packed: B was pre-rounded AND pre-arranged by gemm_tf32_pack_b() into the exact UMMA tile
//      layout, [n-tile][k-step] blocks of b_tile bytes -> ONE cp.async.bulk (TMA engine, no tensor map)
//      per stage, completion counted on the stage's full barrier (complete_tx);
//   1  row-major, already tf32-rounded: 16-byte cp.async per chunk (activations: dQKV, dpre);
//   2  row-major fp32 through registers with in-flight rounding (generic callers).
template <bool A_MN, bool B_MN, int BMODE, bool X3>
__global__ void __launch_bounds__(THREADS, 1) gemm_tf32_kernel(const Params p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[MAX_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[MAX_STAGES];
  __shared__ __align__(8) uint64_t tfull_bar[2];
  __shared__ __align__(8) uint64_t tempty_bar[2];
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
  const int n_items = p.tiles_m * p.tiles_n * p.splitk;

  if (tid == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(smem_u32(&full_bar[s]), PROD_THREADS + (BMODE == 0 ? 1 : 0));  // + the expect_tx arrival
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&tfull_bar[b]), 1);
      mbar_init(smem_u32(&tempty_bar[b]), EPI_THREADS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)),
                 "n"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (warp < PROD_WARPS) {
    // ===================== producers =====================
    // Everything that does not depend on the k-step is hoisted: chunk coordinates and swizzled smem
    // offsets are per-thread constants, row pointers / validity / dropout group ids are per work item.
    const int ptid = tid;
    const int LAG = BMODE == 1 ? p.lag : 0;  // cp.async groups kept in flight before a stage is published
    const int a_groups = BM >> 5, b_groups = b_ext >> 5;
    int arow[A_CH], ac[A_CH];
    uint32_t aoff[A_CH];
#pragma unroll
    for (int q = 0; q < A_CH; ++q) {
      chunk_coord<A_MN>(q, ptid, BM, arow[q], ac[q]);
      aoff[q] = A_MN ? mn_off(arow[q], ac[q], a_groups) : k_off(arow[q], ac[q]);
    }
    int brow[B_CH], bc[B_CH];
    uint32_t boff[B_CH];
    bool bok[B_CH];
    if (BMODE != 0) {
#pragma unroll
      for (int q = 0; q < B_CH; ++q) {
        bok[q] = chunk_coord<B_MN>(q, ptid, b_ext, brow[q], bc[q]);
        boff[q] = B_MN ? mn_off(brow[q], bc[q], b_groups) : k_off(brow[q], bc[q]);
      }
    }

    Cursor ld, pr;
    ld.init(p, blockIdx.x, n_items);
    pr.init(p, blockIdx.x, n_items);
    // ---- per-item state of the LOAD cursor ----
    int ld_item = -1;
    const float* aptr[A_CH];   // K-major: &A[row(m), 4c] or nullptr (zero row); MN-major: &A[0, m0+4c]
    int avalid[A_CH];          // MN-major: valid floats of the chunk (0..4)
    auto a_item_setup = [&]() {
      ld_item = ld.item;
#pragma unroll
      for (int q = 0; q < A_CH; ++q) {
        if (!A_MN) {
          const int m = ld.m0 + arow[q];
          long r = m;
          bool ok = m < p.M;
          if (p.a_gather) {
            const int g = __ldg(p.a_gather + (ok ? m : 0));
            ok = ok && (g >= 0) && (g < p.a_gather_limit);
            r = g;
          }
          aptr[q] = ok ? p.A + r * (long)p.lda + ac[q] * 4 : nullptr;
          avalid[q] = 4;
          // One thread per row asks the L2 for the whole row (all k-steps of this item): DRAM then sees
          // one sequential burst per gathered row instead of K/32 scattered 128-byte reads.
          if (ok && ac[q] == 0 && p.a_prefetch_bytes)
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(aptr[q]), "r"(p.a_prefetch_bytes) : "memory");
        } else {
          const int m = ld.m0 + ac[q] * 4;
          int nv = p.M - m;
          avalid[q] = nv < 0 ? 0 : (nv > 4 ? 4 : nv);
          aptr[q] = p.A + m;
        }
      }
    };
    // issue the A loads of the step under the load cursor into `dst`
    auto a_load = [&](float4 (&dst)[A_CH]) {
      if (ld.item != ld_item) a_item_setup();
      const int k0 = ld.ks * BK;
      if (!A_MN) {
        if (k0 + BK <= p.K) {  // fast path: whole 128-byte row segments
#pragma unroll
          for (int q = 0; q < A_CH; ++q) {
            dst[q] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (aptr[q] != nullptr) dst[q] = ldg_chunk(aptr[q] + k0, 4);
          }
        } else {
          int nv = p.K - (k0 + ac[0] * 4);
          nv = nv > 4 ? 4 : nv;
#pragma unroll
          for (int q = 0; q < A_CH; ++q) {
            const bool ok = aptr[q] != nullptr && nv > 0;
            dst[q] = ldg_chunk(ok ? aptr[q] + k0 : p.A, ok ? nv : 0);
          }
        }
      } else {
        // MN-major: the storage rows are the k indices (gathered); index loads issued back to back
        int g[A_CH];
        bool ok[A_CH];
#pragma unroll
        for (int q = 0; q < A_CH; ++q) {
          const int k = k0 + arow[q];
          ok[q] = k < p.K && avalid[q] > 0;
          g[q] = k;
          if (p.a_gather) g[q] = __ldg(p.a_gather + (ok[q] ? k : 0));
        }
#pragma unroll
        for (int q = 0; q < A_CH; ++q) {
          if (p.a_gather) ok[q] = ok[q] && g[q] >= 0 && g[q] < p.a_gather_limit;
          dst[q] = ldg_chunk(ok[q] ? aptr[q] + (long)g[q] * p.lda : p.A, ok[q] ? avalid[q] : 0);
        }
      }
    };
    // ---- per-item state of the PROCESS cursor: dropout group id of each chunk at k-step 0 ----
    // (the mask is re-hashed here: measured on B200, fetching precomputed keep bits is not faster)
    int pr_item = -1;
    uint64_t dgrp[A_CH];
    uint64_t dstep = 0;  // group-id increment per k-step
    auto d_item_setup = [&]() {
      pr_item = pr.item;
#pragma unroll
      for (int q = 0; q < A_CH; ++q) {
        const uint64_t srow = A_MN ? (uint64_t)arow[q] : (uint64_t)(pr.m0 + arow[q]);
        const uint64_t scol = A_MN ? (uint64_t)(pr.m0 + ac[q] * 4) : (uint64_t)(ac[q] * 4);
        dgrp[q] = (srow * (uint64_t)p.a_drop_ld + scol) >> 2;
      }
      dstep = A_MN ? ((uint64_t)BK * (uint64_t)p.a_drop_ld) >> 2 : (uint64_t)(BK / 4);
    };
    float4 av[PF][A_CH];
    // prologue: A loads of the first PF-1 steps
#pragma unroll
    for (int u = 0; u < PF - 1; ++u) {
#pragma unroll
      for (int q = 0; q < A_CH; ++q) av[u][q] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ld.valid(n_items)) {
        a_load(av[u]);
        ld.advance(p, n_items, gridDim.x);
      }
    }
    int g = 0;          // flat k-step counter of this CTA
    int published = 0;  // steps whose full barrier has been arrived on
    int s = 0;          // stage of step g, and the parity of its use (no runtime div/mod in the loop)
    uint32_t ph = 0;
    int ps = 0;         // stage of the next step to publish
    while (pr.valid(n_items)) {
#pragma unroll
      for (int u = 0; u < PF; ++u) {
        if (!pr.valid(n_items)) break;
        // ---- A loads for step g+PF-1 into the slot freed last iteration ----
        constexpr int PFM1 = PF - 1;
        if (ld.valid(n_items)) {
          a_load(av[(u + PFM1) % PF]);
          ld.advance(p, n_items, gridDim.x);
        }
        const int k0 = pr.ks * BK;
        if (ptid == 0) dbg_stamp(p.dbg, 0, g, 0);
        mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1u);
        if (ptid == 0) dbg_stamp(p.dbg, 0, g, 1);
        uint8_t* pa = smem + (size_t)s * stage_bytes;
        uint8_t* pb = pa + a_bytes;
        const uint32_t sb = smem_base + (uint32_t)s * stage_bytes + a_bytes;
        if (BMODE == 0) {
          // ---- B: one bulk copy of the pre-packed tile (hi [+ lo]) ----
          if (ptid == 0) {
            const uint32_t bar = smem_u32(&full_bar[s]);
            const uint32_t bytes = NSPLIT * b_tile;
            const size_t blk = ((size_t)pr.tn * p.ksteps_total + pr.ks) * (size_t)(b_tile / 4);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sb),
                         "l"(p.B + blk), "r"(b_tile), "r"(bar)
                         : "memory");
            if (X3)
              asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                               sb + b_tile),
                           "l"(p.B_lo + blk), "r"(b_tile), "r"(bar)
                           : "memory");
          }
        } else if (BMODE == 1) {
          // ---- B: cp.async of pre-rounded row-major data ----
          if (!B_MN) {
            int nvk = p.K - (k0 + bc[0] * 4);
            nvk = nvk < 0 ? 0 : (nvk > 4 ? 4 : nvk);
#pragma unroll
            for (int q = 0; q < B_CH; ++q) {
              if (!bok[q]) break;
              const int n = pr.n0 + brow[q];
              const bool ok = n < p.N;
              const long off = ok ? (long)n * p.ldb + k0 + bc[q] * 4 : 0;
              cp_async16(sb + boff[q], p.B + off, ok ? (uint32_t)nvk * 4u : 0u);
              if (X3) cp_async16(sb + b_tile + boff[q], p.B_lo + off, ok ? (uint32_t)nvk * 4u : 0u);
            }
          } else {
#pragma unroll
            for (int q = 0; q < B_CH; ++q) {
              if (!bok[q]) break;
              const int k = k0 + brow[q];
              const int n = pr.n0 + bc[q] * 4;
              int nv = p.N - n;
              nv = (k < p.K && nv > 0) ? (nv > 4 ? 4 : nv) : 0;
              const long off = nv ? (long)k * p.ldb + n : 0;
              cp_async16(sb + boff[q], p.B + off, (uint32_t)nv * 4u);
              if (X3) cp_async16(sb + b_tile + boff[q], p.B_lo + off, (uint32_t)nv * 4u);
            }
          }
          cp_async_commit();
        }
        // ---- A: mask, round to tf32, store with the UMMA swizzle ----
        if (p.a_drop.on() && pr.item != pr_item) d_item_setup();
#pragma unroll
        for (int q = 0; q < A_CH; ++q) {
          float4 v = av[u][q];
          if (p.a_drop.on()) {
            const float4 f = p.a_drop.factor4_group(dgrp[q] + (uint64_t)pr.ks * dstep);  // scale forced to 1
            v.x *= f.x; v.y *= f.y; v.z *= f.z; v.w *= f.w;
          }
          if (X3) {
            float4 hi, lo;
            split4(v, hi, lo);
            *reinterpret_cast<float4*>(pa + aoff[q]) = hi;
            *reinterpret_cast<float4*>(pa + a_tile + aoff[q]) = lo;
          } else {
            *reinterpret_cast<float4*>(pa + aoff[q]) = rn4(v);
          }
        }
        if (BMODE == 2) {
          // ---- B through registers with in-flight rounding (generic callers) ----
#pragma unroll
          for (int h = 0; h < B_CH; h += 4) {
            float4 bv[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              int nvalid = 0;
              const float* src = p.B;
              if (bok[h + q])
                src = chunk_src<B_MN>(brow[h + q], bc[h + q], p.B, p.ldb, nullptr, 0, pr.n0, p.N, k0, p.K, nvalid);
              bv[q] = ldg_chunk(src, nvalid);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              if (!bok[h + q]) continue;
              if (X3) {
                float4 hi, lo;
                split4(bv[q], hi, lo);
                *reinterpret_cast<float4*>(pb + boff[h + q]) = hi;
                *reinterpret_cast<float4*>(pb + b_tile + boff[h + q]) = lo;
              } else {
                *reinterpret_cast<float4*>(pb + boff[h + q]) = rn4(bv[q]);
              }
            }
          }
        }
        if (ptid == 0) dbg_stamp(p.dbg, 0, g, 2);
        // ---- publish step g-LAG (its cp.async group is complete once <= LAG younger groups are pending) ----
        if (g - LAG >= 0) {
          if (BMODE == 1) cp_async_wait_dyn(LAG);
          fence_proxy_async();  // this thread's st.shared / cp.async writes -> visible to the tensor core
          mbar_arrive(smem_u32(&full_bar[ps]));
          if (++ps == S) ps = 0;
          published = g - LAG + 1;
        }
        if (ptid == 0) dbg_stamp(p.dbg, 0, g, 3);
        pr.advance(p, n_items, gridDim.x);
        ++g;
        if (++s == S) {
          s = 0;
          ph ^= 1u;
        }
      }
    }
    // drain: publish the last LAG steps
    if (BMODE == 1) cp_async_wait<0>();
    fence_proxy_async();
    for (int j = published; j < g; ++j) {
      mbar_arrive(smem_u32(&full_bar[ps]));
      if (++ps == S) ps = 0;
    }
  } else if (warp == MMA_WARP) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = make_idesc(A_MN, B_MN, BN);
      const uint32_t a_sbo = A_MN ? (uint32_t)(BM / 32) * 512u : 1024u;
      const uint32_t b_sbo = B_MN ? (uint32_t)(bn_pad / 32) * 512u : 1024u;
      const uint32_t a_lbo = A_MN ? 512u : 16u;
      const uint32_t b_lbo = B_MN ? 512u : 16u;
      const uint32_t a_lt = A_MN ? 1u : 2u, b_lt = B_MN ? 1u : 2u;
      Cursor cu;
      cu.init(p, blockIdx.x, n_items);
      int g = 0, t = 0, s = 0;
      uint32_t ph = 0;
      while (cu.valid(n_items)) {
        const int buf = t & 1, use = t >> 1;
        mbar_wait_backoff(smem_u32(&tempty_bar[buf]), (uint32_t)((use & 1) ^ 1));  // epilogue drained this buffer
        tc_fence_after();
        const uint32_t tacc = tmem + (uint32_t)buf * 256u;
        bool first = true, last = false;
        while (!last) {
          dbg_stamp(p.dbg, 1, g, 0);
          mbar_wait(smem_u32(&full_bar[s]), ph);
          dbg_stamp(p.dbg, 1, g, 1);
          tc_fence_after();
          const uint32_t sa = smem_base + (uint32_t)s * stage_bytes;
          const uint32_t sb = sa + a_bytes;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // K-major: +32 bytes inside the 128-byte swizzle row; MN-major: 8 k = two 4-row k-groups
            const uint32_t a_addr = sa + (A_MN ? (uint32_t)(2 * k) * a_sbo : (uint32_t)k * 32u);
            const uint32_t b_addr = sb + (B_MN ? (uint32_t)(2 * k) * b_sbo : (uint32_t)k * 32u);
            const uint64_t da = make_desc(a_addr, a_lbo, a_sbo, a_lt), db = make_desc(b_addr, b_lbo, b_sbo, b_lt);
            const uint32_t acc = (first && k == 0) ? 0u : 1u;
            if (X3) {  // small terms first: lo*hi + hi*lo + hi*hi
              umma_tf32(tacc, make_desc(a_addr + a_tile, a_lbo, a_sbo, a_lt), db, idesc, acc);
              umma_tf32(tacc, da, make_desc(b_addr + b_tile, b_lbo, b_sbo, b_lt), idesc, 1u);
              umma_tf32(tacc, da, db, idesc, 1u);
            } else {
              umma_tf32(tacc, da, db, idesc, acc);
            }
          }
          umma_commit(smem_u32(&empty_bar[s]));  // frees the stage when these MMAs retire
          dbg_stamp(p.dbg, 1, g, 2);
          first = false;
          last = cu.advance(p, n_items, gridDim.x);
          ++g;
          if (++s == S) {
            s = 0;
            ph ^= 1u;
          }
        }
        umma_commit(smem_u32(&tfull_bar[buf]));  // accumulator of this item complete
        ++t;
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue (warps 9-12) =====================
    const int ew = warp & 3;  // TMEM lane quarter this warp may access
    Cursor cu;
    cu.init(p, blockIdx.x, n_items);
    int t = 0;
    const float alpha = p.alpha;
    while (cu.valid(n_items)) {
      const int buf = t & 1, use = t >> 1;
      const int m0 = cu.m0, n0 = cu.n0;
      if (ew == 0 && lane == 0) dbg_stamp(p.dbg, 2, t, 0);
      mbar_wait_backoff(smem_u32(&tfull_bar[buf]), (uint32_t)(use & 1));
      if (ew == 0 && lane == 0) dbg_stamp(p.dbg, 2, t, 1);
      tc_fence_after();
      const int row = m0 + ew * 32 + lane;
      const uint32_t tbase = tmem + ((uint32_t)(ew * 32) << 16) + (uint32_t)buf * 256u;
      float* crow = p.C + (long)row * p.ldc + n0;
      const bool vec_ok = ((p.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0) && ((n0 & 3) == 0);
      const bool vec8_ok = vec_ok && ((p.ldc & 7) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 31) == 0) &&
                           ((n0 & 7) == 0) && p.out_mode == 0;
      for (int c0 = 0; c0 < BN; c0 += 16) {
        float v[16];
        tmem_ld16(tbase + (uint32_t)c0, v);  // warp-collective: executed by all lanes
        if (row < p.M) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] *= alpha;
          if (vec8_ok && n0 + c0 + 15 < p.N) {
            // two 256-bit stores: each fills a whole 32-byte sector of the row
            asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(crow + c0), "f"(v[0]), "f"(v[1]),
                         "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
                         : "memory");
            asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(crow + c0 + 8), "f"(v[8]), "f"(v[9]),
                         "f"(v[10]), "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15])
                         : "memory");
            continue;
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int col = c0 + q * 4;
            if (n0 + col >= p.N) break;
            float4 o = make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
            if (vec_ok && n0 + col + 3 < p.N) {
              float4* dst = reinterpret_cast<float4*>(crow + col);
              if (p.out_mode == 0) {
                *dst = o;
              } else if (p.out_mode == 1) {
                float4 c = *dst;
                c.x += o.x; c.y += o.y; c.z += o.z; c.w += o.w;
                *dst = c;
              } else {
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(o.x), "f"(o.y),
                             "f"(o.z), "f"(o.w)
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
      mbar_arrive(smem_u32(&tempty_bar[buf]));  // buffer may be overwritten by the MMA warp
      if (ew == 0 && lane == 0) dbg_stamp(p.dbg, 2, t, 2);
      // skip to this CTA's next item
      cu.ks = cu.ks_end - 1;
      cu.advance(p, n_items, gridDim.x);
      ++t;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TMEM_COLS) : "memory");
  }
}

__global__ void zero_matrix_kernel(float* C, int ldc, int M, int N) {
  long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i < (long)M * N) C[(i / N) * (long)ldc + (i % N)] = 0.0f;
}

// Writes B (row-major fp32) as [n-tile][k-step] blocks in the exact shared-memory tile layout of the
// kernel (swizzle included), rounded to tf32 (hi) and optionally the 3xTF32 low part (lo).
__global__ void pack_b_kernel(float* __restrict__ dst_hi, float* __restrict__ dst_lo, const float* __restrict__ B,
                              int ldb, int b_mn, int N, int K, int BN, int b_ext, int tiles_n, int ksteps,
                              int chunks_per_tile) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  const long total = (long)tiles_n * ksteps * chunks_per_tile;
  if (i >= total) return;
  const int L = (int)(i % chunks_per_tile);
  const long blk = i / chunks_per_tile;
  const int ks = (int)(blk % ksteps), tn = (int)(blk / ksteps);
  const int n0 = tn * BN, k0 = ks * BK;
  float v[4] = {0.f, 0.f, 0.f, 0.f};
  if (!b_mn) {
    const int r = L >> 3, c = (L & 7) ^ (r & 7);
    const int n = n0 + r;
    if (r < b_ext && n < N) {
      for (int e = 0; e < 4; ++e) {
        const int k = k0 + c * 4 + e;
        if (k < K) v[e] = B[(long)n * ldb + k];
      }
    }
  } else {
    const int groups = b_ext >> 5;
    const int atom = L >> 5, within = L & 31;
    const int kr_lo = within >> 3, pos = within & 7;
    const int c7 = (((pos >> 1) ^ kr_lo) << 1) | (pos & 1);
    const int kg = atom / groups, ng = atom - kg * groups;
    const int kr = kg * 4 + kr_lo, c = ng * 8 + c7;
    const int k = k0 + kr;
    if (kr < BK && k < K) {
      for (int e = 0; e < 4; ++e) {
        const int n = n0 + c * 4 + e;
        if (n < N) v[e] = B[(long)k * ldb + n];
      }
    }
  }
  float4 hi, lo;
  hi.x = round_tf32_bits(v[0]); hi.y = round_tf32_bits(v[1]); hi.z = round_tf32_bits(v[2]); hi.w = round_tf32_bits(v[3]);
  reinterpret_cast<float4*>(dst_hi)[i] = hi;
  if (dst_lo) {
    lo.x = round_tf32_bits(v[0] - hi.x); lo.y = round_tf32_bits(v[1] - hi.y);
    lo.z = round_tf32_bits(v[2] - hi.z); lo.w = round_tf32_bits(v[3] - hi.w);
    reinterpret_cast<float4*>(dst_lo)[i] = lo;
  }
}

int g_num_sms = 0;
long long* g_dbg = nullptr;
int g_dbg_target = 0, g_dbg_count = 0;  // only the g_dbg_target-th GEMM launch after arming is traced

struct BGeom {
  bool b_mn;
  int BN, b_ext, tiles_n, ksteps;
  size_t b_tile;  // bytes per packed block
};
BGeom b_geom(int N, int K, bool transB) {
  BGeom g;
  g.b_mn = !transB;
  const int ntiles_n = ceil_div(N, 256);
  // MN-major B is staged in 32-column swizzle atoms -> keep the UMMA N a whole number of atoms
  const int bn_quant = g.b_mn ? 32 : 16;
  g.BN = ceil_div(ceil_div(N, ntiles_n), bn_quant) * bn_quant;
  g.b_ext = g.b_mn ? ((g.BN + 31) & ~31) : g.BN;
  g.tiles_n = ceil_div(N, g.BN);
  g.ksteps = ceil_div(K, BK);
  g.b_tile = align_up((size_t)g.b_ext * BK * 4, 1024);
  return g;
}

}  // namespace

void gemm_tf32_set_debug(long long* buf, int target) {
  g_dbg = buf;
  g_dbg_target = target;
  g_dbg_count = 0;
}

bool gemm_tf32_eligible(const GemmOperandA& A, const float* B, int ldb, int M, int N, int K) {
  // 16-byte loads need 4-float aligned rows; tiny or unaligned problems take the fp32 FMA kernel.
  const bool aligned = (A.lda % 4 == 0) && (ldb % 4 == 0) && ((reinterpret_cast<uintptr_t>(A.ptr) & 15) == 0) &&
                       ((reinterpret_cast<uintptr_t>(B) & 15) == 0) && (!A.drop.on() || A.drop_ld % 4 == 0);
  return aligned && K >= 8 && (long)M * N * K >= (1L << 18);
}

int gemm_tf32_bn(int N, int K, bool transB) { return b_geom(N, K, transB).BN; }

size_t gemm_tf32_packed_floats(int N, int K, bool transB) {
  const BGeom g = b_geom(N, K, transB);
  return (size_t)g.tiles_n * g.ksteps * (g.b_tile / 4);
}

int gemm_tf32_pack_b(float* dst_hi, float* dst_lo, const float* B, int ldb, bool transB, int N, int K,
                     cudaStream_t st) {
  const BGeom g = b_geom(N, K, transB);
  const int chunks_per_tile = (int)(g.b_tile / 16);
  const long total = (long)g.tiles_n * g.ksteps * chunks_per_tile;
  pack_b_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(dst_hi, dst_lo, B, ldb, g.b_mn ? 1 : 0, N, K, g.BN,
                                                                 g.b_ext, g.tiles_n, g.ksteps, chunks_per_tile);
  EBK_LAUNCH_CHECK();
  return EBK_OK;
}

int gemm_tf32(const GemmOperandA& A, const float* B, int ldb, bool transB, float* C, int ldc, int M, int N, int K,
              float beta, cudaStream_t st, int b_mode, bool x3, const float* B_lo) {
  if (M <= 0 || N <= 0) return EBK_OK;
  EBK_CHECK_ARG(K >= 0 && A.ptr && B && C, "gemm_tf32: null operand");
  if (b_mode != GEMM_B_PACKED && !gemm_tf32_eligible(A, B, ldb, M, N, K))
    return gemm_f32(A, B, ldb, transB, C, ldc, M, N, K, beta, st);
  EBK_CHECK_ARG(!(x3 && b_mode != GEMM_B_RAW && B_lo == nullptr), "gemm_tf32: 3xTF32 with a pre-split B needs B_lo");
  if (g_num_sms == 0) {
    int dev = 0;
    EBK_CUDA(cudaGetDevice(&dev));
    EBK_CUDA(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
  }

  Params p;
  p.A = A.ptr; p.lda = A.lda; p.a_gather = A.gather; p.a_gather_limit = A.gather_limit; p.a_drop = A.drop;
  p.a_drop_ld = A.drop_ld;
  p.alpha = A.drop.on() ? A.drop.scale : 1.0f;  // mask in the operand, scale on the accumulator
  p.a_drop.scale = 1.0f;
  // whole-row L2 prefetch pays when the rows are scattered (gather) or strided; rows must be 16B multiples
  p.a_prefetch_bytes = (!A.trans && K >= 64) ? (uint32_t)((K * 4) & ~15) : 0u;
  p.dbg = (g_dbg != nullptr && g_dbg_count++ == g_dbg_target) ? g_dbg : nullptr;
  p.B = B; p.B_lo = B_lo; p.ldb = ldb; p.C = C; p.ldc = ldc; p.M = M; p.N = N; p.K = K;
  const bool a_mn = A.trans;
  const BGeom bg = b_geom(N, K, transB);
  const bool b_mn = bg.b_mn;
  p.BN = bg.BN;
  const size_t stage_bytes = ((size_t)BM * BK * 4 + bg.b_tile) * (x3 ? 2 : 1);
  const size_t budget = 200 * 1024;
  int stages = (int)(budget / stage_bytes);
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  if (stages < 2) stages = 2;
  p.stages = stages;
  p.lag = stages >= 4 ? 2 : (stages >= 3 ? 1 : 0);
  p.ksteps_total = bg.ksteps;
  p.tiles_m = ceil_div(M, BM);
  p.tiles_n = bg.tiles_n;
  const long tiles = (long)p.tiles_m * p.tiles_n;
  int splitk = 1;
  if (tiles < g_num_sms && p.ksteps_total >= 16) {
    splitk = (int)(g_num_sms / tiles);  // fill the machine in ONE wave of equal items
    int maxsplit = p.ksteps_total / 8;
    if (splitk > maxsplit) splitk = maxsplit;
    if (splitk < 1) splitk = 1;
  }
  p.ksteps_per_split = ceil_div(p.ksteps_total, splitk);
  splitk = ceil_div(p.ksteps_total, p.ksteps_per_split);
  p.splitk = splitk;
  p.out_mode = splitk > 1 ? 2 : (beta != 0.0f ? 1 : 0);
  if (splitk > 1 && beta == 0.0f) {
    long n = (long)M * N;
    zero_matrix_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(C, ldc, M, N);
    EBK_LAUNCH_CHECK();
  }
  const long n_items = tiles * splitk;
  const int grid = (int)(n_items < g_num_sms ? n_items : g_num_sms);
  const size_t smem = (size_t)stages * stage_bytes + 1024;
#define LAUNCH4(AMN_, BMN_, BMODE_, X3_)                                                                 \
  {                                                                                                      \
    EBK_CUDA(cudaFuncSetAttribute(gemm_tf32_kernel<AMN_, BMN_, BMODE_, X3_>,                              \
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));              \
    gemm_tf32_kernel<AMN_, BMN_, BMODE_, X3_><<<grid, THREADS, smem, st>>>(p);                            \
  }
#define LAUNCH3(AMN_, BMN_, BMODE_)            \
  {                                            \
    if (x3) LAUNCH4(AMN_, BMN_, BMODE_, true)  \
    else LAUNCH4(AMN_, BMN_, BMODE_, false)    \
  }
#define LAUNCH(AMN_, BMN_)                                        \
  {                                                               \
    if (b_mode == GEMM_B_PACKED) LAUNCH3(AMN_, BMN_, 0)           \
    else if (b_mode == GEMM_B_ROUNDED) LAUNCH3(AMN_, BMN_, 1)     \
    else LAUNCH3(AMN_, BMN_, 2)                                   \
  }
  if (!a_mn && !b_mn) LAUNCH(false, false)
  else if (!a_mn && b_mn) LAUNCH(false, true)
  else if (a_mn && !b_mn) LAUNCH(true, false)
  else LAUNCH(true, true)
#undef LAUNCH
#undef LAUNCH3
#undef LAUNCH4
  EBK_LAUNCH_CHECK();
  return EBK_OK;
}

}  // namespace ebk

// All-TMA tcgen05 TF32 GEMM for sm_100a (the TRAINING hot path):
//   C[M,N] (+)= alpha * opA(A)[M,K] . opB(B)[K,N],  fp32 storage, tf32 tensor cores, fp32 accumulate in TMEM.
//
// Why a second GEMM next to gemm_tf32_sm100.cu: ncu showed the register-path GEMM (gather + dropout hash
// + rounding in 16 producer warps) to be ISSUE-bound -- 4160 warp instructions per k-step against the
// 1920 issue slots of the 480-cycle MMA window.  Here every operand is a dense fp32 matrix that the
// layer before it already wrote masked and rounded to tf32 (embed_rows_kernel, attention / attpool
// epilogues), so the operand feed needs no ALU work at all: ONE thread issues cp.async.bulk.tensor
// (TMA) copies that land in shared memory in the UMMA swizzled layouts, one thread issues tcgen05.mma,
// four warps drain the accumulators.  192 threads per CTA, persistent, one CTA per SM.
//
// Operand layouts in shared memory (per pipeline stage, BK = 32 fp32 = one 128-byte swizzle row):
//   K-major  (A not transposed / B transposed: storage [mn, k], k contiguous)
//            ONE 2-D box {32 k, rows}, CU_TENSOR_MAP_SWIZZLE_128B  -> rows x 128 B, UMMA SWIZZLE_128B,
//            SBO = 1024 (8-row atoms), k-advance = +32 B per MMA.
//   MN-major (A transposed / B not transposed: storage [k, mn], mn contiguous)
//            one 2-D box {32 mn, 32 k} per group of 32 mn, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B -> 4 KB
//            blocks [k][32 mn] = the only UMMA layout for 32-bit MN-major operands
//            (SWIZZLE_128B_BASE32B: atoms of 4 k-rows x 128 B, 32-byte units XOR-swizzled by the k-row);
//            blocks of consecutive mn groups follow each other: LBO = 4096, SBO = 512, k-advance = +1024 B.
//   Out-of-range box elements are zero-filled by the TMA unit, so no tail code exists for M, N or K.
//
// PAIR (cta_group::2): measured, the 1-CTA kernel tops out near 62 % tensor-pipe activity because every byte
// of A and B is written to AND read from the SM's shared memory once per MMA (154 KB per k-step of a
// 256 x 240 tile against ~128 B/clk).  In PAIR mode two CTAs of a cluster (one TPC) form one 512-row
// (MT = 2) tile: each stages its own A rows and only HALF of the B tile, and the leader CTA issues
// tcgen05.mma.cta_group::2 (M = 256) that feeds both tensor cores from both shared memories.  The peer's
// TMA loads complete on the LEADER's full barrier (cp.async.bulk.tensor ... .cta_group::2), the leader's
// tcgen05.commit multicasts to both CTAs' empty / accumulator-full barriers, both epilogues release the
// accumulator on the leader's barrier.  (Plain operand MULTICAST between 1-CTA tiles was tried first: it
// cuts L2 traffic but not shared-memory traffic and measured no gain, so it is not kept.)
//
// MT = 1: 128 x BN tiles, double-buffered accumulator (2 x 256 TMEM columns): the epilogue of tile i
//         overlaps the main loop of tile i+1.
// MT = 2: 256 x BN tiles (two M=128 MMAs per k-step sharing the B tile, 2 x 256 TMEM columns, single
//         buffered): halves the L2->SM traffic of the B operand; used where B is the re-read operand.
#include <cuda.h>  // CUtensorMap types only: the encoder is fetched through cudaGetDriverEntryPoint (no libcuda link)

#include <stdlib.h>
#include <string.h>

#include <unordered_map>

#include "attention_tiles.cuh"
#include "ebk_common.cuh"

namespace ebk {
namespace {

constexpr int BM = 128;
constexpr int BK = 32;
constexpr int UMMA_K = 8;
constexpr int MAX_STAGES = 8;
constexpr int THREADS = 192;       // warp 0: TMA producer, warp 1: MMA issuer, warps 2-5: epilogue
constexpr int EPI_THREADS = 128;
// GEMMs with a heavy fused epilogue (dropout hash + pooling term + rounding) run 8 epilogue warps, two per TMEM lane
// quarter splitting the tile's 32-column chunks: one warp alone on its scheduler is latency-bound (ncu: 20k cycles per
// 128x208 tile against a 5k-cycle main loop with 4 warps)
constexpr int MAX_EPI_WARPS = 16;
constexpr int MAX_THREADS = 64 + 32 * MAX_EPI_WARPS;
// fused-attention variant: 16 epilogue warps (four per TMEM lane quarter).  The per-(sequence, head) attention is a long
// dependent chain of mma.sync / shuffles / exp (measured ~12k cycles per unit on one warp), so 8 units run per CTA at a
// time and each unit is split between TWO warps (one 16-row half of the 32x32 products each).
constexpr int ATT_EPI_WARPS = 16;
constexpr int ATT_THREADS = 64 + 32 * ATT_EPI_WARPS;
constexpr int TMEM_COLS = 512;
constexpr uint32_t SPIN_LIMIT = 1u << 27;

struct TParams {
  float* C; int ldc;
  int M, N, K;
  int BN;              // UMMA N of a tile (multiple of 16, <= 256)
  int b_rows;          // B rows (K-major) or padded columns (MN-major, multiple of 32) staged per tile
  int stages;
  int tiles_m, tiles_n, splitk;
  int ksteps_total, ksteps_per_split;
  int out_mode;        // 0 store, 1 load-add-store (beta = 1), 2 red.add (split-K)
  float alpha;
  GemmEpilogue epi;    // optional fused epilogue (out_mode 0 only)
  int rv_smem;         // 1: the rowvec rows of each epilogue warp are staged in shared memory (L >= 16)
  int c_tma;           // 1: C has a tensor map -> the epilogue stores through shared memory + TMA
  int tile_rows;       // matrix rows between consecutive m-tiles of one CTA (BM * MT; fused attention: spt * L <= 128)
  int epi_warps;       // epilogue warps of the plain GEMM: 4, 8 or 16 (1, 2 or 4 per TMEM lane quarter)
  int stg_bufs;        // store staging tiles per epilogue warp: 2, or 1 when 16 warps share the shared memory
  // ---- fused QKV projection + attention epilogue (ADH > 0): the tile is spt whole sequences x HPT whole heads
  int att_L, att_spt, att_nh, att_nseq;
  float* att_qkv_t;    // [n_seq, nh, 3, 32, ST] tf32 Q|K|V tiles saved for the backward pass (NULL: inference)
  float* att_y;        // [n_seq * L, nh * DH] attention output (dropout-masked, tf32-rounded)
  Dropout att_drop;
};
// fused-attention geometry for head dim DH: HR heads are processed per epilogue round, HPT heads per tile
template <int DH> struct AttGeo {
  static constexpr int HPT = (DH <= 20) ? 4 : 2;
  static constexpr int HR = 2;
  static constexpr int ROUNDS = HPT / HR;
  static constexpr int BN = HPT * 3 * DH;                 // 240 (dh 20), 192 (16), 144 (24), 192 (32)
  static constexpr int RCOLS = HR * 3 * DH;               // accumulator columns per round (multiple of 8)
  static constexpr int MAT = att::Cfg<DH>::MAT;
  static constexpr int SPT_MAX = 4;                       // sequences per 128-row tile handled (one per epilogue warp)
  static constexpr uint32_t TILE_BYTES = SPT_MAX * HR * 3 * MAT * 4;
  static constexpr uint32_t SMEM = TILE_BYTES;             // (the 32x32 probability tile overlays the unit's own Q | K)
  static_assert(2 * MAT >= att::LP * att::PS, "the score tile overlays Q and K");
};
constexpr int RV_ART = 4;                   // articles a warp's 32 rows can span when L >= 16
constexpr int RV_WARP_FLOATS = RV_ART * 256; // per warp, per m-subtile

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
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
// The single producer / MMA threads share their SM sub-partitions with epilogue warps: after a few failed polls
// they back off, so that a long wait (epilogue-bound GEMMs, a full stage ring) does not burn the issue slots the
// epilogue needs.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > 4) __nanosleep(40);
    if (spins > SPIN_LIMIT) __trap();  // a protocol bug traps instead of hanging the GPU
  }
}
__device__ __forceinline__ void mbar_wait_backoff(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    __nanosleep(64);
    if (++spins > SPIN_LIMIT) __trap();
  }
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
      "l"(map), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}

// PAIR mode: the copy lands in THIS CTA's shared memory, its completion is counted on `bar`, a
// shared::cluster address that may belong to the peer (leader) CTA
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
      "l"(map), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
// shared::cluster address of the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// Arrive on a barrier of another CTA of the cluster.  Only used to hand an accumulator buffer (TMEM) back to the MMA warp:
// that hand-over is ordered by the tcgen05 fences on both sides, no generic-memory data travels with it, so the default
// .release.cta form is enough (a .release.cluster arrive costs MEMBAR.ALL.GPU + ERRBAR per warp per tile).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// cta_group::2 commit: arrives on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void umma_tf32_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

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
// UMMA shared-memory descriptor (cute::UMMA::SmemDescriptor): start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version=1 [46,48), layout type [61,64) (2 = SWIZZLE_128B, 1 = SWIZZLE_128B_BASE32B).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                              uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout_type << 61;
  return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D=f32 [4,6)=1, A,B=tf32 [7,10)=[10,13)=2,
// a_major bit15, b_major bit16 (1 = MN-major), N>>3 [17,23), M>>4 [24,29).
__device__ __forceinline__ uint32_t make_idesc(bool a_mn, bool b_mn, int n, int m = BM) {
  uint32_t d = 0;
  d |= 1u << 4;
  d |= 2u << 7;
  d |= 2u << 10;
  d |= (a_mn ? 1u : 0u) << 15;
  d |= (b_mn ? 1u : 0u) << 16;
  d |= (uint32_t)(n >> 3) << 17;
  d |= (uint32_t)(m >> 4) << 24;
  return d;
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

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void epi_bar_sync() {   // the epilogue warps of the fused-attention variant
  asm volatile("bar.sync 1, %0;" ::"n"(32 * ATT_EPI_WARPS) : "memory");
}

// Walks this CTA's work items (m-tile, n-tile, k-split; n fastest so the CTAs that run at the same
// time share the A rows in L2) k-step by k-step.
template <int MT, bool PAIR>
struct Cursor {
  int item, ks, ks_end, m0, n0;
  int rm;  // rank of this CTA inside its pair (PAIR: an item is a tile of 2 x MT x 128 rows)
  __device__ __forceinline__ void load(const TParams& p) {
    const int per_m = p.tiles_n * p.splitk;
    const int tm = item / per_m, rem = item - tm * per_m;
    const int tn = rem / p.splitk, sp = rem - tn * p.splitk;
    m0 = (tm * (PAIR ? 2 : 1) + rm) * p.tile_rows;
    n0 = tn * p.BN;
    ks = sp * p.ksteps_per_split;
    ks_end = min(p.ksteps_total, ks + p.ksteps_per_split);
  }
  __device__ __forceinline__ void init(const TParams& p, int first, int n_items, int rm_) {
    item = first;
    rm = rm_;
    ks = ks_end = m0 = n0 = 0;
    if (item < n_items) load(p);
  }
  __device__ __forceinline__ bool valid(int n_items) const { return item < n_items; }
  __device__ __forceinline__ bool advance(const TParams& p, int n_items, int stride) {  // true: item finished
    if (++ks < ks_end) return false;
    item += stride;
    if (item < n_items) load(p);
    return true;
  }
  __device__ __forceinline__ void next_item(const TParams& p, int n_items, int stride) {
    item += stride;
    if (item < n_items) load(p);
  }
};

// Per-row operands of the fused epilogue, fetched before the accumulator is ready.
struct EpiRow {
  float rs;          // rowscale[row]
  const float* rv;   // rowvec row of this thread (shared-memory slab or global), indexed by tile column
  uint32_t rv_sa;    // shared-state-space address of that row when it lives in the slab (p.rv_smem)
};

// fused epilogue on 16 consecutive columns [n0 + c0, n0 + c0 + 16) of one row (see GemmEpilogue)
__device__ __forceinline__ void epi_apply16(const TParams& p, const EpiRow& er, float* v, int row, int n0, int c0) {
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] *= p.alpha;
  if (p.epi.rowscale != nullptr && er.rv != nullptr) {
    // + rowscale[row] * rowvec[row / L, col]   (AttLayer2 backward: w_t * d_out[n, :])
    if (p.rv_smem) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float4 x;
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w)
                     : "r"(er.rv_sa + (uint32_t)(c0 + 4 * q) * 4u));
        v[q * 4] = fmaf(er.rs, x.x, v[q * 4]); v[q * 4 + 1] = fmaf(er.rs, x.y, v[q * 4 + 1]);
        v[q * 4 + 2] = fmaf(er.rs, x.z, v[q * 4 + 2]); v[q * 4 + 3] = fmaf(er.rs, x.w, v[q * 4 + 3]);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (n0 + c0 + i < p.N) v[i] = fmaf(er.rs, __ldg(er.rv + c0 + i), v[i]);
    }
  }
  if (p.epi.drop.on()) {
    // inverted-dropout mask and scale of element (row, col): index row * drop_ld + col
    const uint64_t g0 = ((uint64_t)row * (uint64_t)p.epi.drop_ld + (uint64_t)(n0 + c0)) >> 2;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 f = p.epi.drop.factor4_group(g0 + q);
      v[q * 4] *= f.x; v[q * 4 + 1] *= f.y; v[q * 4 + 2] *= f.z; v[q * 4 + 3] *= f.w;
    }
  }
  if (p.epi.round_out) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = round_tf32_bits(v[i]);
  }
}

// row-per-thread global stores of 16 columns (no tensor map for C, or the 16-column tail of a tile)
__device__ __forceinline__ void store16_direct(const TParams& p, float* crow, const float* v, int n0, int c0, bool vec_ok,
                                               bool vec8_ok) {
  if (vec8_ok && n0 + c0 + 15 < p.N) {
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(crow + c0), "f"(v[0]), "f"(v[1]), "f"(v[2]),
                 "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
                 : "memory");
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(crow + c0 + 8), "f"(v[8]), "f"(v[9]),
                 "f"(v[10]), "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15])
                 : "memory");
    return;
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
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(o.x), "f"(o.y), "f"(o.z), "f"(o.w)
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

template <bool A_MN, bool B_MN, int MT, bool PAIR, int ADH>
__global__ void __launch_bounds__(ADH > 0 ? ATT_THREADS : MAX_THREADS, 1)
    gemm_tma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmC, const TParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[MAX_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[MAX_STAGES];
  __shared__ __align__(8) uint64_t tfull_bar[2];
  __shared__ __align__(8) uint64_t tempty_bar[2];
  __shared__ uint32_t tmem_slot;
  constexpr int NBUF = MT == 1 ? 2 : 1;  // accumulator buffers

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int S = p.stages;
  const int BN = p.BN;
  constexpr uint32_t a_bytes = (uint32_t)MT * BM * BK * 4;
  const uint32_t b_bytes = (uint32_t)p.b_rows * BK * 4;          // multiple of 1024 (b_rows % 8 == 0)
  const uint32_t stage_bytes = a_bytes + b_bytes;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int n_items = p.tiles_m * p.tiles_n * p.splitk;
  constexpr int csize = PAIR ? 2 : 1;
  const int rm = PAIR ? (int)cluster_ctarank() : 0;   // 0 = leader (issues the MMAs of the pair)
  const int first_item = (int)blockIdx.x / csize, item_stride = (int)gridDim.x / csize;

  if (tid == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&tfull_bar[b]), 1);
      // PAIR: both CTAs' epilogues release the leader
      mbar_init(smem_u32(&tempty_bar[b]), (uint32_t)((ADH > 0 ? 32 * ATT_EPI_WARPS : 32 * p.epi_warps) * csize));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    if (p.c_tma) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmC) : "memory");
  }
  if (warp == 1) {
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)),
                   "n"(TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)),
                   "n"(TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();  // every barrier of the cluster exists before anyone signals a peer
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer (one thread) =====================
    if (lane == 0) {
      Cursor<MT, PAIR> cu;
      cu.init(p, first_item, n_items, rm);
      int s = 0;
      uint32_t ph = 0;
      while (cu.valid(n_items)) {
        const int k0 = cu.ks * BK;
        mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1u);
        const uint32_t sa = smem_base + (uint32_t)s * stage_bytes;
        const uint32_t sb = sa + a_bytes;
        if (!PAIR) {
          const uint32_t bar = smem_u32(&full_bar[s]);
          mbar_expect_tx(bar, stage_bytes);
          if (!A_MN) {
            tma_load_2d(sa, &tmA, k0, cu.m0, bar);                       // box {32 k, 128*MT rows}
          } else {
#pragma unroll
            for (int g = 0; g < 4 * MT; ++g) tma_load_2d(sa + (uint32_t)g * 4096u, &tmA, cu.m0 + 32 * g, k0, bar);
          }
          if (!B_MN) {
            tma_load_2d(sb, &tmB, k0, cu.n0, bar);                       // box {32 k, b_rows rows}
          } else {
            const int groups = p.b_rows >> 5;
            for (int g = 0; g < groups; ++g) tma_load_2d(sb + (uint32_t)g * 4096u, &tmB, cu.n0 + 32 * g, k0, bar);
          }
        } else {
          // both CTAs' bytes are counted on the LEADER's barrier; this CTA stages its own A rows and its
          // half of the B tile (columns n0 + rm * BN/2 ...) in its own shared memory
          const uint32_t bar = mapa_u32(smem_u32(&full_bar[s]), 0);
          if (rm == 0) mbar_expect_tx(smem_u32(&full_bar[s]), 2u * stage_bytes);
          const int nb0 = cu.n0 + rm * (BN >> 1);
          if (!A_MN) {
            tma_load_2d_2sm(sa, &tmA, k0, cu.m0, bar);
          } else {
#pragma unroll
            for (int g = 0; g < 4 * MT; ++g) tma_load_2d_2sm(sa + (uint32_t)g * 4096u, &tmA, cu.m0 + 32 * g, k0, bar);
          }
          if (!B_MN) {
            tma_load_2d_2sm(sb, &tmB, k0, nb0, bar);                     // box {32 k, b_rows = BN/2 rows}
          } else {
            const int groups = p.b_rows >> 5;
            for (int g = 0; g < groups; ++g) tma_load_2d_2sm(sb + (uint32_t)g * 4096u, &tmB, nb0 + 32 * g, k0, bar);
          }
        }
        cu.advance(p, n_items, item_stride);
        if (++s == S) {
          s = 0;
          ph ^= 1u;
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0 && rm == 0) {   // PAIR: only the leader CTA issues
      const uint32_t idesc = make_idesc(A_MN, B_MN, BN, PAIR ? 2 * BM : BM);
      constexpr uint32_t a_lbo = A_MN ? 4096u : 16u, a_sbo = A_MN ? 512u : 1024u, a_lt = A_MN ? 1u : 2u;
      constexpr uint32_t b_lbo = B_MN ? 4096u : 16u, b_sbo = B_MN ? 512u : 1024u, b_lt = B_MN ? 1u : 2u;
      constexpr uint32_t a_kadv = A_MN ? 1024u : 32u, b_kadv = B_MN ? 1024u : 32u;  // bytes per UMMA_K
      Cursor<MT, PAIR> cu;
      cu.init(p, first_item, n_items, rm);
      int t = 0, s = 0;
      uint32_t ph = 0;
      while (cu.valid(n_items)) {
        const int buf = NBUF == 2 ? (t & 1) : 0, use = NBUF == 2 ? (t >> 1) : t;
        mbar_wait_backoff(smem_u32(&tempty_bar[buf]), (uint32_t)((use & 1) ^ 1));  // epilogue drained this buffer
        tc_fence_after();
        const uint32_t tacc = tmem + (NBUF == 2 ? (uint32_t)buf * 256u : 0u);
        bool first = true, last = false;
        while (!last) {
          mbar_wait(smem_u32(&full_bar[s]), ph);
          tc_fence_after();
          const uint32_t sa = smem_base + (uint32_t)s * stage_bytes;
          const uint32_t sb = sa + a_bytes;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t db = make_desc(sb + (uint32_t)k * b_kadv, b_lbo, b_sbo, b_lt);
            const uint32_t acc = (first && k == 0) ? 0u : 1u;
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
              const uint32_t a_addr = sa + (uint32_t)mt * (uint32_t)(BM * BK * 4) + (uint32_t)k * a_kadv;
              if (PAIR) umma_tf32_2sm(tacc + (uint32_t)mt * 256u, make_desc(a_addr, a_lbo, a_sbo, a_lt), db, idesc, acc);
              else umma_tf32(tacc + (uint32_t)mt * 256u, make_desc(a_addr, a_lbo, a_sbo, a_lt), db, idesc, acc);
            }
          }
          // frees the stage when these MMAs retire (PAIR: in both CTAs)
          if (PAIR) umma_commit_2sm(smem_u32(&empty_bar[s]));
          else umma_commit(smem_u32(&empty_bar[s]));
          first = false;
          last = cu.advance(p, n_items, item_stride);
          if (++s == S) {
            s = 0;
            ph ^= 1u;
          }
        }
        // accumulator(s) of this item complete (PAIR: each CTA's epilogue drains its own 128 x MT rows)
        if (PAIR) umma_commit_2sm(smem_u32(&tfull_bar[buf]));
        else umma_commit(smem_u32(&tfull_bar[buf]));
        ++t;
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue (warps 2-5; TMEM lane quarter = warp & 3) =====================
    // TMEM -> registers -> (fused ops) -> 32x32 swizzled staging tile in shared memory -> TMA store
    // (cp.async.bulk.tensor, or its .add reduction for beta = 1 / split-K).  A thread owns one accumulator
    // ROW, so direct global stores touch 32 different cache lines per instruction -- measured, they kept the
    // LSU busy ~8k cycles per 256x240 tile with the tensor pipe idle; the TMA store writes whole 128-byte
    // lines and clips at the matrix edge by itself.
    const int ew = warp & 3;
    Cursor<MT, PAIR> cu;
    cu.init(p, first_item, n_items, rm);
    int t = 0;
    if constexpr (ADH > 0) {
      // ============ fused attention epilogue (north star: projection -> per-head softmax(QK^T/sqrt(dh))^T V) ============
      // The accumulator tile is spt whole sequences x HPT whole heads, columns ordered (head, Q|K|V, d) by the
      // permuted weight operand.  Per round of HR heads: every epilogue warp moves its 32 accumulator rows
      // TMEM -> registers -> [seq][head][Q|K|V][32 tok][ST] tiles in shared memory (tf32-rounded), then warp w runs the
      // attention of sequence w on mma.sync and writes dropout(Y) rows; when training, the Q|K|V tiles go to global
      // memory as they are (one 3*MAT bulk copy per (sequence, head)) for the backward kernel.
      using G = AttGeo<ADH>;
      using AC = att::Cfg<ADH>;
      static_assert(MT == 1, "fused attention uses 128-row tiles with a double-buffered accumulator");
      constexpr int DH = ADH, ST = AC::ST, MAT = AC::MAT;
      const int lane_g = lane >> 2, lane_t = lane & 3;
      const int L = p.att_L, spt = p.att_spt, nh = p.att_nh, D = nh * DH;
      float* tiles = reinterpret_cast<float*>(smem_raw + (smem_base - smem_u32(smem_raw)) + (size_t)S * stage_bytes);
      const float inv = rsqrtf((float)DH);
      const int eg = (warp - 2) >> 2;                          // 0..3: the four warps of a TMEM lane quarter
      const int etid = (warp - 2) * 32 + lane;
      for (int i = etid; i < (int)(G::TILE_BYTES / 16); i += 32 * ATT_EPI_WARPS)   // zero rows past L (and all padding) once
        reinterpret_cast<float4*>(tiles)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      const int r_loc = ew * 32 + lane;                      // accumulator row of this thread inside the tile
      const int seq_loc = r_loc / L, tok = r_loc - seq_loc * L;
      const bool row_ok = seq_loc < spt;
      // compute: unit = (sequence s_loc of the tile, head my_hh of the round), shared by the two warps half = 0 / 1
      const int s_loc = (warp - 2) & 3, my_hh = eg & 1, half = eg >> 1;
      const int pair_bar = 2 + s_loc * 2 + my_hh;              // named barrier of the unit's two warps (ids 2..9)
      static_assert(G::HR == 2 && ATT_EPI_WARPS == 16, "two warps per (sequence, head) unit and round");
      while (cu.valid(n_items)) {
        const int buf = t & 1, use = t >> 1;
        const int seq0 = cu.m0 / L;                          // first sequence of this CTA's tile
        const int head0 = cu.n0 / (3 * DH);
        mbar_wait_backoff(smem_u32(&tfull_bar[buf]), (uint32_t)(use & 1));
        tc_fence_after();
        const uint32_t tbase = tmem + ((uint32_t)(ew * 32) << 16) + (uint32_t)buf * 256u;
#pragma unroll 1
        for (int rd = 0; rd < G::ROUNDS; ++rd) {
          epi_bar_sync();   // previous round's attention (and its bulk copies) are done with the tiles in every warp
          float* rowbase = tiles + (size_t)seq_loc * (G::HR * 3 * MAT) + tok * ST;
          // the four warps of a lane quarter split the round's columns in 8-column chunks (c mod 4)
#pragma unroll
          for (int c = 0; c < G::RCOLS / 8; ++c) {
            if ((c & 3) != eg) continue;                       // warp-uniform
            float v[8];
            tmem_ld8(tbase + (uint32_t)(rd * G::RCOLS + c * 8), v);   // warp-collective
            if (row_ok) {
#pragma unroll
              for (int j = 0; j < 2; ++j) {
                const int col = c * 8 + j * 4;                 // compile-time after unrolling
                const int hh = col / (3 * DH), rem = col - hh * 3 * DH, m = rem / DH, d = rem - m * DH;
                float4 o = make_float4(round_tf32_bits(v[j * 4] * p.alpha), round_tf32_bits(v[j * 4 + 1] * p.alpha),
                                       round_tf32_bits(v[j * 4 + 2] * p.alpha), round_tf32_bits(v[j * 4 + 3] * p.alpha));
                *reinterpret_cast<float4*>(rowbase + (hh * 3 + m) * MAT + d) = o;
              }
            }
          }
          if (rd == G::ROUNDS - 1) {
            // every accumulator column of this tile has been read: hand the TMEM buffer back to the MMA warp
            tc_fence_before();
            if (PAIR) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty_bar[buf]), 0));
            else mbar_arrive(smem_u32(&tempty_bar[buf]));
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // tiles are read by bulk copies below
          epi_bar_sync();   // all rows of every sequence are in place
          const long seq = (long)seq0 + s_loc;
          const int head = head0 + rd * G::HR + my_hh;
          if (s_loc < spt && seq < p.att_nseq && head < nh) {
            float* Qs = tiles + (size_t)(s_loc * G::HR + my_hh) * 3 * MAT;
            const float* Ks = Qs + MAT;
            const float* Vs = Ks + MAT;
            const bool save = p.att_qkv_t != nullptr && half == 0;
            if (save && lane == 0) {   // Q | K | V tiles of this (sequence, head) -> HBM for the backward pass
              float* dst = p.att_qkv_t + ((size_t)seq * nh + head) * 3 * MAT;
              asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(Qs)),
                           "r"((uint32_t)(3 * MAT * 4))
                           : "memory");
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            float acc[4][4];
            att::gemm_xyT_half<DH>(acc, Qs, Ks, half, lane_g, lane_t);   // query rows half*16 .. +15
            att::softmax_rows_half(acc, inv, L, lane_t);
            if (save && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            // both warps are done reading Q / K, and the bulk copy has read them: the probabilities overlay them
            asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
            float* Ps = Qs;
            att::store_frag_half(Ps, acc, half, lane_g, lane_t);
            asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
            float o[AC::NT][4];
            att::gemm_smemT_half<DH>(o, Ps, Vs, half, lane_g, lane_t);   // O[k, d] = sum_q P[q, k] V[q, d], k rows half*16 ..
            float* out = p.att_y + seq * L * D + head * DH;
#pragma unroll
            for (int nt = 0; nt < AC::NT; ++nt) {
              const int col = nt * 8 + 2 * lane_t;
              if (col >= DH) continue;
#pragma unroll
              for (int hf = 0; hf < 2; ++hf) {
                const int r = half * 16 + lane_g + hf * 8;
                if (r >= L) continue;
                float2 v = make_float2(o[nt][hf * 2], o[nt][hf * 2 + 1]);
                if (p.att_drop.on()) {  // AttLayer2 only ever reads dropout(y): store it masked and scaled
                  const uint64_t idx = (uint64_t)(seq * L + r) * (uint64_t)D + (uint64_t)(head * DH + col);
                  const float4 f = p.att_drop.factor4_group(idx >> 2);
                  v.x *= (idx & 2ull) ? f.z : f.x;
                  v.y *= (idx & 2ull) ? f.w : f.y;
                }
                *reinterpret_cast<uint2*>(out + (long)r * D + col) =
                    make_uint2(__float_as_uint(round_tf32_bits(v.x)), __float_as_uint(round_tf32_bits(v.y)));
              }
            }
            // (rows past L of the overlaid Q / K tiles now hold probabilities: finite values, which is all the
            // masked softmax columns and the zero V rows need; the backward re-zeroes them when it loads the tiles)
          }
        }
        cu.next_item(p, n_items, item_stride);
        ++t;
      }
      if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      __syncwarp();
    } else {
    const bool vec_base = ((p.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0);
    const bool vec8_base = vec_base && ((p.ldc & 7) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 31) == 0) && p.out_mode == 0;
    uint8_t* tail = smem_raw + (smem_base - smem_u32(smem_raw)) + (size_t)S * stage_bytes;   // behind the stage ring
    // rowvec slab of this TMEM lane quarter (the 32 rows its EG warps share): [MT][RV_ART][BN] floats
    const int EG = p.epi_warps >> 2, eg = (warp - 2) >> 2;   // warps per TMEM lane quarter, this warp's index among them
    const uint32_t stg_warp = (uint32_t)p.stg_bufs * 4096u;
    float* rv_s = reinterpret_cast<float*>(tail + (size_t)p.epi_warps * stg_warp) + (size_t)ew * MT * RV_WARP_FLOATS;
    // the EG warps of a quarter meet on named barrier 1 + ew around their use of the slab
    auto quarter_sync = [&]() {
      if (EG > 1) asm volatile("bar.sync %0, %1;" ::"r"(1 + ew), "r"(32 * EG) : "memory");
      else __syncwarp();
    };
    // store staging of this warp: stg_bufs x [32 rows][128 B], SWIZZLE_128B
    const uint32_t stg = smem_base + (uint32_t)S * stage_bytes + (uint32_t)(warp - 2) * stg_warp;
    int nstore = 0;  // TMA stores issued by this warp (lane 0 tracks the bulk groups)
    // the rowvec slab is filled with 16-byte cp.async when every address involved is 16-byte aligned (tile origins are
    // multiples of 16 columns) and N is a multiple of 4, so that no float4 straddles the last column
    const bool rv_vec = p.rv_smem && p.epi.rowvec != nullptr && (p.epi.rowvec_ld & 3) == 0 && (p.N & 3) == 0 && (BN & 3) == 0 &&
                        (reinterpret_cast<uintptr_t>(p.epi.rowvec) & 15) == 0;
    while (cu.valid(n_items)) {
      const int buf = NBUF == 2 ? (t & 1) : 0, use = NBUF == 2 ? (t >> 1) : t;
      const int n0 = cu.n0;
      EpiRow er[MT];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        er[mt].rs = 0.0f;
        er[mt].rv = nullptr;
        er[mt].rv_sa = 0u;
      }
      if (p.epi.rowscale != nullptr) {
        // operands of the fused epilogue that do not depend on the accumulator: fetched BEFORE waiting
        // for the MMAs, so their latency hides behind the main loop
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          const int row_lo = cu.m0 + mt * BM + ew * 32, row = row_lo + lane;
          er[mt].rs = row < p.M ? __ldg(p.epi.rowscale + row) : 0.0f;
          const int art0 = row_lo / p.epi.L;
          if (p.rv_smem) {
            const int art_last = (min(row_lo + 31, p.M - 1)) / p.epi.L;
            if (rv_vec) {
              // 16-byte asynchronous copies, none waited for here: the slab lands while this warp waits for the MMAs
              const int BN4 = BN >> 2;
              for (int i = eg * 32 + lane; i < RV_ART * BN4; i += 32 * EG) {
                const int a = i / BN4, c = (i - a * BN4) << 2, art = art0 + a;
                const bool ok = art <= art_last && n0 + c < p.N;
                const float* src = ok ? p.epi.rowvec + (long)art * p.epi.rowvec_ld + n0 + c : p.epi.rowvec;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(rv_s + (mt * RV_ART + a) * BN + c)),
                             "l"(src), "r"(ok ? 16u : 0u)
                             : "memory");
              }
            } else {
              for (int a = 0; a < RV_ART; ++a) {
                const int art = art0 + a;
                for (int c = eg * 32 + lane; c < BN; c += 32 * EG) {
                  float x = 0.0f;
                  if (art <= art_last && n0 + c < p.N) x = __ldg(p.epi.rowvec + (long)art * p.epi.rowvec_ld + n0 + c);
                  rv_s[(mt * RV_ART + a) * BN + c] = x;
                }
              }
            }
            er[mt].rv = rv_s + (mt * RV_ART + (min(row, p.M - 1) / p.epi.L - art0)) * BN;
            er[mt].rv_sa = smem_u32(er[mt].rv);
          } else if (row < p.M) {
            er[mt].rv = p.epi.rowvec + (long)(row / p.epi.L) * p.epi.rowvec_ld + n0;
          }
        }
        if (rv_vec) asm volatile("cp.async.commit_group;" ::: "memory");
      }
      mbar_wait_backoff(smem_u32(&tfull_bar[buf]), (uint32_t)(use & 1));
      tc_fence_after();
      if (p.rv_smem) {
        if (rv_vec) asm volatile("cp.async.wait_group 0;" ::: "memory");
        quarter_sync();   // every warp's share of the slab has landed
      }
      const bool vec_ok = vec_base && ((n0 & 3) == 0);
      const bool vec8_ok = vec8_base && ((n0 & 7) == 0);
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        const int row_lo = cu.m0 + mt * BM + ew * 32, row = row_lo + lane;
        const uint32_t tbase = tmem + ((uint32_t)(ew * 32) << 16) + (NBUF == 2 ? (uint32_t)buf * 256u : (uint32_t)mt * 256u);
        int c0 = 0;
        if (p.c_tma) {
          for (; c0 + 32 <= BN && n0 + c0 < p.N; c0 += 32) {
            if ((((c0 >> 5) + t) % EG) != eg) continue;   // the warps of a lane quarter take the 32-column chunks in turn, rotating per tile
            float v[32];
            tmem_ld16(tbase + (uint32_t)c0, v);
            tmem_ld16(tbase + (uint32_t)c0 + 16u, v + 16);
            epi_apply16(p, er[mt], v, row, n0, c0);
            epi_apply16(p, er[mt], v + 16, row, n0, c0 + 16);
            const uint32_t sb = stg + (p.stg_bufs == 2 ? (uint32_t)(nstore & 1) * 4096u : 0u);
            if (nstore >= p.stg_bufs) {  // the store that used this staging tile before must have read it
              if (lane == 0) {
                if (p.stg_bufs == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
              }
              __syncwarp();
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const uint32_t dst = sb + (uint32_t)lane * 128u + (uint32_t)((j ^ (lane & 7)) << 4);
              asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(dst), "f"(v[4 * j]), "f"(v[4 * j + 1]),
                           "f"(v[4 * j + 2]), "f"(v[4 * j + 3])
                           : "memory");
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
              if (p.out_mode == 0)
                asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(&tmC),
                             "r"(n0 + c0), "r"(row_lo), "r"(sb)
                             : "memory");
              else
                asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(&tmC),
                             "r"(n0 + c0), "r"(row_lo), "r"(sb)
                             : "memory");
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            ++nstore;
          }
        }
        // direct path: everything when C has no tensor map, else the 16-column tail of a tile
        float* crow = p.C + (long)row * p.ldc + n0;
        for (; c0 < BN; c0 += 16) {
          if (n0 + c0 >= p.N) break;  // warp-uniform
          if ((((c0 >> 4) + t + 3) % EG) != eg) continue;
          float v[16];
          tmem_ld16(tbase + (uint32_t)c0, v);  // warp-collective
          if (row < p.M) {
            epi_apply16(p, er[mt], v, row, n0, c0);
            store16_direct(p, crow, v, n0, c0, vec_ok, vec8_ok);
          }
        }
      }
      tc_fence_before();
      // buffer may be overwritten by the MMA warp (PAIR: the leader's, which waits for both epilogues)
      if (PAIR) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty_bar[buf]), 0));
      else mbar_arrive(smem_u32(&tempty_bar[buf]));
      if (p.rv_smem) quarter_sync();            // every reader of the rowvec slab is done with it
      else __syncwarp();
      cu.next_item(p, n_items, item_stride);
      ++t;
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // staging must outlive its readers
    __syncwarp();
    }   // plain GEMM epilogue
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();  // no CTA leaves while its peer may still signal its barriers or read its operands
  if (warp == 1) {
    tc_fence_after();
    if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TMEM_COLS) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TMEM_COLS) : "memory");
  }
}

__global__ void zero_c_kernel(float* C, int ldc, int M, int N) {
  long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i < (long)M * N) C[(i / N) * (long)ldc + (i % N)] = 0.0f;
}

// ---- host side: tensor maps --------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
    (void)cudaGetLastError();
  }
  return fn;
}

struct MapKey {
  const void* ptr; int inner, outer, ld, box_outer, mn;
  bool operator==(const MapKey& o) const {
    return ptr == o.ptr && inner == o.inner && outer == o.outer && ld == o.ld && box_outer == o.box_outer && mn == o.mn;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.ptr);
    auto mix = [&h](size_t v) { h ^= v + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2); };
    mix((size_t)k.inner); mix((size_t)k.outer); mix((size_t)k.ld); mix((size_t)k.box_outer); mix((size_t)k.mn);
    return h;
  }
};
thread_local std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_maps;

// 2-D map of a row-major fp32 matrix [outer, inner] (row stride ld floats); box {32, box_outer}.
// mn = 1 selects the 32-byte-atom swizzle of MN-major 32-bit UMMA operands.
int get_map(const float* ptr, int inner, int outer, int ld, int box_outer, bool mn, CUtensorMap* out) {
  const MapKey key{ptr, inner, outer, ld, box_outer, mn ? 1 : 0};
  auto it = g_maps.find(key);
  if (it != g_maps.end()) {
    *out = it->second;
    return EBK_OK;
  }
  EncodeTiledFn enc = encode_fn();
  if (enc == nullptr) {
    set_error("gemm_tma: cuTensorMapEncodeTiled is not available from the driver");
    return EBK_ERR_UNSUPPORTED;
  }
  const cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * 4ull};
  const cuuint32_t box[2] = {32u, (cuuint32_t)box_outer};
  const cuuint32_t estr[2] = {1u, 1u};
  CUtensorMap m;
  const CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE,
                         mn ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("gemm_tma: cuTensorMapEncodeTiled failed (%d) for [%d x %d] ld %d box %d", (int)r, outer, inner, ld, box_outer);
    return EBK_ERR_CUDA;
  }
  if (g_maps.size() > 512) g_maps.clear();
  g_maps.emplace(key, m);
  *out = m;
  return EBK_OK;
}

int g_sms = 0;

}  // namespace

bool gemm_tma_eligible(const float* A, int lda, const float* B, int ldb, int M, int N, int K) {
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  return al(A) && al(B) && (lda % 4 == 0) && (ldb % 4 == 0) && M >= 1 && N >= 1 && K >= 1 && encode_fn() != nullptr;
}

// Operands must already hold tf32-representable values (the tensor core truncates the low 13 mantissa
// bits of whatever it is given).  transA: A is stored [K, M]; transB: B is stored [N, K].
// tall: 1 selects 256-row tiles (MT = 2), 0 128-row tiles, -1 picks by problem size.
// cluster: 0 one CTA per tile, 1 CTA pairs (cta_group::2 MMA, each CTA stages half of B), -1 picks by problem size.
int gemm_tma(const float* A, int lda, bool transA, const float* B, int ldb, bool transB, float* C, int ldc, int M,
             int N, int K, float beta, float alpha, cudaStream_t st, int tall, const GemmEpilogue* epi, int cluster) {
  if (M <= 0 || N <= 0) return EBK_OK;
  EBK_CHECK_ARG(A && B && C && K >= 1, "gemm_tma: null operand or K < 1");
  EBK_CHECK_ARG(gemm_tma_eligible(A, lda, B, ldb, M, N, K), "gemm_tma: operands must be 16-byte aligned with ld %% 4 == 0");
  EBK_CHECK_ARG(beta == 0.0f || beta == 1.0f, "gemm_tma: beta must be 0 or 1");
  if (g_sms == 0) {
    int dev = 0;
    EBK_CUDA(cudaGetDevice(&dev));
    EBK_CUDA(cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const bool a_mn = transA, b_mn = !transB;
  static const int env_cluster = getenv("EBK_GEMM_CLUSTER") ? atoi(getenv("EBK_GEMM_CLUSTER")) : -2;  // experiments
  if (env_cluster >= -1) cluster = env_cluster;
  int bn_max = 256;
  if (tall < 0) {
    // auto: 256-row tiles halve the L2 traffic of B; use them when they still fill the machine.  A problem with fewer
    // 128 x 256 tiles than SMs is latency bound (pipeline fill + one epilogue per CTA: ~19 us for 256 x 256 tiles
    // whatever the size): it gets 128 x <=128 tiles -- more CTAs, 5 stages instead of 2, a quarter of the epilogue each.
    const long tn = ceil_div(N, ceil_div(ceil_div(N, ceil_div(N, 256)), 16) * 16);
    const long t1 = (long)ceil_div(M, BM) * tn, t2 = (long)ceil_div(M, 2 * BM) * tn;
    static const bool small_on = !(getenv("EBK_GEMM_SMALL_TILES") && atoi(getenv("EBK_GEMM_SMALL_TILES")) == 0);
    // (few tiles but a huge K -- the QKV weight gradient -- is not small: it keeps big tiles and splits K)
    if (t1 < g_sms && t1 * (long)ceil_div(K, BK) < 128L * g_sms && small_on) {
      tall = 0;
      bn_max = 128;
    } else {
      tall = (t2 >= 2L * g_sms || t1 < g_sms) ? 1 : 0;
    }
  }
  const int MT = tall ? 2 : 1;
  // cluster: 0 = one CTA per tile, 1 = CTA pairs (cta_group::2), -1 auto.  Measured on B200 (tools/bench_gemm_tma.py):
  // pairs lift 128-row tiles to the speed of 256-row tiles (both halve the B traffic through shared memory) but
  // add nothing on top of them -- 620-660 TFLOP/s tf32 either way, ~90 % of half the sustained cuBLAS bf16 rate --
  // so auto only pairs up 128-row tiles of problems large enough to keep every pair busy.
  if (cluster < 0) cluster = (MT == 1 && (long)ceil_div(M, 2 * BM) * ceil_div(N, 256) >= 2L * (g_sms / 2)) ? 1 : 0;
  const bool pair = cluster >= 1;
  const int cm = pair ? 2 : 1;
  TParams p;
  p.C = C; p.ldc = ldc; p.M = M; p.N = N; p.K = K; p.alpha = alpha;
  p.epi = epi ? *epi : GemmEpilogue{nullptr, nullptr, 0, 1, Dropout{0, 0, 1.0f, nullptr}, 0, false};
  const bool has_epi = p.epi.rowscale != nullptr || p.epi.drop.on() || p.epi.round_out;
  EBK_CHECK_ARG(!has_epi || beta == 0.0f, "gemm_tma: a fused epilogue needs beta == 0");
  EBK_CHECK_ARG(!p.epi.drop.on() || p.epi.drop_ld % 4 == 0, "gemm_tma: dropout epilogue needs drop_ld %% 4 == 0");
  int ntn = ceil_div(N, bn_max);
  {
    // Split-K problems (few tiles, long K: the QKV weight gradient is 3 x 5 tiles of 256 x 240 over K = 192 000): the
    // split count is floor(SMs / tiles), so 15 tiles leave 13 of 148 SMs without a CTA.  One more tile column (3 x 6 tiles
    // of 256 x 208, 8 splits = 144 CTAs) is 10 % slower as a kernel on its own (0.70 -> 0.78 ms) but the train step, where
    // this GEMM runs on the side stream against the HBM-bound optimizer pass, is 0.07-0.1 ms FASTER (measured three times,
    // profiles/r02_bench_ab.md calls 21 / 22; narrower tiles -- 176, 160, 128 -- lose again).  Taken when it lifts the
    // CTA count above 95 % of the SMs.
    static const int env_bn_bigk = getenv("EBK_GEMM_BN_BIGK") ? atoi(getenv("EBK_GEMM_BN_BIGK")) : -1;   // experiments: 0 = off
    const long tm = ceil_div(M, BM * MT * (pair ? 2 : 1)), sms = g_sms / (pair ? 2 : 1);
    auto ctas = [&](int ntn_) {
      const int bn = ceil_div(ceil_div(N, ntn_), 16) * 16;
      const long t = tm * ceil_div(N, bn);
      long sk = t < sms ? sms / t : 1;
      const long maxsplit = ceil_div(K, BK) / 8;
      if (sk > maxsplit) sk = maxsplit;
      return t * (sk < 1 ? 1 : sk);
    };
    if (env_bn_bigk > 0 && K >= 100000) ntn = ceil_div(N, env_bn_bigk);
    else if (env_bn_bigk != 0 && bn_max == 256 && !has_epi && K >= 32768 && tm * ntn < sms && ctas(ntn) * 100 < sms * 95 &&
             ctas(ntn + 1) * 100 >= sms * 95)
      ntn += 1;
  }
  p.BN = ceil_div(ceil_div(N, ntn), 16) * 16;
  p.tiles_n = ceil_div(N, p.BN);
  const int bn_cta = pair ? p.BN / 2 : p.BN;                       // B columns staged by one CTA
  p.b_rows = b_mn ? ((bn_cta + 31) & ~31) : bn_cta;
  p.tiles_m = ceil_div(M, BM * MT * cm);
  p.tile_rows = BM * MT;
  p.att_L = p.att_spt = p.att_nh = p.att_nseq = 0;
  p.att_qkv_t = p.att_y = nullptr;
  p.att_drop = Dropout{0, 0, 1.0f, nullptr};
  p.ksteps_total = ceil_div(K, BK);
  const size_t stage_bytes = (size_t)MT * BM * BK * 4 + (size_t)p.b_rows * BK * 4;
  p.rv_smem = (p.epi.rowscale != nullptr && p.epi.L >= 16) ? 1 : 0;
  static const int env_ew = getenv("EBK_GEMM_EPI_WARPS") ? atoi(getenv("EBK_GEMM_EPI_WARPS")) : 0;   // experiments
  // a fused epilogue is ~20 dependent ALU instructions per element: 4 warps per scheduler hide that latency, 1 or 2 do not
  p.epi_warps = (env_ew == 4 || env_ew == 8 || env_ew == 16) ? env_ew
                : ((p.epi.rowscale != nullptr || p.epi.drop.on()) && MT == 1 && !pair) ? 16 : 4;
  p.stg_bufs = p.epi_warps == 16 ? 1 : 2;
  const size_t STG_BYTES = (size_t)p.epi_warps * p.stg_bufs * 4096;
  const size_t rv_bytes = p.rv_smem ? (size_t)4 * MT * RV_WARP_FLOATS * sizeof(float) : 0;   // one slab per TMEM lane quarter
  p.c_tma = (ldc % 4 == 0 && (reinterpret_cast<uintptr_t>(C) & 15) == 0) ? 1 : 0;
  const size_t budget = 226 * 1024 - 1024 - STG_BYTES - rv_bytes;   // 227 KB per CTA minus barriers / alignment slack
  int stages = (int)(budget / stage_bytes);
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  if (stages < 2) stages = 2;
  p.stages = stages;
  const int csize = cm;
  const int max_clusters = g_sms / csize;
  const long tiles = (long)p.tiles_m * p.tiles_n;         // work items (of a CTA or of a CTA pair)
  int splitk = 1;
  if (tiles < max_clusters && p.ksteps_total >= 16 && !has_epi) {
    splitk = (int)(max_clusters / tiles);  // fill the machine in ONE wave of equal items
    const int maxsplit = p.ksteps_total / (bn_max < 256 ? 16 : 8);   // (small problems: a split must be worth its zero-fill)
    if (splitk > maxsplit) splitk = maxsplit;
    if (splitk < 1) splitk = 1;
  }
  p.ksteps_per_split = ceil_div(p.ksteps_total, splitk);
  splitk = ceil_div(p.ksteps_total, p.ksteps_per_split);
  p.splitk = splitk;
  p.out_mode = splitk > 1 ? 2 : (beta != 0.0f ? 1 : 0);
  if (splitk > 1 && beta == 0.0f) {
    const long n = (long)M * N;
    zero_c_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(C, ldc, M, N);
    EBK_LAUNCH_CHECK();
  }
  CUtensorMap tmA, tmB;
  if (!a_mn) EBK_TRY(get_map(A, K, M, lda, BM * MT, false, &tmA));   // storage [M, K]
  else       EBK_TRY(get_map(A, M, K, lda, 32, true, &tmA));         // storage [K, M]
  if (!b_mn) EBK_TRY(get_map(B, K, N, ldb, p.b_rows, false, &tmB));  // storage [N, K]; PAIR: half of the tile's rows
  else       EBK_TRY(get_map(B, N, K, ldb, 32, true, &tmB));         // storage [K, N]
  const long n_items = tiles * splitk;
  const int n_clusters = (int)(n_items < max_clusters ? n_items : max_clusters);
  const int grid = n_clusters * csize;
  const size_t smem = (size_t)stages * stage_bytes + 1024 + STG_BYTES + rv_bytes;
  CUtensorMap tmC = tmA;
  if (p.c_tma) EBK_TRY(get_map(C, N, M, ldc, 32, false, &tmC));      // box {32 cols, 32 rows}, SWIZZLE_128B
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(64 + 32 * p.epi_warps);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)csize;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = csize > 1 ? 1 : 0;
#define LAUNCH4(AMN_, BMN_, MT_, PAIR_)                                                                          \
  {                                                                                                                \
    EBK_CUDA(cudaFuncSetAttribute(gemm_tma_kernel<AMN_, BMN_, MT_, PAIR_, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                  (int)smem));                                                                     \
    EBK_CUDA(cudaLaunchKernelEx(&cfg, gemm_tma_kernel<AMN_, BMN_, MT_, PAIR_, 0>, tmA, tmB, tmC, p));               \
  }
#define LAUNCH3(AMN_, BMN_, MT_)                   \
  {                                                \
    if (pair) LAUNCH4(AMN_, BMN_, MT_, true)       \
    else LAUNCH4(AMN_, BMN_, MT_, false)           \
  }
#define LAUNCH2(AMN_, BMN_)                  \
  {                                          \
    if (MT == 2) LAUNCH3(AMN_, BMN_, 2)      \
    else LAUNCH3(AMN_, BMN_, 1)              \
  }
  if (!a_mn && !b_mn) LAUNCH2(false, false)
  else if (!a_mn && b_mn) LAUNCH2(false, true)
  else if (a_mn && !b_mn) LAUNCH2(true, false)
  else LAUNCH2(true, true)
#undef LAUNCH2
#undef LAUNCH3
#undef LAUNCH4
  EBK_LAUNCH_CHECK();
  return EBK_OK;
}


// ---- fused QKV projection + attention (the forward of SelfAttention, layers.py:214-252, in ONE kernel) -------------
namespace {
// wp[k, (h, m, d)] = tf32(W[k, m * D + h * dh + d]): head-major column order, so that an accumulator tile of
// AttGeo::BN columns holds Q | K | V of whole heads
__global__ void permute_round_wqkv_kernel(const float* __restrict__ W, float* __restrict__ wp, int Din, int nh, int dh) {
  const int D = nh * dh, N = 3 * D;
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= (long)Din * N) return;
  const int k = (int)(i / N), c = (int)(i - (long)k * N);
  const int h = c / (3 * dh), rem = c - h * 3 * dh, m = rem / dh, d = rem - m * dh;
  wp[i] = round_tf32_bits(W[(long)k * N + m * D + h * dh + d]);
}
}  // namespace

int permute_round_wqkv(float* wp, const float* W, int Din, int nh, int dh, cudaStream_t st) {
  const long n = (long)Din * 3 * nh * dh;
  permute_round_wqkv_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(W, wp, Din, nh, dh);
  EBK_LAUNCH_CHECK();
  return EBK_OK;
}

bool qkv_attn_fused_supported(int L, int dh, int Din, const float* xd, const float* wp, const float* y) {
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  return L >= 1 && L <= 32 && (dh == 16 || dh == 20 || dh == 24 || dh == 32) && Din % 4 == 0 && al(xd) && al(wp) && al(y) &&
         encode_fn() != nullptr;
}
size_t qkv_tiles_floats(int n_seq, int nh, int dh) {
  const int st = (dh == 20) ? 20 : dh + 4;
  return (size_t)n_seq * nh * 3 * 32 * st;
}

// xd [n_seq * L, Din] (tf32 values), wp [Din, nh * 3 * dh] head-major (permute_round_wqkv) ->
//   y [n_seq * L, nh * dh] = tf32(dropout(attention output)), qkv_t (nullable) = the Q | K | V tiles for the backward
int qkv_attn_fused(const float* xd, int Din, const float* wp, int n_seq, int L, int nh, int dh, float* qkv_t, float* y,
                   Dropout drop, cudaStream_t st) {
  if (n_seq <= 0) return EBK_OK;
  EBK_CHECK_ARG(qkv_attn_fused_supported(L, dh, Din, xd, wp, y), "qkv_attn_fused: unsupported shape L=%d dh=%d Din=%d", L, dh, Din);
  if (g_sms == 0) {
    int dev = 0;
    EBK_CUDA(cudaGetDevice(&dev));
    EBK_CUDA(cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const int hpt = dh <= 20 ? 4 : 2, BN = hpt * 3 * dh, Ntot = nh * 3 * dh;
  const int spt = (128 / L) < 4 ? (128 / L) : 4;          // whole sequences per 128-row tile, one per epilogue warp
  const long M = (long)n_seq * L;
  const int n_row_tiles = ceil_div(n_seq, spt);
  const int env_pair = getenv("EBK_FUSED_PAIR") ? atoi(getenv("EBK_FUSED_PAIR")) : -1;   // experiments / tests
  const bool pair = env_pair >= 0 ? env_pair != 0 : (n_row_tiles >= 2 * (g_sms / 2));
  const int cm = pair ? 2 : 1;
  TParams p;
  memset(&p, 0, sizeof(p));
  p.C = nullptr; p.ldc = 0; p.M = (int)M; p.N = Ntot; p.K = Din; p.alpha = 1.0f;
  p.epi = GemmEpilogue{nullptr, nullptr, 0, 1, Dropout{0, 0, 1.0f, nullptr}, 0, false};
  p.BN = BN;
  p.tiles_n = ceil_div(nh, hpt);
  const int bn_cta = pair ? BN / 2 : BN;
  p.b_rows = (bn_cta + 31) & ~31;                             // MN-major operand: padded groups of 32 columns
  p.tiles_m = ceil_div(n_row_tiles, cm);
  p.tile_rows = spt * L;
  p.ksteps_total = ceil_div(Din, BK);
  p.ksteps_per_split = p.ksteps_total;
  p.splitk = 1;
  p.out_mode = 0;
  p.rv_smem = 0;
  p.stg_bufs = 2;
  p.c_tma = 0;
  p.att_L = L; p.att_spt = spt; p.att_nh = nh; p.att_nseq = n_seq;
  p.att_qkv_t = qkv_t; p.att_y = y; p.att_drop = drop;
  const size_t stage_bytes = (size_t)BM * BK * 4 + (size_t)p.b_rows * BK * 4;
  const size_t att_smem = dh == 16 ? AttGeo<16>::SMEM : dh == 20 ? AttGeo<20>::SMEM : dh == 24 ? AttGeo<24>::SMEM : AttGeo<32>::SMEM;
  const size_t budget = 226 * 1024 - 1024 - att_smem;
  int stages = (int)(budget / stage_bytes);
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  EBK_CHECK_ARG(stages >= 2, "qkv_attn_fused: shared memory budget (stage %zu B, attention %zu B)", stage_bytes, att_smem);
  p.stages = stages;
  CUtensorMap tmA, tmB;
  EBK_TRY(get_map(xd, Din, (int)M, Din, BM, false, &tmA));          // storage [M, K], box {32 k, 128 rows}
  EBK_TRY(get_map(wp, Ntot, Din, Ntot, 32, true, &tmB));            // storage [K, N], boxes {32 n, 32 k}
  const long n_items = (long)p.tiles_m * p.tiles_n;
  const int max_clusters = g_sms / cm;
  const int grid = (int)(n_items < max_clusters ? n_items : max_clusters) * cm;
  const size_t smem = (size_t)stages * stage_bytes + 1024 + att_smem;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(ATT_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cm;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = cm > 1 ? 1 : 0;
#define FUSED_LAUNCH(PAIR_, DH_)                                                                                       \
  {                                                                                                                   \
    EBK_CUDA(cudaFuncSetAttribute(gemm_tma_kernel<false, true, 1, PAIR_, DH_>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                  (int)smem));                                                                        \
    EBK_CUDA(cudaLaunchKernelEx(&cfg, gemm_tma_kernel<false, true, 1, PAIR_, DH_>, tmA, tmB, tmA, p));                 \
  }
#define FUSED_DH(DH_) { if (pair) FUSED_LAUNCH(true, DH_) else FUSED_LAUNCH(false, DH_) }
  if (dh == 16) FUSED_DH(16) else if (dh == 20) FUSED_DH(20) else if (dh == 24) FUSED_DH(24) else FUSED_DH(32)
#undef FUSED_DH
#undef FUSED_LAUNCH
  EBK_LAUNCH_CHECK();
  return EBK_OK;
}

}  // namespace ebk

// C-ABI entry points: error plumbing, the sequence encoder (news / user encoder)
// forward + backward composed from the kernels of this directory, and test exports.
#include <stdarg.h>
#include <stdlib.h>

#include <atomic>
#include <vector>

#include "ebk_common.cuh"

namespace ebk {

static thread_local char g_err[512] = "";
// Deferred weight gradient (ebk_seqenc_opts.defer_wgrad): the QKV wgrad GEMM (tensor-pipe bound) is independent of
// everything the optimizer's table pass (HBM bound) needs, so on request it runs on a library-owned side stream,
// forked after the dgrad GEMM, and is joined by ebk_join_deferred() before its output is consumed.  The only state
// kept between calls is "a fork is outstanding on this thread's side stream".
static thread_local bool g_side_pending = false;
static thread_local cudaStream_t g_side = nullptr;
static thread_local cudaEvent_t g_ev_fork = nullptr, g_ev_join = nullptr;
constexpr int GATHER_CHUNKS = 4;
static thread_local cudaEvent_t g_ev_chunk[GATHER_CHUNKS] = {nullptr, nullptr, nullptr, nullptr};
static int side_stream_init() {
  if (g_side == nullptr) {
    EBK_CUDA(cudaStreamCreateWithFlags(&g_side, cudaStreamNonBlocking));
    EBK_CUDA(cudaEventCreateWithFlags(&g_ev_fork, cudaEventDisableTiming));
    EBK_CUDA(cudaEventCreateWithFlags(&g_ev_join, cudaEventDisableTiming));
    for (int c = 0; c < GATHER_CHUNKS; ++c) EBK_CUDA(cudaEventCreateWithFlags(&g_ev_chunk[c], cudaEventDisableTiming));
  }
  return EBK_OK;
}
// per-call options -> kernel-side peer-table descriptor (data parallel, rank-sharded table)
static int peers_from_opts(const ebk_seqenc_opts* o, PeerTables* out) {
  PeerTables pt = {0, 0, {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr}};
  if (o != nullptr && o->peer_tables != nullptr && o->peer_world > 1) {
    EBK_CHECK_ARG(o->peer_world <= 8, "seqenc: at most 8 peer tables (one NVSwitch box), got %d", o->peer_world);
    EBK_CHECK_ARG(o->peer_shard_floats > 0 && o->peer_shard_floats % 4 == 0, "seqenc: peer_shard_floats must be a positive multiple of 4");
    pt.world = o->peer_world;
    pt.shard_floats = o->peer_shard_floats;
    for (int r = 0; r < o->peer_world; ++r) {
      EBK_CHECK_ARG(o->peer_tables[r] != nullptr, "seqenc: peer table %d is NULL", r);
      pt.p[r] = reinterpret_cast<const float*>(o->peer_tables[r]);
    }
  }
  *out = pt;
  return EBK_OK;
}

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// ---- launch counter + event profiler -----------------------------------------------------------
static std::atomic<long long> g_launches{0};
void note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

namespace {
struct ProfRec { int tag; int user; cudaEvent_t a, b; };
bool g_prof = false;
int g_prof_user = 0;  // 0 = news encoder call, 1 = user encoder call (set by the seqenc entry points)
std::vector<ProfRec> g_recs;
std::vector<cudaEvent_t> g_pool;
cudaEvent_t prof_event() {
  if (!g_pool.empty()) { cudaEvent_t e = g_pool.back(); g_pool.pop_back(); return e; }
  cudaEvent_t e; cudaEventCreate(&e); return e;
}
const char* kTagNames[T_NUM_TAGS] = {"qkv_gemm_fwd", "attn_core_fwd", "att_gemm_fwd", "attpool_fwd", "attpool_bwd",
                                     "colsum", "att_wgrad_gemm", "att_dgrad_gemm", "attn_core_bwd", "qkv_wgrad_gemm",
                                     "qkv_dgrad_gemm", "embed_scatter", "score_ce", "adam", "embed_pad",
                                     "conv_gemm_fwd", "conv_dz", "conv_wgrad_gemm", "conv_dgrad_gemm", "catview",
                                     "embed_gather"};
}  // namespace
bool prof_on() { return g_prof; }
void prof_set_group(int g) { g_prof_user = g; }
void prof_begin(int tag, cudaStream_t st) {
  ProfRec r{tag, g_prof_user, prof_event(), prof_event()};
  cudaEventRecord(r.a, st);
  g_recs.push_back(r);
}
void prof_end(int tag, cudaStream_t st) {
  (void)tag;
  cudaEventRecord(g_recs.back().b, st);
}

int gemm_dispatch(int math, const GemmOperandA& A, const float* B, int ldb, bool transB, float* C, int ldc,
                  int M, int N, int K, float beta, cudaStream_t st, int b_mode, const float* B_lo) {
  if (math == EBK_MATH_TF32) return gemm_tf32(A, B, ldb, transB, C, ldc, M, N, K, beta, st, b_mode);
  if (math == EBK_MATH_TF32X3) {
    // only a B that comes with its low parts can skip the in-flight split
    const int mode = (b_mode != GEMM_B_RAW && B_lo == nullptr) ? GEMM_B_RAW : b_mode;
    return gemm_tf32(A, B, ldb, transB, C, ldc, M, N, K, beta, st, mode, true, B_lo);
  }
  if (math == EBK_MATH_FP32) return gemm_f32(A, B, ldb, transB, C, ldc, M, N, K, beta, st);
  set_error("unknown math mode %d", math);
  return EBK_ERR_INVALID;
}

namespace {

// Workspace of one sequence-encoder call.  Saved activations first, backward scratch after.
struct SeqWs {
  float *qkv, *y0, *hbuf, *w;          // saved by forward
  float *dy, *dpre, *da, *dqkv, *dx;   // backward scratch
  float* colsum;                       // column-sum partials
  // weights packed (tf32-rounded, UMMA tile layout) as B operands of the tcgen05 GEMMs:
  float *wqkv_f, *attw_f;              //   forward  (B = W,   MN-major)
  float *wqkv_d, *attw_d;              //   dgrad    (B = W^T, K-major)
  float *wqkv_f_lo, *attw_f_lo;        //   3xTF32 low parts of the forward packs (EBK_MATH_TF32X3)
  float* dqkv_pk;                      // dQKV again, in the packed B layout of the weight-gradient GEMM
  // TMA path (EBK_MATH_TF32): dense tf32-rounded operands
  float *xd;                           //   dropout(gather(table, tok)) or the dense input, [R, Din]
  float *wqkv_r, *attw_r;              //   rounded copies of the weights
  float *wqkv_p;                       //   head-major permuted + rounded Wqkv (fused projection + attention forward)
  float *colpart, *colsum2;            //   per-sequence column-sum partials [n_seq, 2 att] and their reduction scratch
  size_t bytes;
};

SeqWs seq_layout(const ebk_seqenc_desc& d0, void* base) {
  ebk_seqenc_desc d = d0;
  if (d.att <= 0) d.att = 4;   // attention-only mode: the AttLayer2 buffers are unused, keep their sizing well-defined
  const size_t R = (size_t)d.n_seq * d.L, D = (size_t)d.nh * d.dh;
  size_t off = 0;
  auto take = [&](size_t nfloat) {
    float* p = base ? reinterpret_cast<float*>(reinterpret_cast<char*>(base) + off) : nullptr;
    off += align_up(nfloat * sizeof(float), 256);
    return p;
  };
  SeqWs w;
  // Q|K|V of the forward: [R, 3D] rows, or the zero-padded per-(sequence, head) tiles of the fused forward
  size_t qkv_floats = R * 3 * D;
  if (d.dh == 16 || d.dh == 20 || d.dh == 24 || d.dh == 32) {
    const size_t tiled = qkv_tiles_floats(d.n_seq, d.nh, d.dh);
    if (tiled > qkv_floats) qkv_floats = tiled;
  }
  w.qkv = take(qkv_floats);
  w.y0 = take(R * D);
  w.hbuf = take(R * d.att);
  w.w = take(R);
  w.dy = take(R * D);
  w.dpre = take(R * d.att);
  w.da = take(R);
  w.dqkv = take(R * 3 * D);
  w.dx = take(R * (size_t)d.Din);
  w.colsum = take(colsum_partial_floats((int)R, d.att));
  w.wqkv_f = take(gemm_tf32_packed_floats(3 * (int)D, d.Din, false));
  w.attw_f = take(gemm_tf32_packed_floats(d.att, (int)D, false));
  w.wqkv_d = take(gemm_tf32_packed_floats(d.Din, 3 * (int)D, true));
  w.attw_d = take(gemm_tf32_packed_floats((int)D, d.att, true));
  w.wqkv_f_lo = take(gemm_tf32_packed_floats(3 * (int)D, d.Din, false));
  w.attw_f_lo = take(gemm_tf32_packed_floats(d.att, (int)D, false));
  w.dqkv_pk = take(gemm_tf32_packed_floats(3 * (int)D, (int)R, false));
  w.xd = take(R * (size_t)d.Din + 64);  // + 64: TMA boxes of the last row may touch (zero-filled) columns past it
  w.wqkv_r = take((size_t)d.Din * 3 * D + 64);
  w.attw_r = take(D * (size_t)d.att + 64);
  w.wqkv_p = take((size_t)d.Din * 3 * D + 64);
  w.colpart = take((size_t)d.n_seq * 2 * d.att);
  w.colsum2 = take(colsum_partial_floats(d.n_seq, 2 * d.att));
  w.bytes = off;
  return w;
}

// Fused SelfAttention forward (QKV projection with the attention in its epilogue): the forward and the backward of a
// call decide it from the same descriptor + workspace, because it fixes the layout of the saved Q|K|V.
bool fused_attn(const ebk_seqenc_desc& d, const SeqWs& ws) {
  // (read per call: tests and experiments flip these between calls)
  const char* e1 = getenv("EBK_FUSED_ATTN");
  const char* e2 = getenv("EBK_DP_CHUNKED_GATHER");
  const bool on = !(e1 && atoi(e1) == 0) && !(e2 && atoi(e2) != 0);
  return on && qkv_attn_fused_supported(d.L, d.dh, d.Din, ws.xd, ws.wqkv_p, ws.y0);
}

// The all-TMA GEMM path: tf32 tensor-core math, every row stride a multiple of 16 bytes.
bool tma_path(const ebk_seqenc_desc& d, const SeqWs& ws) {
  const int D = d.nh * d.dh;
  return d.math == EBK_MATH_TF32 && d.att % 4 == 0 && gemm_tma_eligible(ws.xd, d.Din, ws.wqkv_r, 3 * D, 1, 1, 1) &&
         (d.att == 0 || gemm_tma_eligible(ws.y0, D, ws.attw_r, d.att, 1, 1, 1));
}

int check_desc(const ebk_seqenc_desc* d) {
  EBK_CHECK_ARG(d != nullptr, "seqenc: null descriptor");
  EBK_CHECK_ARG(d->n_seq >= 0 && d->L >= 1 && d->L <= 64, "seqenc: need n_seq>=0, 1<=L<=64 (n_seq=%d L=%d)", d->n_seq, d->L);
  EBK_CHECK_ARG(d->Din >= 4 && d->Din % 4 == 0, "seqenc: Din=%d must be a positive multiple of 4", d->Din);
  EBK_CHECK_ARG(d->nh >= 1 && d->dh >= 1 && d->dh <= 32, "seqenc: need nh>=1, 1<=dh<=32 (nh=%d dh=%d)", d->nh, d->dh);
  EBK_CHECK_ARG((d->nh * d->dh) % 4 == 0, "seqenc: D=nh*dh=%d must be a multiple of 4", d->nh * d->dh);
  EBK_CHECK_ARG(d->att >= 0, "seqenc: att=%d", d->att);
  EBK_CHECK_ARG(d->dropout >= 0.0f && d->dropout < 1.0f, "seqenc: dropout=%f outside [0,1)", d->dropout);
  EBK_CHECK_ARG((long)d->n_seq * d->L < (1L << 31), "seqenc: n_seq*L overflows int32");
  return EBK_OK;
}

}  // namespace
}  // namespace ebk

using namespace ebk;

extern "C" const char* ebk_last_error(void) { return g_err; }
namespace ebk { void gemm_tf32_set_debug(long long* buf, int target); }
// debugging aid: device buffer of 3*96*4 int64 receiving clock64() stamps of CTA 0 of the target-th
// tcgen05 GEMM launched after this call (NULL disarms)
extern "C" int ebk_debug_gemm_timeline(long long* device_buf, int target) {
  ebk::gemm_tf32_set_debug(device_buf, target);
  return 0;
}
extern "C" int ebk_join_deferred(void* stream) {
  if (g_side_pending) {
    EBK_CUDA(cudaEventRecord(g_ev_join, g_side));
    EBK_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, g_ev_join, 0));
    g_side_pending = false;
  }
  return EBK_OK;
}
// ---- CUDA IPC plumbing for the rank-sharded table (one process per GPU, all on one NVSwitch box) ----------
extern "C" int ebk_ipc_export(const void* ptr, void* handle64, size_t* offset) {
  EBK_CHECK_ARG(ptr && handle64 && offset, "ipc_export: null pointer");
  typedef int (*RangeFn)(unsigned long long*, size_t*, unsigned long long);
  static RangeFn range = nullptr;
  if (range == nullptr) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    EBK_CUDA(cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &q));
    EBK_CHECK_ARG(fn != nullptr && q == cudaDriverEntryPointSuccess, "ipc_export: cuMemGetAddressRange unavailable");
    range = reinterpret_cast<RangeFn>(fn);
  }
  unsigned long long base = 0;
  size_t size = 0;
  const int rc = range(&base, &size, (unsigned long long)(uintptr_t)ptr);
  EBK_CHECK_ARG(rc == 0, "ipc_export: cuMemGetAddressRange failed (%d)", rc);
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
  EBK_CUDA(cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(handle64), reinterpret_cast<void*>((uintptr_t)base)));
  *offset = (size_t)((unsigned long long)(uintptr_t)ptr - base);
  return EBK_OK;
}
extern "C" int ebk_ipc_open(const void* handle64, size_t offset, void** out) {
  EBK_CHECK_ARG(handle64 && out, "ipc_open: null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  void* base = nullptr;
  EBK_CUDA(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
  *out = reinterpret_cast<char*>(base) + offset;
  return EBK_OK;
}
// plain asynchronous device-to-device copy (peer mappings included): lets a rank refresh its replica of the
// rank-sharded table from its peers' memory WITHOUT a collective (inference on one rank after sharded training)
extern "C" int ebk_memcpy_async(void* dst, const void* src, size_t bytes, void* stream) {
  EBK_CHECK_ARG(bytes == 0 || (dst && src), "memcpy_async: null pointer");
  if (bytes) EBK_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return EBK_OK;
}
extern "C" long long ebk_launch_count(void) { return g_launches.load(); }
extern "C" int ebk_prof_enable(int on) {
  g_prof = on != 0;
  if (g_prof) {
    for (auto& r : g_recs) { g_pool.push_back(r.a); g_pool.push_back(r.b); }
    g_recs.clear();
  }
  return EBK_OK;
}
extern "C" int ebk_prof_is_enabled(void) { return g_prof ? 1 : 0; }
extern "C" int ebk_prof_num_tags(void) { return 2 * T_NUM_TAGS; }
extern "C" const char* ebk_prof_tag_name(int slot) {
  static thread_local char buf[64];
  if (slot < 0 || slot >= 2 * T_NUM_TAGS) return "";
  snprintf(buf, sizeof(buf), "%s%s", slot >= T_NUM_TAGS ? "user." : "news.", kTagNames[slot % T_NUM_TAGS]);
  return buf;
}
extern "C" int ebk_prof_collect(double* ms_out, long long* count_out) {
  for (int i = 0; i < 2 * T_NUM_TAGS; ++i) { ms_out[i] = 0.0; count_out[i] = 0; }
  for (auto& r : g_recs) {
    EBK_CUDA(cudaEventSynchronize(r.b));
    float ms = 0.0f;
    EBK_CUDA(cudaEventElapsedTime(&ms, r.a, r.b));
    int slot = r.tag + (r.user ? T_NUM_TAGS : 0);
    ms_out[slot] += ms;
    count_out[slot] += 1;
  }
  return EBK_OK;
}
extern "C" int ebk_version(void) { return 100; }

extern "C" int ebk_device_ok(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  cudaDeviceProp p;
  if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) return 0;
  return p.major == 10 ? 1 : 0;
}

extern "C" size_t ebk_seqenc_workspace_bytes(const ebk_seqenc_desc* d) {
  if (check_desc(d) != EBK_OK) return 0;
  return seq_layout(*d, nullptr).bytes;
}

extern "C" int ebk_seqenc_uses_tma(const ebk_seqenc_desc* d) {
  if (check_desc(d) != EBK_OK) return 0;
  return tma_path(*d, seq_layout(*d, nullptr)) ? 1 : 0;
}

extern "C" int ebk_seqenc_fwd(const ebk_seqenc_desc* d, const int32_t* tok, const float* table_or_x,
                              const float* Wqkv, const float* attW, const float* attb, const float* attq,
                              int training, uint64_t seed1, uint64_t seed2, void* workspace,
                              size_t workspace_bytes, float* out, void* stream) {
  return ebk_seqenc_fwd_opts(d, nullptr, tok, table_or_x, Wqkv, attW, attb, attq, training, seed1, seed2, workspace,
                             workspace_bytes, out, stream);
}

extern "C" int ebk_seqenc_fwd_opts(const ebk_seqenc_desc* d, const ebk_seqenc_opts* opts, const int32_t* tok,
                                   const float* table_or_x, const float* Wqkv, const float* attW, const float* attb,
                                   const float* attq, int training, uint64_t seed1, uint64_t seed2, void* workspace,
                                   size_t workspace_bytes, float* out, void* stream) {
  EBK_TRY(check_desc(d));
  PeerTables peers;
  EBK_TRY(peers_from_opts(opts, &peers));
  if (d->n_seq == 0) return EBK_OK;
  const bool pool = d->att > 0;   // att == 0: stop after the SelfAttention, out = [n_seq*L, D] (no dropout on it)
  EBK_CHECK_ARG(table_or_x && Wqkv && out && workspace && (!pool || (attW && attb && attq)), "seqenc_fwd: null pointer");
  EBK_CHECK_ARG(tok == nullptr || d->V >= 1, "seqenc_fwd: V=%d with a token gather", d->V);
  EBK_CHECK_ARG(tok != nullptr || !(training && d->dropout > 0.0f), "seqenc_fwd: dropout on a dense input is not supported");
  SeqWs ws = seq_layout(*d, workspace);
  if (workspace_bytes < ws.bytes) {
    set_error("seqenc_fwd: workspace %zu < %zu bytes", workspace_bytes, ws.bytes);
    return EBK_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  EBK_TRY(ebk_join_deferred(stream));   // a forgotten deferred weight gradient may still read this workspace
  g_prof_user = tok == nullptr;
  const int R = d->n_seq * d->L, D = d->nh * d->dh;
  const ebk_step_params* sp = opts != nullptr ? opts->step_dev : nullptr;   // device-resident seeds (CUDA graphs)
  Dropout drop1 = make_dropout(training != 0, d->dropout, seed1, sp ? &sp->seed1 : nullptr);
  Dropout drop2 = make_dropout(training != 0, d->dropout, seed2, sp ? &sp->seed2 : nullptr);
  const Dropout none = make_dropout(false, 0.0f, 0);
  if (tma_path(*d, ws)) {
    // ---- all-TMA path: every GEMM operand is materialised dense, masked and tf32-rounded by the layer
    // before it, so the tensor-core kernels spend no issue slots on operand preparation ----
    const bool remote = tok && peers.world > 1;   // rank-sharded table: rows come over NVLink
    EBK_TRY(round_tf32_copy(ws.wqkv_r, Wqkv, (size_t)d->Din * 3 * D, st));
    if (pool) EBK_TRY(round_tf32_copy(ws.attw_r, attW, (size_t)D * d->att, st));
    // (stored rounded to tf32: the attention kernels feed it to mma.sync without touching it again)
    const GemmEpilogue round_epi{nullptr, nullptr, 0, 1, none, 0, true};
    // opt-in (EBK_DP_CHUNKED_GATHER=1): measured on 2 GPUs it does not pay (5.81 vs 5.73 ms/step: the remote share
    // of the gather is small there and four partial GEMM waves cost more); not yet measured on 8
    // the whole gather: plain (one read per position), or through a token CSR built in ebk_seqenc_opts.token_csr_ws (each
    // distinct token's row read once: what pays when the rows are remote)
    auto gather_rows = [&]() -> int {
      if (tok != nullptr && opts != nullptr && opts->token_csr_ws != nullptr) {
        EBK_CHECK_ARG(opts->token_csr_ws_bytes >= token_csr_bytes(R, d->V), "seqenc_fwd: token_csr_ws %zu < %zu bytes",
                      opts->token_csr_ws_bytes, token_csr_bytes(R, d->V));
        TokenCsr csr;
        if (prof_on()) prof_begin(T_EMBED_GATHER, st);
        EBK_TRY(token_csr_build(R, d->V, tok, opts->token_csr_ws, &csr, st));
        EBK_TRY(embed_rows_csr(R, d->Din, d->V, tok, table_or_x, drop1, ws.xd, st, remote ? &peers : nullptr, csr));
        if (prof_on()) prof_end(T_EMBED_GATHER, st);
        return EBK_OK;
      }
      EBK_PROF(T_EMBED_GATHER, embed_rows(R, d->Din, d->V, tok, table_or_x, tok ? drop1 : none, ws.xd, st,
                                          remote ? &peers : nullptr));
      return EBK_OK;
    };
    const char* env_ch = getenv("EBK_DP_CHUNKED_GATHER");
    const bool env_chunked = env_ch != nullptr && atoi(env_ch) != 0;
    if (remote && env_chunked && R >= GATHER_CHUNKS * 4096) {
      // The peer-memory gather is NVLink-bound (630 GB/s, ~0.8 ms on 8 GPUs) and the QKV projection tensor-bound:
      // software-pipeline them over row chunks -- the gather of chunk c+1 runs on the side stream while the GEMM
      // of chunk c runs here.
      EBK_TRY(side_stream_init());
      EBK_CUDA(cudaEventRecord(g_ev_fork, st));
      EBK_CUDA(cudaStreamWaitEvent(g_side, g_ev_fork, 0));
      const int chunk = ceil_div(ceil_div(R, GATHER_CHUNKS), 256) * 256;
      for (int c = 0; c < GATHER_CHUNKS; ++c) {
        const int r0 = c * chunk, rows = (r0 + chunk <= R ? chunk : R - r0);
        if (rows <= 0) break;
        {
          cudaStream_t st = g_side;   // EBK_PROF records its events on `st`
          EBK_PROF(T_EMBED_GATHER, embed_rows(rows, d->Din, d->V, tok + r0, table_or_x, drop1, ws.xd + (size_t)r0 * d->Din, st,
                                              &peers, r0));
        }
        EBK_CUDA(cudaEventRecord(g_ev_chunk[c], g_side));
      }
      for (int c = 0; c < GATHER_CHUNKS; ++c) {
        const int r0 = c * chunk, rows = (r0 + chunk <= R ? chunk : R - r0);
        if (rows <= 0) break;
        EBK_CUDA(cudaStreamWaitEvent(st, g_ev_chunk[c], 0));
        EBK_PROF(T_QKV_FWD, gemm_tma(ws.xd + (size_t)r0 * d->Din, d->Din, false, ws.wqkv_r, 3 * D, false,
                                     ws.qkv + (size_t)r0 * 3 * D, 3 * D, rows, 3 * D, d->Din, 0.0f, 1.0f, st, -1, &round_epi));
      }
    } else if (fused_attn(*d, ws)) {
      // ---- north-star kernel: gather -> [QKV projection + per-head softmax(QK^T/sqrt(dh))^T V in ONE tcgen05 kernel]:
      // the accumulator tile is whole sequences x whole heads (weights permuted head-major), the attention runs in the
      // GEMM epilogue from TMEM through shared memory, Q|K|V reach HBM only as the tiles the backward needs
      EBK_TRY(gather_rows());
      EBK_TRY(permute_round_wqkv(ws.wqkv_p, Wqkv, d->Din, d->nh, d->dh, st));
      float* y0f = pool ? ws.y0 : out;
      EBK_PROF(T_QKV_FWD, qkv_attn_fused(ws.xd, d->Din, ws.wqkv_p, d->n_seq, d->L, d->nh, d->dh, ws.qkv, y0f,
                                         pool ? drop2 : none, st));
      if (!pool) return EBK_OK;
      goto attlayer;
    } else {
      EBK_TRY(gather_rows());
      // (1) Q|K|V = dropout1(gather(table, tok)) . Wqkv        nrms.py:134-139, layers.py:214-230
      EBK_PROF(T_QKV_FWD, gemm_tma(ws.xd, d->Din, false, ws.wqkv_r, 3 * D, false, ws.qkv, 3 * D, R, 3 * D, d->Din, 0.0f,
                                   1.0f, st, -1, &round_epi));
    }
    {
    // (2) attention core; its output is stored as tf32(dropout2(Y0)) -- the only form AttLayer2 reads
    float* y0 = pool ? ws.y0 : out;
    const Dropout dropy = pool ? drop2 : none;
    if (attention_pre_supported(d->L, d->dh, ws.qkv, y0, ws.qkv)) {
      EBK_PROF(T_ATTN_FWD, attention_core_fwd_pre(d->n_seq, d->L, d->nh, d->dh, ws.qkv, y0, dropy, st));
    } else if (attention_mma_supported(d->L, d->dh, ws.qkv, y0, ws.qkv)) {
      EBK_PROF(T_ATTN_FWD, attention_core_fwd_mma(d->n_seq, d->L, d->nh, d->dh, ws.qkv, y0, st, dropy, true));
    } else {
      EBK_CHECK_ARG(!dropy.on(), "seqenc_fwd: dropout needs L <= 32 and dh <= 32 on the tensor-core path");
      EBK_PROF(T_ATTN_FWD, attention_core_fwd(d->n_seq, d->L, d->nh, d->dh, ws.qkv, y0, st));
      if (pool) EBK_TRY(round_tf32_copy(y0, y0, (size_t)R * D, st));
    }
    }
    if (!pool) return EBK_OK;
  attlayer:
    // (3) pre-activation of AttLayer2                                nrms.py:153-156, layers.py:65
    static const int att_fwd_tall = getenv("EBK_ATT_FWD_TALL") ? atoi(getenv("EBK_ATT_FWD_TALL")) : -1;   // experiments
    EBK_PROF(T_ATT_GEMM_FWD, gemm_tma(ws.y0, D, false, ws.attw_r, d->att, false, ws.hbuf, d->att, R, d->att, D, 0.0f,
                                      1.0f, st, att_fwd_tall));
    // (4) tanh, .q, exp, normalise (+1e-7), pool                     layers.py:65-81
    EBK_PROF(T_POOL_FWD, attpool_fwd(d->n_seq, d->L, D, d->att, ws.y0, none, ws.hbuf, attb, attq, ws.w, out, st));
    return EBK_OK;
  }
  // only the all-TMA path gathers through peer mappings: anything else would silently read a stale local shard
  EBK_CHECK_ARG(peers.world <= 1, "seqenc_fwd: peer tables need the all-TMA path (EBK_MATH_TF32, att %% 4 == 0, "
                "16-byte aligned row strides; query ebk_seqenc_uses_tma)");
  // Tensor-core modes: the weights are packed once per call (rounded to tf32, arranged in the GEMM's
  // shared-memory tile layout) and kept in the workspace for the backward pass.
  const bool tc = d->math != EBK_MATH_FP32;
  const bool x3 = d->math == EBK_MATH_TF32X3;
  GemmOperandA ax{table_or_x, d->Din, false, tok, d->V, tok ? drop1 : none, d->Din};
  GemmOperandA ay{ws.y0, D, false, nullptr, 0, drop2, D};
  const bool pk_qkv = tc && gemm_tf32_eligible(ax, Wqkv, 3 * D, R, 3 * D, d->Din);
  const bool pk_att = pool && tc && gemm_tf32_eligible(ay, attW, d->att, R, d->att, D);
  if (pk_qkv) {
    EBK_TRY(gemm_tf32_pack_b(ws.wqkv_f, x3 ? ws.wqkv_f_lo : nullptr, Wqkv, 3 * D, false, 3 * D, d->Din, st));
    if (!x3) EBK_TRY(gemm_tf32_pack_b(ws.wqkv_d, nullptr, Wqkv, 3 * D, true, d->Din, 3 * D, st));
  }
  if (pk_att) {
    EBK_TRY(gemm_tf32_pack_b(ws.attw_f, x3 ? ws.attw_f_lo : nullptr, attW, d->att, false, d->att, D, st));
    if (!x3) EBK_TRY(gemm_tf32_pack_b(ws.attw_d, nullptr, attW, d->att, true, D, d->att, st));
  }
  // (1) Q|K|V = dropout1(gather(table, tok)) . Wqkv        nrms.py:134-139, layers.py:214-230
  EBK_PROF(T_QKV_FWD, gemm_dispatch(d->math, ax, pk_qkv ? ws.wqkv_f : Wqkv, 3 * D, false, ws.qkv, 3 * D, R, 3 * D,
                                    d->Din, 0.0f, st, pk_qkv ? GEMM_B_PACKED : GEMM_B_RAW,
                                    (pk_qkv && x3) ? ws.wqkv_f_lo : nullptr));
  // (2) per-head softmax(QK^T/sqrt(dh)) and the adjoint product    layers.py:231-252
  // training with tf32 contractions: warp-level tensor-core attention; otherwise the exact fp32 kernel
  float* y0o = pool ? ws.y0 : out;
  if (d->math == EBK_MATH_TF32 && training && attention_mma_supported(d->L, d->dh, ws.qkv, y0o, ws.qkv)) {
    EBK_PROF(T_ATTN_FWD, attention_core_fwd_mma(d->n_seq, d->L, d->nh, d->dh, ws.qkv, y0o, st));
  } else {
    EBK_PROF(T_ATTN_FWD, attention_core_fwd(d->n_seq, d->L, d->nh, d->dh, ws.qkv, y0o, st));
  }
  if (!pool) return EBK_OK;
  // (3) pre-activation of AttLayer2: dropout2(Y0) . W              nrms.py:153-156, layers.py:65
  EBK_PROF(T_ATT_GEMM_FWD, gemm_dispatch(d->math, ay, pk_att ? ws.attw_f : attW, d->att, false, ws.hbuf, d->att, R,
                                         d->att, D, 0.0f, st, pk_att ? GEMM_B_PACKED : GEMM_B_RAW,
                                         (pk_att && x3) ? ws.attw_f_lo : nullptr));
  // (4) tanh, .q, exp, normalise (+1e-7), pool                     layers.py:65-81
  EBK_PROF(T_POOL_FWD, attpool_fwd(d->n_seq, d->L, D, d->att, ws.y0, drop2, ws.hbuf, attb, attq, ws.w, out, st));
  return EBK_OK;
}

extern "C" int ebk_seqenc_bwd(const ebk_seqenc_desc* d, const int32_t* tok, const float* table_or_x,
                              const float* Wqkv, const float* attW, const float* attb, const float* attq,
                              int training, uint64_t seed1, uint64_t seed2, void* workspace,
                              size_t workspace_bytes, const float* d_out, float* dWqkv, float* dattW,
                              float* dattb, float* dattq, float* d_table, float* d_x, void* stream) {
  return ebk_seqenc_bwd_opts(d, nullptr, tok, table_or_x, Wqkv, attW, attb, attq, training, seed1, seed2, workspace,
                             workspace_bytes, d_out, dWqkv, dattW, dattb, dattq, d_table, d_x, stream);
}

extern "C" int ebk_seqenc_bwd_opts(const ebk_seqenc_desc* d, const ebk_seqenc_opts* opts, const int32_t* tok,
                                   const float* table_or_x, const float* Wqkv, const float* attW, const float* attb,
                                   const float* attq, int training, uint64_t seed1, uint64_t seed2, void* workspace,
                                   size_t workspace_bytes, const float* d_out, float* dWqkv, float* dattW,
                                   float* dattb, float* dattq, float* d_table, float* d_x, void* stream) {
  EBK_TRY(check_desc(d));
  const bool defer_wgrad = opts != nullptr && opts->defer_wgrad != 0;
  cudaEvent_t table_grad_event = opts != nullptr ? (cudaEvent_t)opts->table_grad_event : nullptr;
  if (d->n_seq == 0) return EBK_OK;
  (void)attb;
  const bool pool = d->att > 0;   // att == 0: d_out is the gradient of the [n_seq*L, D] attention output
  EBK_CHECK_ARG(table_or_x && Wqkv && d_out && workspace && (!pool || (attW && attq)), "seqenc_bwd: null pointer");
  EBK_CHECK_ARG(dWqkv && (!pool || (dattW && dattb && dattq)), "seqenc_bwd: null parameter-gradient pointer");
  EBK_CHECK_ARG(tok == nullptr || d_x == nullptr || d_table == nullptr,
                "seqenc_bwd: with token ids give d_table (scatter here) OR d_x (rows for ebk_embed_adam_step), not both");
  EBK_CHECK_ARG(tok != nullptr || d_table == nullptr, "seqenc_bwd: d_table needs token ids");
  EBK_CHECK_ARG(tok != nullptr || !(training && d->dropout > 0.0f), "seqenc_bwd: dropout on a dense input is not supported");
  SeqWs ws = seq_layout(*d, workspace);
  if (workspace_bytes < ws.bytes) {
    set_error("seqenc_bwd: workspace %zu < %zu bytes", workspace_bytes, ws.bytes);
    return EBK_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  g_prof_user = tok == nullptr;
  const int R = d->n_seq * d->L, D = d->nh * d->dh;
  const ebk_step_params* sp = opts != nullptr ? opts->step_dev : nullptr;   // device-resident seeds (CUDA graphs)
  Dropout drop1 = make_dropout(training != 0, d->dropout, seed1, sp ? &sp->seed1 : nullptr);
  Dropout drop2 = make_dropout(training != 0, d->dropout, seed2, sp ? &sp->seed2 : nullptr);
  const Dropout none = make_dropout(false, 0.0f, 0);

  if (tma_path(*d, ws)) {
    // ---- all-TMA path (see ebk_seqenc_fwd): ws.xd, ws.y0 (= tf32(dropout2(Y0))), ws.wqkv_r, ws.attw_r are
    // the forward's; dpre and dQKV are rounded to tf32 by the kernels that produce them ----
    if (!pool) {
      EBK_TRY(round_tf32_copy(ws.dy, d_out, (size_t)R * D, st));   // the attention kernel multiplies it as tf32
    } else {
    EBK_PROF(T_POOL_BWD, attpool_bwd_fused(d->n_seq, d->L, D, d->att, ws.y0, ws.hbuf, attq, ws.w, d_out, ws.da, ws.dpre,
                                           ws.colpart, st));
    // (the AttLayer2 parameter gradients -- db, dq, dW -- only need y0 / dpre / colpart, which stay valid: they are
    // computed at the END of this call, behind the table-gradient scatter, so that under data parallel they overlap the
    // table gradient's reduce-scatter together with the QKV weight gradient)
    // dY0 = tf32(dropout2'(w_t d_out + dpre W^T)): pooling term, dropout backward and the rounding for the
    // attention kernel all happen in the GEMM epilogue
    const GemmEpilogue dy_epi{ws.w, d_out, D, d->L, drop2, D, true};
    // (128-row tiles: this GEMM is epilogue-bound, so the double-buffered accumulator matters more than B traffic)
    EBK_PROF(T_ATT_DGRAD, gemm_tma(ws.dpre, d->att, false, ws.attw_r, d->att, true, ws.dy, D, R, D, d->att, 0.0f, 1.0f, st,
                                   0, &dy_epi));
    }
    // SelfAttention core backward
    if (attention_pre_supported(d->L, d->dh, ws.qkv, ws.dy, ws.dqkv)) {
      const bool tiled = fused_attn(*d, ws);   // the fused forward saved Q|K|V as per-(sequence, head) tiles
      EBK_PROF(T_ATTN_BWD, attention_core_bwd_pre(d->n_seq, d->L, d->nh, d->dh, ws.qkv, ws.dy, ws.dqkv, st, tiled));
    } else if (attention_mma_supported(d->L, d->dh, ws.qkv, ws.dy, ws.dqkv)) {
      EBK_PROF(T_ATTN_BWD, attention_core_bwd_mma(d->n_seq, d->L, d->nh, d->dh, ws.qkv, ws.dy, none, ws.dqkv, true, st));
    } else {
      EBK_PROF(T_ATTN_BWD, attention_core_bwd(d->n_seq, d->L, d->nh, d->dh, ws.qkv, ws.dy, none, ws.dqkv, true, st));
    }
    // dX = dQKV Wqkv^T first: the table gradient is the large one, and in data parallel its collective can then
    // overlap the weight-gradient GEMM below
    if (d_table != nullptr || d_x != nullptr) {
      float* dx = d_x ? d_x : ws.dx;
      EBK_PROF(T_QKV_DGRAD, gemm_tma(ws.dqkv, 3 * D, false, ws.wqkv_r, 3 * D, true, dx, d->Din, R, d->Din, 3 * D, 0.0f, 1.0f,
                                     st, -1));
      if (tok && d_table) EBK_PROF(T_SCATTER, scatter_rows_add(R, d->Din, d->V, tok, dx, drop1, d_table, st));
    }
    if (tok && table_grad_event) EBK_CUDA(cudaEventRecord(table_grad_event, st));
    if (pool) {
      // db += sum_r dpre_r ; dq += sum_r h_r da_r   (second, deterministic stage over the per-sequence partials)
      EBK_PROF(T_COLSUM, colsum_accum2_ws(d->n_seq, 2 * d->att, d->att, ws.colpart, 2 * d->att, dattb, dattq, ws.colsum2, st));
      // dW += X^T dpre
      EBK_PROF(T_ATT_WGRAD, gemm_tma(ws.y0, D, true, ws.dpre, d->att, false, dattW, d->att, D, d->att, R, 1.0f, 1.0f, st, -1));
    }
    // dWqkv += X^T dQKV  (X = dropout1(gather)); deferred mode: on the side stream, behind the dgrad GEMM
    if (defer_wgrad && tok != nullptr) {
      EBK_TRY(side_stream_init());
      EBK_CUDA(cudaEventRecord(g_ev_fork, st));
      EBK_CUDA(cudaStreamWaitEvent(g_side, g_ev_fork, 0));
      cudaStream_t main_st = st;
      {
        cudaStream_t st = g_side;   // EBK_PROF records its events on `st`
        EBK_PROF(T_QKV_WGRAD, gemm_tma(ws.xd, d->Din, true, ws.dqkv, 3 * D, false, dWqkv, 3 * D, d->Din, 3 * D, R, 1.0f,
                                       1.0f, st, -1));
      }
      (void)main_st;
      g_side_pending = true;
      return EBK_OK;
    }
    EBK_PROF(T_QKV_WGRAD, gemm_tma(ws.xd, d->Din, true, ws.dqkv, 3 * D, false, dWqkv, 3 * D, d->Din, 3 * D, R, 1.0f, 1.0f,
                                   st, -1));
    return EBK_OK;
  }
  // tensor-core mode: the forward left the packed weights in the workspace, and the kernels that
  // produce dpre / dQKV round them to tf32 on store so they can be staged as B operands with cp.async.
  const bool tc = d->math != EBK_MATH_FP32;
  const bool x3 = d->math == EBK_MATH_TF32X3;
  const bool rnd = tc && !x3;
  GemmOperandA adp{ws.dpre, d->att, false, nullptr, 0, none, 0};
  GemmOperandA adq{ws.dqkv, 3 * D, false, nullptr, 0, none, 0};
  const bool pk_att = pool && rnd && gemm_tf32_eligible(adp, attW, d->att, R, D, d->att) &&
                      gemm_tf32_eligible(GemmOperandA{ws.y0, D, false, nullptr, 0, drop2, D}, attW, d->att, R, d->att, D);
  const bool pk_qkv = rnd && gemm_tf32_eligible(adq, Wqkv, 3 * D, R, d->Din, 3 * D) &&
                      gemm_tf32_eligible(GemmOperandA{table_or_x, d->Din, false, tok, d->V, tok ? drop1 : none, d->Din},
                                         Wqkv, 3 * D, R, 3 * D, d->Din);
  const float* dyp = pool ? ws.dy : d_out;
  const Dropout dropy = pool ? drop2 : none;
  if (pool) {
  // AttLayer2 backward (layers.py:55-81)
  EBK_PROF(T_POOL_BWD, attpool_bwd(d->n_seq, d->L, D, d->att, ws.y0, drop2, ws.hbuf, attq, ws.w, d_out, ws.da, ws.dpre,
                      ws.dy, rnd, st));
  EBK_PROF(T_COLSUM, colsum_accum_ws(R, d->att, ws.hbuf, d->att, ws.da, dattq, ws.colsum, st));   // dq = sum_r h_r da_r
  EBK_PROF(T_COLSUM, colsum_accum_ws(R, d->att, ws.dpre, d->att, nullptr, dattb, ws.colsum, st)); // db = sum_r dpre_r
  GemmOperandA ayT{ws.y0, D, true, nullptr, 0, drop2, D};                              // dW += X^T dpre
  EBK_PROF(T_ATT_WGRAD, gemm_dispatch(d->math, ayT, ws.dpre, d->att, false, dattW, d->att, D, d->att, R, 1.0f, st,
                                      rnd ? GEMM_B_ROUNDED : GEMM_B_RAW));
  // dX += dpre W^T
  EBK_PROF(T_ATT_DGRAD, gemm_dispatch(d->math, adp, pk_att ? ws.attw_d : attW, d->att, true, ws.dy, D, R, D, d->att,
                                      1.0f, st, pk_att ? GEMM_B_PACKED : GEMM_B_RAW));
  }
  // SelfAttention core backward (dropout2 mask applied while reading dy)
  GemmOperandA axT{table_or_x, d->Din, true, tok, d->V, tok ? drop1 : none, d->Din};
  const bool pk_dq = rnd && gemm_tf32_eligible(axT, ws.dqkv, 3 * D, d->Din, 3 * D, R) &&
                     (d->dh == 8 || d->dh == 16 || d->dh == 20 || d->dh == 32);
  if (d->math == EBK_MATH_TF32 && training && attention_mma_supported(d->L, d->dh, ws.qkv, dyp, ws.dqkv)) {
    EBK_PROF(T_ATTN_BWD, attention_core_bwd_mma(d->n_seq, d->L, d->nh, d->dh, ws.qkv, dyp, dropy, ws.dqkv, rnd, st,
                                                pk_dq ? ws.dqkv_pk : nullptr,
                                                pk_dq ? gemm_tf32_bn(3 * D, R, false) : 0));
  } else {
    EBK_PROF(T_ATTN_BWD, attention_core_bwd(d->n_seq, d->L, d->nh, d->dh, ws.qkv, dyp, dropy, ws.dqkv, rnd, st,
                                            pk_dq ? ws.dqkv_pk : nullptr,
                                            pk_dq ? gemm_tf32_bn(3 * D, R, false) : 0));
  }
  // dWqkv += X^T dQKV  (X = dropout1(gather))
  EBK_PROF(T_QKV_WGRAD, gemm_dispatch(d->math, axT, pk_dq ? ws.dqkv_pk : ws.dqkv, 3 * D, false, dWqkv, 3 * D, d->Din,
                                      3 * D, R, 1.0f, st,
                                      pk_dq ? GEMM_B_PACKED : (rnd ? GEMM_B_ROUNDED : GEMM_B_RAW)));
  // dX = dQKV Wqkv^T
  if (d_table != nullptr || d_x != nullptr) {
    float* dx = d_x ? d_x : ws.dx;
    EBK_PROF(T_QKV_DGRAD, gemm_dispatch(d->math, adq, pk_qkv ? ws.wqkv_d : Wqkv, 3 * D, true, dx, d->Din, R, d->Din,
                                        3 * D, 0.0f, st, pk_qkv ? GEMM_B_PACKED : GEMM_B_RAW));
    if (tok && d_table) EBK_PROF(T_SCATTER, scatter_rows_add(R, d->Din, d->V, tok, dx, drop1, d_table, st));
  }
  if (tok && table_grad_event) EBK_CUDA(cudaEventRecord(table_grad_event, st));
  return EBK_OK;
}

extern "C" int ebk_gemm(int32_t math, int32_t transA, int32_t transB, int32_t M, int32_t N, int32_t K,
                        const float* A, int32_t lda, const float* B, int32_t ldb, float* C, int32_t ldc,
                        float beta, void* stream) {
  EBK_CHECK_ARG(M >= 0 && N >= 0 && K >= 0 && A && B && C, "gemm: bad argument");
  GemmOperandA a{A, lda, transA != 0, nullptr, 0, make_dropout(false, 0.0f, 0), 0};
  return gemm_dispatch(math, a, B, ldb, transB != 0, C, ldc, M, N, K, beta, (cudaStream_t)stream);
}

extern "C" int ebk_gemm_tma(int32_t transA, int32_t transB, int32_t tall, int32_t M, int32_t N, int32_t K,
                            const float* A, int32_t lda, const float* B, int32_t ldb, float* C, int32_t ldc,
                            float beta, float alpha, void* stream) {
  EBK_CHECK_ARG(M >= 0 && N >= 0 && K >= 1 && A && B && C, "gemm_tma: bad argument");
  // tall: bits 0-3 = tile mode (0 / 1 / 15 = auto), bits 4-7 = cluster mode (0 / 1 / 2 / 15 = auto)
  const int tile = (tall & 15) == 15 ? -1 : (tall & 15), cl = ((tall >> 4) & 15) == 15 ? -1 : ((tall >> 4) & 15);
  return gemm_tma(A, lda, transA != 0, B, ldb, transB != 0, C, ldc, M, N, K, beta, alpha, (cudaStream_t)stream, tile,
                  nullptr, cl);
}

extern "C" int ebk_attention_core_fwd(int32_t n_seq, int32_t L, int32_t nh, int32_t dh, const float* qkv,
                                      float* y, void* stream) {
  EBK_CHECK_ARG(qkv && y, "attention_core_fwd: null pointer");
  return attention_core_fwd(n_seq, L, nh, dh, qkv, y, (cudaStream_t)stream);
}

extern "C" int ebk_attention_core_bwd(int32_t n_seq, int32_t L, int32_t nh, int32_t dh, const float* qkv,
                                      const float* dy, float drop_p, uint64_t drop_seed, float* dqkv,
                                      void* stream) {
  EBK_CHECK_ARG(qkv && dy && dqkv, "attention_core_bwd: null pointer");
  return attention_core_bwd(n_seq, L, nh, dh, qkv, dy, make_dropout(drop_p > 0.0f, drop_p, drop_seed), dqkv,
                            false, (cudaStream_t)stream);
}

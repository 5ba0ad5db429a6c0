// NAML building blocks (reference src/ebrec/models/newsrec/naml.py):
//   ebk_attlayer_*  AttLayer2 alone (+ Dropout on its input)          layers.py:55-81, naml.py:79-84,133-138,167-168
//   ebk_conv1d_*    Embedding -> Dropout -> Conv1D("same") + act      naml.py:155-166, 186-197
//   ebk_catview_*   Embedding(n_cat, dim) -> Dense(F, act)            naml.py:205-252
// The contractions (AttLayer2 projection, the convolution as a K = window*E GEMM, their weight and data
// gradients) run on the tcgen05 GEMM of gemm_tf32_sm100.cu; the kernels here are the HBM-bound pieces.
#include "ebk_common.cuh"

namespace ebk {
namespace {

// ------------------------------------------------------------------------------------------------
// AttLayer2
// ------------------------------------------------------------------------------------------------
struct AttWs {
  float *hbuf, *w, *da, *dpre, *colsum, *w_f, *w_f_lo, *w_d;
  // TMA path (EBK_MATH_TF32): tf32(dropout(x)), tf32(W), per-sequence column partials and their reduction scratch
  float *xd, *w_r, *colpart, *colsum2;
  size_t bytes;
};
AttWs att_layout(const ebk_attlayer_desc& d, void* base) {
  size_t off = 0;
  auto take = [&](size_t nfloat) {
    float* p = base ? reinterpret_cast<float*>(reinterpret_cast<char*>(base) + off) : nullptr;
    off += align_up(nfloat * sizeof(float), 256);
    return p;
  };
  const size_t R = (size_t)d.n_seq * d.L;
  AttWs w;
  w.hbuf = take(R * d.att);
  w.w = take(R);
  w.da = take(R);
  w.dpre = take(R * d.att);
  w.colsum = take(colsum_partial_floats((int)R, d.att));
  w.w_f = take(gemm_tf32_packed_floats(d.att, d.D, false));
  w.w_f_lo = take(gemm_tf32_packed_floats(d.att, d.D, false));
  w.w_d = take(gemm_tf32_packed_floats(d.D, d.att, true));
  w.xd = take(R * d.D + 64);
  w.w_r = take((size_t)d.D * d.att + 64);
  w.colpart = take((size_t)d.n_seq * 2 * d.att);
  w.colsum2 = take(colsum_partial_floats(d.n_seq, 2 * d.att));
  w.bytes = off;
  return w;
}
// all-TMA path of the standalone AttLayer2 (same kernels as the sequence encoder's, api.cu)
bool att_tma(const ebk_attlayer_desc& d, const AttWs& ws, int ld_out) {
  return d.math == EBK_MATH_TF32 && d.att % 4 == 0 && ld_out % 4 == 0 &&
         gemm_tma_eligible(ws.xd, d.D, ws.w_r, d.att, 1, 1, 1);
}
int check_att(const ebk_attlayer_desc* d) {
  EBK_CHECK_ARG(d != nullptr, "attlayer: null descriptor");
  EBK_CHECK_ARG(d->n_seq >= 0 && d->L >= 1 && d->L <= 64, "attlayer: need n_seq>=0, 1<=L<=64 (n_seq=%d L=%d)", d->n_seq, d->L);
  EBK_CHECK_ARG(d->D >= 4 && d->D % 4 == 0 && d->att >= 1, "attlayer: D=%d must be a multiple of 4, att=%d >= 1", d->D, d->att);
  EBK_CHECK_ARG(d->dropout >= 0.f && d->dropout < 1.f, "attlayer: dropout=%f outside [0,1)", d->dropout);
  EBK_CHECK_ARG((long)d->n_seq * d->L < (1L << 31), "attlayer: n_seq*L overflows int32");
  return EBK_OK;
}

// ------------------------------------------------------------------------------------------------
// Conv1D text view.  Padded row space: article n owns rows [n*Lp, (n+1)*Lp), Lp = L + window - 1.
//   XG  [(Q + 2G), E]  dropped embeddings, token l of article n at row G + n*Lp + padL + l, zeros elsewhere
//   dzp [Q, F]         dz (gradient at the conv pre-activation) of token l at row n*Lp + padR + l
// with Q = n_seq*Lp, G = window-1 guard rows, padL = (window-1)/2, padR = window-1-padL.
//   fwd   y[n,l]   = sum_j XG[G + n*Lp + l + j] Wc[j]        -> A row = window*E contiguous floats
//   wgrad dWc[j]  += sum_q XG[G - padR + q + j]^T dzp[q]      -> A^T over the whole padded space, no gather
//   dgrad dX[n,l]  = sum_j' dzp[n*Lp + l + j'] Wc[w-1-j']^T   -> A row = window*F contiguous floats
// ------------------------------------------------------------------------------------------------
struct ConvGeom {
  int Lp, padL, padR, G;
  long Q, R;
};
ConvGeom conv_geom(const ebk_conv1d_desc& d) {
  ConvGeom g;
  g.Lp = d.L + d.window - 1;
  g.padL = (d.window - 1) / 2;
  g.padR = d.window - 1 - g.padL;
  g.G = d.window - 1;
  g.Q = (long)d.n_seq * g.Lp;
  g.R = (long)d.n_seq * d.L;
  return g;
}
struct ConvWs {
  float *xg, *dzp, *dx, *wrev, *colsum, *wc_f, *wrev_f;
  // TMA path (EBK_MATH_TF32): outputs over the whole padded row space + tf32 copies of the kernels
  float *yp, *dxp, *wc_r;
  int32_t* gidx;
  size_t bytes;
};
ConvWs conv_layout(const ebk_conv1d_desc& d, void* base) {
  size_t off = 0;
  auto take = [&](size_t nfloat) {
    float* p = base ? reinterpret_cast<float*>(reinterpret_cast<char*>(base) + off) : nullptr;
    off += align_up(nfloat * sizeof(float), 256);
    return p;
  };
  const ConvGeom g = conv_geom(d);
  ConvWs w;
  w.xg = take((size_t)(g.Q + 2 * g.G) * d.E);
  w.dzp = take((size_t)g.Q * d.F);
  w.dx = take((size_t)g.R * d.E);
  w.wrev = take((size_t)d.window * d.F * d.E);
  w.colsum = take(colsum_partial_floats((int)g.Q, d.F));
  w.wc_f = take(gemm_tf32_packed_floats(d.F, d.window * d.E, false));
  w.wrev_f = take(gemm_tf32_packed_floats(d.E, d.window * d.F, false));
  w.gidx = reinterpret_cast<int32_t*>(take((size_t)g.R));
  w.yp = take((size_t)g.Q * d.F + 64);
  w.dxp = take((size_t)g.Q * d.E + 64);
  w.wc_r = take((size_t)d.window * d.E * d.F + 64);
  w.bytes = off;
  return w;
}
// all-TMA path: the convolution windows are rows of an OVERLAPPING strided view (row stride E, row length
// window*E) of the padded embedding buffer -- a plain 2-D tensor map -- evaluated at every padded row; the
// (window-1) junk rows per article are dropped when the bias/activation pass compacts the result.
bool conv_tma(const ebk_conv1d_desc& d, const ConvWs& ws) {
  return d.math == EBK_MATH_TF32 && gemm_tma_eligible(ws.xg, d.E, ws.wc_r, d.F, 1, 1, 1) && d.n_seq * (d.L + d.window - 1) > d.window;
}
int check_conv(const ebk_conv1d_desc* d) {
  EBK_CHECK_ARG(d != nullptr, "conv1d: null descriptor");
  EBK_CHECK_ARG(d->n_seq >= 0 && d->L >= 1 && d->window >= 1 && d->window <= 16, "conv1d: bad n_seq/L/window (%d/%d/%d)",
                d->n_seq, d->L, d->window);
  EBK_CHECK_ARG(d->E >= 4 && d->E % 4 == 0 && d->F >= 4 && d->F % 4 == 0, "conv1d: E=%d, F=%d must be multiples of 4", d->E, d->F);
  EBK_CHECK_ARG(d->V >= 1, "conv1d: V=%d", d->V);
  EBK_CHECK_ARG(d->dropout >= 0.f && d->dropout < 1.f, "conv1d: dropout=%f outside [0,1)", d->dropout);
  EBK_CHECK_ARG((long)d->n_seq * (d->L + d->window - 1) + 2 * d->window < (1L << 31), "conv1d: padded row count overflows int32");
  return EBK_OK;
}

// XG rows (incl. guard and pad rows, which are zero): 128-bit gather of table rows, dropout mask applied
__global__ void embed_pad_kernel(float4* __restrict__ xg, const int32_t* __restrict__ tok, const float4* __restrict__ table,
                                 long rows, long Q, int L, int Lp, int padL, int G, int E4, int V, Dropout drop,
                                 int round_out) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= rows * E4) return;
  const long row = i / E4;
  const int c4 = (int)(i - row * E4);
  const long q = row - G;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (q >= 0 && q < Q) {
    const long n = q / Lp;
    const int l = (int)(q - n * Lp) - padL;
    if (l >= 0 && l < L) {
      const long r = n * L + l;
      const int t = __ldg(tok + r);
      if (t >= 0 && t < V) {
        v = __ldg(table + (long)t * E4 + c4);
        if (drop.on()) {
          const float4 f = drop.factor4_group((uint64_t)(r * E4 + c4));
          v.x *= f.x; v.y *= f.y; v.z *= f.z; v.w *= f.w;
        }
        if (round_out) {
          v.x = round_tf32_bits(v.x); v.y = round_tf32_bits(v.y); v.z = round_tf32_bits(v.z); v.w = round_tf32_bits(v.w);
        }
      }
    }
  }
  xg[i] = v;
}

// y[r, :] = act(yp[gidx[r], :] + b): drops the junk rows of the padded-space convolution
__global__ void bias_act_gather_kernel(float4* __restrict__ y, const float4* __restrict__ yp, const int32_t* __restrict__ gidx,
                                       const float4* __restrict__ b, long R, int F4, int relu) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= R * F4) return;
  const long r = i / F4;
  const int c4 = (int)(i - r * F4);
  float4 v = yp[(long)gidx[r] * F4 + c4];
  const float4 bb = b[c4];
  v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
  if (relu) {
    v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
  }
  y[i] = v;
}

// dx[r, :] = dxp[gidx[r], :]   (compaction of the padded-space data gradient)
__global__ void gather_rows_kernel(float4* __restrict__ dx, const float4* __restrict__ dxp, const int32_t* __restrict__ gidx,
                                   long R, int E4, long rows_p) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= R * E4) return;
  const long r = i / E4;
  const long q = gidx[r];
  dx[i] = q < rows_p ? dxp[q * E4 + (i - r * E4)] : make_float4(0.f, 0.f, 0.f, 0.f);
}

// dzp rows: dz = dy * dropout_out' * act'(y) at the token rows, zero at the pad rows
__global__ void conv_dz_kernel(float4* __restrict__ dzp, const float4* __restrict__ dy, const float4* __restrict__ y, long Q,
                               int L, int Lp, int padR, int F4, Dropout drop, int relu, int round_out) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= Q * F4) return;
  const long q = i / F4;
  const int c4 = (int)(i - q * F4);
  const long n = q / Lp;
  const int l = (int)(q - n * Lp) - padR;
  float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
  if (l >= 0 && l < L) {
    const long src = (n * L + l) * F4 + c4;
    g = dy[src];
    if (drop.on()) {
      const float4 f = drop.factor4_group((uint64_t)src);
      g.x *= f.x; g.y *= f.y; g.z *= f.z; g.w *= f.w;
    }
    if (relu) {
      const float4 yv = y[src];
      if (!(yv.x > 0.f)) g.x = 0.f;
      if (!(yv.y > 0.f)) g.y = 0.f;
      if (!(yv.z > 0.f)) g.z = 0.f;
      if (!(yv.w > 0.f)) g.w = 0.f;
    }
    if (round_out) {
      g.x = round_tf32_bits(g.x); g.y = round_tf32_bits(g.y); g.z = round_tf32_bits(g.z); g.w = round_tf32_bits(g.w);
    }
  }
  dzp[i] = g;
}

__global__ void conv_gidx_kernel(int32_t* __restrict__ gidx, long R, int L, int Lp) {
  const long r = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (r >= R) return;
  const long n = r / L;
  gidx[r] = (int32_t)(n * Lp + (r - n * L));
}

// wrev[(j'*F + f)*E + e] = Wc[((w-1-j')*E + e)*F + f]
__global__ void conv_wrev_kernel(float* __restrict__ wrev, const float* __restrict__ Wc, int w, int E, int F) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= (long)w * E * F) return;
  const int e = (int)(i % E);
  const long t = i / E;
  const int f = (int)(t % F), jp = (int)(t / F);
  wrev[i] = Wc[((long)(w - 1 - jp) * E + e) * F + f];
}

// ------------------------------------------------------------------------------------------------
// Categorical view
// ------------------------------------------------------------------------------------------------
// T[c, f] = act(b[f] + sum_k emb[c,k] W[k,f]), c < n_cat;  T[n_cat, f] = act(b[f])
__global__ void cat_table_kernel(float* __restrict__ T, const float* __restrict__ emb, const float* __restrict__ W,
                                 const float* __restrict__ b, int n_cat, int dim, int F, int relu) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  const int c = blockIdx.y;
  if (f >= F) return;
  float acc = b[f];
  if (c < n_cat)
    for (int k = 0; k < dim; ++k) acc = fmaf(emb[(long)c * dim + k], W[(long)k * F + f], acc);
  T[(long)c * F + f] = relu ? fmaxf(acc, 0.f) : acc;
}
__global__ void cat_gather_kernel(float* __restrict__ out, int out_ld, const float* __restrict__ T, const int32_t* __restrict__ ids,
                                  long N, int n_cat, int F) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= N * F) return;
  const long n = i / F;
  const int f = (int)(i - n * F);
  int c = ids[n];
  if (c < 0 || c >= n_cat) c = n_cat;
  out[n * out_ld + f] = T[(long)c * F + f];
}
// G[c, f] = sum_{n: id_n == c} d_out[n, f] * act'(T[c, f])   (deterministic: one thread per (c, f))
__global__ void cat_segsum_kernel(float* __restrict__ G, const float* __restrict__ T, const float* __restrict__ d_out,
                                  int dout_ld, const int32_t* __restrict__ ids, long N, int n_cat, int F, int relu) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  const int c = blockIdx.y;
  if (f >= F) return;
  float acc = 0.f;
  if (!relu || T[(long)c * F + f] > 0.f) {
    for (long n = 0; n < N; ++n) {
      int id = __ldg(ids + n);
      if (id < 0 || id >= n_cat) id = n_cat;
      if (id == c) acc += d_out[n * dout_ld + f];
    }
  }
  G[(long)c * F + f] = acc;
}
// db[f] += sum_c G[c,f] (incl. the out-of-range row); dW[k,f] += sum_{c<n_cat} emb[c,k] G[c,f]
__global__ void cat_wgrad_kernel(const float* __restrict__ G, const float* __restrict__ emb, int n_cat, int dim, int F,
                                 float* __restrict__ dW, float* __restrict__ db) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  const int k = blockIdx.y;  // k == dim -> bias
  if (f >= F) return;
  float acc = 0.f;
  if (k == dim) {
    for (int c = 0; c <= n_cat; ++c) acc += G[(long)c * F + f];
    db[f] += acc;
  } else {
    for (int c = 0; c < n_cat; ++c) acc = fmaf(emb[(long)c * dim + k], G[(long)c * F + f], acc);
    dW[(long)k * F + f] += acc;
  }
}
// d_emb[c,k] += sum_f G[c,f] W[k,f]   (one warp per (c,k))
__global__ void cat_dgrad_kernel(const float* __restrict__ G, const float* __restrict__ W, int n_cat, int dim, int F,
                                 float* __restrict__ d_emb) {
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (wid >= n_cat * dim) return;
  const int c = wid / dim, k = wid % dim;
  float acc = 0.f;
  for (int f = lane; f < F; f += 32) acc = fmaf(G[(long)c * F + f], W[(long)k * F + f], acc);
  acc = warp_sum(acc);
  if (lane == 0) d_emb[(long)c * dim + k] += acc;
}

}  // namespace
}  // namespace ebk

using namespace ebk;

// ====================================================================================================
extern "C" size_t ebk_attlayer_workspace_bytes(const ebk_attlayer_desc* d) {
  if (check_att(d) != EBK_OK) return 0;
  return att_layout(*d, nullptr).bytes;
}

extern "C" int ebk_attlayer_fwd(const ebk_attlayer_desc* d, const float* x, const float* W, const float* b, const float* q,
                                int training, uint64_t seed, void* workspace, size_t workspace_bytes, float* out,
                                int32_t out_ld, void* stream) {
  EBK_TRY(check_att(d));
  if (d->n_seq == 0) return EBK_OK;
  EBK_CHECK_ARG(x && W && b && q && out && workspace, "attlayer_fwd: null pointer");
  EBK_CHECK_ARG(out_ld >= d->D, "attlayer_fwd: out_ld=%d < D=%d", out_ld, d->D);
  AttWs ws = att_layout(*d, workspace);
  if (workspace_bytes < ws.bytes) {
    set_error("attlayer_fwd: workspace %zu < %zu bytes", workspace_bytes, ws.bytes);
    return EBK_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int R = d->n_seq * d->L, D = d->D;
  const Dropout drop = make_dropout(training != 0, d->dropout, seed);
  if (att_tma(*d, ws, out_ld)) {
    // tf32(dropout(x)) is materialised once; the GEMM and the pooling read it without touching the mask again
    const Dropout none = make_dropout(false, 0.f, 0);
    EBK_TRY(embed_rows(R, D, 0, nullptr, x, drop, ws.xd, st));
    EBK_TRY(round_tf32_copy(ws.w_r, W, (size_t)D * d->att, st));
    EBK_PROF(T_ATT_GEMM_FWD, gemm_tma(ws.xd, D, false, ws.w_r, d->att, false, ws.hbuf, d->att, R, d->att, D, 0.0f, 1.0f, st, -1));
    EBK_PROF(T_POOL_FWD, attpool_fwd(d->n_seq, d->L, D, d->att, ws.xd, none, ws.hbuf, b, q, ws.w, out, st, out_ld));
    return EBK_OK;
  }
  const bool tc = d->math != EBK_MATH_FP32, x3 = d->math == EBK_MATH_TF32X3;
  GemmOperandA ax{x, D, false, nullptr, 0, drop, D};
  const bool pk = tc && gemm_tf32_eligible(ax, W, d->att, R, d->att, D);
  if (pk) {
    EBK_TRY(gemm_tf32_pack_b(ws.w_f, x3 ? ws.w_f_lo : nullptr, W, d->att, false, d->att, D, st));
    if (!x3) EBK_TRY(gemm_tf32_pack_b(ws.w_d, nullptr, W, d->att, true, D, d->att, st));
  }
  // pre-activation dropout(x) . W                                              layers.py:65
  EBK_PROF(T_ATT_GEMM_FWD, gemm_dispatch(d->math, ax, pk ? ws.w_f : W, d->att, false, ws.hbuf, d->att, R, d->att, D, 0.0f, st,
                                         pk ? GEMM_B_PACKED : GEMM_B_RAW, (pk && x3) ? ws.w_f_lo : nullptr));
  // tanh, .q, exp, normalise (+1e-7), pool                                      layers.py:65-81
  EBK_PROF(T_POOL_FWD, attpool_fwd(d->n_seq, d->L, D, d->att, x, drop, ws.hbuf, b, q, ws.w, out, st, out_ld));
  return EBK_OK;
}

extern "C" int ebk_attlayer_bwd(const ebk_attlayer_desc* d, const float* x, const float* W, const float* q, int training,
                                uint64_t seed, void* workspace, size_t workspace_bytes, const float* d_out,
                                int32_t d_out_ld, float* dW, float* db, float* dq, float* dx, void* stream) {
  EBK_TRY(check_att(d));
  if (d->n_seq == 0) return EBK_OK;
  EBK_CHECK_ARG(x && W && q && d_out && dW && db && dq && dx && workspace, "attlayer_bwd: null pointer");
  EBK_CHECK_ARG(d_out_ld >= d->D, "attlayer_bwd: d_out_ld=%d < D=%d", d_out_ld, d->D);
  AttWs ws = att_layout(*d, workspace);
  if (workspace_bytes < ws.bytes) {
    set_error("attlayer_bwd: workspace %zu < %zu bytes", workspace_bytes, ws.bytes);
    return EBK_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int R = d->n_seq * d->L, D = d->D;
  const Dropout drop = make_dropout(training != 0, d->dropout, seed);
  const Dropout none = make_dropout(false, 0.f, 0);
  if (att_tma(*d, ws, d_out_ld)) {
    // ws.xd / ws.w_r / ws.hbuf / ws.w are the forward's
    EBK_PROF(T_POOL_BWD, attpool_bwd_fused(d->n_seq, d->L, D, d->att, ws.xd, ws.hbuf, q, ws.w, d_out, ws.da, ws.dpre,
                                           ws.colpart, st, d_out_ld));
    EBK_PROF(T_COLSUM, colsum_accum_ws(d->n_seq, d->att, ws.colpart, 2 * d->att, nullptr, db, ws.colsum2, st));
    EBK_PROF(T_COLSUM, colsum_accum_ws(d->n_seq, d->att, ws.colpart + d->att, 2 * d->att, nullptr, dq, ws.colsum2, st));
    EBK_PROF(T_ATT_WGRAD, gemm_tma(ws.xd, D, true, ws.dpre, d->att, false, dW, d->att, D, d->att, R, 1.0f, 1.0f, st, -1));
    // dx = w_t d_out + dpre W^T (gradient w.r.t. the MASKED input; the producer of x applies the mask)
    const GemmEpilogue dx_epi{ws.w, d_out, d_out_ld, d->L, none, 0, false};
    EBK_PROF(T_ATT_DGRAD, gemm_tma(ws.dpre, d->att, false, ws.w_r, d->att, true, dx, D, R, D, d->att, 0.0f, 1.0f, st, 0,
                                   &dx_epi));
    return EBK_OK;
  }
  const bool tc = d->math != EBK_MATH_FP32, x3 = d->math == EBK_MATH_TF32X3;
  const bool rnd = tc && !x3;
  GemmOperandA adp{ws.dpre, d->att, false, nullptr, 0, none, 0};
  const bool pk = rnd && gemm_tf32_eligible(adp, W, d->att, R, D, d->att) &&
                  gemm_tf32_eligible(GemmOperandA{x, D, false, nullptr, 0, drop, D}, W, d->att, R, d->att, D);
  EBK_PROF(T_POOL_BWD, attpool_bwd(d->n_seq, d->L, D, d->att, x, drop, ws.hbuf, q, ws.w, d_out, ws.da, ws.dpre, dx, rnd, st,
                                   d_out_ld));
  EBK_PROF(T_COLSUM, colsum_accum_ws(R, d->att, ws.hbuf, d->att, ws.da, dq, ws.colsum, st));    // dq = sum_r h_r da_r
  EBK_PROF(T_COLSUM, colsum_accum_ws(R, d->att, ws.dpre, d->att, nullptr, db, ws.colsum, st));  // db = sum_r dpre_r
  GemmOperandA axT{x, D, true, nullptr, 0, drop, D};                                             // dW += X^T dpre
  EBK_PROF(T_ATT_WGRAD, gemm_dispatch(d->math, axT, ws.dpre, d->att, false, dW, d->att, D, d->att, R, 1.0f, st,
                                      rnd ? GEMM_B_ROUNDED : GEMM_B_RAW));
  EBK_PROF(T_ATT_DGRAD, gemm_dispatch(d->math, adp, pk ? ws.w_d : W, d->att, true, dx, D, R, D, d->att, 1.0f, st,
                                      pk ? GEMM_B_PACKED : GEMM_B_RAW));                         // dX += dpre W^T
  return EBK_OK;
}

// ====================================================================================================
extern "C" size_t ebk_conv1d_workspace_bytes(const ebk_conv1d_desc* d) {
  if (check_conv(d) != EBK_OK) return 0;
  return conv_layout(*d, nullptr).bytes;
}

extern "C" int ebk_conv1d_fwd(const ebk_conv1d_desc* d, const int32_t* tok, const float* table, const float* Wc,
                              const float* bc, int training, uint64_t seed_in, void* workspace, size_t workspace_bytes,
                              float* y, void* stream) {
  EBK_TRY(check_conv(d));
  if (d->n_seq == 0) return EBK_OK;
  EBK_CHECK_ARG(tok && table && Wc && bc && y && workspace, "conv1d_fwd: null pointer");
  ConvWs ws = conv_layout(*d, workspace);
  if (workspace_bytes < ws.bytes) {
    set_error("conv1d_fwd: workspace %zu < %zu bytes", workspace_bytes, ws.bytes);
    return EBK_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  prof_set_group(0);
  const ConvGeom g = conv_geom(*d);
  const int E = d->E, F = d->F, KW = d->window * E;
  const Dropout drop = make_dropout(training != 0, d->dropout, seed_in);
  const Dropout none = make_dropout(false, 0.f, 0);
  if (prof_on()) prof_begin(T_EMBED_PAD, st);
  {
    const long rows = g.Q + 2 * g.G, n4 = rows * (E / 4);
    embed_pad_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(reinterpret_cast<float4*>(ws.xg), tok,
                                                                   reinterpret_cast<const float4*>(table), rows, g.Q, d->L,
                                                                   g.Lp, g.padL, g.G, E / 4, d->V, drop,
                                                                   conv_tma(*d, ws) ? 1 : 0);
    EBK_LAUNCH_CHECK();
    conv_gidx_kernel<<<(unsigned)((g.R + 255) / 256), 256, 0, st>>>(ws.gidx, g.R, d->L, g.Lp);
    EBK_LAUNCH_CHECK();
  }
  if (prof_on()) prof_end(T_EMBED_PAD, st);
  if (conv_tma(*d, ws)) {
    EBK_TRY(round_tf32_copy(ws.wc_r, Wc, (size_t)KW * F, st));
    // yp[q] = window(q) . Wc for every padded row q (window(q) = XG rows G+q .. G+q+w-1, contiguous)
    EBK_PROF(T_CONV_FWD, gemm_tma(ws.xg + (size_t)g.G * E, E, false, ws.wc_r, F, false, ws.yp, F, (int)g.Q, F, KW, 0.0f, 1.0f,
                                  st, -1));
    const long n4 = g.R * (F / 4);
    bias_act_gather_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(
        reinterpret_cast<float4*>(y), reinterpret_cast<const float4*>(ws.yp), ws.gidx, reinterpret_cast<const float4*>(bc),
        g.R, F / 4, d->relu);
    EBK_LAUNCH_CHECK();
    return EBK_OK;
  }
  const bool tc = d->math != EBK_MATH_FP32, x3 = d->math == EBK_MATH_TF32X3;
  GemmOperandA ax{ws.xg + (size_t)g.G * E, E, false, ws.gidx, (int)g.Q, none, 0};
  const bool pk = tc && !x3 && gemm_tf32_eligible(ax, Wc, F, (int)g.R, F, KW);
  if (pk) EBK_TRY(gemm_tf32_pack_b(ws.wc_f, nullptr, Wc, F, false, F, KW, st));
  EBK_PROF(T_CONV_FWD, gemm_dispatch(d->math, ax, pk ? ws.wc_f : Wc, F, false, y, F, (int)g.R, F, KW, 0.0f, st,
                                     pk ? GEMM_B_PACKED : GEMM_B_RAW));
  EBK_TRY(bias_act(y, bc, g.R * F, F, d->relu, st));
  return EBK_OK;
}

extern "C" int ebk_conv1d_bwd(const ebk_conv1d_desc* d, const int32_t* tok, const float* Wc, const float* y, int training,
                              uint64_t seed_in, uint64_t seed_out, void* workspace, size_t workspace_bytes,
                              const float* dy, float* dWc, float* dbc, float* d_table, void* stream) {
  EBK_TRY(check_conv(d));
  if (d->n_seq == 0) return EBK_OK;
  EBK_CHECK_ARG(tok && Wc && y && dy && dWc && dbc && workspace, "conv1d_bwd: null pointer");
  ConvWs ws = conv_layout(*d, workspace);
  if (workspace_bytes < ws.bytes) {
    set_error("conv1d_bwd: workspace %zu < %zu bytes", workspace_bytes, ws.bytes);
    return EBK_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  prof_set_group(0);
  const ConvGeom g = conv_geom(*d);
  const int E = d->E, F = d->F, w = d->window, KW = w * E;
  const Dropout drop_in = make_dropout(training != 0, d->dropout, seed_in);
  const Dropout drop_out = make_dropout(training != 0, d->dropout, seed_out);
  const Dropout none = make_dropout(false, 0.f, 0);
  const bool tc = d->math != EBK_MATH_FP32, x3 = d->math == EBK_MATH_TF32X3;
  const bool rnd = tc && !x3;
  if (prof_on()) prof_begin(T_CONV_DZ, st);
  {
    const long n4 = g.Q * (F / 4);
    conv_dz_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(reinterpret_cast<float4*>(ws.dzp),
                                                                 reinterpret_cast<const float4*>(dy),
                                                                 reinterpret_cast<const float4*>(y), g.Q, d->L, g.Lp, g.padR,
                                                                 F / 4, drop_out, d->relu, rnd ? 1 : 0);
    EBK_LAUNCH_CHECK();
  }
  if (prof_on()) prof_end(T_CONV_DZ, st);
  EBK_PROF(T_COLSUM, colsum_accum_ws((int)g.Q, F, ws.dzp, F, nullptr, dbc, ws.colsum, st));  // pad rows are zero
  if (conv_tma(*d, ws)) {
    // ws.xg holds tf32(dropout(gather)) from the forward, dzp was rounded by conv_dz_kernel
    EBK_PROF(T_CONV_WGRAD, gemm_tma(ws.xg + (size_t)(g.G - g.padR) * E, E, true, ws.dzp, F, false, dWc, F, KW, F, (int)g.Q, 1.0f,
                                    1.0f, st, -1));
    if (d_table != nullptr) {
      const long nw = (long)w * E * F;
      conv_wrev_kernel<<<(unsigned)((nw + 255) / 256), 256, 0, st>>>(ws.wrev, Wc, w, E, F);
      EBK_LAUNCH_CHECK();
      EBK_TRY(round_tf32_copy(ws.wc_r, ws.wrev, (size_t)nw, st));   // wc_r is free again: reuse it for tf32(wrev)
      // dxp[q] = dz-window(q) . wrev for the padded rows whose window stays inside dzp
      const long rows_p = g.Q - (w - 1);
      EBK_PROF(T_CONV_DGRAD, gemm_tma(ws.dzp, F, false, ws.wc_r, E, false, ws.dxp, E, (int)rows_p, E, w * F, 0.0f, 1.0f, st, -1));
      const long n4 = g.R * (E / 4);
      gather_rows_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(reinterpret_cast<float4*>(ws.dx),
                                                                       reinterpret_cast<const float4*>(ws.dxp), ws.gidx, g.R,
                                                                       E / 4, rows_p);
      EBK_LAUNCH_CHECK();
      EBK_PROF(T_SCATTER, scatter_rows_add((int)g.R, E, d->V, tok, ws.dx, drop_in, d_table, st));
    }
    return EBK_OK;
  }
  // dWc += XG_window^T dzp over the whole padded row space (guard rows make every window readable)
  GemmOperandA axT{ws.xg + (size_t)(g.G - g.padR) * E, E, true, nullptr, 0, none, 0};
  EBK_PROF(T_CONV_WGRAD, gemm_dispatch(d->math, axT, ws.dzp, F, false, dWc, F, KW, F, (int)g.Q, 1.0f, st,
                                       rnd ? GEMM_B_ROUNDED : GEMM_B_RAW));
  if (d_table != nullptr) {
    const long nw = (long)w * E * F;
    conv_wrev_kernel<<<(unsigned)((nw + 255) / 256), 256, 0, st>>>(ws.wrev, Wc, w, E, F);
    EBK_LAUNCH_CHECK();
    GemmOperandA adz{ws.dzp, F, false, ws.gidx, (int)g.Q, none, 0};
    const bool pk = rnd && gemm_tf32_eligible(adz, ws.wrev, E, (int)g.R, E, w * F);
    if (pk) EBK_TRY(gemm_tf32_pack_b(ws.wrev_f, nullptr, ws.wrev, E, false, E, w * F, st));
    EBK_PROF(T_CONV_DGRAD, gemm_dispatch(d->math, adz, pk ? ws.wrev_f : ws.wrev, E, false, ws.dx, E, (int)g.R, E, w * F, 0.0f,
                                         st, pk ? GEMM_B_PACKED : GEMM_B_RAW));
    EBK_PROF(T_SCATTER, scatter_rows_add((int)g.R, E, d->V, tok, ws.dx, drop_in, d_table, st));
  }
  return EBK_OK;
}

// ====================================================================================================
extern "C" size_t ebk_catview_workspace_bytes(int32_t n_cat, int32_t F) {
  if (n_cat < 1 || F < 1) return 0;
  return 2 * align_up((size_t)(n_cat + 1) * F * sizeof(float), 256);
}

extern "C" int ebk_catview_fwd(int32_t N, int32_t n_cat, int32_t dim, int32_t F, int32_t relu, const int32_t* ids,
                               const float* emb, const float* W, const float* b, void* workspace, size_t workspace_bytes,
                               float* out, int32_t out_ld, void* stream) {
  EBK_CHECK_ARG(N >= 0 && n_cat >= 1 && dim >= 1 && F >= 1, "catview_fwd: bad shape N=%d n_cat=%d dim=%d F=%d", N, n_cat, dim, F);
  if (N == 0) return EBK_OK;
  EBK_CHECK_ARG(ids && emb && W && b && out && workspace && out_ld >= F, "catview_fwd: null pointer or out_ld < F");
  if (workspace_bytes < ebk_catview_workspace_bytes(n_cat, F)) {
    set_error("catview_fwd: workspace %zu < %zu bytes", workspace_bytes, ebk_catview_workspace_bytes(n_cat, F));
    return EBK_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  prof_set_group(0);
  float* T = reinterpret_cast<float*>(workspace);
  if (prof_on()) prof_begin(T_CATVIEW, st);
  cat_table_kernel<<<dim3(ceil_div(F, 128), n_cat + 1), 128, 0, st>>>(T, emb, W, b, n_cat, dim, F, relu);
  EBK_LAUNCH_CHECK();
  const long n = (long)N * F;
  cat_gather_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(out, out_ld, T, ids, N, n_cat, F);
  EBK_LAUNCH_CHECK();
  if (prof_on()) prof_end(T_CATVIEW, st);
  return EBK_OK;
}

extern "C" int ebk_catview_bwd(int32_t N, int32_t n_cat, int32_t dim, int32_t F, int32_t relu, const int32_t* ids,
                               const float* emb, const float* W, void* workspace, size_t workspace_bytes,
                               const float* d_out, int32_t d_out_ld, float* d_emb, float* dW, float* db, void* stream) {
  EBK_CHECK_ARG(N >= 0 && n_cat >= 1 && dim >= 1 && F >= 1, "catview_bwd: bad shape N=%d n_cat=%d dim=%d F=%d", N, n_cat, dim, F);
  if (N == 0) return EBK_OK;
  EBK_CHECK_ARG(ids && emb && W && d_out && d_emb && dW && db && workspace && d_out_ld >= F,
                "catview_bwd: null pointer or d_out_ld < F");
  if (workspace_bytes < ebk_catview_workspace_bytes(n_cat, F)) {
    set_error("catview_bwd: workspace %zu < %zu bytes", workspace_bytes, ebk_catview_workspace_bytes(n_cat, F));
    return EBK_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  prof_set_group(0);
  float* T = reinterpret_cast<float*>(workspace);
  float* G = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + align_up((size_t)(n_cat + 1) * F * sizeof(float), 256));
  if (prof_on()) prof_begin(T_CATVIEW, st);
  cat_segsum_kernel<<<dim3(ceil_div(F, 128), n_cat + 1), 128, 0, st>>>(G, T, d_out, d_out_ld, ids, N, n_cat, F, relu);
  EBK_LAUNCH_CHECK();
  cat_wgrad_kernel<<<dim3(ceil_div(F, 128), dim + 1), 128, 0, st>>>(G, emb, n_cat, dim, F, dW, db);
  EBK_LAUNCH_CHECK();
  cat_dgrad_kernel<<<ceil_div(n_cat * dim * 32, 128), 128, 0, st>>>(G, W, n_cat, dim, F, d_emb);
  EBK_LAUNCH_CHECK();
  if (prof_on()) prof_end(T_CATVIEW, st);
  return EBK_OK;
}

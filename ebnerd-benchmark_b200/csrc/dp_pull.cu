// Data-parallel optimizer step of the rank-sharded embedding table with the gradient reduction FUSED into the Adam
// pass over NVLink peer memory (no reference counterpart: the reference is single-device; replaces the
// ncclReduceScatter of the 768 MB fp32 table gradient + the Adam pass over this rank's shard).
//
// Every rank scatters its row-sparse table gradient into its own dense [V, E] buffer (scatter_rows_add) and publishes
// one byte per table row: "this rank touched row v this step" (ebk_dp_token_flags, all-gathered -- 250 KB per rank).
// The owner of a shard then runs ONE kernel over its 1/world slice of theta / m / v: for every 16-byte chunk it adds
// the gradient chunks of exactly those ranks whose flag for the row is set -- local chunk from HBM, remote chunks
// through the peers' CUDA-IPC mappings over NVLink, in fixed rank order (deterministic) -- and applies the Keras-form
// Adam update (identical arithmetic to ebk_adam_keras_step).  Wire volume = touched rows only (54 % of the table for
// uniform synthetic tokens, a few % for Zipfian text) instead of the whole dense table, and the transfer overlaps the
// update arithmetic chunk by chunk instead of preceding it.
#include "ebk_common.cuh"

namespace ebk {
namespace {

struct PullPeers {
  const float4* g[8];      // table-gradient buffers of every rank (this process's mappings; g[rank] is local)
};

__global__ void token_flags_kernel(int R, int V, const int32_t* __restrict__ tok, uint8_t* __restrict__ flags) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  const int t = tok[r];
  if (t >= 0 && t < V) flags[t] = 1;
}

template <int WORLD>
__global__ void __launch_bounds__(256) adam_pull_kernel(float4* __restrict__ theta, float4* __restrict__ m, float4* __restrict__ v,
                                                         float4* __restrict__ g_local, PullPeers peers,
                                                         const uint8_t* __restrict__ flags, size_t v_pad, int rank, int E4,
                                                         size_t lo4, size_t n4, float alpha, const float* __restrict__ alpha_dev,
                                                         float omb1, float omb2, float eps) {
  if (alpha_dev != nullptr) alpha = __ldg(alpha_dev);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const size_t gi = lo4 + i;                 // float4 index inside the table
    const size_t row = gi / (size_t)E4;
    float4 part[WORLD];
    bool on[WORLD];
    // issue every needed load first (remote ones take microseconds), then reduce in rank order
#pragma unroll
    for (int p = 0; p < WORLD; ++p) {
      on[p] = flags[(size_t)p * v_pad + row] != 0;
      part[p] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (on[p]) {
        const float4* src = peers.g[p] + gi;
        asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                     : "=f"(part[p].x), "=f"(part[p].y), "=f"(part[p].z), "=f"(part[p].w)
                     : "l"(src));
      }
    }
    float4 th = theta[gi], mm = m[gi], vv = v[gi];
    float4 gg = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int p = 0; p < WORLD; ++p) {
      gg.x += part[p].x; gg.y += part[p].y; gg.z += part[p].z; gg.w += part[p].w;
    }
#define UPD(c)                                  \
  mm.c += (gg.c - mm.c) * omb1;                 \
  vv.c += (gg.c * gg.c - vv.c) * omb2;          \
  th.c -= (mm.c * alpha) / (sqrtf(vv.c) + eps);
    UPD(x) UPD(y) UPD(z) UPD(w)
#undef UPD
    theta[gi] = th;
    m[gi] = mm;
    v[gi] = vv;
    if (on[rank]) g_local[gi] = make_float4(0.f, 0.f, 0.f, 0.f);   // this rank's own contribution is consumed
  }
}

}  // namespace
}  // namespace ebk

using namespace ebk;

extern "C" int ebk_dp_token_flags(int32_t R, int32_t V, const int32_t* tok, uint8_t* flags, size_t v_pad, void* stream) {
  EBK_CHECK_ARG(R >= 0 && V >= 1 && flags && (R == 0 || tok) && v_pad >= (size_t)V, "dp_token_flags: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  EBK_CUDA(cudaMemsetAsync(flags, 0, v_pad, st));
  if (R > 0) {
    token_flags_kernel<<<(R + 255) / 256, 256, 0, st>>>(R, V, tok, flags);
    EBK_LAUNCH_CHECK();
  }
  return EBK_OK;
}

extern "C" int ebk_adam_pull_step(float* theta, float* m, float* v, const void* const* grads, const uint8_t* flags_all,
                                  size_t v_pad, int32_t world, int32_t rank, int32_t E, size_t lo_float, size_t n_float,
                                  float alpha, const ebk_step_params* step_dev, double beta1, double beta2, float eps,
                                  void* stream) {
  EBK_CHECK_ARG(theta && m && v && grads && flags_all, "adam_pull: null pointer");
  EBK_CHECK_ARG(world >= 2 && world <= 8 && rank >= 0 && rank < world, "adam_pull: world=%d rank=%d", world, rank);
  EBK_CHECK_ARG(E >= 4 && E % 4 == 0 && lo_float % 4 == 0 && n_float % 4 == 0, "adam_pull: E, shard offset and size must be multiples of 4");
  if (n_float == 0) return EBK_OK;
  PullPeers peers;
  for (int p = 0; p < 8; ++p) {
    peers.g[p] = p < world ? reinterpret_cast<const float4*>(grads[p]) : nullptr;
    EBK_CHECK_ARG(p >= world || (grads[p] != nullptr && (reinterpret_cast<uintptr_t>(grads[p]) & 15) == 0), "adam_pull: gradient mapping %d", p);
  }
  cudaStream_t st = (cudaStream_t)stream;
  const float omb1 = (float)(1.0 - beta1), omb2 = (float)(1.0 - beta2);
  const size_t n4 = n_float / 4, lo4 = lo_float / 4;
  const float* alpha_dev = step_dev ? &step_dev->alpha : nullptr;
  size_t blocks = (n4 + 255) / 256;
  const size_t cap = 148 * 8;
  const unsigned grid = (unsigned)(blocks < cap ? blocks : cap);
  float4* gl = reinterpret_cast<float4*>(const_cast<void*>(grads[rank]));
  if (prof_on()) prof_begin(T_ADAM, st);
#define RUN(W_)                                                                                                          \
  adam_pull_kernel<W_><<<grid, 256, 0, st>>>(reinterpret_cast<float4*>(theta), reinterpret_cast<float4*>(m),            \
                                             reinterpret_cast<float4*>(v), gl, peers, flags_all, v_pad, rank, E / 4, lo4, \
                                             n4, alpha, alpha_dev, omb1, omb2, eps)
  switch (world) {
    case 2: RUN(2); break;
    case 3: RUN(3); break;
    case 4: RUN(4); break;
    case 5: RUN(5); break;
    case 6: RUN(6); break;
    case 7: RUN(7); break;
    default: RUN(8); break;
  }
#undef RUN
  if (prof_on()) prof_end(T_ADAM, st);
  EBK_LAUNCH_CHECK();
  return EBK_OK;
}

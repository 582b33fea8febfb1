// Aggregation / gather kernels (sm_100a), HBM-bound.
//
//   agg_csr_sum : out[v] = sum_{k in row v, ascending} src[eid[k]]   (scatter_sum, models/GNN.py:3-21, :99;
//                 also the backward of the x[row] / x[col] gathers with the source-/destination-side CSR)
//   gather_rows : out[m] = src[idx[m]]                               (x[row], x[col]; backward of scatter_sum)
//   edge_geometry: [pos[dst]-pos[src], L1]                           (models/GNN.py:299-302)
//
// Rows are summed SEQUENTIALLY in ascending edge id - the order the CPU reference's
// index_add_ uses - so the result is bit-identical to the reference whatever the
// in-degree.  Parallelism comes from rows x feature lanes: a group of LPR lanes owns
// one row and each lane owns 4*VPL consecutive features (128-bit accesses).  Up to
// 4 neighbour rows are in flight per lane before the ordered adds retire them.
#include "common.cuh"

namespace gnc {

template <int LPR, int VPL, bool ACC>
__device__ __forceinline__ void agg_rows(
    const int32_t* __restrict__ rowptr, const int32_t* __restrict__ eid, const float* __restrict__ src,
    int64_t ld_src, int64_t N, int D4 /* D/4 */, float* __restrict__ out, int64_t ld_out, int64_t warp_global,
    int64_t warps_total) {
  constexpr int ROWS_PER_WARP = 32 / LPR;
  const int lane = threadIdx.x & 31;
  const int sub = lane / LPR, sl = lane % LPR;
  for (int64_t v0 = warp_global * ROWS_PER_WARP; v0 < N; v0 += warps_total * ROWS_PER_WARP) {
    const int64_t v = v0 + sub;
    if (v >= N) continue;
    const int32_t beg = __ldg(rowptr + v), end = __ldg(rowptr + v + 1);
    float4 acc[VPL];
#pragma unroll
    for (int q = 0; q < VPL; ++q) acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    int32_t k = beg;
    for (; k + 4 <= end; k += 4) {
      int32_t e[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) e[u] = __ldg(eid + k + u);
      float4 t[4][VPL];
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int q = 0; q < VPL; ++q) {
          const int c4 = sl + q * LPR;
          t[u][q] = (c4 < D4) ? ldg_stream(reinterpret_cast<const float4*>(src + (int64_t)e[u] * ld_src) + c4)
                              : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int q = 0; q < VPL; ++q) {
          acc[q].x += t[u][q].x; acc[q].y += t[u][q].y; acc[q].z += t[u][q].z; acc[q].w += t[u][q].w;
        }
    }
    {  // tail of up to 3 rows, still loaded together
      const int rem = end - k;
      int32_t e[3];
      float4 t[3][VPL];
#pragma unroll
      for (int u = 0; u < 3; ++u) e[u] = (u < rem) ? __ldg(eid + k + u) : 0;
#pragma unroll
      for (int u = 0; u < 3; ++u)
#pragma unroll
        for (int q = 0; q < VPL; ++q) {
          const int c4 = sl + q * LPR;
          t[u][q] = (u < rem && c4 < D4)
                        ? ldg_stream(reinterpret_cast<const float4*>(src + (int64_t)e[u] * ld_src) + c4)
                        : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
      for (int u = 0; u < 3; ++u)
        if (u < rem) {
#pragma unroll
          for (int q = 0; q < VPL; ++q) {
            acc[q].x += t[u][q].x; acc[q].y += t[u][q].y; acc[q].z += t[u][q].z; acc[q].w += t[u][q].w;
          }
        }
    }
#pragma unroll
    for (int q = 0; q < VPL; ++q) {
      const int c4 = sl + q * LPR;
      if (c4 < D4) {
        float4* o = reinterpret_cast<float4*>(out + v * ld_out) + c4;
        if (ACC) {
          const float4 p = *o;
          acc[q].x += p.x; acc[q].y += p.y; acc[q].z += p.z; acc[q].w += p.w;
        }
        stg_stream(o, acc[q]);
      }
    }
  }
}

template <int LPR, int VPL, bool ACC>
__global__ void __launch_bounds__(256) agg_csr_sum_vec_kernel(
    const int32_t* __restrict__ rowptr, const int32_t* __restrict__ eid, const float* __restrict__ src,
    int64_t ld_src, int64_t N, int D4 /* D/4 */, float* __restrict__ out, int64_t ld_out) {
  agg_rows<LPR, VPL, ACC>(rowptr, eid, src, ld_src, N, D4, out, ld_out, ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5,
                          ((int64_t)gridDim.x * blockDim.x) >> 5);
}

// Two segmented sums over the SAME rows in one launch (the backward of the x[row] / x[col] gathers: the edge gradient summed
// by source and by destination, models/GNN.py:57-58): even warps take CSR A, odd warps CSR B, both walk the node rows in
// step, so on locality-preserving topologies (grids) a source row fetched for one sum is still in L2 for the other - one
// pass over the [E, D] tensor from DRAM instead of two.  Same per-row arithmetic as the single form (same bits).
template <int LPR, int VPL>
__global__ void __launch_bounds__(256) agg_csr_sum_pair_kernel(
    const int32_t* __restrict__ rowptrA, const int32_t* __restrict__ eidA, float* __restrict__ outA,
    const int32_t* __restrict__ rowptrB, const int32_t* __restrict__ eidB, float* __restrict__ outB,
    const float* __restrict__ src, int64_t ld_src, int64_t N, int D4, int64_t ld_out) {
  const int64_t wg = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, wt = ((int64_t)gridDim.x * blockDim.x) >> 5;
  if (wg & 1) agg_rows<LPR, VPL, false>(rowptrB, eidB, src, ld_src, N, D4, outB, ld_out, wg >> 1, wt >> 1);
  else agg_rows<LPR, VPL, false>(rowptrA, eidA, src, ld_src, N, D4, outA, ld_out, wg >> 1, wt >> 1);
}

// Any D / alignment: one thread per (row, feature).
template <bool ACC>
__global__ void agg_csr_sum_scalar_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ eid,
                                          const float* __restrict__ src, int64_t ld_src, int64_t N, int D,
                                          float* __restrict__ out, int64_t ld_out) {
  const int64_t total = N * D;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t v = i / D;
    const int c = (int)(i - v * D);
    float acc = 0.f;
    for (int32_t k = rowptr[v]; k < rowptr[v + 1]; ++k) acc += src[(int64_t)eid[k] * ld_src + c];
    if (ACC) acc += out[v * ld_out + c];
    out[v * ld_out + c] = acc;
  }
}

template <int LPR, int VPL, bool ACC>
__global__ void __launch_bounds__(256) gather_rows_vec_kernel(const float* __restrict__ src, int64_t ld_src,
                                                              const int32_t* __restrict__ idx, int64_t M, int D4,
                                                              float* __restrict__ out, int64_t ld_out) {
  constexpr int ROWS_PER_WARP = 32 / LPR;
  const int lane = threadIdx.x & 31;
  const int sub = lane / LPR, sl = lane % LPR;
  const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t warps_total = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t m0 = warp_global * ROWS_PER_WARP; m0 < M; m0 += warps_total * ROWS_PER_WARP) {
    const int64_t m = m0 + sub;
    if (m >= M) continue;
    const int64_t r = __ldg(idx + m);
#pragma unroll
    for (int q = 0; q < VPL; ++q) {
      const int c4 = sl + q * LPR;
      if (c4 < D4) {
        float4 t = __ldg(reinterpret_cast<const float4*>(src + r * ld_src) + c4);
        float4* o = reinterpret_cast<float4*>(out + m * ld_out) + c4;
        if (ACC) {
          const float4 p = *o;
          t.x += p.x; t.y += p.y; t.z += p.z; t.w += p.w;
        }
        stg_stream(o, t);
      }
    }
  }
}

template <bool ACC>
__global__ void gather_rows_scalar_kernel(const float* __restrict__ src, int64_t ld_src,
                                          const int32_t* __restrict__ idx, int64_t M, int D,
                                          float* __restrict__ out, int64_t ld_out) {
  const int64_t total = M * D;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t m = i / D;
    const int c = (int)(i - m * D);
    float t = src[(int64_t)idx[m] * ld_src + c];
    if (ACC) t += out[m * ld_out + c];
    out[m * ld_out + c] = t;
  }
}

__global__ void edge_geometry_kernel(const float* __restrict__ pos, int P, const int32_t* __restrict__ src,
                                     const int32_t* __restrict__ dst, int64_t E, float* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < E; e += stride) {
    const float* ps = pos + (int64_t)src[e] * P;
    const float* pd = pos + (int64_t)dst[e] * P;
    float l1 = 0.f;
    float* o = out + e * (P + 1);
    for (int d = 0; d < P; ++d) {       // torch.sum over dim 1 of |diff|: left-to-right for P <= 8
      const float r = pd[d] - ps[d];
      o[d] = r;
      l1 += fabsf(r);
    }
    o[P] = l1;
  }
}

// out[m, :] = act( sum_s T_s[idx_s[m], :] + bias )  for up to 3 gathered tables; D = 4 * D4 <= 128 * 4.
// Used when an operand of a linear layer is a lookup into a small table (edge classes of grid
// graphs): the product with the weights is taken once per table row, and the layer becomes this
// streaming gather-add (see GraphNet._forward_tc).
struct GatherAddArgs {
  const float* tab[4];
  const int32_t* idx[4];
  long long ld[4];
  int nsrc;
  const float* bias;
  int relu;
  float* out;
  long long ld_out;
  long long M;
  int D4;
};

__global__ void __launch_bounds__(256) gather_add_rows_kernel(const GatherAddArgs a) {
  const int lane = threadIdx.x & 31;
  const long long warp_global = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long warps_total = ((long long)gridDim.x * blockDim.x) >> 5;
  for (int c4 = lane; c4 < a.D4; c4 += 32) {
    const float4 b = a.bias ? __ldg(reinterpret_cast<const float4*>(a.bias) + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
    for (long long m0 = warp_global * 4; m0 < a.M; m0 += warps_total * 4) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = b;
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        if (s < a.nsrc) {
          float4 t[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const long long m = m0 + u < a.M ? m0 + u : a.M - 1;
            const long long r = a.idx[s] ? (long long)__ldg(a.idx[s] + m) : m;
            t[u] = a.idx[s] ? __ldg(reinterpret_cast<const float4*>(a.tab[s] + r * a.ld[s]) + c4)
                            : ldg_stream(reinterpret_cast<const float4*>(a.tab[s] + r * a.ld[s]) + c4);
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) { v[u].x += t[u].x; v[u].y += t[u].y; v[u].z += t[u].z; v[u].w += t[u].w; }
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (m0 + u < a.M) {
          if (a.relu) { v[u].x = fmaxf(v[u].x, 0.f); v[u].y = fmaxf(v[u].y, 0.f); v[u].z = fmaxf(v[u].z, 0.f); v[u].w = fmaxf(v[u].w, 0.f); }
          stg_stream(reinterpret_cast<float4*>(a.out + (m0 + u) * a.ld_out) + c4, v[u]);
        }
      }
    }
  }
}

// Readout of a batch of graphs with DIFFERENT node counts into the dense [B, num_nodes] input of the classifier head
// (reference models/GNN.py:331, 339-340 flattens [N, 1] and needs N == num_nodes exactly; main.py:65-66 sizes the head
// by resize_value // 2 for superpixel graphs whose node count varies).  Graph b contributes its first
// min(n_b, num_nodes) node outputs; missing entries are zero.  Equal to the reference's flatten when n_b == num_nodes.
__global__ void segment_readout_kernel(const float* __restrict__ y, const int32_t* __restrict__ node_ptr, int B, int num_nodes,
                                       float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * num_nodes) return;
  const int b = (int)(i / num_nodes), j = (int)(i - (long long)b * num_nodes);
  const int lo = node_ptr[b], n = node_ptr[b + 1] - lo;
  out[i] = j < n ? y[lo + j] : 0.f;
}
__global__ void segment_readout_bwd_kernel(const float* __restrict__ dout, const int32_t* __restrict__ node_ptr, int B,
                                           int num_nodes, float* __restrict__ dy) {
  // one thread per (graph, local node): nodes past num_nodes were truncated and receive zero
  const int b = blockIdx.y;
  const int lo = node_ptr[b], n = node_ptr[b + 1] - lo;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x)
    dy[lo + j] = j < num_nodes ? dout[(long long)b * num_nodes + j] : 0.f;
}

static inline unsigned row_grid(int64_t rows, int rows_per_block) {
  int64_t blocks = ceil_div<int64_t>(rows, rows_per_block);
  const int64_t cap = (int64_t)kNumSMs * 8;     // 8 resident 256-thread CTAs per SM
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (unsigned)blocks;
}

template <bool ACC>
static int launch_agg(const int32_t* rowptr, const int32_t* eid, const float* src, int64_t ld_src, int64_t N, int D,
                      float* out, int64_t ld_out, cudaStream_t st) {
  const bool vec = (D % 4 == 0) && (ld_src % 4 == 0) && (ld_out % 4 == 0) && aligned16(src) && aligned16(out) &&
                   D <= 512;
  if (!vec) {
    int64_t blocks = ceil_div<int64_t>(N * D, 256);
    if (blocks > (int64_t)kNumSMs * 8) blocks = (int64_t)kNumSMs * 8;
    agg_csr_sum_scalar_kernel<ACC><<<(unsigned)(blocks < 1 ? 1 : blocks), 256, 0, st>>>(rowptr, eid, src, ld_src,
                                                                                        N, D, out, ld_out);
    return check_launch("agg_csr_sum_scalar_kernel");
  }
  const int D4 = D / 4;
#define GNC_AGG(LPR, VPL)                                                                                  \
  agg_csr_sum_vec_kernel<LPR, VPL, ACC><<<row_grid(N, 8 * (32 / LPR)), 256, 0, st>>>(rowptr, eid, src, ld_src, N, \
                                                                                     D4, out, ld_out)
  if (D4 <= 4) GNC_AGG(4, 1);
  else if (D4 <= 8) GNC_AGG(8, 1);
  else if (D4 <= 16) GNC_AGG(16, 1);
  else if (D4 <= 32) GNC_AGG(32, 1);
  else if (D4 <= 64) GNC_AGG(32, 2);
  else GNC_AGG(32, 4);
#undef GNC_AGG
  return check_launch("agg_csr_sum_vec_kernel");
}

template <bool ACC>
static int launch_gather(const float* src, int64_t ld_src, const int32_t* idx, int64_t M, int D, float* out,
                         int64_t ld_out, cudaStream_t st) {
  const bool vec = (D % 4 == 0) && (ld_src % 4 == 0) && (ld_out % 4 == 0) && aligned16(src) && aligned16(out) &&
                   D <= 512;
  if (!vec) {
    int64_t blocks = ceil_div<int64_t>(M * D, 256);
    if (blocks > (int64_t)kNumSMs * 8) blocks = (int64_t)kNumSMs * 8;
    gather_rows_scalar_kernel<ACC><<<(unsigned)(blocks < 1 ? 1 : blocks), 256, 0, st>>>(src, ld_src, idx, M, D, out,
                                                                                        ld_out);
    return check_launch("gather_rows_scalar_kernel");
  }
  const int D4 = D / 4;
#define GNC_GATHER(LPR, VPL)                                                                                 \
  gather_rows_vec_kernel<LPR, VPL, ACC><<<row_grid(M, 8 * (32 / LPR)), 256, 0, st>>>(src, ld_src, idx, M, D4, out, \
                                                                                     ld_out)
  if (D4 <= 4) GNC_GATHER(4, 1);
  else if (D4 <= 8) GNC_GATHER(8, 1);
  else if (D4 <= 16) GNC_GATHER(16, 1);
  else if (D4 <= 32) GNC_GATHER(32, 1);
  else if (D4 <= 64) GNC_GATHER(32, 2);
  else GNC_GATHER(32, 4);
#undef GNC_GATHER
  return check_launch("gather_rows_vec_kernel");
}

}  // namespace gnc

using namespace gnc;

extern "C" {

int gnc_agg_csr_sum_f32(const int32_t* rowptr, const int32_t* eid, const float* src, int64_t ld_src, int64_t N,
                        int D, float* out, int64_t ld_out, int accumulate, gnc_stream_t stream) {
  GNC_REQUIRE(N >= 0 && D > 0 && ld_src >= D && ld_out >= D, "agg_csr_sum: bad sizes");
  if (N == 0) return GNC_OK;
  GNC_REQUIRE(rowptr && out, "agg_csr_sum: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  return accumulate ? launch_agg<true>(rowptr, eid, src, ld_src, N, D, out, ld_out, st)
                    : launch_agg<false>(rowptr, eid, src, ld_src, N, D, out, ld_out, st);
}

int gnc_agg_csr_sum_pair_f32(const int32_t* rowptr_a, const int32_t* eid_a, float* out_a, const int32_t* rowptr_b,
                             const int32_t* eid_b, float* out_b, const float* src, int64_t ld_src, int64_t N, int D,
                             int64_t ld_out, gnc_stream_t stream) {
  GNC_REQUIRE(N >= 0 && D > 0 && ld_src >= D && ld_out >= D, "agg_csr_sum_pair: bad sizes");
  if (N == 0) return GNC_OK;
  GNC_REQUIRE(rowptr_a && rowptr_b && out_a && out_b, "agg_csr_sum_pair: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec = (D % 4 == 0) && D <= 128 && (ld_src % 4 == 0) && (ld_out % 4 == 0) && aligned16(src) && aligned16(out_a) && aligned16(out_b);
  if (!vec) {                                         // shapes outside the paired kernel: the two sums one after the other
    if (int rc = launch_agg<false>(rowptr_a, eid_a, src, ld_src, N, D, out_a, ld_out, st)) return rc;
    return launch_agg<false>(rowptr_b, eid_b, src, ld_src, N, D, out_b, ld_out, st);
  }
  const int D4 = D / 4;
#define GNC_AGGP(LPR)                                                                                                  \
  agg_csr_sum_pair_kernel<LPR, 1><<<row_grid(2 * N, 8 * (32 / LPR)), 256, 0, st>>>(rowptr_a, eid_a, out_a, rowptr_b, eid_b, out_b, \
                                                                                    src, ld_src, N, D4, ld_out)
  if (D4 <= 4) GNC_AGGP(4);
  else if (D4 <= 8) GNC_AGGP(8);
  else if (D4 <= 16) GNC_AGGP(16);
  else GNC_AGGP(32);
#undef GNC_AGGP
  return check_launch("agg_csr_sum_pair_kernel");
}

int gnc_gather_rows_f32(const float* src, int64_t ld_src, const int32_t* idx, int64_t M, int D, float* out,
                        int64_t ld_out, int accumulate, gnc_stream_t stream) {
  GNC_REQUIRE(M >= 0 && D > 0 && ld_src >= D && ld_out >= D, "gather_rows: bad sizes");
  if (M == 0) return GNC_OK;
  GNC_REQUIRE(src && idx && out, "gather_rows: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  return accumulate ? launch_gather<true>(src, ld_src, idx, M, D, out, ld_out, st)
                    : launch_gather<false>(src, ld_src, idx, M, D, out, ld_out, st);
}

int gnc_gather_add_rows_f32(const float* const* tables, const int32_t* const* idx, const int64_t* ld, int nsrc,
                            const float* bias, int relu, int64_t M, int D, float* out, int64_t ld_out,
                            gnc_stream_t stream) {
  GNC_REQUIRE(nsrc >= 1 && nsrc <= 4 && tables && idx && ld && out && M >= 0 && D > 0 && D % 4 == 0 && ld_out >= D,
              "gather_add_rows: need 1..4 sources, D % 4 == 0");
  if (M == 0) return GNC_OK;
  GatherAddArgs a;
  for (int s = 0; s < 4; ++s) {
    a.tab[s] = s < nsrc ? tables[s] : nullptr;
    a.idx[s] = s < nsrc ? idx[s] : nullptr;
    a.ld[s] = s < nsrc ? ld[s] : 0;
    if (s < nsrc) GNC_REQUIRE(a.tab[s] && aligned16(a.tab[s]) && a.ld[s] % 4 == 0, "gather_add_rows: tables must be 16-byte aligned rows");
  }
  GNC_REQUIRE(aligned16(out) && ld_out % 4 == 0 && (!bias || aligned16(bias)), "gather_add_rows: out / bias alignment");
  a.nsrc = nsrc; a.bias = bias; a.relu = relu; a.out = out; a.ld_out = ld_out; a.M = M; a.D4 = D / 4;
  gather_add_rows_kernel<<<row_grid(M, 32), 256, 0, (cudaStream_t)stream>>>(a);
  return check_launch("gather_add_rows_kernel");
}

int gnc_segment_readout_f32(const float* y, const int32_t* node_ptr, int B, int num_nodes, float* out, gnc_stream_t stream) {
  GNC_REQUIRE(y && node_ptr && out && B >= 0 && num_nodes >= 1, "segment_readout: bad arguments");
  if (B == 0) return GNC_OK;
  const long long total = (long long)B * num_nodes;
  segment_readout_kernel<<<(unsigned)ceil_div<long long>(total, 256), 256, 0, (cudaStream_t)stream>>>(y, node_ptr, B, num_nodes, out);
  return check_launch("segment_readout_kernel");
}

int gnc_segment_readout_bwd_f32(const float* dout, const int32_t* node_ptr, int B, int num_nodes, float* dy, gnc_stream_t stream) {
  GNC_REQUIRE(dout && node_ptr && dy && B >= 0 && num_nodes >= 1, "segment_readout_bwd: bad arguments");
  if (B == 0) return GNC_OK;
  GNC_REQUIRE(B <= 65535, "segment_readout_bwd: at most 65535 graphs per call");
  segment_readout_bwd_kernel<<<dim3(4, (unsigned)B), 256, 0, (cudaStream_t)stream>>>(dout, node_ptr, B, num_nodes, dy);
  return check_launch("segment_readout_bwd_kernel");
}

int gnc_edge_geometry_f32(const float* pos, int P, const int32_t* src, const int32_t* dst, int64_t E, float* out,
                          gnc_stream_t stream) {
  GNC_REQUIRE(P >= 1 && P <= 8 && E >= 0, "edge_geometry: need 1 <= P <= 8");
  if (E == 0) return GNC_OK;
  GNC_REQUIRE(pos && src && dst && out, "edge_geometry: null pointer");
  int64_t blocks = ceil_div<int64_t>(E, 256);
  if (blocks > (int64_t)kNumSMs * 8) blocks = (int64_t)kNumSMs * 8;
  edge_geometry_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(pos, P, src, dst, E, out);
  return check_launch("edge_geometry_kernel");
}

}  // extern "C"

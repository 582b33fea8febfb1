// First / last layers of the GraphNet MLPs whose contraction is too thin for a GEMM tile:
//   * K <= 8 inputs  (node encoder: 3 pixel channels; edge encoder: 1 + space_dim geometry
//     values - reference models/GNN.py:233-234, models/MLP.py:24): HBM-bound streaming kernels,
//     one pass, 128-bit stores.  y = act(x W^T + b);  backward fuses ReLU mask, bias gradient
//     and weight gradient in one pass over (dY, Y, x) - there is no data gradient (x is input).
#include "common.cuh"

namespace gnc {

constexpr int kMaxK = 8;

// y[m, 4l..4l+3] for lane l (N <= 128 * VPL covered by VPL passes); x row broadcast-loaded
template <int K>
__global__ void __launch_bounds__(256) narrowk_fwd_kernel(const float* __restrict__ X, long long ldx, long long M,
                                                          const float* __restrict__ W, long long ldw,
                                                          const float* __restrict__ bias, int N4, int relu,
                                                          float* __restrict__ Y, long long ldy) {
  const int lane = threadIdx.x & 31;
  const long long warp_global = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long warps_total = ((long long)gridDim.x * blockDim.x) >> 5;
  for (int c4 = lane; c4 < N4; c4 += 32) {
    float w[4][K], b[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      b[j] = bias ? __ldg(bias + 4 * c4 + j) : 0.f;
#pragma unroll
      for (int k = 0; k < K; ++k) w[j][k] = __ldg(W + (long long)(4 * c4 + j) * ldw + k);
    }
    for (long long m = warp_global; m < M; m += warps_total) {
      float x[K];
#pragma unroll
      for (int k = 0; k < K; ++k) x[k] = __ldg(X + m * ldx + k);
      float o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < K; ++k) acc = fmaf(x[k], w[j][k], acc);
        acc += b[j];
        o[j] = relu ? fmaxf(acc, 0.f) : acc;
      }
      stg_stream(reinterpret_cast<float4*>(Y + m * ldy) + c4, make_float4(o[0], o[1], o[2], o[3]));
    }
  }
}

// partial[b][k][n] (k < K: weight gradient column k; k == K: bias gradient) over the block's rows,
// dZ = dY * (Y > 0) when Y != NULL.  N % 4 == 0, N <= 128.
template <int K>
__global__ void __launch_bounds__(256) narrowk_wgrad_kernel(const float* __restrict__ dY, long long lddy,
                                                            const float* __restrict__ Y, long long ldy,
                                                            const float* __restrict__ X, long long ldx, long long M,
                                                            int N4, float* __restrict__ partial) {
  __shared__ float4 red[8][K + 1][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long rows_per_block = ceil_div<long long>(M, gridDim.x);
  const long long rbeg = (long long)blockIdx.x * rows_per_block;
  long long rend = rbeg + rows_per_block;
  if (rend > M) rend = M;
  float4 s[K + 1];
#pragma unroll
  for (int k = 0; k <= K; ++k) s[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (lane < N4) {
    for (long long r = rbeg + warp; r < rend; r += 8) {
      float4 d = ldg_stream(reinterpret_cast<const float4*>(dY + r * lddy) + lane);
      if (Y) {
        const float4 y = ldg_stream(reinterpret_cast<const float4*>(Y + r * ldy) + lane);
        d.x = y.x > 0.f ? d.x : 0.f; d.y = y.y > 0.f ? d.y : 0.f;
        d.z = y.z > 0.f ? d.z : 0.f; d.w = y.w > 0.f ? d.w : 0.f;
      }
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const float x = __ldg(X + r * ldx + k);
        s[k].x = fmaf(d.x, x, s[k].x); s[k].y = fmaf(d.y, x, s[k].y);
        s[k].z = fmaf(d.z, x, s[k].z); s[k].w = fmaf(d.w, x, s[k].w);
      }
      s[K].x += d.x; s[K].y += d.y; s[K].z += d.z; s[K].w += d.w;
    }
  }
#pragma unroll
  for (int k = 0; k <= K; ++k) red[warp][k][lane] = s[k];
  __syncthreads();
  for (int i = threadIdx.x; i < (K + 1) * N4; i += blockDim.x) {
    const int k = i / N4, c4 = i - k * N4;
    float4 t = red[0][k][c4];
#pragma unroll
    for (int w = 1; w < 8; ++w) { const float4 u = red[w][k][c4]; t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w; }
    reinterpret_cast<float4*>(partial + ((long long)blockIdx.x * (K + 1) + k) * (N4 * 4))[c4] = t;
  }
}

// dW[n, k] (+)= sum_b partial[b][k][n];  db[n] (+)= sum_b partial[b][K][n]
__global__ void narrowk_reduce_kernel(const float* __restrict__ partial, long long blocks, int K, int N,
                                      float* __restrict__ dW, long long lddw, float* __restrict__ db, int accumulate) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (K + 1) * N) return;
  const int k = i / N, n = i - k * N;
  float s = 0.f;
  for (long long b = 0; b < blocks; ++b) s += partial[(b * (K + 1) + k) * N + n];
  if (k < K) {
    if (dW) { float* d = dW + (long long)n * lddw + k; *d = accumulate ? *d + s : s; }
  } else if (db) {
    db[n] = accumulate ? db[n] + s : s;
  }
}

static long long narrow_blocks(long long M) {
  long long b = ceil_div<long long>(M, 64);
  if (b > (long long)kNumSMs * 4) b = (long long)kNumSMs * 4;
  return b < 1 ? 1 : b;
}

}  // namespace gnc

using namespace gnc;

extern "C" {

int gnc_linear_narrowk_fwd_f32(const float* X, int64_t ldx, int64_t M, int K, const float* W, int64_t ldw,
                               const float* bias, int N, int relu, float* Y, int64_t ldy, gnc_stream_t stream) {
  GNC_REQUIRE(K >= 1 && K <= kMaxK && N > 0 && N % 4 == 0 && M >= 0 && ldx >= K && ldw >= K && ldy >= N,
              "linear_narrowk_fwd: need 1 <= K <= 8, N % 4 == 0");
  if (M == 0) return GNC_OK;                          // an edge-less graph (1 x 1 image): empty tensors have no pointers
  GNC_REQUIRE(X && W && Y, "linear_narrowk_fwd: null pointer");
  GNC_REQUIRE(ldy % 4 == 0 && aligned16(Y), "linear_narrowk_fwd: Y rows must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  long long blocks = ceil_div<long long>(M, 8);
  if (blocks > (long long)kNumSMs * 8) blocks = (long long)kNumSMs * 8;
  const int N4 = N / 4;
#define GNC_NK(KK) case KK: narrowk_fwd_kernel<KK><<<(unsigned)blocks, 256, 0, st>>>(X, ldx, M, W, ldw, bias, N4, relu, Y, ldy); break;
  switch (K) { GNC_NK(1) GNC_NK(2) GNC_NK(3) GNC_NK(4) GNC_NK(5) GNC_NK(6) GNC_NK(7) GNC_NK(8) }
#undef GNC_NK
  return check_launch("narrowk_fwd_kernel");
}

int64_t gnc_linear_narrowk_wgrad_workspace(int64_t M, int N, int K) { return narrow_blocks(M) * (int64_t)(K + 1) * N; }

int gnc_linear_narrowk_wgrad_f32(const float* dY, int64_t lddy, const float* Y, int64_t ldy, const float* X, int64_t ldx,
                                 int64_t M, int N, int K, float* dW, int64_t lddw, float* db, int accumulate,
                                 float* work, int64_t work_elems, gnc_stream_t stream) {
  GNC_REQUIRE(K >= 1 && K <= kMaxK && N > 0 && N % 4 == 0 && N <= 128 && M >= 0 && (M == 0 || (dY && X)) && lddy >= N && ldx >= K,
              "linear_narrowk_wgrad: need 1 <= K <= 8, N % 4 == 0, N <= 128");
  GNC_REQUIRE(lddy % 4 == 0 && (M == 0 || aligned16(dY)) && (!Y || (ldy % 4 == 0 && aligned16(Y))) && (!dW || lddw >= K),
              "linear_narrowk_wgrad: dY / Y rows must be 16-byte aligned");
  const long long blocks = M > 0 ? narrow_blocks(M) : 0;
  if (!work || !aligned16(work) || work_elems < blocks * (long long)(K + 1) * N)
    return fail(GNC_EWORKSPACE, "%s", "linear_narrowk_wgrad: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  if (M > 0) {
    const int N4 = N / 4;
#define GNC_NK(KK) case KK: narrowk_wgrad_kernel<KK><<<(unsigned)blocks, 256, 0, st>>>(dY, lddy, Y, ldy, X, ldx, M, N4, work); break;
    switch (K) { GNC_NK(1) GNC_NK(2) GNC_NK(3) GNC_NK(4) GNC_NK(5) GNC_NK(6) GNC_NK(7) GNC_NK(8) }
#undef GNC_NK
    if ((rc = check_launch("narrowk_wgrad_kernel"))) return rc;
  }
  narrowk_reduce_kernel<<<(unsigned)ceil_div<int>((K + 1) * N, 256), 256, 0, st>>>(work, blocks, K, N, dW, lddw, db, accumulate);
  return check_launch("narrowk_reduce_kernel");
}

}  // extern "C"

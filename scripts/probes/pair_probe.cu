// Hardware probe (not part of libgnc): validates the mechanics the chained-MLP kernel relies on
//   * a CTA pair (cluster of 2) sharing one tcgen05.mma.cta_group::2 (M = 256, N = 128),
//     B split by N across the two CTAs' shared memory, A read from each CTA's own TMEM;
//   * kind::f16 with bf16 operands split three ways (a = a1 + a2 + a3, 6 products) for fp32 parity;
//   * remote mbarrier arrives (peer -> leader) and the multicast tcgen05.commit.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o pair_probe pair_probe.cu
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <vector>

constexpr int kD = 128;
constexpr int kImgBytes = 64 * 128;          // one (piece, K-block) image: 64 weight rows x 128 bytes
constexpr int kOffBar = 6 * kImgBytes;       // 3 pieces x 2 K-blocks
constexpr int kSmem = kOffBar + 64 + 1024;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t local_bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}" ::"r"(local_bar), "r"(cta) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "W1:\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra.uni D1;\n\tbra.uni W1;\n\tD1:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;   // SWIZZLE_128B
  return d;
}
// kind::f16, bf16 x bf16 -> fp32, K-major A and B, M = 256 (pair), N = 128
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((256u >> 4) << 24);

constexpr uint32_t kIdescF16 = (1u << 4) | (0u << 7) | (0u << 10) | ((128u >> 3) << 17) | ((256u >> 4) << 24);
__device__ __forceinline__ void umma_bf16_ts2(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t acc, uint32_t idesc = kIdesc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc),
      "r"(idesc), "r"(acc) : "memory");
}
// (x0, x1) * scale -> two packed fp16 pairs, p1 + p2 == x * scale to 22 bits
__device__ __forceinline__ void split2h(float x0, float x1, float scale, uint32_t& p1, uint32_t& p2) {
  x0 *= scale; x1 *= scale;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(p1) : "f"(x1), "f"(x0));
  float h0, h1;
  asm("{.reg .f16 lo, hi; mov.b32 {lo, hi}, %2; cvt.f32.f16 %0, lo; cvt.f32.f16 %1, hi;}" : "=f"(h0), "=f"(h1) : "r"(p1));
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(p2) : "f"(x1 - h1), "f"(x0 - h0));
}
__device__ __forceinline__ void umma_commit2(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* u) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(u[0]), "r"(u[1]), "r"(u[2]), "r"(u[3]), "r"(u[4]), "r"(u[5]), "r"(u[6]), "r"(u[7]), "r"(u[8]), "r"(u[9]),
      "r"(u[10]), "r"(u[11]), "r"(u[12]), "r"(u[13]), "r"(u[14]), "r"(u[15]) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* r) {
  uint32_t* u = reinterpret_cast<uint32_t*>(r);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
        "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]),
        "=r"(u[16]), "=r"(u[17]), "=r"(u[18]), "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]), "=r"(u[23]),
        "=r"(u[24]), "=r"(u[25]), "=r"(u[26]), "=r"(u[27]), "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31])
      : "r"(taddr));
}

// (x0, x1) -> three packed bf16 pairs with p1 + p2 + p3 == x exactly (low half = x0)
__device__ __forceinline__ void split3(float x0, float x1, uint32_t& p1, uint32_t& p2, uint32_t& p3) {
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p1) : "f"(x1), "f"(x0));
  float r0 = x0 - __uint_as_float(p1 << 16), r1 = x1 - __uint_as_float(p1 & 0xffff0000u);
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p2) : "f"(r1), "f"(r0));
  r0 -= __uint_as_float(p2 << 16); r1 -= __uint_as_float(p2 & 0xffff0000u);
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p3) : "f"(r1), "f"(r0));
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
pair_probe_kernel(const float* __restrict__ A, const float* __restrict__ W, float* __restrict__ Y, int pieces) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  const uint32_t rank = cluster_rank();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t a_full = base + kOffBar, d_full = base + kOffBar + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + kOffBar + 16);

  if (threadIdx.x == 0) {
    mbar_init(a_full, 256);     // every thread of both CTAs
    mbar_init(d_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  // B images: this CTA holds output features n = 64*rank .. 64*rank+63
  for (int item = threadIdx.x; item < 64 * 16; item += 128) {
    const int c = item & 15, nl = item >> 4;           // 16-byte chunk of 8 k-values, local weight row
    const int n = (int)rank * 64 + nl;
    const int kb = c >> 3, cc = c & 7;
    uint32_t p1[4], p2[4], p3[4];
    for (int j = 0; j < 4; ++j) {
      const float x0 = W[n * kD + c * 8 + 2 * j], x1 = W[n * kD + c * 8 + 2 * j + 1];
      if (pieces == 12) { split2h(x0, x1, 256.f, p1[j], p2[j]); p3[j] = 0; }
      else split3(x0, x1, p1[j], p2[j], p3[j]);
    }
    const uint32_t off = (uint32_t)kb * kImgBytes + (uint32_t)((nl >> 3) * 1024 + (nl & 7) * 128 + ((cc ^ (nl & 7)) << 4));
    *reinterpret_cast<uint4*>(sm + 0 * 2 * kImgBytes + off) = make_uint4(p1[0], p1[1], p1[2], p1[3]);
    *reinterpret_cast<uint4*>(sm + 1 * 2 * kImgBytes + off) = make_uint4(p2[0], p2[1], p2[2], p2[3]);
    *reinterpret_cast<uint4*>(sm + 2 * 2 * kImgBytes + off) = make_uint4(p3[0], p3[1], p3[2], p3[3]);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t kTmemA = 256;                     // A pieces: [256,320) [320,384) [384,448)

  // A operand: thread = row, packed bf16 pairs, 64 columns per piece
  {
    const long long row = (long long)blockIdx.x * 128 + threadIdx.x;   // blockIdx.x = 128-row slice (rank within the pair tile)
    const float* a = A + row * kD;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
    for (int g = 0; g < 4; ++g) {                      // 32 k-values -> 16 packed columns per piece
      uint32_t p1[16], p2[16], p3[16];
      for (int j = 0; j < 16; ++j) {
        if (pieces == 12) { split2h(a[g * 32 + 2 * j], a[g * 32 + 2 * j + 1], 16.f, p1[j], p2[j]); p3[j] = 0; }
        else split3(a[g * 32 + 2 * j], a[g * 32 + 2 * j + 1], p1[j], p2[j], p3[j]);
      }
      tmem_st16(lane_addr + kTmemA + 0 * 64 + g * 16, p1);
      tmem_st16(lane_addr + kTmemA + 1 * 64 + g * 16, p2);
      tmem_st16(lane_addr + kTmemA + 2 * 64 + g * 16, p3);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    tc_fence_before();
    mbar_arrive_cluster(a_full, 0);                    // leader's barrier
  }
  if (rank == 0 && threadIdx.x == 0) {
    mbar_wait_cluster(a_full, 0);
    tc_fence_after();
    const uint64_t db = make_desc(base);
    bool first = true;
    for (int ks = 0; ks < 8; ++ks) {                   // K = 16 per instruction
      const int kb = ks >> 2, k = ks & 3;
      auto bdesc = [&](int piece) { return db + (uint64_t)((piece * 2 * kImgBytes + kb * kImgBytes + k * 32) >> 4); };
      auto apiece = [&](int piece) { return tmem_base + kTmemA + (uint32_t)piece * 64 + (uint32_t)ks * 8; };
      // (a_i, w_j) products, smallest first
      const int combos[6][2] = {{0, 2}, {2, 0}, {1, 1}, {0, 1}, {1, 0}, {0, 0}};
      for (int q = 0; q < 6; ++q) {
        if (pieces == 12) { if (q < 3) continue; }          // fp16: a1w2, a2w1, a1w1 only
        else if (combos[q][0] >= pieces || combos[q][1] >= pieces) continue;
        umma_bf16_ts2(tmem_base, apiece(combos[q][0]), bdesc(combos[q][1]), first ? 0u : 1u, pieces == 12 ? kIdescF16 : kIdesc);
        first = false;
      }
    }
    umma_commit2(d_full);
  }
  mbar_wait_cluster(d_full, 0);
  tc_fence_after();
  {
    const long long row = (long long)blockIdx.x * 128 + threadIdx.x;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
    for (int ch = 0; ch < 4; ++ch) {
      float r[32];
      tmem_ld32(lane_addr + ch * 32, r);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 32; ++j) Y[row * kD + ch * 32 + j] = pieces == 12 ? r[j] * (1.0f / 4096.f) : r[j];
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
  (void)lane;
}

int main() {
  const int M = 256 * 4;   // 4 pair tiles -> 8 CTAs
  std::vector<float> A(M * kD), W(kD * kD), Y(M * kD);
  srand(1);
  for (auto& v : A) v = (float)rand() / RAND_MAX * 4.f - 2.f;
  for (auto& v : W) v = ((float)rand() / RAND_MAX - 0.5f) * 0.25f;
  float *dA, *dW, *dY;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dW, W.size() * 4); cudaMalloc(&dY, Y.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(pair_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
  const int modes[4] = {1, 2, 3, 12};
  for (int mi = 0; mi < 4 * 3; ++mi) {
    const int pieces = modes[mi % 4];
    if (mi == 4) { for (auto& v : A) v *= 0.01f; cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice); printf("-- A scaled by 0.01\n"); }
    if (mi == 8) { for (auto& v : A) v *= 1e4f; cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice); printf("-- A scaled by 100 (|a| <= 200)\n"); }
    cudaMemset(dY, 0, Y.size() * 4);
    pair_probe_kernel<<<M / 128, 128, kSmem>>>(dA, dW, dY, pieces);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(Y.data(), dY, Y.size() * 4, cudaMemcpyDeviceToHost);
    double num = 0, den = 0, maxabs = 0;
    double err_q[4] = {0, 0, 0, 0};   // per quadrant: (row half of the pair tile) x (N half)
    for (int m = 0; m < M; ++m)
      for (int n = 0; n < kD; ++n) {
        double ref = 0;
        for (int k = 0; k < kD; ++k) ref += (double)A[m * kD + k] * (double)W[n * kD + k];
        const double d = Y[m * kD + n] - ref;
        num += d * d; den += ref * ref;
        if (fabs(d) > maxabs) maxabs = fabs(d);
        const int qd = ((m >> 7) & 1) * 2 + (n >> 6);
        if (fabs(d) > err_q[qd]) err_q[qd] = fabs(d);
      }
    printf("pieces=%d rel_l2=%.3e max_abs=%.3e quadrant_max_abs=[%.2e %.2e %.2e %.2e]\n", pieces, sqrt(num / den), maxabs,
           err_q[0], err_q[1], err_q[2], err_q[3]);
  }
  return 0;
}

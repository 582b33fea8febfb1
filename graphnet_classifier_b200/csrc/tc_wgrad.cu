// tcgen05 3xTF32 weight gradient for the width-128 layers:
//
//   dW[n, k] (+)= sum_m dZ[m, n] * X[m, k]        n, k in [0, 128),  m in [0, M)
//
// (backward of y = x W^T: models/MLP.py:24-27 under autograd).  The contraction runs over the
// ROWS, so both operands stream from HBM exactly once and the 128 x 128 result stays in one
// TMEM accumulator for the whole life of the CTA; a deterministic cross-CTA reduction follows.
//
// Per 32-row block ("stage") and per CTA (persistent, one per SM, 12 warps):
//   warps 0-3  Z path : cp.async ring (3 x 16 KB) of raw dZ rows -> each thread n reads COLUMN n
//                       of the block from smem (bank = n mod 32, conflict-free), splits hi/lo and
//                       writes the A operand dZ^T (lane = n, column = m) into TMEM with tcgen05.st
//                       (4 stages x {hi,lo} x 32 columns).  No transposition pass is needed: the
//                       row-major block already is "lane = column index" for this operand.
//   warps 4-7  X path : cp.async ring (3 x 16 KB) of raw X rows -> thread k reads COLUMN k of the
//                       block, splits hi/lo and writes row k of the UMMA K-major SWIZZLE_128B
//                       image (B operand X^T: N = k, K = m), 3 stages x 32 KB.  (kind::tf32 with an
//                       MN-major B descriptor returned zeros on B200, so the transposition is done
//                       by the column read, which is bank-conflict-free.)
//   warp  8    MMA    : one thread, 4 K-steps x 3 split products per stage, kind::tf32,
//                       A from TMEM, B from smem descriptors, D accumulates in TMEM.
//   warps 12-15 drain : the tensor core's fp32 accumulation rounds toward zero, so a long chain
//                       of accumulations drifts (measured 1e-5 after ~500 MMAs, growing linearly).
//                       The MMA therefore alternates between two TMEM accumulators every 4 stages
//                       (128 rows, 48 MMAs) and these warps add each finished partial into fp32
//                       REGISTERS with round-to-nearest adds (thread n holds row n of dW), then
//                       write this CTA's slice of the workspace at the end.
#include "common.cuh"

namespace gnc {
namespace tcw {

constexpr int kD = 128;
constexpr int kRows = 32;                       // rows per stage (K extent of one stage)
constexpr int kRawBytes = kRows * kD * 4;       // 16 KB
constexpr int kRing = 3;                        // cp.async ring depth per path
constexpr int kBStages = 3;
constexpr int kAStages = 4;
constexpr int kThreads = 16 * 32;
constexpr int kGroup = 4;                       // stages accumulated in the tensor core before a flush
constexpr int kRegsMma = 40, kRegsDrain = 200;  // launch: 128 per thread (65536 / 512)
constexpr int kTmemCols = 512;                  // [0,256): two accumulators; [256, 512): A stages (64 columns each)
constexpr int kTmemA = 256;

constexpr int kOffZraw = 0;                                   // 3 x 16 KB
constexpr int kOffXraw = kOffZraw + kRing * kRawBytes;        // 49152
constexpr int kOffB = kOffXraw + kRing * kRawBytes;           // 98304: [stage][hi|lo][16 KB]
constexpr int kOffBar = kOffB + kBStages * 2 * kRawBytes;     // 196608
constexpr int kSmemBytes = kOffBar + 256 + 1024;

struct Params {
  const float* dZ; long long lddz;
  const float* X; long long ldx;
  long long M;
  float* ws;                      // [grid][128][128] partial dW, then [grid][128] partial column sums of dZ
  int want_db;
  long long blocks_per_cta;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, 0x989680;\n\t"
      "@p bra.uni WAIT_DONE;\n\t"
      "bra.uni WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float* r) {
  const uint32_t* u = reinterpret_cast<const uint32_t*>(r);
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
__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

// K-major SWIZZLE_128B descriptor: rows of 128 bytes, 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// kind::tf32, fp32 accumulate, A (TMEM) and B K-major, M = 128, N = 128
constexpr uint32_t kInstrDesc = (1u << 4) | (2u << 7) | (2u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);

// copy one 32-row x 512-byte block global -> smem (row-major, unpadded); rows >= M are zero-filled
__device__ __forceinline__ void issue_block_copy(const float* src, long long ld, long long row0, long long M,
                                                 uint32_t dst, int t /*0..127*/) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int f = t + 128 * i;               // float4 index inside the block
    const int r = f >> 5, c4 = f & 31;
    const long long row = row0 + r;
    const long long rc = row < M ? row : M - 1;
    const uint32_t nbytes = row < M ? 16u : 0u;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + (uint32_t)(f * 16)),
                 "l"(src + rc * ld + c4 * 4), "r"(nbytes) : "memory");
  }
}

__global__ void __launch_bounds__(kThreads, 1) tc_wgrad_kernel(const Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = base + kOffBar;
  auto a_full = [&](int s) { return bar0 + 8u * s; };            // 4
  auto a_empty = [&](int s) { return bar0 + 32u + 8u * s; };     // 4
  auto b_full = [&](int s) { return bar0 + 64u + 8u * s; };      // 3
  auto b_empty = [&](int s) { return bar0 + 96u + 8u * s; };     // 3
  auto d_full = [&](int d) { return bar0 + 128u + 8u * d; };     // 2
  auto d_empty = [&](int d) { return bar0 + 144u + 8u * d; };    // 2
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + kOffBar + 176);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kAStages; ++s) { mbar_init(a_full(s), 128); mbar_init(a_empty(s), 1); }
    for (int s = 0; s < kBStages; ++s) { mbar_init(b_full(s), 128); mbar_init(b_empty(s), 1); }
    for (int d = 0; d < 2; ++d) { mbar_init(d_full(d), 1); mbar_init(d_empty(d), 128); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const long long nblocks_total = (p.M + kRows - 1) / kRows;
  const long long blk0 = (long long)blockIdx.x * p.blocks_per_cta;
  long long nblk = nblocks_total - blk0;
  if (nblk > p.blocks_per_cta) nblk = p.blocks_per_cta;
  if (nblk < 0) nblk = 0;

  if (warp < 4) {
    // ======================= Z path: dZ^T -> TMEM A operand =======================
    const int t = threadIdx.x;                 // 0..127 = column n = TMEM lane
    const int q = warp;
    const uint32_t zraw = base + kOffZraw;
    for (int b = 0; b < kRing; ++b) {
      if (b < nblk) issue_block_copy(p.dZ, p.lddz, (blk0 + b) * kRows, p.M, zraw + b * kRawBytes, t);
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    int rb = 0;
    float cs_sum = 0.f, cs_comp = 0.f;         // bias gradient: Kahan sum over blocks of per-block column sums
    for (long long it = 0; it < nblk; ++it) {
      asm volatile("cp.async.wait_group %0;" ::"n"(kRing - 1) : "memory");
      asm volatile("bar.sync 1, 128;" ::: "memory");          // every Z thread's copies of block `it` landed
      const float* tile = reinterpret_cast<const float*>(sm + kOffZraw + rb * kRawBytes);
      const int s = (int)(it & 3);
      const uint32_t ph = (uint32_t)((it >> 2) & 1);
      mbar_wait(a_empty(s), ph ^ 1u);
      tc_fence_after();
      const uint32_t ta = tmem_base + kTmemA + (uint32_t)s * 64 + ((uint32_t)(q * 32) << 16);
      float blk_sum = 0.f;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        float hi[16], lo[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float v = tile[(half * 16 + j) * kD + t];     // column t of the block: bank = t mod 32
          blk_sum += v;
          hi[j] = tf32_rna(v);
          lo[j] = tf32_rna(v - hi[j]);
        }
        tmem_st16(ta + half * 16, hi);
        tmem_st16(ta + 32 + half * 16, lo);
      }
      {
        const float y = blk_sum - cs_comp;
        const float tsum = cs_sum + y;
        cs_comp = (tsum - cs_sum) - y;
        cs_sum = tsum;
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
      mbar_arrive(a_full(s));
      asm volatile("bar.sync 1, 128;" ::: "memory");          // all Z threads finished reading the raw tile
      if (it + kRing < nblk) issue_block_copy(p.dZ, p.lddz, (blk0 + it + kRing) * kRows, p.M, zraw + rb * kRawBytes, t);
      asm volatile("cp.async.commit_group;" ::: "memory");
      rb = (rb + 1 == kRing) ? 0 : rb + 1;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (p.want_db) p.ws[(long long)gridDim.x * kD * kD + (long long)blockIdx.x * kD + t] = cs_sum;
  } else if (warp < 8) {
    // ======================= X path: X rows -> MN-major B operand images =======================
    const int t = threadIdx.x - 128;
    const uint32_t xraw = base + kOffXraw;
    for (int b = 0; b < kRing; ++b) {
      if (b < nblk) issue_block_copy(p.X, p.ldx, (blk0 + b) * kRows, p.M, xraw + b * kRawBytes, t);
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    int rb = 0;
    for (long long it = 0; it < nblk; ++it) {
      asm volatile("cp.async.wait_group %0;" ::"n"(kRing - 1) : "memory");
      asm volatile("bar.sync 2, 128;" ::: "memory");
      const float* tile = reinterpret_cast<const float*>(sm + kOffXraw + rb * kRawBytes);
      const int s = (int)(it % kBStages);
      const uint32_t ph = (uint32_t)((it / kBStages) & 1);
      mbar_wait(b_empty(s), ph ^ 1u);
      const uint32_t hi_base = base + kOffB + (uint32_t)s * 2 * kRawBytes;
      const uint32_t lo_base = hi_base + kRawBytes;
      {
        // thread t = output column k of dW = row k of the B image; its 32 values are column k of the block
        const uint32_t row_off = (uint32_t)((t >> 3) * 1024 + (t & 7) * 128);
#pragma unroll
        for (int c = 0; c < 8; ++c) {                          // 16-byte chunk c holds m = 4c .. 4c+3
          float4 v, hi, lo;
          v.x = tile[(4 * c + 0) * kD + t];
          v.y = tile[(4 * c + 1) * kD + t];
          v.z = tile[(4 * c + 2) * kD + t];
          v.w = tile[(4 * c + 3) * kD + t];
          hi.x = tf32_rna(v.x); hi.y = tf32_rna(v.y); hi.z = tf32_rna(v.z); hi.w = tf32_rna(v.w);
          lo.x = tf32_rna(v.x - hi.x); lo.y = tf32_rna(v.y - hi.y); lo.z = tf32_rna(v.z - hi.z); lo.w = tf32_rna(v.w - hi.w);
          const uint32_t off = row_off + (uint32_t)((c ^ (t & 7)) << 4);
          asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(hi_base + off), "f"(hi.x), "f"(hi.y), "f"(hi.z), "f"(hi.w) : "memory");
          asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(lo_base + off), "f"(lo.x), "f"(lo.y), "f"(lo.z), "f"(lo.w) : "memory");
        }
      }
      fence_proxy_async();
      mbar_arrive(b_full(s));
      asm volatile("bar.sync 2, 128;" ::: "memory");
      if (it + kRing < nblk) issue_block_copy(p.X, p.ldx, (blk0 + it + kRing) * kRows, p.M, xraw + rb * kRawBytes, t);
      asm volatile("cp.async.commit_group;" ::: "memory");
      rb = (rb + 1 == kRing) ? 0 : rb + 1;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  } else if (warp < 12) {
    // ======================= MMA issuer (warp 8; warps 9-11 pad the warpgroup) =======================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsMma));
    if (warp == 8 && lane == 0) {
#pragma unroll 1
      for (long long it = 0; it < nblk; ++it) {
        const long long grp = it / kGroup;
        const int in_grp = (int)(it % kGroup);
        const int d = (int)(grp & 1);
        if (in_grp == 0) {
          mbar_wait(d_empty(d), (uint32_t)((grp >> 1) & 1) ^ 1u);
          tc_fence_after();
        }
        const int sa = (int)(it & 3), sb = (int)(it % kBStages);
        mbar_wait(a_full(sa), (uint32_t)((it >> 2) & 1));
        mbar_wait(b_full(sb), (uint32_t)((it / kBStages) & 1));
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)d * kD;
        const uint32_t a_hi = tmem_base + kTmemA + (uint32_t)sa * 64;
        const uint32_t a_lo = a_hi + 32;
        const uint32_t b_hi = base + kOffB + (uint32_t)sb * 2 * kRawBytes;
        const uint32_t b_lo = b_hi + kRawBytes;
#pragma unroll
        for (int g = 0; g < 4; ++g) {                              // 8 rows of the block per MMA (K = 8)
          const uint64_t dbh = make_desc(b_hi + g * 32);           // 8 tf32 = 32 bytes along K in the swizzle row
          const uint64_t dbl = make_desc(b_lo + g * 32);
          umma_tf32_ts(d_tmem, a_lo + 8 * g, dbh, kInstrDesc, (in_grp | g) != 0);
          umma_tf32_ts(d_tmem, a_hi + 8 * g, dbl, kInstrDesc, 1);
          umma_tf32_ts(d_tmem, a_hi + 8 * g, dbh, kInstrDesc, 1);
        }
        umma_commit(a_empty(sa));
        umma_commit(b_empty(sb));
        if (in_grp == kGroup - 1 || it + 1 == nblk) umma_commit(d_full(d));
      }
    }
  } else {
    // ======================= drain: TMEM partials -> fp32 registers (RN) -> workspace =======================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsDrain));
    const int q = warp & 3;
    const int n = q * 32 + lane;                                   // row of dW held by this thread
    float acc[kD];
#pragma unroll
    for (int j = 0; j < kD; ++j) acc[j] = 0.f;
    const long long ngroups = (nblk + kGroup - 1) / kGroup;
#pragma unroll 1
    for (long long grp = 0; grp < ngroups; ++grp) {
      const int d = (int)(grp & 1);
      mbar_wait(d_full(d), (uint32_t)((grp >> 1) & 1));
      tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t)d * kD + ((uint32_t)(q * 32) << 16);
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        float r[32];
        tmem_ld32(taddr + ch * 32, r);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[ch * 32 + j] += r[j];
      }
      tc_fence_before();
      mbar_arrive(d_empty(d));
    }
    float* out = p.ws + ((long long)blockIdx.x * kD + n) * kD;
#pragma unroll
    for (int j = 0; j < kD / 4; ++j)
      *reinterpret_cast<float4*>(out + 4 * j) = make_float4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)kTmemCols) : "memory");
  }
}

__global__ void tc_wgrad_reduce_kernel(const float* __restrict__ ws, int parts, float* __restrict__ dW, long long lddw,
                                       int accumulate, float* __restrict__ db, int accumulate_db) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;      // 0 .. 128*128
  if (i >= kD * kD) return;
  if (db && i < kD) {
    const float* wdb = ws + (long long)parts * kD * kD;
    float sb = 0.f;
    for (int c = 0; c < parts; ++c) sb += wdb[c * kD + i];
    db[i] = accumulate_db ? db[i] + sb : sb;
  }
  float s = 0.f;
  for (int c = 0; c < parts; ++c) s += ws[(long long)c * kD * kD + i];
  const int n = i >> 7, k = i & 127;
  float* d = dW + n * lddw + k;
  *d = accumulate ? (*d + s) : s;
}

}  // namespace tcw
}  // namespace gnc

using namespace gnc;

extern "C" {

int64_t gnc_tc_wgrad_workspace(int64_t M) {
  (void)M;
  return (int64_t)kNumSMs * (tcw::kD * tcw::kD + tcw::kD);
}

int gnc_tc_wgrad_f32(const float* dZ, int64_t lddz, const float* X, int64_t ldx, int64_t M, int N, int K, float* dW,
                     int64_t lddw, int accumulate, float* db, float* work, int64_t work_elems, gnc_stream_t stream) {
  GNC_REQUIRE(N == tcw::kD && K == tcw::kD, "tc_wgrad: specialised for 128 x 128 weights");
  GNC_REQUIRE(dZ && X && dW && M >= 0 && lddz >= N && ldx >= K && lddw >= K, "tc_wgrad: bad arguments");
  GNC_REQUIRE(lddz % 4 == 0 && ldx % 4 == 0 && aligned16(dZ) && aligned16(X), "tc_wgrad: rows must be 16-byte aligned");
  if (!work || work_elems < gnc_tc_wgrad_workspace(M)) return fail(GNC_EWORKSPACE, "%s", "tc_wgrad: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  static SmemAttrOnce smem_attr;
  if (int rc_attr = smem_attr.ensure(tcw::tc_wgrad_kernel, tcw::kSmemBytes, "tc_wgrad")) return rc_attr;
  const long long nblocks = (M + tcw::kRows - 1) / tcw::kRows;
  long long grid = nblocks < kNumSMs ? nblocks : kNumSMs;
  if (grid < 1) grid = 1;
  tcw::Params p;
  p.dZ = dZ; p.lddz = lddz; p.X = X; p.ldx = ldx; p.M = M; p.ws = work;
  p.blocks_per_cta = (nblocks + grid - 1) / grid;
  p.want_db = db ? 1 : 0;
  tcw::tc_wgrad_kernel<<<(unsigned)grid, tcw::kThreads, tcw::kSmemBytes, st>>>(p);
  int rc = check_launch("tc_wgrad_kernel");
  if (rc) return rc;
  tcw::tc_wgrad_reduce_kernel<<<tcw::kD * tcw::kD / 256, 256, 0, st>>>(work, (int)grid, dW, lddw, accumulate, db, accumulate);
  return check_launch("tc_wgrad_reduce_kernel");
}

}  // extern "C"

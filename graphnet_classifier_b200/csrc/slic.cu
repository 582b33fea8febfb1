// SLIC superpixel segmentation (Achanta et al., "SLIC Superpixels Compared to State-of-the-art
// Superpixel Methods", TPAMI 2012) for batches of images - the stage the reference delegates to
// scikit-image (reference utils/image_to_graph/image_to_graph_superpixel.py:31:
// slic(img, n_segments=100, compactness=10, start_label=0)).
//
// PARITY UNPINNED against the reference: scikit-image is neither vendored nor version-pinned by the reference
// (requirements.txt:12) and is not installed here, so there are no reference labels to compare with.  What is pinned
// is the algorithm itself: oracle/slic.py restates it step by step on the CPU and tests/test_gpu_slic.py compares
// (>= 99.5 % identical labels, every difference a near-tie of the fp32 distance).  The steps, with scikit-image's
// documented conventions:
//   1. step = sqrt(H W / n_segments); ny = round(H / step) x nx = round(W / step) grid cells, K = ny nx centres;
//      spatial scale S = max(H / ny, W / nx);
//   2. sRGB -> CIELAB (D65), divided by `compactness`;
//   3. centre k = gy nx + gx starts at the middle of its cell with the colour of the pixel under it;
//   4. `iters` Lloyd iterations (10 in scikit-image): every pixel (at (y + 0.5, x + 0.5)) takes the nearest of the
//      centres of the 3 x 3 grid cells around its own cell (centres move by less than one step) under
//      d = |dLab|^2 + |dyx|^2 / S^2 - ranked by |c|^2 - 2 p.c, one fp32 FMA chain (slic_score), ties to the lowest
//      centre index; a centre moves
//      to the mean of its pixels; the sums are 2^-20 fixed-point integers, so the result does not depend on the order
//      the atomics arrive in (deterministic); a centre without pixels stays;
//   5. one more assignment gives the labels 0..K-1.
// Connectivity enforcement (scikit-image's post-pass) is the separate pass in csrc/slic_connect.cu, applied by
// utils/image_to_graph/slic.py by default as scikit-image does.
// Two device forms with identical labels: the streaming form (one launch per iteration over the whole batch) and, where
// the shape allows it, one CTA per image running the whole loop in one launch (slic_image_kernel below).
#include <stdlib.h>

#include "common.cuh"

namespace gnc {

struct SlicDims {
  int B, H, W, ny, nx, K, dbg;
  float step, inv_compactness;
};

__device__ __forceinline__ float srgb_to_linear(float c) {
  return c > 0.04045f ? powf((c + 0.055f) / 1.055f, 2.4f) : c / 12.92f;
}
__device__ __forceinline__ float lab_f(float t) {
  return t > 0.008856f ? cbrtf(t) : 7.787f * t + 16.0f / 116.0f;
}

// Nearest centre in the scaled (Lab / compactness, yx / S) space.  |p - c|^2 = |p|^2 - 2 p.c + |c|^2 and |p|^2 is common
// to all candidates of a pixel, so candidates are ranked by the SCORE |c|^2 - 2 p.c: five FMAs per pixel and candidate
// against a centre prepared once per iteration (-2 c and |c|^2).  Every kernel evaluates the same FMA chain in the
// same order, so all device forms round identically.
struct SlicCen { float m2L, m2A, m2B, m2x, m2y, cc; };
__device__ __forceinline__ SlicCen slic_prepare(float cL, float cA, float cB, float cy, float cx, float inv_step) {
  const float ys = cy * inv_step, xs = cx * inv_step;
  SlicCen c;
  c.m2L = -2.f * cL; c.m2A = -2.f * cA; c.m2B = -2.f * cB; c.m2x = -2.f * xs; c.m2y = -2.f * ys;
  c.cc = __fmaf_rn(xs, xs, __fmaf_rn(ys, ys, __fmaf_rn(cB, cB, __fmaf_rn(cA, cA, cL * cL))));
  return c;
}
// first link of the chain (shared by the pixels of a row), then the rest; ys / xs = (pixel centre) / S
__device__ __forceinline__ float slic_score_row(const SlicCen& c, float ys) { return __fmaf_rn(c.m2y, ys, c.cc); }
__device__ __forceinline__ float slic_score(const SlicCen& c, float row, float L, float A, float B, float xs) {
  return __fmaf_rn(c.m2B, B, __fmaf_rn(c.m2A, A, __fmaf_rn(c.m2L, L, __fmaf_rn(c.m2x, xs, row))));
}
__device__ __forceinline__ void rgb_to_lab(float r, float g, float b, float inv_c, float& L, float& A, float& Bc) {
  const float X = (0.412453f * r + 0.357580f * g + 0.180423f * b) / 0.95047f;
  const float Y = (0.212671f * r + 0.715160f * g + 0.072169f * b);
  const float Z = (0.019334f * r + 0.119193f * g + 0.950227f * b) / 1.08883f;
  const float fx = lab_f(X), fy = lab_f(Y), fz = lab_f(Z);
  L = (116.0f * fy - 16.0f) * inv_c;
  A = 500.0f * (fx - fy) * inv_c;
  Bc = 200.0f * (fy - fz) * inv_c;
}

// rgb uint8 -> (L, a, b) / compactness
__global__ void slic_lab_kernel(const uint8_t* __restrict__ img, long long npix, float inv_c, float* __restrict__ lab) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += stride) {
    const float r = srgb_to_linear(img[p * 3 + 0] * (1.0f / 255.0f));
    const float g = srgb_to_linear(img[p * 3 + 1] * (1.0f / 255.0f));
    const float b = srgb_to_linear(img[p * 3 + 2] * (1.0f / 255.0f));
    rgb_to_lab(r, g, b, inv_c, lab[p * 3 + 0], lab[p * 3 + 1], lab[p * 3 + 2]);
  }
}

// centres: [B, K, 5] = (L, a, b, y, x)
__global__ void slic_init_kernel(const float* __restrict__ lab, SlicDims d, float* __restrict__ centers) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= d.B * d.K) return;
  const int b = i / d.K, k = i - b * d.K;
  const int gy = k / d.nx, gx = k - gy * d.nx;
  const float cy = (gy + 0.5f) * d.H / d.ny, cx = (gx + 0.5f) * d.W / d.nx;
  int py = (int)cy, px = (int)cx;
  py = py < d.H ? py : d.H - 1; px = px < d.W ? px : d.W - 1;
  const float* l = lab + (((long long)b * d.H + py) * d.W + px) * 3;
  float* c = centers + (long long)i * 5;
  c[0] = l[0]; c[1] = l[1]; c[2] = l[2]; c[3] = cy; c[4] = cx;
}

constexpr double kFix = 1048576.0;   // 2^20 fixed-point scale for deterministic sums

// Assignment + accumulation.  Each thread owns kRun consecutive pixels of one image row:
//  * their Lab values arrive as 128-bit loads and their labels leave as 128-bit stores (a thread-per-pixel
//    layout with 12-byte pixels makes every scalar access a 32-sector request);
//  * the open runs of a warp's 32 lanes (32 * kRun consecutive pixels) are merged by a segmented warp
//    reduction before they go to the global accumulators;
//  * the fixed-point sums of the current label run stay in registers, so the 6 global 64-bit atomics are
//    issued once per run, not once per pixel (superpixels are ~25 pixels wide; integer sums: order-independent).
// kRun = 1 is the per-pixel form of the same arithmetic (debug switch, used by the tests as the reference).
template <int kRun>
__global__ void __launch_bounds__(256, 2) slic_assign_kernel(const float* __restrict__ lab, SlicDims d,
                                                          const float* __restrict__ centers, int32_t* __restrict__ labels,
                                                          long long* __restrict__ acc /*[B,K,6]*/) {
  const int gpr = (d.W + kRun - 1) / kRun;                     // pixel groups per row
  const long long ngroups = (long long)d.B * d.H * gpr;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const float inv_step = 1.0f / d.step;
  const bool vec = kRun % 4 == 0 && (d.W % kRun == 0) && ((reinterpret_cast<uintptr_t>(lab) | reinterpret_cast<uintptr_t>(labels)) & 15u) == 0;
  const int lane = threadIdx.x & 31;
  for (long long base = (long long)blockIdx.x * blockDim.x; base < ngroups; base += stride) {   // warp-uniform trip count
    const long long gi0 = base + threadIdx.x;
    const bool active = gi0 < ngroups;
    const long long gi = active ? gi0 : ngroups - 1;
    const long long row = gi / gpr;                            // b * H + y
    const int x0 = (int)(gi - row * gpr) * kRun;
    const int b = (int)(row / d.H);
    const int y = (int)(row - (long long)b * d.H);
    const int n = !active ? 0 : (d.W - x0 < kRun ? d.W - x0 : kRun);
    const long long p0 = row * d.W + x0;
    const int gy = (int)((long long)y * d.ny / d.H);
    const float* cb = centers + (long long)b * d.K * 5;
    float px[kRun * 3];
    int out[kRun];
    if (vec && active) {
      const float4* src = reinterpret_cast<const float4*>(lab + p0 * 3);
#pragma unroll
      for (int j = 0; j < (kRun * 3) / 4; ++j) {
        const float4 v = __ldg(src + j);
        px[4 * j] = v.x; px[4 * j + 1] = v.y; px[4 * j + 2] = v.z; px[4 * j + 3] = v.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < kRun * 3; ++j) px[j] = j < n * 3 ? lab[p0 * 3 + j] : 0.f;
    }
    int run_k = -1;                                            // key of the open run: b * K + k
    long long sL = 0, sA = 0, sB = 0, sx = 0, cnt = 0;
    auto flush_to = [&](int key, long long vL, long long vA, long long vB, long long vy, long long vx, long long vn) {
      long long* a = acc + (long long)key * 6;
      atomicAdd(reinterpret_cast<unsigned long long*>(a + 0), (unsigned long long)vL);
      atomicAdd(reinterpret_cast<unsigned long long*>(a + 1), (unsigned long long)vA);
      atomicAdd(reinterpret_cast<unsigned long long*>(a + 2), (unsigned long long)vB);
      atomicAdd(reinterpret_cast<unsigned long long*>(a + 3), (unsigned long long)vy);   // sum of 2*(y+0.5)
      atomicAdd(reinterpret_cast<unsigned long long*>(a + 4), (unsigned long long)vx);
      atomicAdd(reinterpret_cast<unsigned long long*>(a + 5), (unsigned long long)vn);
    };
    auto flush = [&]() {
      if (acc && cnt > 0) flush_to(run_k, sL, sA, sB, cnt * (2 * y + 1), sx, cnt);
    };
#pragma unroll
    for (int i = 0; i < kRun; ++i) {
      if (i < n) {
        const int x = x0 + i;
        const float L = px[3 * i], A = px[3 * i + 1], Bc = px[3 * i + 2];
        const int gx = (int)((long long)x * d.nx / d.W);
        float best = 3.4e38f;
        int best_k = gy * d.nx + gx;
#pragma unroll 1
        for (int dy = -1; dy <= 1; ++dy) {
          const int yy = gy + dy;
          if (yy < 0 || yy >= d.ny) continue;
#pragma unroll
          for (int dx = -1; dx <= 1; ++dx) {
            const int xx = gx + dx;
            if (xx < 0 || xx >= d.nx) continue;
            const int k = yy * d.nx + xx;
            const float* c = cb + (long long)k * 5;
            const SlicCen pc = slic_prepare(__ldg(c), __ldg(c + 1), __ldg(c + 2), __ldg(c + 3), __ldg(c + 4), inv_step);
            const float dist = slic_score(pc, slic_score_row(pc, (y + 0.5f) * inv_step), L, A, Bc, (x + 0.5f) * inv_step);
            if (dist < best) { best = dist; best_k = k; }     // ties: lowest centre index (scan order)
          }
        }
        out[i] = best_k;
        if (acc) {
          if (b * d.K + best_k != run_k) {
            flush();
            run_k = b * d.K + best_k; sL = sA = sB = sx = cnt = 0;
          }
          // v * 2^20 is exact in fp32, so this equals llrint((double)v * 2^20) without the (slow) fp64 pipe
          sL += __float2ll_rn(L * (float)kFix);
          sA += __float2ll_rn(A * (float)kFix);
          sB += __float2ll_rn(Bc * (float)kFix);
          sx += 2 * x + 1;                                     // 2*(x+0.5)
          cnt += 1;
        }
      }
    }
    // The open runs of neighbouring lanes usually carry the same label (a warp covers 32 * kRun consecutive
    // pixels): segmented warp reduction, one set of atomics per segment instead of per lane.
    if (acc) {
      const unsigned kFull = 0xffffffffu;
      const int key = (cnt > 0) ? run_k : -1;
      long long v[6] = {sL, sA, sB, cnt * (2 * y + 1), sx, cnt};
      const int prev = __shfl_up_sync(kFull, key, 1);
      const bool head = lane == 0 || prev != key;
      const unsigned heads = __ballot_sync(kFull, head);
      const unsigned higher = lane == 31 ? 0u : (heads & ~((2u << lane) - 1u));
      const int end = higher ? __ffs(higher) - 2 : 31;         // last lane of this lane's segment
#pragma unroll
      for (int delta = 1; delta < 32; delta <<= 1) {
#pragma unroll
        for (int j = 0; j < 6; ++j) {
          const long long t = __shfl_down_sync(kFull, v[j], delta);
          if (lane + delta <= end) v[j] += t;
        }
      }
      if (head && key >= 0) flush_to(key, v[0], v[1], v[2], v[3], v[4], v[5]);
    }
    if (vec && active) {
      int4* dst = reinterpret_cast<int4*>(labels + p0);
#pragma unroll
      for (int j = 0; j < kRun / 4; ++j) dst[j] = make_int4(out[4 * j], out[4 * j + 1], out[4 * j + 2], out[4 * j + 3]);
    } else {
#pragma unroll
      for (int i = 0; i < kRun; ++i)
        if (i < n) labels[p0 + i] = out[i];
    }
  }
}

__global__ void slic_update_kernel(SlicDims d, long long* __restrict__ acc, float* __restrict__ centers) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= d.B * d.K) return;
  long long* a = acc + (long long)i * 6;
  const long long n = a[5];
  if (n > 0) {
    float* c = centers + (long long)i * 5;
    const double inv = 1.0 / (double)n;
    c[0] = (float)((double)a[0] / kFix * inv);
    c[1] = (float)((double)a[1] / kFix * inv);
    c[2] = (float)((double)a[2] / kFix * inv);
    c[3] = (float)((double)a[3] * 0.5 * inv);
    c[4] = (float)((double)a[4] * 0.5 * inv);
  }
#pragma unroll
  for (int j = 0; j < 6; ++j) a[j] = 0;
}


// ---- one CTA per image: colour conversion, every Lloyd iteration and the final assignment in ONE launch --------------
// The streaming form above is one launch per iteration over the whole batch: every pixel re-reads its 9 candidate
// centres from global memory, writes a label map nobody reads until the last iteration and sends its sums to global
// 64-bit atomics (~2 ms per iteration for 1024 images of 256 x 256: 12 x the HBM time of its 16 bytes per pixel).
// An image's centres (K x 5 floats) and accumulators (K x 6 fixed-point sums) fit in shared memory, and images are
// independent, so a CTA can keep both there and run the whole loop for its image without leaving the SM:
//   * sRGB -> linear through a 256-entry table built with the streaming form's expression (same Lab bits);
//   * a thread owns runs of 8 pixels: the 3 candidate grid rows (and their scaled y distances) are found once per run,
//     the 3 candidate columns once per pixel, centres come from shared memory, the distance is the same FMA chain;
//   * sums leave the registers once per label run and land in shared-memory atomics (no warp merge: measured slower); the centre update is a __syncthreads() away instead of a launch away;
//   * labels are written once, after the last iteration.
// Same integer sums, same update arithmetic, same tie rule: the labels are those of the streaming form bit for bit
// (tests/test_gpu_graph_build.py compares them).
constexpr int kImgThreads = 256;
constexpr int kImgMaxK = 1024;

template <int kMinBlocks>
__global__ void __launch_bounds__(kImgThreads, kMinBlocks) slic_image_kernel(const uint8_t* __restrict__ img, SlicDims d, int iters,
                                                                   float* __restrict__ lab_ws, int32_t* __restrict__ labels) {
  extern __shared__ unsigned char slic_smem[];
  long long* acc = reinterpret_cast<long long*>(slic_smem);                 // [K][6]
  float* cen = reinterpret_cast<float*>(acc + (size_t)d.K * 6);            // [K][8]: prepared centres -2 (L, a, b, x / S, y / S), |c|^2 (32-byte records)
  float* lut = cen + (size_t)d.K * 8;                                      // [256]
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
  const int K = d.K, gpr = d.W >> 3, G = d.H * gpr;
  const long long pix0 = (long long)b * d.H * d.W;
  float* lab = lab_ws + pix0 * 3;
  const float inv_step = 1.0f / d.step;

  for (int i = tid; i < 256; i += kImgThreads) lut[i] = srgb_to_linear(i * (1.0f / 255.0f));
  for (int i = tid; i < K * 6; i += kImgThreads) acc[i] = 0;
  __syncthreads();
  // Lab of the image's pixels (kept in global memory: 12 bytes per pixel, re-read from L2 by every iteration)
  for (int g = tid; g < G; g += kImgThreads) {
    const long long p0 = (long long)g * 8;
    const uint2* src = reinterpret_cast<const uint2*>(img + (pix0 + p0) * 3);
    uint32_t w[6];
#pragma unroll
    for (int j = 0; j < 3; ++j) { const uint2 v = __ldg(src + j); w[2 * j] = v.x; w[2 * j + 1] = v.y; }
    float o[24];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int b0 = 3 * i, b1 = 3 * i + 1, b2 = 3 * i + 2;
      const float r = lut[(w[b0 >> 2] >> (8 * (b0 & 3))) & 255u], gg = lut[(w[b1 >> 2] >> (8 * (b1 & 3))) & 255u];
      const float bb = lut[(w[b2 >> 2] >> (8 * (b2 & 3))) & 255u];
      rgb_to_lab(r, gg, bb, d.inv_compactness, o[3 * i], o[3 * i + 1], o[3 * i + 2]);
    }
    float4* dst = reinterpret_cast<float4*>(lab + p0 * 3);
#pragma unroll
    for (int j = 0; j < 6; ++j) dst[j] = make_float4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
  }
  __syncthreads();
  for (int k = tid; k < K; k += kImgThreads) {                             // regular-grid initial centres
    const int gy = k / d.nx, gx = k - gy * d.nx;
    const float cy = (gy + 0.5f) * d.H / d.ny, cx = (gx + 0.5f) * d.W / d.nx;
    int py = (int)cy, px = (int)cx;
    py = py < d.H ? py : d.H - 1; px = px < d.W ? px : d.W - 1;
    const float* l = lab + ((long long)py * d.W + px) * 3;
    const SlicCen pc = slic_prepare(l[0], l[1], l[2], cy, cx, inv_step);
    cen[8 * k] = pc.m2L; cen[8 * k + 1] = pc.m2A; cen[8 * k + 2] = pc.m2B; cen[8 * k + 3] = pc.m2x;
    cen[8 * k + 4] = pc.m2y; cen[8 * k + 5] = pc.cc;
  }
  __syncthreads();

  // 64-bit adds are two native 32-bit shared-memory atomics with the carry passed on (integer sums: any order gives the
  // same total); position and count sums stay below 2^32 per image for every shape this kernel is launched for
  auto add64 = [&](int k, int j, long long v) {
    unsigned* a = reinterpret_cast<unsigned*>(acc + (size_t)k * 6 + j);
    const unsigned lo = (unsigned)v, hi = (unsigned)((unsigned long long)v >> 32);
    if (d.dbg & 2) return;
    const unsigned old = atomicAdd(a, lo);
    const unsigned h2 = hi + (((old + lo) < old) ? 1u : 0u);
    if (h2) atomicAdd(a + 1, h2);
  };
  auto add32 = [&](int k, int j, int v) {
    if (d.dbg & 2) return;
    atomicAdd(reinterpret_cast<unsigned*>(acc + (size_t)k * 6 + j), (unsigned)v);
  };
  // A thread owns an 8-pixel run in each of kRowsT consecutive rows and a warp an 8-pixel-wide strip of 32 * kRowsT
  // rows: its lanes share the candidate grid columns (uniform control flow in the candidate loop), a thread's label
  // usually survives from one row to the next (its sums stay in registers across the rows: |v| <= 108 / compactness
  // with compactness >= 4 keeps 32 pixels below 2^31 in 2^-20 fixed point), and the warp meets only a few labels.
  constexpr int kRowsT = 4;
  const int Gt = ((d.H + 32 * kRowsT - 1) / (32 * kRowsT)) * 32 * gpr;     // work items: (band, strip, lane)
  for (int it = 0; it <= iters; ++it) {
    const bool last = it == iters;
    for (int base = 0; base < Gt; base += kImgThreads) {                   // warp-uniform trip count
      const int gi = base + tid;
      const int band = gi / (gpr * 32), rem = gi - band * gpr * 32;
      const int x0 = (rem >> 5) << 3, yb = band * (32 * kRowsT) + (rem & 31) * kRowsT;
      // candidate grid columns of the run: its pixels lie in cell gx_lo or, past x_b, in gx_lo + 1 (the launcher
      // takes this kernel only for grid cells at least 8 pixels wide)
      // (32-bit: the launcher takes this kernel for K <= 1024 and H, W < 2^15 only, so the products stay below 2^31)
      const int gx_lo = (int)((unsigned)(x0 * d.nx) / (unsigned)d.W), gx_hi = (int)((unsigned)((x0 + 7) * d.nx) / (unsigned)d.W);
      const int x_b = (int)((unsigned)((gx_lo + 1) * d.W + d.nx - 1) / (unsigned)d.nx);   // first x of cell gx_lo + 1
      const int ib = x_b - x0;                                             // pixels i >= ib lie in cell gx_lo + 1
      const unsigned m_hi = ib >= 8 ? 0u : (0xffu << (ib < 0 ? 0 : ib)) & 0xffu;
      const float xf0 = x0 + 0.5f;
      float xs[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) xs[i] = (xf0 + (float)i) * inv_step;     // xf0 + i is exact: the value of (x + 0.5f)
#pragma unroll
      for (int i = 0; i < 8; ++i) asm volatile("" : "+f"(xs[i]));         // keep them in registers (the compiler re-derived them at every use)
      int run_k = -1, sL = 0, sA = 0, sB = 0, sy = 0, sx = 0, cnt = 0;     // the thread's open label run
#pragma unroll 1
      for (int r = 0; r < kRowsT; ++r) {
        const bool active = yb + r < d.H;
        const int y = active ? yb + r : d.H - 1;
        const int gy = (int)((unsigned)(y * d.ny) / (unsigned)d.H);
        float px[24];
        {
          const float4* src = reinterpret_cast<const float4*>(lab + ((long long)y * d.W + x0) * 3);
#pragma unroll
          for (int j = 0; j < 6; ++j) {
            const float4 v = src[j];
            px[4 * j] = v.x; px[4 * j + 1] = v.y; px[4 * j + 2] = v.z; px[4 * j + 3] = v.w;
          }
        }
        const float ys = (y + 0.5f) * inv_step;
        int out[8];
        float best[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { best[i] = 3.4e38f; out[i] = gy * d.nx + gx_lo + (int)((m_hi >> i) & 1u); }
        // candidates outermost (ascending centre index, as the per-pixel scan visits them): a prepared centre is read
        // from shared memory once per run instead of once per pixel
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy) {
          const int yy = gy + dy;
          if (yy < 0 || yy >= d.ny) continue;
#pragma unroll 1
          for (int xx = gx_lo - 1; xx <= gx_hi + 1; ++xx) {
            if (xx < 0 || xx >= d.nx) continue;
            const int k = yy * d.nx + xx;
            const float4 c4 = *reinterpret_cast<const float4*>(cen + 8 * k);      // -2 (L, a, b, x / S)
            const float2 c2 = *reinterpret_cast<const float2*>(cen + 8 * k + 4);  // -2 y / S, |c|^2
            SlicCen pc;
            pc.m2L = c4.x; pc.m2A = c4.y; pc.m2B = c4.z; pc.m2x = c4.w; pc.m2y = c2.x; pc.cc = c2.y;
            const float row = slic_score_row(pc, ys);
            // pixels this centre is a candidate of: the column left of the run's first cell serves the pixels of that
            // cell only, the column right of its second cell the pixels of the second cell only
            const unsigned cand = xx == gx_lo - 1 ? (~m_hi & 0xffu) : (xx == gx_lo + 2 ? m_hi : 0xffu);
            if (cand == 0xffu) {                 // the usual case (uniform over the warp: its lanes share x0): no mask
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float dist = slic_score(pc, row, px[3 * i], px[3 * i + 1], px[3 * i + 2], xs[i]);
                if (dist < best[i]) { best[i] = dist; out[i] = k; }
              }
            } else {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float dist = slic_score(pc, row, px[3 * i], px[3 * i + 1], px[3 * i + 2], xs[i]);
                if (((cand >> i) & 1u) && dist < best[i]) { best[i] = dist; out[i] = k; }
              }
            }
          }
        }
        if (last) {
          if (active) {
            int4* dst = reinterpret_cast<int4*>(labels + pix0 + (long long)y * d.W + x0);
            dst[0] = make_int4(out[0], out[1], out[2], out[3]);
            dst[1] = make_int4(out[4], out[5], out[6], out[7]);
          }
          continue;
        }
        if ((d.dbg & 1) || !active) continue;
        // a label run that closes inside the thread goes to the accumulators directly (boundaries only)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (out[i] != run_k) {
            if (cnt > 0) {
              add64(run_k, 0, sL); add64(run_k, 1, sA); add64(run_k, 2, sB);
              add32(run_k, 3, sy); add32(run_k, 4, sx); add32(run_k, 5, cnt);
            }
            run_k = out[i]; sL = sA = sB = sy = sx = cnt = 0;
          }
          sL += __float2int_rn(px[3 * i] * (float)kFix);
          sA += __float2int_rn(px[3 * i + 1] * (float)kFix);
          sB += __float2int_rn(px[3 * i + 2] * (float)kFix);
          sy += 2 * y + 1;
          sx += 2 * (x0 + i) + 1;
          cnt += 1;
        }
      }
      // the run still open goes to the accumulators like the others.  (Merging the lanes' open runs per label first -
      // match.any + nine redux.sync, one lane per label issuing the atomics - was measured: 7.32 against 6.72 ms per
      // 1024 images; so was a second register accumulator for the label on the other side of a boundary: 7.5 - 8.3 ms.)
      if (!last && !(d.dbg & 1) && cnt > 0) {
        add64(run_k, 0, sL); add64(run_k, 1, sA); add64(run_k, 2, sB);
        add32(run_k, 3, sy); add32(run_k, 4, sx); add32(run_k, 5, cnt);
      }
    }
    if (last) break;
    __syncthreads();
    for (int k = tid; k < K; k += kImgThreads) {                           // centre update (slic_update_kernel's arithmetic)
      long long* a = acc + (size_t)k * 6;
      const long long n = a[5];
      if (n > 0) {
        const double inv = 1.0 / (double)n;
        const SlicCen pc = slic_prepare((float)((double)a[0] / kFix * inv), (float)((double)a[1] / kFix * inv),
                                        (float)((double)a[2] / kFix * inv), (float)((double)a[3] * 0.5 * inv),
                                        (float)((double)a[4] * 0.5 * inv), inv_step);
        cen[8 * k] = pc.m2L; cen[8 * k + 1] = pc.m2A; cen[8 * k + 2] = pc.m2B; cen[8 * k + 3] = pc.m2x;
        cen[8 * k + 4] = pc.m2y; cen[8 * k + 5] = pc.cc;
      }
#pragma unroll
      for (int j = 0; j < 6; ++j) a[j] = 0;
    }
    __syncthreads();
  }
}

}  // namespace gnc

using namespace gnc;

static int g_slic_run = 8;     // pixels per thread in the assignment kernel (1 = the per-pixel form; debug)

static SlicDims slic_dims(int B, int H, int W, int n_segments, float compactness) {
  SlicDims d;
  d.B = B; d.H = H; d.W = W;
  const double step = sqrt((double)H * W / (double)(n_segments > 0 ? n_segments : 1));
  int ny = (int)floor(H / step + 0.5), nx = (int)floor(W / step + 0.5);
  d.ny = ny < 1 ? 1 : ny; d.nx = nx < 1 ? 1 : nx;
  d.K = d.ny * d.nx;
  const double sy = (double)H / d.ny, sx = (double)W / d.nx;
  d.step = (float)(sy > sx ? sy : sx);
  d.inv_compactness = 1.0f / compactness;
  d.dbg = getenv("GNC_SLIC_DBG") ? atoi(getenv("GNC_SLIC_DBG")) : 0;      // timing experiments (results invalid)
  return d;
}

extern "C" {

// Debug: 1 = per-pixel form of the streaming assignment kernel, -8 = streaming form with runs of 8 pixels (one launch per
// iteration), anything else = default (one CTA per image where the shape allows it, else the streaming form).  Same labels.
int gnc_debug_slic_run_length(int run) {
  g_slic_run = run == 1 ? 1 : (run == -8 ? -8 : 8);
  return GNC_OK;
}

int gnc_slic_num_centers(int H, int W, int n_segments) {
  if (H <= 0 || W <= 0) return 0;
  return slic_dims(1, H, W, n_segments, 10.f).K;
}

// work: bytes = gnc_slic_workspace_bytes(B, H, W, n_segments)
int64_t gnc_slic_workspace_bytes(int B, int H, int W, int n_segments) {
  if (B <= 0 || H <= 0 || W <= 0) return 0;
  const SlicDims d = slic_dims(B, H, W, n_segments, 10.f);
  return (int64_t)B * H * W * 3 * 4 + (int64_t)B * d.K * 5 * 4 + (int64_t)B * d.K * 6 * 8 + 256;
}

int gnc_slic_labels_u8(const uint8_t* img, int B, int H, int W, int n_segments, float compactness, int iters,
                       int32_t* labels, void* work, gnc_stream_t stream) {
  GNC_REQUIRE(B > 0 && H > 0 && W > 0 && n_segments > 0 && compactness > 0.f && iters >= 0, "slic: bad arguments");
  GNC_REQUIRE(img && labels && work, "slic: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const SlicDims d = slic_dims(B, H, W, n_segments, compactness);
  const long long npix = (long long)B * H * W;
  float* lab = reinterpret_cast<float*>(work);
  if (g_slic_run == 8 && d.K <= kImgMaxK && W % 8 == 0 && d.nx * 8 <= W && (double)H * W * (2.0 * (H > W ? H : W) + 1.0) < 4.0e9 && compactness >= 4.f && ((reinterpret_cast<uintptr_t>(img) | reinterpret_cast<uintptr_t>(labels) |
                                                         reinterpret_cast<uintptr_t>(lab)) & 15u) == 0) {
    // one CTA per image, the whole loop in one launch
    const int smem = d.K * (6 * 8 + 8 * 4) + 256 * 4;
    static SmemAttrOnce smem_attr2, smem_attr3, smem_attr4;
    static const int min_blocks = getenv("GNC_SLIC_MINB") ? atoi(getenv("GNC_SLIC_MINB")) : 3;   // 256 threads x 3 CTAs per SM (80 registers, 24 warps): 6.60 ms per 1024 images; 4 CTAs (64 registers, spills) 6.99, 2 CTAs 6.98
    if (min_blocks == 2) {
      if (int rc_attr = smem_attr2.ensure(slic_image_kernel<2>, 227 * 1024, "slic_image")) return rc_attr;
      slic_image_kernel<2><<<(unsigned)B, kImgThreads, smem, st>>>(img, d, iters, lab, labels);
    } else if (min_blocks == 4) {
      if (int rc_attr = smem_attr4.ensure(slic_image_kernel<4>, 227 * 1024, "slic_image")) return rc_attr;
      slic_image_kernel<4><<<(unsigned)B, kImgThreads, smem, st>>>(img, d, iters, lab, labels);
    } else {
      if (int rc_attr = smem_attr3.ensure(slic_image_kernel<3>, 227 * 1024, "slic_image")) return rc_attr;
      slic_image_kernel<3><<<(unsigned)B, kImgThreads, smem, st>>>(img, d, iters, lab, labels);
    }
    return check_launch("slic_image_kernel");
  }
  float* centers = lab + npix * 3;
  long long* acc = reinterpret_cast<long long*>(reinterpret_cast<uintptr_t>(centers + (long long)B * d.K * 5 + 1) & ~(uintptr_t)7);
  cudaError_t e = cudaMemsetAsync(acc, 0, (size_t)B * d.K * 6 * 8, st);
  if (e != cudaSuccess) return fail(GNC_ECUDA, "slic memset: %s", cudaGetErrorString(e));
  long long blocks = ceil_div<long long>(npix, 256);
  if (blocks > (long long)kNumSMs * 16) blocks = (long long)kNumSMs * 16;
  const int run = g_slic_run == 1 ? 1 : 8;
  long long ablocks = ceil_div<long long>((long long)B * H * ceil_div<int>(W, run), 256);   // one thread per pixel run
  if (ablocks > (long long)kNumSMs * 32) ablocks = (long long)kNumSMs * 32;
  int rc;
  slic_lab_kernel<<<(unsigned)blocks, 256, 0, st>>>(img, npix, d.inv_compactness, lab);
  if ((rc = check_launch("slic_lab_kernel"))) return rc;
  const unsigned kb = (unsigned)ceil_div<int>(B * d.K, 128);
  slic_init_kernel<<<kb, 128, 0, st>>>(lab, d, centers);
  if ((rc = check_launch("slic_init_kernel"))) return rc;
  for (int it = 0; it < iters; ++it) {
    if (run == 1) slic_assign_kernel<1><<<(unsigned)ablocks, 256, 0, st>>>(lab, d, centers, labels, acc);
    else slic_assign_kernel<8><<<(unsigned)ablocks, 256, 0, st>>>(lab, d, centers, labels, acc);
    if ((rc = check_launch("slic_assign_kernel"))) return rc;
    slic_update_kernel<<<kb, 128, 0, st>>>(d, acc, centers);
    if ((rc = check_launch("slic_update_kernel"))) return rc;
  }
  if (run == 1) slic_assign_kernel<1><<<(unsigned)ablocks, 256, 0, st>>>(lab, d, centers, labels, nullptr);   // final labels
  else slic_assign_kernel<8><<<(unsigned)ablocks, 256, 0, st>>>(lab, d, centers, labels, nullptr);
  return check_launch("slic_assign_kernel");
}

}  // extern "C"

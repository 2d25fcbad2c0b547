// microbenchmark: staging a row-segment footprint into shared memory with cp.async.bulk (1-D TMA, one instruction per row
// segment) versus cp.async 16 B (LDGSTS, one instruction per 4 floats).  Per "channel" a CTA copies ROWS segments of SEG floats.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
constexpr int THREADS = 256, ROWS = 54, CH = 23, D = 3;

__device__ __forceinline__ void mbar_init(unsigned a, int n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(n)); }
__device__ __forceinline__ void mbar_expect(unsigned a, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned a, unsigned parity) {
  asm volatile("{\n .reg .pred p;\n W: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE;\n bra W;\n DONE:\n}\n" ::"r"(a), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}

template <int MODE>  // 0 = bulk, 1 = LDGSTS
__global__ void __launch_bounds__(THREADS) stage(const float* __restrict__ src, float* out, int W, int H, int seg) {
  extern __shared__ float4 sm4[];
  float* sm = reinterpret_cast<float*>(sm4);
  __shared__ uint64_t bar[D];
  const unsigned sm_s = (unsigned)__cvta_generic_to_shared(sm), bar_s = (unsigned)__cvta_generic_to_shared(bar);
  const int stage_f = ROWS * seg;
  const int tid = threadIdx.x;
  // this CTA's footprint: rows y0..y0+ROWS/2 of two "directions", segment start x0 (16 B aligned)
  const int x0 = (blockIdx.x * 32) % (W - seg - 32), y0 = (blockIdx.y * 16) % (H - ROWS);
  const size_t plane = (size_t)W * H;
  const float* base = src + (size_t)blockIdx.z * CH * plane;
  if (MODE == 0 && tid == 0) { for (int s = 0; s < D; ++s) mbar_init(bar_s + 8 * s, 1); }
  __syncthreads();
  float acc = 0.f;
  auto issue = [&](int c, int s) {
    if (c >= CH) { if (MODE == 1) asm volatile("cp.async.commit_group;"); return; }
    const float* p = base + (size_t)c * plane;
    if (MODE == 0) {
      if (tid == 0) mbar_expect(bar_s + 8 * s, stage_f * 4);
      if (tid < ROWS) bulk_g2s(sm_s + 4 * (s * stage_f + tid * seg), p + (size_t)(y0 + tid) * W + x0 + ((tid * 4) & 12), seg * 4, bar_s + 8 * s);
    } else {
      const int n4 = stage_f / 4, s4 = seg / 4;
      for (int k = tid; k < n4; k += THREADS) {
        const int r = k / s4, e = k - r * s4;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sm_s + 4 * (s * stage_f) + 16 * k), "l"(p + (size_t)(y0 + r) * W + x0 + ((r * 4) & 12) + 4 * e));
      }
      asm volatile("cp.async.commit_group;");
    }
  };
  issue(0, 0); issue(1, 1);
  for (int c = 0; c < CH; ++c) {
    const int s = c % D;
    if (MODE == 0) mbar_wait(bar_s + 8 * s, (c / D) & 1); else asm volatile("cp.async.wait_group 1;");
    __syncthreads();
    issue(c + 2, (c + 2) % D);
    // consume: 8 LDS per thread
    const float* b = sm + s * stage_f;
#pragma unroll
    for (int k = 0; k < 8; ++k) acc += b[(tid * 5 + k * 131) % stage_f];
  }
  if (acc == 1.2345f) out[0] = acc;
}

int main() {
  const int W = 512, H = 256, N = 16;
  float* src; float* out;
  CK(cudaMalloc(&src, (size_t)N * CH * W * H * 4)); CK(cudaMalloc(&out, 1024));
  CK(cudaMemset(src, 0, (size_t)N * CH * W * H * 4));
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int seg : {16, 28, 40, 64}) {
    const int smem = D * ROWS * seg * 4;
    CK(cudaFuncSetAttribute(stage<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CK(cudaFuncSetAttribute(stage<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    dim3 grid(16, 16, N);
    for (int mode = 0; mode < 2; ++mode) {
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(a);
        for (int r = 0; r < 5; ++r) { if (mode == 0) stage<0><<<grid, THREADS, smem>>>(src, out, W, H, seg); else stage<1><<<grid, THREADS, smem>>>(src, out, W, H, seg); }
        cudaEventRecord(b); CK(cudaEventSynchronize(b));
      }
      float ms; cudaEventElapsedTime(&ms, a, b); ms /= 5;
      double bytes = (double)grid.x * grid.y * grid.z * CH * ROWS * seg * 4;
      printf("seg %2d floats x %d rows  %-6s : %.3f ms  %.1f GB/s staged  (%.0f copies/us chip-wide)\n", seg, ROWS, mode == 0 ? "bulk" : "ldgsts", ms, bytes / ms / 1e6,
             mode == 0 ? (double)grid.x * grid.y * grid.z * CH * ROWS / ms / 1e3 : 0.0);
    }
  }
  CK(cudaGetLastError());
  return 0;
}

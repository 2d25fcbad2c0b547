// mb_bulk2.cu — throughput of the async-proxy copy engine for ROW-SEGMENT sized transfers on B200:
//   mode 0: cp.async.bulk global -> shared (UBLKCP.S.G), completion on an mbarrier per stage
//   mode 1: cp.reduce.async.bulk shared -> global .add.f32 (bulk_group completion)
//   mode 2: cp.async 16 B (LDGSTS) of the same bytes, for reference
// Each CTA keeps S stages of R rows x SEG bytes in flight; IW warps issue (one row per lane).  No consumer work: this is
// the engine / issue rate, not a kernel.   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o mb_bulk2 tools/mb_bulk2.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
constexpr int THREADS = 128;
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
__device__ __forceinline__ void bulk_red(float* dst, unsigned src, unsigned bytes) {
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(THREADS) k(float* __restrict__ buf, int W, int H, int planes, int R, int seg_f, int S, int iters, int IW) {
  extern __shared__ float4 sm4[];
  __shared__ uint64_t bar[16];
  const unsigned sm_s = (unsigned)__cvta_generic_to_shared(sm4), bar_s = (unsigned)__cvta_generic_to_shared(bar);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int stage_f = R * seg_f;
  if (tid == 0) for (int s = 0; s < S; ++s) mbar_init(bar_s + 8 * s, 1);
  for (int i = tid; i < S * stage_f; i += THREADS) reinterpret_cast<float*>(sm4)[i] = 1.0f;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  const size_t plane = (size_t)W * H;
  // a pseudo-random tile position per (CTA, iteration): rows y0.., segment start x0 (16 B aligned)
  unsigned h = blockIdx.x * 2654435761u + 12345u;
  for (int it = 0; it < iters; ++it) {
    const int s = it % S;
    h = h * 1664525u + 1013904223u;
    const int x0 = ((h >> 8) % (W - seg_f - 4)) & ~3, y0 = (h >> 20) % (H - R), pl = (h >> 3) % planes;
    float* base = buf + pl * plane + (size_t)y0 * W + x0;
    if (MODE == 0) {
      if (it >= S) mbar_wait(bar_s + 8 * s, ((it / S) - 1) & 1);  // previous use of this stage has landed
      __syncwarp();
      if (tid == 0) mbar_expect(bar_s + 8 * s, stage_f * 4);
      __syncthreads();
      for (int r = tid; r < R; r += 32 * IW)
        if (warp < IW) bulk_g2s(sm_s + 4 * (s * stage_f + r * seg_f), base + (size_t)r * W, seg_f * 4, bar_s + 8 * s);
    } else if (MODE == 1) {
      if (warp < IW) {
        if (it >= S) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(3) : "memory");
        for (int r = tid; r < R; r += 32 * IW) bulk_red(base + (size_t)r * W, sm_s + 4 * (s * stage_f + r * seg_f), seg_f * 4);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    } else {
      const int s4 = seg_f / 4, n4 = R * s4;
      for (int q = tid; q < n4; q += THREADS) {
        const int r = q / s4, e = q - r * s4;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sm_s + 4 * (s * stage_f) + 16 * q), "l"(base + (size_t)r * W + 4 * e));
      }
      asm volatile("cp.async.commit_group;");
      asm volatile("cp.async.wait_group 3;");
    }
  }
  if (MODE == 0) for (int s = 0; s < S && s < iters; ++s) { const int last = ((iters - 1 - s) / S) * S + s; mbar_wait(bar_s + 8 * s, (last / S) & 1); }
  if (MODE == 1) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  if (MODE == 2) asm volatile("cp.async.wait_group 0;");
  __syncthreads();
  if (reinterpret_cast<float*>(sm4)[tid] == 1.2345f) buf[0] = 1.f;
}

int main() {
  const int W = 512, H = 256, planes = 16 * 46;
  float* buf; CK(cudaMalloc(&buf, (size_t)planes * W * H * 4)); CK(cudaMemset(buf, 0, (size_t)planes * W * H * 4));
  int nsm; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  const int iters = 400;
  for (int mode = 0; mode < 3; ++mode)
    for (int ctas : {2, 4})
      for (int R : {32, 64})
        for (int seg_f : {28, 48, 64})
          for (int IW : {1, 2}) {
            if (mode == 2 && IW == 2) continue;
            if (R == 32 && IW == 2) continue;
            const int S = 4, smem = S * R * seg_f * 4;
            auto fn = mode == 0 ? k<0> : mode == 1 ? k<1> : k<2>;
            CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            const int grid = nsm * ctas;
            fn<<<grid, THREADS, smem>>>(buf, W, H, planes, R, seg_f, S, iters, IW);
            CK(cudaDeviceSynchronize());
            cudaEventRecord(a);
            fn<<<grid, THREADS, smem>>>(buf, W, H, planes, R, seg_f, S, iters, IW);
            cudaEventRecord(b); CK(cudaEventSynchronize(b));
            float ms; cudaEventElapsedTime(&ms, a, b);
            const double copies = (double)grid * iters * R, bytes = copies * seg_f * 4;
            printf("%-6s CTAs/SM %d rows %2d seg %3d B issue-warps %d : %7.3f ms  %7.1f GB/s  %5.2f cycles/row/SM (1.965 GHz)\n",
                   mode == 0 ? "g2s" : mode == 1 ? "red" : "ldgsts", ctas, R, seg_f * 4, IW, ms, bytes / ms / 1e6, ms * 1e-3 * 1.965e9 / (copies / nsm));
          }
  CK(cudaGetLastError());
  return 0;
}

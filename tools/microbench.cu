// microbench.cu — B200 primitives that decide the scatter/gather kernel design (DESIGN.md "Measured primitives").
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o microbench tools/microbench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

constexpr int TPB = 256;
constexpr int ITERS = 256;
constexpr int SM_WORDS = 8192;  // 32 KB of 4-byte words

// mode 0: int32 smem atomicAdd; 1: int64 smem atomicAdd; 2: float LDS+FADD+STS (no atomic, racy but measures rate);
// 3: float atomicAdd (CAS loop); pattern: stride between lanes (1 = conflict-free consecutive), jitter: pseudo-random rows
template <int MODE>
__global__ void smem_rmw(float* out, int stride, int rowjit) {
  __shared__ unsigned long long sm64[SM_WORDS / 2];
  unsigned* sm = (unsigned*)sm64;
  float* smf = (float*)sm64;
  for (int i = threadIdx.x; i < SM_WORDS; i += TPB) sm[i] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned h = threadIdx.x * 2654435761u + blockIdx.x;
  float v = 1.0f + lane;
  for (int it = 0; it < ITERS; ++it) {
    h = h * 1664525u + 1013904223u;
    // address: lanes consecutive*stride within a row, row chosen per-lane when rowjit (rows of 97 words)
    int row = rowjit ? ((h >> 20) % rowjit) : 0;
    int a = (warp * 513 + it * 37 + row * 97 + lane * stride) % (MODE == 1 ? SM_WORDS / 2 : SM_WORDS);
    if (MODE == 0) atomicAdd(&sm[a], (unsigned)(int)v);
    if (MODE == 1) atomicAdd(&sm64[a], (unsigned long long)(long long)v);
    if (MODE == 2) { float t = smf[a]; smf[a] = t + v; }
    if (MODE == 3) atomicAdd(&smf[a], v);
  }
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = smf[7] + (float)sm[9];
}

// global RED: each warp adds 32 floats; pattern 0: consecutive (1 line), 1: 16 rows x 2 lanes (rough flow), v4 variant
__global__ void gred(float* buf, size_t n, int rows, int use_v4) {
  const int lane = threadIdx.x & 31;
  size_t gw = (size_t)(blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5));
  for (int it = 0; it < 32; ++it) {
    size_t base = ((gw * 32 + it) * 4099) % (n - 70000);
    if (use_v4) {
      size_t a = (base & ~3ull) + (size_t)(lane % (32 / rows)) * 4 + (size_t)(lane / (32 / rows)) * 2048;
      asm volatile("red.global.add.v4.f32 [%0], {%1,%1,%1,%1};" ::"l"(buf + a), "f"(1.0f) : "memory");
    } else {
      size_t a = base + (size_t)(lane % (32 / rows)) + (size_t)(lane / (32 / rows)) * 2048;
      atomicAdd(buf + a, 1.0f);
    }
  }
}

// gather: each warp loads 32 floats spread over `rows` rows (row pitch 2048 floats), consecutive within a row
__global__ void gather(const float* __restrict__ buf, float* out, size_t n, int rows) {
  const int lane = threadIdx.x & 31;
  size_t gw = (size_t)(blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5));
  float acc = 0;
  size_t base = (gw * 64) % (n - 200000);
  for (int it = 0; it < 64; ++it) {
    size_t a = base + (size_t)it * 2048 * 0 + (size_t)(it % 23) * 4096 * 23 % 1 + (size_t)(lane % (32 / rows)) + (size_t)(lane / (32 / rows)) * 2048 + it;
    acc += __ldg(buf + a);
  }
  if (acc == 123.456f) out[0] = acc;
}

int main() {
  float* out; CK(cudaMalloc(&out, 1 << 20));
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  int nsm; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("SMs %d clock %d kHz\n", nsm, clk);
  const int grid = nsm * 4;
  auto run = [&](auto kern, const char* name, int stride, int jit) {
    kern<<<grid, TPB>>>(out, stride, jit); cudaDeviceSynchronize();
    cudaEventRecord(a); for (int r = 0; r < 5; ++r) kern<<<grid, TPB>>>(out, stride, jit); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); ms /= 5;
    double warp_ops = (double)grid * (TPB / 32) * ITERS;  // warp-level RMW instructions
    double per_sm_cyc = ms * 1e-3 * 1.965e9 / (warp_ops / nsm);
    printf("%-28s stride %2d rowjit %2d : %8.3f ms  %6.2f cycles/warp-op/SM (at 1.965 GHz)\n", name, stride, jit, ms, per_sm_cyc);
  };
  for (int jit : {0, 16}) for (int stride : {1, 2, 32}) {
    run(smem_rmw<0>, "smem atomicAdd int32", stride, jit);
    run(smem_rmw<1>, "smem atomicAdd int64", stride, jit);
    run(smem_rmw<2>, "smem LDS+FADD+STS", stride, jit);
    run(smem_rmw<3>, "smem atomicAdd float(CAS)", stride, jit);
  }
  size_t n = (size_t)256 << 20; float* buf; CK(cudaMalloc(&buf, n * 4)); CK(cudaMemset(buf, 0, n * 4));
  for (int v4 : {0, 1}) for (int rows : {1, 4, 16}) {
    int g = nsm * 64;
    gred<<<g, 256>>>(buf, n, rows, v4); cudaDeviceSynchronize();
    cudaEventRecord(a); for (int r = 0; r < 3; ++r) gred<<<g, 256>>>(buf, n, rows, v4); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); ms /= 3;
    double lanes = (double)g * 8 * 32 * 32 * (v4 ? 4 : 1);
    printf("global RED %s rows %2d : %8.3f ms  %8.2f Gfloat-adds/s\n", v4 ? "v4.f32" : "f32   ", rows, ms, lanes / ms / 1e6);
  }
  for (int rows : {1, 2, 4, 8, 16, 32}) {
    int g = nsm * 128;
    gather<<<g, 256>>>(buf, out, n, rows); cudaDeviceSynchronize();
    cudaEventRecord(a); for (int r = 0; r < 3; ++r) gather<<<g, 256>>>(buf, out, n, rows); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); ms /= 3;
    double ldg = (double)g * 8 * 64;
    printf("gather LDG.32 rows %2d : %8.3f ms  %6.2f cycles/warp-LDG/SM\n", rows, ms, ms * 1e-3 * 1.965e9 / (ldg / nsm));
  }
  CK(cudaGetLastError());
  return 0;
}

// mb_tex.cu — can the TEXTURE path carry the 4-tap gather?  tex2Dgather returns the 2x2 quad of one fp32 plane in ONE
// instruction (exact values, no filtering).  This is the forward warp+blend of config 2 written on texture objects over
// pitch-linear planes ([C*H, W] per clip), flows = smooth pseudo-random field with |grad| ~ 0.6 (like bench.make_inputs).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o mb_tex tools/mb_tex.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s (line %d)\n", #x, cudaGetErrorString(e), __LINE__); return 1; } } while (0)

constexpr int N = 16, C = 23, H = 256, W = 512;

struct Tex { cudaTextureObject_t t[2][N]; };

// MODE 0: tex2Dgather; MODE 1: 4 x tex2D point fetch; MODE 2: 4 x __ldg (generic kernel's way); MODE 3: hybrid, channels
// below NLDG through __ldg (LSU pipe), the rest through tex2Dgather (TEX pipe); MODE 4: tex2Dgather + 4 shared-memory integer
// atomics per (direction, channel) (the backward's scatter), MODE 5: the atomics alone, MODE 6: tex2Dgather + 4 LDS
template <int MODE, int PW, int NLDG = 0>
__global__ void __launch_bounds__(256) fwd(const __grid_constant__ Tex T, const float* __restrict__ src0, const float* __restrict__ src1,
                                           const float* __restrict__ fx, const float* __restrict__ fy, float* __restrict__ out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int PH = 32 / PW;
  // CTA = 8 warps as 4 x 2 patches
  const int j = blockIdx.x * (4 * PW) + (warp & 3) * PW + (lane % PW);
  const int i = blockIdx.y * (2 * PH) + (warp >> 2) * PH + (lane / PW);
  const int n = blockIdx.z;
  const size_t pix = ((size_t)n * H + i) * W + j;
  float X[2], Y[2], w[2][4];
#pragma unroll
  for (int d = 0; d < 2; ++d) {
    const float sx = d ? 1.f : -1.f;
    float ix = (float)j + sx * fx[pix + (size_t)d * N * H * W], iy = (float)i + sx * fy[pix + (size_t)d * N * H * W];
    ix = fminf(fmaxf(ix, 0.f), (float)(W - 1));
    iy = fminf(fmaxf(iy, 0.f), (float)(H - 1));
    const float x0 = floorf(ix), y0 = floorf(iy), tx = ix - x0, ty = iy - y0;
    X[d] = x0, Y[d] = y0;
    w[d][0] = (1 - tx) * (1 - ty), w[d][1] = tx * (1 - ty), w[d][2] = (1 - tx) * ty, w[d][3] = tx * ty;
  }
  float* o = out + ((size_t)n * C * H + i) * W + j;
  __shared__ int acc[2][2048];
  unsigned sa[2];
  if (MODE >= 4) {
    for (int k = threadIdx.x; k < 4096; k += 256) (&acc[0][0])[k] = 0;
    __syncthreads();
#pragma unroll
    for (int d = 0; d < 2; ++d) sa[d] = (unsigned)__cvta_generic_to_shared(&acc[d][(((int)Y[d] & 31) * 44 + ((int)X[d] & 31)) & 2047]);
  }
#pragma unroll
  for (int c = 0; c < C; ++c) {
    float r = 0.f;
#pragma unroll
    for (int d = 0; d < 2; ++d) {
      float a, b, cc, dd;
      if (MODE == 5) {
        a = b = cc = dd = w[d][0];
      } else if (MODE == 0 || MODE == 4 || MODE == 6 || (MODE == 3 && c >= NLDG)) {
        const float4 q = tex2Dgather<float4>(T.t[d][n], X[d] + 1.0f, Y[d] + 1.0f + (float)(c * H), 0);
        a = q.w, b = q.z, cc = q.x, dd = q.y;  // (x0,y0) (x0+1,y0) (x0,y0+1) (x0+1,y0+1)
      } else if (MODE == 1) {
        const float yy = Y[d] + 0.5f + (float)(c * H), xx = X[d] + 0.5f;
        a = tex2D<float>(T.t[d][n], xx, yy), b = tex2D<float>(T.t[d][n], xx + 1.f, yy);
        cc = tex2D<float>(T.t[d][n], xx, yy + 1.f), dd = tex2D<float>(T.t[d][n], xx + 1.f, yy + 1.f);
      } else {
        const float* s = (d ? src1 : src0) + ((size_t)(n * C + c) * H + (int)Y[d]) * W + (int)X[d];
        const bool xi = (int)X[d] + 1 < W, yi = (int)Y[d] + 1 < H;
        a = __ldg(s), b = xi ? __ldg(s + 1) : 0.f, cc = yi ? __ldg(s + W) : 0.f, dd = (xi && yi) ? __ldg(s + W + 1) : 0.f;
      }
      r += a * w[d][0] + b * w[d][1] + cc * w[d][2] + dd * w[d][3];
      if (MODE == 4 || MODE == 5) {
        const unsigned ad = sa[d] + (c & 1) * 0;
        asm volatile("red.shared.add.s32 [%0], %1;" ::"r"(ad), "r"(__float_as_int(a * w[d][0] + 12582912.f)) : "memory");
        asm volatile("red.shared.add.s32 [%0+4], %1;" ::"r"(ad), "r"(__float_as_int(b * w[d][1] + 12582912.f)) : "memory");
        asm volatile("red.shared.add.s32 [%0+176], %1;" ::"r"(ad), "r"(__float_as_int(cc * w[d][2] + 12582912.f)) : "memory");
        asm volatile("red.shared.add.s32 [%0+180], %1;" ::"r"(ad), "r"(__float_as_int(dd * w[d][3] + 12582912.f)) : "memory");
      }
      if (MODE == 6) {
        float l0, l1, l2, l3;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(l0) : "r"(sa[d]));
        asm volatile("ld.shared.f32 %0, [%1+4];" : "=f"(l1) : "r"(sa[d]));
        asm volatile("ld.shared.f32 %0, [%1+176];" : "=f"(l2) : "r"(sa[d]));
        asm volatile("ld.shared.f32 %0, [%1+180];" : "=f"(l3) : "r"(sa[d]));
        r += l0 * a + l1 * b + l2 * cc + l3 * dd;
      }
    }
    __stcs(o + (size_t)c * H * W, r);
  }
}

int main() {
  const size_t plane = (size_t)H * W, tot = (size_t)N * C * plane;
  float *s0, *s1, *fx, *fy, *out;
  CK(cudaMalloc(&s0, tot * 4)); CK(cudaMalloc(&s1, tot * 4)); CK(cudaMalloc(&out, tot * 4));
  CK(cudaMalloc(&fx, 2 * N * plane * 4)); CK(cudaMalloc(&fy, 2 * N * plane * 4));
  {  // host data: sources = index pattern (for the ordering check), flows = bilinear upsample of a 16-px lattice, sigma 8 px
    std::vector<float> h(tot);
    for (size_t k = 0; k < tot; ++k) h[k] = (float)((k * 2654435761u) >> 8 & 0xffff) / 65536.f;
    CK(cudaMemcpy(s0, h.data(), tot * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(s1, h.data(), tot * 4, cudaMemcpyHostToDevice));
    std::vector<float> f(2 * N * plane);
    for (int comp = 0; comp < 2; ++comp) {
      const int LH = H / 16 + 2, LW = W / 16 + 2;
      std::vector<float> lat((size_t)2 * N * LH * LW);
      srand(123 + comp);
      for (auto& v : lat) {  // Box-Muller
        const float u1 = (rand() + 1.f) / (RAND_MAX + 2.f), u2 = (rand() + 1.f) / (RAND_MAX + 2.f);
        v = 8.f * sqrtf(-2.f * logf(u1)) * cosf(6.2831853f * u2);
      }
      for (int dn = 0; dn < 2 * N; ++dn)
        for (int i = 0; i < H; ++i)
          for (int j = 0; j < W; ++j) {
            const float y = i * (LH - 1.f) / (H - 1.f), x = j * (LW - 1.f) / (W - 1.f);
            const int y0 = (int)y, x0 = (int)x, y1 = y0 + 1 < LH ? y0 + 1 : y0, x1 = x0 + 1 < LW ? x0 + 1 : x0;
            const float ty = y - y0, tx = x - x0;
            const float* L = &lat[(size_t)dn * LH * LW];
            f[(size_t)dn * plane + (size_t)i * W + j] = (1 - ty) * ((1 - tx) * L[y0 * LW + x0] + tx * L[y0 * LW + x1]) + ty * ((1 - tx) * L[y1 * LW + x0] + tx * L[y1 * LW + x1]);
          }
      CK(cudaMemcpy(comp ? fy : fx, f.data(), f.size() * 4, cudaMemcpyHostToDevice));
    }
  }
  Tex T;
  for (int d = 0; d < 2; ++d)
    for (int n = 0; n < N; ++n) {
      cudaResourceDesc rd = {};
      rd.resType = cudaResourceTypePitch2D;
      rd.res.pitch2D.devPtr = (d ? s1 : s0) + (size_t)n * C * plane;
      rd.res.pitch2D.desc = cudaCreateChannelDesc<float>();
      rd.res.pitch2D.width = W;
      rd.res.pitch2D.height = C * H;
      rd.res.pitch2D.pitchInBytes = W * 4;
      cudaTextureDesc td = {};
      td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
      td.filterMode = cudaFilterModePoint;
      td.readMode = cudaReadModeElementType;
      td.normalizedCoords = 0;
      CK(cudaCreateTextureObject(&T.t[d][n], &rd, &td, nullptr));
    }
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  std::vector<float> ref(tot), got(tot);
  auto run = [&](auto kern, dim3 grid, const char* name, bool is_ref) -> int {
    kern<<<grid, 256>>>(T, s0, s1, fx, fy, out);
    CK(cudaDeviceSynchronize());
    cudaEventRecord(a);
    for (int r = 0; r < 10; ++r) kern<<<grid, 256>>>(T, s0, s1, fx, fy, out);
    cudaEventRecord(b); CK(cudaEventSynchronize(b));
    float ms; cudaEventElapsedTime(&ms, a, b); ms /= 10;
    CK(cudaMemcpy(got.data(), out, tot * 4, cudaMemcpyDeviceToHost));
    double md = 0;
    if (is_ref) ref = got; else for (size_t k = 0; k < tot; k += 7) md = fmax(md, fabs((double)got[k] - ref[k]));
    printf("%-34s : %.3f ms   %.2f cycles per warp-(dir,channel) per SM   maxdiff vs ldg %.2e\n", name, ms, ms * 1e-3 * 1.965e9 * 148 / ((double)N * plane / 32 * 2 * C), md);
    return 0;
  };
  run(fwd<2, 8>, dim3(W / 32, H / 8, N), "ldg x4, 8x4 patch", true);
  run(fwd<0, 8>, dim3(W / 32, H / 8, N), "tex2Dgather, 8x4 patch", false);
  run(fwd<0, 16>, dim3(W / 64, H / 4, N), "tex2Dgather, 16x2 patch", false);
  run(fwd<0, 4>, dim3(W / 16, H / 16, N), "tex2Dgather, 4x8 patch", false);
  run(fwd<0, 32>, dim3(W / 128, H / 2, N), "tex2Dgather, 32x1 patch", false);
  run(fwd<1, 8>, dim3(W / 32, H / 8, N), "tex2D point x4, 8x4 patch", false);
  run(fwd<4, 8>, dim3(W / 32, H / 8, N), "tld4 + 4 ATOMS", false);
  run(fwd<5, 8>, dim3(W / 32, H / 8, N), "4 ATOMS alone", false);
  run(fwd<6, 8>, dim3(W / 32, H / 8, N), "tld4 + 4 LDS", false);
  run(fwd<3, 8, 4>, dim3(W / 32, H / 8, N), "hybrid 4 ldg + 19 tld4", false);
  run(fwd<3, 8, 7>, dim3(W / 32, H / 8, N), "hybrid 7 ldg + 16 tld4", false);
  run(fwd<3, 8, 10>, dim3(W / 32, H / 8, N), "hybrid 10 ldg + 13 tld4", false);
  run(fwd<3, 16, 7>, dim3(W / 64, H / 4, N), "hybrid 7 ldg + 16 tld4, 16x2", false);
  CK(cudaGetLastError());
  return 0;
}

// I-cache capacity probe for sm_100a: a loop whose body is N independent-ish FFMAs of straight-line code.
// Prints cycles per instruction for body sizes from 4 KB to 256 KB (16 B per SASS instruction), 1 and 4 warps per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O1 -o tools/_ab/icache_probe tools/icache_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#define F1(a, b, c, d) a = fmaf(a, x, y); b = fmaf(b, x, y); c = fmaf(c, x, y); d = fmaf(d, x, y);
#define F4(a, b, c, d) F1(a, b, c, d) F1(a, b, c, d) F1(a, b, c, d) F1(a, b, c, d)
#define F16 F4(a0, a1, a2, a3) F4(a4, a5, a6, a7) F4(a0, a1, a2, a3) F4(a4, a5, a6, a7)   /* 64 FFMA = 1 KB */
#define K4 F16 F16 F16 F16
#define K16 K4 K4 K4 K4
template <int KB>
__global__ void probe(float* out, long long* cyc, int iters, float x, float y) {
  float a0 = threadIdx.x, a1 = 1, a2 = 2, a3 = 3, a4 = 4, a5 = 5, a6 = 6, a7 = 7;
  long long t0 = 0;
  for (int it = 0; it < iters + 1; it++) {
    if (it == 1) t0 = clock64();          // first iteration warms
    if (KB >= 4) { K4 }
    if (KB >= 8) { K4 }
    if (KB >= 16) { K4 K4 }
    if (KB >= 24) { K4 K4 }
    if (KB >= 32) { K4 K4 }
    if (KB >= 40) { K4 K4 }
    if (KB >= 48) { K4 K4 }
    if (KB >= 64) { K16 }
    if (KB >= 96) { K16 K16 }
    if (KB >= 128) { K16 K16 }
    if (KB >= 136) { K4 K4 }
    if (KB >= 144) { K4 K4 }
    if (KB >= 152) { K4 K4 }
    if (KB >= 160) { K4 K4 }
    if (KB >= 176) { K16 }
    if (KB >= 192) { K16 }
    if (KB >= 256) { K16 K16 K16 K16 }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
template <int KB>
void run(int warps, float* out, long long* cyc) {
  const int iters = 8, blocks = 148;
  probe<KB><<<blocks, 32 * warps>>>(out, cyc, iters, 1.0001f, 0.5f);
  probe<KB><<<blocks, 32 * warps>>>(out, cyc, iters, 1.0001f, 0.5f);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double s = 0;
  for (int i = 0; i < blocks; i++) s += h[i];
  s /= blocks;
  printf("body %3d KB  warps/SM %d : %7.3f cycles per instruction per warp (%.0f cycles per 128-byte line)\n", KB, warps,
         s / (iters * KB * 64.0), s / (iters * KB * 8.0));
}
int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 8);
  for (int w : {4}) {
    run<4>(w, out, cyc); run<8>(w, out, cyc); run<16>(w, out, cyc); run<24>(w, out, cyc); run<32>(w, out, cyc);
    run<40>(w, out, cyc); run<48>(w, out, cyc); run<64>(w, out, cyc); run<96>(w, out, cyc); run<128>(w, out, cyc);
    run<136>(w, out, cyc); run<144>(w, out, cyc); run<152>(w, out, cyc); run<160>(w, out, cyc); run<176>(w, out, cyc);
    run<192>(w, out, cyc); run<256>(w, out, cyc);
  }
  return 0;
}

// Instruction-cache capacity probe: a straight-line block of N independent-ish integer instructions (16 bytes each), executed in a loop
// by 1 or 4 warps of one CTA; cycles per instruction vs code size shows where the per-sub-partition L0 and the per-SM L1 instruction
// caches end.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o icache_lab icache_lab.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int N> __global__ void __launch_bounds__(128, 1) probe(unsigned *out, long long *cyc, int iters, int nwarps) {
  unsigned x = threadIdx.x, y = out[0];
  const int warp = threadIdx.x >> 5;
  if (warp >= nwarps)
    return;
  long long t0 = 0;
  for (int it = 0; it < iters + 1; it++) {
    if (it == 1)
      t0 = clock64();
#pragma unroll
    for (int i = 0; i < N / 4; i++) {
      asm volatile("add.u32 %0, %0, %1;" : "+r"(x) : "r"(y));
      asm volatile("xor.b32 %0, %0, %1;" : "+r"(y) : "r"(x));
      asm volatile("add.u32 %0, %0, 7;" : "+r"(x));
      asm volatile("shl.b32 %0, %0, 1;" : "+r"(y));
    }
  }
  const long long t1 = clock64();
  if ((threadIdx.x & 31) == 0)
    cyc[warp] = t1 - t0;
  out[threadIdx.x + 1] = x + y;
}
template <int N> void run() {
  unsigned *out;
  long long *cyc;
  cudaMalloc(&out, 4096);
  cudaMemset(out, 0, 4096);
  cudaMalloc(&cyc, 64);
  for (int nw : {1, 4}) {
    const int iters = 50;
    probe<N><<<1, 128>>>(out, cyc, iters, nw);
    cudaDeviceSynchronize();
    long long h[4];
    cudaMemcpy(h, cyc, 32, cudaMemcpyDeviceToHost);
    printf("code %4d KB  warps %d: %.2f cycles / instruction   %s\n", N * 16 / 1024, nw, (double)h[0] / iters / N, cudaGetErrorString(cudaGetLastError()));
  }
  cudaFree(out);
  cudaFree(cyc);
}
int main() {
  run<256>();
  run<512>();
  run<1024>();
  run<2048>();
  run<3072>();
  run<4096>();
  run<6144>();
  run<8192>();
  run<12288>();
  run<16384>();
  run<32768>();
  return 0;
}

// Development aid: raw FP64 throughput of the DMMA shapes and of plain DFMA on this GPU
// (register-only loops, no memory traffic).  nvcc -arch=sm_100a -O3 dmma_peak.cu -o dmma_peak
#include <cstdio>
#include <cuda_runtime.h>

template <int SHAPE>
__global__ void __launch_bounds__(256) k_dmma(double* out, int iters) {
  double a[8], b[4], c[8][4];
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + i;
  for (int i = 0; i < 4; ++i) b[i] = threadIdx.x * 2e-3 + i;
  for (int i = 0; i < 8; ++i)
    for (int j = 0; j < 4; ++j) c[i][j] = 0.0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (SHAPE == 0) {  // m8n8k4: 2 accumulators
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a[i]), "d"(b[0]));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c[i][2]), "+d"(c[i][3]) : "d"(a[i]), "d"(b[1]));
      } else if (SHAPE == 1) {  // m16n8k4
        asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                     : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(a[i]), "d"(a[(i + 1) & 7]), "d"(b[0]));
      } else if (SHAPE == 2) {  // m16n8k8
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                     : "d"(a[i]), "d"(a[(i + 1) & 7]), "d"(a[(i + 2) & 7]), "d"(a[(i + 3) & 7]), "d"(b[0]), "d"(b[1]));
      } else if (SHAPE == 3) {  // m16n8k16
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                     : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                     : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
      } else {  // DFMA: 4 per i
#pragma unroll
        for (int j = 0; j < 4; ++j) c[i][j] = fma(a[i], b[j], c[i][j]);
      }
    }
  }
  double s = 0;
  for (int i = 0; i < 8; ++i)
    for (int j = 0; j < 4; ++j) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int SHAPE>
void run(const char* name, double flop_per_warp_iter, int blocks_per_sm) {
  int dev = 0, sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  double* out;
  const int blocks = sms * blocks_per_sm;
  cudaMalloc(&out, (size_t)blocks * 256 * 8);
  const int iters = 20000;
  k_dmma<SHAPE><<<blocks, 256>>>(out, 100);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k_dmma<SHAPE><<<blocks, 256>>>(out, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double flops = flop_per_warp_iter * iters * (double)blocks * 8.0;
  printf("%-10s blocks/SM %d: %.2f ms, %.2f TFLOP/s (%s)\n", name, blocks_per_sm, ms, flops / (ms * 1e-3) / 1e12, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
}

int main() {
  for (int bps = 1; bps <= 4; bps *= 2) {
    run<0>("m8n8k4", 16.0 * 2 * 8 * 8 * 4, bps);
    run<1>("m16n8k4", 8.0 * 2 * 16 * 8 * 4, bps);
    run<2>("m16n8k8", 8.0 * 2 * 16 * 8 * 8, bps);
    run<3>("m16n8k16", 8.0 * 2 * 16 * 8 * 16, bps);
    run<4>("dfma", 32.0 * 2 * 32, bps);
  }
  return 0;
}

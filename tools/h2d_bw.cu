// H2D bandwidth probe: pinned vs write-combined pinned host memory, one or two concurrent copy streams.
#include <cstdio>
#include <cstring>
#include <cuda_runtime.h>
static float run(void* h, void* d, size_t bytes, int streams, int reps) {
  cudaStream_t st[4];
  for (int i = 0; i < streams; ++i) cudaStreamCreateWithFlags(&st[i], cudaStreamNonBlocking);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  size_t part = bytes / streams;
  for (int w = 0; w < 2; ++w)
    for (int i = 0; i < streams; ++i) cudaMemcpyAsync((char*)d + i * part, (char*)h + i * part, part, cudaMemcpyHostToDevice, st[i]);
  cudaDeviceSynchronize();
  cudaEventRecord(e0, 0);
  cudaStreamSynchronize(0);
  auto t0 = e0;
  (void)t0;
  float best = 1e9f;
  for (int r = 0; r < reps; ++r) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a, st[0]);
    for (int i = 1; i < streams; ++i) cudaStreamWaitEvent(st[i], a, 0);
    for (int i = 0; i < streams; ++i) cudaMemcpyAsync((char*)d + i * part, (char*)h + i * part, part, cudaMemcpyHostToDevice, st[i]);
    for (int i = 1; i < streams; ++i) { cudaEvent_t c; cudaEventCreate(&c); cudaEventRecord(c, st[i]); cudaStreamWaitEvent(st[0], c, 0); }
    cudaEventRecord(b, st[0]);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    if (ms < best) best = ms;
  }
  return best;
}
int main() {
  const size_t bytes = 256u << 20;
  void *d, *hp, *hw;
  cudaMalloc(&d, bytes);
  cudaHostAlloc(&hp, bytes, cudaHostAllocDefault);
  cudaHostAlloc(&hw, bytes, cudaHostAllocWriteCombined);
  memset(hp, 1, bytes); memset(hw, 1, bytes);
  for (int s = 1; s <= 2; ++s) {
    float a = run(hp, d, bytes, s, 5), b = run(hw, d, bytes, s, 5);
    printf("streams=%d pinned %.2f ms (%.1f GB/s)  write-combined %.2f ms (%.1f GB/s)\n", s, a, bytes / a / 1e6, b, bytes / b / 1e6);
  }
  // D2H
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a); cudaMemcpyAsync(hp, d, bytes, cudaMemcpyDeviceToHost); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  printf("d2h pinned %.2f ms (%.1f GB/s)\n", ms, bytes / ms / 1e6);
  return 0;
}

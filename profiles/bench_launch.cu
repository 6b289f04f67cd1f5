// Host cost of kernel launches vs one graph launch on this box (profiles/ only).
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)
struct Big { char b[320]; };
__global__ void k_empty(int* p) { if (p && threadIdx.x == 9999) *p = 1; }
__global__ void k_big(Big b, int* p) { if (p && threadIdx.x == 9999) *p = b.b[0]; }
int main() {
  cudaStream_t s; CK(cudaStreamCreate(&s));
  int* d; CK(cudaMalloc(&d, 4)); Big big{};
  auto now = [] { return std::chrono::steady_clock::now(); };
  auto us = [](auto a, auto b) { return std::chrono::duration<double, std::micro>(b - a).count(); };
  for (int i = 0; i < 100; ++i) k_empty<<<1, 32, 0, s>>>(d);
  CK(cudaStreamSynchronize(s));
  for (int rep = 0; rep < 3; ++rep) {
    auto t0 = now();
    for (int i = 0; i < 900; ++i) k_empty<<<148, 256, 0, s>>>(d);
    auto t1 = now(); CK(cudaStreamSynchronize(s)); auto t2 = now();
    printf("900 small launches: host %.2f us/launch, total incl. sync %.2f us/launch\n", us(t0, t1) / 900, us(t0, t2) / 900);
  }
  { auto t0 = now(); for (int i = 0; i < 900; ++i) k_big<<<148, 256, 0, s>>>(big, d); auto t1 = now(); CK(cudaStreamSynchronize(s));
    printf("900 launches with 320 B params: host %.2f us/launch\n", us(t0, t1) / 900); }
  // bursts of 9 with a sync in between (the extraction pattern)
  { double acc = 0; for (int r = 0; r < 100; ++r) { auto t0 = now(); for (int i = 0; i < 9; ++i) k_empty<<<148, 256, 0, s>>>(d); auto t1 = now(); acc += us(t0, t1); CK(cudaStreamSynchronize(s)); }
    printf("burst of 9 launches after idle: host %.2f us/burst\n", acc / 100); }
  cudaGraph_t g; cudaGraphExec_t ge;
  CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
  for (int i = 0; i < 9; ++i) k_empty<<<148, 256, 0, s>>>(d);
  CK(cudaStreamEndCapture(s, &g)); CK(cudaGraphInstantiate(&ge, g, 0));
  for (int i = 0; i < 10; ++i) CK(cudaGraphLaunch(ge, s));
  CK(cudaStreamSynchronize(s));
  { double acc = 0; for (int r = 0; r < 100; ++r) { auto t0 = now(); CK(cudaGraphLaunch(ge, s)); auto t1 = now(); acc += us(t0, t1); CK(cudaStreamSynchronize(s)); }
    printf("graph of 9 kernels after idle: host %.2f us/launch\n", acc / 100); }
  { auto t0 = now(); for (int i = 0; i < 200; ++i) CK(cudaGraphLaunch(ge, s)); auto t1 = now(); CK(cudaStreamSynchronize(s)); auto t2 = now();
    printf("graph of 9 kernels back to back: host %.2f us/launch, device %.2f us/graph (%.2f us/kernel)\n", us(t0, t1) / 200, us(t0, t2) / 200, us(t0, t2) / 1800); }
  { auto t0 = now(); for (int i = 0; i < 1800; ++i) k_empty<<<148, 256, 0, s>>>(d); CK(cudaStreamSynchronize(s)); auto t2 = now();
    printf("stream of 1800 kernels: %.2f us/kernel end to end\n", us(t0, t2) / 1800); }
  return 0;
}

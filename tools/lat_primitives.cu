// micro-benchmark of API primitives on the GPU box
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
__global__ void k_empty() {}
__global__ void k_read(const double* in, double* out, int n) { int i = threadIdx.x; if (i < n) out[i] = in[i] * 2.0; }
__global__ void k_chain(const unsigned* idx, const double* in, double* out, int n) { int i = threadIdx.x; if (i < n) out[i] = in[idx[i]] + 1.0; }
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main() {
  cudaStream_t st; cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
  double *h_in, *h_out; unsigned* h_idx; cudaHostAlloc(&h_in, 4096, 0); cudaHostAlloc(&h_out, 4096, 0); cudaHostAlloc(&h_idx, 4096, 0);
  for (int i = 0; i < 64; i++) { h_in[i] = i; h_idx[i] = (i * 7) % 64; }
  double *d_in, *d_out; cudaMalloc(&d_in, 4096); cudaMalloc(&d_out, 4096);
  const int N = 2000; double t;
  auto rep = [&](const char* name, auto fn) { for (int i = 0; i < 100; i++) fn(); double t0 = now(); for (int i = 0; i < N; i++) fn(); printf("%-46s %.2f us\n", name, (now() - t0) / N * 1e6); };
  rep("empty kernel + sync", [&] { k_empty<<<1, 32, 0, st>>>(); cudaStreamSynchronize(st); });
  rep("cudaPointerGetAttributes (pinned)", [&] { cudaPointerAttributes a; cudaPointerGetAttributes(&a, h_in); });
  rep("kernel in-place pinned read+write + sync", [&] { k_read<<<1, 64, 0, st>>>(h_in, h_out, 64); cudaStreamSynchronize(st); });
  rep("kernel chained pinned reads (idx->in) + sync", [&] { k_chain<<<1, 64, 0, st>>>(h_idx, h_in, h_out, 64); cudaStreamSynchronize(st); });
  rep("H2D 512B + kernel(dev) + D2H 512B + sync", [&] { cudaMemcpyAsync(d_in, h_in, 512, cudaMemcpyHostToDevice, st); k_read<<<1, 64, 0, st>>>(d_in, d_out, 64); cudaMemcpyAsync(h_out, d_out, 512, cudaMemcpyDeviceToHost, st); cudaStreamSynchronize(st); });
  rep("kernel(dev in) writes pinned out + sync", [&] { k_read<<<1, 64, 0, st>>>(d_in, h_out, 64); cudaStreamSynchronize(st); });
  rep("2 kernels in-place + sync", [&] { k_read<<<1, 64, 0, st>>>(h_in, h_out, 64); k_read<<<1, 64, 0, st>>>(h_in, h_out + 64, 64); cudaStreamSynchronize(st); });
  // spin on a flag written by the kernel instead of cudaStreamSynchronize
  volatile double* flag = h_out + 200; 
  rep("kernel in-place + spin on last output word", [&] { h_out[63] = -1.0; k_read<<<1, 64, 0, st>>>(h_in, h_out, 64); while (((volatile double*)h_out)[63] < 0) {} });
  cudaStreamSynchronize(st);
  (void)flag; (void)t; return 0; }

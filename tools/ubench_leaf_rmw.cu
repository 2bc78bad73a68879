// Microbenchmark: what HBM bandwidth does a 2 KB-granular read-modify-write reach on this GPU, sequential vs random
// leaf order? (Sets the practical roofline of apply_update_kernel, whose unit of work is one 2 KB map leaf.)
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/ubench tools/ubench_leaf_rmw.cu && /tmp/ubench
#include <cstdio>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <random>
#include <cuda_runtime.h>
__device__ __forceinline__ void ld256(const float* p, float (&v)[8]) {
  asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(v[0]),"=f"(v[1]),"=f"(v[2]),"=f"(v[3]),"=f"(v[4]),"=f"(v[5]),"=f"(v[6]),"=f"(v[7]) : "l"(p));
}
__device__ __forceinline__ void st256(float* p, const float (&v)[8]) {
  asm volatile("st.global.v8.f32 [%8], {%0,%1,%2,%3,%4,%5,%6,%7};" :: "f"(v[0]),"f"(v[1]),"f"(v[2]),"f"(v[3]),"f"(v[4]),"f"(v[5]),"f"(v[6]),"f"(v[7]),"l"(p) : "memory");
}
template <int READ, int WRITE, int PF>
__global__ void __launch_bounds__(256) rmw(float* pool, const uint32_t* order, uint32_t n) {
  const int lane = threadIdx.x & 31;
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t i = warp; i < n; i += nw) {
    float* p = pool + size_t(order[i]) * 512 + lane * 16;
    if (PF && i + PF * nw < n) { const float* q = pool + size_t(order[i + PF * nw]) * 512 + lane * 16; asm volatile("prefetch.global.L2 [%0];" :: "l"(q)); }
    float a[8], b[8];
    if (READ) { ld256(p, a); ld256(p + 8, b); } else { for (int k = 0; k < 8; ++k) { a[k] = float(i); b[k] = float(lane); } }
    for (int k = 0; k < 8; ++k) { a[k] += 1.0f; b[k] += 2.0f; }
    if (WRITE) { st256(p, a); st256(p + 8, b); }
    else if (a[0] == 123.456f) pool[0] = b[3];
  }
}
int main() {
  const uint32_t pool_leaves = 1u << 19, n = 356000;   // 1 GB pool, cfg2-sized touched set
  float* pool; uint32_t* d_order;
  cudaMalloc(&pool, size_t(pool_leaves) * 2048); cudaMemset(pool, 0, size_t(pool_leaves) * 2048);
  cudaMalloc(&d_order, n * 4);
  std::vector<uint32_t> seq(n), rnd(n), chunk(n);
  for (uint32_t i = 0; i < n; ++i) seq[i] = i;
  std::mt19937 g(1); rnd = seq; std::shuffle(rnd.begin(), rnd.end(), g);
  // chunked: random order of 32-leaf chunks (what a roughly sorted entry list looks like)
  { std::vector<uint32_t> c(n / 32); for (uint32_t i = 0; i < c.size(); ++i) c[i] = i; std::shuffle(c.begin(), c.end(), g);
    for (uint32_t i = 0; i < n; ++i) chunk[i] = (i / 32 < c.size()) ? c[i / 32] * 32 + (i % 32) : i; }
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto run = [&](const char* name, const std::vector<uint32_t>& o, auto kern, int grid, double bytes_per_leaf) {
    cudaMemcpy(d_order, o.data(), n * 4, cudaMemcpyHostToDevice);
    float best = 1e9;
    for (int it = 0; it < 6; ++it) { cudaEventRecord(e0); kern<<<grid, 256>>>(pool, d_order, n); cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (it) best = std::min(best, ms); }
    printf("%-34s grid %5d  %.3f ms  %.0f GB/s\n", name, grid, best, n * bytes_per_leaf / best / 1e6);
  };
  for (int grid : {148 * 4, 148 * 8}) {
    run("rmw sequential", seq, rmw<1, 1, 0>, grid, 4096);
    run("rmw random leaves", rnd, rmw<1, 1, 0>, grid, 4096);
    run("rmw random 32-leaf chunks", chunk, rmw<1, 1, 0>, grid, 4096);
    run("rmw random leaves + L2 prefetch 2", rnd, rmw<1, 1, 2>, grid, 4096);
    run("read-only random", rnd, rmw<1, 0, 0>, grid, 2048);
    run("write-only random", rnd, rmw<0, 1, 0>, grid, 2048);
    run("read-only sequential", seq, rmw<1, 0, 0>, grid, 2048);
  }
  return 0;
}

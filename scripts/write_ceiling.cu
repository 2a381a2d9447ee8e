// write_ceiling.cu — calibration: what does a B200 sustain for PURE HBM writes, and which
// store flavour / layout / occupancy gets the mailbox-append pattern closest to it?
// (MEASURED_PEAKS.json's 6575 GB/s is a copy: half reads, half writes.)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o write_ceiling write_ceiling.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

enum { PLAIN = 0, CS = 1, WT = 2, NOALLOC = 3, EVICT_FIRST = 4 };
template <int F> __device__ __forceinline__ void st_v8(void* dst, uint32_t v, uint64_t pol) {
  if (F == PLAIN) asm volatile("st.global.v8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(dst), "r"(v) : "memory");
  if (F == CS) { asm volatile("st.global.cs.v4.b32 [%0], {%1,%1,%1,%1};" ::"l"(dst), "r"(v) : "memory"); asm volatile("st.global.cs.v4.b32 [%0+16], {%1,%1,%1,%1};" ::"l"(dst), "r"(v) : "memory"); }
  if (F == WT) { asm volatile("st.global.wt.v4.b32 [%0], {%1,%1,%1,%1};" ::"l"(dst), "r"(v) : "memory"); asm volatile("st.global.wt.v4.b32 [%0+16], {%1,%1,%1,%1};" ::"l"(dst), "r"(v) : "memory"); }
  if (F == NOALLOC) asm volatile("st.global.L1::no_allocate.v8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(dst), "r"(v) : "memory");
  if (F == EVICT_FIRST) asm volatile("st.global.L2::cache_hint.v8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1}, %2;" ::"l"(dst), "r"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ void st_v4(void* dst, uint32_t v) {
  asm volatile("st.global.v4.b32 [%0], {%1,%1,%1,%1};" ::"l"(dst), "r"(v) : "memory");
}
__global__ void linear_v4(uint4* p, size_t n16, uint32_t v) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) st_v4(p + i, v);
}
// chunked: each warp owns one `chunk`-byte region at a time (the mailbox pattern), 32 B per lane per instruction
template <int F>
__global__ void chunked_v8(unsigned char* p, size_t n_chunks, uint32_t chunk, uint32_t stride, uint32_t v) {
  uint64_t pol = 0;
  if (F == EVICT_FIRST) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  const uint32_t lane = threadIdx.x & 31;
  const size_t w = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5, nw = ((size_t)gridDim.x * blockDim.x) >> 5;
  for (size_t c = w; c < n_chunks; c += nw) {
    unsigned char* base = p + c * stride;
    for (uint32_t o = lane * 32; o < chunk; o += 1024) st_v8<F>(base + o, v, pol);
  }
}
// persistent grid + dynamic claiming: each warp takes T chunks at a time from a global counter
__global__ void chunked_dyn(unsigned char* p, size_t n_chunks, uint32_t chunk, uint32_t stride, uint32_t v, unsigned long long* counter, uint32_t T) {
  const uint32_t lane = threadIdx.x & 31;
  for (;;) {
    unsigned long long c0 = 0;
    if (lane == 0) c0 = atomicAdd(counter, (unsigned long long)T);
    c0 = __shfl_sync(0xffffffffu, c0, 0);
    if (c0 >= n_chunks) break;
    for (uint32_t t = 0; t < T && c0 + t < n_chunks; t++) {
      unsigned char* base = p + (c0 + t) * stride;
      for (uint32_t o = lane * 32; o < chunk; o += 1024) st_v8<PLAIN>(base + o, v, 0);
    }
  }
}
int main() {
  const size_t bytes = (size_t)4 << 30;
  unsigned char* d; CK(cudaMalloc(&d, bytes));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms;
  auto report = [&](const char* name, size_t b, int reps) { printf("%-58s %8.1f GB/s\n", name, b * (double)reps / (ms * 1e-3) / 1e9); };
  for (int it = 0; it < 2; it++) {
    cudaEventRecord(e0); for (int r = 0; r < 10; r++) CK(cudaMemsetAsync(d, r, bytes)); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
  }
  report("cudaMemset 4 GiB", bytes, 10);
  for (int g : {148 * 4, 148 * 8, 148 * 16, 148 * 64}) {
    cudaEventRecord(e0); for (int r = 0; r < 10; r++) linear_v4<<<g, 256>>>((uint4*)d, bytes / 16, r); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1); char nm[64]; snprintf(nm, 64, "linear st.v4 grid=%d", g); report(nm, bytes, 10);
  }
  const size_t n_chunks = 65536;
  auto run = [&](auto kern, const char* flav, uint32_t chunk, uint32_t stride, int g) {
    const size_t b = n_chunks * chunk;
    cudaEventRecord(e0);
    for (int r = 0; r < 40; r++) {
      // mailbox layout: stride 32 KiB, append offset rotates; paged layout: stride == chunk, page base rotates
      unsigned char* base = stride == 32768 ? d + (size_t)(r % (32768 / chunk)) * chunk : d + (size_t)(r % 4) * n_chunks * chunk;
      kern<<<g, 256>>>(base, n_chunks, chunk, stride, r);
    }
    cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    char nm[96]; snprintf(nm, 96, "chunk=%5u stride=%5u grid=%4d %s", chunk, stride, g, flav); report(nm, b, 40);
  };
  for (uint32_t chunk : {8192u}) for (uint32_t stride : {32768u, chunk}) for (int g : {148 * 3, 148 * 4, 148 * 8, 148 * 16, 148 * 32, 148 * 64, 8192}) {
    run(chunked_v8<PLAIN>, "plain", chunk, stride, g);
    if (g <= 148 * 4) run(chunked_v8<EVICT_FIRST>, "L2::evict_first", chunk, stride, g);
  }
  // fewer threads per CTA, more CTAs (128-thread CTAs)
  for (uint32_t stride : {32768u, 8192u}) for (int g : {148 * 16, 148 * 64, 16384}) {
    const size_t b = n_chunks * 8192;
    cudaEventRecord(e0);
    for (int r = 0; r < 40; r++) chunked_v8<PLAIN><<<g, 128>>>(stride == 32768 ? d + (size_t)(r % 4) * 8192 : d + (size_t)(r % 4) * n_chunks * 8192, n_chunks, 8192, stride, r);
    cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    char nm[96]; snprintf(nm, 96, "chunk= 8192 stride=%5u grid=%5d x128thr plain", stride, g); report(nm, b, 40);
  }
  {
    unsigned long long* ctr; CK(cudaMalloc(&ctr, 8 * 64));
    for (uint32_t stride : {32768u, 8192u}) for (uint32_t T : {1u, 4u, 16u}) for (int g : {148 * 3, 148 * 4, 148 * 8}) {
      const size_t b = n_chunks * 8192;
      cudaEventRecord(e0);
      for (int r = 0; r < 40; r++) {
        CK(cudaMemsetAsync(ctr, 0, 8));
        chunked_dyn<<<g, 256>>>(stride == 32768 ? d + (size_t)(r % 4) * 8192 : d + (size_t)(r % 4) * n_chunks * 8192, n_chunks, 8192, stride, r, ctr, T);
      }
      cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
      char nm[96]; snprintf(nm, 96, "DYNAMIC chunk=8192 stride=%5u grid=%4d claim=%u", stride, g, T); report(nm, b, 40);
    }
  }
  for (uint32_t chunk : {32768u}) for (int g : {148 * 4, 148 * 8}) { run(chunked_v8<PLAIN>, "plain", chunk, 32768, g); run(chunked_v8<EVICT_FIRST>, "L2::evict_first", chunk, 32768, g); }
  CK(cudaGetLastError());
  return 0;
}

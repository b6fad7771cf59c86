// Zero-copy (GPU stores into pinned host memory) write bandwidth: contiguous vs scattered records.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/zc_probe tools/exp/zc_probe.cu && /tmp/zc_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void contiguous(const uint4* __restrict__ src, uint4* __restrict__ dst, long long chunks) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < chunks; i += (long long)gridDim.x * blockDim.x) dst[i] = src[i];
}
// records of REC bytes at stride STRIDE bytes in dst, every `every`-th record written; LANES = REC / 16 lanes per record
template <int REC, int STRIDE>
__global__ void records(const uint4* __restrict__ src, uint8_t* __restrict__ dst, long long nrec, int every) {
  constexpr int LANES = REC / 16, PER_WARP = 32 / LANES;
  const int lane = threadIdx.x & 31, sub = lane / LANES, part = lane % LANES;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, warps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long k0 = warp * PER_WARP; k0 < nrec; k0 += warps * PER_WARP) {
    const long long k = k0 + sub;
    if (sub < PER_WARP && k < nrec) reinterpret_cast<uint4*>(dst + k * every * STRIDE)[part] = src[k * LANES + part];
  }
}

// 128-byte lines at the given record indices (8 lanes per record)
__global__ void lines_at(const uint4* __restrict__ src, uint8_t* __restrict__ dst, const int* __restrict__ idx, long long nrec) {
  const int lane = threadIdx.x & 31, sub = lane >> 3, part = lane & 7;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, warps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long k0 = warp * 4; k0 < nrec; k0 += warps * 4) {
    const long long k = k0 + sub;
    if (k < nrec) reinterpret_cast<uint4*>(dst + (long long)idx[k] * 128)[part] = src[k * 8 + part];
  }
}

int main() {
  const long long n = 1 << 20;
  const size_t bytes = (size_t)n * 256;
  uint8_t *h, *d;
  CK(cudaHostAlloc(&h, bytes, cudaHostAllocDefault));
  CK(cudaMalloc(&d, bytes));
  CK(cudaMemset(d, 1, bytes));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto report = [&](const char* what, double payload, int reps) {
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
    printf("%-64s %.3f ms  %.1f GB/s\n", what, ms, payload / ms / 1e6);
  };
  const int reps = 5;
  const size_t cb = (size_t)46 << 20;
  CK(cudaMemcpy(h, d, cb, cudaMemcpyDeviceToHost));
  cudaEventRecord(e0); for (int r = 0; r < reps; r++) cudaMemcpyAsync(h, d, cb, cudaMemcpyDeviceToHost); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
  report("DMA 46 MiB", (double)cb, reps);
  for (int g : {148, 148 * 8}) {
    cudaEventRecord(e0); for (int r = 0; r < reps; r++) contiguous<<<g, 256>>>((const uint4*)d, (uint4*)h, cb / 16); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
    char b[96]; snprintf(b, 96, "zero-copy contiguous 46 MiB, grid %d", g); report(b, (double)cb, reps);
  }
  const long long nrec = n / 4;
  cudaEventRecord(e0); for (int r = 0; r < reps; r++) records<176, 176><<<148 * 8, 256>>>((const uint4*)d, h, nrec, 4); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
  report("zero-copy 176-B records, stride 176, every 4th", nrec * 176.0, reps);
  cudaEventRecord(e0); for (int r = 0; r < reps; r++) records<160, 176><<<148 * 8, 256>>>((const uint4*)d, h, nrec, 4); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
  report("zero-copy 160-B records, stride 176, every 4th", nrec * 160.0, reps);
  cudaEventRecord(e0); for (int r = 0; r < reps; r++) records<128, 128><<<148 * 8, 256>>>((const uint4*)d, h, nrec, 4); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
  report("zero-copy 128-B records, stride 128 (aligned), every 4th", nrec * 128.0, reps);
  cudaEventRecord(e0); for (int r = 0; r < reps; r++) records<256, 256><<<148 * 8, 256>>>((const uint4*)d, h, nrec, 4); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
  report("zero-copy 256-B records, stride 256 (aligned), every 4th", nrec * 256.0, reps);
  cudaEventRecord(e0); for (int r = 0; r < reps; r++) records<64, 64><<<148 * 8, 256>>>((const uint4*)d, h, nrec, 4); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
  report("zero-copy 64-B records, stride 64 (aligned), every 4th", nrec * 64.0, reps);
  cudaEventRecord(e0); for (int r = 0; r < reps; r++) records<16, 16><<<148 * 8, 256>>>((const uint4*)d, h, n, 1); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
  report("zero-copy 16-B records contiguous (sel plane)", n * 16.0, reps);
  cudaEventRecord(e0); for (int r = 0; r < reps; r++) records<176, 176><<<148 * 8, 256>>>((const uint4*)d, h, n, 1); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
  report("zero-copy 176-B records, every record", n * 176.0, reps);
  {
    // 2^18 of 2^20 line slots, chosen at random; written in ascending order, in random order, and in "interleaved runs"
    // (the order the step's lists have: runs of ~32 ascending indices from many warps)
    int* hidx = new int[nrec]; int* didx; CK(cudaMalloc(&didx, nrec * 4));
    unsigned long long s = 88172645463325252ull;
    auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; };
    for (long long i = 0; i < nrec; i++) hidx[i] = (int)(i * 4 + rnd() % 4);
    CK(cudaMemcpy(didx, hidx, nrec * 4, cudaMemcpyHostToDevice));
    cudaEventRecord(e0); for (int r = 0; r < reps; r++) lines_at<<<148 * 8, 256>>>((const uint4*)d, h, didx, nrec); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
    report("zero-copy 128-B lines, random 1-in-4 slots, ascending order", nrec * 128.0, reps);
    for (long long i = nrec - 1; i > 0; i--) { long long j = rnd() % (i + 1); int t = hidx[i]; hidx[i] = hidx[j]; hidx[j] = t; }
    CK(cudaMemcpy(didx, hidx, nrec * 4, cudaMemcpyHostToDevice));
    cudaEventRecord(e0); for (int r = 0; r < reps; r++) lines_at<<<148 * 8, 256>>>((const uint4*)d, h, didx, nrec); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
    report("zero-copy 128-B lines, same slots, random order", nrec * 128.0, reps);
  }
  return 0;
}

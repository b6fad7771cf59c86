// Practical HBM write / mixed rates on B200 with the access styles the step kernels use.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/exp/write_bw tools/exp/write_bw.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void st_v4(uint4* dst, size_t n16) {
  size_t stride = (size_t)gridDim.x * blockDim.x;
  uint4 v = make_uint4(1, 2, 3, threadIdx.x);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) dst[i] = v;
}
__global__ void copy_v4(uint4* dst, const uint4* src, size_t n16, int wr_per_rd) {
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) {
    uint4 v = src[i];
    for (int k = 0; k < wr_per_rd; k++) dst[i + k * n16] = v;
  }
}
// bulk stores from shared memory, one 4608+7680 B pair per warp iteration (the main pass's shape)
__global__ void bulk_store(uint8_t* dst, size_t bytes, int chunk) {
  extern __shared__ __align__(128) uint8_t smem[];
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* buf = smem + (size_t)warp * chunk;
  for (int i = lane * 16; i < chunk; i += 512) *reinterpret_cast<uint4*>(buf + i) = make_uint4(i, 1, 2, 3);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  size_t nchunks = bytes / chunk;
  size_t wid = (size_t)blockIdx.x * (blockDim.x >> 5) + warp, wcnt = (size_t)gridDim.x * (blockDim.x >> 5);
  if (lane == 0) {
    for (size_t c = wid; c < nchunks; c += wcnt) {
      uint32_t s = (uint32_t)__cvta_generic_to_shared(buf);
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + c * chunk), "r"(s), "r"(chunk) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

template <typename F> float timed(F f, int reps = 20) {
  for (int i = 0; i < 3; i++) f();
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  for (int i = 0; i < reps; i++) f();
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  return ms / reps;
}

int main() {
  size_t N = (size_t)768 << 20;
  uint8_t *src, *dst;
  cudaMalloc(&src, N); cudaMalloc(&dst, N);
  cudaMemset(src, 1, N);
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  size_t W = (size_t)512 << 20;
  float t = timed([&] { cudaMemsetAsync(dst, 0, W); });
  printf("cudaMemset 512 MB: %.1f us %.2f TB/s\n", t * 1e3, W / t / 1e9);
  for (int bpsm : {8, 16, 32}) {
    t = timed([&] { st_v4<<<sms * bpsm, 256>>>((uint4*)dst, W / 16); });
    printf("st.v4 write 512 MB (%d blocks/SM): %.1f us %.2f TB/s\n", bpsm, t * 1e3, W / t / 1e9);
  }
  t = timed([&] { copy_v4<<<sms * 16, 256>>>((uint4*)dst, (const uint4*)src, (W / 2) / 16, 1); });
  printf("copy 256->256 MB: %.1f us %.2f TB/s\n", t * 1e3, W / t / 1e9);
  size_t R = (size_t)160 << 20;
  t = timed([&] { copy_v4<<<sms * 16, 256>>>((uint4*)dst, (const uint4*)src, R / 16, 2); });
  printf("read 160 write 320 MB: %.1f us %.2f TB/s\n", t * 1e3, 3 * R / t / 1e9);
  for (int chunk : {4608, 7680, 12288}) {
    int warps = 4, ctas = 3;
    size_t smem = (size_t)warps * chunk;
    cudaFuncSetAttribute(bulk_store, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    size_t bytes = (W / chunk) * chunk;
    t = timed([&] { bulk_store<<<sms * ctas, warps * 32, smem>>>(dst, bytes, chunk); });
    printf("bulk store chunks of %d B (12 warps/SM): %.1f us %.2f TB/s\n", chunk, t * 1e3, bytes / t / 1e9);
    ctas = 8;
    t = timed([&] { bulk_store<<<sms * ctas, warps * 32, smem>>>(dst, bytes, chunk); });
    printf("bulk store chunks of %d B (32 warps/SM): %.1f us %.2f TB/s\n", chunk, t * 1e3, bytes / t / 1e9);
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}

// Does cp.async.bulk.wait_group.read (TMA store read-out) also wait for outstanding cp.async (LDGSTS) groups?
// Both compile to DEPBAR.LE SB0 in SASS.  One warp: [optional] cold cp.async of 16 B per lane + commit, then a bulk
// store of 2 KB + commit + wait_group.read 0, timed with clock64.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void k(const char* cold, char* out, long long* t, int with_cp, size_t stride) {
  __shared__ __align__(128) char sm[4096];
  const int lane = threadIdx.x;
  for (int i = lane; i < 2048 / 4; i += 32) reinterpret_cast<int*>(sm)[i] = i;
  __syncwarp();
  const uint32_t sa = (uint32_t)__cvta_generic_to_shared(sm);
  long long t0 = clock64();
  if (with_cp) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa + 2048 + lane * 16), "l"(cold + (size_t)(lane + 32 * blockIdx.x) * stride) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  long long t1 = clock64();
  if (lane == 0) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 2048;" ::"l"(out + (size_t)blockIdx.x * 2048), "r"(sa) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }
  long long t2 = clock64();
  if (with_cp) asm volatile("cp.async.wait_group 0;" ::: "memory");
  long long t3 = clock64();
  if (lane == 0) { t[blockIdx.x * 3] = t1 - t0; t[blockIdx.x * 3 + 1] = t2 - t1; t[blockIdx.x * 3 + 2] = t3 - t2; }
}

int main() {
  const size_t stride = 1 << 20, n = 64;
  char *cold, *out;
  long long* t;
  cudaMalloc(&cold, stride * 32 * n);
  cudaMalloc(&out, 2048 * n);
  cudaMallocManaged(&t, n * 3 * sizeof(long long));
  for (int with_cp = 0; with_cp < 2; ++with_cp)
    for (int rep = 0; rep < 3; ++rep) {
      cudaMemset(cold, rep, stride * 32 * n);   // evicts nothing useful, but keeps the lines out of L1
      k<<<(unsigned)n, 32>>>(cold, out, t, with_cp, stride);
      if (cudaDeviceSynchronize() != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
      double a = 0, b = 0, c = 0;
      for (size_t i = 0; i < n; ++i) { a += t[i * 3]; b += t[i * 3 + 1]; c += t[i * 3 + 2]; }
      printf("with_cp=%d rep=%d: issue %.0f  bulk store+wait_group.read %.0f  cp.async.wait_group after %.0f cycles\n", with_cp, rep, a / n, b / n, c / n);
    }
  return 0;
}

// Hop latency of a verdict chain through shared-memory bytes (the leaderboard replay's early phase in miniature):
// step s is "decided" by warp owner(s) once steps < s are decided; every warp needs every verdict.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o smem_chain smem_chain.cu && ./smem_chain
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void chain(int steps, int mode, int sleep_ns, long long* out) {
  extern __shared__ uint8_t smem[];
  volatile uint8_t* dec = smem;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, W = blockDim.x >> 5;
  for (int i = threadIdx.x; i < steps; i += blockDim.x) dec[i] = 0;
  __syncthreads();
  const long long t0 = clock64();
  if (mode == 0) {  // every warp visits every step: wait for it, or decide it
    for (int s = 0; s < steps; ++s) {
      const int own = (s * 7 + (s >> 5)) % W;
      if (own == warp) {
        if (lane == 0) dec[s] = 1;
        __syncwarp();
      } else {
        while (dec[s] == 0) { if (sleep_ns) __nanosleep(sleep_ns); }
      }
    }
  } else {          // words of 32 steps: lane e polls step 32w+e; the warp stops at its own steps only
    for (int w0 = 0; w0 < steps; w0 += 32) {
      uint32_t mine = __ballot_sync(0xffffffffu, (((w0 + lane) * 7 + ((w0 + lane) >> 5)) % W) == warp);
      uint32_t mm = 0xffffffffu;
      while (mm) {
        const uint32_t left = mm & mine;
        const int o = left ? __ffs(left) - 1 : 32;
        const uint32_t seg = mm & ~mine & (o >= 32 ? 0xffffffffu : ((1u << o) - 1u));
        if (seg) {
          const bool poll = (seg >> lane) & 1u;
          uint8_t d = 1;
          while (true) {
            if (poll) d = dec[w0 + lane];
            if (__all_sync(0xffffffffu, d != 0)) break;
            if (sleep_ns) __nanosleep(sleep_ns);
          }
          mm &= ~seg;
        }
        if (o < 32) {
          if (lane == 0) dec[w0 + o] = 1;
          __syncwarp();
          mm &= ~(1u << o);
        }
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) out[0] = clock64() - t0;
}

int main() {
  long long* d;
  cudaMalloc(&d, 8);
  const int steps = 8192;
  for (int mode = 0; mode < 2; ++mode)
    for (int W : {2, 4, 8, 16, 32})
      for (int ns : {0, 32}) {
        long long h = 0;
        for (int rep = 0; rep < 2; ++rep) {
          chain<<<1, 32 * W, steps>>>(steps, mode, ns, d);
          cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
        }
        printf("mode %d warps %2d sleep %2d ns: %6.0f clocks per step\n", mode, W, ns, (double)h / steps);
      }
  return cudaGetLastError() != cudaSuccess;
}

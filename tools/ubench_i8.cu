// Micro-benchmark: issue rate of tcgen05.mma kind::i8 (UTCIMMA) and kind::f16 (bf16) on B200, one CTA per
// SM, operands = whatever is in shared memory (SWIZZLE_128B K-major descriptors), accumulators in TMEM.
// Gives the roofline denominator for the int8-slice scoring kernel (MEASURED_PEAKS.json has bf16 only).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_i8 ubench_i8.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); return 1;}}while(0)

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

template <int KIND /*0 = i8, 1 = bf16*/, int N>
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int iters, unsigned* sink) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* tiles = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"r"(smem_u32(&tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  // M = 128, N, K-major A/B, SWIZZLE_128B; i8: S8 x S8 -> S32; bf16: BF16 x BF16 -> F32
  const uint32_t idesc = (KIND == 0) ? ((2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24))
                                     : ((1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24));
  const unsigned a0 = smem_u32(tiles), b0 = a0 + 16384;
  const uint64_t HI = ((uint64_t)(64u | (1u << 14) | (2u << 29))) << 32;
  if (warp == 0 && lane == 0) {
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const uint64_t da = HI | (uint64_t)((((a0 + (j & 3) * 32) & 0x3FFFFu) >> 4) | 0x10000u);
        const uint64_t db = HI | (uint64_t)((((b0 + (j & 3) * 32) & 0x3FFFFu) >> 4) | 0x10000u);
        const unsigned acc = (unsigned)((j & 1) * 256 % 512);
        if (KIND == 0)
          asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n"
                       ::"r"(acc), "l"(da), "l"(db), "r"(idesc), "r"(1) : "memory");
        else
          asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                       ::"r"(acc), "l"(da), "l"(db), "r"(idesc), "r"(1) : "memory");
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&bar)) : "memory");
    unsigned ok = 0;
    while (!ok)
      asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0,1,0,p;\n}\n" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
    sink[blockIdx.x] = ok;
  }
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(0u) : "memory");
  }
}

template <int KIND, int N>
int run(const char* name, int sms, unsigned* sink) {
  const size_t smem = 16384 + 32768 + 2048;
  CK(cudaFuncSetAttribute(mma_rate_kernel<KIND, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  mma_rate_kernel<KIND, N><<<sms, 128, smem>>>(200, sink); CK(cudaDeviceSynchronize());
  cudaEventRecord(e0); mma_rate_kernel<KIND, N><<<sms, 128, smem>>>(iters, sink); cudaEventRecord(e1);
  CK(cudaEventSynchronize(e1));
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double k_per = (KIND == 0) ? 32.0 : 16.0;
  const double macs = 128.0 * N * k_per * 8.0 * iters * sms;
  printf("%-22s N=%3d: %.3f ms  %.1f TOP/s (2*MAC)  %.0f MAC/clk/SM @1.965GHz  %.1f clk/MMA\n", name, N, ms,
         2 * macs / ms * 1e-9, macs / (ms * 1e-3) / sms / 1.965e9, ms * 1e-3 * 1.965e9 / (8.0 * iters));
  return 0;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  unsigned* sink; CK(cudaMalloc(&sink, 4096));
  const int sms = p.multiProcessorCount;
  printf("device %s SMs %d\n", p.name, sms);
  if (run<0, 64>("tcgen05 kind::i8", sms, sink)) return 1;
  if (run<0, 128>("tcgen05 kind::i8", sms, sink)) return 1;
  if (run<0, 256>("tcgen05 kind::i8", sms, sink)) return 1;
  if (run<1, 64>("tcgen05 kind::f16 bf16", sms, sink)) return 1;
  if (run<1, 256>("tcgen05 kind::f16 bf16", sms, sink)) return 1;
  return 0;
}

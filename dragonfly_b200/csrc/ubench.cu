// Live roofline denominators for bench.py (dfb_measure_peak): the issue rate of tcgen05.mma kind::i8 and of fp64
// DMMA.8x8x4 on the device the bench runs on, measured in the same process as the timed step instead of being
// quoted from an earlier run (MEASURED_PEAKS.json carries HBM GB/s and bf16 TF/s only).  Not on the product path.
#include "common.cuh"

namespace dfb {

__device__ __forceinline__ unsigned ub_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

// One CTA per SM; operands are whatever sits in shared memory (SWIZZLE_128B K-major descriptors), accumulators in
// tensor memory; one thread issues 8 x iters MMAs of shape M128 N256 K32 (signed int8 -> int32) and waits for the
// commit.
__global__ void __launch_bounds__(128, 1) ub_i8_rate_kernel(int iters, unsigned* sink) {
  constexpr int N = 256;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* tiles = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(ub_smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"r"(ub_smem_u32(&tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
  const unsigned a0 = ub_smem_u32(tiles), b0 = a0 + 16384;
  const uint64_t HI = ((uint64_t)(64u | (1u << 14) | (2u << 29))) << 32;
  if (warp == 0 && lane == 0) {
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const uint64_t da = HI | (uint64_t)((((a0 + (j & 3) * 32) & 0x3FFFFu) >> 4) | 0x10000u);
        const uint64_t db = HI | (uint64_t)((((b0 + (j & 3) * 32) & 0x3FFFFu) >> 4) | 0x10000u);
        const unsigned acc = (unsigned)((j & 1) * 256 % 512);
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n"
                     ::"r"(acc), "l"(da), "l"(db), "r"(idesc), "r"(1) : "memory");
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(ub_smem_u32(&bar)) : "memory");
    unsigned ok = 0;
    while (!ok)
      asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0,1,0,p;\n}\n" : "=r"(ok) : "r"(ub_smem_u32(&bar)) : "memory");
    sink[blockIdx.x] = ok;
  }
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(0u) : "memory");
  }
}

// 16 warps per SM, 8 independent DMMA.8x8x4 accumulators per thread
__global__ void ub_dmma_rate_kernel(double* out, int iters) {
  double c[8][2];
#pragma unroll
  for (int i = 0; i < 8; i++) { c[i][0] = 0; c[i][1] = 0; }
  const double av = 1.0 + threadIdx.x * 1e-9, bv = 1.0 - threadIdx.x * 1e-9;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(av), "d"(bv));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace dfb

using namespace dfb;

extern "C" int dfb_measure_peak(int device, int what, double* out) {
  if (out == nullptr || (what != DFB_PEAK_TCGEN05_I8 && what != DFB_PEAK_DMMA_F64)) { set_error("bad measure_peak arguments"); return -1; }
  DFB_CUDA_OK(cudaSetDevice(device));
  int sms = 0;
  DFB_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
  cudaEvent_t e0, e1;
  DFB_CUDA_OK(cudaEventCreate(&e0));
  DFB_CUDA_OK(cudaEventCreate(&e1));
  void* scratch = nullptr;
  DFB_CUDA_OK(cudaMalloc(&scratch, sizeof(double) * (size_t)sms * 512 + 4096));
  float best = 1e30f;
  double work = 0.0;
  if (what == DFB_PEAK_TCGEN05_I8) {
    const size_t smem = 16384 + 32768 + 2048;
    DFB_CUDA_OK(cudaFuncSetAttribute(ub_i8_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int iters = 8000;
    ub_i8_rate_kernel<<<sms, 128, smem>>>(200, (unsigned*)scratch);
    DFB_CUDA_OK(cudaDeviceSynchronize());
    for (int rep = 0; rep < 3; rep++) {
      DFB_CUDA_OK(cudaEventRecord(e0));
      ub_i8_rate_kernel<<<sms, 128, smem>>>(iters, (unsigned*)scratch);
      DFB_CUDA_OK(cudaEventRecord(e1));
      DFB_CUDA_OK(cudaEventSynchronize(e1));
      float ms = 0.f;
      DFB_CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
      if (ms < best) best = ms;
    }
    work = 2.0 * 128.0 * 256.0 * 32.0 * 8.0 * iters * sms;        // int8 ops (2 per MAC)
  } else {
    const int iters = 8000, threads = 512;
    ub_dmma_rate_kernel<<<sms, threads>>>((double*)scratch, 100);
    DFB_CUDA_OK(cudaDeviceSynchronize());
    for (int rep = 0; rep < 3; rep++) {
      DFB_CUDA_OK(cudaEventRecord(e0));
      ub_dmma_rate_kernel<<<sms, threads>>>((double*)scratch, iters);
      DFB_CUDA_OK(cudaEventRecord(e1));
      DFB_CUDA_OK(cudaEventSynchronize(e1));
      float ms = 0.f;
      DFB_CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
      if (ms < best) best = ms;
    }
    work = 2.0 * 256.0 * 8.0 * (double)iters * (threads / 32) * sms;   // flops
  }
  cudaFree(scratch);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *out = work / (best * 1e-3) * 1e-12;       // TOP/s or TFLOP/s
  return 0;
}

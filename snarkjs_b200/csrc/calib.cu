// calib.cu — calibration micro-kernels for the integer-pipe roofline: the measured rate of
// IMAD.WIDE.U32 (the instruction that carries >95% of the work) and of back-to-back register-resident
// Montgomery multiplies.  bench.py reports kernel throughput against these measured peaks.
#include <cuda_runtime.h>
#include "fp.cuh"
namespace sb {
__global__ void __launch_bounds__(256) k_calib_imad(uint64_t* out, int iters, uint32_t a, uint32_t b) {
    uint64_t acc[8];
#pragma unroll
    for (int j = 0; j < 8; j++) acc[j] = threadIdx.x + j;
    uint32_t x = a + threadIdx.x, y = b;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
#pragma unroll
            for (int j = 0; j < 8; j++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[j]) : "r"((uint32_t)acc[(j + 3) & 7]), "r"(y));   // data dependent multiplicand: no strength reduction
        }
    }
    uint64_t s = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) s ^= acc[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void __launch_bounds__(256) k_calib_modmul(uint32_t* out, int iters) {
    typedef Fp<BnFq> F;
    F a = F::one(), b = F::r2(), c = F::one(), d = F::r2();
    a.v[0] += threadIdx.x; c.v[1] += blockIdx.x + 3 * threadIdx.x; b.v[2] ^= threadIdx.x; d.v[3] += 7 * threadIdx.x;   // every chain is per-thread (nothing for the uniform datapath)
    for (int it = 0; it < iters; it++) { a = F::mul(a, b); c = F::mul(c, d); b = F::mul(b, a); d = F::mul(d, c); }
    F r = F::add(F::add(a, b), F::add(c, d));
    out[blockIdx.x * blockDim.x + threadIdx.x] = r.v[0] ^ r.v[7];
}
// returns operations per second (what = 0: IMAD.WIDE.U32, 1: BN254 Fq Montgomery multiplies), <0 on error
double calibrate(int what, cudaStream_t stream) {
    int dev = 0, sms = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int blocks = sms * 8, threads = 256;
    void* buf = nullptr; if (cudaMalloc(&buf, (size_t)blocks * threads * 8) != cudaSuccess) return -1;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = what == 0 ? 4096 : 512;
    double best = 0;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0, stream);
        if (what == 0) k_calib_imad<<<blocks, threads, 0, stream>>>((uint64_t*)buf, iters, 12345u + rep, 0x9e3779b9u);
        else k_calib_modmul<<<blocks, threads, 0, stream>>>((uint32_t*)buf, iters);
        cudaEventRecord(e1, stream);
        if (cudaEventSynchronize(e1) != cudaSuccess) { best = -1; break; }
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
        double ops = (double)blocks * threads * iters * (what == 0 ? 64.0 : 4.0);
        if (rep > 0 && ms > 0) best = ops / (ms * 1e-3) > best ? ops / (ms * 1e-3) : best;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(buf);
    return best;
}
}

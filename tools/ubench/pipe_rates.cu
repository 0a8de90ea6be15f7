// pipe_rates.cu -- per-SM issue rate of the instructions the fast beam kernel leans on (sm_100a).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_rates pipe_rates.cu && ./pipe_rates
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

template <int OP> __global__ void k(float *out, int iters, float seed)
{
    float a = seed + threadIdx.x, b = seed * 3.f + threadIdx.x, c = seed * 5.f, d = seed * 7.f;
    uint32_t u = 0, v = 0;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int r = 0; r < 16; r++) {
            if (OP == 0) {            // F2FP.BF16.F32.PACK_AB, 4 independent chains
                asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u) : "f"(a), "f"(b));
                asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(v) : "f"(c), "f"(d));
                a += __uint_as_float(u << 16) * 1e-30f; c += __uint_as_float(v << 16) * 1e-30f;
            } else if (OP == 1) {     // FFMA reference (same dependent adds as OP 0 without the cvt)
                a += b * 1e-30f; c += d * 1e-30f;
            } else if (OP == 2) {     // MUFU.EX2
                asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(a) : "f"(a));
                asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(c) : "f"(c));
            } else if (OP == 3) {     // integer RN split: hi = rn_bf16(x) by LOP3/IADD
                uint32_t xa = __float_as_uint(a), xc = __float_as_uint(c);
                xa = (xa + 0x7fffu + ((xa >> 16) & 1u)) & 0xffff0000u;
                xc = (xc + 0x7fffu + ((xc >> 16) & 1u)) & 0xffff0000u;
                a += __uint_as_float(xa) * 1e-30f; c += __uint_as_float(xc) * 1e-30f;
            }
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = a + c + (float)(t1 - t0) * 0.f;
    if (threadIdx.x == 0) reinterpret_cast<long long *>(out + gridDim.x * blockDim.x)[blockIdx.x] = t1 - t0;
}

template <int OP> void run(const char *name, int warps_per_sm, int per_iter)
{
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int threads = warps_per_sm * 32, iters = 2000;
    float *out; cudaMalloc(&out, (size_t)sms * threads * 4 + sms * 8);
    k<OP><<<sms, threads>>>(out, 10, 1.f);
    k<OP><<<sms, threads>>>(out, iters, 1.f);
    cudaDeviceSynchronize();
    long long cyc; cudaMemcpy(&cyc, out + (size_t)sms * threads, 8, cudaMemcpyDeviceToHost);
    double instr = (double)iters * 16 * per_iter * warps_per_sm;      // warp-instructions of the probed kind per SM
    printf("%-28s warps/SM %2d : %.2f clk per warp-instruction per SM (%.1f lanes/clk/SM)\n", name, warps_per_sm, cyc / instr, 32.0 * instr / cyc);
    cudaFree(out);
}

int main()
{
    for (int w : {8, 16, 32}) {
        if (w == 8) { run<0>("F2FP.BF16.F32.PACK_AB", 8, 2); run<1>("FFMA (loop skeleton)", 8, 2); run<2>("MUFU.EX2", 8, 2); run<3>("int RN-to-bf16 (4 ALU ops)", 8, 2); }
        if (w == 16) { run<0>("F2FP.BF16.F32.PACK_AB", 16, 2); run<1>("FFMA (loop skeleton)", 16, 2); run<2>("MUFU.EX2", 16, 2); run<3>("int RN-to-bf16 (4 ALU ops)", 16, 2); }
        if (w == 32) { run<0>("F2FP.BF16.F32.PACK_AB", 32, 2); run<1>("FFMA (loop skeleton)", 32, 2); run<2>("MUFU.EX2", 32, 2); run<3>("int RN-to-bf16 (4 ALU ops)", 32, 2); }
    }
    return 0;
}

// pipe_probe.cu -- issue-rate probe for the integer pipes of one SM sub-partition (sm_100a).
// Eight independent chains per thread, 64 instructions per chain per trip; 16 warps per SM sub-partition.
// Prints warp instructions per cycle per sub-partition for: add only (ALU pipe), mad only (FMA pipe), and mixes.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_probe pipe_probe.cu && ./pipe_probe
#include <cstdio>
#include <cuda_runtime.h>

#define CHAINS 8
#define REP 32

#define OP_ADD(X, Y) asm volatile("add.s32 %0, %0, %1;" : "+r"(X) : "r"(Y))
#define OP_MAD(X, Y) asm volatile("mad.lo.s32 %0, %0, %2, %1;" : "+r"(X) : "r"(Y), "r"(m))
#define OP_SHF(X, Y) asm volatile("shf.r.wrap.b32 %0, %0, %1, 7;" : "+r"(X) : "r"(Y))
#define OP_LOP(X, Y) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(X) : "r"(Y), "r"(m))
#define OP_SEL(X, Y) asm volatile("{.reg .pred p; setp.gt.s32 p, %0, %1; selp.s32 %0, %1, %2, p;}" : "+r"(X) : "r"(Y), "r"(m))
#define OP_FMA(X, Y) asm volatile("fma.rn.f32 %0, %0, %2, %1;" : "+f"(*(float *)&X) : "f"(*(float *)&Y), "f"(1.0001f))

template <int MODE> __global__ void __launch_bounds__(256) k_probe(int *out, int m, int iters)
{
    int x[CHAINS], y[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) { x[c] = threadIdx.x + c; y[c] = threadIdx.x * 3 + c; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < REP; ++r) {
#pragma unroll
            for (int c = 0; c < CHAINS; ++c) {
                int &X = (r & 1) ? y[c] : x[c];
                int &Y = (r & 1) ? x[c] : y[c];
                if (MODE == 0) OP_ADD(X, Y);
                else if (MODE == 1) OP_MAD(X, Y);
                else if (MODE == 2) { if (c & 1) OP_ADD(X, Y); else OP_MAD(X, Y); }
                else if (MODE == 3) { if (c & 3) OP_ADD(X, Y); else OP_MAD(X, Y); }
                else if (MODE == 4) OP_SHF(X, Y);
                else if (MODE == 5) OP_LOP(X, Y);
                else if (MODE == 6) OP_SEL(X, Y);
                else if (MODE == 7) { if (c & 1) OP_SHF(X, Y); else OP_MAD(X, Y); }
                else if (MODE == 8) { if ((c & 3) == 0) OP_MAD(X, Y); else if ((c & 3) == 1) OP_FMA(X, Y); else OP_ADD(X, Y); }
                else if (MODE == 9) OP_FMA(X, Y);
                else if (MODE == 10) { if (c & 1) OP_ADD(X, Y); else OP_FMA(X, Y); }
                else if (MODE == 11) { if (c & 1) OP_MAD(X, Y); else OP_FMA(X, Y); }
            }
        }
    }
    int s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += x[c] ^ y[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE> void run(const char *name, int *d, int sms, double mhz)
{
    const int iters = 2000, blocks = sms * 8; // 8 CTAs x 8 warps = 64 warps per SM = 16 per sub-partition
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    k_probe<MODE><<<blocks, 256>>>(d, 3, 10);
    cudaEventRecord(a);
    k_probe<MODE><<<blocks, 256>>>(d, 3, iters);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    const double inst = (double)iters * REP * CHAINS * (256 / 32) * 8; // warp instructions per SM (probe body only)
    const double cycles = ms * 1e-3 * mhz * 1e6;
    printf("%-34s %8.3f ms  %.3f warp-inst/cycle/sub-partition\n", name, ms, inst / cycles / 4.0);
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double mhz = khz / 1000.0;
    printf("%s, %d SMs, %.0f MHz (max; rates assume the clock holds)\n", p.name, p.multiProcessorCount, mhz);
    int *d;
    cudaMalloc(&d, (size_t)p.multiProcessorCount * 8 * 256 * sizeof(int));
    run<0>("add (ALU)", d, p.multiProcessorCount, mhz);
    run<1>("mad.lo (FMA pipe, integer)", d, p.multiProcessorCount, mhz);
    run<2>("1 add : 1 mad", d, p.multiProcessorCount, mhz);
    run<3>("3 add : 1 mad", d, p.multiProcessorCount, mhz);
    run<4>("shf (ALU)", d, p.multiProcessorCount, mhz);
    run<5>("lop3 (ALU)", d, p.multiProcessorCount, mhz);
    run<6>("setp+selp pair (counted as 1)", d, p.multiProcessorCount, mhz);
    run<7>("1 shf : 1 mad", d, p.multiProcessorCount, mhz);
    run<8>("2 add : 1 mad : 1 ffma", d, p.multiProcessorCount, mhz);
    run<9>("ffma", d, p.multiProcessorCount, mhz);
    run<10>("1 add : 1 ffma", d, p.multiProcessorCount, mhz);
    run<11>("1 mad : 1 ffma", d, p.multiProcessorCount, mhz);
    return 0;
}

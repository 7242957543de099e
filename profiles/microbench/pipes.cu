// pipes.cu — throughput of the integer instructions the pack / search loops are made of, per SM sub-partition, and
// whether alu-pipe and fma-pipe instructions issue side by side.  nvcc -arch=sm_100a -O3 -o pipes pipes.cu && ./pipes
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096
#define UNROLL 8
#ifndef THREADS
#define THREADS 1024
#endif
#ifndef BLOCKS
#define BLOCKS 1
#endif

template <int OP>
__global__ void __launch_bounds__(THREADS, 2048 / THREADS >= 2 && BLOCKS >= 2 ? 2 : 1) k(uint32_t *out, uint32_t seed, long long *cycles)
{
    uint32_t a[UNROLL], b = seed | 1u, c = seed * 3u + 7u;
#pragma unroll
    for (int i = 0; i < UNROLL; ++i) a[i] = threadIdx.x * 17u + i + seed;
    __shared__ uint32_t sm[1024]; sm[threadIdx.x] = seed;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < UNROLL; ++i) {
            uint32_t x = a[i];
            if (OP == 0) asm volatile("shr.u32 %0, %0, 1;" : "+r"(x));                        // SHF
            if (OP == 1) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(x) : "r"(0x80000001u)); // IMAD.HI
            if (OP == 2) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(c)); // IMAD
            if (OP == 3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(b), "r"(c)); // LOP3
            if (OP == 4) asm volatile("prmt.b32 %0, %0, %1, 0x2103;" : "+r"(x) : "r"(b));     // PRMT
            if (OP == 5) asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(c));   // SHF funnel, variable
            if (OP == 6) {                                                                     // LOP3 + IMAD alternating
                if (i & 1) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(c));
                else asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(b), "r"(c));
            }
            if (OP == 7) {                                                                     // LOP3 + IMAD.HI alternating
                if (i & 1) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(x) : "r"(0x80000001u));
                else asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(b), "r"(c));
            }
            if (OP == 8) asm volatile("shl.b32 %0, %0, 3;" : "+r"(x));                         // SHL (IMAD.SHL or SHF?)
            if (OP == 9) asm volatile("add.u32 %0, %0, %1;" : "+r"(x) : "r"(b));               // IADD3
            if (OP == 10) {                                                                    // 3 ALU : 1 FMA
                if ((i & 3) == 3) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(c));
                else asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(b), "r"(c));
            }
            if (OP == 11) {                                                                    // 64-bit multiply (IMAD.WIDE)
                unsigned long long w;
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w) : "r"(x), "r"(b));
                x = (uint32_t)(w >> 32) ^ (uint32_t)w;
            }
            if (OP == 12) {                                                                    // 2 LOP3 : 1 IMAD : 1 LDS
                if ((i & 3) == 3) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(c));
                else if ((i & 3) == 2) { x ^= sm[x & 1023u]; }
                else asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(b), "r"(c));
            }
            if (OP == 13) {                                                                    // LOP3 : IMAD : FFMA
                if (i % 3 == 0) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(b), "r"(c));
                else if (i % 3 == 1) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(c));
                else { float f = __uint_as_float(x); asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f) : "f"(1.0001f), "f"(0.5f)); x = __float_as_uint(f); }
            }
            if (OP == 14) {                                                                    // 1 LOP3 : 2 IMAD
                if (i % 3 == 0) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(b), "r"(c));
                else asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(c));
            }
            a[i] = x;
        }
    }
    const long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < UNROLL; ++i) s ^= a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int OP> void run(const char *name, uint32_t *d_out, long long *d_cyc)
{
    k<OP><<<148 * BLOCKS, THREADS>>>(d_out, 12345u, d_cyc);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<OP><<<148 * BLOCKS, THREADS>>>(d_out, 12345u, d_cyc);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k<OP>, THREADS, 0);
    long long cyc = 0;
    cudaMemcpy(&cyc, d_cyc, sizeof(cyc), cudaMemcpyDeviceToHost);
    // 32 warps per SM = 8 per sub-partition; instructions per sub-partition = 8 warps x ITERS x UNROLL
    const double per_smsp = (THREADS / 128.0) * BLOCKS * ITERS * UNROLL;      // warp instructions per sub-partition
    printf("%-30s %8lld cycles (block 0)  %.3f ms  occupancy %d blocks/SM  %.3f warp-instr/cycle/SMSP by the event time at 1.965 GHz\n",
           name, cyc, ms, occ, per_smsp / (ms * 1e-3 * 1.965e9));
}

int main()
{
    uint32_t *d_out; long long *d_cyc;
    cudaMalloc(&d_out, 148 * BLOCKS * THREADS * 4); cudaMalloc(&d_cyc, 8);
    run<0>("SHF (shr by constant)", d_out, d_cyc);
    run<1>("IMAD.HI (mul.hi.u32)", d_out, d_cyc);
    run<2>("IMAD (mad.lo)", d_out, d_cyc);
    run<3>("LOP3", d_out, d_cyc);
    run<4>("PRMT", d_out, d_cyc);
    run<5>("SHF funnel variable", d_out, d_cyc);
    run<6>("LOP3 / IMAD alternating", d_out, d_cyc);
    run<7>("LOP3 / IMAD.HI alternating", d_out, d_cyc);
    run<8>("SHL by constant", d_out, d_cyc);
    run<9>("IADD", d_out, d_cyc);
    run<10>("3 LOP3 : 1 IMAD", d_out, d_cyc);
    run<11>("IMAD.WIDE + 2 ALU", d_out, d_cyc);
    run<12>("2 LOP3 : 1 IMAD : 1 LDS(+LOP)", d_out, d_cyc);
    run<13>("LOP3 : IMAD : FFMA", d_out, d_cyc);
    run<14>("1 LOP3 : 2 IMAD", d_out, d_cyc);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

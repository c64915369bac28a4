// tools/microbench/pipes.cu -- integer pipe throughput probes for sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu ; run on the GPU box.
// Each kernel runs `iters` x UNROLL copies of one instruction pattern in registers on
// every SM (8 blocks x 256 threads) and reports warp-instructions per clock per SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define REP8(x) x x x x x x x x

template <int V>
__global__ void __launch_bounds__(256) probe(uint32_t seed, int iters, unsigned long long* out, long long* clk)
{
    uint32_t a = seed + threadIdx.x, b = seed * 2654435761u + blockIdx.x;
    uint32_t r0 = a, r1 = b, r2 = a ^ b, r3 = a + b, r4 = a * 3, r5 = b * 5, r6 = a * 7, r7 = b * 9;
    uint32_t r8 = a + 1, r9 = b + 2, r10 = a + 3, r11 = b + 4, r12 = a + 5, r13 = b + 6, r14 = a + 7, r15 = b + 8;
    unsigned long long q0 = a, q1 = b, q2 = r2, q3 = r3, q4 = r4, q5 = r5, q6 = r6, q7 = r7;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (V == 0) {           // plain IMAD.WIDE.U32, 8 independent 64-bit accumulators, data-dependent multiplier
            REP8(asm volatile(
                "{ .reg .b64 t; mov.b64 t, {%0,%1}; mad.wide.u32 t, %2, %16, t; mov.b64 {%0,%1}, t; }"
                "{ .reg .b64 t; mov.b64 t, {%2,%3}; mad.wide.u32 t, %4, %16, t; mov.b64 {%2,%3}, t; }"
                "{ .reg .b64 t; mov.b64 t, {%4,%5}; mad.wide.u32 t, %6, %16, t; mov.b64 {%4,%5}, t; }"
                "{ .reg .b64 t; mov.b64 t, {%6,%7}; mad.wide.u32 t, %8, %16, t; mov.b64 {%6,%7}, t; }"
                "{ .reg .b64 t; mov.b64 t, {%8,%9}; mad.wide.u32 t, %10, %16, t; mov.b64 {%8,%9}, t; }"
                "{ .reg .b64 t; mov.b64 t, {%10,%11}; mad.wide.u32 t, %12, %16, t; mov.b64 {%10,%11}, t; }"
                "{ .reg .b64 t; mov.b64 t, {%12,%13}; mad.wide.u32 t, %14, %16, t; mov.b64 {%12,%13}, t; }"
                "{ .reg .b64 t; mov.b64 t, {%14,%15}; mad.wide.u32 t, %0, %16, t; mov.b64 {%14,%15}, t; }"
                : "+r"(r0), "+r"(r1), "+r"(r2), "+r"(r3), "+r"(r4), "+r"(r5), "+r"(r6), "+r"(r7),
                  "+r"(r8), "+r"(r9), "+r"(r10), "+r"(r11), "+r"(r12), "+r"(r13), "+r"(r14), "+r"(r15) : "r"(b));)
        } else if (V == 1) {    // carry chains: two interleaved rows of 4 IMAD.WIDE.U32.X each (as mul_full emits)
            REP8(asm volatile(
                "mad.lo.cc.u32 %0,%16,%17,%0; madc.hi.cc.u32 %1,%16,%17,%1; madc.lo.cc.u32 %2,%16,%17,%2; madc.hi.cc.u32 %3,%16,%17,%3;"
                "madc.lo.cc.u32 %4,%16,%17,%4; madc.hi.cc.u32 %5,%16,%17,%5; madc.lo.cc.u32 %6,%16,%17,%6; madc.hi.u32 %7,%16,%17,%7;"
                "mad.lo.cc.u32 %8,%16,%17,%8; madc.hi.cc.u32 %9,%16,%17,%9; madc.lo.cc.u32 %10,%16,%17,%10; madc.hi.cc.u32 %11,%16,%17,%11;"
                "madc.lo.cc.u32 %12,%16,%17,%12; madc.hi.cc.u32 %13,%16,%17,%13; madc.lo.cc.u32 %14,%16,%17,%14; madc.hi.u32 %15,%16,%17,%15;"
                : "+r"(r0), "+r"(r1), "+r"(r2), "+r"(r3), "+r"(r4), "+r"(r5), "+r"(r6), "+r"(r7),
                  "+r"(r8), "+r"(r9), "+r"(r10), "+r"(r11), "+r"(r12), "+r"(r13), "+r"(r14), "+r"(r15) : "r"(a), "r"(b));)
        } else if (V == 2) {    // IMAD 32-bit low, 8 independent
            REP8(asm volatile("mad.lo.u32 %0,%1,%9,%0; mad.lo.u32 %1,%2,%9,%1; mad.lo.u32 %2,%3,%9,%2; mad.lo.u32 %3,%4,%9,%3;"
                              "mad.lo.u32 %4,%5,%9,%4; mad.lo.u32 %5,%6,%9,%5; mad.lo.u32 %6,%7,%9,%6; mad.lo.u32 %7,%0,%9,%7;"
                : "+r"(r0), "+r"(r1), "+r"(r2), "+r"(r3), "+r"(r4), "+r"(r5), "+r"(r6), "+r"(r7) : "r"(a), "r"(b));)
        } else if (V == 3) {    // IMAD.HI
            REP8(asm volatile("mad.hi.u32 %0,%1,%9,%0; mad.hi.u32 %1,%2,%9,%1; mad.hi.u32 %2,%3,%9,%2; mad.hi.u32 %3,%4,%9,%3;"
                              "mad.hi.u32 %4,%5,%9,%4; mad.hi.u32 %5,%6,%9,%5; mad.hi.u32 %6,%7,%9,%6; mad.hi.u32 %7,%0,%9,%7;"
                : "+r"(r0), "+r"(r1), "+r"(r2), "+r"(r3), "+r"(r4), "+r"(r5), "+r"(r6), "+r"(r7) : "r"(a), "r"(b));)
        } else if (V == 4) {    // IADD3 independent
            REP8(asm volatile("add.u32 %0,%0,%1; add.u32 %1,%1,%2; add.u32 %2,%2,%3; add.u32 %3,%3,%4;"
                              "add.u32 %4,%4,%5; add.u32 %5,%5,%6; add.u32 %6,%6,%7; add.u32 %7,%7,%0;"
                : "+r"(r0), "+r"(r1), "+r"(r2), "+r"(r3), "+r"(r4), "+r"(r5), "+r"(r6), "+r"(r7) : "r"(a), "r"(b));)
        } else if (V == 5) {    // IADD3.X carry chain, 2 interleaved chains of 8
            REP8(asm volatile(
                "add.cc.u32 %0,%0,%16; addc.cc.u32 %1,%1,%17; addc.cc.u32 %2,%2,%16; addc.cc.u32 %3,%3,%17;"
                "addc.cc.u32 %4,%4,%16; addc.cc.u32 %5,%5,%17; addc.cc.u32 %6,%6,%16; addc.u32 %7,%7,%17;"
                "add.cc.u32 %8,%8,%16; addc.cc.u32 %9,%9,%17; addc.cc.u32 %10,%10,%16; addc.cc.u32 %11,%11,%17;"
                "addc.cc.u32 %12,%12,%16; addc.cc.u32 %13,%13,%17; addc.cc.u32 %14,%14,%16; addc.u32 %15,%15,%17;"
                : "+r"(r0), "+r"(r1), "+r"(r2), "+r"(r3), "+r"(r4), "+r"(r5), "+r"(r6), "+r"(r7),
                  "+r"(r8), "+r"(r9), "+r"(r10), "+r"(r11), "+r"(r12), "+r"(r13), "+r"(r14), "+r"(r15) : "r"(a), "r"(b));)
        } else if (V == 6) {    // funnel shift
            REP8(asm volatile("shf.l.wrap.b32 %0,%0,%1,%8; shf.l.wrap.b32 %1,%1,%2,%8; shf.l.wrap.b32 %2,%2,%3,%8; shf.l.wrap.b32 %3,%3,%4,%8;"
                              "shf.l.wrap.b32 %4,%4,%5,%8; shf.l.wrap.b32 %5,%5,%6,%8; shf.l.wrap.b32 %6,%6,%7,%8; shf.l.wrap.b32 %7,%7,%0,%8;"
                : "+r"(r0), "+r"(r1), "+r"(r2), "+r"(r3), "+r"(r4), "+r"(r5), "+r"(r6), "+r"(r7) : "r"(a & 7), "r"(b));)
        } else if (V == 7) {    // LOP3 xor
            REP8(asm volatile("and.b32 %0,%0,%1; or.b32 %1,%1,%2; xor.b32 %2,%2,%3; and.b32 %3,%3,%4;"
                              "or.b32 %4,%4,%5; xor.b32 %5,%5,%6; and.b32 %6,%6,%7; or.b32 %7,%7,%0;"
                : "+r"(r0), "+r"(r1), "+r"(r2), "+r"(r3), "+r"(r4), "+r"(r5), "+r"(r6), "+r"(r7) : "r"(a), "r"(b));)
        } else if (V == 8) {    // alternate IMAD.WIDE (fma pipe) and IADD3 (alu pipe): do they overlap?
            REP8(asm volatile(
                "{ .reg .b64 t; mov.b64 t, {%0,%1}; mad.wide.u32 t, %2, %12, t; mov.b64 {%0,%1}, t; } add.u32 %8,%8,%9;"
                "{ .reg .b64 t; mov.b64 t, {%2,%3}; mad.wide.u32 t, %4, %12, t; mov.b64 {%2,%3}, t; } add.u32 %9,%9,%10;"
                "{ .reg .b64 t; mov.b64 t, {%4,%5}; mad.wide.u32 t, %6, %12, t; mov.b64 {%4,%5}, t; } add.u32 %10,%10,%11;"
                "{ .reg .b64 t; mov.b64 t, {%6,%7}; mad.wide.u32 t, %0, %12, t; mov.b64 {%6,%7}, t; } add.u32 %11,%11,%8;"
                : "+r"(r0), "+r"(r1), "+r"(r2), "+r"(r3), "+r"(r4), "+r"(r5), "+r"(r6), "+r"(r7),
                  "+r"(r8), "+r"(r9), "+r"(r10), "+r"(r11) : "r"(b));)
        } else if (V == 9) {    // carry chains + ALU work interleaved (4 IMAD.WIDE.X + 4 shf)
            REP8(asm volatile(
                "mad.lo.cc.u32 %0,%12,%13,%0; madc.hi.cc.u32 %1,%12,%13,%1; shf.l.wrap.b32 %8,%8,%9,%14;"
                "madc.lo.cc.u32 %2,%12,%13,%2; madc.hi.cc.u32 %3,%12,%13,%3; shf.l.wrap.b32 %9,%9,%10,%14;"
                "madc.lo.cc.u32 %4,%12,%13,%4; madc.hi.cc.u32 %5,%12,%13,%5; shf.l.wrap.b32 %10,%10,%11,%14;"
                "madc.lo.cc.u32 %6,%12,%13,%6; madc.hi.u32 %7,%12,%13,%7; shf.l.wrap.b32 %11,%11,%8,%14;"
                : "+r"(r0), "+r"(r1), "+r"(r2), "+r"(r3), "+r"(r4), "+r"(r5), "+r"(r6), "+r"(r7),
                  "+r"(r8), "+r"(r9), "+r"(r10), "+r"(r11) : "r"(a), "r"(b), "r"(a & 7));)
        } else if (V == 10) {   // IMAD.WIDE, both multiplier operands from other accumulators (register bank pressure)
            REP8(asm volatile(
                "{ .reg .b64 t; mov.b64 t, {%0,%1}; mad.wide.u32 t, %2, %7, t; mov.b64 {%0,%1}, t; }"
                "{ .reg .b64 t; mov.b64 t, {%2,%3}; mad.wide.u32 t, %4, %9, t; mov.b64 {%2,%3}, t; }"
                "{ .reg .b64 t; mov.b64 t, {%4,%5}; mad.wide.u32 t, %6, %11, t; mov.b64 {%4,%5}, t; }"
                "{ .reg .b64 t; mov.b64 t, {%6,%7}; mad.wide.u32 t, %8, %13, t; mov.b64 {%6,%7}, t; }"
                "{ .reg .b64 t; mov.b64 t, {%8,%9}; mad.wide.u32 t, %10, %15, t; mov.b64 {%8,%9}, t; }"
                "{ .reg .b64 t; mov.b64 t, {%10,%11}; mad.wide.u32 t, %12, %1, t; mov.b64 {%10,%11}, t; }"
                "{ .reg .b64 t; mov.b64 t, {%12,%13}; mad.wide.u32 t, %14, %3, t; mov.b64 {%12,%13}, t; }"
                "{ .reg .b64 t; mov.b64 t, {%14,%15}; mad.wide.u32 t, %0, %5, t; mov.b64 {%14,%15}, t; }"
                : "+r"(r0), "+r"(r1), "+r"(r2), "+r"(r3), "+r"(r4), "+r"(r5), "+r"(r6), "+r"(r7),
                  "+r"(r8), "+r"(r9), "+r"(r10), "+r"(r11), "+r"(r12), "+r"(r13), "+r"(r14), "+r"(r15) : "r"(b));)
        }
    }
    long long t1 = clock64();
    unsigned long long x = q0 ^ q1 ^ q2 ^ q3 ^ q4 ^ q5 ^ q6 ^ q7 ^ r0 ^ r1 ^ r2 ^ r3 ^ r4 ^ r5 ^ r6 ^ r7 ^ r8 ^ r9 ^ r10 ^ r11 ^ r12 ^ r13 ^ r14 ^ r15;
    if (x == 0x1234567ull) out[0] = x;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int V>
static void run(const char* name, int per_iter, int warps_per_sm_blocks)
{
    int dev = 0; cudaDeviceProp prop; cudaGetDeviceProperties(&prop, dev);
    const int blocks = prop.multiProcessorCount * warps_per_sm_blocks, threads = 256;
    unsigned long long* d_out; long long* d_clk;
    cudaMalloc(&d_out, 8); cudaMalloc(&d_clk, blocks * 8);
    const int iters = 4000;
    probe<V><<<blocks, threads>>>(1, 100, d_out, d_clk);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    probe<V><<<blocks, threads>>>(2, iters, d_out, d_clk);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long* h = new long long[blocks];
    cudaMemcpy(h, d_clk, blocks * 8, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < blocks; ++i) avg += h[i]; avg /= blocks;
    // warp-instructions per SM = blocks_per_sm * 8 warps * iters * per_iter
    const double winst = (double)warps_per_sm_blocks * 8 * iters * per_iter;
    printf("%-44s blocks/SM=%d  %.3f warp-inst/clk/SM  (%.1f thread-ops/clk/SM)  %.2f Tops/s  [%s]\n",
           name, warps_per_sm_blocks, winst / avg, 32.0 * winst / avg,
           (double)blocks * threads * iters * per_iter / (ms * 1e-3) / 1e12, cudaGetErrorString(cudaGetLastError()));
    cudaFree(d_out); cudaFree(d_clk); delete[] h;
}

int main()
{
    for (int b : {1, 2, 8}) {
        printf("---- %d blocks of 256 threads per SM ----\n", b);
        if (b == 1) { run<0>("IMAD.WIDE.U32 independent", 64, 1); run<1>("IMAD.WIDE.U32.X carry chains (2x4)", 64, 1); run<2>("IMAD lo", 64, 1);
                      run<3>("IMAD.HI", 64, 1); run<4>("IADD3", 64, 1); run<5>("IADD3.X chains (2x8)", 128, 1); run<6>("SHF funnel", 64, 1);
                      run<7>("LOP3", 64, 1); run<8>("IMAD.WIDE + IADD3 alternating", 64, 1); run<9>("IMAD.WIDE.X chain + SHF", 64, 1);
                      run<10>("IMAD.WIDE.U32 varied operands", 64, 1); }
        if (b == 2) { run<0>("IMAD.WIDE.U32 independent", 64, 2); run<1>("IMAD.WIDE.U32.X carry chains (2x4)", 64, 2); run<2>("IMAD lo", 64, 2);
                      run<3>("IMAD.HI", 64, 2); run<4>("IADD3", 64, 2); run<5>("IADD3.X chains (2x8)", 128, 2); run<6>("SHF funnel", 64, 2);
                      run<7>("LOP3", 64, 2); run<8>("IMAD.WIDE + IADD3 alternating", 64, 2); run<9>("IMAD.WIDE.X chain + SHF", 64, 2);
                      run<10>("IMAD.WIDE.U32 varied operands", 64, 2); }
        if (b == 8) { run<0>("IMAD.WIDE.U32 independent", 64, 8); run<1>("IMAD.WIDE.U32.X carry chains (2x4)", 64, 8); run<2>("IMAD lo", 64, 8);
                      run<3>("IMAD.HI", 64, 8); run<4>("IADD3", 64, 8); run<5>("IADD3.X chains (2x8)", 128, 8); run<6>("SHF funnel", 64, 8);
                      run<7>("LOP3", 64, 8); run<8>("IMAD.WIDE + IADD3 alternating", 64, 8); run<9>("IMAD.WIDE.X chain + SHF", 64, 8);
                      run<10>("IMAD.WIDE.U32 varied operands", 64, 8); }
    }
    return 0;
}

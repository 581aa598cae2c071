// pipes.cu -- issue-rate microbenchmark of a LONE warp on one SM sub-partition (sm_100a):
// how many cycles per warp instruction for the integer ALU pipe, the FMA pipe and their mixes,
// with 32 and with 16 active lanes.  Feeds the cost model in DESIGN.md (decoder, latency regime).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o pipes pipes.cu && ./pipes
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int kIters = 2048;

#define CHAINS8(OP)                                                                                     \
    OP(a0) OP(a1) OP(a2) OP(a3) OP(a4) OP(a5) OP(a6) OP(a7)

template <int kTest>
__global__ void __launch_bounds__(32) probe(uint32_t *out, uint64_t *cycles, uint32_t seed, int active)
{
    if ((int)threadIdx.x >= active) return;
    uint32_t a0 = seed + threadIdx.x, a1 = a0 * 3u, a2 = a0 * 5u, a3 = a0 * 7u, a4 = a0 * 11u, a5 = a0 * 13u,
             a6 = a0 * 17u, a7 = a0 * 19u;
    uint32_t b = seed | 1u, c = seed * 77u + 5u;
    __shared__ uint32_t sm[32];
    sm[threadIdx.x] = (threadIdx.x * 7u + seed) & 31u;
    __syncwarp();
    const uint64_t t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < kIters; ++i) {
        if (kTest == 0) {           // LOP3, 8 independent chains
#define OP(x) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(b), "r"(c));
            CHAINS8(OP) CHAINS8(OP)
#undef OP
        } else if (kTest == 1) {    // IMAD (three register operands)
#define OP(x) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(c));
            CHAINS8(OP) CHAINS8(OP)
#undef OP
        } else if (kTest == 2) {    // alternating LOP3 / IMAD
#define OP(x) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(b), "r"(c));
#define OQ(x) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(c));
            OP(a0) OQ(a1) OP(a2) OQ(a3) OP(a4) OQ(a5) OP(a6) OQ(a7) OP(a0) OQ(a1) OP(a2) OQ(a3) OP(a4) OQ(a5) OP(a6) OQ(a7)
#undef OP
#undef OQ
        } else if (kTest == 3) {    // SHF (funnel shift, variable amount)
#define OP(x) asm volatile("shf.l.wrap.b32 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(c));
            CHAINS8(OP) CHAINS8(OP)
#undef OP
        } else if (kTest == 4) {    // PRMT
#define OP(x) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(c));
            CHAINS8(OP) CHAINS8(OP)
#undef OP
        } else if (kTest == 5) {    // dependent LOP3 chain (latency)
#define OP(x) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a0) : "r"(b), "r"(c));
            CHAINS8(OP) CHAINS8(OP)
#undef OP
        } else if (kTest == 6) {    // dependent IMAD chain (latency)
#define OP(x) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a0) : "r"(b), "r"(c));
            CHAINS8(OP) CHAINS8(OP)
#undef OP
        } else if (kTest == 7) {    // IMAD.HI
#define OP(x) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(x) : "r"(b));
            CHAINS8(OP) CHAINS8(OP)
#undef OP
        } else if (kTest == 8) {    // FFMA, register form
#define OP(x) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(*(float *)&x) : "f"(*(float *)&b), "f"(*(float *)&c));
            CHAINS8(OP) CHAINS8(OP)
#undef OP
        } else if (kTest == 9) {    // 3 LOP3 : 1 IMAD (the decoder's mix)
#define OP(x) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(b), "r"(c));
#define OQ(x) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(c));
            OP(a0) OP(a1) OQ(a2) OP(a3) OP(a4) OP(a5) OQ(a6) OP(a7) OP(a0) OP(a1) OQ(a2) OP(a3) OP(a4) OP(a5) OQ(a6) OP(a7)
#undef OP
#undef OQ
        } else if (kTest == 10) {   // IADD3 with three register operands
#define OP(x) asm volatile("{.reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2;}" : "+r"(x) : "r"(b), "r"(c));
            CHAINS8(OP) CHAINS8(OP)
#undef OP
        } else if (kTest == 11) {   // dependent chain alternating LOP3 -> IMAD (cross-pipe latency)
#define OP(x) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a0) : "r"(b), "r"(c));
#define OQ(x) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a0) : "r"(b), "r"(c));
            OP(a0) OQ(a0) OP(a0) OQ(a0) OP(a0) OQ(a0) OP(a0) OQ(a0) OP(a0) OQ(a0) OP(a0) OQ(a0) OP(a0) OQ(a0) OP(a0) OQ(a0)
#undef OP
#undef OQ
        } else if (kTest == 12) {   // SEL on a predicate + ISETP pairs
#define OP(x) asm volatile("{.reg .pred p; setp.lt.u32 p, %0, %1; selp.u32 %0, %2, %0, p;}" : "+r"(x) : "r"(b), "r"(c));
            CHAINS8(OP)
#undef OP
        } else if (kTest == 13) {   // 2 chains only of LOP3 (ILP 2)
#define OP(x) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(b), "r"(c));
            OP(a0) OP(a1) OP(a0) OP(a1) OP(a0) OP(a1) OP(a0) OP(a1) OP(a0) OP(a1) OP(a0) OP(a1) OP(a0) OP(a1) OP(a0) OP(a1)
#undef OP
        } else if (kTest == 14) {   // IMAD with an immediate multiplier (shift-like)
#define OP(x) asm volatile("mad.lo.u32 %0, %0, 65537, %1;" : "+r"(x) : "r"(c));
            CHAINS8(OP) CHAINS8(OP)
#undef OP
        } else if (kTest == 20) {   // dependent chain through a predicate: ISETP -> SEL
#define OP(x) asm volatile("{.reg .pred p; setp.lt.s32 p, %0, %1; selp.u32 %0, %2, %1, p; xor.b32 %0, %0, %1;}" : "+r"(a0) : "r"(b), "r"(c));
            CHAINS8(OP)
#undef OP
        } else if (kTest == 21) {   // the same decision on a sign mask: SHF -> LOP3
#define OP(x) asm volatile("{.reg .u32 t; shr.s32 t, %0, 31; lop3.b32 %0, %2, t, %1, 0xd8; xor.b32 %0, %0, %1;}" : "+r"(a0) : "r"(b), "r"(c));
            CHAINS8(OP)
#undef OP
        } else if (kTest == 22) {   // one tree level on predicates: IMAD -> ISETP -> SEL -> SEL
#define OP(x) asm volatile("{.reg .pred p, q; .reg .u32 t, u; mad.lo.u32 t, %0, %1, %2; setp.lt.s32 p, t, 0; add.u32 u, t, %1; setp.lt.s32 q, u, 0; selp.u32 t, %1, %2, p; @q selp.u32 t, %2, %0, p; mov.u32 %0, t;}" : "+r"(a0) : "r"(b), "r"(c));
            CHAINS8(OP)
#undef OP
        } else if (kTest == 23) {   // one tree level on masks: IMAD -> SHF -> LOP3 -> LOP3
#define OP(x) asm volatile("{.reg .u32 t, u, v; mad.lo.u32 t, %0, %1, %2; shr.s32 u, t, 31; lop3.b32 v, %1, u, %2, 0xd8; lop3.b32 %0, v, u, %0, 0xd8;}" : "+r"(a0) : "r"(b), "r"(c));
            CHAINS8(OP)
#undef OP
        } else if (kTest == 24) {   // quotient chain: I2F -> FMUL -> F2I
#define OP(x) asm volatile("{.reg .f32 f; cvt.rz.f32.u32 f, %0; mul.f32 f, f, %1; cvt.rzi.u32.f32 %0, f;}" : "+r"(a0) : "f"(1.0001f));
            CHAINS8(OP)
#undef OP
        } else if (kTest == 25) {   // dependent IMAD.HI
#define OP(x) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(a0) : "r"(0xFFFFFFF0u));
            CHAINS8(OP) CHAINS8(OP)
#undef OP
        } else if (kTest == 26) {   // dependent FADD
#define OP(x) asm volatile("add.f32 %0, %0, %1;" : "+f"(*(float *)&a0) : "f"(1.5f));
            CHAINS8(OP) CHAINS8(OP)
#undef OP
        } else if (kTest == 27) {   // dependent shared-memory load (pointer chase inside the lane's own word)
            a0 = ((volatile uint32_t *)sm)[a0 & 31u]; a0 = ((volatile uint32_t *)sm)[a0 & 31u];
            a0 = ((volatile uint32_t *)sm)[a0 & 31u]; a0 = ((volatile uint32_t *)sm)[a0 & 31u];
            a0 = ((volatile uint32_t *)sm)[a0 & 31u]; a0 = ((volatile uint32_t *)sm)[a0 & 31u];
            a0 = ((volatile uint32_t *)sm)[a0 & 31u]; a0 = ((volatile uint32_t *)sm)[a0 & 31u];
        } else if (kTest == 28) {   // dependent SHFL
#define OP(x) a0 = __shfl_sync(0xFFFFFFFFu, a0, (a0 + 1u) & 15u);
            CHAINS8(OP)
#undef OP
        } else if (kTest == 29) {   // dependent funnel shift with a data-dependent amount
#define OP(x) asm volatile("shf.l.wrap.b32 %0, %0, %1, %0;" : "+r"(a0) : "r"(b));
            CHAINS8(OP) CHAINS8(OP)
#undef OP
        } else if (kTest == 30) {   // predicate produced by LOP3 (and + setp.ne fused) -> SEL
#define OP(x) asm volatile("{.reg .pred p; .reg .u32 t; and.b32 t, %0, 0x8000; setp.ne.u32 p, t, 0; selp.u32 %0, %2, %1, p; xor.b32 %0, %0, %1;}" : "+r"(a0) : "r"(b), "r"(c));
            CHAINS8(OP)
#undef OP
        } else if (kTest == 15) {   // add that ptxas may turn into IMAD.IADD / IADD3
#define OP(x) asm volatile("add.u32 %0, %0, %1;" : "+r"(x) : "r"(b));
            CHAINS8(OP) CHAINS8(OP)
#undef OP
        }
    }
    const uint64_t t1 = clock64();
    out[threadIdx.x] = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;
    if (threadIdx.x == 0) *cycles = t1 - t0;
}

template <int kTest>
void run(const char *name, int per_iter, uint32_t *d_out, uint64_t *d_cyc)
{
    for (int active : {32, 16}) {
        for (int warps_per_sm : {1, 8}) {
            // warps_per_sm = 8: grid of 8 x 148 one-warp CTAs, two warps per scheduler
            const int grid = warps_per_sm == 1 ? 1 : 8 * 148;
            probe<kTest><<<grid, 32>>>(d_out, d_cyc, 12345u, active);
            probe<kTest><<<grid, 32>>>(d_out, d_cyc, 12345u, active);
            cudaDeviceSynchronize();
            uint64_t cyc = 0;
            cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost);
            printf("%-34s lanes %2d  warps/SM %d : %6.3f cycles per warp instruction\n", name, active, warps_per_sm,
                   (double)cyc / ((double)kIters * per_iter));
        }
    }
}

int main()
{
    uint32_t *d_out;
    uint64_t *d_cyc;
    cudaMalloc(&d_out, 4096);
    cudaMalloc(&d_cyc, 8);
    run<0>("LOP3 x8 chains", 16, d_out, d_cyc);
    run<13>("LOP3 x2 chains", 16, d_out, d_cyc);
    run<5>("LOP3 dependent", 16, d_out, d_cyc);
    run<1>("IMAD rrr x8 chains", 16, d_out, d_cyc);
    run<14>("IMAD imm x8 chains", 16, d_out, d_cyc);
    run<6>("IMAD dependent", 16, d_out, d_cyc);
    run<7>("IMAD.HI x8 chains", 16, d_out, d_cyc);
    run<2>("LOP3/IMAD alternating", 16, d_out, d_cyc);
    run<9>("3 LOP3 : 1 IMAD", 16, d_out, d_cyc);
    run<11>("LOP3->IMAD dependent", 16, d_out, d_cyc);
    run<3>("SHF x8 chains", 16, d_out, d_cyc);
    run<4>("PRMT x8 chains", 16, d_out, d_cyc);
    run<10>("IADD3 rrr x8 chains", 16, d_out, d_cyc);
    run<15>("add rr x8 chains", 16, d_out, d_cyc);
    run<12>("ISETP+SEL x8 chains", 16, d_out, d_cyc);
    run<8>("FFMA rrr x8 chains", 16, d_out, d_cyc);
    printf("--- dependent chains: cycles per ELEMENT of the chain (loop overhead ~10 cycles per 8 elements included)\n");
    run<20>("ISETP -> SEL -> LOP3", 8, d_out, d_cyc);
    run<30>("LOP3.P -> SEL -> LOP3", 8, d_out, d_cyc);
    run<21>("SHF -> LOP3 -> LOP3", 8, d_out, d_cyc);
    run<22>("IMAD -> ISETP -> SEL -> @SEL", 8, d_out, d_cyc);
    run<23>("IMAD -> SHF -> LOP3 -> LOP3", 8, d_out, d_cyc);
    run<24>("I2F -> FMUL -> F2I", 8, d_out, d_cyc);
    run<25>("IMAD.HI dependent", 16, d_out, d_cyc);
    run<26>("FADD dependent", 16, d_out, d_cyc);
    run<27>("LDS dependent", 8, d_out, d_cyc);
    run<28>("SHFL dependent", 8, d_out, d_cyc);
    run<29>("SHF dependent", 16, d_out, d_cyc);
    return 0;
}

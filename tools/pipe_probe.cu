// pipe_probe.cu -- which issue pipe does an instruction share?  (B200, sm_100a)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/pipe_probe tools/pipe_probe.cu && /tmp/pipe_probe
// Every warp runs a loop of 8 independent chains of instruction X, of instruction Y, and of X and Y interleaved (16 chains).
// If X and Y issue on different pipes the interleaved loop takes about max(tX, tY); if they share one, about tX + tY.
// Used to decide which integer work of the step kernel can leave the ALU pipe (board_ops.cuh: add_on_fma_pipe and friends).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

enum Op { LOP3, IMAD, PRMT, VABSDIFF4, FADD, SHF, VIMNMX3, POPC, I2FP };

template <int kOp>
__device__ __forceinline__ uint32_t apply(uint32_t x, uint32_t y, uint32_t z)
{
    uint32_t d;
    if (kOp == LOP3) asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(d) : "r"(x), "r"(y), "r"(z));
    else if (kOp == IMAD) asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(x), "r"(y), "r"(z));
    else if (kOp == PRMT) asm volatile("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(x), "r"(y), "r"(z));
    else if (kOp == VABSDIFF4) asm volatile("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(d) : "r"(x), "r"(y), "r"(z));
    else if (kOp == FADD) { float f; asm volatile("add.f32 %0, %1, %2;" : "=f"(f) : "f"(__uint_as_float(x)), "f"(__uint_as_float(y))); d = __float_as_uint(f); }
    else if (kOp == SHF) asm volatile("shf.l.wrap.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(x), "r"(y), "r"(z));
    else if (kOp == VIMNMX3) d = __vimin3_u16x2(x, y, z);
    else if (kOp == POPC) { asm volatile("popc.b32 %0, %1;" : "=r"(d) : "r"(x)); d += y; }
    else { float f; asm volatile("cvt.rn.f32.u32 %0, %1;" : "=f"(f) : "r"(x)); d = __float_as_uint(f); }
    return d;
}

template <int kX, int kY, bool kUseX, bool kUseY>
__global__ void __launch_bounds__(256) probe(uint32_t *out, long long *cycles, int iters, uint32_t seed)
{
    uint32_t a[8], b[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = seed + threadIdx.x * 8 + i, b[i] = seed * 3 + threadIdx.x + i;
    const uint32_t y = seed | 1u, z = seed + 5u;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (kUseX) a[i] = apply<kX>(a[i], y, z);
            if (kUseY) b[i] = apply<kY>(b[i], y, z);
        }
    }
    const long long t1 = clock64();
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) r ^= a[i] ^ b[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int kX, int kY>
void run(const char *nx, const char *ny, uint32_t *out, long long *cyc)
{
    const int iters = 4096, blocks = 148 * 8;  // 8 blocks x 8 warps per SM: 16 warps per scheduler
    long long h[3] = {0, 0, 0};
    for (int v = 0; v < 3; ++v) {
        if (v == 0) probe<kX, kY, true, false><<<blocks, 256>>>(out, cyc, iters, 12345u);
        if (v == 1) probe<kX, kY, false, true><<<blocks, 256>>>(out, cyc, iters, 12345u);
        if (v == 2) probe<kX, kY, true, true><<<blocks, 256>>>(out, cyc, iters, 12345u);
        cudaDeviceSynchronize();
        cudaMemcpy(&h[v], cyc, sizeof(long long), cudaMemcpyDeviceToHost);
    }
    const long long alone = h[0] > h[1] ? h[0] : h[1];
    printf("%-10s alone %8lld cycles   %-10s alone %8lld   interleaved %8lld = %.2f x the slower one alone -> %s\n", nx, h[0], ny, h[1], h[2],
           (double)h[2] / alone, (double)h[2] / alone < 1.25 ? "different pipes" : "they share a pipe (or contend)");
}

int main()
{
    uint32_t *out;
    long long *cyc;
    cudaMalloc(&out, 148 * 8 * 256 * 4);
    cudaMalloc(&cyc, 148 * 8 * 8);
    run<LOP3, IMAD>("LOP3", "IMAD", out, cyc);
    run<LOP3, PRMT>("LOP3", "PRMT", out, cyc);
    run<LOP3, VABSDIFF4>("LOP3", "VABSDIFF4", out, cyc);
    run<IMAD, VABSDIFF4>("IMAD", "VABSDIFF4", out, cyc);
    run<LOP3, FADD>("LOP3", "FADD", out, cyc);
    run<IMAD, FADD>("IMAD", "FADD", out, cyc);
    run<LOP3, SHF>("LOP3", "SHF", out, cyc);
    run<LOP3, VIMNMX3>("LOP3", "VIMNMX3", out, cyc);
    run<LOP3, POPC>("LOP3", "POPC", out, cyc);
    run<LOP3, I2FP>("LOP3", "I2FP", out, cyc);
    run<IMAD, I2FP>("IMAD", "I2FP", out, cyc);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

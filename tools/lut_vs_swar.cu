// lut_vs_swar.cu -- the measurement BASELINE.json's north_star leaves open: "a shared-memory row-transition LUT or a
// register-resident row merge (whichever ncu shows is faster)".
//
// Two kernels with identical inputs and outputs -- board (16 bytes) + action (1 byte) in, moved board (16 bytes) + the
// reward_fn_normal gain (4 bytes) out, one thread per game, M = 2^24 games -- that differ only in how the four lines
// of a move are pushed:
//   move_swar : the product's lane-parallel byte-SWAR push (board_ops.cuh: move_board_sel + fusion_gain), 256-thread blocks
//   move_lut  : the classic 2048 bitboard method: every line packed to 16 bits (four 4-bit cells) indexes a 65536-entry
//               table of pushed lines (u16, 128 KiB) and one of the consumed exponents (u8, 64 KiB), both in shared
//               memory (192 KiB => one persistent 1024-thread block per SM); the direction is handled by the same kind
//               of selector-driven byte-permute network as in the SWAR kernel, so the comparison is push vs push.
// 4-bit cells cannot hold the exponents 16 and 17 the reference allows (game_numba.py:23-45), so a product LUT kernel
// would additionally need a SWAR fallback for such boards; here the boards are kept below 16.
// The program checks that both kernels agree on every game, then times them with CUDA events (and runs under ncu for
// the pipe utilisation).  Build + run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -o /tmp/lut_vs_swar tools/lut_vs_swar.cu && /tmp/lut_vs_swar
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "../ml2048_b200/csrc/board_ops.cuh"

using namespace ml2048;

#define CK(x)                                                                                 \
    do {                                                                                      \
        cudaError_t e_ = (x);                                                                 \
        if (e_ != cudaSuccess) {                                                              \
            fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); \
            exit(1);                                                                          \
        }                                                                                     \
    } while (0)

__device__ const uint32_t d_move_sel[4 * kMoveSelRow] = ML2048_MOVE_SEL_TABLE;

// lines as WORDS (byte 0 = the cell next to the wall) and back: {in1a, in1b, in2a, in2b, out1a, out1b, out2a, out2b}
//   t0 = P(r0,r2,in1a) t1 = P(r1,r3,in1a) t2 = P(r0,r2,in1b) t3 = P(r1,r3,in1b)
//   L0 = P(t0,t1,in2a) L1 = P(t0,t1,in2b) L2 = P(t2,t3,in2a) L3 = P(t2,t3,in2b)       and the same wiring back
__device__ const uint32_t d_line_sel[4 * 8] = {
    0x3210u, 0x7654u, 0x3210u, 0x7654u, 0x3210u, 0x7654u, 0x3210u, 0x7654u,  // left : lines are the rows
    0x0123u, 0x4567u, 0x3210u, 0x7654u, 0x0123u, 0x4567u, 0x3210u, 0x7654u,  // right: rows, bytes reversed
    0x5140u, 0x7362u, 0x5140u, 0x7362u, 0x5140u, 0x7362u, 0x5140u, 0x7362u,  // up   : columns (transpose)
    0x1504u, 0x3726u, 0x1504u, 0x3726u, 0x6273u, 0x4051u, 0x5140u, 0x7362u,  // down : columns, bottom first
};

__global__ void __launch_bounds__(256) move_swar(const uint4 *boards, const uint8_t *actions, uint4 *out, uint32_t *gain, int64_t n)
{
    const int64_t g = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (g >= n) return;
    const uint4 b = boards[g];
    uint32_t r0 = b.x, r1 = b.y, r2 = b.z, r3 = b.w;
    Fusions f;
    move_board_sel(r0, r1, r2, r3, d_move_sel + (actions[g] & 3u) * kMoveSelRow, f);
    out[g] = make_uint4(r0, r1, r2, r3);
    gain[g] = (uint32_t)fusion_gain(f);
}

__device__ __forceinline__ uint32_t pack_line(uint32_t w)  // four bytes < 16 -> four nibbles
{
    const uint32_t x = w | (w >> 4);
    return prmt(x, 0u, 0x4420);
}

__device__ __forceinline__ uint32_t unpack_line(uint32_t v)  // four nibbles -> four bytes
{
    const uint32_t z = prmt(v, 0u, 0x1100);  // (B0, B0, B1, B1)
    return ((z & 0x000f000fu) | ((z >> 4) & 0x0f000f00u));
}

__global__ void __launch_bounds__(1024, 1) move_lut(const uint4 *boards, const uint8_t *actions, uint4 *out, uint32_t *gain, int64_t n,
                                                    const uint16_t *g_rows, const uint8_t *g_fuse)
{
    extern __shared__ __align__(16) uint8_t smem[];
    uint16_t *s_rows = reinterpret_cast<uint16_t *>(smem);  // 128 KiB
    uint8_t *s_fuse = smem + 65536 * 2;                     // 64 KiB
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(g_rows);
        uint4 *dst = reinterpret_cast<uint4 *>(s_rows);
        for (int i = threadIdx.x; i < 65536 * 2 / 16; i += 1024) dst[i] = src[i];
        src = reinterpret_cast<const uint4 *>(g_fuse);
        dst = reinterpret_cast<uint4 *>(s_fuse);
        for (int i = threadIdx.x; i < 65536 / 16; i += 1024) dst[i] = src[i];
    }
    __syncthreads();
    for (int64_t g = (int64_t)blockIdx.x * 1024 + threadIdx.x; g < n; g += (int64_t)gridDim.x * 1024) {
        const uint4 b = boards[g];
        const uint32_t *sel = d_line_sel + (actions[g] & 3u) * 8;
        const uint4 sa = __ldg(reinterpret_cast<const uint4 *>(sel));
        const uint4 sb = __ldg(reinterpret_cast<const uint4 *>(sel + 4));
        const uint32_t t0 = prmt_sign(b.x, b.z, sa.x), t1 = prmt_sign(b.y, b.w, sa.x);
        const uint32_t t2 = prmt_sign(b.x, b.z, sa.y), t3 = prmt_sign(b.y, b.w, sa.y);
        const uint32_t i0 = pack_line(prmt_sign(t0, t1, sa.z)), i1 = pack_line(prmt_sign(t0, t1, sa.w));
        const uint32_t i2 = pack_line(prmt_sign(t2, t3, sa.z)), i3 = pack_line(prmt_sign(t2, t3, sa.w));
        const uint32_t L0 = unpack_line(s_rows[i0]), L1 = unpack_line(s_rows[i1]);
        const uint32_t L2 = unpack_line(s_rows[i2]), L3 = unpack_line(s_rows[i3]);
        // consumed exponents, two nibbles per line, as one word: byte i = line i
        const uint32_t fz = (uint32_t)s_fuse[i0] | ((uint32_t)s_fuse[i1] << 8) | ((uint32_t)s_fuse[i2] << 16) | ((uint32_t)s_fuse[i3] << 24);
        Fusions f;
        // Fusions holds exponent + 127 per fused candidate (board_ops.cuh)
        const uint32_t e1 = fz & 0x0f0f0f0fu, e2 = (fz >> 4) & 0x0f0f0f0fu;
        f.first = (e1 + kLo7) & nonzero_mask(e1);
        f.second = (e2 + kLo7) & nonzero_mask(e2);
        const uint32_t u0 = prmt_sign(L0, L2, sb.x), u1 = prmt_sign(L1, L3, sb.x);
        const uint32_t u2 = prmt_sign(L0, L2, sb.y), u3 = prmt_sign(L1, L3, sb.y);
        out[g] = make_uint4(prmt_sign(u0, u1, sb.z), prmt_sign(u0, u1, sb.w), prmt_sign(u2, u3, sb.z), prmt_sign(u2, u3, sb.w));
        gain[g] = (uint32_t)fusion_gain(f);
    }
}

__global__ void fill_inputs(uint4 *boards, uint8_t *actions, int64_t n)
{
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    const u32x4 a = philox4x32_10((uint32_t)g, (uint32_t)(g >> 32), 1u, 0u, 0x2048u, 0u);
    const u32x4 c = philox4x32_10((uint32_t)g, (uint32_t)(g >> 32), 2u, 0u, 0x2048u, 0u);
    uint32_t w[4] = {a.x, a.y, a.z, a.w}, e[4] = {c.x, c.y, c.z, c.w}, r[4];
    for (int i = 0; i < 4; ++i) {
        uint32_t row = 0;
        for (int j = 0; j < 4; ++j) {
            const uint32_t v = (w[i] >> (8 * j)) & 0xffu;
            const uint32_t cell = ((e[i] >> (8 * j)) & 3u) == 0u ? 0u : 1u + v % 6u + ((v >> 5) == 0u ? v % 9u : 0u);  // 1..6 mostly, up to 14
            row |= (cell > 15u ? 15u : cell) << (8 * j);
        }
        r[i] = row;
    }
    boards[g] = make_uint4(r[0], r[1], r[2], r[3]);
    actions[g] = (uint8_t)(c.x >> 30);
}

// host: the 65536-entry tables, by the rules of _push_row (game_numba.py:48-90) on one line of four 4-bit cells
static void build_tables(std::vector<uint16_t> &rows, std::vector<uint8_t> &fuse)
{
    rows.resize(65536);
    fuse.resize(65536);
    for (uint32_t idx = 0; idx < 65536; ++idx) {
        int cells[4], outc[4] = {0, 0, 0, 0}, n = 0, consumed[2] = {0, 0}, nf = 0;
        for (int j = 0; j < 4; ++j) cells[j] = (idx >> (4 * j)) & 15;
        int prev = 0;
        for (int j = 0; j < 4; ++j) {
            const int v = cells[j];
            if (!v) continue;
            if (prev && prev == v) {
                outc[n - 1] = (v + 1) & 15;  // 15 + 15 does not fit four bits; the input generator stops at 14
                consumed[nf++] = v;
                prev = 0;
            } else {
                outc[n++] = v;
                prev = v;
            }
        }
        rows[idx] = (uint16_t)(outc[0] | outc[1] << 4 | outc[2] << 8 | outc[3] << 12);
        fuse[idx] = (uint8_t)(consumed[0] | consumed[1] << 4);
    }
}

int main(int argc, char **argv)
{
    const int64_t n = argc > 1 ? atoll(argv[1]) : (1ll << 24);
    const int reps = argc > 2 ? atoi(argv[2]) : 20;
    uint4 *boards, *out_a, *out_b;
    uint8_t *actions, *d_fuse;
    uint32_t *gain_a, *gain_b;
    uint16_t *d_rows;
    CK(cudaMalloc(&boards, n * 16));
    CK(cudaMalloc(&out_a, n * 16));
    CK(cudaMalloc(&out_b, n * 16));
    CK(cudaMalloc(&actions, n));
    CK(cudaMalloc(&gain_a, n * 4));
    CK(cudaMalloc(&gain_b, n * 4));
    CK(cudaMalloc(&d_rows, 65536 * 2));
    CK(cudaMalloc(&d_fuse, 65536));
    std::vector<uint16_t> rows;
    std::vector<uint8_t> fuse;
    build_tables(rows, fuse);
    CK(cudaMemcpy(d_rows, rows.data(), 65536 * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_fuse, fuse.data(), 65536, cudaMemcpyHostToDevice));
    fill_inputs<<<(unsigned)((n + 255) / 256), 256>>>(boards, actions, n);
    CK(cudaDeviceSynchronize());

    int dev = 0, sms = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int lut_smem = 65536 * 3;
    CK(cudaFuncSetAttribute(move_lut, cudaFuncAttributeMaxDynamicSharedMemorySize, lut_smem));
    const unsigned grid_swar = (unsigned)((n + 255) / 256);

    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    float ms_swar = 0, ms_lut = 0;
    for (int pass = 0; pass < 2; ++pass) {  // pass 0 warms up
        CK(cudaEventRecord(e0));
        for (int i = 0; i < reps; ++i) move_swar<<<grid_swar, 256>>>(boards, actions, out_a, gain_a, n);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms_swar, e0, e1));
        CK(cudaEventRecord(e0));
        for (int i = 0; i < reps; ++i) move_lut<<<sms, 1024, lut_smem>>>(boards, actions, out_b, gain_b, n, d_rows, d_fuse);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms_lut, e0, e1));
    }
    CK(cudaGetLastError());

    // both kernels must agree on every game
    std::vector<uint32_t> ha((size_t)n * 4), hb((size_t)n * 4), ga((size_t)n), gb((size_t)n);
    CK(cudaMemcpy(ha.data(), out_a, n * 16, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hb.data(), out_b, n * 16, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(ga.data(), gain_a, n * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(gb.data(), gain_b, n * 4, cudaMemcpyDeviceToHost));
    int64_t bad = 0, fused = 0;
    for (int64_t i = 0; i < n; ++i) {
        bad += (ha[4 * i] != hb[4 * i]) | (ha[4 * i + 1] != hb[4 * i + 1]) | (ha[4 * i + 2] != hb[4 * i + 2]) | (ha[4 * i + 3] != hb[4 * i + 3]) |
               (ga[i] != gb[i]);
        fused += ga[i] != 0;
    }
    const double bytes = 37.0 * (double)n;  // 16 + 1 read, 16 + 4 written
    printf("{\"games\": %lld, \"reps\": %d, \"mismatches\": %lld, \"moves_with_fusion\": %lld, \"swar_us\": %.1f, \"lut_us\": %.1f, "
           "\"swar_gbs\": %.0f, \"lut_gbs\": %.0f, \"lut_over_swar\": %.3f}\n",
           (long long)n, reps, (long long)bad, (long long)fused, ms_swar * 1e3 / reps, ms_lut * 1e3 / reps, bytes / (ms_swar * 1e-3 / reps) / 1e9,
           bytes / (ms_lut * 1e-3 / reps) / 1e9, ms_lut / ms_swar);
    return bad ? 2 : 0;
}

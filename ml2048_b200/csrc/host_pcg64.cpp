// host_pcg64.cpp -- the handful of numpy Generator(PCG64) draws the reference makes per prepare()/step()
// (reference: src/ml2048/game_numba.py:589-591, 622-626, 670), re-implemented on the host in C so that the replay
// ("bit-exact") mode does not pay numpy's per-call overhead: Generator.permuted over the (1024,16) table alone costs
// ~470 us per refresh in numpy, ~15 us here.
//
// These are restatements of published algorithms -- PCG64 XSL-RR 128/64 (O'Neill), Lemire's bounded integers, masked
// rejection sampling, Fisher-Yates -- arranged to consume the bit stream exactly like numpy 1.17+ / 2.x does.  Nothing
// here is trusted blindly: ml2048_b200/host_rng.py checks the output against the installed numpy at start-up and falls
// back to calling numpy if a single draw differs, and tests/test_host_rng.py compares long streams.
#include <stdint.h>

#include "../../include/ml2048_b200.h"

namespace {

typedef unsigned __int128 u128;

inline u128 make128(uint64_t hi, uint64_t lo) { return ((u128)hi << 64) | lo; }

const u128 kMult = make128(2549297995355413924ULL, 4865540595714422341ULL);  // PCG_DEFAULT_MULTIPLIER_128

inline uint64_t rotr64(uint64_t v, unsigned r) { return (v >> r) | (v << ((-r) & 63)); }

inline uint64_t next64(ml2048_pcg64 *g)
{
    u128 s = make128(g->state_hi, g->state_lo);
    s = s * kMult + make128(g->inc_hi, g->inc_lo);  // advance, then output from the NEW state (setseq_128_xsl_rr_64)
    g->state_hi = (uint64_t)(s >> 64);
    g->state_lo = (uint64_t)s;
    return rotr64(g->state_hi ^ g->state_lo, (unsigned)(g->state_hi >> 58));
}

// numpy hands out the two halves of one 64-bit draw as two 32-bit draws, low half first
inline uint32_t next32(ml2048_pcg64 *g)
{
    if (g->has_uint32) {
        g->has_uint32 = 0;
        return g->uinteger;
    }
    const uint64_t v = next64(g);
    g->has_uint32 = 1;
    g->uinteger = (uint32_t)(v >> 32);
    return (uint32_t)v;
}

// random_interval(max): uniform in [0, max] by masked rejection on 32-bit draws (max < 2^32)
inline uint32_t interval32(ml2048_pcg64 *g, uint32_t max)
{
    if (max == 0) return 0;
    uint32_t mask = max;
    mask |= mask >> 1;
    mask |= mask >> 2;
    mask |= mask >> 4;
    mask |= mask >> 8;
    mask |= mask >> 16;
    uint32_t v;
    while ((v = (next32(g) & mask)) > max) {
    }
    return v;
}

}  // namespace

extern "C" {

// Generator.random(): float64 in [0,1) from the top 53 bits of one 64-bit draw
double ml2048_pcg64_random(ml2048_pcg64 *g) { return (double)(next64(g) >> 11) * (1.0 / 9007199254740992.0); }

// Generator.integers(0, high) for 0 < high <= 2^32 - 1: Lemire's multiply-shift with rejection on 32-bit draws
int64_t ml2048_pcg64_integers(ml2048_pcg64 *g, int64_t high)
{
    if (high <= 1) return 0;
    const uint32_t rng = (uint32_t)(high - 1), excl = rng + 1u;
    uint64_t m = (uint64_t)next32(g) * excl;
    uint32_t leftover = (uint32_t)m;
    if (leftover < excl) {
        const uint32_t threshold = (0xFFFFFFFFu - rng) % excl;
        while (leftover < threshold) {
            m = (uint64_t)next32(g) * excl;
            leftover = (uint32_t)m;
        }
    }
    return (int64_t)(m >> 32);
}

// Generator.random(dtype=float32, out=...): 24 bits of one 32-bit draw each
void ml2048_pcg64_random_f32(ml2048_pcg64 *g, float *out, int64_t n)
{
    for (int64_t i = 0; i < n; ++i) out[i] = (float)(next32(g) >> 8) * (1.0f / 16777216.0f);
}

// Generator.permuted(x, axis=1, out=x) for a C-contiguous uint8 (rows, cols) array: every row is shuffled in place,
// rows in order, Fisher-Yates from the last element down, j = random_interval(i)
void ml2048_pcg64_permuted_rows_u8(ml2048_pcg64 *g, uint8_t *x, int64_t rows, int64_t cols)
{
    for (int64_t r = 0; r < rows; ++r) {
        uint8_t *row = x + r * cols;
        for (int64_t i = cols - 1; i >= 1; --i) {
            const uint32_t j = interval32(g, (uint32_t)i);
            const uint8_t t = row[j];
            row[j] = row[i];
            row[i] = t;
        }
    }
}

}  // extern "C"

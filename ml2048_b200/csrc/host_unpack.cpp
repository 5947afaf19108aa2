// host_unpack.cpp -- the host half of the one-byte-per-game result flags (ml2048_pack_flags): a caller that wants the
// reference's arrays (valid_actions bool[M,4] game_numba.py:540, terminated :546, invalid :547) on the host receives ONE byte
// per game over PCIe instead of six and expands it here, a slice at a time while the next slices are still in flight.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include <thread>
#include <vector>

#include "../../include/ml2048_b200.h"

namespace {

// bits 0..3 -> four 0/1 bytes: the multiply drops bit i at 8i (and copies elsewhere that the mask removes)
inline uint32_t spread4(uint32_t bits) { return ((bits & 15u) * 0x00204081u) & 0x01010101u; }

#if defined(__x86_64__) && defined(__GNUC__)
#define ML2048_HAVE_AVX2_PATH 1
// eight games per iteration: every packed byte replicated four times (in-lane byte shuffle), tested against the bit of its
// position, turned into 0/1 -- 32 output bytes per store
__attribute__((target("avx2"))) int64_t unpack_valid_avx2(const uint8_t *packed, int64_t lo, int64_t hi, uint8_t *valid4)
{
    const __m256i spread = _mm256_setr_epi8(0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 6, 6, 6, 6, 7, 7, 7, 7);
    const __m256i bit = _mm256_set1_epi32(0x08040201);
    const __m256i one = _mm256_set1_epi8(1);
    int64_t i = lo;
    // the expanded array is written once and read by somebody else later: non-temporal stores when the rows are 32-byte aligned
    // (no read-for-ownership traffic next to the PCIe writes that are landing in the same memory)
    const bool stream = ((reinterpret_cast<uintptr_t>(valid4) + 4 * (uintptr_t)lo) & 31u) == 0u;
    for (; i + 8 <= hi; i += 8) {
        long long eight;
        memcpy(&eight, packed + i, 8);
        const __m256i b = _mm256_shuffle_epi8(_mm256_set1_epi64x(eight), spread);
        const __m256i hit = _mm256_cmpeq_epi8(_mm256_and_si256(b, bit), bit);
        const __m256i out = _mm256_and_si256(hit, one);
        if (stream) _mm256_stream_si256(reinterpret_cast<__m256i *>(valid4 + 4 * i), out);
        else _mm256_storeu_si256(reinterpret_cast<__m256i *>(valid4 + 4 * i), out);
    }
    if (stream) _mm_sfence();
    return i;
}
#endif

// one flag bit of eight games at a time
inline void unpack_bit(const uint8_t *packed, int64_t lo, int64_t hi, unsigned shift, uint8_t *out)
{
    int64_t i = lo;
    for (; i + 8 <= hi; i += 8) {
        uint64_t eight;
        memcpy(&eight, packed + i, 8);
        eight = (eight >> shift) & 0x0101010101010101ull;
        memcpy(out + i, &eight, 8);
    }
    for (; i < hi; ++i) out[i] = (packed[i] >> shift) & 1u;
}

void unpack_range(const uint8_t *packed, int64_t lo, int64_t hi, uint8_t *valid4, uint8_t *terminated, uint8_t *invalid)
{
    if (valid4) {
        int64_t i = lo;
#if defined(ML2048_HAVE_AVX2_PATH)
        static const bool avx2 = __builtin_cpu_supports("avx2");
        if (avx2) i = unpack_valid_avx2(packed, lo, hi, valid4);
#endif
        for (; i < hi; ++i) {
            const uint32_t w = spread4(packed[i]);
            memcpy(valid4 + 4 * i, &w, 4);
        }
    }
    if (terminated) unpack_bit(packed, lo, hi, 4u, terminated);
    if (invalid) unpack_bit(packed, lo, hi, 5u, invalid);
}

}  // namespace

extern "C" void ml2048_unpack_flags(const uint8_t *packed, int64_t num_games, uint8_t *valid4, uint8_t *terminated, uint8_t *invalid,
                                    int32_t threads)
{
    if (!packed || num_games <= 0) return;
    const int64_t min_per_thread = 1 << 16;
    int64_t t = threads > 0 ? threads : 1;
    if (t > (num_games + min_per_thread - 1) / min_per_thread) t = (num_games + min_per_thread - 1) / min_per_thread;
    if (t <= 1) {
        unpack_range(packed, 0, num_games, valid4, terminated, invalid);
        return;
    }
    const int64_t per = ((num_games + t - 1) / t + 63) / 64 * 64;
    std::vector<std::thread> pool;
    for (int64_t k = 1; k < t; ++k) {
        const int64_t lo = k * per, hi = lo + per < num_games ? lo + per : num_games;
        if (lo < hi) pool.emplace_back(unpack_range, packed, lo, hi, valid4, terminated, invalid);
    }
    unpack_range(packed, 0, per < num_games ? per : num_games, valid4, terminated, invalid);
    for (auto &th : pool) th.join();
}

// The same for a result that arrives in slices: slice k is complete when CUDA event events[k] (recorded on the copy stream
// after the slice's packed bytes) has fired.  The worker threads are started ONCE per call; each waits for the event of a
// slice itself and expands its share of it, so that the expansion of slice k overlaps the transfers of the slices behind it
// and no thread is created per slice.  Returns 0 or the cudaError_t of a failed wait.
extern "C" int ml2048_unpack_flags_sliced(const uint8_t *packed, int32_t num_slices, const int64_t *slice_lo, const int64_t *slice_hi,
                                          void *const *events, uint8_t *valid4, uint8_t *terminated, uint8_t *invalid, int32_t threads)
{
    if (!packed || num_slices <= 0 || !slice_lo || !slice_hi) return 0;
    const int t = threads > 0 ? (threads < 64 ? threads : 64) : 1;
    int status[64] = {0};
    auto work = [&](int who) {
        for (int k = 0; k < num_slices; ++k) {
            if (events && events[k]) {
                const cudaError_t e = cudaEventSynchronize(static_cast<cudaEvent_t>(events[k]));
                if (e != cudaSuccess) {
                    status[who] = (int)e;
                    return;
                }
            }
            const int64_t lo = slice_lo[k], n = slice_hi[k] - lo;
            const int64_t per = ((n + t - 1) / t + 63) / 64 * 64;  // whole 64-game groups: no two threads share a cache line of flags
            const int64_t a = lo + who * per, b = a + per < lo + n ? a + per : lo + n;
            if (a < b) unpack_range(packed, a, b, valid4, terminated, invalid);
        }
    };
    std::vector<std::thread> pool;
    for (int who = 1; who < t; ++who) pool.emplace_back(work, who);
    work(0);
    for (auto &th : pool) th.join();
    for (int who = 0; who < t; ++who)
        if (status[who]) return status[who];
    return 0;
}

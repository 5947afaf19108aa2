// board_ops.cuh -- register-resident 2048 board arithmetic for sm_100a.
//
// A board is four 32-bit words (one per row), one byte per cell holding the tile exponent
// (0 = empty), cell = row*4+col with column 0 in the least significant byte -- i.e. exactly the 16
// bytes of the reference's `board` field read as a little-endian uint4 (game_numba.py:13-20, :542).
// Everything here is branch-free byte-SWAR on those four words: PRMT (__byte_perm) for the data
// movement, carry-free adds for the per-byte zero tests (cell values are <= 17, so `byte + 0x7f`
// never carries into the next byte).  No local memory, no lookup tables.
#pragma once
#include <stdint.h>

namespace ml2048 {

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t s) { return __byte_perm(a, b, s); }

// PRMT in its generic PTX form: a selector nibble with bit 3 set replicates the SIGN of the selected byte
// over the whole result byte (__byte_perm masks that bit away, so this needs inline PTX).
__device__ __forceinline__ uint32_t prmt_sign(uint32_t a, uint32_t b, uint32_t s)
{
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(s));
    return d;
}

constexpr uint32_t kHi = 0x80808080u;
constexpr uint32_t kLo7 = 0x7f7f7f7fu;

// 0x80 in every byte whose cell is non-empty
__device__ __forceinline__ uint32_t occupied_flags(uint32_t row) { return (row + kLo7) & kHi; }

// What one move fused.  `gain` is the reference's reward_fn_normal (game_numba.py:408-438),
// `rank` reward_fn_rank (:469-484), `count` the merged.sum() of reward_fn_maxcell (:502), and
// `log` the `merged` array itself as sixteen 4-bit counters (at most 8 fusions per move).
struct Fusions {
    uint32_t gain;
    uint32_t rank;
    uint32_t count;
    unsigned long long log;
};

template <bool kLog>
__device__ __forceinline__ void note_fusion(Fusions &f, uint32_t k)
{
    f.gain += 2u << k;  // two tiles of exponent k became one tile worth 2^(k+1)
    f.rank += k + 1u;
    f.count += 1u;
    if (kLog) {
        if (k < 16u)  // merged[] has 16 slots (game_numba.py:543); tile 65536 is never reached
            f.log += 1ull << (4u * k);
    }
}

// One line of four cells pushed toward byte 0.  Same result as the reference's _push_row
// (game_numba.py:48-90): stable compaction of the tiles, then equal neighbours fuse once, in
// order from the wall, a fused tile never fusing again.
template <bool kLog>
__device__ __forceinline__ uint32_t push_line(uint32_t w, Fusions &f)
{
    // stable compaction: close the gap at byte 2, then 1, then 0
    if ((w & 0x00ff0000u) == 0u) w = prmt(w, 0u, 0x4310);
    if ((w & 0x0000ff00u) == 0u) w = prmt(w, 0u, 0x4320);
    if ((w & 0x000000ffu) == 0u) w = prmt(w, 0u, 0x4321);
    // compacted cells a,b,c,d : x holds a^b, b^c, c^d, d
    const uint32_t x = w ^ (w >> 8);
    const bool ab = ((x & 0x000000ffu) == 0u) && ((w & 0x000000ffu) != 0u);
    const bool bc = ((x & 0x0000ff00u) == 0u) && ((w & 0x0000ff00u) != 0u) && !ab;
    const bool cd = ((x & 0x00ff0000u) == 0u) && ((w & 0x00ff0000u) != 0u) && !bc;
    const uint32_t a = w & 0xffu, b = (w >> 8) & 0xffu, c = (w >> 16) & 0xffu;
    if (ab) {
        note_fusion<kLog>(f, a);
        w = prmt(w, 0u, 0x4320) + 0x00000001u;  // [a+1, c, d, 0]
    }
    if (bc) {
        note_fusion<kLog>(f, b);
        w = prmt(w, 0u, 0x4310) + 0x00000100u;  // [a, b+1, d, 0]
    }
    if (cd) {
        note_fusion<kLog>(f, c);
        // after an a/b fusion the pair sits in bytes 1,2: [a+1, c+1, 0, 0]; otherwise [a, b, c+1, 0]
        w = ab ? ((w & 0x0000ffffu) + 0x00000100u) : ((w & 0x00ffffffu) + 0x00010000u);
    }
    return w;
}

__device__ __forceinline__ void transpose4x4(uint32_t &r0, uint32_t &r1, uint32_t &r2, uint32_t &r3)
{
    const uint32_t t0 = prmt(r0, r1, 0x5140);  // r0.0 r1.0 r0.1 r1.1
    const uint32_t t1 = prmt(r2, r3, 0x5140);
    const uint32_t t2 = prmt(r0, r1, 0x7362);  // r0.2 r1.2 r0.3 r1.3
    const uint32_t t3 = prmt(r2, r3, 0x7362);
    r0 = prmt(t0, t1, 0x5410);
    r1 = prmt(t0, t1, 0x7632);
    r2 = prmt(t2, t3, 0x5410);
    r3 = prmt(t2, t3, 0x7632);
}

// The move itself: direction dispatch of _step_kernel (game_numba.py:93-134) without a branch.
// action bit 1 = vertical (work on the transposed board), bit 0 = toward the high end (reverse lines).
template <bool kLog>
__device__ __forceinline__ void move_board(uint32_t &r0, uint32_t &r1, uint32_t &r2, uint32_t &r3, uint32_t action,
                                           Fusions &f)
{
    const bool vertical = (action & 2u) != 0u;
    const uint32_t flip = (action & 1u) ? 0x0123u : 0x3210u;
    uint32_t c0 = r0, c1 = r1, c2 = r2, c3 = r3;
    transpose4x4(c0, c1, c2, c3);
    uint32_t l0 = prmt(vertical ? c0 : r0, 0u, flip);
    uint32_t l1 = prmt(vertical ? c1 : r1, 0u, flip);
    uint32_t l2 = prmt(vertical ? c2 : r2, 0u, flip);
    uint32_t l3 = prmt(vertical ? c3 : r3, 0u, flip);
    l0 = prmt(push_line<kLog>(l0, f), 0u, flip);
    l1 = prmt(push_line<kLog>(l1, f), 0u, flip);
    l2 = prmt(push_line<kLog>(l2, f), 0u, flip);
    l3 = prmt(push_line<kLog>(l3, f), 0u, flip);
    c0 = l0, c1 = l1, c2 = l2, c3 = l3;
    transpose4x4(c0, c1, c2, c3);
    r0 = vertical ? c0 : l0;
    r1 = vertical ? c1 : l1;
    r2 = vertical ? c2 : l2;
    r3 = vertical ? c3 : l3;
}

// Valid-action mask, one byte per direction (left,right,up,down), as the little-endian word the
// reference stores in `valid_actions` (game_numba.py:259-289).  A direction is valid iff some tile
// can slide into an empty cell or fuse with its neighbour on that axis -- the predicate
// _line_movable (:215-256) enumerates pairwise, proven equal to "the move changes the board" for
// all 18^4 lines (tests/test_oracle_golden.py::test_line_table_exhaustive).
__device__ __forceinline__ uint32_t valid_mask(uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3)
{
    const uint32_t n0 = occupied_flags(r0), n1 = occupied_flags(r1), n2 = occupied_flags(r2), n3 = occupied_flags(r3);
    const uint32_t z0 = n0 ^ kHi, z1 = n1 ^ kHi, z2 = n2 ^ kHi, z3 = n3 ^ kHi;
    // slides along rows: an empty cell with a tile further from the wall
    uint32_t s, left, right;
    s = z0 | (z0 << 8); left = (s | (s << 16)) & n0;
    s = z1 | (z1 << 8); left |= (s | (s << 16)) & n1;
    s = z2 | (z2 << 8); left |= (s | (s << 16)) & n2;
    s = z3 | (z3 << 8); left |= (s | (s << 16)) & n3;
    s = z0 | (z0 >> 8); right = (s | (s >> 16)) & n0;
    s = z1 | (z1 >> 8); right |= (s | (s >> 16)) & n1;
    s = z2 | (z2 >> 8); right |= (s | (s >> 16)) & n2;
    s = z3 | (z3 >> 8); right |= (s | (s >> 16)) & n3;
    // slides along columns
    const uint32_t up = (z0 & (n1 | n2 | n3)) | (z1 & (n2 | n3)) | (z2 & n3);
    const uint32_t down = (z3 & (n0 | n1 | n2)) | (z2 & (n0 | n1)) | (z1 & n0);
    // fusions: equal neighbours where the first one is a tile (0x20 marks empty cells so they never match)
    const uint32_t q0 = z0 >> 2, q1 = z1 >> 2, q2 = z2 >> 2, q3 = z3 >> 2;
    const uint32_t hz = ((((r0 ^ (r0 >> 8)) | q0) + kLo7) & (((r1 ^ (r1 >> 8)) | q1) + kLo7) &
                         (((r2 ^ (r2 >> 8)) | q2) + kLo7) & (((r3 ^ (r3 >> 8)) | q3) + kLo7)) & kHi;
    const uint32_t vt = ((((r0 ^ r1) | q0) + kLo7) & (((r1 ^ r2) | q1) + kLo7) & (((r2 ^ r3) | q2) + kLo7)) & kHi;
    const bool hfuse = hz != kHi, vfuse = vt != kHi;
    const uint32_t l = (left != 0u) || hfuse, r = (right != 0u) || hfuse;
    const uint32_t u = (up != 0u) || vfuse, d = (down != 0u) || vfuse;
    return l | (r << 8) | (u << 16) | (d << 24);
}

// Write `value` into cell `cell` (0..15) of the board.
__device__ __forceinline__ void put_cell(uint32_t &r0, uint32_t &r1, uint32_t &r2, uint32_t &r3, uint32_t cell, uint32_t value)
{
    const uint32_t v = value << ((cell & 3u) * 8u);
    const uint32_t row = cell >> 2;
    r0 |= (row == 0u) ? v : 0u;
    r1 |= (row == 1u) ? v : 0u;
    r2 |= (row == 2u) ? v : 0u;
    r3 |= (row == 3u) ? v : 0u;
}

// Replay-mode spawn position: first entry of the permutation row `perm` (16 bytes, a permutation of
// 0..15) whose cell is empty -- the table walk of _spawn2 (game_numba.py:198-204) done as a 16-lane
// byte gather: PRMT looks every entry up in the 16-byte "empty" table (z0..z3, 0x80 = empty).
// Returns the cell index, or 16 when the board is full.
__device__ __forceinline__ uint32_t first_empty_in_order(uint4 perm, uint32_t z0, uint32_t z1, uint32_t z2, uint32_t z3)
{
    uint32_t hit[4];
    const uint32_t pw[4] = {perm.x, perm.y, perm.z, perm.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t p = pw[i];
        uint32_t s = p & 0x07070707u;
        s |= s >> 4;                                 // byte0 = p0|p1<<4, byte2 = p2|p3<<4
        const uint32_t sel = prmt(s, 0u, 0x4420);    // selector nibbles (p0,p1,p2,p3) mod 8
        const uint32_t lo = prmt(z0, z1, sel);       // cells 0..7
        const uint32_t hi = prmt(z2, z3, sel);       // cells 8..15
        const uint32_t up = prmt_sign(p << 4, 0u, 0xba98);  // 0xff where p >= 8
        hit[i] = (lo & ~up) | (hi & up);
    }
    uint32_t h = hit[0], p = pw[0];
    if (h == 0u) { h = hit[1]; p = pw[1]; }
    if (h == 0u) { h = hit[2]; p = pw[2]; }
    if (h == 0u) { h = hit[3]; p = pw[3]; }
    if (h == 0u) return 16u;
    const uint32_t sh = (uint32_t)(__ffs((int)h) - 1) & ~7u;
    return (p >> sh) & 0xffu;
}

// 16-bit mask of empty cells from the per-row 0x80 flags (multiply gathers bits 7,15,23,31 into a nibble)
__device__ __forceinline__ uint32_t empties16(uint32_t z0, uint32_t z1, uint32_t z2, uint32_t z3)
{
    const uint32_t m = 0x00204081u;
    return ((z0 * m) >> 28) | (((z1 * m) >> 24) & 0xf0u) | (((z2 * m) >> 20) & 0xf00u) | (((z3 * m) >> 16) & 0xf000u);
}

// index of the k-th (0-based) set bit of a 16-bit mask; k < popc(mask)
__device__ __forceinline__ uint32_t kth_set_bit16(uint32_t mask, uint32_t k)
{
    uint32_t pos = 0u, c;
    c = __popc(mask & 0xffu);
    if (k >= c) { k -= c; pos = 8u; mask >>= 8; }
    c = __popc(mask & 0xfu);
    if (k >= c) { k -= c; pos += 4u; mask >>= 4; }
    c = __popc(mask & 0x3u);
    if (k >= c) { k -= c; pos += 2u; mask >>= 2; }
    c = mask & 1u;
    if (k >= c) { pos += 1u; }
    return pos;
}

// max tile exponent of a board
__device__ __forceinline__ uint32_t max_cell(uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3)
{
    uint32_t m = __vmaxu4(__vmaxu4(r0, r1), __vmaxu4(r2, r3));
    m = __vmaxu4(m, m >> 16);
    m = __vmaxu4(m, m >> 8);
    return m & 0xffu;
}

// Philox4x32-10 (Salmon et al., SC'11), counter-based: out = f(counter, key)
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key)
{
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
        const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += W0;
        key.y += W1;
    }
    return ctr;
}

}  // namespace ml2048

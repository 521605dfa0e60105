// Shared device/host helpers of the B200 cube kernels.
//
// Everything marked CUBE_HD compiles for both the device and the host, so the
// per-instance arithmetic of the kernels (byte-permute cubie updates, lazy
// twist reduction, sticker assembly) can be run on the CPU by the test-only
// emulation harness in tests/host_emul/ before any GPU time is spent.  The
// product library never calls the host versions.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define CUBE_HD __host__ __device__ __forceinline__
#else
#define CUBE_HD inline
#endif

// ---- byte permute / funnel shift (PRMT / SHF on sm_100a) --------------------
CUBE_HD uint32_t cube_prmt(uint32_t x, uint32_t y, uint32_t sel)
{
#if defined(__CUDA_ARCH__)
    return __byte_perm(x, y, sel);
#else
    uint64_t v = (uint64_t)x | ((uint64_t)y << 32);
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) r |= (uint32_t)((v >> (8 * ((sel >> (4 * i)) & 7))) & 0xffu) << (8 * i);
    return r;
#endif
}

CUBE_HD uint32_t cube_funnel_r(uint32_t lo, uint32_t hi, uint32_t shift)   // ((hi:lo) >> shift), shift < 32
{
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, shift);
#else
    uint64_t v = (uint64_t)lo | ((uint64_t)hi << 32);
    return (uint32_t)(v >> (shift & 31));
#endif
}

#include "cube_tables.cuh"

// ---- geometry ---------------------------------------------------------------
template <int SIZE> struct CubeGeom;
template <> struct CubeGeom<2> {
    static constexpr int S = 24, A = 6, R = 7, C = 21, D = 147, WORDS = 6, NCYC = 3, MW = CUBE_W2;
};
template <> struct CubeGeom<3> {
    static constexpr int S = 54, A = 12, R = 20, C = 24, D = 480, WORDS = 14, NCYC = 5, MW = CUBE_W3;
};

// ---- cubie state of the fused scramble kernel --------------------------------
// corners: byte = piece | twist_sum << 3   (twist_sum reduced lazily, see below)
// edges  : byte = piece | flip << 4
struct CubieState {
    uint32_t c0, c1;          // corner slots 0-3, 4-7
    uint32_t e0, e1, e2;      // edge slots 0-3 (U layer), 4-7 (middle), 8-11 (D layer); unused for 2x2x2
};

CUBE_HD void cubie_init(CubieState& s)
{
    s.c0 = 0x03020100u; s.c1 = 0x07060504u;
    s.e0 = 0x03020100u; s.e1 = 0x07060504u; s.e2 = 0x0b0a0908u;
}

// one face turn; tbl is the [word][CUBE_MOVE_ROWS] move-word table, m < CUBE_MOVE_ROWS
template <int SIZE>
CUBE_HD void cubie_move(CubieState& s, const uint32_t* tbl, uint32_t m)
{
    const uint32_t* t = tbl + m;
    uint32_t n0 = cube_prmt(s.c0, s.c1, t[0 * CUBE_MOVE_ROWS]) + t[2 * CUBE_MOVE_ROWS];
    uint32_t n1 = cube_prmt(s.c0, s.c1, t[1 * CUBE_MOVE_ROWS]) + t[3 * CUBE_MOVE_ROWS];
    s.c0 = n0; s.c1 = n1;
    if (SIZE == 3) {
        uint32_t tt = cube_prmt(s.e0, s.e2, t[6 * CUBE_MOVE_ROWS]);
        uint32_t m0 = cube_prmt(s.e0, s.e1, t[4 * CUBE_MOVE_ROWS]) ^ t[8 * CUBE_MOVE_ROWS];
        uint32_t m2 = cube_prmt(s.e2, s.e1, t[5 * CUBE_MOVE_ROWS]) ^ t[9 * CUBE_MOVE_ROWS];
        uint32_t m1 = cube_prmt(s.e1, tt, t[7 * CUBE_MOVE_ROWS]);
        s.e0 = m0; s.e1 = m1; s.e2 = m2;
    }
}

// The twist field (5 bits) only accumulates: every turn adds 0, 1 or 2 per corner.
// Because 4 == 1 (mod 3), f -> (f & 3) + (f >> 2) preserves f mod 3 and maps
// f <= 31 to <= 10 (and f <= 26 to <= 8), so one 4-op fold every 8 turns keeps the
// field from overflowing: 8 + 2*8 = 24 <= 31.
constexpr int kTwistFoldPeriod = 8;

CUBE_HD uint32_t cubie_fold_twist(uint32_t c)
{
    return (c & 0x1f1f1f1fu) + ((c >> 2) & 0x38383838u);
}

// exact twist mod 3 (field <= 31 on entry): two folds bring it to <= 4, then one conditional -3
CUBE_HD uint32_t cubie_reduce_twist(uint32_t c)
{
    c = cubie_fold_twist(cubie_fold_twist(c));
    uint32_t ge3 = ((c + 0x28282828u) >> 6) & 0x01010101u;      // (f + 5) >= 8  <=>  f >= 3
    return c - ge3 * 24u;                                        // 3 << 3
}

// colour words (bytes k = 0..2) of every slot -> sticker rows; luts have 32 entries
template <int SIZE>
CUBE_HD void cubie_to_stickers(const CubieState& s, const uint32_t* corner_lut, const uint32_t* edge_lut,
                               uint32_t* words /* CubeGeom<SIZE>::WORDS */)
{
    uint32_t L[20];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        L[q] = corner_lut[(s.c0 >> (8 * q)) & 0x1fu];
        L[4 + q] = corner_lut[(s.c1 >> (8 * q)) & 0x1fu];
    }
    if (SIZE == 3) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            L[8 + q] = edge_lut[(s.e0 >> (8 * q)) & 0x1fu];
            L[12 + q] = edge_lut[(s.e1 >> (8 * q)) & 0x1fu];
            L[16 + q] = edge_lut[(s.e2 >> (8 * q)) & 0x1fu];
        }
        cube_assemble3(L, words);
    } else {
        cube_assemble2(L, words);
    }
}

template <int SIZE>
CUBE_HD bool cubie_is_identity(const CubieState& s)     // twists must be reduced first
{
    bool ok = (s.c0 == 0x03020100u) & (s.c1 == 0x07060504u);
    if (SIZE == 3) ok = ok & (s.e0 == 0x03020100u) & (s.e1 == 0x07060504u) & (s.e2 == 0x0b0a0908u);
    return ok;
}

// ---- sticker-level helpers ----------------------------------------------------
// face uniformity (py333.py:229-233 / py222 isSolved) over a row held as bytes
template <int SIZE>
CUBE_HD bool stickers_solved(const uint8_t* row)
{
    constexpr int K = CubeGeom<SIZE>::S / 6;
    uint32_t diff = 0;
#pragma unroll
    for (int f = 0; f < 6; ++f) {
#pragma unroll
        for (int j = 1; j < K; ++j) diff |= (uint32_t)(row[f * K + j] ^ row[f * K]);
    }
    return diff == 0;
}

// error codes of the C ABI
#include "../../include/cube_b200.h"

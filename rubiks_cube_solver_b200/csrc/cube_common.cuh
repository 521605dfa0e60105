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
    // raw PRMT: __byte_perm() masks the selector with 0x7777 first (one extra LOP3 per permute);
    // every selector in this library keeps the sign-replicate bits clear by construction
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(x), "r"(y), "r"(sel));
    return r;
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

// x >> 16 on the FMA pipe (IMAD.HI) instead of the ALU pipe (SHF): the scramble kernel is
// bound by the ALU pipe (PRMT / LOP3 / SHF share it) while the FMA pipe idles.
CUBE_HD uint32_t cube_hi16(uint32_t x)
{
#if defined(__CUDA_ARCH__)
    uint32_t r;
    asm("mul.hi.u32 %0, %1, 65536;" : "=r"(r) : "r"(x));
    return r;
#else
    return x >> 16;
#endif
}

// sum of byte products + c (IDP.4A): measured to issue beside PRMT (tools/microbench_pipes.cu), i.e. it
// runs on the FMA pipe -- used to pull ONE byte out of a register (weights = 0 except one) and scale
// and offset it in the same instruction, where a PRMT + multiply + add would load the ALU pipe
CUBE_HD uint32_t cube_dp4a(uint32_t x, uint32_t w, uint32_t c)
{
#if defined(__CUDA_ARCH__)
    return __dp4a(x, w, c);
#else
    uint32_t r = c;
    for (int i = 0; i < 4; ++i) r += ((x >> (8 * i)) & 0xffu) * ((w >> (8 * i)) & 0xffu);
    return r;
#endif
}

CUBE_HD uint32_t cube_shr2(uint32_t x)          // x >> 2, also on the FMA pipe
{
#if defined(__CUDA_ARCH__)
    uint32_t r;
    asm("mul.hi.u32 %0, %1, 1073741824;" : "=r"(r) : "r"(x));
    return r;
#else
    return x >> 2;
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
// edges  : byte = piece | a << 4 | b << 5 with flip = a ^ b (U-layer flips toggle a, D-layer flips b)
struct CubieState {
    uint32_t c0, c1;          // corner slots 0-3, 4-7
    uint32_t e0, e1, e2;      // edge slots 0-3 (U layer), 4-7 (middle), 8-11 (D layer); unused for 2x2x2
};

CUBE_HD void cubie_init(CubieState& s)
{
    s.c0 = 0x03020100u; s.c1 = 0x07060504u;
    s.e0 = 0x03020100u; s.e1 = 0x07060504u; s.e2 = 0x0b0a0908u;
}

// one face turn; tbl is the [word][CUBE_MOVE_ROWS] move-word table, byte_off = 4 * move row
template <int SIZE>
CUBE_HD void cubie_move_at(CubieState& s, const uint32_t* tbl, uint32_t byte_off)
{
    const uint32_t* t = reinterpret_cast<const uint32_t*>(reinterpret_cast<const char*>(tbl) + byte_off);
    // packed move words (gen_tables.py packed_words_3): PRMT only reads selector bits 15:0
    // D-layer corner q+4 sits under U-layer corner q and a side-face turn twists them in opposite
    // senses, so the D-layer twist word is 2*B (mod 3): one multiply-add, no table word
    const uint32_t A = t[0 * CUBE_MOVE_ROWS], B = t[1 * CUBE_MOVE_ROWS];
    const uint32_t n0 = cube_prmt(s.c0, s.c1, A) + B;
    const uint32_t n1 = cube_prmt(s.c0, s.c1, cube_hi16(A)) + 2u * B;
    s.c0 = n0; s.c1 = n1;
    if (SIZE == 3) {
        const uint32_t D = t[2 * CUBE_MOVE_ROWS], E = t[3 * CUBE_MOVE_ROWS], F = t[4 * CUBE_MOVE_ROWS];
        const uint32_t tt = cube_prmt(s.e0, s.e2, E);
        const uint32_t m0 = cube_prmt(s.e0, s.e1, D) ^ (F & 0x10101010u);
        const uint32_t m2 = cube_prmt(s.e2, s.e1, cube_hi16(D)) ^ (F & 0x20202020u);     // flip = bit4 ^ bit5
        const uint32_t m1 = cube_prmt(s.e1, tt, cube_hi16(E));
        s.e0 = m0; s.e1 = m1; s.e2 = m2;
    }
}

template <int SIZE>
CUBE_HD void cubie_move(CubieState& s, const uint32_t* tbl, uint32_t m)      // m < CUBE_MOVE_ROWS
{
    cubie_move_at<SIZE>(s, tbl, m * 4u);
}

// The twist field (5 bits) only accumulates: a turn adds 0, 1 or 2 to a corner byte that ends up
// in the U layer (c0 += B) and 0, 2 or 4 to one that ends up in the D layer (c1 += 2*B), and bytes
// migrate between the two registers.  Because 4 == 1 (mod 3), f -> (f & 3) + (f >> 2) preserves
// f mod 3 and maps f <= 26 to <= 8.  BOTH registers are folded after every 4 turns, so whatever
// route a byte takes it holds <= 8 + 4*4 = 24 before a fold, and <= 8 + 3*4 = 20 after the up to
// 3 turns that may follow the last fold.  (Folding the two registers on different schedules is
// wrong: a byte can dodge the folds by changing layer.)
CUBE_HD uint32_t cubie_fold_twist(uint32_t c)
{
    // f - 3 * (f >> 2) == (f & 3) + (f >> 2); t holds (f >> 2) at the field's position (bits 3..5)
    const uint32_t t = cube_shr2(c) & 0x38383838u;
    return c - 3u * t;
}

// exact twist mod 3 (field <= 31 on entry): two folds bring it to <= 4, then one conditional -3
CUBE_HD uint32_t cubie_reduce_twist(uint32_t c)
{
    c = cubie_fold_twist(cubie_fold_twist(c));
#if defined(__CUDA_ARCH__)
    uint32_t sh6;                                                // >> 6 on the FMA pipe
    asm("mul.hi.u32 %0, %1, 67108864;" : "=r"(sh6) : "r"(c + 0x28282828u));
#else
    const uint32_t sh6 = (c + 0x28282828u) >> 6;
#endif
    uint32_t ge3 = sh6 & 0x01010101u;                            // (f + 5) >= 8  <=>  f >= 3
    return c - ge3 * 24u;                                        // 3 << 3
}

CUBE_HD uint32_t cube_lut_at(const uint32_t* lut, uint32_t byte_off)
{
    return *reinterpret_cast<const uint32_t*>(reinterpret_cast<const char*>(lut) + byte_off);
}

// colour words (bytes k = 0..2) of every slot -> sticker rows; 32 corner / 64 edge LUT entries
template <int SIZE>
CUBE_HD void cubie_to_stickers(const CubieState& s, const uint32_t* corner_lut, const uint32_t* edge_lut,
                               uint32_t* words /* CubeGeom<SIZE>::WORDS */)
{
    // every cubie byte is < 64 here (twists reduced), so byte * 4 stays inside its byte:
    // one multiply per register, then one byte-extract per slot gives the LUT byte offset
    uint32_t L[20];
    const uint32_t c0 = s.c0 * 4u, c1 = s.c1 * 4u;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        L[q] = cube_lut_at(corner_lut, cube_prmt(c0, 0u, 0x4440u + q));
        L[4 + q] = cube_lut_at(corner_lut, cube_prmt(c1, 0u, 0x4440u + q));
    }
    if (SIZE == 3) {
        const uint32_t e0 = s.e0 * 4u, e1 = s.e1 * 4u, e2 = s.e2 * 4u;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            L[8 + q] = cube_lut_at(edge_lut, cube_prmt(e0, 0u, 0x4440u + q));
            L[12 + q] = cube_lut_at(edge_lut, cube_prmt(e1, 0u, 0x4440u + q));
            L[16 + q] = cube_lut_at(edge_lut, cube_prmt(e2, 0u, 0x4440u + q));
        }
        cube_assemble3(L, words);
    } else {
        cube_assemble2(L, words);
    }
}

template <int SIZE>
CUBE_HD bool cubie_is_identity(const CubieState& s)     // twists must be reduced first
{
    // corners first: a random state passes this with probability ~1e-8, so the edge test
    // (which has to fold the two flip bits) almost never executes
    bool ok = (s.c0 == 0x03020100u) & (s.c1 == 0x07060504u);
    if (SIZE == 3 && ok) {
        const uint32_t e0 = (s.e0 ^ ((s.e0 >> 1) & 0x10101010u)) & 0x1f1f1f1fu;
        const uint32_t e1 = (s.e1 ^ ((s.e1 >> 1) & 0x10101010u)) & 0x1f1f1f1fu;
        const uint32_t e2 = (s.e2 ^ ((s.e2 >> 1) & 0x10101010u)) & 0x1f1f1f1fu;
        ok = (e0 == 0x03020100u) & (e1 == 0x07060504u) & (e2 == 0x0b0a0908u);
    }
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

// Per-thread bodies of the cube kernels, host+device (CUBE_HD) so that the test-only
// emulation harness (tests/host_emul/) can run exactly this code on the CPU.
#pragma once
#include "cube_common.cuh"

// ---- K1: fused scramble -------------------------------------------------------------------
template <int SIZE>
CUBE_HD void scramble_apply_word(CubieState& st, const uint32_t* s_tbl, uint32_t w)
{
    // byte offsets of the four table rows: (m & 15) * 4 per byte (rows 12..15 are identity),
    // then one byte-extract per move
    const uint32_t w4 = (w * 4u) & 0x3c3c3c3cu;
#pragma unroll
    for (int k = 0; k < 4; ++k) cubie_move_at<SIZE>(st, s_tbl, cube_prmt(w4, 0u, 0x4440u + k));
}

// moves of the tile are a flat byte image in s_moves (row `tid` at byte tid*depth, any alignment)
template <int SIZE>
CUBE_HD void scramble_run_staged(CubieState& st, int tid, int depth, const uint8_t* s_moves, const uint32_t* s_tbl)
{
    const uint32_t* mw = reinterpret_cast<const uint32_t*>(s_moves);
    const uint32_t r = (uint32_t)tid * (uint32_t)depth;
    const uint32_t wi = r >> 2, sh = (r & 3u) << 3;
    const int nfull = depth >> 2, tail = depth & 3;
    uint32_t lo = mw[wi];
#pragma unroll 2
    for (int j = 0; j < nfull; ++j) {
        const uint32_t hi = mw[wi + j + 1];             // may run <= 4 bytes past the row (tile is padded)
        const uint32_t w = cube_funnel_r(lo, hi, sh);
        lo = hi;
        scramble_apply_word<SIZE>(st, s_tbl, w);
        st.c0 = cubie_fold_twist(st.c0);                 // every 4 turns, see cubie_fold_twist
        st.c1 = cubie_fold_twist(st.c1);
    }
    if (tail) {                                          // last 1..3 moves (uniform over the grid)
        const uint32_t w = cube_funnel_r(lo, mw[wi + nfull + 1], sh);
        const uint32_t w4 = (w * 4u) & 0x3c3c3c3cu;
        for (int k = 0; k < tail; ++k) cubie_move_at<SIZE>(st, s_tbl, cube_prmt(w4, 0u, 0x4440u + k));
    }
}

// reduce, judge, expand to stickers and store the row into the shared output tile
template <int SIZE>
CUBE_HD bool scramble_finish(CubieState& st, int tid, const uint32_t* s_clut, const uint32_t* s_elut, uint8_t* s_out)
{
    st.c0 = cubie_reduce_twist(st.c0);
    st.c1 = cubie_reduce_twist(st.c1);
    const bool ok = cubie_is_identity<SIZE>(st);
    uint32_t words[CubeGeom<SIZE>::WORDS];
    cubie_to_stickers<SIZE>(st, s_clut, s_elut, words);
    if (SIZE == 3) {
        // 54-byte rows: odd rows start 2 bytes off a word boundary -> shift by selector, no branch
        const uint32_t odd = (uint32_t)tid & 1u;
        const uint32_t sel = odd ? 0x5432u : 0x3210u;
        uint8_t* rowp = s_out + 54 * tid;
        uint32_t* wbase = reinterpret_cast<uint32_t*>(rowp + 2 * odd);
#pragma unroll
        for (int j = 0; j < 13; ++j) wbase[j] = cube_prmt(words[j], words[j + 1], sel);
        *reinterpret_cast<uint16_t*>(rowp + (odd ? 0 : 52)) = (uint16_t)(odd ? words[0] : words[13]);
    } else {
        // 24-byte rows as three 8-byte stores: a half-warp's 16 rows hit 16 distinct bank pairs
        uint64_t* rowp = reinterpret_cast<uint64_t*>(s_out + 24 * tid);
#pragma unroll
        for (int j = 0; j < 3; ++j) rowp[j] = (uint64_t)words[2 * j] | ((uint64_t)words[2 * j + 1] << 32);
    }
    return ok;
}

// ---- K2: one face turn of a sticker row in shared memory ------------------------------------
template <int SIZE>
CUBE_HD void walk_turn(uint8_t* row, uint32_t m, const uint32_t* s_cyc)
{
#pragma unroll
    for (int c = 0; c < CubeGeom<SIZE>::NCYC; ++c) {
        const uint32_t w = s_cyc[c * CUBE_MOVE_ROWS + m];
        const uint32_t a = w & 0xffu, b = (w >> 8) & 0xffu, cc = (w >> 16) & 0xffu, d = w >> 24;
        const uint8_t vb = row[b], vc = row[cc], vd = row[d], va = row[a];
        row[a] = vb; row[b] = vc; row[cc] = vd; row[d] = va;
    }
}

// face uniformity of a 2-byte-aligned row: s[i] == s[i-1] for every i that does not start a face
template <int SIZE>
CUBE_HD bool row_solved(const uint8_t* row)
{
    constexpr int S = CubeGeom<SIZE>::S, K = S / 6;
    const uint16_t* h = reinterpret_cast<const uint16_t*>(row);
    uint32_t diff = 0;
    uint32_t prev = 0;
#pragma unroll
    for (int i = 0; i < S / 2; ++i) {
        const uint32_t cur = h[i];
        const uint32_t b0 = cur & 0xffu, b1 = cur >> 8;
        if ((2 * i) % K != 0) diff |= b0 ^ prev;
        if ((2 * i + 1) % K != 0) diff |= b1 ^ b0;
        prev = b1;
    }
    return diff == 0;
}

// ---- K3: one-hot columns and vectors ----------------------------------------------------------
template <int DTYPE> struct OneHot;      // V = elements per 16-byte vector
template <> struct OneHot<0> { static constexpr int V = 8, ESIZE = 2; };   // bf16
template <> struct OneHot<1> { static constexpr int V = 4, ESIZE = 4; };   // f32
template <> struct OneHot<2> { static constexpr int V = 16, ESIZE = 1; };  // u8

// 16 bytes with element `p` set to 1
template <int DTYPE>
CUBE_HD void onehot_set(uint32_t* w, int p)
{
    if (DTYPE == 0) {
        const uint32_t v = (p & 1) ? 0x3f800000u : 0x00003f80u;
#pragma unroll
        for (int q = 0; q < 4; ++q) w[q] |= ((p >> 1) == q) ? v : 0u;
    } else if (DTYPE == 1) {
#pragma unroll
        for (int q = 0; q < 4; ++q) w[q] |= (p == q) ? 0x3f800000u : 0u;
    } else {
        const uint32_t v = 1u << (8 * (p & 3));
#pragma unroll
        for (int q = 0; q < 4; ++q) w[q] |= ((p >> 2) == q) ? v : 0u;
    }
}

// The 16-byte vector that starts `off` elements into segment `seg` of a stream of one-hot
// segments of width C; col[s] = position of the 1 in segment s (255 = none).  A vector touches at
// most two segments (V <= C).  Branch-free: onehot_set ignores positions outside [0, V), and a
// 255 column always lands outside; col[] must be readable one entry past the last segment.
template <int DTYPE, int C>
CUBE_HD void onehot_vector_at(int seg, int off, const uint8_t* col, uint32_t* w)
{
    w[0] = w[1] = w[2] = w[3] = 0u;
    onehot_set<DTYPE>(w, (int)col[seg] - off);
    if (C % OneHot<DTYPE>::V != 0) onehot_set<DTYPE>(w, C - off + (int)col[seg + 1]);   // segments straddle vectors
}

// vector number v of the stream
template <int DTYPE, int C>
CUBE_HD void onehot_vector(int v, const uint8_t* col, int n_seg, uint32_t* w)
{
    (void)n_seg;
    const int e0 = v * OneHot<DTYPE>::V;
    const int seg = e0 / C;
    onehot_vector_at<DTYPE, C>(seg, e0 - seg * C, col, w);
}

// column of the 1 for one (row, slot): 3x3x3 returns the column, 2x2x2 returns cubelet | ori << 4
template <int SIZE>
CUBE_HD uint32_t onehot_code(const uint8_t* row, int slot, const uint32_t* s_def, const uint8_t* s_lut0,
                             const uint8_t* s_lut1)
{
    const uint32_t d = s_def[slot];
    const uint32_t a = row[d & 0xffu], b = row[(d >> 8) & 0xffu], c = row[(d >> 16) & 0xffu];
    if (SIZE == 3 && slot >= 8) return s_lut1[(a + 10u * b) & 127u];
    return s_lut0[(a + 2u * b + 10u * c) & 127u];
}

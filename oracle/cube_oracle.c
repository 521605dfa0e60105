/* Plain-C restatement of the reference's cube hot path.
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): used by tests/, by
 * __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference
 * legs as the checker and the timed CPU baseline -- never by the product.
 *
 * Follows, sticker by sticker, what the reference computes:
 *   orc_step      cube_env.py:71-111   (doMove_3 py333.py:220-222 + isSolved_3 :229-233)
 *   orc_scramble  cube_env.py:50-69, 187-191 (a move sequence from solved / from given states)
 *   orc_solved    py333.py:229-233 / py222 isSolved  (face uniformity)
 *   orc_columns   getOP_3 py333.py:224-227 + pos_to_state_3 :235-246; 2x2x2 cube_env.py:141-147
 *   orc_expand    cube_env.py:212-238  (all A children + their one-hot columns + solved flags)
 * Instances are independent; loops over instances are OpenMP-parallel so the
 * baseline can use every host core ("cores" in bench.py = omp threads used).
 * Pinned against oracle/cube_np.py (itself pinned against the reference) in
 * tests/test_oracle.py.
 */
#include <stdint.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "cube_oracle_tables.h"

static int n_stickers(int size) { return size == 3 ? 54 : 24; }
static int n_actions(int size) { return size == 3 ? 12 : 6; }
static const unsigned char* move_row(int size, int a)
{
    return size == 3 ? ORC_MOVES_3[a] : ORC_MOVES_2[a];
}

static int solved_one(int size, const uint8_t* s)
{
    int k = n_stickers(size) / 6;
    for (int f = 0; f < 6; ++f)
        for (int j = 1; j < k; ++j)
            if (s[f * k + j] != s[f * k]) return 0;
    return 1;
}

static void move_one(int size, const uint8_t* in, int a, uint8_t* out)
{
    const unsigned char* row = move_row(size, a);
    int S = n_stickers(size);
    for (int i = 0; i < S; ++i) out[i] = in[row[i]];
}

/* column of the single 1 in every one-hot row; cols has 20 (3x3x3) or 7 (2x2x2) entries */
static void columns_one(int size, const uint8_t* s, uint8_t* cols)
{
    if (size == 3) {
        for (int q = 0; q < 8; ++q) {
            int h = s[ORC_CORNER_DEFS_3[q][0]] + 2 * s[ORC_CORNER_DEFS_3[q][1]]
                  + 10 * s[ORC_CORNER_DEFS_3[q][2]];
            cols[q] = h < 64 ? ORC_CORNER_COL_3[h] : 0;
        }
        for (int q = 0; q < 12; ++q) {
            int h = s[ORC_EDGE_DEFS_3[q][0]] + 10 * s[ORC_EDGE_DEFS_3[q][1]];
            cols[8 + q] = h < 64 ? ORC_EDGE_COL_3[h] : 0;
        }
    } else {
        for (int c = 0; c < 7; ++c) cols[c] = 255;
        for (int p = 0; p < 7; ++p) {
            int h = s[ORC_PIECE_DEFS_2[p][0]] + 2 * s[ORC_PIECE_DEFS_2[p][1]]
                  + 10 * s[ORC_PIECE_DEFS_2[p][2]];
            int cubelet = h < 64 ? ORC_PIECE_INDS_2[h][0] : 0;
            int ori = h < 64 ? ORC_PIECE_INDS_2[h][1] : 0;
            cols[cubelet] = (uint8_t)(3 * p + ori);
        }
    }
}

int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void orc_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int orc_step(int size, uint8_t* states, const uint8_t* actions, int64_t n,
             uint8_t* solved, float* reward)
{
    int S = n_stickers(size), A = n_actions(size);
    int bad = 0;
#pragma omp parallel for schedule(static) reduction(| : bad)
    for (int64_t i = 0; i < n; ++i) {
        uint8_t tmp[54];
        if (actions[i] >= A) { bad |= 1; continue; }
        move_one(size, states + i * S, actions[i], tmp);
        memcpy(states + i * S, tmp, (size_t)S);
        int ok = solved_one(size, tmp);
        if (solved) solved[i] = (uint8_t)ok;
        if (reward) reward[i] = ok ? 1.0f : -1.0f;
    }
    return bad ? -1 : 0;
}

/* init == NULL: every instance starts solved.  per_step_solved may be NULL. */
int orc_scramble(int size, const uint8_t* init, const uint8_t* moves, int64_t n, int depth,
                 uint8_t* states_out, uint8_t* solved, float* reward, uint8_t* per_step_solved,
                 int64_t* solved_count)
{
    int S = n_stickers(size), A = n_actions(size), K = S / 6;
    int bad = 0;
    int64_t cnt = 0;
#pragma omp parallel for schedule(static) reduction(| : bad) reduction(+ : cnt)
    for (int64_t i = 0; i < n; ++i) {
        uint8_t a[54], b[54];
        if (init) memcpy(a, init + i * S, (size_t)S);
        else for (int j = 0; j < S; ++j) a[j] = (uint8_t)(j / K);
        uint8_t *cur = a, *nxt = b;
        for (int k = 0; k < depth; ++k) {
            int m = moves[i * depth + k];
            if (m >= A) { bad |= 1; break; }
            move_one(size, cur, m, nxt);
            uint8_t* t = cur; cur = nxt; nxt = t;
            if (per_step_solved) per_step_solved[i * depth + k] = (uint8_t)solved_one(size, cur);
        }
        memcpy(states_out + i * S, cur, (size_t)S);
        int ok = solved_one(size, cur);
        if (solved) solved[i] = (uint8_t)ok;
        if (reward) reward[i] = ok ? 1.0f : -1.0f;
        cnt += ok;
    }
    if (solved_count) *solved_count = cnt;
    return bad ? -1 : 0;
}

int orc_solved(int size, const uint8_t* states, int64_t n, uint8_t* solved, float* reward)
{
    int S = n_stickers(size);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        int ok = solved_one(size, states + i * S);
        if (solved) solved[i] = (uint8_t)ok;
        if (reward) reward[i] = ok ? 1.0f : -1.0f;
    }
    return 0;
}

/* cols: [n,20] or [n,7] */
int orc_columns(int size, const uint8_t* states, int64_t n, uint8_t* cols)
{
    int S = n_stickers(size), R = size == 3 ? 20 : 7;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) columns_one(size, states + i * S, cols + i * R);
    return 0;
}

/* one-hot as uint8 [n,R,C] (C = 24 or 21) */
int orc_encode_u8(int size, const uint8_t* states, int64_t n, uint8_t* out)
{
    int S = n_stickers(size), R = size == 3 ? 20 : 7, C = size == 3 ? 24 : 21;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        uint8_t cols[20];
        columns_one(size, states + i * S, cols);
        uint8_t* o = out + i * R * C;
        memset(o, 0, (size_t)(R * C));
        for (int r = 0; r < R; ++r)
            if (cols[r] != 255) o[r * C + cols[r]] = 1;
    }
    return 0;
}

/* children [n,A,S] (may be NULL), cols [n,A,R] (may be NULL), solved [n,A] */
int orc_expand(int size, const uint8_t* states, int64_t n, uint8_t* children, uint8_t* cols,
               uint8_t* solved)
{
    int S = n_stickers(size), A = n_actions(size), R = size == 3 ? 20 : 7;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        uint8_t tmp[54];
        for (int a = 0; a < A; ++a) {
            move_one(size, states + i * S, a, tmp);
            if (children) memcpy(children + (i * A + a) * S, tmp, (size_t)S);
            if (cols) columns_one(size, tmp, cols + (i * A + a) * R);
            if (solved) solved[i * A + a] = (uint8_t)solved_one(size, tmp);
        }
    }
    return 0;
}

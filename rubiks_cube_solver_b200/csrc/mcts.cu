// Batched MCTS tree store (sm_100a): the reference's MCTS (mcts.py:17-154) for B cubes at once, one
// thread per tree, every tree at the same simulation index.
//   mcts_traverse_kernel   traverse (mcts.py:52-81): walk from the root through expanded nodes --
//       random action while a node's children have no visits (mcts.py:69-70), otherwise the first
//       argmax of U + W - L with U = cpuct * P * (sqrt(sum N) / (1 + N)) rounded to float32 after
//       every operation like the reference's expression under NumPy >= 2 (mcts.py:142-152) -- adding
//       the virtual loss (mcts.py:77), stepping the real cube (env.step) and following the STORED key
//       of the chosen child, until a key that is not in the tree: that cube is the leaf.
//   mcts_update_kernel     expand's bookkeeping (mcts.py:103-110: the entry is stored under the real
//       observation's key and overwrites an existing one), backpropagate (mcts.py:115-130: W = max,
//       L -= 150, N += 1) and train's exit test (mcts.py:45-50: first solved child of the new leaf).
// The leaf batch itself (children, codes, done flags) is cube_expand_codes; the value / policy come
// from the caller's network.  Keys are the one-hot rows' column indices: 20 bytes (3x3x3) or 7 + 1
// padding byte (2x2x2).
//
// What bounds these kernels is the chain of DEPENDENT global loads along a path, not bandwidth (a tree
// is ~8 KB, 65 536 trees are 0.5 GB: every step of every tree misses L2).  Round 1 paid three round
// trips per step (scan of all node keys, the node's statistics, the child's key) with byte-wise strided
// loads and kept the cube in a local-memory array.  Now:
//   * dict lookups are memoised per edge: child_slot[node, a] holds the slot of the child's node once a
//     traversal has found it, child_seen[node, a] how many node slots have already been compared with
//     that child's key without a match (slots only ever get appended and their keys never change), so a
//     step is ONE round trip (the node's N / P / W / L rows and its two memo rows, read as vectors) and
//     a key scan only ever looks at slots added since the last visit of that edge;
//   * the update kernel does not scan either when the real leaf's key is the key the traversal just
//     failed to find (always, unless the lossy 3x3x3 encoding aliases two cubes);
//   * the cubes of a block live in a shared-memory tile that is loaded and stored coalesced; a move is
//     the five / three sticker 4-cycles of walk_turn (cube_threads.cuh), in place.
// One thread per tree keeps 2048 independent chains in flight per SM.
#include <cuda_runtime.h>
#include "cube_kernels.h"
#include "cube_threads.cuh"

namespace {

constexpr int kBlock = 128;

template <int SIZE> struct MctsGeom;
template <> struct MctsGeom<2> { static constexpr int S = 24, A = 6, KW = 2; };    // key words
template <> struct MctsGeom<3> { static constexpr int S = 54, A = 12, KW = 5; };

template <int KW>
__device__ __forceinline__ void load_key(const uint8_t* p, uint32_t* key)
{
    const uint32_t* q = reinterpret_cast<const uint32_t*>(p);
#pragma unroll
    for (int w = 0; w < KW; ++w) key[w] = q[w];
}

// first slot in [from, n) whose key equals `key`, -1 if none; the loads of a batch are independent
template <int KW>
__device__ __forceinline__ int find_slot(const uint32_t* node_key, int from, int n, const uint32_t* key)
{
    for (int i0 = from; i0 < n; i0 += 4) {
        uint32_t k[4][KW];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int i = (i0 + j < n) ? i0 + j : n - 1;
#pragma unroll
            for (int w = 0; w < KW; ++w) k[j][w] = node_key[i * KW + w];
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            bool same = i0 + j < n;
#pragma unroll
            for (int w = 0; w < KW; ++w) same &= (k[j][w] == key[w]);
            if (same) return i0 + j;
        }
    }
    return -1;
}

// the A entries of a [.., A] row of 4-byte elements, as 8- / 16-byte vectors (rows are 24 / 48 bytes)
template <int A, class T>
__device__ __forceinline__ void load_row(const T* base, size_t row, T* out)   // no __restrict__: L is re-read after its own update
{
    if (A == 12) {
        const int4* p = reinterpret_cast<const int4*>(base + row);
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const int4 v = p[j];
            out[4 * j] = *reinterpret_cast<const T*>(&v.x); out[4 * j + 1] = *reinterpret_cast<const T*>(&v.y);
            out[4 * j + 2] = *reinterpret_cast<const T*>(&v.z); out[4 * j + 3] = *reinterpret_cast<const T*>(&v.w);
        }
    } else {
        const int2* p = reinterpret_cast<const int2*>(base + row);
#pragma unroll
        for (int j = 0; j < A / 2; ++j) {
            const int2 v = p[j];
            out[2 * j] = *reinterpret_cast<const T*>(&v.x); out[2 * j + 1] = *reinterpret_cast<const T*>(&v.y);
        }
    }
}

template <int SIZE>
__global__ void __launch_bounds__(kBlock)
mcts_traverse_kernel(cube_mcts_tree_t t, float cpuct, int virtual_loss)
{
    using G = MctsGeom<SIZE>;
    constexpr int S = G::S, A = G::A, KW = G::KW;
    __shared__ __align__(16) uint8_t s_rows[kBlock * S];
    __shared__ uint32_t s_cyc[CubeGeom<SIZE>::NCYC * CUBE_MOVE_ROWS];
    const int tid = threadIdx.x;
    const long long b0 = (long long)blockIdx.x * kBlock;
    const int cnt = (t.n_trees - b0) < kBlock ? (int)(t.n_trees - b0) : kBlock;
    for (int i = tid; i < CubeGeom<SIZE>::NCYC * CUBE_MOVE_ROWS; i += kBlock) s_cyc[i] = (SIZE == 3) ? kCycles3[i] : kCycles2[i];
    {   // the block's root rows: one contiguous piece of root_state (4-byte aligned: kBlock * S is a multiple of 4)
        const uint32_t* src = reinterpret_cast<const uint32_t*>(t.root_state + b0 * S);
        const int nw = cnt * S / 4;
        for (int i = tid; i < nw; i += kBlock) reinterpret_cast<uint32_t*>(s_rows)[i] = src[i];
        for (int i = nw * 4 + tid; i < cnt * S; i += kBlock) s_rows[i] = t.root_state[b0 * S + i];
    }
    __syncthreads();

    const long long b = b0 + tid;
    if (tid < cnt) {
        int d = 0;
        if (t.active[b]) {
            const int M = t.n_slots;
            uint8_t* row_p = s_rows + S * tid;
            const uint32_t* nkeys = reinterpret_cast<const uint32_t*>(t.node_key) + (size_t)b * M * KW;
            const int n = t.n_nodes[b];
            int rp = t.rand_ptr[b];
            uint32_t key[KW];
            load_key<KW>(t.root_key + (size_t)b * KW * 4, key);
            int seen = 0;
            int slot = find_slot<KW>(nkeys, 0, n, key);              // the root is slot 0 from the second simulation on
            if (slot < 0) seen = n;
            while (slot >= 0) {
                const size_t row = ((size_t)b * M + slot) * A;
                int nn[A], ll[A];
                float pp[A], ww[A];
                load_row<A>(t.N, row, nn);
                load_row<A>(t.P, row, pp);
                load_row<A>(t.W, row, ww);
                load_row<A>(t.L, row, ll);
                uint8_t memo_slot[A], memo_seen[A];
                {
                    const uint16_t* ps = reinterpret_cast<const uint16_t*>(t.child_slot + row);      // rows of 6 / 12 bytes
                    const uint16_t* pn = reinterpret_cast<const uint16_t*>(t.child_seen + row);
#pragma unroll
                    for (int j = 0; j < A / 2; ++j) {
                        const uint16_t vs = ps[j], vn = pn[j];
                        memo_slot[2 * j] = (uint8_t)vs; memo_slot[2 * j + 1] = (uint8_t)(vs >> 8);
                        memo_seen[2 * j] = (uint8_t)vn; memo_seen[2 * j + 1] = (uint8_t)(vn >> 8);
                    }
                }
                int total = 0;
#pragma unroll
                for (int a = 0; a < A; ++a) total += nn[a];
                int act = 0;
                if (total == 0) {                                                    // mcts.py:69-70
                    act = (rp < t.rand_cap) ? t.rand_table[(size_t)b * t.rand_cap + rp] : 0;
                    if (rp >= t.rand_cap) atomicOr(t.flags, 2);
                    if (act >= A) { atomicOr(t.flags, 8); act = 0; }                  // a draw outside 0..A-1: caller's error
                    ++rp;
                } else {                                                             // mcts.py:142-152
                    const double root = __dsqrt_rn((double)total);
                    float best = 0.0f;
#pragma unroll
                    for (int a = 0; a < A; ++a) {
                        const float tt = __double2float_rn(__ddiv_rn(root, 1.0 + (double)nn[a]));
                        const float u = __fmul_rn(__fmul_rn(cpuct, pp[a]), tt);
                        const float score = __fsub_rn(__fadd_rn(u, ww[a]), (float)ll[a]);
                        if (a == 0 || score > best) { best = score; act = a; }
                    }
                }
                if (d >= t.path_cap) { atomicOr(t.flags, 1); break; }
                t.path_node[(size_t)b * t.path_cap + d] = (uint8_t)slot;
                t.path_action[(size_t)b * t.path_cap + d] = (uint8_t)act;
                int l_act = 0, m_slot = 255, m_seen = 0;
#pragma unroll
                for (int a = 0; a < A; ++a)
                    if (a == act) { l_act = ll[a]; m_slot = memo_slot[a]; m_seen = memo_seen[a]; }
                t.L[row + act] = l_act + virtual_loss;                                // mcts.py:77
                walk_turn<SIZE>(row_p, (uint32_t)act, s_cyc);                         // env.step(act) on the real cube
                ++d;
                if (m_slot != 255) { slot = m_slot; continue; }                       // the child's node, found earlier
                load_key<KW>(t.child_key + (row + act) * (size_t)(KW * 4), key);
                slot = find_slot<KW>(nkeys, m_seen, n, key);
                if (slot >= 0) t.child_slot[row + act] = (uint8_t)slot;
                else { t.child_seen[row + act] = (uint8_t)n; seen = n; }
            }
            t.rand_ptr[b] = rp;
            uint32_t* mk = reinterpret_cast<uint32_t*>(t.miss_key) + (size_t)b * KW;
#pragma unroll
            for (int w = 0; w < KW; ++w) mk[w] = key[w];
            t.miss_seen[b] = seen;
        }
        t.path_len[b] = d;
    }
    __syncthreads();
    {   // the leaves (the roots of inactive trees) leave as one contiguous piece
        uint32_t* dst = reinterpret_cast<uint32_t*>(t.leaf_state + b0 * S);
        const int nw = cnt * S / 4;
        for (int i = tid; i < nw; i += kBlock) dst[i] = reinterpret_cast<const uint32_t*>(s_rows)[i];
        for (int i = nw * 4 + tid; i < cnt * S; i += kBlock) t.leaf_state[b0 * S + i] = s_rows[i];
    }
}

template <int SIZE>
__global__ void __launch_bounds__(kBlock)
mcts_update_kernel(cube_mcts_tree_t t, const uint8_t* __restrict__ leaf_key, const uint8_t* __restrict__ child_key_new,
                   const uint8_t* __restrict__ child_done_new, const float* __restrict__ value,
                   const float* __restrict__ policy, float value_min, int sim_index, int8_t* __restrict__ actions_out, int* __restrict__ n_actions, int* __restrict__ n_sims,
                   int* __restrict__ n_active)
{
    using G = MctsGeom<SIZE>;
    constexpr int A = G::A, KW = G::KW;
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    bool still_active = false;
    if (b < t.n_trees && t.active[b]) {
        still_active = true;
        const int M = t.n_slots;
        uint32_t key[KW], miss[KW];
        load_key<KW>(leaf_key + (size_t)b * KW * 4, key);
        load_key<KW>(t.miss_key + (size_t)b * KW * 4, miss);
        uint32_t* nkeys = reinterpret_cast<uint32_t*>(t.node_key) + (size_t)b * M * KW;
        int n = t.n_nodes[b];
        bool same = true;
#pragma unroll
        for (int w = 0; w < KW; ++w) same &= (key[w] == miss[w]);
        // the traversal compared `miss` with slots [0, miss_seen) ... and, through the edge's memo, with all n of them
        int slot = (same && t.miss_seen[b] >= n) ? -1 : find_slot<KW>(nkeys, 0, n, key);
        bool overflow = false;
        if (slot < 0) {
            if (n >= M) { atomicOr(t.flags, 4); overflow = true; }
            else { slot = n; t.n_nodes[b] = n + 1; }
        }
        if (!overflow) {
#pragma unroll
            for (int w = 0; w < KW; ++w) nkeys[slot * KW + w] = key[w];
            const size_t row = ((size_t)b * M + slot) * A;
            const uint32_t* ckn = reinterpret_cast<const uint32_t*>(child_key_new) + (size_t)b * A * KW;
            uint32_t* ck = reinterpret_cast<uint32_t*>(t.child_key) + row * KW;
#pragma unroll
            for (int i = 0; i < A * KW; ++i) ck[i] = ckn[i];
            int first_done = -1;
#pragma unroll
            for (int a = A - 1; a >= 0; --a) {
                const uint8_t dn = child_done_new[(size_t)b * A + a];
                t.child_done[row + a] = dn;
                if (dn) first_done = a;
                t.P[row + a] = policy[(size_t)b * A + a];
                t.W[row + a] = value_min;
                t.N[row + a] = 0;
                t.L[row + a] = 0;
                t.child_slot[row + a] = 255;                                     // an overwritten node forgets its memos
                t.child_seen[row + a] = 0;
            }
            const float v = value[b];
            const int len = t.path_len[b];
            for (int d = 0; d < len; ++d) {                                          // mcts.py:122-129
                const size_t r = ((size_t)b * M + t.path_node[(size_t)b * t.path_cap + d]) * A + t.path_action[(size_t)b * t.path_cap + d];
                const float w = t.W[r];
                t.W[r] = (v > w) ? v : w;
                t.L[r] -= 150;
                t.N[r] += 1;
            }
            if (first_done >= 0) {                                                   // mcts.py:45-50
                const size_t o = (size_t)b * (t.path_cap + 1);
                for (int d = 0; d < len; ++d) actions_out[o + d] = (int8_t)t.path_action[(size_t)b * t.path_cap + d];
                actions_out[o + len] = (int8_t)first_done;
                n_actions[b] = len + 1;
                n_sims[b] = sim_index + 1;
                t.active[b] = 0;
                still_active = false;
            }
        }
    }
    if (n_active) {                                                                  // trees still searching after this simulation
        const unsigned bal = __ballot_sync(0xffffffffu, still_active);
        if ((threadIdx.x & 31) == 0 && bal) atomicAdd(n_active, __popc(bal));
    }
}

}  // namespace

namespace cube {

int launch_mcts_traverse(int size, const cube_mcts_tree_t& t, float cpuct, int virtual_loss, cudaStream_t stream)
{
    if (t.n_trees == 0) return 0;
    const unsigned blocks = (unsigned)((t.n_trees + kBlock - 1) / kBlock);
    if (size == 3) mcts_traverse_kernel<3><<<blocks, kBlock, 0, stream>>>(t, cpuct, virtual_loss);
    else mcts_traverse_kernel<2><<<blocks, kBlock, 0, stream>>>(t, cpuct, virtual_loss);
    return (int)cudaGetLastError();
}

int launch_mcts_update(int size, const cube_mcts_tree_t& t, const uint8_t* leaf_key, const uint8_t* child_key_new,
                       const uint8_t* child_done_new, const float* value, const float* policy, float value_min,
                       int sim_index, int8_t* actions_out, int* n_actions, int* n_sims, int* n_active, cudaStream_t stream)
{
    if (t.n_trees == 0) return 0;
    const unsigned blocks = (unsigned)((t.n_trees + kBlock - 1) / kBlock);
    if (size == 3)
        mcts_update_kernel<3><<<blocks, kBlock, 0, stream>>>(t, leaf_key, child_key_new, child_done_new, value, policy,
                                                             value_min, sim_index, actions_out, n_actions, n_sims, n_active);
    else
        mcts_update_kernel<2><<<blocks, kBlock, 0, stream>>>(t, leaf_key, child_key_new, child_done_new, value, policy,
                                                             value_min, sim_index, actions_out, n_actions, n_sims, n_active);
    return (int)cudaGetLastError();
}

}  // namespace cube

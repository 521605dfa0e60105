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
// The leaf batch itself (children, one-hot rows, done flags) is cube_expand; the value / policy come
// from the caller's network.  Keys are the one-hot rows' column indices: 20 bytes (3x3x3) or 7 + 1
// padding byte (2x2x2), compared word by word against the tree's node slots (at most one node per
// simulation, so a linear scan of <= 255 slots).
#include <cuda_runtime.h>
#include "cube_kernels.h"
#include "cube_common.cuh"

namespace {

template <int SIZE> struct MctsGeom;
template <> struct MctsGeom<2> { static constexpr int S = 24, A = 6, KW = 2, GS = 24; };    // key words, gather row stride
template <> struct MctsGeom<3> { static constexpr int S = 54, A = 12, KW = 5, GS = 56; };

template <int KW>
__device__ __forceinline__ int find_slot(const uint32_t* __restrict__ node_key, int n, const uint32_t* key)
{
    for (int i = 0; i < n; ++i) {
        const uint32_t* k = node_key + i * KW;
        bool same = true;
#pragma unroll
        for (int w = 0; w < KW; ++w) same &= (k[w] == key[w]);
        if (same) return i;
    }
    return -1;
}

template <int SIZE>
__global__ void __launch_bounds__(128)
mcts_traverse_kernel(cube_mcts_tree_t t, float cpuct, int virtual_loss)
{
    using G = MctsGeom<SIZE>;
    constexpr int S = G::S, A = G::A, KW = G::KW;
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= t.n_trees) return;
    if (!t.active[b]) { t.path_len[b] = 0; return; }
    const int M = t.n_slots;
    uint8_t st[56];
    for (int i = 0; i < S; ++i) st[i] = t.root_state[(size_t)b * S + i];
    uint32_t key[KW];
    for (int w = 0; w < KW; ++w) key[w] = reinterpret_cast<const uint32_t*>(t.root_key)[(size_t)b * KW + w];
    const uint32_t* nkeys = reinterpret_cast<const uint32_t*>(t.node_key) + (size_t)b * M * KW;
    const int n = t.n_nodes[b];
    int rp = t.rand_ptr[b];
    int d = 0;
    while (true) {
        const int slot = find_slot<KW>(nkeys, n, key);
        if (slot < 0) break;
        const size_t row = ((size_t)b * M + slot) * A;
        int total = 0;
        for (int a = 0; a < A; ++a) total += t.N[row + a];
        int act = 0;
        if (total == 0) {                                                    // mcts.py:69-70
            act = (rp < t.rand_cap) ? t.rand_table[(size_t)b * t.rand_cap + rp] : 0;
            if (rp >= t.rand_cap) atomicOr(t.flags, 2);
            ++rp;
        } else {                                                             // mcts.py:142-152
            const double root = __dsqrt_rn((double)total);
            float best = 0.0f;
            for (int a = 0; a < A; ++a) {
                const float tt = __double2float_rn(__ddiv_rn(root, 1.0 + (double)t.N[row + a]));
                const float u = __fmul_rn(__fmul_rn(cpuct, t.P[row + a]), tt);
                const float score = __fsub_rn(__fadd_rn(u, t.W[row + a]), (float)t.L[row + a]);
                if (a == 0 || score > best) { best = score; act = a; }
            }
        }
        if (d >= t.path_cap) { atomicOr(t.flags, 1); break; }
        t.path_node[(size_t)b * t.path_cap + d] = (uint8_t)slot;
        t.path_action[(size_t)b * t.path_cap + d] = (uint8_t)act;
        t.L[row + act] += virtual_loss;                                       // mcts.py:77
        {                                                                    // env.step(act) on the real cube
            const uint8_t* g = (SIZE == 3 ? kGather3 : kGather2) + act * G::GS;
            uint8_t nx[56];
            for (int i = 0; i < S; ++i) nx[i] = st[g[i]];
            for (int i = 0; i < S; ++i) st[i] = nx[i];
        }
        const uint32_t* ck = reinterpret_cast<const uint32_t*>(t.child_key) + (row + act) * KW;
        for (int w = 0; w < KW; ++w) key[w] = ck[w];
        ++d;
    }
    t.path_len[b] = d;
    t.rand_ptr[b] = rp;
    for (int i = 0; i < S; ++i) t.leaf_state[(size_t)b * S + i] = st[i];
}

template <int SIZE>
__global__ void __launch_bounds__(128)
mcts_update_kernel(cube_mcts_tree_t t, const uint8_t* __restrict__ leaf_key, const uint8_t* __restrict__ child_key_new,
                   const uint8_t* __restrict__ child_done_new, const float* __restrict__ value,
                   const float* __restrict__ policy, float value_min, int sim_index, int8_t* __restrict__ actions_out,
                   int* __restrict__ n_actions, int* __restrict__ n_sims)
{
    using G = MctsGeom<SIZE>;
    constexpr int A = G::A, KW = G::KW;
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= t.n_trees || !t.active[b]) return;
    const int M = t.n_slots;
    uint32_t key[KW];
    for (int w = 0; w < KW; ++w) key[w] = reinterpret_cast<const uint32_t*>(leaf_key)[(size_t)b * KW + w];
    uint32_t* nkeys = reinterpret_cast<uint32_t*>(t.node_key) + (size_t)b * M * KW;
    int n = t.n_nodes[b];
    int slot = find_slot<KW>(nkeys, n, key);
    if (slot < 0) {
        if (n >= M) { atomicOr(t.flags, 4); return; }
        slot = n;
        t.n_nodes[b] = n + 1;
    }
    for (int w = 0; w < KW; ++w) nkeys[slot * KW + w] = key[w];
    const size_t row = ((size_t)b * M + slot) * A;
    const uint32_t* ckn = reinterpret_cast<const uint32_t*>(child_key_new) + (size_t)b * A * KW;
    uint32_t* ck = reinterpret_cast<uint32_t*>(t.child_key) + row * KW;
    for (int i = 0; i < A * KW; ++i) ck[i] = ckn[i];
    int first_done = -1;
    for (int a = A - 1; a >= 0; --a) {
        const uint8_t dn = child_done_new[(size_t)b * A + a];
        t.child_done[row + a] = dn;
        if (dn) first_done = a;
        t.P[row + a] = policy[(size_t)b * A + a];
        t.W[row + a] = value_min;
        t.N[row + a] = 0;
        t.L[row + a] = 0;
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
    }
}

}  // namespace

namespace cube {

int launch_mcts_traverse(int size, const cube_mcts_tree_t& t, float cpuct, int virtual_loss, cudaStream_t stream)
{
    if (t.n_trees == 0) return 0;
    const unsigned blocks = (unsigned)((t.n_trees + 127) / 128);
    if (size == 3) mcts_traverse_kernel<3><<<blocks, 128, 0, stream>>>(t, cpuct, virtual_loss);
    else mcts_traverse_kernel<2><<<blocks, 128, 0, stream>>>(t, cpuct, virtual_loss);
    return (int)cudaGetLastError();
}

int launch_mcts_update(int size, const cube_mcts_tree_t& t, const uint8_t* leaf_key, const uint8_t* child_key_new,
                       const uint8_t* child_done_new, const float* value, const float* policy, float value_min,
                       int sim_index, int8_t* actions_out, int* n_actions, int* n_sims, cudaStream_t stream)
{
    if (t.n_trees == 0) return 0;
    const unsigned blocks = (unsigned)((t.n_trees + 127) / 128);
    if (size == 3)
        mcts_update_kernel<3><<<blocks, 128, 0, stream>>>(t, leaf_key, child_key_new, child_done_new, value, policy,
                                                          value_min, sim_index, actions_out, n_actions, n_sims);
    else
        mcts_update_kernel<2><<<blocks, 128, 0, stream>>>(t, leaf_key, child_key_new, child_done_new, value, policy,
                                                          value_min, sim_index, actions_out, n_actions, n_sims);
    return (int)cudaGetLastError();
}

}  // namespace cube

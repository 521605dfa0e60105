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
// What bounds these kernels is neither bandwidth nor arithmetic but the chain of DEPENDENT global loads
// along a path (a tree is ~8 KB, 65 536 trees are 0.5 GB: every step of every tree misses L2), and max-backup
// plus virtual loss make the tree a chain, so at simulation s a traversal is ~s steps long.  Measured history
// of the traversal (65 536 2x2x2 trees, ncu at simulation ~30, profiles/r02_mcts_*):
//   round 1   thread per tree, scan of all node keys + statistics + child key per step (3 round trips),
//             byte-wise strided loads, the cube in a local-memory array                            ~260 us
//   v2        per-edge memo of the dict lookup, vector loads, cube in shared memory; every thread in
//             its own while-loop: the 32 threads of a warp drift apart and serialise                165 us
//   v3        the same in lockstep (one step per iteration, reconverged): 13 warps per SM, ~640 dependent
//             instructions per step, and some tree of the warp is at its leaf -- scanning keys -- in
//             nearly every iteration                                                                171 us
//   v4        8 lanes per tree (one action per lane, shuffle reductions): occupancy 55 %, but 3.3 times the
//             instructions (every lane executes every iteration): issue-bound                       211 us
//   v5 (this) thread per tree in lockstep with NO key scan in the traversal at all: the dict lookup
//             `key in children_and_data` is resolved EAGERLY when a node is stored (update kernel): child_slot
//             [node, a] is the slot of the child's node or 255 = not in the tree, kept exact by (i) looking
//             the new node's A child keys up among the existing nodes and (ii) pointing every existing edge
//             whose child key is the new key at the new slot.  A step is ONE round trip (the node's N / P / W /
//             L rows and its child_slot row, read as vectors).
// The update's back-propagation uses fire-and-forget reductions (RED) on a path that is fetched 16 steps at
// a time, so it has no load -> store chain; the cubes of a block live in a shared-memory tile that is loaded
// and stored coalesced, a move is the five / three sticker 4-cycles of walk_turn (cube_threads.cuh), in place.
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

// first slot in [from, n) whose key equals `key`, -1 if none.  No early exit: every load of the scan is
// independent of every compare, so they are all in flight together (a tree's keys are contiguous)
template <int KW>
__device__ __forceinline__ int find_slot(const uint32_t* node_key, int from, int n, const uint32_t* key)
{
    int found = 0x7fffffff;
#pragma unroll 4
    for (int i = from; i < n; ++i) {
        bool same = true;
#pragma unroll
        for (int w = 0; w < KW; ++w) same &= (node_key[i * KW + w] == key[w]);
        found = (same && i < found) ? i : found;
    }
    return found == 0x7fffffff ? -1 : found;
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
    // one traversal = one simulation: the device-side simulation index (for CUDA-graph replays of the
    // whole simulation, where the host cannot pass a fresh index) is advanced here and read by the update
    if (t.sim_counter && blockIdx.x == 0 && tid == 0) *t.sim_counter += 1;
    __syncthreads();

    const long long b = b0 + tid;
    const bool mine = tid < cnt && t.active[b];
    const int M = t.n_slots;
    uint8_t* row_p = s_rows + S * tid;
    const size_t tree0 = mine ? (size_t)b * M : 0;
    int rp = 0, d = 0, slot = -1;
    if (mine) {
        const int n = t.n_nodes[b];
        rp = t.rand_ptr[b];
        if (n > 0) {                             // the root is slot 0 from the second simulation on: one compare, else a scan
            const uint32_t* nkeys = reinterpret_cast<const uint32_t*>(t.node_key) + tree0 * KW;
            uint32_t key[KW];
            load_key<KW>(t.root_key + (size_t)b * KW * 4, key);
            bool at0 = true;
#pragma unroll
            for (int w = 0; w < KW; ++w) at0 &= (nkeys[w] == key[w]);
            slot = at0 ? 0 : find_slot<KW>(nkeys, 1, n, key);
        }
    }
    // the warp's trees advance in lockstep: one step per iteration, reconverged at its end
    while (__any_sync(0xffffffffu, slot >= 0)) {
        if (slot >= 0) {
            const size_t row = (tree0 + slot) * A;
            int nn[A], ll[A];
            float pp[A], ww[A];
            load_row<A>(t.N, row, nn);
            load_row<A>(t.P, row, pp);
            load_row<A>(t.W, row, ww);
            load_row<A>(t.L, row, ll);
            uint8_t memo[A];
            {
                const uint16_t* ps = reinterpret_cast<const uint16_t*>(t.child_slot + row);      // rows of 6 / 12 bytes
#pragma unroll
                for (int j = 0; j < A / 2; ++j) {
                    const uint16_t vs = ps[j];
                    memo[2 * j] = (uint8_t)vs; memo[2 * j + 1] = (uint8_t)(vs >> 8);
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
            if (d >= t.path_cap) {
                atomicOr(t.flags, 1);
                slot = -1;
            } else {
                t.path_node[(size_t)b * t.path_cap + d] = (uint8_t)slot;
                t.path_action[(size_t)b * t.path_cap + d] = (uint8_t)act;
                int l_act = 0, next = 255;
#pragma unroll
                for (int a = 0; a < A; ++a)
                    if (a == act) { l_act = ll[a]; next = memo[a]; }
                t.L[row + act] = l_act + virtual_loss;                            // mcts.py:77
                walk_turn<SIZE>(row_p, (uint32_t)act, s_cyc);                     // env.step(act) on the real cube
                ++d;
                slot = next == 255 ? -1 : next;          // `stored child key in children_and_data` (mcts.py:57), resolved by the update
            }
        }
        __syncwarp();
    }
    if (mine) t.rand_ptr[b] = rp;
    if (tid < cnt) t.path_len[b] = d;
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
                   const float* __restrict__ policy, float value_min, int sim_index, int8_t* __restrict__ actions_out,
                   int* __restrict__ n_actions, int* __restrict__ n_sims, int* __restrict__ n_active)
{
    using G = MctsGeom<SIZE>;
    constexpr int A = G::A, KW = G::KW;
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    bool still_active = false;
    if (t.sim_counter) sim_index = *t.sim_counter;
    if (b < t.n_trees && t.active[b]) {
        still_active = true;
        const int M = t.n_slots;
        const size_t tree0 = (size_t)b * M;
        uint32_t* nkeys = reinterpret_cast<uint32_t*>(t.node_key) + tree0 * KW;
        uint32_t* ckeys = reinterpret_cast<uint32_t*>(t.child_key) + tree0 * A * KW;      // the tree's edges, contiguous
        const uint8_t* pn = t.path_node + (size_t)b * t.path_cap;
        const uint8_t* pa = t.path_action + (size_t)b * t.path_cap;
        const int n = t.n_nodes[b];
        const int len = t.path_len[b];
        uint32_t key[KW], followed[KW];
        load_key<KW>(leaf_key + (size_t)b * KW * 4, key);
        // the key the traversal followed last and did not find: the stored child key of its last edge (the
        // root's key for an empty path).  The leaf is stored under the REAL observation's key (mcts.py:103-110);
        // the two differ only when the lossy 3x3x3 encoding aliases two cubes, and only then is a scan needed.
        if (len > 0) load_key<KW>(reinterpret_cast<const uint8_t*>(ckeys + ((size_t)pn[len - 1] * A + pa[len - 1]) * KW), followed);
        else load_key<KW>(t.root_key + (size_t)b * KW * 4, followed);
        bool same = true;
#pragma unroll
        for (int w = 0; w < KW; ++w) same &= (key[w] == followed[w]);
        int slot = same ? -1 : find_slot<KW>(nkeys, 0, n, key);
        const bool is_new = slot < 0;
        bool overflow = false;
        if (is_new) {
            if (n >= M) { atomicOr(t.flags, 4); overflow = true; }
            else { slot = n; t.n_nodes[b] = n + 1; }
        }
        if (!overflow) {
#pragma unroll
            for (int w = 0; w < KW; ++w) nkeys[slot * KW + w] = key[w];
            const size_t row = (tree0 + slot) * A;
            const uint32_t* ckn = reinterpret_cast<const uint32_t*>(child_key_new) + (size_t)b * A * KW;
            uint32_t* ck = ckeys + (size_t)slot * A * KW;
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
            }
            // (i) the stored node's own edges: is the child already a node?  (an overwritten node gets fresh memos
            // too.)  ONE pass over the tree's node keys against all A child keys, which sit in registers; the old
            // slots are never written by this kernel, so they are read through the non-coherent path, which lets
            // the compiler keep a whole batch of loads in flight (the slot stored above is compared from registers)
            uint32_t cw[A * KW];
#pragma unroll
            for (int i = 0; i < A * KW; ++i) cw[i] = ckn[i];
            int cs[A];
#pragma unroll
            for (int a = 0; a < A; ++a) {
                bool self = true;
#pragma unroll
                for (int w = 0; w < KW; ++w) self &= (cw[a * KW + w] == key[w]);
                cs[a] = self ? slot : 255;
            }
#pragma unroll 4
            for (int i = 0; i < n; ++i) {
                uint32_t kk[KW];
#pragma unroll
                for (int w = 0; w < KW; ++w) kk[w] = __ldg(nkeys + i * KW + w);
#pragma unroll
                for (int a = 0; a < A; ++a) {
                    bool hit = true;
#pragma unroll
                    for (int w = 0; w < KW; ++w) hit &= (cw[a * KW + w] == kk[w]);
                    cs[a] = hit ? i : cs[a];
                }
            }
#pragma unroll
            for (int a = 0; a < A; ++a) t.child_slot[row + a] = (uint8_t)cs[a];
            // (ii) a NEW key: every existing edge that leads to it now leads to a node.  The old nodes' child keys
            // are read-only here as well.
            if (is_new) {
                const int n_edges = n * A;
#pragma unroll 8
                for (int e = 0; e < n_edges; ++e) {
                    bool hit = true;
#pragma unroll
                    for (int w = 0; w < KW; ++w) hit &= (__ldg(ckeys + (size_t)e * KW + w) == key[w]);
                    if (hit) t.child_slot[tree0 * A + e] = (uint8_t)slot;
                }
            }
            const float v = value[b];
            // back-propagation (mcts.py:122-129: W = max(W, v), L -= 150, N += 1 per edge of the path) as
            // fire-and-forget reductions: one thread owns the tree, so there is no contention, and an edge
            // that occurs twice on a path (a bounce) gets both of its updates; the path arrives 16 steps at
            // a time.  The action list of a solved tree is copied from the same registers.
            const size_t o = (size_t)b * (t.path_cap + 1);
            for (int d0 = 0; d0 < len; d0 += 16) {
                uint32_t nw[4], aw[4];
                if (((t.path_cap | d0) & 3) == 0) {                              // word loads (path_cap is a multiple of 4)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        nw[j] = (d0 + 4 * j < len) ? reinterpret_cast<const uint32_t*>(pn + d0)[j] : 0u;
                        aw[j] = (d0 + 4 * j < len) ? reinterpret_cast<const uint32_t*>(pa + d0)[j] : 0u;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        nw[j] = aw[j] = 0u;
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            if (d0 + 4 * j + k < len) {
                                nw[j] |= (uint32_t)pn[d0 + 4 * j + k] << (8 * k);
                                aw[j] |= (uint32_t)pa[d0 + 4 * j + k] << (8 * k);
                            }
                    }
                }
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    if (d0 + k < len) {
                        const uint32_t node = (nw[k >> 2] >> (8 * (k & 3))) & 0xffu, act = (aw[k >> 2] >> (8 * (k & 3))) & 0xffu;
                        const size_t r = (tree0 + node) * A + act;
                        if (v >= 0.0f) atomicMax(reinterpret_cast<int*>(t.W + r), __float_as_int(v));
                        else atomicMin(reinterpret_cast<unsigned int*>(t.W + r), __float_as_uint(v));
                        atomicAdd(t.L + r, -150);
                        atomicAdd(t.N + r, 1);
                        if (first_done >= 0) actions_out[o + d0 + k] = (int8_t)act;
                    }
                }
            }
            if (first_done >= 0) {                                                   // mcts.py:45-50
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
        if ((threadIdx.x & 31) == 0 && bal) atomicAdd(n_active + sim_index, __popc(bal));
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

"""Batched MCTS leaf expansion: what `MCTS.expand` (mcts.py:83-113) computes for ONE leaf --
`model.predict(leaf)` (value + softmax policy, model.py:78-91), the A children reached by
`env.step(i)` on copies of the env and their `done` flags (mcts.py:96-101) -- for a batch of
leaves from many trees at once (`expand_leaves`), and the whole search for many cubes in lock-step
(`BatchedMCTS`: the tree store of mcts.py:17-154 -- dict keyed by the observation, max-backup,
virtual loss -- as device tensors, one tree per cube, one leaf batch per simulation).

One kernel launch produces the leaves' network input in the net's dtype, the child sticker
rows and the done flags (C ABI cube_expand; the register-resident leaf kernel for 2x2x2); the
caller's net is evaluated once on the whole batch.
"""
import torch

from . import ops


@torch.no_grad()
def expand_leaves(model, cube_size, leaves, obs_dtype=torch.bfloat16, want_child_onehot=False, model_device=None):
    """leaves: uint8 [N, S] sticker rows on a CUDA device (the `env.sim_cube` of every leaf).

    Returns dict(value float32 [N], policy float32 [N, A] (softmax, as model.predict),
    children uint8 [N, A, S], done bool [N, A], leaf_onehot [N, R, C], child_onehot or None).
    `W = [value_min] * A`, `N = L = [0] * A` of the reference's node tuple are constants the caller
    fills in.
    """
    res = ops.expand(cube_size, leaves, dtype=obs_dtype, want_children=True, want_child_onehot=want_child_onehot,
                     want_parent_onehot=True)
    dev = leaves.device
    mdev = dev if model_device is None else torch.device(model_device)
    x = res["parent_onehot"].to(mdev)
    param = next(iter(model.parameters()), None)
    if param is not None and x.dtype != param.dtype:
        x = x.to(param.dtype)
    value, logits = model(x)
    policy = torch.nn.functional.softmax(logits.float(), dim=-1)
    return dict(value=value.float().reshape(-1).to(dev), policy=policy.to(dev), children=res["children"],
                done=res["solved"].bool(), leaf_onehot=res["parent_onehot"], child_onehot=res["child_onehot"])


class BatchedMCTS(object):
    """`MCTS` of the reference (mcts.py:17-154) for B cubes at once: every tree runs the same
    simulation index together, so one simulation is a vectorised traversal, ONE leaf batch through
    `cube_expand` + the network, and a vectorised back-propagation.

    Per tree the reference keeps `children_and_data[key] = (children keys, P, W, N, L, done)` with
    key = np.array2string(observation) (mcts.py:103-110).  Here a tree is a slab of node slots
    (one simulation adds at most one node): the observation key is the vector of one-hot column
    indices (bytes [R]; equal observations <=> equal vectors), found through a 64-bit hash per
    slot and confirmed byte for byte.  Everything the reference does is kept, quirks included:
      * traverse (mcts.py:52-81) follows the STORED child key of the chosen action while the real
        cube is stepped next to it (`env.step`); the leaf is expanded from the real cube and stored
        under the real observation's key -- overwriting that entry if it exists (the lossy 3x3x3
        corner encoding can alias two cubes);
      * a node whose children have no visits picks a random action (mcts.py:69-70); the draws of
        tree b are those of `random.Random(seeds[b]).randint(0, A-1)` in order, i.e. what the
        reference draws after `random.seed(seeds[b])`;
      * otherwise argmax of U + W - L with U = cpuct * P * (sqrt(sum N) / (1 + N)), evaluated in
        float32 after every operation as the reference's expression does under NumPy >= 2
        (mcts.py:142-150), first maximum wins (mcts.py:152);
      * the way down adds `virtual_loss_const` to L, the way up removes the literal 150
        (mcts.py:77, 128), W = max(W, value) (mcts.py:124-126), N += 1;
      * `train` returns path + [first done child of the new leaf] (mcts.py:45-50); that tree stops.
    """

    def __init__(self, model, cube_size, num_sim=50, cpuct=1.0, virtual_loss_const=150, value_min=-10.0,
                 obs_dtype=torch.float32, model_device=None):
        self.model, self.cube_size, self.num_sim = model, cube_size, int(num_sim)
        self.cpuct, self.loss_const, self.value_min = float(cpuct), int(virtual_loss_const), float(value_min)
        self.obs_dtype, self.model_device = obs_dtype, model_device
        self.s, self.a = ops.N_STICKERS[cube_size], ops.N_ACTIONS[cube_size]
        self.r, self.c = ops.STATE_DIM[cube_size]

    # -- keys ---------------------------------------------------------------------------------------
    def _codes(self, onehot_u8):
        """one-hot uint8 [..., R, C] -> column indices uint8 [..., R] (the observation's key)."""
        return onehot_u8.argmax(dim=-1).to(torch.uint8)

    def _hash(self, codes):
        mult = self._mult
        return (codes.long() * mult).sum(dim=-1)

    @staticmethod
    def draw_table(seeds, action_dim, count):
        """[len(seeds), count] int64: the first `count` draws of random.Random(seed).randint(0, A-1)."""
        import random
        rows = []
        for sd in seeds:
            rng = random.Random(sd)
            rows.append([rng.randint(0, action_dim - 1) for _ in range(count)])
        return torch.tensor(rows, dtype=torch.int64)

    @torch.no_grad()
    def run(self, roots, seeds=None, rand_table=None):
        """roots: uint8 [B, S] sticker rows (CUDA).  seeds: per-tree seeds of Python's `random`
        (or rand_table: int64 [B, num_sim + 1] of pre-drawn actions).  Returns dict(solved bool [B],
        actions int64 [B, num_sim + 1] (-1 padded), n_actions, n_sims, n_nodes int64 [B],
        root_N int32 [B, A], root_W float32 [B, A], root_L int32 [B, A])."""
        dev = roots.device
        b, m, a, r = roots.shape[0], self.num_sim + 1, self.a, self.r
        mdev = dev if self.model_device is None else torch.device(self.model_device)
        if rand_table is None:
            if seeds is None:
                raise ValueError("BatchedMCTS.run needs `seeds` or `rand_table`")
            rand_table = self.draw_table(seeds, a, m)
        rand_table = rand_table.to(dev)
        gen = torch.Generator(device="cpu").manual_seed(0x5eed)
        self._mult = (torch.randint(1, 2 ** 62, (r,), generator=gen, dtype=torch.int64) | 1).to(dev)
        ar_b = torch.arange(b, device=dev)
        ar_a = torch.arange(a, device=dev)

        node_key = torch.zeros((b, m, r), dtype=torch.uint8, device=dev)
        node_hash = torch.zeros((b, m), dtype=torch.int64, device=dev)
        child_key = torch.zeros((b, m, a, r), dtype=torch.uint8, device=dev)
        child_done = torch.zeros((b, m, a), dtype=torch.bool, device=dev)
        P = torch.zeros((b, m, a), dtype=torch.float32, device=dev)
        W = torch.zeros((b, m, a), dtype=torch.float32, device=dev)
        N = torch.zeros((b, m, a), dtype=torch.int32, device=dev)
        L = torch.zeros((b, m, a), dtype=torch.int32, device=dev)
        n_nodes = torch.zeros(b, dtype=torch.int64, device=dev)
        rand_ptr = torch.zeros(b, dtype=torch.int64, device=dev)
        active = torch.ones(b, dtype=torch.bool, device=dev)
        actions_out = torch.full((b, m), -1, dtype=torch.int64, device=dev)
        n_actions = torch.zeros(b, dtype=torch.int64, device=dev)
        n_sims = torch.full((b,), self.num_sim, dtype=torch.int64, device=dev)
        slot_ids = torch.arange(m, device=dev)
        root_key = self._codes(ops.encode(self.cube_size, roots, dtype=torch.uint8))
        cpuct32 = torch.tensor(self.cpuct, dtype=torch.float32, device=dev)
        noop = torch.full((b,), 12, dtype=torch.int64, device=dev)

        def lookup(key, mask):
            """slot of `key` in each tree (or -1) for trees in `mask`."""
            h = self._hash(key)
            hit = (node_hash == h[:, None]) & (slot_ids[None, :] < n_nodes[:, None])
            cand = torch.where(hit, slot_ids[None, :], torch.full_like(hit, m, dtype=torch.int64)).min(dim=1).values
            ok = (cand < m) & mask
            safe = cand.clamp(max=m - 1)
            same = (node_key[ar_b, safe] == key).all(dim=-1)        # confirm byte for byte
            return torch.where(ok & same, safe, torch.full_like(safe, -1))

        for sim in range(self.num_sim):
            if not bool(active.any()):
                break
            # ---- traverse (mcts.py:52-81) ----
            cur_state = roots.clone()
            cur_key = root_key.clone()
            walking = active.clone()
            path_nodes, path_acts = [], []
            while True:
                slot = lookup(cur_key, walking)
                walking = walking & (slot >= 0)
                if not bool(walking.any()):
                    break
                node = slot.clamp(min=0)
                p_, w_, n_, l_ = P[ar_b, node], W[ar_b, node], N[ar_b, node], L[ar_b, node]
                total = n_.sum(dim=1)
                t = (total.double().sqrt()[:, None] / (1.0 + n_.double())).float()
                score = ((cpuct32 * p_) * t + w_) - l_.float()
                best = score.max(dim=1, keepdim=True).values
                a_best = torch.where(score == best, ar_a[None, :], torch.full_like(score, a, dtype=torch.int64)).min(dim=1).values
                use_rand = walking & (total == 0)
                a_rand = rand_table[ar_b, rand_ptr.clamp(max=m - 1)]
                rand_ptr = rand_ptr + use_rand.long()
                act = torch.where(use_rand, a_rand, a_best)
                path_nodes.append(torch.where(walking, node, torch.full_like(node, -1)))
                path_acts.append(act)
                wi = walking.nonzero(as_tuple=True)[0]
                L[wi, node[wi], act[wi]] += self.loss_const                              # mcts.py:77
                ops.step(self.cube_size, cur_state, torch.where(walking, act, noop).to(torch.uint8))
                cur_key = torch.where(walking[:, None], child_key[ar_b, node, act], cur_key)
            # ---- expand the leaves of the active trees (mcts.py:83-113) ----
            res = ops.expand(self.cube_size, cur_state, dtype=torch.uint8, want_child_onehot=True, want_parent_onehot=True)
            leaf_key = self._codes(res["parent_onehot"])
            x = res["parent_onehot"].to(mdev).to(self.obs_dtype)
            value, logits = self.model(x)
            value = value.float().reshape(-1).to(dev)
            policy = torch.nn.functional.softmax(logits.float(), dim=-1).to(dev)
            existing = lookup(leaf_key, active)                      # the real observation may alias an entry
            slot = torch.where(existing >= 0, existing, n_nodes.clamp(max=m - 1))
            ai = active.nonzero(as_tuple=True)[0]
            si = slot[ai]
            node_key[ai, si] = leaf_key[ai]
            node_hash[ai, si] = self._hash(leaf_key[ai])
            child_key[ai, si] = self._codes(res["child_onehot"][ai])
            child_done[ai, si] = res["solved"][ai].bool()
            P[ai, si] = policy[ai]
            W[ai, si] = self.value_min
            N[ai, si] = 0
            L[ai, si] = 0
            n_nodes = n_nodes + (active & (existing < 0)).long()
            # ---- back-propagate (mcts.py:115-130) ----
            for node, act in zip(path_nodes, path_acts):
                wi = (node >= 0).nonzero(as_tuple=True)[0]
                ni, ci = node[wi], act[wi]
                W[wi, ni, ci] = torch.maximum(W[wi, ni, ci], value[wi])
                L[wi, ni, ci] -= 150
                N[wi, ni, ci] += 1
            # ---- solved? (mcts.py:45-50) ----
            done_new = child_done[ar_b, slot] & active[:, None]
            first = torch.where(done_new, ar_a[None, :], torch.full_like(done_new, a, dtype=torch.int64)).min(dim=1).values
            solved_now = first < a
            depth = torch.zeros(b, dtype=torch.int64, device=dev)
            for node, act in zip(path_nodes, path_acts):
                on = (node >= 0) & solved_now
                oi = on.nonzero(as_tuple=True)[0]
                actions_out[oi, depth[oi]] = act[oi]
                depth = depth + on.long()
            oi = solved_now.nonzero(as_tuple=True)[0]
            actions_out[oi, depth[oi]] = first[oi]
            n_actions = torch.where(solved_now, depth + 1, n_actions)
            n_sims = torch.where(solved_now, torch.full_like(n_sims, sim + 1), n_sims)
            active = active & ~solved_now

        root_slot = lookup(root_key, torch.ones(b, dtype=torch.bool, device=dev)).clamp(min=0)
        return dict(solved=n_actions > 0, actions=actions_out, n_actions=n_actions, n_sims=n_sims, n_nodes=n_nodes,
                    root_N=N[ar_b, root_slot], root_W=W[ar_b, root_slot], root_L=L[ar_b, root_slot])

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
import ctypes

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


class _Tree(ctypes.Structure):
    """cube_mcts_tree_t of include/cube_b200.h."""
    _fields_ = [("n_trees", ctypes.c_int32), ("n_slots", ctypes.c_int32), ("path_cap", ctypes.c_int32),
                ("rand_cap", ctypes.c_int32)] + [
        (name, ctypes.c_void_p) for name in (
            "node_key", "child_key", "child_done", "P", "W", "N", "L", "n_nodes", "active", "root_state", "root_key",
            "rand_table", "rand_ptr", "path_node", "path_action", "path_len", "leaf_state", "flags",
            "child_slot", "sim_counter")]


class BatchedMCTS(object):
    """`MCTS` of the reference (mcts.py:17-154) for B cubes at once: every tree runs the same
    simulation index together, so one simulation is ONE traversal kernel (C ABI cube_mcts_traverse,
    a thread per tree), ONE leaf batch through `cube_expand` + the network, and ONE update kernel
    (cube_mcts_update: insert, back-propagate, solved test).

    Per tree the reference keeps `children_and_data[key] = (children keys, P, W, N, L, done)` with
    key = np.array2string(observation) (mcts.py:103-110).  Here a tree is a slab of node slots on the
    device (one simulation adds at most one node); the observation key is the vector of one-hot
    column indices (equal observations <=> equal vectors).  Everything the reference does is kept,
    quirks included:
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
                 obs_dtype=torch.float32, model_device=None, path_cap=None, graph=False):
        if not 1 <= int(num_sim) <= 254:
            raise ValueError("num_sim must be in 1..254 (node slots are addressed with one byte)")
        self.model, self.cube_size, self.num_sim = model, cube_size, int(num_sim)
        self.cpuct, self.loss_const, self.value_min = float(cpuct), int(virtual_loss_const), float(value_min)
        self.obs_dtype, self.model_device = obs_dtype, model_device
        self.graph = bool(graph)      # replay the simulations as ONE captured CUDA graph (the model must be capturable)
        self._ws = {}                 # cached tree store (+ graph) of the last batch shape
        self.s, self.a = ops.N_STICKERS[cube_size], ops.N_ACTIONS[cube_size]
        self.r, self.c = ops.STATE_DIM[cube_size]
        self.key_bytes = 20 if cube_size == 3 else 8
        self.path_cap = int(path_cap) if path_cap else 8 * (self.num_sim + 1)
        self.rand_cap = self.path_cap          # a node revisited inside one traversal draws again

    def _codes(self, onehot_u8):
        """one-hot uint8 [..., R, C] -> key bytes uint8 [..., KEY]: the column index of every row
        (2x2x2: 7 indices + one zero byte, so keys are whole words)."""
        codes = onehot_u8.argmax(dim=-1).to(torch.uint8)
        if self.key_bytes != self.r:
            pad = torch.zeros(codes.shape[:-1] + (self.key_bytes - self.r,), dtype=torch.uint8, device=codes.device)
            codes = torch.cat((codes, pad), dim=-1)
        return codes.contiguous()

    @staticmethod
    def draw_table(seeds, action_dim, count):
        """[len(seeds), count] uint8: the first `count` draws of random.Random(seed).randint(0, A-1)."""
        import random
        rows = []
        for sd in seeds:
            rng = random.Random(sd)
            rows.append([rng.randint(0, action_dim - 1) for _ in range(count)])
        return torch.tensor(rows, dtype=torch.uint8)

    def _workspace(self, b, dev, rand_cap):
        """The device tree store and the per-run bookkeeping for B trees, kept between runs: a search over a batch
        of the same shape re-arms a dozen small tensors instead of allocating and zero-filling ~8 KB per tree, and
        (graph=True) replays the simulation graph captured by the first run."""
        key = (b, str(dev), rand_cap)
        ws = self._ws.get(key)
        if ws is not None:
            return ws
        m, a, kb = self.num_sim + 1, self.a, self.key_bytes

        def z(shape, dtype):
            return torch.zeros(shape, dtype=dtype, device=dev)

        T = dict(node_key=z((b, m, kb), torch.uint8), child_key=z((b, m, a, kb), torch.uint8),
                 child_done=z((b, m, a), torch.uint8), P=z((b, m, a), torch.float32), W=z((b, m, a), torch.float32),
                 N=z((b, m, a), torch.int32), L=z((b, m, a), torch.int32), n_nodes=z((b,), torch.int32),
                 active=torch.ones(b, dtype=torch.uint8, device=dev), root_state=z((b, self.s), torch.uint8),
                 root_key=z((b, kb), torch.uint8), rand_table=z((b, rand_cap), torch.uint8),
                 rand_ptr=z((b,), torch.int32), path_node=z((b, self.path_cap), torch.uint8),
                 path_action=z((b, self.path_cap), torch.uint8), path_len=z((b,), torch.int32),
                 leaf_state=z((b, self.s), torch.uint8), flags=z((1,), torch.int32),
                 child_slot=torch.full((b, m, a), 255, dtype=torch.uint8, device=dev),
                 sim_counter=torch.full((1,), -1, dtype=torch.int32, device=dev))
        ws = dict(T=T, tree=_Tree(b, m, self.path_cap, rand_cap, *[T[name].data_ptr() for name, _ in _Tree._fields_[4:]]),
                  actions=torch.full((b, self.path_cap + 1), -1, dtype=torch.int8, device=dev), n_actions=z((b,), torch.int32),
                  n_sims=torch.full((b,), self.num_sim, dtype=torch.int32, device=dev), still=z((self.num_sim,), torch.int32),
                  still_host=torch.full((self.num_sim,), -1, dtype=torch.int32).pin_memory(), graph=None, keep=None)
        self._ws.clear()                                   # one batch shape at a time: a tree store is ~8 KB per tree
        self._ws[key] = ws
        return ws

    def release(self):
        """Drop the cached tree store (and the captured graph)."""
        self._ws.clear()

    @torch.no_grad()
    def run(self, roots, seeds=None, rand_table=None, timers=None):
        """roots: uint8 [B, S] sticker rows (CUDA).  seeds: per-tree seeds of Python's `random`
        (or rand_table: [B, draws] pre-drawn actions < A, `rand_cap` = 8 * (num_sim + 1) by default).  Returns dict(solved bool [B],
        actions int64 [B, max(num_sim + 1, longest list)] (-1 padded), n_actions, n_sims, n_nodes int64 [B],
        root_N int32 [B, A], root_W float32 [B, A], root_L int32 [B, A])."""
        from . import _lib
        lib = _lib.load()
        roots = roots.contiguous()
        dev = roots.device
        b, m, a = roots.shape[0], self.num_sim + 1, self.a
        mdev = dev if self.model_device is None else torch.device(self.model_device)
        if rand_table is None:
            if seeds is None:
                raise ValueError("BatchedMCTS.run needs `seeds` or `rand_table`")
            rand_table = self.draw_table(seeds, a, self.rand_cap)
        rand_table = rand_table.to(device=dev, dtype=torch.uint8).contiguous()
        if rand_table.dim() != 2 or rand_table.shape[0] != b:
            raise ValueError("rand_table must be [B, draws per tree]")
        if rand_table.numel() and int(rand_table.max()) >= a:
            raise IndexError("rand_table holds action indices >= %d" % a)           # they index the node's child rows
        ws = self._workspace(b, dev, rand_table.shape[1])
        T, tree, actions, n_actions, n_sims = ws["T"], ws["tree"], ws["actions"], ws["n_actions"], ws["n_sims"]
        still, still_host = ws["still"], ws["still_host"]
        # re-arm: only what a search reads before it writes (node rows are initialised when a node is stored)
        T["root_state"].copy_(roots)
        T["root_key"].copy_(self._codes(ops.encode(self.cube_size, roots, dtype=torch.uint8)))
        T["rand_table"].copy_(rand_table)
        for name in ("n_nodes", "rand_ptr", "path_len", "flags"):
            T[name].zero_()
        T["active"].fill_(1)
        T["sim_counter"].fill_(-1)
        actions.fill_(-1)
        n_actions.zero_()
        n_sims.fill_(self.num_sim)
        still.zero_()
        still_host.fill_(-1)
        from ._timing import Phases
        phase = Phases(timers, dev)
        # Early exit without stalling the launch queue: every update kernel counts the trees that are still
        # searching into still[sim]; the counts travel to pinned host memory behind every simulation and are
        # looked at (never waited for) a few simulations later.
        direct = self.obs_dtype in ops.ONEHOT_DTYPES

        def simulate():
            """One simulation of every tree on the current stream; the simulation index lives on the device."""
            st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            with phase("traverse_ms"):
                _lib.check(lib.cube_mcts_traverse(self.cube_size, ctypes.byref(tree), self.cpuct, self.loss_const, st),
                           "cube_mcts_traverse")
            # the node keys come out of the expansion kernel as compact codes (cube_expand_codes), the leaf's
            # observation directly in the net's dtype: no one-hot rows of the children, no argmax passes
            with phase("expand_ms"):
                res = ops.expand_codes(self.cube_size, T["leaf_state"], parent_dtype=self.obs_dtype if direct else torch.uint8,
                                       want_reward=False)
            with phase("net_ms"):
                value, logits = self.model(res["parent_onehot"].to(mdev).to(self.obs_dtype))
            with phase("glue_ms"):
                value = value.float().reshape(-1).to(dev).contiguous()
                policy = torch.nn.functional.softmax(logits.float(), dim=-1).to(dev).contiguous()
            with phase("update_ms"):
                _lib.check(lib.cube_mcts_update(self.cube_size, ctypes.byref(tree), res["parent_codes"].data_ptr(),
                                                res["child_codes"].data_ptr(), res["solved"].data_ptr(), value.data_ptr(),
                                                policy.data_ptr(), self.value_min, 0, actions.data_ptr(), n_actions.data_ptr(),
                                                n_sims.data_ptr(), still.data_ptr(), st),
                           "cube_mcts_update")
            still_host.copy_(still, non_blocking=True)
            return res, value, policy                                    # kept alive by the caller (graph replays reuse them)

        with torch.cuda.device(dev):
            landed, sim = [], 0
            use_graph = self.graph and timers is None
            while sim < self.num_sim:
                done = False
                while landed and landed[0][1].query():
                    if int(landed.pop(0)[2]) == 0:
                        done = True
                if done:
                    break
                if use_graph and ws["graph"] is None and sim == 1:
                    # the very first simulation ran eagerly (library warm-up); from here on -- and in every later
                    # run over a batch of this shape -- one captured simulation is replayed
                    ws["graph"] = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(ws["graph"]):
                        ws["keep"] = simulate()
                if use_graph and ws["graph"] is not None:
                    ws["graph"].replay()
                else:
                    simulate()
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream(dev))
                landed.append((sim, ev, still_host[sim:sim + 1]))
                sim += 1
        phase.finish()
        flags = int(T["flags"].item())
        if flags:
            raise RuntimeError("BatchedMCTS: capacity exceeded (flags=%d: 1 path_cap, 2 rand_table, 4 node slots, "
                               "8 rand_table action out of range)" % flags)
        # the root's statistics, for inspection / tests
        ar = torch.arange(b, device=dev)
        same = (T["node_key"] == T["root_key"][:, None, :]).all(dim=-1) & (torch.arange(m, device=dev)[None, :] < T["n_nodes"][:, None])
        root_slot = same.int().argmax(dim=1)
        # a traversal may revisit nodes (U then U' is back at the root's key), so a returned path can be longer
        # than the node count: never cut an action list (the reference returns actions_to_leaf + [i] whole)
        width = max(m, int(n_actions.max())) if b else m
        return dict(solved=n_actions > 0, actions=actions[:, :width].long(), n_actions=n_actions.long(), n_sims=n_sims.long(),
                    n_nodes=T["n_nodes"].long(), root_N=T["N"][ar, root_slot], root_W=T["W"][ar, root_slot],
                    root_L=T["L"][ar, root_slot])

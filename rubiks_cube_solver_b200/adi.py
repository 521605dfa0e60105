"""Batched ADI sample generation: the reference's get_random_samples / get_target_value
(gym-cube/gym_cube/envs/cube_env.py:177-252) for many cubes at once.

The reference emits one sample after EVERY move of every scramble (cube_env.py:190-194),
so the parents of an ADI batch are all scramble prefixes.  Here the prefixes are produced
step by step on the device (K2, one launch per depth level, step-major so each level is a
contiguous slab), expanded into all children + their one-hot rows by K3, evaluated by the
caller's network in one batch, and reduced to (target_value, target_policy, error) by kernel K4
(C ABI cube_adi_targets) with the reference's rules: first solved child a -> (1.0, a) (cube_env.py:217-220), otherwise
max_a(V(child_a) + (-1.0)) with the first maximum winning (cube_env.py:240-246), and
error = |V(state) - target| * depth**(-temperature) in float64 (cube_env.py:247-251).
"""
import torch

from . import ops


def scramble_prefixes(cube_size, moves):
    """All prefix states of every scramble.

    moves [n, depth] uint8 (CUDA).  Returns (trail [depth, n_pad, S] uint8 step-major,
    n_pad) where n_pad rounds n up to a multiple of 16 with no-op rows, so every level is
    a 16-byte aligned slab.
    """
    n, depth = moves.shape
    dev = moves.device
    n_pad = (n + 15) // 16 * 16
    s = ops.N_STICKERS[cube_size]
    moves_t = torch.full((depth, n_pad), 12, dtype=torch.uint8, device=dev)      # 12 = no-op row
    moves_t[:, :n] = moves.t()
    trail = torch.empty((depth, n_pad, s), dtype=torch.uint8, device=dev)
    prev = ops.solved_states(cube_size, n_pad, dev)
    for k in range(depth):
        ops.walk(cube_size, prev, moves_t[k], out=trail[k])
        prev = trail[k]
    return trail, n_pad


def assemble_targets(cube_size, child_values, solved, parent_values, depth_of_row, temperature):
    """cube_env.py:239-252 for a batch, on the device (kernel K4, C ABI cube_adi_targets).
    child_values [P, A] float32 = V(child); solved [P, A]; parent_values [P] float32; depth_of_row [P]
    (scramble counts).  Returns (target_value float32 [P], target_policy int64 [P], error float64 [P])."""
    return ops.adi_targets(cube_size, child_values, solved, parent_values, depth_of_row, temperature)


@torch.no_grad()
def generate_samples(cube_size, moves, model, temperature, model_device=None, onehot_dtype=torch.float32,
                     forward_chunk=1 << 16):
    """ADI samples for every prefix of every scramble in `moves` ([n, depth] uint8, CUDA).

    `model(x)` must return (value [B,1], policy) like DeepCube.forward (model.py:31-45).
    Results are cube-major (cube 0 depth 1..d, cube 1 ...), the order in which the
    reference appends to its replay buffer.
    """
    n, depth = moves.shape
    dev = moves.device
    a = ops.N_ACTIONS[cube_size]
    r, c = ops.STATE_DIM[cube_size]
    trail, n_pad = scramble_prefixes(cube_size, moves)
    parents = trail.view(depth * n_pad, -1)
    res = ops.expand(cube_size, parents, dtype=onehot_dtype, want_parent_onehot=True)
    p = parents.shape[0]
    mdev = dev if model_device is None else torch.device(model_device)

    def values_of(x):
        out = []
        for i in range(0, x.shape[0], forward_chunk):
            v, _ = model(x[i:i + forward_chunk].to(mdev).float())
            out.append(v.reshape(-1).to(dev))
        return torch.cat(out)

    child_v = values_of(res["child_onehot"].view(p * a, r, c)).view(p, a)
    parent_v = values_of(res["parent_onehot"])
    depth_of_row = torch.arange(1, depth + 1, device=dev).repeat_interleave(n_pad)
    tv, tp, err = assemble_targets(cube_size, child_v, res["solved"], parent_v, depth_of_row, temperature)

    # step-major (k, cube) -> cube-major (cube, k), dropping the padding cubes
    def cube_major(t):
        return t.view(depth, n_pad, *t.shape[1:])[:, :n].transpose(0, 1).reshape(n * depth, *t.shape[1:])

    state_u8 = cube_major(ops.encode(cube_size, parents, dtype=torch.uint8))
    return dict(state_u8=state_u8, stickers=cube_major(parents), target_value=cube_major(tv),
                target_policy=cube_major(tp), error=cube_major(err),
                scramble_count=cube_major(depth_of_row), final_stickers=trail[depth - 1, :n],
                child_solved=cube_major(res["solved"]))

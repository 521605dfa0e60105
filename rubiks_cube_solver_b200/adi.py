"""Batched ADI sample generation: the reference's get_random_samples / get_target_value
(gym-cube/gym_cube/envs/cube_env.py:177-252) for many cubes at once.

The reference emits one sample after EVERY move of every scramble (cube_env.py:190-194),
so the parents of an ADI batch are all scramble prefixes, cube by cube.  Here

* ONE launch writes every prefix of every scramble, already cube-major (C ABI
  cube_scramble_prefixes) -- round 1 issued one cube_walk launch per depth level and then
  transposed five tensors from step-major to cube-major;
* the parents are expanded chunk by chunk into all children + their one-hot rows (K3,
  cube_expand) IN THE DTYPE THE NET COMPUTES IN (a bf16 DeepCube reads K3's buffer as it is, no
  `.float()` pass), and every chunk is evaluated by the caller's network right away, while its
  one-hot rows are still in the 126 MB L2: the [P, A, D] batch (48 GB in bf16 for 4 Mi parents)
  never exists as a whole and never round-trips through HBM;
* kernel K4 (C ABI cube_adi_targets) reduces child values to (target_value, target_policy,
  error) with the reference's rules: first solved child a -> (1.0, a) (cube_env.py:217-220),
  otherwise max_a(V(child_a) + (-1.0)) with the first maximum winning (cube_env.py:240-246), and
  error = |V(state) - target| * depth**(-temperature) in float64 (cube_env.py:247-251).
"""
import torch

from . import ops
from ._timing import Phases as _Phases


DEFAULT_FORWARD_CHUNK = 1 << 17      # parents per expand -> forward chunk (measured: tools/adi_iteration.py)


def scramble_prefixes(cube_size, moves):
    """All prefix states of every scramble, cube-major: [n, depth, S] uint8 with
    out[i, k] = state of cube i after moves[i, 0..k] (cube_env.py:187-191).

    One launch (cube_scramble_prefixes) up to depth 526 (3x3x3) / 1159 (2x2x2); deeper scrambles
    fall back to one cube_walk launch per level."""
    n, depth = moves.shape
    dev = moves.device
    s = ops.N_STICKERS[cube_size]
    if depth <= ops.prefixes_max_depth(cube_size):
        return ops.scramble_prefixes(cube_size, moves.contiguous())[0]
    n_pad = (n + 15) // 16 * 16                      # every level a 16-byte aligned slab
    moves_t = torch.full((depth, n_pad), 12, dtype=torch.uint8, device=dev)      # 12 = no-op row
    moves_t[:, :n] = moves.t()
    trail = torch.empty((depth, n_pad, s), dtype=torch.uint8, device=dev)
    prev = ops.solved_states(cube_size, n_pad, dev)
    for k in range(depth):
        ops.walk(cube_size, prev, moves_t[k], out=trail[k])
        prev = trail[k]
    return trail[:, :n].transpose(0, 1).contiguous()


def assemble_targets(cube_size, child_values, solved, parent_values, depth_of_row, temperature):
    """cube_env.py:239-252 for a batch, on the device (kernel K4, C ABI cube_adi_targets).
    child_values [P, A] float32 = V(child); solved [P, A]; parent_values [P] float32; depth_of_row [P]
    (scramble counts).  Returns (target_value float32 [P], target_policy int64 [P], error float64 [P])."""
    return ops.adi_targets(cube_size, child_values, solved, parent_values, depth_of_row, temperature)


def param_dtype(model):
    """The dtype the net computes in: that of its first floating-point parameter (None: no parameters)."""
    for p in model.parameters() if hasattr(model, "parameters") else ():
        if p.is_floating_point():
            return p.dtype
    return None


def model_dtype(model, default=torch.float32):
    """The one-hot dtype to write for `model`: its own dtype when K3 can write it (bf16 / f32), else `default`."""
    d = param_dtype(model)
    return d if d in ops.ONEHOT_DTYPES else default


@torch.no_grad()
def generate_samples(cube_size, moves, model, temperature, model_device=None, onehot_dtype=None,
                     forward_chunk=DEFAULT_FORWARD_CHUNK, timers=None):
    """ADI samples for every prefix of every scramble in `moves` ([n, depth] uint8, CUDA).

    `model(x)` must return (value [B,1], policy) like DeepCube.forward (model.py:31-45).  `onehot_dtype`
    defaults to the dtype the model computes in (bf16 / f32), so the net reads K3's buffer directly.
    `forward_chunk` = parents per expand -> forward chunk (x A children x D elements).  Measured on B200 with
    DeepCube [1024, 256, 128] in bf16 (tools/adi_iteration.py): the net, not HBM, bounds the iteration (218 ms of
    227 ms for 4 Mi parents), chunks small enough to stay in the 126 MB L2 (8192 parents = 94 MB) lose more to
    launch overhead than they save, so the default is 131 072 parents (1.5 GB of bf16 one-hot rows per chunk);
    what chunking does buy is that the [P, A, D] batch (48 GB) never exists as a whole.  Results are cube-major (cube 0 depth 1..d, cube 1
    ...), the order in which the reference appends to its replay buffer.  `timers`: an optional dict that
    receives the milliseconds spent per phase (prefixes / expand / net / targets), device-timed.
    """
    n, depth = moves.shape
    dev = moves.device
    a = ops.N_ACTIONS[cube_size]
    r, c = ops.STATE_DIM[cube_size]
    if onehot_dtype is None:
        onehot_dtype = model_dtype(model)
    mdev = dev if model_device is None else torch.device(model_device)
    in_dtype = param_dtype(model)                                         # cast only when the net needs it
    phase = _Phases(timers, dev)

    def value_of(x):
        x = x.to(mdev)
        if in_dtype is not None and x.dtype != in_dtype:
            x = x.to(in_dtype)
        v, _ = model(x)
        return v.reshape(-1).float().to(dev)

    with phase("prefixes_ms"):
        trail = scramble_prefixes(cube_size, moves)                       # [n, depth, S]
    parents = trail.view(n * depth, -1)
    p = parents.shape[0]
    chunk = max(16, int(forward_chunk) // 16 * 16)                        # slices stay 16-byte aligned
    state = torch.empty((p, r, c), dtype=onehot_dtype, device=dev)        # the samples' `state`: the parents' one-hot
    child_solved = torch.empty((p, a), dtype=torch.uint8, device=dev)
    child_v = torch.empty((p, a), dtype=torch.float32, device=dev)
    parent_v = torch.empty(p, dtype=torch.float32, device=dev)
    buf = torch.empty((min(chunk, p), a, r, c), dtype=onehot_dtype, device=dev)
    for i in range(0, p, chunk):
        cnt = min(chunk, p - i)
        with phase("expand_ms"):
            ops.expand(cube_size, parents[i:i + cnt], dtype=onehot_dtype, child_onehot=buf[:cnt],
                       parent_onehot=state[i:i + cnt], solved=child_solved[i:i + cnt], want_reward=False)
        with phase("net_ms"):
            child_v[i:i + cnt] = value_of(buf[:cnt].view(cnt * a, r, c)).view(cnt, a)
            parent_v[i:i + cnt] = value_of(state[i:i + cnt])
    with phase("targets_ms"):
        depth_of_row = torch.arange(1, depth + 1, device=dev, dtype=torch.int32).repeat(n)
        tv, tp, err = assemble_targets(cube_size, child_v, child_solved, parent_v, depth_of_row, temperature)
    phase.finish()
    return dict(state=state, stickers=parents, target_value=tv, target_policy=tp, error=err,
                scramble_count=depth_of_row.long(), final_stickers=trail[:, depth - 1], child_solved=child_solved,
                child_values=child_v, parent_values=parent_v)

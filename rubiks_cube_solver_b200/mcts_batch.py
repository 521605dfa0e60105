"""Batched MCTS leaf expansion: what `MCTS.expand` (mcts.py:83-113) computes for ONE leaf --
`model.predict(leaf)` (value + softmax policy, model.py:78-91), the A children reached by
`env.step(i)` on copies of the env and their `done` flags (mcts.py:96-101) -- for a batch of
leaves from many trees at once.  The tree itself (dict keyed by np.array2string, max-backup,
virtual loss; mcts.py:17-154) stays with the caller and is out of scope here.

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

"""CPU restatement of the reference's MCTS (mcts.py:17-154) -- TEST INFRASTRUCTURE ONLY.

One tree per cube, exactly as the reference keeps it: a dict keyed by the one-hot observation
(the reference keys by ``np.array2string(obs)``; the observation's bytes are the same key, the
arrays are far below numpy's summarisation threshold), whose entries are
(children keys, policy P, value W, visit count N, virtual loss L, done flags) (mcts.py:23-28,
103-110).  ``train`` = traverse (mcts.py:52-81) -> expand (mcts.py:83-113) -> backpropagate
(mcts.py:115-130) and returns the action list when a child of the new leaf is solved
(mcts.py:45-50).

Arithmetic follows what the reference's expressions evaluate to under NumPy >= 2 (the version
the golden vectors in tests/golden/mcts_*.npz were generated with, by running the reference's own
mcts.py -- oracle/gen_golden.py): P and the network value are float32, Python floats / ints are
"weak", so U = cpuct * P * (sqrt(sum N) / (1 + N)) and U + W - L are rounded to float32 after
every operation (mcts.py:142-150), W = max(W, value) keeps float32 values (mcts.py:124-126), and
``max(range(A), key=...)`` takes the first maximum (mcts.py:152).  The quirks are kept: the
virtual loss added on the way down is ``virtual_loss_const`` but the amount removed on the way
up is the literal 150 (mcts.py:77, 128); a node whose children have no visits yet picks a uniformly
random action from Python's ``random`` (mcts.py:69-70) -- here from the ``rng`` handed in, a
``random.Random`` (``random.seed(s)`` + the module functions draw the same numbers).
"""
import copy
import math

import numpy as np


class MCTSRef(object):
    def __init__(self, predict, action_dim, loss_constant=150, cpuct=1.0, value_min=-10.0, rng=None):
        self.predict = predict                  # obs -> (value np.float32 array [1], policy np.float32 [A])
        self.action_dim = action_dim
        self.loss_constant = loss_constant
        self.cpuct = cpuct
        self.value_min = value_min
        self.rng = rng
        self.nodes = dict()

    @staticmethod
    def key(obs):
        return np.ascontiguousarray(obs).tobytes()

    def train(self, state, env):
        sim_env = copy.deepcopy(env)                                         # mcts.py:37
        path, actions, leaf = self.traverse(state, sim_env)
        value = self.expand(leaf, sim_env)
        self.backpropagate(path, actions, value)
        done = self.nodes[self.key(leaf)][5]
        for i in range(self.action_dim):                                      # mcts.py:45-50
            if done[i]:
                actions.append(i)
                return actions
        return None

    def traverse(self, state, env):
        path, actions = [], []
        cur_arr, cur = state, self.key(state)
        while True:
            if cur not in self.nodes:                                         # mcts.py:66-67
                return path, actions, cur_arr
            children, P, W, N, L, _ = self.nodes[cur]
            if sum(N) == 0:
                a = self.rng.randint(0, self.action_dim - 1)                 # mcts.py:69-70
            else:
                a = self.most_promising(cur)
            path.append(cur)
            actions.append(a)
            L[a] += self.loss_constant                                        # mcts.py:77
            cur_arr, _, _, _ = env.step(a)
            cur = children[a]

    def expand(self, state, env):
        value, policy = self.predict(state)                                   # mcts.py:92
        children, done = [], []
        for a in range(self.action_dim):                                      # mcts.py:96-101
            child = copy.deepcopy(env)
            obs, _, d, _ = child.step(a)
            children.append(self.key(obs))
            done.append(bool(d))
        self.nodes[self.key(state)] = (children, policy, [np.float32(self.value_min)] * self.action_dim,
                                       [0] * self.action_dim, [0] * self.action_dim, done)
        return np.float32(np.asarray(value).reshape(-1)[0])

    def backpropagate(self, path, actions, value):
        for node, a in zip(reversed(path), reversed(actions)):                # mcts.py:122-129
            _, _, W, N, L, _ = self.nodes[node]
            W[a] = max(W[a], value)
            L[a] -= 150
            N[a] += 1

    def most_promising(self, node):
        _, P, W, N, L, _ = self.nodes[node]
        total = sum(N)
        best, best_score = 0, None
        for i in range(self.action_dim):                                      # mcts.py:142-150, float32 after every op
            t = np.float32(math.sqrt(total) / (1 + N[i]))
            u = np.float32(np.float32(np.float32(self.cpuct) * np.float32(P[i])) * t)
            score = np.float32(np.float32(u + np.float32(W[i])) - np.float32(L[i]))
            if best_score is None or score > best_score:                      # first maximum wins (mcts.py:152)
                best, best_score = i, score
        return best


def solve(predict, env, obs, num_sim, rng, **kw):
    """The MCTS branch of test.py's trial for ONE time step (test.py:139-147): up to num_sim
    simulations from the root.  Returns (action list or None, simulations run, tree)."""
    tree = MCTSRef(predict, env.action_dim, rng=rng, **kw)
    for k in range(num_sim):
        actions = tree.train(obs, env)
        if actions is not None:
            return actions, k + 1, tree
    return None, num_sim, tree

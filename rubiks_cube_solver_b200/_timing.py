"""CUDA-event stopwatch for the phases of a multi-launch host loop (ADI iteration, MCTS search)."""
import torch


class Phases(object):
    """CUDA-event stopwatch for the phases of one generate_samples call (bench.py's ADI iteration)."""

    def __init__(self, timers, device):
        self.timers, self.device, self.pairs = timers, device, []

    def __call__(self, name):
        return _Phase(self, name)

    def finish(self):
        if self.timers is None:
            return
        torch.cuda.synchronize(self.device)
        for name, e0, e1 in self.pairs:
            self.timers[name] = self.timers.get(name, 0.0) + e0.elapsed_time(e1)


class _Phase(object):
    def __init__(self, owner, name):
        self.o, self.name = owner, name

    def __enter__(self):
        if self.o.timers is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record(torch.cuda.current_stream(self.o.device))

    def __exit__(self, *exc):
        if self.o.timers is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record(torch.cuda.current_stream(self.o.device))
            self.o.pairs.append((self.name, self.e0, e1))

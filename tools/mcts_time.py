"""Time bench.py's batched-MCTS entry alone:  python tools/mcts_time.py"""
import sys, json
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import torch, bench
print(json.dumps(bench.measure_mcts(torch, torch.device('cuda', 0)), indent=1))

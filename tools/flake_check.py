"""Runs one GPU test function many times in one process (flakiness check).  usage: flake_check.py <test name> [n]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests import test_gpu_parity as T
name, n = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 20
dev = torch.device('cuda', 0)
bad = 0
for i in range(n):
  try:
    getattr(T, name)(dev)
  except AssertionError as e:
    bad += 1
    print('FAIL', i, str(e)[:300])
print(f'{name}: {n - bad}/{n} passed')

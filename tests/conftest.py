import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)


def pytest_configure(config):
  config.addinivalue_line('markers', 'gpu: needs a CUDA device (run with -m gpu on the B200 box)')


@pytest.fixture(scope='session')
def cuda_device():
  import torch
  if not torch.cuda.is_available():
    pytest.skip('no CUDA device')
  from nerfds_b200 import build
  build.build()          # no-op when lib/BUILD_INFO.json matches the sources; compiles (nvcc, sm_100a) otherwise
  return torch.device('cuda', 0)

"""TrainState stand-in: the two fields of the reference's state that the path
reads (hypernerf/model_utils.py:28-52; evaluation.py:119-120)."""
from __future__ import annotations

from dataclasses import dataclass, field
from types import SimpleNamespace
from typing import Any, Dict, Optional


@dataclass
class TrainState:
  """``state.optimizer.target['model']`` and ``state.extra_params``."""
  optimizer: Any
  nerf_alpha: Optional[float] = None
  warp_alpha: Optional[float] = None
  hyper_alpha: Optional[float] = None
  hyper_sheet_alpha: Optional[float] = None
  norm_loss_weight: Optional[float] = None
  norm_input_alpha: Optional[float] = None
  norm_voxel_lr: Optional[float] = None
  norm_voxel_ratio: Optional[float] = None

  @property
  def extra_params(self) -> Dict[str, Any]:      # model_utils.py:41-52
    return {
        'nerf_alpha': self.nerf_alpha, 'warp_alpha': self.warp_alpha,
        'hyper_alpha': self.hyper_alpha, 'hyper_sheet_alpha': self.hyper_sheet_alpha,
        'norm_loss_weight': self.norm_loss_weight, 'norm_input_alpha': self.norm_input_alpha,
        'norm_voxel_lr': self.norm_voxel_lr, 'norm_voxel_ratio': self.norm_voxel_ratio,
    }

  @classmethod
  def create(cls, params: Dict, extra_params: Dict[str, float]) -> 'TrainState':
    opt = SimpleNamespace(target={'model': params}, state=SimpleNamespace(step=0))
    return cls(optimizer=opt, **{k: extra_params.get(k) for k in (
        'nerf_alpha', 'warp_alpha', 'hyper_alpha', 'hyper_sheet_alpha', 'norm_loss_weight',
        'norm_input_alpha', 'norm_voxel_lr', 'norm_voxel_ratio')})

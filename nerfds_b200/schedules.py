"""Scalar schedules of the training/eval drivers (hypernerf/schedules.py).

The ray-marching path consumes five scheduled scalars (`nerf_alpha`,
`warp_alpha`, `hyper_alpha`, `hyper_sheet_alpha`, `norm_input_alpha`,
model_utils.py:42-52).  A restored checkpoint carries their values at the step
it was written; this module evaluates the gin-configured schedules at any
other step (train.py:278-290 builds them with `schedules.from_config`).
Host-side float math only -- nothing here runs per ray.

`from_config` accepts what gin hands the reference: None, a `(type, *args)`
tuple/list, or a `{'type': ..., **kwargs}` mapping (schedules.py:26-48).
"""
from __future__ import annotations

import math
from collections.abc import Mapping
from typing import Any, Callable, Dict, Optional, Sequence


class Schedule:
  def get(self, step):
    raise NotImplementedError

  def __call__(self, step):
    return self.get(step)


class NoneSchedule(Schedule):
  """schedules.py:63-67 -- a disabled scalar."""

  def get(self, step):
    return None


class ConstantSchedule(Schedule):
  def __init__(self, value):
    self.value = value

  def get(self, step):
    if self.value is None:          # schedules.py:79-81: ('constant', None) disables the scalar
      return None
    return float(self.value)


class LinearSchedule(Schedule):
  """schedules.py:84-98: lerp over num_steps then hold; num_steps == 0 -> final."""

  def __init__(self, initial_value, final_value, num_steps):
    self.initial_value, self.final_value, self.num_steps = initial_value, final_value, num_steps

  def get(self, step):
    if self.num_steps == 0:
      return float(self.final_value)
    a = min(step / self.num_steps, 1.0)
    return (1.0 - a) * self.initial_value + a * self.final_value


class ExponentialSchedule(Schedule):
  """schedules.py:101-124: geometric decay reaching final_value at step num_steps-1."""

  def __init__(self, initial_value, final_value, num_steps, eps=1e-10):
    if initial_value <= final_value:
      raise ValueError('Final value must be less than initial value.')
    self.initial_value, self.final_value, self.num_steps, self.eps = initial_value, final_value, num_steps, eps

  def get(self, step):
    if step >= self.num_steps:
      return float(self.final_value)
    base = max(self.final_value, self.eps) / self.initial_value
    return self.initial_value * base ** (step / (self.num_steps - 1))


class CosineEasingSchedule(Schedule):
  """schedules.py:127-142."""

  def __init__(self, initial_value, final_value, num_steps):
    self.initial_value, self.final_value, self.num_steps = initial_value, final_value, num_steps

  def get(self, step):
    x = min(max(min(step / self.num_steps, 1.0), 0.0), 1.0)
    return self.initial_value + (self.final_value - self.initial_value) * 0.5 * (1 + math.cos(math.pi * x + math.pi))


class StepSchedule(Schedule):
  """schedules.py:145-169: staircase decay."""

  def __init__(self, initial_value, decay_interval, decay_factor, max_decays, final_value=None):
    self.initial_value, self.decay_interval = initial_value, decay_interval
    self.decay_factor, self.max_decays = decay_factor, max_decays
    self.final_value = initial_value * decay_factor ** max_decays if final_value is None else final_value

  def get(self, step):
    phase = step // self.decay_interval
    return self.final_value if phase >= self.max_decays else self.initial_value * self.decay_factor ** phase


class PiecewiseSchedule(Schedule):
  """schedules.py:172-185: [(length, schedule), ...]; the last piece runs forever.

  Piece k starts at the cumulative length of the pieces before it and sees
  steps re-based to its own start; a step on a boundary belongs to the later
  piece (searchsorted side='right').
  """

  def __init__(self, schedules: Sequence):
    self.schedules = [from_config(s) for _, s in schedules]
    acc, self.milestones = 0, []
    for length, _ in schedules:
      acc += length
      self.milestones.append(acc)
    self.milestones = self.milestones[:-1]

  def get(self, step):
    idx = sum(1 for m in self.milestones if m <= step)
    base = self.milestones[idx - 1] if idx >= 1 else 0
    return self.schedules[idx].get(step - base)


class DelayedSchedule(Schedule):
  """schedules.py:188-201: sine ramp multiplier over the first delay_steps."""

  def __init__(self, base_schedule, delay_steps, delay_mult):
    self.base_schedule, self.delay_steps, self.delay_mult = from_config(base_schedule), delay_steps, delay_mult

  def get(self, step):
    t = min(max(step / self.delay_steps, 0.0), 1.0)
    rate = self.delay_mult + (1 - self.delay_mult) * math.sin(0.5 * math.pi * t)
    return rate * self.base_schedule(step)


SCHEDULE_MAP: Dict[str, Callable[..., Schedule]] = {
    'constant': ConstantSchedule,
    'linear': LinearSchedule,
    'exponential': ExponentialSchedule,
    'cosine_easing': CosineEasingSchedule,
    'step': StepSchedule,
    'piecewise': PiecewiseSchedule,
    'delayed': DelayedSchedule,
}


def from_config(schedule: Any) -> Schedule:
  if schedule is None:
    return NoneSchedule()
  if isinstance(schedule, Schedule):
    return schedule
  if isinstance(schedule, (tuple, list)):
    kind, *args = schedule
    return SCHEDULE_MAP[kind](*args)
  if isinstance(schedule, Mapping):
    d = dict(schedule)
    return SCHEDULE_MAP[d.pop('type')](**d)
  raise ValueError(f'Unknown type {type(schedule)}.')


# TrainConfig / SpecularConfig attribute -> extra_params key (train.py:278-290, 317-328)
PATH_SCHEDULES = {
    'TrainConfig.nerf_alpha_schedule': 'nerf_alpha',
    'TrainConfig.warp_alpha_schedule': 'warp_alpha',
    'TrainConfig.hyper_alpha_schedule': 'hyper_alpha',
    'TrainConfig.hyper_sheet_alpha_schedule': 'hyper_sheet_alpha',
    'SpecularConfig.norm_input_alpha_schedule': 'norm_input_alpha',
}


# class defaults of the schedule attributes (configs.py:65-68, 219-225): used when the gin file does not bind them
SCHEDULE_DEFAULTS = {'SpecularConfig.norm_input_alpha_schedule': {'type': 'constant', 'value': 4}}


def extra_params_at(bindings: Mapping, step: int) -> Dict[str, Optional[float]]:
  """The scheduled scalars of `TrainState.extra_params` at `step`, from parsed gin bindings."""
  return {name: from_config(bindings.get(key, SCHEDULE_DEFAULTS.get(key))).get(step)
          for key, name in PATH_SCHEDULES.items()}

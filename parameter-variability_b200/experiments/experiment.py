"""Light-weight stand-ins for the reference's experiment schema objects that the hot path reads
(reference src/parvar/experiments/experiment.py:17-57).  The reference's pydantic objects are
accepted everywhere these are (duck typing on ``.add_noise/.cv/.type`` and ``.id``)."""
from __future__ import annotations

from dataclasses import dataclass
from enum import Enum


class DistributionType(str, Enum):
    NORMAL = "normal"
    LOGNORMAL = "lognormal"


@dataclass
class Noise:
    add_noise: bool
    cv: float
    type: DistributionType = DistributionType.NORMAL


@dataclass
class Observable:
    id: str
    starting_value: float = 0.0

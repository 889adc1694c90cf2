"""The knobs of the reference's global `CUSTOM_HYP` that the hot path reads, with the same names and defaults
(/root/reference/custom_hyperparams.py:35-47, :52, :119-123).  Unlike the reference's module this one imports on
Python >= 3.11 (dataclass defaults via default_factory)."""
from dataclasses import dataclass, field
from typing import List


@dataclass
class FusionParams:
    CLIP_FUSION_SCORES: bool = True
    LOGITS_USE_PIECEWISE_FUNCTION: bool = True
    DISTANCE_USE_FROM_ZERO_TO_THR: bool = False
    DISTANCE_USE_IN_DISTRIBUTION_TO_DEFINE_LIMITS: bool = True


@dataclass
class ClustersParams:
    MIN_SAMPLES: int = 3
    RANGE_OF_CLUSTERS: List[int] = field(default_factory=lambda: list(range(2, 15)))
    VISUALIZE: bool = False
    REMOVE_ORPHANS: bool = False


@dataclass
class Hyperparams:
    IOU_THRESHOLD: float = 0.5
    GOOD_NUM_SAMPLES: int = 25
    MIN_NUMBER_OF_SAMPLES_FOR_THR: int = 5
    clusters: ClustersParams = field(default_factory=ClustersParams)
    fusion: FusionParams = field(default_factory=FusionParams)
    BENCHMARK_MODE: bool = False


CUSTOM_HYP = Hyperparams()

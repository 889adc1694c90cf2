"""The knobs of the reference's global `CUSTOM_HYP` that the hot path reads, with the same names and defaults
(/root/reference/custom_hyperparams.py:35-47, :52, :65-76, :112-113, :119-123).  Unlike the reference's module this one imports on
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
class RankParams:
    """custom_hyperparams.py:65-76: how the distances of an unknown proposal to every class are folded into one rank."""
    RANK_BOXES_OPERATION: str = "entropy"        # mean, max, min, sum, geometric_mean, entropy
    USE_OOD_THR_TO_REMOVE_PROPS: bool = False


@dataclass
class UnkParams:
    RANK_BOXES: bool = True
    rank: RankParams = field(default_factory=RankParams)


@dataclass
class Hyperparams:
    IOU_THRESHOLD: float = 0.5
    GOOD_NUM_SAMPLES: int = 25
    MIN_NUMBER_OF_SAMPLES_FOR_THR: int = 5
    clusters: ClustersParams = field(default_factory=ClustersParams)
    fusion: FusionParams = field(default_factory=FusionParams)
    unk: UnkParams = field(default_factory=UnkParams)
    BENCHMARK_MODE: bool = False


CUSTOM_HYP = Hyperparams()

"""tod_b200 — B200-native (sm_100a) implementation of TOD's detection hot path: ORB descriptor k-NN Hamming matching
against the trained object DB + geometric guess generation, behind the reference's DescriptorMatcher /
GuessGenerator cell surface.  All compute lives in libtod_b200.so (hand-written CUDA, C-ABI in include/tod_b200.h);
this package is the Python binding.  There is no CPU fallback."""
from . import capi, dbio  # noqa: F401
from .cells import (DescriptorMatcher, FeatureDescriptor, GuessGenerator, Trainer, comm_unique_id, depth_to_3d, detector_from_ork, fill_adjacency, ork_parameters,  # noqa: F401
                    score_hypotheses)

__all__ = ["capi", "dbio", "DescriptorMatcher", "FeatureDescriptor", "GuessGenerator", "Trainer", "comm_unique_id", "depth_to_3d", "detector_from_ork", "fill_adjacency", "ork_parameters",
           "score_hypotheses"]

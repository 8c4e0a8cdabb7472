"""quicksand_b200 -- B200-native many-chain MCMC engine behind quicksand's infer/MCMC/kernel API.

The compute lives in ``libqsb200.so`` (hand-written CUDA for sm_100a behind the plain-C ABI of
``include/qsb200.h``); this package is the host-side mirror of the reference's interface for the
hot path.  There is no CPU fallback: without the built library every compute call raises.
"""
from . import programs
from .api import (AnnealingKernel, TraceMHKernel, Autocorrelation, DriftKernel, Expectation, HARMKernel, HMCKernel, MAP, MCMC, MixtureKernel,
                  Sampler, Group, Comm, Samples, SampleVector, infer, measure_fp64_peak, source_hash)
from .programs import Program

__all__ = ["programs", "Program", "infer", "MCMC", "HMCKernel", "DriftKernel", "HARMKernel", "MixtureKernel",
           "AnnealingKernel", "TraceMHKernel", "Samples", "MAP", "Expectation", "Autocorrelation", "Sampler", "Group", "Comm", "SampleVector", "measure_fp64_peak", "source_hash"]

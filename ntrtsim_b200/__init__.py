"""ntrtsim_b200 -- B200-native batched tensegrity step behind the NTRTsim modelling API.

csrc/   CUDA kernels (sm_100a) + the C ABI of include/ntrtsim_b200.h  -> lib/libntrtsim_b200.so
host/   C++ mirror of the reference's tgStructure / tgBuildSpec / tgModel / tgSimulation API -> lib/libtgb_host.so
"""
from . import _capi as capi  # noqa: F401
from .world import BatchedWorld, FlatModel  # noqa: F401
from .examples import (build_example, escape_transform_actions, tree_counts, EXAMPLES, CpgController, tetraspine_cpg,  # noqa: F401
                       tetraspine_param_rows, flemons_cpg, flemons_param_rows, serialized_spine_params)

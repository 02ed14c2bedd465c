"""tapir_b200 -- B200-native (sm_100a) implementation of tapir's Townsend phylogenetic-
informativeness hot path, behind tapir's own Python API.

Layout: ``csrc/`` CUDA kernels + the C ABI (``include/tapir_cuda.h``); ``engine`` the batched
multi-locus driver; ``compute`` / ``base`` / ``db`` the host-side mirror of the reference's
``tapir.compute`` / ``tapir.base`` / ``tapir.db`` modules for this path.
"""
__version__ = "0.1.0"

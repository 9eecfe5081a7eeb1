"""B200-native backend for the search hot path of lab-1806-vec-db.

Host-side mirror of the reference's index_algorithm / distance interface (same names, argument
meaning and error behaviour) over the C ABI in include/vdb_b200.h. All compute runs in
hand-written sm_100a CUDA kernels (csrc/); there is no CPU fallback.
"""
from ._lib import VdbError, lib  # noqa: F401
from .index import (CandidatePair, DeviceVecSet, FlatIndex, HNSWConfig, HNSWIndex, IVFConfig, IVFIndex, KMeans, KMeansConfig,  # noqa: F401
                    PQConfig, PQTable, calc_dist, calc_dist_batch, dist_cache, gather_dist, hnsw_rand_levels, init_devices,
                    k_means_init,
                    pq_groups)

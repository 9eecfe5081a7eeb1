// index.cuh — device mirrors of PQTable / IVFIndex and the k-means launchers.
#pragma once
#include "dataset.cuh"

namespace vdb {}  // filled in by kmeans.cu / pq.cu / ivf.cu

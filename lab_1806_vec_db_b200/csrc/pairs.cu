// pairs.cu — distances of explicit (a, b) vector pairs: K2b exact rerank, K3 row cache, K10 HNSW
// candidate evaluation, and the calc_dist primitive.
//
// One warp per pair, coalesced loads along the dimension, FP32 tree reduction. Replaces
//   DistanceAdapter<[T],[T]>::distance            reference src/distance/mod.rs:106-113
//   DistanceAdapter<(&[T],f32),(&[T],f32)>        src/distance/mod.rs:120-129 (cached forms :54-57, :67-69)
//   DistanceAlgorithm::dist_cache                 src/distance/mod.rs:31-36
#include <type_traits>

#include "dataset.cuh"
#include "scanmath.cuh"

namespace vdb {

enum PairMode { PM_L2 = 0, PM_COSINE = 1, PM_DOT = 2, PM_L2_CACHED = 3, PM_COSINE_CACHED = 4, PM_SQNORM = 5, PM_NORM = 6, PM_L2_SCANORDER = 7, PM_COS_SCANORDER = 8 };

struct PairParams {
    const void* A;           // rows of A
    uint64_t strideA;        // elements between rows
    const uint32_t* idxA;    // optional gather index (row of A for pair j); nullptr -> j (or j / divA)
    const void* B;
    uint64_t strideB;
    const uint32_t* idxB;
    const float* cacheA;     // per A-row cache (cached modes), indexed like A rows
    const float* cacheB;
    const uint8_t* valid;    // optional per-pair mask (invalid pairs are skipped, out = 0)
    bool vec4;               // f32 rows, dim % 4 == 0, 16-byte aligned bases and strides: use 128-bit loads
    uint32_t chunk;          // scan-order modes: elements per lane and step (4 for f32 rows, 16 for u8 rows)
    uint32_t dim;
    uint64_t npairs;
    const uint64_t* npairs_dev;  // optional: the pair count lives on the device (npairs is then the grid bound)
    float* out;
};

template <typename TA, typename TB, int MODE>
__global__ void __launch_bounds__(256) pair_dist_kernel(const PairParams p) {
    const int lane = threadIdx.x & 31;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const uint64_t npairs = p.npairs_dev ? *p.npairs_dev : p.npairs;
    for (uint64_t j = warp; j < npairs; j += nwarps) {
        if (p.valid && !p.valid[j]) {
            if (lane == 0) p.out[j] = 0.f;
            continue;
        }
        const uint64_t ia = p.idxA ? p.idxA[j] : j;
        const uint64_t ib = p.idxB ? p.idxB[j] : j;
        const TA* a = (const TA*)p.A + ia * p.strideA;
        const TB* b = (const TB*)p.B + ib * p.strideB;
        float s0 = 0.f, s1 = 0.f, s2 = 0.f;
        // u8 rows in the scan-order modes: exact integer sums like the streaming scan (scanmath.cuh), reduced in
        // integers and converted once - the same bits whatever the order
        constexpr bool INT_SUMS = (MODE == PM_L2_SCANORDER || MODE == PM_COS_SCANORDER) && std::is_same<TB, uint8_t>::value;
        if constexpr (INT_SUMS) {
            uint32_t si = 0, xi = 0;
            for (uint32_t e = lane; e < p.dim; e += 32) {
                const uint32_t xv = b[e], qv = (uint32_t)a[e];
                if (MODE == PM_L2_SCANORDER) {
                    const uint32_t d = xv > qv ? xv - qv : qv - xv;
                    si += d * d;
                } else {
                    si += xv * qv;
                    xi += xv * xv;
                }
            }
            s0 = (float)__reduce_add_sync(0xffffffffu, si);
            if (MODE == PM_COS_SCANORDER) s1 = (float)__reduce_add_sync(0xffffffffu, xi);
        } else
        if (MODE == PM_L2_SCANORDER || MODE == PM_COS_SCANORDER) {
            // same per-lane chains (float4 chunk c = it*32 + lane, even / odd elements, see scanmath.cuh) and the same
            // xor-butterfly as the streaming scan kernel (flat_scan.cu), so both Flat paths return bit-identical distances
            if (p.vec4) {
                // 128-bit loads, 8 row chunks in flight per lane before the first use
                const uint32_t nvec = p.dim >> 2;
                const float4* a4 = reinterpret_cast<const float4*>(a);
                const float4* b4 = reinterpret_cast<const float4*>(b);
                f32x2 t0 = 0ull, t1 = 0ull;
                const uint64_t pol = l2_evict_first_policy();
                for (uint32_t c0 = 0; c0 < nvec; c0 += 256) {
                    float4 xb[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const uint32_t c = c0 + u * 32 + lane;
                        xb[u] = c < nvec ? ldg_stream_f4_ef(b4 + c, pol) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const uint32_t c = c0 + u * 32 + lane;
                        if (c < nvec) {
                            const float4 q = __ldg(a4 + c);
                            const f32x2 x01 = pk2f(xb[u].x, xb[u].y), x23 = pk2f(xb[u].z, xb[u].w);
                            const f32x2 q01 = pk2f(q.x, q.y), q23 = pk2f(q.z, q.w);
                            t0 = chunk_acc<MODE == PM_L2_SCANORDER>(t0, x01, x23, q01, q23);
                            if (MODE == PM_COS_SCANORDER) t1 = chunk_acc<false>(t1, x01, x23, x01, x23);
                        }
                    }
                }
                s0 = sum2(t0);
                s1 = sum2(t1);
            } else {
                ScalarChains c0, c1;
                for (uint32_t c = lane; c * p.chunk < p.dim; c += 32) {
                    for (uint32_t i = 0; i < p.chunk; ++i) {
                        const uint32_t e = c * p.chunk + i;
                        if (e < p.dim) {
                            const float xv = (float)b[e], qv = (float)a[e];
                            if (MODE == PM_L2_SCANORDER) {
                                c0.l2(e, xv, qv);
                            } else {
                                c0.dot(e, xv, qv);
                                c1.dot(e, xv, xv);
                            }
                        }
                    }
                }
                s0 = c0.total();
                s1 = c1.total();
            }
        } else
        for (uint32_t e = lane; e < p.dim; e += 32) {
            const float x = (float)a[e];
            const float y = (MODE == PM_SQNORM || MODE == PM_NORM) ? x : (float)b[e];
            if (MODE == PM_L2) {
                const float d = x - y;
                s0 = fmaf(d, d, s0);
            } else {
                s0 = fmaf(x, y, s0);
                if (MODE == PM_COSINE) {
                    s1 = fmaf(x, x, s1);
                    s2 = fmaf(y, y, s2);
                }
            }
        }
        if constexpr (!INT_SUMS) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                s0 += __shfl_xor_sync(0xffffffffu, s0, o);
                if (MODE == PM_COSINE) {
                    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
                    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
                }
                if (MODE == PM_COS_SCANORDER) s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            }
        }
        if (lane == 0) {
            float r = s0;
            if (MODE == PM_COSINE) r = 1.0f - s0 / fmaxf(sqrtf(s1) * sqrtf(s2), 1e-10f);
            if (MODE == PM_L2_CACHED) r = (p.cacheA[ia] + p.cacheB[ib]) - 2.0f * s0;
            if (MODE == PM_COSINE_CACHED) r = 1.0f - s0 / fmaxf(p.cacheA[ia] * p.cacheB[ib], 1e-10f);
            if (MODE == PM_NORM) r = sqrtf(s0);
            // the streaming scan's cosine: 1 - dot / max(sqrt(x.x) * ||q||, 1e-10) with ||q|| from prepare_queries
            if (MODE == PM_COS_SCANORDER) r = 1.0f - s0 / fmaxf(sqrtf(s1) * p.cacheA[ia], 1e-10f);
            p.out[j] = r;
        }
    }
}

template <typename TA, typename TB>
static void launch_pairs_t(int mode, const PairParams& p, cudaStream_t st) {
    if (p.npairs == 0) return;
    const uint32_t grid = (uint32_t)std::min<uint64_t>(ceil_div<uint64_t>(p.npairs, 8), (uint64_t)sm_count() * 16);
    switch (mode) {
        case PM_L2: pair_dist_kernel<TA, TB, PM_L2><<<grid, 256, 0, st>>>(p); break;
        case PM_COSINE: pair_dist_kernel<TA, TB, PM_COSINE><<<grid, 256, 0, st>>>(p); break;
        case PM_DOT: pair_dist_kernel<TA, TB, PM_DOT><<<grid, 256, 0, st>>>(p); break;
        case PM_L2_CACHED: pair_dist_kernel<TA, TB, PM_L2_CACHED><<<grid, 256, 0, st>>>(p); break;
        case PM_COSINE_CACHED: pair_dist_kernel<TA, TB, PM_COSINE_CACHED><<<grid, 256, 0, st>>>(p); break;
        case PM_SQNORM: pair_dist_kernel<TA, TB, PM_SQNORM><<<grid, 256, 0, st>>>(p); break;
        case PM_L2_SCANORDER: pair_dist_kernel<TA, TB, PM_L2_SCANORDER><<<grid, 256, 0, st>>>(p); break;
        case PM_COS_SCANORDER: pair_dist_kernel<TA, TB, PM_COS_SCANORDER><<<grid, 256, 0, st>>>(p); break;
        default: pair_dist_kernel<TA, TB, PM_NORM><<<grid, 256, 0, st>>>(p); break;
    }
    VDB_LAUNCHED();
}

// a_f32: A rows are f32 (prepared queries) even when the dataset is u8
void launch_pairs(int mode, bool a_is_f32, int b_dtype, const PairParams& p, cudaStream_t st) {
    ProfScope prof("rerank", st);
    if (a_is_f32) {
        if (b_dtype == VDB_F32) launch_pairs_t<float, float>(mode, p, st);
        else launch_pairs_t<float, uint8_t>(mode, p, st);
    } else {
        if (b_dtype == VDB_F32) launch_pairs_t<float, float>(mode, p, st);
        else launch_pairs_t<uint8_t, uint8_t>(mode, p, st);
    }
}

// K3: per-row dist_cache (L2Sqr -> ||v||^2, Cosine -> ||v||)
void row_cache(const vdb_dataset* ds, float* d_out, cudaStream_t st) {
    PairParams p{};
    p.A = p.B = ds->d_rows;
    p.strideA = p.strideB = ds->pitch;
    p.dim = ds->dim;
    p.npairs = ds->n;
    p.out = d_out;
    const int mode = ds->metric == VDB_L2SQR ? PM_SQNORM : PM_NORM;
    if (ds->dtype == VDB_F32) launch_pairs_t<float, float>(mode, p, st);
    else launch_pairs_t<uint8_t, uint8_t>(mode, p, st);
}

// exact (difference form / 3-dot cosine) distances between raw queries [nq][dim] of the dataset dtype and
// listed local rows: pair j = (query qidx[j], row rid[j])
void exact_pair_distances(const vdb_dataset* ds, const void* d_queries, const uint32_t* d_qidx,
                          const uint32_t* d_rid, uint64_t npairs, float* d_out, cudaStream_t st) {
    PairParams p{};
    p.A = d_queries;
    p.strideA = ds->dim;
    p.idxA = d_qidx;
    p.B = ds->d_rows;
    p.strideB = ds->pitch;
    p.idxB = d_rid;
    p.dim = ds->dim;
    p.npairs = npairs;
    p.out = d_out;
    launch_pairs(ds->metric == VDB_L2SQR ? PM_L2 : PM_COSINE, false, ds->dtype, p, st);
}

// same with an explicit query pitch and a validity mask (K2 rerank)
void exact_pair_distances_masked(const vdb_dataset* ds, const void* d_queries, uint32_t qpitch, const uint32_t* d_qidx,
                                 const uint32_t* d_rid, const uint8_t* d_valid, uint64_t npairs, float* d_out,
                                 cudaStream_t st, const uint64_t* d_npairs, const float* d_qnorm) {
    PairParams p{};
    p.A = d_queries;
    p.strideA = qpitch;
    p.idxA = d_qidx;
    p.B = ds->d_rows;
    p.strideB = ds->pitch;
    p.idxB = d_rid;
    p.valid = d_valid;
    p.dim = ds->dim;
    p.npairs = npairs;
    p.npairs_dev = d_npairs;
    p.out = d_out;
    // the queries passed here are always the f32 copy; u8 rows are walked in the scan kernel's 16-element steps
    const bool scan_order = ds->metric == VDB_L2SQR || d_qnorm != nullptr;
    p.cacheA = d_qnorm;
    p.chunk = ds->dtype == VDB_F32 ? 4 : 16;
    p.vec4 = scan_order && ds->dtype == VDB_F32 && ds->dim % 4 == 0 && qpitch % 4 == 0 && ds->pitch % 4 == 0 &&
             ((uintptr_t)d_queries & 15) == 0 && ((uintptr_t)ds->d_rows & 15) == 0;
    const int so_mode = ds->metric == VDB_L2SQR ? PM_L2_SCANORDER : PM_COS_SCANORDER;
    launch_pairs(scan_order ? so_mode : (ds->metric == VDB_L2SQR ? PM_L2 : PM_COSINE), true, ds->dtype, p, st);
}

void cached_pair_distances(const vdb_dataset* ds, const void* d_queries, const float* d_qcache,
                           const float* d_rowcache, const uint32_t* d_qidx, const uint32_t* d_rid,
                           uint64_t npairs, float* d_out, cudaStream_t st) {
    PairParams p{};
    p.A = d_queries;
    p.strideA = ds->dim;
    p.idxA = d_qidx;
    p.B = ds->d_rows;
    p.strideB = ds->pitch;
    p.idxB = d_rid;
    p.cacheA = d_qcache;
    p.cacheB = d_rowcache;
    p.dim = ds->dim;
    p.npairs = npairs;
    p.out = d_out;
    launch_pairs(ds->metric == VDB_L2SQR ? PM_L2_CACHED : PM_COSINE_CACHED, false, ds->dtype, p, st);
}

void raw_pair_distances(const void* d_a, const void* d_b, uint64_t count, uint32_t dim, int dtype, int metric,
                        float* d_out, cudaStream_t st) {
    PairParams p{};
    p.A = d_a;
    p.B = d_b;
    p.strideA = p.strideB = dim;
    p.dim = dim;
    p.npairs = count;
    p.out = d_out;
    const int mode = metric == VDB_L2SQR ? PM_L2 : (metric == VDB_COSINE ? PM_COSINE : PM_DOT);
    launch_pairs(mode, false, dtype, p, st);
}

// expands CSR offsets into a per-pair query index
__global__ void expand_offsets_kernel(const uint64_t* __restrict__ off, uint32_t nq, uint32_t* __restrict__ qidx) {
    const uint32_t q = blockIdx.x;
    if (q >= nq) return;
    for (uint64_t j = off[q] + threadIdx.x; j < off[q + 1]; j += blockDim.x) qidx[j] = q;
}
void expand_offsets(const uint64_t* d_off, uint32_t nq, uint32_t* d_qidx, cudaStream_t st) {
    if (nq == 0) return;
    expand_offsets_kernel<<<nq, 128, 0, st>>>(d_off, nq, d_qidx);
    VDB_LAUNCHED();
}

}  // namespace vdb

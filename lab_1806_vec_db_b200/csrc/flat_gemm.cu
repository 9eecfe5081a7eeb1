// flat_gemm.cu — K2: batched Flat L2 kNN as a tensor-core contraction + exact FP32 rerank.
//
// Replaces FlatIndex::knn (reference src/index_algorithm/flat_index.rs:48-57) for large query batches
// (the additive batch entry; the reference batches with rayon, examples/bench.rs:414-418).
//
// Pipeline (all on the caller's stream):
//   1. side arrays per dataset: ||x||^2 and ||x|| (fp32), built once and cached on the handle
//   2. SAMPLE pass  : S' over a strided row sample, per query the j-th smallest S' becomes tau_q
//   3. FILTER pass  : for every (query, row): S' = ||x||^2 - 2 q.x - c ||q|| ||x||  (a LOWER bound of
//                     d(q,x) - ||q||^2: c bounds the TF32 rounding of the contraction);
//                     rows with S' < tau_q are appended to the query's candidate list
//   4. RERANK       : exact difference-form distances of the candidates (pairs.cu), top-k by (d, id)
//   5. CHECK        : the candidate set is provably complete iff  d_k - ||q||^2 < tau_q  (every excluded
//                     row has d - ||q||^2 >= S' >= tau_q); queries that fail (or overflowed their list)
//                     are re-run through the exact streaming scan (K1) — still on the GPU.
// Results are therefore identical to the exact scan: the tensor cores only PRUNE.
//
// The contraction kernel is hand-written for sm_100a: TMA (cp.async.bulk.tensor, 128B swizzle) feeds a
// 4-stage shared-memory ring, one elected thread issues tcgen05.mma.kind::tf32 (M=128 queries x N=256 rows,
// K=8 per instruction) into double-buffered TMEM accumulators, four epilogue warps read them back with
// tcgen05.ld (thread = query, registers = rows) and apply the per-query threshold in registers.
#include <cuda.h>

#include <cmath>
#include <cstdlib>
#include <algorithm>
#include <mutex>
#include <vector>

#include "index.cuh"
#include "tc.cuh"
#include "topk.cuh"

namespace vdb {

// ---- tile configuration ---------------------------------------------------------------------------------
constexpr int GM = 128;              // queries per tile (TMEM lanes)
constexpr int GN = 256;              // database rows per tile (TMEM columns per accumulator)
constexpr int GK = 32;               // fp32 elements per k-block = one 128-byte swizzle row
constexpr int G_A_BYTES = GM * GK * 4;   // 16 KB: 128 queries x 128 B
constexpr int G_THREADS = 192;       // warps 0-3 epilogue, warp 4 TMA producer, warp 5 MMA issuer
constexpr int G_TMEM_COLS = 512;     // 2 accumulators x 256 columns
constexpr int G_MAX_STAGES = 6;
// CTAS = 1: one CTA computes a 128 x 256 tile. CTAS = 2: a CTA pair (cta_group::2) computes 256 x 256, each CTA
// loads its own 128 queries and HALF of the 256 database rows, so the shared-memory fill per MMA cycle drops 1.5x.
template <int CTAS> struct GemmCfg {
    static constexpr int B_ROWS = GN / CTAS;                    // database rows loaded per CTA and k-block
    static constexpr int B_BYTES = B_ROWS * GK * 4;
    static constexpr int STAGE_BYTES = G_A_BYTES + B_BYTES;     // 48 KB / 32 KB
    static constexpr int STAGES = CTAS == 1 ? 4 : 6;            // 192 KB either way
    static constexpr uint32_t SMEM = 1024 /*align*/ + STAGES * STAGE_BYTES + 2 * 2 * GN * 4 /*norm tiles*/ + 512 /*barriers*/;
    // instruction descriptor (cute::UMMA::InstrDescriptor): D=F32, A=B=TF32, both K-major, N>>3, M>>4
    static constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(GN >> 3) << 17) |
                                      ((uint32_t)((GM * CTAS) >> 4) << 24);
};
constexpr uint32_t G_SAMPLE = 32768;  // sampled rows for the thresholds
constexpr int G_TOPJ = 16;             // order statistics kept in registers by the sample pass (mode 2)
constexpr int G_ITEMQ = 4;              // depth of the dynamic-scheduler item queue
constexpr uint32_t G_NO_ITEM = 0xFFFFFFFFu;
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;  // clears the CTA-pair peer bit of a shared::cluster address (-> even CTA)

// ---- kernel ----------------------------------------------------------------------------------------------------
// one unit of work: query rows [q0, q_end) (at most GM * CTAS of them) against database rows [r0, r_end)
struct GemmItem {
    uint32_t q0, q_end;
    uint64_t r0, r_end;
};
struct ItemView {
    uint32_t q0, q_end, ntile, slab;
    uint64_t r0, r_end;
};

struct GemmParams {
    uint32_t nq;            // queries
    uint64_t nrows;         // rows addressed by the B tensor map (sample rows or all rows)
    uint32_t row_stride;    // database row of B row i is i * row_stride
    uint32_t kblocks;       // ceil(dim / 32)
    uint32_t ntiles;        // ceil(nrows / GN)
    uint32_t nqt;           // query-tile UNITS: ceil(nq / (GM * CTAS))
    uint32_t tiles_per_slab;
    uint32_t nslabs;        // slabs covered by this launch
    uint32_t slab0;         // first slab of this launch (the filter pass may be cut into row parts)
    const float* sqnorm;    // [n] ||x||^2
    const float* rnorm;     // [n] ||x||
    const float* qcm;       // [nq] L2Sqr: c * ||q|| (pruning-bound coefficient times the query norm); cosine: 1/||q||
    const float* qnorm;     // [nq] cosine: ||q||
    float kc;               // cosine: 1 - (bound on the cosine-distance error)
    // mode 0: store keys (S', sample index) to out_keys[nq][nrows]
    // mode 2: per (query, slab) the G_TOPJ smallest S' as keys to out_keys[nq][nslabs][G_TOPJ]
    uint64_t* out_keys;
    // mode 1: filter
    const float* tau;       // [nq]
    uint32_t* cand_cnt;     // [nq]
    uint64_t* cand;         // [nq][cap]  (S' bits << 32 | local row)
    uint32_t cap;
    uint32_t* work_counter; // dynamic item scheduler (zeroed before the launch)
    // table mode (IVF: one rectangular block per (list, query group)): explicit items instead of the slab x unit grid
    const GemmItem* items;  // nullptr -> the grid of plan_gemm
    uint32_t nitems;
    const uint32_t* qmap;   // table mode: row of the gathered query matrix -> query that owns the candidate list
};

template <int CTAS>
__device__ __forceinline__ ItemView decode_item(const GemmParams& p, uint32_t item) {
    ItemView v;
    if (p.items) {
        const GemmItem it = p.items[item];
        v.q0 = it.q0;
        v.q_end = it.q_end;
        v.r0 = it.r0;
        v.r_end = it.r_end;
        v.slab = 0;  // mode 2: every gathered query row belongs to exactly one item
    } else {
        const uint32_t sl = item / p.nqt;
        v.slab = p.slab0 + sl;
        v.q0 = (item - sl * p.nqt) * GM * CTAS;
        v.q_end = p.nq;
        v.r0 = (uint64_t)v.slab * p.tiles_per_slab * GN;
        v.r_end = min(p.nrows, v.r0 + (uint64_t)p.tiles_per_slab * GN);
    }
    v.ntile = (uint32_t)((v.r_end - v.r0 + GN - 1) / GN);
    return v;
}

template <int MODE, int CTAS, int METRIC>
__global__ void __launch_bounds__(G_THREADS, 1)
flat_gemm_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_x, const GemmParams p) {
    using Cfg = GemmCfg<CTAS>;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* stage_base = smem;
    float* norm_tiles = reinterpret_cast<float*>(smem + STAGES * Cfg::STAGE_BYTES);  // [2 acc][2 arrays][GN]
    uint64_t* bars = reinterpret_cast<uint64_t*>(norm_tiles + 2 * 2 * GN);
    uint64_t* full_bar = bars;                       // [STAGES]  (pair mode: the leader's copy is the live one)
    uint64_t* empty_bar = bars + G_MAX_STAGES;       // [STAGES]  per CTA
    uint64_t* tfull_bar = bars + 2 * G_MAX_STAGES;   // [2]       per CTA
    uint64_t* tempty_bar = tfull_bar + 2;            // [2]       (pair mode: the leader's copy is the live one)
    uint64_t* iq_full = tempty_bar + 2;              // [G_ITEMQ] per CTA: an item id has been published
    uint64_t* iq_empty = iq_full + G_ITEMQ;          // [G_ITEMQ] (leader's copy is live): every reader took it
    uint32_t* item_ring = reinterpret_cast<uint32_t*>(iq_empty + G_ITEMQ);  // [G_ITEMQ]
    uint32_t* tmem_slot = item_ring + G_ITEMQ;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t cta_rank = CTAS == 2 ? cluster_ctarank() : 0u;
    const bool leader = cta_rank == 0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tfull_bar[a], 1);
            mbar_init(&tempty_bar[a], 4 * CTAS);  // one arrive per epilogue warp of every CTA in the group
        }
        for (int i = 0; i < G_ITEMQ; ++i) {
            mbar_init(&iq_full[i], 1);
            // readers of a published item: MMA thread + 4 epilogue warps (+ the peer's producer and 4 epilogue warps)
            mbar_init(&iq_empty[i], CTAS == 1 ? 5 : 10);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 5) {
        if (CTAS == 2) tmem_alloc2(tmem_slot, G_TMEM_COLS);
        else tmem_alloc(tmem_slot, G_TMEM_COLS);
    }
    tc_fence_before();
    __syncthreads();
    if (CTAS == 2) cluster_sync_all();  // the peer's barriers are initialised before any remote arrive / TMA
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const uint32_t items = p.items ? p.nitems : p.nqt * p.nslabs;
    // ---- dynamic scheduler: the leader's producer draws item ids from a global counter and publishes them to
    // every role of the CTA (pair) through a small shared-memory queue. Items are taken in global order, so the
    // CTAs in flight always work on neighbouring slabs (shared in L2) no matter how their speeds drift apart.
    const uint32_t leader_empty_base = CTAS == 2 ? mapa_cluster(smem_u32(iq_empty), 0) : smem_u32(iq_empty);
    auto take_item = [&](uint32_t n) -> uint32_t {  // called by ONE thread of a reader role, n = 0, 1, 2, ...
        const uint32_t slot = n % G_ITEMQ, ph = (n / G_ITEMQ) & 1;
        if (CTAS == 2) mbar_wait_cluster(&iq_full[slot], ph);
        else mbar_wait(&iq_full[slot], ph);
        const uint32_t item = ((volatile uint32_t*)item_ring)[slot];
        if (CTAS == 2) mbar_arrive_cluster_release(leader_empty_base + slot * 8);
        else mbar_arrive(&iq_empty[slot]);
        return item;
    };

    if (warp == 4) {
        // ===== TMA producer (every CTA loads its own queries and its share of the database rows) =====
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (uint32_t n = 0;; ++n) {
                uint32_t item;
                if (leader) {
                    const uint32_t slot = n % G_ITEMQ, ph = (n / G_ITEMQ) & 1;
                    mbar_wait_cluster(&iq_empty[slot], ph ^ 1);
                    item = atomicAdd(p.work_counter, 1u);
                    if (item >= items) item = G_NO_ITEM;
                    item_ring[slot] = item;
                    mbar_arrive(&iq_full[slot]);
                    if (CTAS == 2) {
                        st_cluster_u32(mapa_cluster(smem_u32(&item_ring[slot]), 1), item);
                        mbar_arrive_cluster_release(mapa_cluster(smem_u32(&iq_full[slot]), 1));
                    }
                } else {
                    item = take_item(n);
                }
                if (item == G_NO_ITEM) break;
                const ItemView iv = decode_item<CTAS>(p, item);
                const int qrow = (int)(iv.q0 + cta_rank * GM);
                for (uint32_t t = 0; t < iv.ntile; ++t) {
                    const int brow = (int)(iv.r0 + (uint64_t)t * GN + cta_rank * Cfg::B_ROWS);
                    for (uint32_t kb = 0; kb < p.kblocks; ++kb) {
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        const uint32_t sa = smem_u32(stage_base + stage * Cfg::STAGE_BYTES);
                        if (CTAS == 1) {
                            mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
                            tma_load_2d(sa, &map_q, (int)(kb * GK), qrow, &full_bar[stage]);
                            tma_load_2d(sa + G_A_BYTES, &map_x, (int)(kb * GK), brow, &full_bar[stage]);
                        } else {
                            // both CTAs' bytes are accounted on the LEADER's barrier
                            if (leader) mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES * CTAS);
                            const uint32_t lbar = smem_u32(&full_bar[stage]) & PEER_MASK;
                            tma_load_2d_2sm(sa, &map_q, (int)(kb * GK), qrow, lbar);
                            tma_load_2d_2sm(sa + G_A_BYTES, &map_x, (int)(kb * GK), brow, lbar);
                        }
                        if (++stage == STAGES) stage = 0, phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 5) {
        // ===== MMA issuer (one elected thread of the leader CTA) =====
        if (lane == 0 && leader) {
            uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
            for (uint32_t n = 0;; ++n) {
                const uint32_t item = take_item(n);
                if (item == G_NO_ITEM) break;
                const ItemView iv = decode_item<CTAS>(p, item);
                for (uint32_t t = 0; t < iv.ntile; ++t) {
                    mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + acc * GN;
                    for (uint32_t kb = 0; kb < p.kblocks; ++kb) {
                        mbar_wait(&full_bar[stage], phase);
                        tc_fence_after();
                        const uint32_t sa = smem_u32(stage_base + stage * Cfg::STAGE_BYTES);
                        const uint64_t da = umma_desc(sa), db = umma_desc(sa + G_A_BYTES);
#pragma unroll
                        for (int k = 0; k < GK / 8; ++k) {  // advance 32 bytes (2 x 16B units) per K=8 step
                            if (CTAS == 1) umma_tf32(d_tmem, da + 2 * k, db + 2 * k, Cfg::IDESC, (kb | k) != 0);
                            else umma_tf32_2sm(d_tmem, da + 2 * k, db + 2 * k, Cfg::IDESC, (kb | k) != 0);
                        }
                        // frees the smem stage (in both CTAs) when these MMAs retire
                        if (CTAS == 1) umma_commit(&empty_bar[stage]);
                        else umma_commit_2sm(&empty_bar[stage]);
                        if (++stage == STAGES) stage = 0, phase ^= 1;
                    }
                    // accumulator ready for the epilogue warps (of both CTAs)
                    if (CTAS == 1) umma_commit(&tfull_bar[acc]);
                    else umma_commit_2sm(&tfull_bar[acc]);
                    if (++acc == 2) acc = 0, acc_phase ^= 1;
                }
            }
        }
    } else {
        // ===== epilogue: warps 0-3, thread = query (TMEM lane), registers = database rows =====
        uint32_t acc = 0, acc_phase = 0;
        const uint32_t lane_base = (uint32_t)warp * 32;
        for (uint32_t n = 0;; ++n) {
            uint32_t item = 0;
            if (lane == 0) item = take_item(n);
            item = __shfl_sync(0xffffffffu, item, 0);
            if (item == G_NO_ITEM) break;
            const ItemView iv = decode_item<CTAS>(p, item);
            const uint32_t slab = iv.slab;
            const uint32_t q = iv.q0 + cta_rank * GM + threadIdx.x;   // row of the (gathered) query matrix
            const bool qok = q < iv.q_end;
            const uint32_t oq = (qok && p.qmap) ? p.qmap[q] : q;        // query that owns the candidate list
            const float cq = qok ? p.qcm[q] : 0.f;
            const float qn = (METRIC == VDB_COSINE && qok) ? p.qnorm[q] : 0.f;
            float tau = 0.f;
            if (MODE == 1) tau = qok ? p.tau[q] : __uint_as_float(0xff800000u);  // -inf: nothing passes
            float best[MODE == 2 ? G_TOPJ : 1];  // mode 2: this query's smallest scores of the slab, ascending
            if (MODE == 2) {
#pragma unroll
                for (int i = 0; i < (MODE == 2 ? G_TOPJ : 1); ++i) best[i] = __uint_as_float(0x7f800000u);
            }
            for (uint32_t t = 0; t < iv.ntile; ++t) {
                // stage the row-norm tiles of this N-tile (2 x GN floats) while the MMAs run
                float* sq = norm_tiles + acc * 2 * GN;
                float* rn = sq + GN;
                const uint64_t tile_row0 = iv.r0 + (uint64_t)t * GN;
                for (uint32_t c = threadIdx.x; c < GN; c += 128) {
                    const uint64_t brow = tile_row0 + c;
                    const bool ok = brow < iv.r_end;
                    const uint64_t row = brow * p.row_stride;
                    sq[c] = ok ? p.sqnorm[row] : __uint_as_float(0x7f800000u);  // +inf: never a candidate
                    rn[c] = ok ? p.rnorm[row] : 0.f;
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
                mbar_wait(&tfull_bar[acc], acc_phase);
                tc_fence_after();
                const uint32_t taddr = tmem_base + (lane_base << 16) + acc * GN;
#pragma unroll 1
                for (int c0 = 0; c0 < GN; c0 += 32) {
                    uint32_t v[32];
                    tmem_ld32(taddr + c0, v);
                    if (qok) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const float dot = __uint_as_float(v[j]);
                            float s;
                            if (METRIC == VDB_L2SQR) {
                                // S' = ||x||^2 - c||q|| ||x|| - 2 q.x   (lower bound of d - ||q||^2)
                                s = fmaf(-2.0f, dot, fmaf(-cq, rn[c0 + j], sq[c0 + j]));
                            } else {
                                // S' = (1 - bound) - q.x / (||q|| ||x||)  (lower bound of the cosine distance); rows whose
                                // norm product falls under the reference's 1e-10 clamp are always kept (sq = 1/||x||)
                                s = fmaf(-(cq * sq[c0 + j]), dot, p.kc);
                                if (qn * rn[c0 + j] < 2e-10f) s = __uint_as_float(0xff800000u);
                                if (tile_row0 + c0 + j >= iv.r_end) s = __uint_as_float(0x7f800000u);  // padding row
                            }
                            if (MODE == 0) {
                                const uint64_t brow = tile_row0 + c0 + j;
                                if (brow < iv.r_end) p.out_keys[(uint64_t)q * p.nrows + brow] = make_key(s, (uint32_t)brow);
                            } else if (MODE == 2) {
                                if (s < best[G_TOPJ - 1]) {  // rare after the first few hundred rows: bubble s into place
                                    float v = s;
#pragma unroll
                                    for (int i = 0; i < G_TOPJ; ++i) {
                                        const float lo = fminf(best[i], v);
                                        v = fmaxf(best[i], v);
                                        best[i] = lo;
                                    }
                                }
                            } else if (s < tau) {
                                const uint32_t pos = atomicAdd(&p.cand_cnt[oq], 1u);
                                if (pos < p.cap)
                                    p.cand[(uint64_t)oq * p.cap + pos] =
                                        ((uint64_t)__float_as_uint(s) << 32) | (uint32_t)(tile_row0 + c0 + j);
                            }
                        }
                    }
                }
                // hand the accumulator back to the MMA issuer: one arrive per warp, on the leader's barrier
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (CTAS == 1) mbar_arrive(&tempty_bar[acc]);
                    else mbar_arrive_cluster(smem_u32(&tempty_bar[acc]) & PEER_MASK);
                }
                if (++acc == 2) acc = 0, acc_phase ^= 1;
            }
            if (MODE == 2 && qok) {
#pragma unroll
                for (int i = 0; i < (MODE == 2 ? G_TOPJ : 1); ++i)
                    p.out_keys[((uint64_t)q * p.nslabs + slab) * G_TOPJ + i] = make_key(best[i], slab * G_TOPJ + i);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (CTAS == 2) cluster_sync_all();  // neither CTA may exit (or free TMEM) while its peer still uses it
    if (warp == 5) {
        if (CTAS == 2) tmem_dealloc2(tmem_base, G_TMEM_COLS);
        else tmem_dealloc(tmem_base, G_TMEM_COLS);
    }
}

// ---- host side ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) mean_kernel(const float* __restrict__ v, uint64_t n, uint64_t stride, float* out) {
    __shared__ float red[32];
    float s = 0.f;
    uint32_t cnt = 0;
    for (uint64_t i = (uint64_t)threadIdx.x * stride; i < n; i += (uint64_t)blockDim.x * stride) s += v[i], ++cnt;
    __shared__ uint32_t total;
    if (threadIdx.x == 0) total = 0;
    __syncthreads();
    atomicAdd(&total, cnt);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < 32; ++w) t += red[w];
        *out = t / (float)max(total, 1u);
    }
}

// round-to-nearest TF32 copy: the tensor core then consumes exactly representable operands, so the only operand
// error is this rounding (2^-11 relative) instead of the hardware's truncation (2^-10)
__global__ void round_tf32_kernel(const float* __restrict__ src, float* __restrict__ dst, uint64_t count) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t r;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(src[i]));
        dst[i] = __uint_as_float(r);
    }
}
__global__ void u8_to_f32_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst, uint64_t count) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (uint64_t)gridDim.x * blockDim.x)
        dst[i] = (float)src[i];
}
static void u8_to_f32(const uint8_t* src, float* dst, uint64_t count, cudaStream_t st) {
    if (count == 0) return;
    u8_to_f32_kernel<<<(uint32_t)std::min<uint64_t>(ceil_div<uint64_t>(count, 256), (uint64_t)sm_count() * 32), 256, 0, st>>>(src, dst, count);
    VDB_LAUNCHED();
}
// [nq][dim] u8 -> [nq][qpitch] f32 (zero padded)
__global__ void u8_rows_to_f32_kernel(const uint8_t* __restrict__ src, uint32_t dim, uint32_t qpitch, float* __restrict__ dst) {
    const uint32_t q = blockIdx.x;
    for (uint32_t e = threadIdx.x; e < qpitch; e += blockDim.x) dst[(size_t)q * qpitch + e] = e < dim ? (float)src[(size_t)q * dim + e] : 0.f;
}
static void round_tf32(const float* src, float* dst, uint64_t count, cudaStream_t st) {
    if (count == 0) return;
    round_tf32_kernel<<<(uint32_t)std::min<uint64_t>(ceil_div<uint64_t>(count, 256), (uint64_t)sm_count() * 32), 256, 0, st>>>(src, dst, count);
    VDB_LAUNCHED();
}

// stratified random sample: one row per bucket of n/ns consecutive rows, at a hashed offset inside the bucket
// (a plain stride would alias with periodic data)
__global__ void gather_sample_kernel(const float* __restrict__ rows_tf32, const float* __restrict__ sqnorm,
                                     const float* __restrict__ rnorm, uint64_t n, uint32_t pitch, uint32_t ns,
                                     float* __restrict__ out, float* __restrict__ out_sq, float* __restrict__ out_rn) {
    const uint32_t i = blockIdx.x;
    if (i >= ns) return;
    const uint64_t lo = (uint64_t)i * n / ns, hi = (uint64_t)(i + 1) * n / ns;
    uint64_t h = (i + 0x9E3779B97F4A7C15ull) * 0xBF58476D1CE4E5B9ull;
    h ^= h >> 31;
    h *= 0x94D049BB133111EBull;
    h ^= h >> 29;
    const uint64_t row = lo + h % (hi > lo ? hi - lo : 1);
    for (uint32_t e = threadIdx.x; e < pitch; e += blockDim.x) out[(size_t)i * pitch + e] = rows_tf32[row * pitch + e];
    if (threadIdx.x == 0) {
        out_sq[i] = sqnorm[row];
        out_rn[i] = rnorm[row];
    }
}

__global__ void rinv_kernel(const float* __restrict__ rn, uint64_t n, float* __restrict__ out) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        out[i] = rn[i] > 0.f ? 1.0f / rn[i] : 0.f;
}

static std::mutex g_side_mu;
// ||x||^2 and ||x|| per row, cached on the (logically const) dataset handle
static void ensure_side_arrays(const vdb_dataset* cds, cudaStream_t st) {
    vdb_dataset* ds = const_cast<vdb_dataset*>(cds);
    std::lock_guard<std::mutex> lk(g_side_mu);
    if (ds->d_sqnorm && ds->side_n == ds->n) return;
    if (ds->d_sqnorm) cudaFree(ds->d_sqnorm);
    if (ds->d_lo) cudaFree(ds->d_lo);
    if (ds->d_tf32) cudaFree(ds->d_tf32);
    if (ds->d_sample) cudaFree(ds->d_sample);
    if (ds->d_sample_sq) cudaFree(ds->d_sample_sq);
    if (ds->d_sample_rn) cudaFree(ds->d_sample_rn);
    ds->d_sqnorm = ds->d_lo = ds->d_tf32 = ds->d_sample = ds->d_sample_sq = ds->d_sample_rn = nullptr;
    VDB_CUDA(cudaMalloc(&ds->d_sqnorm, ds->n * 4));
    VDB_CUDA(cudaMalloc(&ds->d_lo, ds->n * 4));
    VDB_CUDA(cudaMalloc(&ds->d_tf32, ds->n * (size_t)ds->pitch * 4));
    if (ds->dtype == VDB_F32) round_tf32((const float*)ds->d_rows, ds->d_tf32, ds->n * ds->pitch, st);
    else u8_to_f32((const uint8_t*)ds->d_rows, ds->d_tf32, ds->n * ds->pitch, st);  // 8-bit values are exact in TF32
    // ~3 % of the shard, so the sample pass stays a small fixed fraction of the filter pass on every shard size
    ds->sample_n = (uint32_t)std::min<uint64_t>(std::min<uint64_t>(G_SAMPLE, ds->n / 2), std::max<uint64_t>(2048, ds->n / 30));
    VDB_CUDA(cudaMalloc(&ds->d_sample, (size_t)ds->sample_n * ds->pitch * 4));
    VDB_CUDA(cudaMalloc(&ds->d_sample_sq, (size_t)ds->sample_n * 4));
    VDB_CUDA(cudaMalloc(&ds->d_sample_rn, (size_t)ds->sample_n * 4));
    vdb_dataset tmp = *ds;
    tmp.metric = VDB_COSINE;
    row_cache(&tmp, ds->d_lo, st);  // ||x||
    if (ds->metric == VDB_L2SQR) {
        tmp.metric = VDB_L2SQR;
        row_cache(&tmp, ds->d_sqnorm, st);  // ||x||^2
    } else {
        rinv_kernel<<<(uint32_t)std::min<uint64_t>(ceil_div<uint64_t>(ds->n, 256), 65535), 256, 0, st>>>(ds->d_lo, ds->n, ds->d_sqnorm);
        VDB_LAUNCHED();  // 1/||x|| (0 for zero rows)
    }
    gather_sample_kernel<<<ds->sample_n, 256, 0, st>>>(ds->d_tf32, ds->d_sqnorm, ds->d_lo, ds->n, ds->pitch, ds->sample_n,
                                                       ds->d_sample, ds->d_sample_sq, ds->d_sample_rn);
    VDB_LAUNCHED();
    {
        DevBuf m(4, st);
        mean_kernel<<<1, 1024, 0, st>>>(ds->d_lo, ds->n, std::max<uint64_t>(1, ds->n / 65536), m.as<float>());
        VDB_LAUNCHED();
        VDB_CUDA(cudaMemcpyAsync(&ds->mean_norm, m.p, 4, cudaMemcpyDeviceToHost, st));
        VDB_CUDA(cudaStreamSynchronize(st));
    }
    ds->side_n = ds->n;
}

static int gemm_ctas();
// tiling of one pass: query-tile units x row slabs (small slabs, query tile fastest: the CTAs in flight share a few
// row tiles and all query tiles in L2)
static void plan_gemm(GemmParams& p, int ctas_sel = 0) {
    const uint32_t ctas = (uint32_t)(ctas_sel ? ctas_sel : gemm_ctas());
    const uint32_t sms = (uint32_t)sm_count();
    p.ntiles = (uint32_t)ceil_div<uint64_t>(p.nrows, GN);
    p.nqt = ceil_div<uint32_t>(p.nq, GM * ctas);
    static const uint32_t tps_env = getenv("VDB_GEMM_TPS") ? (uint32_t)atoi(getenv("VDB_GEMM_TPS")) : 0;
    const uint32_t tps = tps_env ? tps_env : 4u;
    p.tiles_per_slab = std::max(1u, std::min(tps, ceil_div(p.ntiles * p.nqt, sms / ctas)));
    p.nslabs = ceil_div(p.ntiles, p.tiles_per_slab);
}

static int gemm_ctas() {
    static const int v = getenv("VDB_GEMM_CTAS") ? atoi(getenv("VDB_GEMM_CTAS")) : 2;
    return v == 1 ? 1 : 2;
}

template <int MODE, int CTAS, int METRIC>
static void launch_gemm_t(const CUtensorMap& mq, const CUtensorMap& mx, GemmParams p, cudaStream_t st) {
    using Cfg = GemmCfg<CTAS>;
    auto kern = flat_gemm_kernel<MODE, CTAS, METRIC>;
    static std::atomic<size_t> configured[VDB_MAX_DEVICES];
    ensure_dyn_smem(kern, Cfg::SMEM, configured);
    const uint32_t sms = (uint32_t)sm_count();
    const uint32_t units = std::min(sms / CTAS, p.items ? p.nitems : p.nqt * p.nslabs);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(units * CTAS);
    cfg.blockDim = dim3(G_THREADS);
    cfg.dynamicSmemBytes = Cfg::SMEM;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CTAS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    DevBuf counter(4, st);
    VDB_CUDA(cudaMemsetAsync(counter.p, 0, 4, st));
    p.work_counter = counter.as<uint32_t>();
    ProfScope prof("flat_gemm", st);
    VDB_CUDA(cudaLaunchKernelEx(&cfg, kern, mq, mx, p));
    VDB_LAUNCHED();
}

// mode 0 = store every score, 1 = filter, 2 = per-slab smallest scores; `p` must have been planned (plan_gemm)
static void launch_gemm(int mode, int metric, const CUtensorMap& mq, const CUtensorMap& mx, const GemmParams& p,
                        cudaStream_t st, int ctas = 0) {
    const int sel = ((ctas ? ctas : gemm_ctas()) == 2 ? 6 : 0) + mode * 2 + (metric == VDB_COSINE ? 1 : 0);
    switch (sel) {
        case 0: launch_gemm_t<0, 1, VDB_L2SQR>(mq, mx, p, st); break;
        case 1: launch_gemm_t<0, 1, VDB_COSINE>(mq, mx, p, st); break;
        case 2: launch_gemm_t<1, 1, VDB_L2SQR>(mq, mx, p, st); break;
        case 3: launch_gemm_t<1, 1, VDB_COSINE>(mq, mx, p, st); break;
        case 4: launch_gemm_t<2, 1, VDB_L2SQR>(mq, mx, p, st); break;
        case 5: launch_gemm_t<2, 1, VDB_COSINE>(mq, mx, p, st); break;
        case 6: launch_gemm_t<0, 2, VDB_L2SQR>(mq, mx, p, st); break;
        case 7: launch_gemm_t<0, 2, VDB_COSINE>(mq, mx, p, st); break;
        case 8: launch_gemm_t<1, 2, VDB_L2SQR>(mq, mx, p, st); break;
        case 9: launch_gemm_t<1, 2, VDB_COSINE>(mq, mx, p, st); break;
        case 10: launch_gemm_t<2, 2, VDB_L2SQR>(mq, mx, p, st); break;
        default: launch_gemm_t<2, 2, VDB_COSINE>(mq, mx, p, st); break;
    }
}

// c * ||q|| per query; c bounds |S'_tf32 - S'_exact| / (||q|| ||x||): two TF32 operand roundings (2^-10 each,
// truncation) + their product + K fp32 accumulation steps (K * 2^-23), times 2 for the -2 q.x term.
// Query side of the tensor path in one pass (one warp per query): padded fp32 copy (exact rerank), TF32-rounded copy
// (MMA operand) and, for L2Sqr, ||q||^2 and c * ||q||. The norm is summed exactly like row_cache's PM_SQNORM kernel
// (lane-strided fma chain + xor butterfly), so thresholds do not depend on which of the two produced it.
template <typename T>
__global__ void __launch_bounds__(256) tq_prepare_kernel(const T* __restrict__ src, uint32_t nq, uint32_t dim, uint32_t qpitch,
                                                         float c, float* __restrict__ qcopy, float* __restrict__ qround,
                                                         float* __restrict__ qsq, float* __restrict__ qcm) {
    const uint32_t q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (q >= nq) return;
    const T* row = src + (size_t)q * dim;
    float s = 0.f;
    for (uint32_t e = lane; e < qpitch; e += 32) {
        const float v = e < dim ? (float)row[e] : 0.f;
        uint32_t r;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
        qcopy[(size_t)q * qpitch + e] = v;
        qround[(size_t)q * qpitch + e] = __uint_as_float(r);
        if (e < dim) s = fmaf(v, v, s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0 && qsq) {
        qsq[q] = s;
        qcm[q] = c * sqrtf(s);
    }
}
__global__ void qcm_kernel(const float* __restrict__ qsq, uint32_t nq, float c, float* __restrict__ qcm) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < nq) qcm[q] = c * sqrtf(qsq[q]);
}
// cosine: qcm = 1/||q|| with ||q|| exactly as the streaming scan computes it (prepare_queries)
__global__ void qinv_kernel(const float* __restrict__ qn, uint32_t nq, float* __restrict__ qinv) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < nq) qinv[q] = qn[q] > 0.f ? 1.0f / qn[q] : 0.f;
}
// tau_q = (j0-th smallest sampled S') + margin. j0 is chosen so that the k-th best S' of the shard is <= S'_(j0)
// with high probability; because S <= S' + 2 c||q|| ||x||, the margin (2.5 x the pruning bound at the mean row
// norm) keeps the k-th EXACT distance inside the threshold. The check kernel verifies it per query afterwards.
__global__ void tau_from_keys_kernel(const uint64_t* __restrict__ keys, uint32_t nq, uint32_t j, uint32_t j0,
                                     const float* __restrict__ qcm, float mean_norm, float* __restrict__ tau) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < nq) tau[q] = key_dist(keys[(size_t)q * j + (j0 - 1)]) + 2.5f * (qcm ? qcm[q] * mean_norm : mean_norm);
}
// exclusive scan of min(cnt, cap) over the queries (one block; nq is at most a few 100k)
__global__ void __launch_bounds__(1024) cand_offsets_kernel(const uint32_t* __restrict__ cnt, uint32_t nq, uint32_t cap,
                                                            uint64_t* __restrict__ off) {
    __shared__ uint64_t warp_sums[32];
    __shared__ uint64_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t base = 0; base < nq; base += blockDim.x) {
        const uint32_t q = base + threadIdx.x;
        const uint64_t v = q < nq ? min(cnt[q], cap) : 0u;
        uint64_t x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint64_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) warp_sums[warp] = x;
        __syncthreads();
        if (warp == 0) {
            uint64_t w = warp_sums[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint64_t y = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += y;
            }
            warp_sums[lane] = w;
        }
        __syncthreads();
        const uint64_t before = carry + (warp ? warp_sums[warp - 1] : 0) + x - v;
        if (q < nq) off[q] = before;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry = before + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) off[nq] = carry;
}
// ---- row-part pipeline of the filter pass: candidates [prev[q], cur[q]) of every query (one row part) -------------
// exclusive scan of the part's candidate counts; prev == nullptr -> 0
__global__ void __launch_bounds__(1024) part_offsets_kernel(const uint32_t* __restrict__ prev, const uint32_t* __restrict__ cur,
                                                            uint32_t nq, uint32_t cap, uint64_t* __restrict__ off) {
    __shared__ uint64_t warp_sums[32];
    __shared__ uint64_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t base = 0; base < nq; base += blockDim.x) {
        const uint32_t q = base + threadIdx.x;
        const uint64_t v = q < nq ? min(cur[q], cap) - (prev ? min(prev[q], cap) : 0u) : 0u;
        uint64_t x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint64_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) warp_sums[warp] = x;
        __syncthreads();
        if (warp == 0) {
            uint64_t w = warp_sums[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint64_t y = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += y;
            }
            warp_sums[lane] = w;
        }
        __syncthreads();
        const uint64_t before = carry + (warp ? warp_sums[warp - 1] : 0) + x - v;
        if (q < nq) off[q] = before;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry = before + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) off[nq] = carry;
}
__global__ void part_to_pairs_kernel(const uint64_t* __restrict__ cand, const uint32_t* __restrict__ prev,
                                     const uint32_t* __restrict__ cur, uint32_t cap, const uint64_t* __restrict__ off,
                                     uint32_t* __restrict__ qidx, uint32_t* __restrict__ rid) {
    const uint32_t q = blockIdx.x;
    const uint32_t j0 = prev ? min(prev[q], cap) : 0u, j1 = min(cur[q], cap);
    const uint64_t o = off[q];
    for (uint32_t j = j0 + threadIdx.x; j < j1; j += blockDim.x) {
        qidx[o + (j - j0)] = q;
        rid[o + (j - j0)] = (uint32_t)cand[(uint64_t)q * cap + j];
    }
}
// exact distances of the part's pairs -> final keys, written over the candidate entries they came from
__global__ void part_rekey_kernel(const float* __restrict__ dist, const uint32_t* __restrict__ qidx,
                                  const uint32_t* __restrict__ rid, const uint64_t* __restrict__ off, uint32_t nq,
                                  const uint32_t* __restrict__ prev, uint32_t cap, uint32_t id_base,
                                  uint64_t* __restrict__ cand) {
    const uint64_t n = off[nq];
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t q = qidx[i];
        const uint32_t j = (prev ? min(prev[q], cap) : 0u) + (uint32_t)(i - off[q]);
        cand[(uint64_t)q * cap + j] = make_key(dist[i], id_base + rid[i]);
    }
}
// candidate (S', local row) lists -> dense rerank inputs at off[q]
__global__ void cand_to_pairs_kernel(const uint64_t* __restrict__ cand, const uint32_t* __restrict__ cnt, uint32_t cap,
                                     const uint64_t* __restrict__ off, const uint32_t* __restrict__ pos_to_row,
                                     uint32_t* __restrict__ qidx, uint32_t* __restrict__ rid) {
    const uint32_t q = blockIdx.x;
    const uint32_t c = min(cnt[q], cap);
    const uint64_t o = off[q];
    for (uint32_t j = threadIdx.x; j < c; j += blockDim.x) {
        qidx[o + j] = q;
        const uint32_t pos = (uint32_t)cand[(uint64_t)q * cap + j];
        rid[o + j] = pos_to_row ? pos_to_row[pos] : pos;  // IVF: list-order position -> row id
    }
}
__global__ void rekey_dev_count_kernel(const float* __restrict__ dist, const uint32_t* __restrict__ ids, uint32_t id_base,
                                       const uint64_t* __restrict__ count, uint64_t* __restrict__ keys) {
    const uint64_t n = *count;
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (uint64_t)gridDim.x * blockDim.x)
        keys[j] = make_key(dist[j], ids[j] + id_base);
}
__global__ void overflow_kernel(const uint32_t* __restrict__ cnt, uint32_t nq, uint32_t cap, uint32_t* __restrict__ flag) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < nq) flag[q] = cnt[q] > cap ? 1u : 0u;
}
// completeness check; failing queries are appended to redo[]
// keys / overflow are indexed by the LOCAL query i of the checked range [q0, q0 + cnt) (a row-sharded search checks
// only the slice of queries a GPU owns); tau / qsq by the batch index q0 + i, which is also what redo[] receives.
__global__ void check_kernel(const uint64_t* __restrict__ keys, uint32_t q0, uint32_t cnt, uint32_t k, uint64_t n,
                             const uint32_t* __restrict__ overflow, const float* __restrict__ tau,
                             const float* __restrict__ qsq, uint32_t force_mod, uint32_t* __restrict__ redo,
                             uint32_t* __restrict__ nredo) {
    // qsq == nullptr: cosine (scores bound the distance itself); else L2Sqr (scores bound d - ||q||^2)
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cnt) return;
    const uint32_t q = q0 + i;
    const uint32_t need = (uint32_t)((uint64_t)k < n ? (uint64_t)k : n);
    bool ok = !(overflow && overflow[i]);
    if (ok && need) {
        const uint64_t kk = keys[(size_t)i * k + (need - 1)];
        ok = kk != KEY_NONE;  // fewer than `need` candidates survived the filter
        if (ok) {
            const float dk = key_dist(kk);
            const float shift = qsq ? qsq[q] : 0.f;
            const float slack = 2e-5f * (fabsf(dk) + shift + fabsf(tau[q]));
            ok = (dk - shift) < tau[q] - slack;
        }
    }
    if (force_mod && q % force_mod == 0) ok = false;
    if (!ok) redo[atomicAdd(nredo, 1u)] = q;
}
__global__ void gather_rows_kernel(const uint8_t* __restrict__ src, uint32_t row_bytes, const uint32_t* __restrict__ idx,
                                   uint32_t cnt, uint8_t* __restrict__ dst) {
    const uint32_t i = blockIdx.x;
    if (i >= cnt) return;
    for (uint32_t e = threadIdx.x; e < row_bytes; e += blockDim.x)
        dst[(size_t)i * row_bytes + e] = src[(size_t)idx[i] * row_bytes + e];
}
__global__ void scatter_keys_kernel(const uint64_t* __restrict__ src, uint32_t k, const uint32_t* __restrict__ idx,
                                    uint32_t cnt, uint64_t* __restrict__ dst) {
    const uint32_t i = blockIdx.x;
    if (i >= cnt) return;
    for (uint32_t e = threadIdx.x; e < k; e += blockDim.x) dst[(size_t)idx[i] * k + e] = src[(size_t)i * k + e];
}

constexpr uint32_t G_MAX_K = 1024;

bool flat_gemm_supported(const vdb_dataset* ds, uint32_t nq, uint32_t k) {
    return (ds->dtype == VDB_F32 || ds->dtype == VDB_U8) && k >= 1 && k <= G_MAX_K && ds->n >= 65536 &&
           nq >= 1 && ((uintptr_t)ds->d_rows & 15) == 0;
}

std::atomic<uint64_t> g_gemm_redo{0};  // queries that needed the exact fallback (instrumentation)
std::atomic<uint64_t> g_gemm_cands{0}; // candidates reranked (instrumentation)
std::atomic<uint64_t> g_gemm_queries{0};
// test hook (vdb_debug_force_redo): every query whose index is a multiple of m fails the completeness check, so
// the exact re-scan of flagged queries (single-GPU and sharded) can be exercised on purpose; 0 = off
std::atomic<uint32_t> g_debug_force_redo{0};

// ---- phases (also exported one by one for the row-sharded search, sharded.py) -------------------------------
// j0: smallest order statistic of an `ns`-row uniform sample whose rank among the `n` rows is >= k with high
// probability: P(rank < k) = P(Poisson(k*ns/n) >= j0) < 2e-3
uint32_t tensor_j0(uint32_t k, uint64_t ns, uint64_t n, double eps) {
    uint32_t j0 = 1;
    const double x = (double)k * (double)ns / (double)n;
    double term = exp(-x), cdf = term;
    while (1.0 - cdf >= eps && j0 < 4096) {
        term *= x / j0;
        cdf += term;
        ++j0;
    }
    return (uint32_t)std::min<uint64_t>(j0, ns);
}

}  // namespace vdb

// per-call query context of the tensor path: padded fp32 copy (exact rerank), TF32-rounded copy (MMA operand),
// ||q||^2, c||q|| and the TMA descriptor of the rounded copy
struct vdb_tq {
    const vdb_dataset* ds = nullptr;
    const void* d_queries = nullptr;
    uint32_t nq = 0, qpitch = 0;
    cudaStream_t st = nullptr;
    vdb::DevBuf qcopy, qround, qsq, qcm, cnt;
    vdb::QueryTile qtile;   // cosine: ||q|| exactly as the streaming scan computes it
    float kc = 0.f, cbound = 0.f;
    CUtensorMap mq;
    uint32_t cap = 0;
    int ctas = 2;   // CTAs per work unit: pairs (M = 256 queries), single CTAs (M = 128) for batches of <= 128 queries
};

namespace vdb {

vdb_tq* tensor_begin(const vdb_dataset* ds, const void* d_queries, uint32_t nq, cudaStream_t st) {
    VDB_REQUIRE(flat_gemm_supported(ds, nq, 1), "tensor-core Flat path: unsupported dataset");
    ensure_side_arrays(ds, st);
    const uint32_t dim = ds->dim;
    // pruning-bound coefficient: both operands rounded to TF32 (2^-11 each, products then exact in fp32) plus
    // dim fp32 accumulation steps, times 2 for the -2 q.x term
    // (u8 rows and queries are exact in TF32: only the accumulation term remains)
    const float c = ds->dtype == VDB_F32 ? 2.0f * (ldexpf(1.0f, -10) * 1.001f + (float)dim * ldexpf(1.0f, -23))
                                         : 2.0f * (ldexpf(1.0f, -20) + (float)dim * ldexpf(1.0f, -23));
    auto tq = new vdb_tq();
    try {
        tq->ds = ds;
        tq->d_queries = d_queries;
        tq->nq = nq;
        tq->st = st;
        tq->qpitch = round_up(dim, 4u);
        const size_t qbytes = (size_t)nq * tq->qpitch * 4;
        tq->qcopy = DevBuf(qbytes, st);
        tq->qround = DevBuf(qbytes, st);
        tq->qsq = DevBuf((size_t)nq * 4, st);
        tq->qcm = DevBuf((size_t)nq * 4, st);
        tq->cnt = DevBuf((size_t)nq * 4, st);
        const bool l2 = ds->metric == VDB_L2SQR;
        float* qsq = l2 ? tq->qsq.as<float>() : nullptr;
        const uint32_t pgrid = ceil_div(nq, 8u);   // 8 warps per CTA, one query each
        if (ds->dtype == VDB_U8)
            tq_prepare_kernel<uint8_t><<<pgrid, 256, 0, st>>>((const uint8_t*)d_queries, nq, dim, tq->qpitch, c,
                                                              tq->qcopy.as<float>(), tq->qround.as<float>(), qsq, tq->qcm.as<float>());
        else
            tq_prepare_kernel<float><<<pgrid, 256, 0, st>>>((const float*)d_queries, nq, dim, tq->qpitch, c,
                                                            tq->qcopy.as<float>(), tq->qround.as<float>(), qsq, tq->qcm.as<float>());
        VDB_LAUNCHED();
        if (!l2) {
            // cosine distance error <= (c/2) from the contraction + the reciprocal products of the epilogue
            tq->cbound = 0.5f * c + 1e-6f;
            tq->kc = 1.0f - tq->cbound;
            tq->qtile = prepare_queries(ds, d_queries, nq, st);
            qinv_kernel<<<ceil_div(nq, 256u), 256, 0, st>>>(tq->qtile.qcache.as<float>(), nq, tq->qcm.as<float>());
            VDB_LAUNCHED();
        }
        tq->mq = make_map(tq->qround.as<float>(), dim, nq, (uint64_t)tq->qpitch * 4, GM);
        // a batch that fits one 128-query tile runs on single CTAs: a CTA pair would spend half of its MMAs on padding
        // (the pass is then HBM-bound on the TF32 copy of the rows instead of tensor-bound)
        static const bool small_single = !(getenv("VDB_GEMM_SMALL_PAIRS") && atoi(getenv("VDB_GEMM_SMALL_PAIRS")));
        tq->ctas = (small_single && nq <= (uint32_t)GM) ? 1 : gemm_ctas();
    } catch (...) {
        delete tq;
        throw;
    }
    return tq;
}
void tensor_end(vdb_tq* tq) { delete tq; }
void tensor_info(const vdb_dataset* ds, uint64_t* n, uint32_t* sample_n, float* mean_norm, cudaStream_t st) {
    VDB_REQUIRE(flat_gemm_supported(ds, 1, 1), "tensor-core Flat path: unsupported dataset");
    ensure_side_arrays(ds, st);
    if (n) *n = ds->n;
    if (sample_n) *sample_n = ds->sample_n;
    if (mean_norm) *mean_norm = ds->mean_norm;
}

static GemmParams base_params(const vdb_tq* tq) {
    GemmParams p{};
    p.nq = tq->nq;
    p.kblocks = ceil_div(tq->ds->dim, (uint32_t)GK);
    p.qcm = tq->qcm.as<float>();
    p.qnorm = tq->ds->metric == VDB_COSINE ? tq->qtile.qcache.as<float>() : nullptr;
    p.kc = tq->kc;
    return p;
}

// SAMPLE: this shard's j smallest sampled S' keys per query, [nq][j] ascending
void tensor_sample_keys(vdb_tq* tq, uint32_t j, uint64_t* d_jkeys) {
    const vdb_dataset* ds = tq->ds;
    cudaStream_t st = tq->st;
    const uint64_t ns = ds->sample_n;
    VDB_REQUIRE(j >= 1 && j <= ns, "sample order statistic %u out of range (sample %llu)", j, (unsigned long long)ns);
    const CUtensorMap ms = make_map(ds->d_sample, ds->dim, ns, (uint64_t)ds->pitch * 4, GN / tq->ctas);
    GemmParams ps = base_params(tq);
    ps.nrows = ns;
    ps.row_stride = 1;
    ps.sqnorm = ds->d_sample_sq;
    ps.rnorm = ds->d_sample_rn;
    plan_gemm(ps, tq->ctas);
    if (j <= (uint32_t)G_TOPJ) {
        // the epilogue keeps each query's G_TOPJ smallest scores per slab in registers: nothing but
        // nq * nslabs * G_TOPJ keys ever reaches HBM
        DevBuf part((size_t)tq->nq * ps.nslabs * G_TOPJ * 8, st);
        ps.out_keys = part.as<uint64_t>();
        launch_gemm(2, ds->metric, tq->mq, ms, ps, st, tq->ctas);
        launch_merge_keys(part.as<uint64_t>(), ps.nslabs, tq->nq, G_TOPJ, false, j, d_jkeys, nullptr, nullptr, nullptr, st);
    } else {
        DevBuf skeys((size_t)tq->nq * ns * 8, st);
        ps.out_keys = skeys.as<uint64_t>();
        launch_gemm(0, ds->metric, tq->mq, ms, ps, st, tq->ctas);
        launch_merge_keys(skeys.as<uint64_t>(), 1, tq->nq, (uint32_t)ns, false, j, d_jkeys, nullptr, nullptr, nullptr, st);
    }
}

// TAU: merge `nlists` shards' [nq][j] sample keys (list-major) and set tau_q = S'_(j0) + margin
void tensor_tau(vdb_tq* tq, const uint64_t* d_lists, uint32_t nlists, uint32_t j, uint32_t j0, float mean_norm,
                float* d_tau) {
    cudaStream_t st = tq->st;
    VDB_REQUIRE(j0 >= 1 && j0 <= (uint64_t)j * nlists, "j0 out of range");
    const uint32_t jj = std::min<uint64_t>(j0, (uint64_t)j * nlists);
    const bool cosine = tq->ds->metric == VDB_COSINE;
    if (nlists == 1) {   // a single ascending [nq][j] list: its jj-th entry is the order statistic, nothing to merge
        tau_from_keys_kernel<<<ceil_div(tq->nq, 256u), 256, 0, st>>>(d_lists, tq->nq, j, jj, cosine ? nullptr : tq->qcm.as<float>(),
                                                                      cosine ? tq->cbound : mean_norm, d_tau);
        VDB_LAUNCHED();
        return;
    }
    DevBuf merged((size_t)tq->nq * jj * 8, st);
    launch_merge_sorted(d_lists, nlists, tq->nq, j, jj, merged.as<uint64_t>(), nullptr, nullptr, nullptr, st);  // per-shard lists are ascending
    tau_from_keys_kernel<<<ceil_div(tq->nq, 256u), 256, 0, st>>>(merged.as<uint64_t>(), tq->nq, jj, jj,
                                                                  cosine ? nullptr : tq->qcm.as<float>(),
                                                                  cosine ? tq->cbound : mean_norm, d_tau);
    VDB_LAUNCHED();
}

// side stream of the filter pass (per host thread and device): reranks row part i under the contraction of part i + 1
struct SideStream {
    int device = -1;
    cudaStream_t s = nullptr;
    cudaEvent_t ev = nullptr, ev2 = nullptr;
    void ensure(int dev) {
        if (device == dev) return;
        release();
        VDB_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
        VDB_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        VDB_CUDA(cudaEventCreateWithFlags(&ev2, cudaEventDisableTiming));
        device = dev;
    }
    void release() {
        if (device < 0) return;
        cudaStreamDestroy(s);
        cudaEventDestroy(ev);
        cudaEventDestroy(ev2);
        device = -1;
    }
    ~SideStream() { release(); }
};
// drains the side stream if the filter pass unwinds early (its scratch is freed in stream order on the MAIN stream)
struct SideDrain {
    cudaStream_t s;
    bool armed;
    ~SideDrain() {
        if (armed) cudaStreamSynchronize(s);
    }
};
// sum of min(cnt, cap) over the queries (instrumentation: candidates reranked)
__global__ void __launch_bounds__(1024) cand_total_kernel(const uint32_t* __restrict__ cnt, uint32_t nq, uint32_t cap,
                                                          uint64_t* __restrict__ out) {
    __shared__ unsigned long long sum;
    if (threadIdx.x == 0) sum = 0;
    __syncthreads();
    unsigned long long mine = 0;
    for (uint32_t q = threadIdx.x; q < nq; q += blockDim.x) mine += min(cnt[q], cap);
    atomicAdd(&sum, mine);
    __syncthreads();
    if (threadIdx.x == 0) *out = sum;
}

// FILTER + RERANK: rows with S' < tau_q -> exact distances -> this shard's k best keys per query.
// d_overflow[q] = 1 when the candidate list of q overflowed (its result is then incomplete).
void tensor_filter_keys(vdb_tq* tq, uint32_t k, uint32_t j0_local_hint, const float* d_tau, uint64_t* d_keys,
                        uint32_t* d_overflow, uint64_t* d_cand_total) {
    const vdb_dataset* ds = tq->ds;
    cudaStream_t st = tq->st;
    const uint32_t nq = tq->nq, dim = ds->dim;
    const uint64_t ns = ds->sample_n;
    const uint32_t cap = (uint32_t)next_pow2((uint32_t)std::min<uint64_t>(
        ds->n, std::max<uint64_t>(8ull * std::max(j0_local_hint, 1u) * (ds->n / ns), 8192)));
    tq->cap = cap;
    DevBuf cand((size_t)nq * cap * 8, st);
    VDB_CUDA(cudaMemsetAsync(tq->cnt.p, 0, (size_t)nq * 4, st));
    const CUtensorMap mx = make_map(ds->d_tf32, dim, ds->n, (uint64_t)ds->pitch * 4, GN / tq->ctas);
    GemmParams pf = base_params(tq);
    pf.sqnorm = ds->d_sqnorm;
    pf.rnorm = ds->d_lo;
    pf.nrows = ds->n;
    pf.row_stride = 1;
    pf.tau = d_tau;
    pf.cand_cnt = tq->cnt.as<uint32_t>();
    pf.cand = cand.as<uint64_t>();
    pf.cap = cap;
    plan_gemm(pf, tq->ctas);
    // The contraction is tensor-bound and leaves HBM idle, the exact rerank of its candidates is an HBM-bound gather:
    // the pass is cut into row parts (slab ranges, so nothing is streamed twice) and the rerank of part i runs on a
    // side stream under the contraction of part i + 1 (the rerank CTAs fit next to the one persistent contraction
    // CTA per SM: 192 threads x <= 87 registers and no free shared memory needed). A part's candidates are the entries
    // [snap[i-1][q], snap[i][q]) of every query's list (counter snapshots taken between the launches); their exact
    // keys are written back over the entries they came from, so the final selection reads one [nq][cap] array.
    const uint32_t total_slabs = pf.nslabs;
    const char* parts_s = getenv("VDB_GEMM_PARTS");   // read per call: the probe scripts sweep it
    const uint32_t parts_env = parts_s ? (uint32_t)atoi(parts_s) : 0;
    const uint64_t tile_units = (uint64_t)pf.ntiles * pf.nqt;   // 256 x 256 (or 128 x 256) score tiles of the pass
    // measured (scripts/probe_parts.py, probe_shard_phases.py): 1M x 960, 10k queries 39.2 -> 38.6 ms with 3 parts (the
    // step is power-capped, so hiding the gathers slows the contraction by almost as much); a 125k-row shard with its
    // own thresholds 9.6 -> 8.0 ms; the same shard under the global thresholds of an 8-way split 4.87 -> 4.77 ms;
    // 8 parts cost more in launch tails than they hide
    uint32_t parts = parts_env ? parts_env : (tile_units >= 8192 ? 3u : 1u);
    parts = std::max(1u, std::min(parts, total_slabs));
    static thread_local SideStream side;
    if (parts > 1) side.ensure(ds->device);
    cudaStream_t rs = parts > 1 ? side.s : st;   // rerank stream
    DevBuf snaps((size_t)parts * nq * 4, st);
    const uint64_t total = (uint64_t)nq * cap;  // capacity bound; the live counts stay on the device
    DevBuf off((size_t)(nq + 1) * 8, st), qidx(total * 4, st), rid(total * 4, st), dist(total * 4, st);
    SideDrain drain{rs, parts > 1};
    if (parts > 1) {   // the side stream must not touch the scratch before the allocations above are ordered
        VDB_CUDA(cudaEventRecord(side.ev, st));
        VDB_CUDA(cudaStreamWaitEvent(rs, side.ev, 0));
    }
    for (uint32_t part = 0; part < parts; ++part) {
        const uint32_t s0 = (uint32_t)((uint64_t)total_slabs * part / parts), s1 = (uint32_t)((uint64_t)total_slabs * (part + 1) / parts);
        GemmParams pp = pf;
        pp.slab0 = s0;
        pp.nslabs = s1 - s0;
        launch_gemm(1, ds->metric, tq->mq, mx, pp, st, tq->ctas);
        uint32_t* snap = snaps.as<uint32_t>() + (size_t)part * nq;
        const uint32_t* prev = part ? snap - nq : nullptr;
        VDB_CUDA(cudaMemcpyAsync(snap, tq->cnt.p, (size_t)nq * 4, cudaMemcpyDeviceToDevice, st));
        if (parts > 1) {
            VDB_CUDA(cudaEventRecord(side.ev, st));
            VDB_CUDA(cudaStreamWaitEvent(rs, side.ev, 0));
        }
        // exact rerank of the part's candidates (compacted: only the valid pairs are touched)
        part_offsets_kernel<<<1, 1024, 0, rs>>>(prev, snap, nq, cap, off.as<uint64_t>());
        VDB_LAUNCHED();
        part_to_pairs_kernel<<<nq, 256, 0, rs>>>(cand.as<uint64_t>(), prev, snap, cap, off.as<uint64_t>(), qidx.as<uint32_t>(),
                                                 rid.as<uint32_t>());
        VDB_LAUNCHED();
        exact_pair_distances_masked(ds, tq->qcopy.p, tq->qpitch, qidx.as<uint32_t>(), rid.as<uint32_t>(), nullptr, total,
                                    dist.as<float>(), rs, off.as<uint64_t>() + nq,
                                    ds->metric == VDB_COSINE ? tq->qtile.qcache.as<float>() : nullptr);
        part_rekey_kernel<<<(uint32_t)sm_count() * 8, 256, 0, rs>>>(dist.as<float>(), qidx.as<uint32_t>(), rid.as<uint32_t>(),
                                                                   off.as<uint64_t>(), nq, prev, cap, (uint32_t)ds->id_base,
                                                                   cand.as<uint64_t>());
        VDB_LAUNCHED();
    }
    if (parts > 1) {
        VDB_CUDA(cudaEventRecord(side.ev2, rs));
        VDB_CUDA(cudaStreamWaitEvent(st, side.ev2, 0));
    }
    drain.armed = false;   // from here on the main stream is ordered after the side stream
    // the k best exact keys of every query's list (its first min(cnt, cap) entries)
    launch_merge_keys(cand.as<uint64_t>(), 1, nq, cap, false, k, d_keys, nullptr, nullptr, nullptr, st, nullptr,
                      tq->cnt.as<uint32_t>());
    if (d_overflow) {
        overflow_kernel<<<ceil_div(nq, 256u), 256, 0, st>>>(tq->cnt.as<uint32_t>(), nq, cap, d_overflow);
        VDB_LAUNCHED();
    }
    if (d_cand_total) {
        cand_total_kernel<<<1, 1024, 0, st>>>(tq->cnt.as<uint32_t>(), nq, cap, d_cand_total);
        VDB_LAUNCHED();
    }
}

// CHECK: the (merged) result of q is provably exact iff no shard overflowed and d_k - ||q||^2 < tau_q
void tensor_check(vdb_tq* tq, const uint64_t* d_keys, uint32_t k, uint64_t n_total, const float* d_tau,
                  const uint32_t* d_overflow, uint32_t* d_redo, uint32_t* d_nredo) {
    tensor_check_range(tq, 0, tq->nq, d_keys, k, n_total, d_tau, d_overflow, d_redo, d_nredo);
}
// the same for the queries [q0, q0 + cnt) of the batch: d_keys [cnt][k] and d_overflow [cnt] are local to the range,
// d_tau is the whole batch's; d_redo receives batch indices
void tensor_check_range(vdb_tq* tq, uint32_t q0, uint32_t cnt, const uint64_t* d_keys, uint32_t k, uint64_t n_total,
                        const float* d_tau, const uint32_t* d_overflow, uint32_t* d_redo, uint32_t* d_nredo) {
    VDB_CUDA(cudaMemsetAsync(d_nredo, 0, 4, tq->st));
    if (cnt == 0) return;
    check_kernel<<<ceil_div(cnt, 256u), 256, 0, tq->st>>>(d_keys, q0, cnt, k, n_total, d_overflow, d_tau,
                                                          tq->ds->metric == VDB_COSINE ? nullptr : tq->qsq.as<float>(),
                                                          g_debug_force_redo.load(), d_redo, d_nredo);
    VDB_LAUNCHED();
}

// One chunk of the batch, enqueued WITHOUT a host synchronisation: sample -> tau -> filter -> rerank -> check.
struct ChunkState {
    vdb_tq* tq = nullptr;
    DevBuf redo, nredo, ctotal;
    const void* d_queries = nullptr;
    uint64_t* d_keys = nullptr;
    uint32_t nq = 0;
    uint32_t h_redo = 0;
    uint64_t h_cands = 0;
};

static void chunk_enqueue(const vdb_dataset* ds, const void* d_queries, uint32_t nq, uint32_t k, uint64_t* d_keys,
                          cudaStream_t st, ChunkState& cs) {
    cs.d_queries = d_queries;
    cs.d_keys = d_keys;
    cs.nq = nq;
    cs.tq = tensor_begin(ds, d_queries, nq, st);
    const uint32_t j0 = tensor_j0(k, ds->sample_n, ds->n);
    DevBuf jkeys((size_t)nq * j0 * 8, st), tau((size_t)nq * 4, st), overflow((size_t)nq * 4, st);
    cs.redo = DevBuf((size_t)nq * 4, st);
    cs.nredo = DevBuf(4, st);
    cs.ctotal = DevBuf(8, st);
    tensor_sample_keys(cs.tq, j0, jkeys.as<uint64_t>());
    tensor_tau(cs.tq, jkeys.as<uint64_t>(), 1, j0, j0, ds->mean_norm, tau.as<float>());
    tensor_filter_keys(cs.tq, k, j0, tau.as<float>(), d_keys, overflow.as<uint32_t>(), cs.ctotal.as<uint64_t>());
    tensor_check(cs.tq, d_keys, k, ds->n, tau.as<float>(), overflow.as<uint32_t>(), cs.redo.as<uint32_t>(),
                 cs.nredo.as<uint32_t>());
}

// worker streams of the chunk pipeline (per host thread and device)
struct ChunkStreams {
    int device = -1;
    cudaStream_t s[2] = {nullptr, nullptr};
    cudaEvent_t fork = nullptr, join[2] = {nullptr, nullptr};
    void ensure(int dev) {
        if (device == dev) return;
        release();
        for (int i = 0; i < 2; ++i) {
            VDB_CUDA(cudaStreamCreateWithFlags(&s[i], cudaStreamNonBlocking));
            VDB_CUDA(cudaEventCreateWithFlags(&join[i], cudaEventDisableTiming));
        }
        VDB_CUDA(cudaEventCreateWithFlags(&fork, cudaEventDisableTiming));
        device = dev;
    }
    void release() {
        if (device < 0) return;
        for (int i = 0; i < 2; ++i) {
            cudaStreamDestroy(s[i]);
            cudaEventDestroy(join[i]);
        }
        cudaEventDestroy(fork);
        device = -1;
    }
    ~ChunkStreams() { release(); }
};

// Batches are cut into chunks of at most 16384 queries (bounds the candidate / rerank scratch, ~5 GB per chunk); all
// completeness checks are read back once, after the last chunk. VDB_GEMM_CHUNKS=c (c > 1) additionally splits the
// batch into c chunks that alternate between two worker streams, so the tail of one chunk (rerank gathers, merges,
// check) runs under the contraction of the next. Measured on B200 (1M x 960, 10k queries): 39.4 ms unsplit, 39.9 ms
// with 4 chunks, 55.8 ms with 8 — the rerank gathers and the contraction compete for the same L2 bandwidth and every
// chunk re-streams the database, so the pipeline is OFF by default.
constexpr uint32_t G_QUERY_CHUNK = 16384;

void flat_gemm_keys(const vdb_dataset* ds, const void* d_queries, uint32_t nq, uint32_t k, uint64_t* d_keys,
                    cudaStream_t st) {
    VDB_REQUIRE(flat_gemm_supported(ds, nq, k), "tensor-core Flat path: unsupported dataset/k");
    static const uint32_t chunks_env = getenv("VDB_GEMM_CHUNKS") ? (uint32_t)atoi(getenv("VDB_GEMM_CHUNKS")) : 0;
    uint32_t csize = (chunks_env > 1 && nq >= 4096) ? std::max(1024u, round_up(ceil_div(nq, chunks_env), 256u)) : nq;
    csize = std::min(csize, G_QUERY_CHUNK);
    const uint32_t nchunks = ceil_div(nq, csize);
    const size_t row_bytes = (size_t)ds->dim * ds->elem_size();
    std::vector<ChunkState> cs(nchunks);
    static thread_local ChunkStreams ws;
    const bool pipelined = nchunks > 1 && chunks_env > 1;
    auto cleanup = [&]() {
        for (auto& c : cs)
            if (c.tq) tensor_end(c.tq), c.tq = nullptr;
    };
    try {
        if (pipelined) {
            ws.ensure(ds->device);
            VDB_CUDA(cudaEventRecord(ws.fork, st));
            for (int i = 0; i < 2; ++i) VDB_CUDA(cudaStreamWaitEvent(ws.s[i], ws.fork, 0));
        }
        for (uint32_t c = 0; c < nchunks; ++c) {
            const uint32_t q0 = c * csize, cn = std::min(csize, nq - q0);
            chunk_enqueue(ds, (const uint8_t*)d_queries + (size_t)q0 * row_bytes, cn, k, d_keys + (size_t)q0 * k,
                          pipelined ? ws.s[c & 1] : st, cs[c]);
        }
        if (pipelined)
            for (int i = 0; i < 2; ++i) {
                VDB_CUDA(cudaEventRecord(ws.join[i], ws.s[i]));
                VDB_CUDA(cudaStreamWaitEvent(st, ws.join[i], 0));
            }
        for (auto& c : cs) {
            VDB_CUDA(cudaMemcpyAsync(&c.h_redo, c.nredo.p, 4, cudaMemcpyDeviceToHost, st));
            VDB_CUDA(cudaMemcpyAsync(&c.h_cands, c.ctotal.p, 8, cudaMemcpyDeviceToHost, st));
        }
        VDB_CUDA(cudaStreamSynchronize(st));
        for (auto& c : cs) {
            g_gemm_redo += c.h_redo;
            g_gemm_cands += c.h_cands;
            g_gemm_queries += c.nq;
            if (c.h_redo) {  // exact streaming scan for the queries whose candidate set could not be proven complete
                DevBuf rq((size_t)c.h_redo * row_bytes, st), rkeys((size_t)c.h_redo * k * 8, st);
                gather_rows_kernel<<<c.h_redo, 128, 0, st>>>((const uint8_t*)c.d_queries, (uint32_t)row_bytes, c.redo.as<uint32_t>(),
                                                             c.h_redo, rq.as<uint8_t>());
                VDB_LAUNCHED();
                flat_scan_keys(ds, rq.p, c.h_redo, k, rkeys.as<uint64_t>(), st);
                scatter_keys_kernel<<<c.h_redo, 128, 0, st>>>(rkeys.as<uint64_t>(), k, c.redo.as<uint32_t>(), c.h_redo, c.d_keys);
                VDB_LAUNCHED();
            }
        }
    } catch (...) {
        cudaDeviceSynchronize();  // worker streams may still hold work that references the chunk scratch
        cleanup();
        throw;
    }
    cleanup();
}


// ---- IVF probe scan on the tensor cores ---------------------------------------------------------------------------
// A batch of queries probing nlist lists is a block-sparse contraction: list l's rows x the queries that probe l.
// Rows and norms are kept a second time in LIST ORDER (position p <-> members[p]) so every list is a contiguous row
// range the TMA can tile; the queries are gathered list by list into one matrix; the work items of the contraction
// kernel are (list, tile of gathered queries, range of row tiles). Query tiles of <= 128 rows run on single CTAs
// (M = 128), larger ones on CTA pairs (M = 256). Thresholds:
//   k/16 small (j0 <= 16): a stratified 1/16 sample of every list (also kept in list order) is scored first, the
//     j0-th smallest sampled S' over the probed lists (+ margin) is tau, as in the Flat path;
//   else: tau = the k-th smallest EXACT distance over the first S rows of the visit sequence (complete by construction).
// Candidates are reranked with the FP32 list scan's arithmetic; queries failing the completeness check or overflowing
// their candidate list are redone by the FP32 list scan.
constexpr uint32_t IVF_SAMPLE_RATE = 16;

__global__ void gather_round_rows_kernel(const float* __restrict__ rows_tf32, const float* __restrict__ colA,
                                         const float* __restrict__ rn, const uint32_t* __restrict__ members, uint64_t n,
                                         uint32_t pitch, float* __restrict__ out, float* __restrict__ outA,
                                         float* __restrict__ outR) {
    const uint64_t p = blockIdx.x;
    if (p >= n) return;
    const uint64_t row = members[p];
    const float4* src = reinterpret_cast<const float4*>(rows_tf32 + row * pitch);
    float4* dst = reinterpret_cast<float4*>(out + p * pitch);
    for (uint32_t e = threadIdx.x; e < pitch / 4; e += blockDim.x) dst[e] = src[e];
    if (threadIdx.x == 0) {
        outA[p] = colA[row];
        outR[p] = rn[row];
    }
}
// sample row s of list l = one hashed position out of the s-th bucket of IVF_SAMPLE_RATE consecutive list positions
__global__ void ivf_sample_gather_kernel(const float* __restrict__ rows_lo, const float* __restrict__ colA_lo,
                                         const float* __restrict__ rn_lo, const uint64_t* __restrict__ offsets,
                                         const uint64_t* __restrict__ soff, uint32_t nlist, uint32_t pitch,
                                         float* __restrict__ out, float* __restrict__ outA, float* __restrict__ outR) {
    const uint64_t s = blockIdx.x;
    uint32_t lo = 0, hi = nlist;  // largest l with soff[l] <= s
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) / 2;
        if (soff[mid] <= s) lo = mid;
        else hi = mid;
    }
    const uint64_t b0 = offsets[lo] + (s - soff[lo]) * IVF_SAMPLE_RATE;
    const uint64_t b1 = min(offsets[lo + 1], b0 + IVF_SAMPLE_RATE);
    uint64_t h = (s + 0x9E3779B97F4A7C15ull) * 0xBF58476D1CE4E5B9ull;
    h ^= h >> 31;
    h *= 0x94D049BB133111EBull;
    h ^= h >> 29;
    const uint64_t pos = b0 + h % (b1 - b0);
    const float4* src = reinterpret_cast<const float4*>(rows_lo + pos * pitch);
    float4* dst = reinterpret_cast<float4*>(out + s * pitch);
    for (uint32_t e = threadIdx.x; e < pitch / 4; e += blockDim.x) dst[e] = src[e];
    if (threadIdx.x == 0) {
        outA[s] = colA_lo[pos];
        outR[s] = rn_lo[pos];
    }
}

static std::mutex g_ivf_side_mu;
static void ensure_ivf_side(const vdb_dataset* ds, const vdb_ivf* civf, const std::vector<uint64_t>& h_off, cudaStream_t st) {
    vdb_ivf* ivf = const_cast<vdb_ivf*>(civf);
    std::lock_guard<std::mutex> lk(g_ivf_side_mu);
    if (ivf->d_rows_lo) return;
    ensure_side_arrays(ds, st);
    float *rows = nullptr, *colA = nullptr, *rn = nullptr;
    VDB_CUDA(cudaMalloc(&rows, ds->n * (size_t)ds->pitch * 4));
    VDB_CUDA(cudaMalloc(&colA, ds->n * 4));
    VDB_CUDA(cudaMalloc(&rn, ds->n * 4));
    gather_round_rows_kernel<<<(uint32_t)ds->n, 128, 0, st>>>(ds->d_tf32, ds->d_sqnorm, ds->d_lo, ivf->d_members, ds->n, ds->pitch,
                                                             rows, colA, rn);
    VDB_LAUNCHED();
    ivf->h_samp_off.assign(ivf->nlist + 1, 0);
    for (uint32_t l = 0; l < ivf->nlist; ++l)
        ivf->h_samp_off[l + 1] = ivf->h_samp_off[l] + ceil_div<uint64_t>(h_off[l + 1] - h_off[l], IVF_SAMPLE_RATE);
    ivf->samp_n = ivf->h_samp_off[ivf->nlist];
    VDB_CUDA(cudaMalloc(&ivf->d_samp_rows, std::max<uint64_t>(ivf->samp_n, 1) * (size_t)ds->pitch * 4));
    VDB_CUDA(cudaMalloc(&ivf->d_samp_colA, std::max<uint64_t>(ivf->samp_n, 1) * 4));
    VDB_CUDA(cudaMalloc(&ivf->d_samp_rn, std::max<uint64_t>(ivf->samp_n, 1) * 4));
    if (ivf->samp_n) {
        DevBuf soff((size_t)(ivf->nlist + 1) * 8, st);
        VDB_CUDA(cudaMemcpyAsync(soff.p, ivf->h_samp_off.data(), (size_t)(ivf->nlist + 1) * 8, cudaMemcpyHostToDevice, st));
        ivf_sample_gather_kernel<<<(uint32_t)ivf->samp_n, 128, 0, st>>>(rows, colA, rn, ivf->d_offsets, soff.as<uint64_t>(),
                                                                       ivf->nlist, ds->pitch, ivf->d_samp_rows,
                                                                       ivf->d_samp_colA, ivf->d_samp_rn);
        VDB_LAUNCHED();
        VDB_CUDA(cudaStreamSynchronize(st));
    }
    VDB_CUDA(cudaStreamSynchronize(st));
    ivf->d_colA_lo = colA;
    ivf->d_rn_lo = rn;
    ivf->d_rows_lo = rows;
}

// first S rows of every query's visit sequence (probed lists in probe order) as rerank pairs
__global__ void ivf_subset_kernel(const uint64_t* __restrict__ probes, uint32_t nprobe, const uint64_t* __restrict__ offsets,
                                  const uint32_t* __restrict__ members, uint32_t S, uint32_t* __restrict__ qidx,
                                  uint32_t* __restrict__ rid, uint8_t* __restrict__ valid) {
    const uint32_t q = blockIdx.x;
    for (uint32_t s = threadIdx.x; s < S; s += blockDim.x) {
        uint32_t left = s, row = 0;
        bool ok = false;
        for (uint32_t j = 0; j < nprobe && !ok; ++j) {
            const uint64_t pk = probes[(size_t)q * nprobe + j];
            if (pk == KEY_NONE) break;
            const uint32_t c = key_id(pk);
            const uint64_t len = offsets[c + 1] - offsets[c];
            if (left < len) {
                row = members[offsets[c] + left];
                ok = true;
            } else {
                left -= (uint32_t)len;
            }
        }
        qidx[(size_t)q * S + s] = q;
        rid[(size_t)q * S + s] = row;
        valid[(size_t)q * S + s] = ok;
    }
}
// tau' in the units of the pruning score: L2Sqr: d_k - ||q||^2, cosine: d_k; +inf when fewer than k rows are probed
__global__ void ivf_tau_kernel(const uint64_t* __restrict__ keys, uint32_t nq, uint32_t k, const float* __restrict__ qsq,
                               float* __restrict__ tau) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    const uint64_t kk = keys[(size_t)q * k + (k - 1)];
    if (kk == KEY_NONE) {
        tau[q] = __uint_as_float(0x7f800000u);
        return;
    }
    const float dk = key_dist(kk);
    const float shift = qsq ? qsq[q] : 0.f;
    tau[q] = (dk - shift) + 2e-5f * (fabsf(dk) + shift) + 1e-30f;
}
// the 16 smallest sampled scores of every probed list of a query, side by side: [nq][nprobe][G_TOPJ]
__global__ void ivf_collect_sample_kernel(const uint64_t* __restrict__ skeys, const uint32_t* __restrict__ gpos, uint32_t nprobe,
                                          uint64_t* __restrict__ out) {
    const uint32_t q = blockIdx.x;
    for (uint32_t e = threadIdx.x; e < nprobe * G_TOPJ; e += blockDim.x) {
        const uint32_t g = gpos[(size_t)q * nprobe + e / G_TOPJ];
        out[(size_t)q * nprobe * G_TOPJ + e] = g == 0xffffffffu ? KEY_NONE : skeys[(size_t)g * G_TOPJ + e % G_TOPJ];
    }
}
// tau = S'_(j0) of the sample + margin; +inf when the probed lists hold fewer than j0 sampled rows
__global__ void ivf_sample_tau_kernel(const uint64_t* __restrict__ keys, uint32_t nq, uint32_t j0, const float* __restrict__ qcm,
                                      float margin, float* __restrict__ tau) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    const uint64_t kk = keys[(size_t)q * j0 + (j0 - 1)];
    const float s = kk == KEY_NONE ? __uint_as_float(0x7f800000u) : key_dist(kk);
    tau[q] = s + 2.5f * (qcm ? qcm[q] * margin : margin);
}
__global__ void gather_query_side_kernel(const float* __restrict__ qround, uint32_t qpitch, const float* __restrict__ qcm,
                                         const float* __restrict__ qnorm, const uint32_t* __restrict__ qmap, uint32_t G,
                                         float* __restrict__ outq, float* __restrict__ out_qcm, float* __restrict__ out_qnorm) {
    const uint32_t g = blockIdx.x;
    if (g >= G) return;
    const uint32_t q = qmap[g];
    const float4* src = reinterpret_cast<const float4*>(qround + (size_t)q * qpitch);
    float4* dst = reinterpret_cast<float4*>(outq + (size_t)g * qpitch);
    for (uint32_t e = threadIdx.x; e < qpitch / 4; e += blockDim.x) dst[e] = src[e];
    if (threadIdx.x == 0) {
        out_qcm[g] = qcm[q];
        if (qnorm) out_qnorm[g] = qnorm[q];
    }
}
__global__ void gather_f32_kernel(const float* __restrict__ src, const uint32_t* __restrict__ idx, uint32_t cnt,
                                  float* __restrict__ dst) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < cnt) dst[i] = src[idx[i]];
}
// completeness: no overflow, and (unless the threshold is complete by construction) the k-th exact distance lies
// strictly inside the threshold; tau = +inf keeps every probed row, so short visit sets pass
__global__ void ivf_check_kernel(const uint64_t* __restrict__ keys, uint32_t nq, uint32_t k, const uint32_t* __restrict__ cnt,
                                 uint32_t cap, const float* __restrict__ tau, const float* __restrict__ qsq, int by_construction,
                                 uint32_t* __restrict__ redo, uint32_t* __restrict__ nredo) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    bool ok = cnt[q] <= cap;
    if (ok && !by_construction && tau[q] != __uint_as_float(0x7f800000u)) {
        const uint64_t kk = keys[(size_t)q * k + (k - 1)];
        ok = kk != KEY_NONE;
        if (ok) {
            const float dk = key_dist(kk);
            const float shift = qsq ? qsq[q] : 0.f;
            const float slack = 2e-5f * (fabsf(dk) + shift + fabsf(tau[q]));
            ok = (dk - shift) < tau[q] - slack;
        }
    }
    if (!ok) redo[atomicAdd(nredo, 1u)] = q;
}

bool ivf_tensor_keys(const vdb_dataset* ds, const vdb_ivf* ivf, const void* d_queries, const uint64_t* d_probes,
                     const std::vector<uint64_t>& h_probes, const std::vector<uint64_t>& h_off, uint32_t nq, uint32_t nprobe,
                     uint32_t k, uint64_t* d_keys, cudaStream_t st) {
    if (!flat_gemm_supported(ds, nq, k) || k > 512) return false;
    ensure_ivf_side(ds, ivf, h_off, st);
    // ---- host: group the (query, list) pairs by list, build the gathered query order and the work items ----
    std::vector<std::vector<uint32_t>> by_list(ivf->nlist);   // entries: q * nprobe + j
    for (uint32_t q = 0; q < nq; ++q)
        for (uint32_t j = 0; j < nprobe; ++j) {
            const uint64_t pk = h_probes[(size_t)q * nprobe + j];
            if (pk != KEY_NONE) by_list[key_id(pk)].push_back(q * nprobe + j);
        }
    std::vector<uint32_t> qmap, gpos((size_t)nq * nprobe, 0xffffffffu);
    std::vector<GemmItem> items[2], sitems[2];   // [0]: single CTAs (<= 128 gathered queries), [1]: CTA pairs
    qmap.reserve((size_t)nq * nprobe);
    static const uint32_t force_ctas = getenv("VDB_IVF_CTAS") ? (uint32_t)atoi(getenv("VDB_IVF_CTAS")) : 0;
    const uint32_t tiles_per_item = 8;
    for (uint32_t l = 0; l < ivf->nlist; ++l) {
        const uint64_t r0 = h_off[l], r1 = h_off[l + 1];
        if (r1 == r0 || by_list[l].empty()) continue;
        const uint32_t g0 = (uint32_t)qmap.size();
        for (uint32_t e : by_list[l]) {
            gpos[e] = (uint32_t)qmap.size();
            qmap.push_back(e / nprobe);
        }
        const uint32_t g1 = (uint32_t)qmap.size();
        const uint64_t s0 = ivf->h_samp_off[l], s1 = ivf->h_samp_off[l + 1];
        // query tiles: 256 rows on a CTA pair, a remainder of <= 128 rows on a single CTA
        for (uint32_t g = g0; g < g1;) {
            const uint32_t left = g1 - g;
            const uint32_t pair = force_ctas ? force_ctas - 1 : (left > GM ? 1u : 0u);
            const uint32_t len = std::min(left, pair ? 2u * GM : (uint32_t)GM);
            for (uint64_t r = r0; r < r1; r += (uint64_t)tiles_per_item * GN)
                items[pair].push_back(GemmItem{g, g + len, r, std::min<uint64_t>(r1, r + (uint64_t)tiles_per_item * GN)});
            sitems[pair].push_back(GemmItem{g, g + len, s0, s1});
            g += len;
        }
    }
    if (qmap.empty()) {
        VDB_CUDA(cudaMemsetAsync(d_keys, 0xff, (size_t)nq * k * 8, st));
        return true;
    }
    const uint32_t G = (uint32_t)qmap.size();
    vdb_tq* tq = tensor_begin(ds, d_queries, nq, st);
    try {
        const bool cosine = ds->metric == VDB_COSINE;
        // ---- gathered query matrix, its per-row scalars, item tables ----
        DevBuf d_qmap((size_t)G * 4, st), qg((size_t)G * tq->qpitch * 4, st), qcm_g((size_t)G * 4, st), qn_g((size_t)G * 4, st),
            tau_g((size_t)G * 4, st), tau((size_t)nq * 4, st);
        VDB_CUDA(cudaMemcpyAsync(d_qmap.p, qmap.data(), (size_t)G * 4, cudaMemcpyHostToDevice, st));
        gather_query_side_kernel<<<G, 128, 0, st>>>(tq->qround.as<float>(), tq->qpitch, tq->qcm.as<float>(),
                                                    cosine ? tq->qtile.qcache.as<float>() : nullptr, d_qmap.as<uint32_t>(), G,
                                                    qg.as<float>(), qcm_g.as<float>(), qn_g.as<float>());
        VDB_LAUNCHED();
        const CUtensorMap mq = make_map(qg.as<float>(), ds->dim, G, (uint64_t)tq->qpitch * 4, GM);
        GemmParams base{};
        base.nq = G;
        base.kblocks = ceil_div(ds->dim, (uint32_t)GK);
        base.qcm = qcm_g.as<float>();
        base.qnorm = cosine ? qn_g.as<float>() : nullptr;
        base.kc = tq->kc;
        base.row_stride = 1;
        base.qmap = d_qmap.as<uint32_t>();
        base.nslabs = 1;
        auto run_items = [&](int mode, const std::vector<GemmItem>* tabs, const float* rows, uint64_t nrows, GemmParams p) {
            for (uint32_t pair = 0; pair < 2; ++pair) {
                if (tabs[pair].empty()) continue;
                DevBuf d_items(tabs[pair].size() * sizeof(GemmItem), st);
                VDB_CUDA(cudaMemcpyAsync(d_items.p, tabs[pair].data(), tabs[pair].size() * sizeof(GemmItem),
                                         cudaMemcpyHostToDevice, st));
                const CUtensorMap mx = make_map(rows, ds->dim, nrows, (uint64_t)ds->pitch * 4, GN / (pair + 1));
                p.nrows = nrows;
                p.items = d_items.as<GemmItem>();
                p.nitems = (uint32_t)tabs[pair].size();
                launch_gemm(mode, ds->metric, mq, mx, p, st, (int)pair + 1);
            }
        };
        // ---- thresholds ----
        const uint32_t j0 = tensor_j0(k, 1u << 20, (uint64_t)IVF_SAMPLE_RATE << 20);
        static const int force_subset = getenv("VDB_IVF_SUBSET") ? atoi(getenv("VDB_IVF_SUBSET")) : 0;
        const bool by_sample = j0 <= (uint32_t)G_TOPJ && ivf->samp_n > 0 && !force_subset;
        if (by_sample) {
            DevBuf skeys((size_t)G * G_TOPJ * 8, st), d_gpos((size_t)nq * nprobe * 4, st),
                ckeys((size_t)nq * nprobe * G_TOPJ * 8, st), jkeys((size_t)nq * j0 * 8, st);
            VDB_CUDA(cudaMemsetAsync(skeys.p, 0xff, (size_t)G * G_TOPJ * 8, st));
            VDB_CUDA(cudaMemcpyAsync(d_gpos.p, gpos.data(), (size_t)nq * nprobe * 4, cudaMemcpyHostToDevice, st));
            GemmParams ps = base;
            ps.sqnorm = ivf->d_samp_colA;
            ps.rnorm = ivf->d_samp_rn;
            ps.out_keys = skeys.as<uint64_t>();
            run_items(2, sitems, ivf->d_samp_rows, ivf->samp_n, ps);
            ivf_collect_sample_kernel<<<nq, 128, 0, st>>>(skeys.as<uint64_t>(), d_gpos.as<uint32_t>(), nprobe, ckeys.as<uint64_t>());
            VDB_LAUNCHED();
            launch_merge_keys(ckeys.as<uint64_t>(), 1, nq, nprobe * G_TOPJ, false, j0, jkeys.as<uint64_t>(), nullptr, nullptr,
                              nullptr, st);
            ivf_sample_tau_kernel<<<ceil_div(nq, 256u), 256, 0, st>>>(jkeys.as<uint64_t>(), nq, j0,
                                                                     cosine ? nullptr : tq->qcm.as<float>(),
                                                                     cosine ? tq->cbound : ds->mean_norm, tau.as<float>());
            VDB_LAUNCHED();
        } else {
            // S balances the two exact-distance passes: S subset rows against ~k * visited / S surviving candidates
            const double visited = (double)nprobe * (double)ds->n / (double)ivf->nlist;
            const uint32_t S = force_subset > 1
                                   ? (uint32_t)force_subset
                                   : std::min(4096u, (uint32_t)next_pow2(std::max<uint32_t>(
                                                         std::max(256u, 2 * k), (uint32_t)std::sqrt((double)k * visited))));
            const uint64_t cnt = (uint64_t)nq * S;
            DevBuf qidx(cnt * 4, st), rid(cnt * 4, st), valid(cnt, st), dist(cnt * 4, st), skeys(cnt * 8, st), kkeys((size_t)nq * k * 8, st);
            ivf_subset_kernel<<<nq, 128, 0, st>>>(d_probes, nprobe, ivf->d_offsets, ivf->d_members, S, qidx.as<uint32_t>(),
                                                  rid.as<uint32_t>(), valid.as<uint8_t>());
            VDB_LAUNCHED();
            exact_pair_distances_masked(ds, tq->qcopy.p, tq->qpitch, qidx.as<uint32_t>(), rid.as<uint32_t>(), valid.as<uint8_t>(),
                                        cnt, dist.as<float>(), st, nullptr, cosine ? tq->qtile.qcache.as<float>() : nullptr);
            rekey_based(dist.as<float>(), rid.as<uint32_t>(), 0, valid.as<uint8_t>(), cnt, skeys.as<uint64_t>(), st);
            launch_merge_keys(skeys.as<uint64_t>(), 1, nq, S, false, k, kkeys.as<uint64_t>(), nullptr, nullptr, nullptr, st);
            ivf_tau_kernel<<<ceil_div(nq, 256u), 256, 0, st>>>(kkeys.as<uint64_t>(), nq, k, cosine ? nullptr : tq->qsq.as<float>(),
                                                              tau.as<float>());
            VDB_LAUNCHED();
        }
        gather_f32_kernel<<<ceil_div(G, 256u), 256, 0, st>>>(tau.as<float>(), d_qmap.as<uint32_t>(), G, tau_g.as<float>());
        VDB_LAUNCHED();
        // ---- filter pass over the probed (list, query tile) blocks ----
        const uint32_t cap = 8192;
        DevBuf cand((size_t)nq * cap * 8, st);
        VDB_CUDA(cudaMemsetAsync(tq->cnt.p, 0, (size_t)nq * 4, st));
        {
            GemmParams pf = base;
            pf.sqnorm = ivf->d_colA_lo;
            pf.rnorm = ivf->d_rn_lo;
            pf.tau = tau_g.as<float>();
            pf.cand_cnt = tq->cnt.as<uint32_t>();
            pf.cand = cand.as<uint64_t>();
            pf.cap = cap;
            run_items(1, items, ivf->d_rows_lo, ds->n, pf);
        }
        // ---- exact rerank (the FP32 list scan's arithmetic) and top-k ----
        const uint64_t total = (uint64_t)nq * cap;
        DevBuf off((size_t)(nq + 1) * 8, st), qidx(total * 4, st), rid(total * 4, st), dist(total * 4, st), keys2(total * 8, st);
        cand_offsets_kernel<<<1, 1024, 0, st>>>(tq->cnt.as<uint32_t>(), nq, cap, off.as<uint64_t>());
        VDB_LAUNCHED();
        cand_to_pairs_kernel<<<nq, 256, 0, st>>>(cand.as<uint64_t>(), tq->cnt.as<uint32_t>(), cap, off.as<uint64_t>(),
                                                 ivf->d_members, qidx.as<uint32_t>(), rid.as<uint32_t>());
        VDB_LAUNCHED();
        const uint64_t* d_total = off.as<uint64_t>() + nq;
        exact_pair_distances_masked(ds, tq->qcopy.p, tq->qpitch, qidx.as<uint32_t>(), rid.as<uint32_t>(), nullptr, total,
                                    dist.as<float>(), st, d_total, cosine ? tq->qtile.qcache.as<float>() : nullptr);
        rekey_dev_count_kernel<<<(uint32_t)sm_count() * 8, 256, 0, st>>>(dist.as<float>(), rid.as<uint32_t>(),
                                                                        (uint32_t)ds->id_base, d_total, keys2.as<uint64_t>());
        VDB_LAUNCHED();
        launch_merge_keys(keys2.as<uint64_t>(), 1, nq, cap, false, k, d_keys, nullptr, nullptr, nullptr, st, off.as<uint64_t>());
        // ---- completeness check; failing queries: FP32 list scan ----
        DevBuf redo((size_t)nq * 4, st), nredo(4, st);
        VDB_CUDA(cudaMemsetAsync(nredo.p, 0, 4, st));
        ivf_check_kernel<<<ceil_div(nq, 256u), 256, 0, st>>>(d_keys, nq, k, tq->cnt.as<uint32_t>(), cap, tau.as<float>(),
                                                            cosine ? nullptr : tq->qsq.as<float>(), by_sample ? 0 : 1,
                                                            redo.as<uint32_t>(), nredo.as<uint32_t>());
        VDB_LAUNCHED();
        uint32_t h_n = 0;
        uint64_t h_cands = 0;
        VDB_CUDA(cudaMemcpyAsync(&h_n, nredo.p, 4, cudaMemcpyDeviceToHost, st));
        VDB_CUDA(cudaMemcpyAsync(&h_cands, d_total, 8, cudaMemcpyDeviceToHost, st));
        VDB_CUDA(cudaStreamSynchronize(st));
        g_gemm_redo += h_n;
        g_gemm_cands += h_cands;
        g_gemm_queries += nq;
        if (h_n) {
            std::vector<uint32_t> sel(h_n);
            VDB_CUDA(cudaMemcpyAsync(sel.data(), redo.p, (size_t)h_n * 4, cudaMemcpyDeviceToHost, st));
            VDB_CUDA(cudaStreamSynchronize(st));
            std::sort(sel.begin(), sel.end());
            DevBuf rk((size_t)h_n * k * 8, st), dsel((size_t)h_n * 4, st);
            ivf_list_major_subset(ds, ivf, d_queries, d_probes, sel.data(), h_n, nprobe, k, rk.as<uint64_t>(), st);
            VDB_CUDA(cudaMemcpyAsync(dsel.p, sel.data(), (size_t)h_n * 4, cudaMemcpyHostToDevice, st));
            scatter_keys_kernel<<<h_n, 128, 0, st>>>(rk.as<uint64_t>(), k, dsel.as<uint32_t>(), h_n, d_keys);
            VDB_LAUNCHED();
            VDB_CUDA(cudaStreamSynchronize(st));
        }
    } catch (...) {
        tensor_end(tq);
        throw;
    }
    tensor_end(tq);
    return true;
}

// debug / test entry: S' keys of every (query, sampled row) pair, [nq][nrows] (mode 0 of the kernel)
void flat_gemm_store(const vdb_dataset* ds, const void* d_queries, uint32_t nq, uint32_t row_stride, float c,
                     uint64_t* d_out_keys, cudaStream_t st) {
    VDB_REQUIRE(ds->dtype == VDB_F32 && ds->metric == VDB_L2SQR && ds->dim % 4 == 0 && ((uintptr_t)d_queries & 15) == 0 &&
                    row_stride >= 1,
                "flat_gemm_store: f32 L2Sqr rows, dim %% 4 == 0 and aligned queries only");
    ensure_side_arrays(ds, st);
    const uint64_t ns = ds->n / row_stride;
    DevBuf qsq((size_t)nq * 4, st), qcm((size_t)nq * 4, st);
    vdb_dataset qd = *ds;
    qd.d_rows = const_cast<void*>(d_queries);
    qd.n = nq;
    qd.pitch = ds->dim;
    qd.metric = VDB_L2SQR;
    row_cache(&qd, qsq.as<float>(), st);
    qcm_kernel<<<ceil_div(nq, 256u), 256, 0, st>>>(qsq.as<float>(), nq, c, qcm.as<float>());
    VDB_LAUNCHED();
    GemmParams p{};
    p.nq = nq;
    p.kblocks = ceil_div(ds->dim, (uint32_t)GK);
    p.sqnorm = ds->d_sqnorm;
    p.rnorm = ds->d_lo;
    p.qcm = qcm.as<float>();
    p.nrows = ns;
    p.row_stride = row_stride;
    p.out_keys = d_out_keys;
    const CUtensorMap mq = make_map(d_queries, ds->dim, nq, (uint64_t)ds->dim * 4, GM);
    const CUtensorMap ms = make_map(ds->d_rows, ds->dim, ns, ds->pitch_bytes() * row_stride, GN / gemm_ctas());
    plan_gemm(p);
    launch_gemm(0, VDB_L2SQR, mq, ms, p, st);
}

}  // namespace vdb

// common.cuh — shared host/device helpers for the sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <stdexcept>
#include <string>

#include "../../include/vdb_b200.h"

namespace vdb {

// ---------------------------------------------------------------------------------------------
// error plumbing: kernels/launchers throw, the C ABI layer catches and returns a status code
// ---------------------------------------------------------------------------------------------
struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

[[noreturn]] inline void fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    throw Error(code, buf);
}

#define VDB_CUDA(expr)                                                                      \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess)                                                              \
            ::vdb::fail(_e == cudaErrorMemoryAllocation ? VDB_ENOMEM : VDB_ECUDA,           \
                        "CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__, __LINE__, \
                        cudaGetErrorString(_e));                                            \
    } while (0)

#define VDB_REQUIRE(cond, ...)                              \
    do {                                                    \
        if (!(cond)) ::vdb::fail(VDB_EINVAL, __VA_ARGS__);  \
    } while (0)

extern std::atomic<uint64_t> g_launches;
#define VDB_LAUNCHED()                      \
    do {                                    \
        ::vdb::g_launches.fetch_add(1);     \
        VDB_CUDA(cudaGetLastError());       \
    } while (0)

// ---- optional per-kernel timing (bench.py's roofline leg): CUDA events on the launching stream ----
void prof_record(const char* name, cudaStream_t st, bool begin);
extern bool g_prof_on;
// page-locked staging memory that grows and is reused (one per host thread and purpose: thread_local). Copies from / to
// pageable vectors are staged by the driver and serialise with the host; small per-call tables go through this instead.
struct PinnedStage {
    void* p = nullptr;
    size_t cap = 0;
    void* get(size_t bytes) {
        if (bytes > cap) {
            if (p) cudaFreeHost(p);
            p = nullptr;
            cap = 0;
            VDB_CUDA(cudaHostAlloc(&p, bytes * 2, cudaHostAllocPortable));
            cap = bytes * 2;
        }
        return p;
    }
    ~PinnedStage() {
        if (p) cudaFreeHost(p);
    }
};

struct ProfScope {
    const char* name;
    cudaStream_t st;
    ProfScope(const char* n, cudaStream_t s) : name(n), st(s) {
        if (g_prof_on) prof_record(name, st, true);
    }
    ~ProfScope() {
        if (g_prof_on) prof_record(name, st, false);
    }
};

// keep freed scratch memory in the stream-ordered pool instead of returning it to the driver on
// every synchronisation (the default release threshold is 0)
inline void pool_init() {
    static thread_local int done_dev = -1;
    int dev;
    if (cudaGetDevice(&dev) != cudaSuccess || dev == done_dev) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        uint64_t keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    done_dev = dev;
}

// stream-ordered scratch allocation (re-entrant: every call owns its workspace)
struct DevBuf {
    void* p = nullptr;
    cudaStream_t s = nullptr;
    DevBuf() = default;
    DevBuf(size_t bytes, cudaStream_t st) : s(st) {
        pool_init();
        if (bytes) VDB_CUDA(cudaMallocAsync(&p, bytes, st));
    }
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), s(o.s) { o.p = nullptr; }
    DevBuf& operator=(DevBuf&& o) noexcept {
        release();
        p = o.p;
        s = o.s;
        o.p = nullptr;
        return *this;
    }
    void release() {
        if (p) cudaFreeAsync(p, s);
        p = nullptr;
    }
    ~DevBuf() { release(); }
    template <class T> T* as() const { return static_cast<T*>(p); }
};

inline int sm_count() {
    static thread_local int cached_dev = -1, cached = 0;
    int dev;
    VDB_CUDA(cudaGetDevice(&dev));
    if (dev != cached_dev) {
        VDB_CUDA(cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev));
        cached_dev = dev;
    }
    return cached;
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) applies to ONE device: the largest size configured so far is
// remembered per (kernel instantiation, device) - `state` is a function-local static array of the caller - so a
// process that drives several GPUs (vdb_init) configures every kernel once on each of them.
constexpr int VDB_MAX_DEVICES = 64;
template <class K>
inline void ensure_dyn_smem(K kern, size_t bytes, std::atomic<size_t> (&state)[VDB_MAX_DEVICES]) {
    int dev = 0;
    VDB_CUDA(cudaGetDevice(&dev));
    std::atomic<size_t>& s = state[dev & (VDB_MAX_DEVICES - 1)];
    if (bytes <= s.load(std::memory_order_acquire)) return;
    VDB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    size_t cur = s.load();
    while (cur < bytes && !s.compare_exchange_weak(cur, bytes)) {}
}

inline uint32_t next_pow2(uint32_t v) {
    uint32_t p = 1;
    while (p < v) p <<= 1;
    return p;
}
template <class T> __host__ __device__ inline T ceil_div(T a, T b) { return (a + b - 1) / b; }
template <class T> __host__ __device__ inline T round_up(T a, T b) { return ceil_div(a, b) * b; }

// ---------------------------------------------------------------------------------------------
// sortable keys: (distance, id) -> u64 whose integer order equals CandidatePair's order
// (reference src/index_algorithm/candidate_pair.rs:36-40 with ordered-float's total order:
//  -0 == +0, NaN greater than everything).
// ---------------------------------------------------------------------------------------------
constexpr uint64_t KEY_NONE = 0xFFFFFFFFFFFFFFFFull;

__host__ __device__ __forceinline__ uint32_t f32_order_bits(float d) {
#ifdef __CUDA_ARCH__
    d = d + 0.0f;  // -0 -> +0
    uint32_t b = __float_as_uint(d);
#else
    d = d + 0.0f;
    uint32_t b;
    memcpy(&b, &d, 4);
#endif
    if ((b & 0x7fffffffu) > 0x7f800000u) b = 0x7fc00000u;  // every NaN -> one greatest value
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__host__ __device__ __forceinline__ float f32_from_order_bits(uint32_t o) {
    uint32_t b = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
#ifdef __CUDA_ARCH__
    return __uint_as_float(b);
#else
    float d;
    memcpy(&d, &b, 4);
    return d;
#endif
}
__host__ __device__ __forceinline__ uint64_t make_key(float d, uint32_t id) {
    return ((uint64_t)f32_order_bits(d) << 32) | id;
}
__host__ __device__ __forceinline__ float key_dist(uint64_t k) {
    return f32_from_order_bits((uint32_t)(k >> 32));
}
__host__ __device__ __forceinline__ uint32_t key_id(uint64_t k) { return (uint32_t)k; }

#ifdef __CUDACC__
// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
// streaming 128-bit load: read-only path, do not allocate in L1 (rows are touched once)
__device__ __forceinline__ uint4 ldg_stream_u4(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ float4 ldg_stream_f4(const void* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}

// the same with an L2 evict-first policy: gathers that run next to the tensor-core contraction (rerank of one row part
// under the contraction of the next) must not push the contraction's query / row tiles out of L2
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ float4 ldg_stream_f4_ef(const void* p, uint64_t pol) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p), "l"(pol));
    return r;
}

// Reduce V per-lane partials across the warp so that lane L ends up holding the warp-wide total
// of value index (L >> (5 - log2 V)). V-1 + (5 - log2 V) shuffles instead of 5*V.
template <int V, typename T>
__device__ __forceinline__ T warp_reduce_scatter(T (&v)[V], int lane) {
    static_assert(V >= 1 && V <= 32 && (V & (V - 1)) == 0, "V must be a power of two <= 32");
    int offset = 16;
#pragma unroll
    for (int n = V; n > 1; n >>= 1) {
        const int half = n >> 1;
        const bool upper = (lane & offset) != 0;
#pragma unroll
        for (int j = 0; j < half; ++j) {
            T send = upper ? v[j] : v[j + half];
            T keep = upper ? v[j + half] : v[j];
            v[j] = keep + __shfl_xor_sync(0xffffffffu, send, offset);
        }
        offset >>= 1;
    }
    T r = v[0];
#pragma unroll
    for (; offset >= 1; offset >>= 1) r += __shfl_xor_sync(0xffffffffu, r, offset);
    return r;
}
template <int V> struct Log2 { static constexpr int value = 1 + Log2<V / 2>::value; };
template <> struct Log2<1> { static constexpr int value = 0; };

// In-place ascending bitonic sort of `n` (power of two) u64 keys in shared memory by the whole CTA.
// `nseg` independent segments of n keys each (segment s at base + s*stride) are sorted together so
// the barrier count does not grow with the number of segments.
__device__ __forceinline__ void cta_bitonic_sort(uint64_t* base, uint32_t n, uint32_t nseg,
                                                 uint32_t stride) {
    const uint32_t half = n >> 1;
    const uint32_t hshift = 31 - __clz(half);  // n is a power of two: divisions become shifts
    const uint32_t total = half * nseg;
    for (uint32_t size = 2; size <= n; size <<= 1) {
        for (uint32_t j = size >> 1; j > 0; j >>= 1) {
            for (uint32_t t = threadIdx.x; t < total; t += blockDim.x) {
                const uint32_t seg = t >> hshift, p = t & (half - 1);
                const uint32_t i = ((p & ~(j - 1)) << 1) | (p & (j - 1));  // insert a 0 bit at position log2(j)
                const uint32_t l = i | j;
                uint64_t* a = base + (size_t)seg * stride;
                const uint64_t x = a[i], y = a[l];
                const bool up = (i & size) == 0;
                if ((x > y) == up) {
                    a[i] = y;
                    a[l] = x;
                }
            }
            __syncthreads();
        }
    }
}
#endif  // __CUDACC__

}  // namespace vdb

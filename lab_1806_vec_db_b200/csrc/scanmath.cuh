// scanmath.cuh — the per-lane arithmetic of every "scan order" distance kernel (K1 flat_scan, K9 ivf scans, the
// K2b rerank in pairs.cu): packed FP32 pairs for f32 rows, exact integer byte arithmetic for u8 rows (further down).
//
// FFMA2/FADD2 (PTX fma/sub.rn.f32x2 on a 64-bit register pair) do two lanes' worth of FP32 work per issue slot.
// Measured on B200 (scripts/ubench/fp32_rate.cu): scalar FFMA/FADD reach 1 warp instruction per clock and scheduler
// = 128 lane-ops/clk/SM, the packed forms 0.5 per clock = the same 128 lane-ops/clk/SM. So they do not raise the FP32
// ceiling, they free issue slots: the HBM-bound variants of the scan (1-2 queries per row byte, k = 100, cosine)
// gained 4-9 % with them, the FP32-bound 8-query variant did not (and pays for the second accumulator chain).
// NB ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even with --fmad=false, unlike the scalar forms, so the
// bit-exact kernels (k-means assignment, PQ encode: rustc's unfused arithmetic) must not use them.
//
// Summation order (shared by all of these kernels, which is why the tensor-core path's rerank and the IVF probe
// scans return the streaming scan's distance BITS): a lane owns the 4-element chunks c = it*32 + lane of a row;
// inside a chunk the elements 0 and 2 go to the EVEN chain, 1 and 3 to the ODD chain, each chain an fma sequence in chunk order; the lane's total is
// even + odd; lanes are combined by the xor butterfly (warp_reduce_scatter pairs partners the same way).
// Each half of a packed instruction is an IEEE fma/add, so a scalar kernel that follows the same two chains
// produces the same bits (pairs.cu does for unaligned operands).
#pragma once
#include "common.cuh"

namespace vdb {

typedef unsigned long long f32x2;  // two f32 in one 64-bit register pair: {lo = even element, hi = odd element}

__device__ __forceinline__ f32x2 pk2(uint32_t lo, uint32_t hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
    return r;
}
__device__ __forceinline__ f32x2 pk2f(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ float sum2(f32x2 v) {
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
    return lo + hi;
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

// One 16-byte load of an f32 row -> its (x0,x1), (x2,x3) pairs
__device__ __forceinline__ void row_pairs(const uint4& u, f32x2& x01, f32x2& x23) {
    x01 = pk2(u.x, u.y);
    x23 = pk2(u.z, u.w);
}

// a += the chunk's contribution: L2Sqr -> (x - q)^2, otherwise x * q
template <bool L2>
__device__ __forceinline__ f32x2 chunk_acc(f32x2 a, f32x2 x01, f32x2 x23, f32x2 q01, f32x2 q23) {
    if constexpr (L2) {
        const f32x2 d01 = sub2(x01, q01), d23 = sub2(x23, q23);
        a = fma2(d01, d01, a);
        a = fma2(d23, d23, a);
    } else {
        a = fma2(x01, q01, a);
        a = fma2(x23, q23, a);
    }
    return a;
}

// ---- u8 rows: exact integer arithmetic ------------------------------------------------------------------------------
// Rows and queries of a u8 set are bytes, so (x - q)^2, x * q and x * x are small integers: four of them per
// instruction with VABSDIFF4 + IDP.4A into a u32 accumulator (960 x 255^2 < 2^26; the entry points require
// dim <= 65536 so the sum cannot wrap). Integer sums are order-independent, i.e. the scan, the IVF scans and the rerank
// agree bit for bit by construction, and the total is converted to f32 once (exact below 2^24, which is also where the
// reference's sequential f32 sum of the same integers is exact).
__device__ __forceinline__ uint32_t u8x4_l2(uint32_t x, uint32_t q, uint32_t acc) {
    const uint32_t d = __vabsdiffu4(x, q);
    return __dp4a(d, d, acc);
}
__device__ __forceinline__ uint32_t u8x4_dot(uint32_t x, uint32_t q, uint32_t acc) { return __dp4a(x, q, acc); }
template <bool L2>
__device__ __forceinline__ uint32_t u8x16_acc(uint32_t acc, const uint4& x, const uint4& q) {
    if constexpr (L2) {
        acc = u8x4_l2(x.x, q.x, acc);
        acc = u8x4_l2(x.y, q.y, acc);
        acc = u8x4_l2(x.z, q.z, acc);
        acc = u8x4_l2(x.w, q.w, acc);
    } else {
        acc = u8x4_dot(x.x, q.x, acc);
        acc = u8x4_dot(x.y, q.y, acc);
        acc = u8x4_dot(x.z, q.z, acc);
        acc = u8x4_dot(x.w, q.w, acc);
    }
    return acc;
}

// the two chains of chunk_acc with scalar instructions on a float2 (x = even chain, y = odd chain): bit-identical
// (measured in the 8-query scan: same speed as the packed forms)
__device__ __forceinline__ void unpk2(f32x2 v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
template <bool L2>
__device__ __forceinline__ float2 chunk_acc_s(float2 a, const float4& x, const float4& q) {
    if constexpr (L2) {
        const float d0 = x.x - q.x, d1 = x.y - q.y, d2 = x.z - q.z, d3 = x.w - q.w;
        a.x = fmaf(d0, d0, a.x);
        a.y = fmaf(d1, d1, a.y);
        a.x = fmaf(d2, d2, a.x);
        a.y = fmaf(d3, d3, a.y);
    } else {
        a.x = fmaf(x.x, q.x, a.x);
        a.y = fmaf(x.y, q.y, a.y);
        a.x = fmaf(x.z, q.z, a.x);
        a.y = fmaf(x.w, q.w, a.y);
    }
    return a;
}

// the same two chains element by element (element parity picks the chain); bit-identical to chunk_acc
struct ScalarChains {
    float even = 0.f, odd = 0.f;
    __device__ __forceinline__ void l2(uint32_t e, float x, float q) {
        const float d = x - q;
        if (e & 1) odd = fmaf(d, d, odd);
        else even = fmaf(d, d, even);
    }
    __device__ __forceinline__ void dot(uint32_t e, float x, float q) {
        if (e & 1) odd = fmaf(x, q, odd);
        else even = fmaf(x, q, even);
    }
    __device__ __forceinline__ float total() const { return even + odd; }
};

}  // namespace vdb

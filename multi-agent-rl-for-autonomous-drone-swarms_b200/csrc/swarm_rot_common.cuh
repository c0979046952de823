// swarm_rot_common.cuh -- helpers shared by the rotation-pass kernels (swarm_step_rot.cu: N <= 32, one drone per
// lane; swarm_step_rotx.cu: N = 64 / 128, several drones per lane): mbarrier / TMA bulk-copy wrappers, packed-key
// LOP3s, min/max merge and sorting networks.
#pragma once
#include "swarm_device.cuh"

namespace swarm {
namespace {


constexpr int kD = 37;            // 9 + 4 * 3 + 4 * 4
constexpr int kTileBytes = 32 * kD * 4;

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
// global -> shared bulk copy (TMA engine), bytes % 16 == 0, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, unsigned src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// 1: the group inputs arrive by TMA bulk copies (one lane issues them, mbarrier completion);
// 0: by per-lane 16-byte cp.async (LDGSTS), completion by cp.async.wait_group
#ifndef SWARM_ROT_TMA_LOADS
#define SWARM_ROT_TMA_LOADS 1
#endif
#if !SWARM_ROT_TMA_LOADS
__device__ __forceinline__ void cp_async16(unsigned saddr, const void* g) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(g) : "memory");
}
#endif
__device__ __forceinline__ void cp_async4(unsigned saddr, const void* g) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(saddr), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

#ifndef SWARM_ROT_OBST_SQKEY
#define SWARM_ROT_OBST_SQKEY 1
#endif
// (a & mask) | c   and   (a & ~mask) | (b & mask)  as single LOP3s
template <unsigned MASK>
__device__ __forceinline__ unsigned and_or(unsigned a, unsigned c) {
    unsigned r;
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(r) : "r"(a), "n"(MASK), "r"(c));  // (a & b) | c
    return r;
}
template <unsigned MASK>
__device__ __forceinline__ unsigned merge_low(unsigned a, unsigned b) {
    unsigned r;
    asm("lop3.b32 %0, %1, %2, %3, 0xD8;" : "=r"(r) : "r"(a), "r"(b), "n"(MASK));  // (a & ~c) | (b & c)
    return r;
}

// float32 -> float64 without the quarter-rate conversion unit (XU pipe: F2F, MUFU -- the busiest pipe of the pair
// rounds): the float64 bit pattern of a normal a >= 0 is  bits(a) * 2^29 + 0x3800000000000000  (mantissa and
// exponent shifted into place, exponent re-biased), ONE wide multiply-add on the FMA pipe (IMAD.WIDE.U32).
// Zero / denormal inputs come out as some value below 2^-125, which cannot change the float64 sum of squares
// once that sum is >= 2^-28 (the rotation pass checks exactly that and falls back to the exact path otherwise).
// Which conversions take this route is a tuning knob (measured: N = 128, six conversions per pair -- all of them;
// N <= 32 -- none: there the wide multiply's own pipe becomes the bottleneck, 0.1487 vs 0.1457 ms per C4 step):
//   SWARM_ROT_CVT_SQ   bit k: the k-th square of a pair distance (x, y, z)
//   SWARM_ROT_CVT_FORM bit 0 / 1: the forward / backward distance of the formation sum
#ifndef SWARM_ROT_INT_CVT
#define SWARM_ROT_INT_CVT 0
#endif
#ifndef SWARM_ROT_CVT_SQ
#define SWARM_ROT_CVT_SQ (SWARM_ROT_INT_CVT ? 7 : 0)
#endif
#ifndef SWARM_ROT_CVT_FORM
#define SWARM_ROT_CVT_FORM (SWARM_ROT_INT_CVT ? 3 : 0)
#endif
template <bool INT>
__device__ __forceinline__ double f64_of_pos_f32(float a) {
    if (INT) {
        const unsigned b = __float_as_uint(a);
        return __longlong_as_double((long long)((unsigned long long)b * 0x20000000ull + 0x3800000000000000ull));
    }
    return (double)a;
}
// the same for a NEGATIVE normal float32 (the negated distances of the packed rounds): the sign bit of the float32
// pattern lands on bit 60 of the product, so the constant takes 2^60 back out and sets bit 63
__device__ __forceinline__ double f64_of_neg_f32(float a) {
    if (SWARM_ROT_CVT_FORM & 1) {
        const unsigned b = __float_as_uint(a);
        return __longlong_as_double((long long)((unsigned long long)b * 0x20000000ull + 0xA800000000000000ull));
    }
    return (double)a;
}
__device__ __forceinline__ float sumsq1d_fast(float x, float y, float z) {
    const float px = __fmul_rn(x, x), py = __fmul_rn(y, y), pz = __fmul_rn(z, z);
    return __double2float_rn(__dadd_rn(__dadd_rn(f64_of_pos_f32<(SWARM_ROT_CVT_SQ & 1) != 0>(px),
                                                 f64_of_pos_f32<(SWARM_ROT_CVT_SQ & 2) != 0>(py)),
                                       f64_of_pos_f32<(SWARM_ROT_CVT_SQ & 4) != 0>(pz)));
}

// ---- packed float32 pairs (Blackwell FADD2 / FMUL2 / FFMA2: two IEEE round-to-nearest operations per issue slot)
#ifndef SWARM_ROT_PACKED
#define SWARM_ROT_PACKED 1
#endif
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
// squares (dx^2, dy^2, dz^2) of q - p from two packed subtractions / multiplications; .w rides along unused
__device__ __forceinline__ void sq_diff3(const float4& q, f32x2 p_xy, f32x2 p_z0, float& sx, float& sy, float& sz) {
    const f32x2 dxy = sub2(pack2(q.x, q.y), p_xy), dzw = sub2(pack2(q.z, q.w), p_z0);
    float junk;
    unpack2(mul2(dxy, dxy), sx, sy);
    unpack2(mul2(dzw, dzw), sz, junk);
}
// MINUS the sum of squares of np.linalg.norm(vec3) (BLAS sdot: float64 accumulate, cast back), squares given
__device__ __forceinline__ float neg_sumsq1d_of_squares(float px, float py, float pz) {
    return __double2float_rn(-__dadd_rn(__dadd_rn(f64_of_pos_f32<(SWARM_ROT_CVT_SQ & 1) != 0>(px),
                                                  f64_of_pos_f32<(SWARM_ROT_CVT_SQ & 2) != 0>(py)),
                                        f64_of_pos_f32<(SWARM_ROT_CVT_SQ & 4) != 0>(pz)));
}
// sqrt_rn_fast (swarm_device.cuh) of two values at once, on NEGATED inputs and with NEGATED results: with
// ns = -s, y = rsqrt(s):  g' = ns y = -g,  r' = g' g' + ns = -(s - g g),  d' = r' h + g' = -(r h + g) -- every
// operation is the mirror image of the scalar sequence (round-to-nearest is symmetric), so -d' has the same bits,
// and no operand needs a sign flip (the packed instructions have no negate modifier in PTX)
__device__ __forceinline__ f32x2 neg_sqrt_rn_fast2(float ns0, float ns1) {
    float y0, y1;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y0) : "f"(-ns0));
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y1) : "f"(-ns1));
    const f32x2 ns = pack2(ns0, ns1), y = pack2(y0, y1);
    const f32x2 g = mul2(ns, y);
    const f32x2 h = mul2(y, pack2(0.5f, 0.5f));
    const f32x2 r = fma2(g, g, ns);
    return fma2(r, h, g);
}

// atomicAdd(counter, 1) by ONE lane, without the compiler's warp-aggregation wrapper (vote / popc / elect /
// broadcast shuffle): that wrapper consumes the result at once, so the warp sits out the full round trip of an
// atomic on a contended address (12 % of the step kernel's stall samples).  `zero` must be a value that IS zero
// but that the compiler cannot prove uniform (threadIdx.y of a one-dimensional CTA): it keeps ptxas from
// recognising a uniform address.  The result is only waited for where it is first used.
__device__ __forceinline__ unsigned atom_inc_lane(unsigned* counter, unsigned zero) {
    unsigned r;
    asm volatile("atom.global.add.u32 %0, [%1], 1;" : "=r"(r) : "l"(counter + zero) : "memory");
    return r;
}
__device__ __forceinline__ unsigned atom_add_lane(unsigned* counter, unsigned n, unsigned zero) {
    unsigned r;
    asm volatile("atom.global.add.u32 %0, [%1], %2;" : "=r"(r) : "l"(counter + zero), "r"(n) : "memory");
    return r;
}

__device__ __forceinline__ unsigned umin3(unsigned a, unsigned b, unsigned c) { return min(min(a, b), c); }

// sorted k0 <= k1 <= k2 <= k3 (4 smallest keys so far)  <-  two more keys
__device__ __forceinline__ void merge2(unsigned x, unsigned y, unsigned& k0, unsigned& k1, unsigned& k2, unsigned& k3) {
    const unsigned lo = min(x, y), hi = max(x, y);
    const unsigned n0 = min(k0, lo);
    const unsigned n1 = umin3(max(k0, lo), k1, hi);
    const unsigned n2 = umin3(max(k0, hi), max(k1, lo), k2);
    const unsigned n3 = umin3(max(k1, hi), max(k2, lo), k3);
    k0 = n0; k1 = n1; k2 = n2; k3 = n3;
}
__device__ __forceinline__ void merge1(unsigned x, unsigned& k0, unsigned& k1, unsigned& k2, unsigned& k3) {
    const unsigned n3 = min(k3, max(k2, x));
    const unsigned n2 = min(k2, max(k1, x));
    const unsigned n1 = min(k1, max(k0, x));
    k0 = min(k0, x); k1 = n1; k2 = n2; k3 = n3;
}
__device__ __forceinline__ void merge1_5(unsigned x, unsigned& k0, unsigned& k1, unsigned& k2, unsigned& k3, unsigned& k4) {
    const unsigned n4 = min(k4, max(k3, x));
    const unsigned n3 = min(k3, max(k2, x));
    const unsigned n2 = min(k2, max(k1, x));
    const unsigned n1 = min(k1, max(k0, x));
    k0 = min(k0, x); k1 = n1; k2 = n2; k3 = n3; k4 = n4;
}

__device__ __forceinline__ void cex(unsigned& a, unsigned& b) {
    const unsigned lo = min(a, b), hi = max(a, b);
    a = lo; b = hi;
}
// Batcher odd-even merge sort (19 / 5 compare-exchanges); unused outputs are dead-code eliminated
__device__ __forceinline__ void sort8(unsigned (&k)[8]) {
    cex(k[0], k[1]); cex(k[2], k[3]); cex(k[4], k[5]); cex(k[6], k[7]);
    cex(k[0], k[2]); cex(k[1], k[3]); cex(k[4], k[6]); cex(k[5], k[7]);
    cex(k[1], k[2]); cex(k[5], k[6]);
    cex(k[0], k[4]); cex(k[1], k[5]); cex(k[2], k[6]); cex(k[3], k[7]);
    cex(k[2], k[4]); cex(k[3], k[5]);
    cex(k[1], k[2]); cex(k[3], k[4]); cex(k[5], k[6]);
}
__device__ __forceinline__ void sort4(unsigned (&k)[4]) {
    cex(k[0], k[1]); cex(k[2], k[3]);
    cex(k[0], k[2]); cex(k[1], k[3]);
    cex(k[1], k[2]);
}


}  // namespace
}  // namespace swarm

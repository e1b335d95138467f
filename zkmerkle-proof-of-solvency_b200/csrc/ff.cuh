// BN254 prime-field arithmetic for sm_100a: 254-bit elements as 8 x 32-bit limbs in registers, Montgomery form
// (R = 2^256) -- bit-compatible with gnark-crypto's fp.Element / fr.Element memory layout (4 x u64 LE limbs), which
// is what crosses the C-ABI (include/zkpor_b200.h).  Replaces the out-of-tree gnark-crypto ecc/bn254/{fp,fr}
// arithmetic reached from src/prover/prover/prover.go:269 (reference pins it at go.mod:57-60).
//
// Device path: carry-chain PTX (mad.lo.cc / madc.hi.cc pairs, which ptxas fuses into IMAD.WIDE.U32.X), modulus
// limbs as immediates, interleaved CIOS on two register-pair-aligned accumulators (see Fe::mul).
// Host path: the same code against an emulated carry flag; used for one-off constants and the O(1) tail of an MSM.
#pragma once
#include <cstdint>
#include <cstring>

#ifdef __CUDACC__
#define FF_HD __host__ __device__ __forceinline__
#define FF_D __device__ __forceinline__
#else
#define FF_HD inline
#define FF_D inline
#endif

namespace ff {

// ------------------------------------------------------------------------------------------------ parameters
struct FpParams {   // base field q
    FF_HD static constexpr uint32_t M(int i) { constexpr uint32_t t[8] = {0xd87cfd47u, 0x3c208c16u, 0x6871ca8du, 0x97816a91u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u}; return t[i]; }
    static constexpr uint32_t INV = 0xe4866389u;   // -q^-1 mod 2^32
    FF_HD static constexpr uint32_t R2(int i) { constexpr uint32_t t[8] = {0x538afa89u, 0xf32cfc5bu, 0xd44501fbu, 0xb5e71911u, 0x0a417ff6u, 0x47ab1effu, 0xcab8351fu, 0x06d89f71u}; return t[i]; }
    FF_HD static constexpr uint32_t ONE(int i) { constexpr uint32_t t[8] = {0xc58f0d9du, 0xd35d438du, 0xf5c70b3du, 0x0a78eb28u, 0x7879462cu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u}; return t[i]; }
    // R^3 mod q, the modulus in 9 x 30-bit limbs and q^-1 mod 2^30: constants of the divstep inversion (Fe::inv)
    FF_HD static constexpr uint32_t R3(int i) { constexpr uint32_t t[8] = {0xda1530dfu, 0xb1cd6dafu, 0xa7283db6u, 0x62f210e6u, 0x0ada0afbu, 0xef7f0b0cu, 0x2d592544u, 0x20fd6e90u}; return t[i]; }
    FF_HD static constexpr int32_t M30(int i) { constexpr int32_t t[9] = {0x187cfd47, 0x3082305b, 0x71ca8d3, 0x205aa45a, 0x1585d97, 0x116da06, 0x1a029b85, 0x139cb84c, 0x3064}; return t[i]; }
    static constexpr uint32_t INV30 = 0x1b799c77u;
};
struct FrParams {   // scalar field r
    FF_HD static constexpr uint32_t M(int i) { constexpr uint32_t t[8] = {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u}; return t[i]; }
    static constexpr uint32_t INV = 0xefffffffu;
    FF_HD static constexpr uint32_t R2(int i) { constexpr uint32_t t[8] = {0xae216da7u, 0x1bb8e645u, 0xe35c59e3u, 0x53fe3ab1u, 0x53bb8085u, 0x8c49833du, 0x7f4e44a5u, 0x0216d0b1u}; return t[i]; }
    FF_HD static constexpr uint32_t ONE(int i) { constexpr uint32_t t[8] = {0x4ffffffbu, 0xac96341cu, 0x9f60cd29u, 0x36fc7695u, 0x7879462eu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u}; return t[i]; }
    FF_HD static constexpr uint32_t R3(int i) { constexpr uint32_t t[8] = {0xb4bf0040u, 0x5e94d8e1u, 0x1cfbb6b8u, 0x2a489cbeu, 0xa19fcfedu, 0x893cc664u, 0x7fcc657cu, 0x0cf8594bu}; return t[i]; }
    FF_HD static constexpr int32_t M30(int i) { constexpr int32_t t[9] = {0x30000001, 0xf87d64f, 0x1b970914, 0xcfa121e, 0x1585d28, 0x116da06, 0x1a029b85, 0x139cb84c, 0x3064}; return t[i]; }
    static constexpr uint32_t INV30 = 0x10000001u;
};

// ------------------------------------------------------------------------------------------------ PTX carry chains
// On the device these are single PTX instructions sharing the hardware carry flag.  On the host the same call
// sequence runs against an emulated flag, so the exact limb schedule of mul()/add()/sub() is unit-tested on the CPU
// (tests/test_host_ff.py) before it ever reaches a GPU.
namespace ptx {
#ifdef __CUDA_ARCH__
FF_D uint32_t add_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
FF_D uint32_t addc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
FF_D uint32_t addc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
FF_D uint32_t sub_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
FF_D uint32_t subc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
FF_D uint32_t subc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
FF_D uint32_t mul_lo(uint32_t a, uint32_t b) { return a * b; }
FF_D uint32_t mul_hi(uint32_t a, uint32_t b) { return __umulhi(a, b); }
FF_D uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
FF_D uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
FF_D uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
#else
inline uint32_t &cc() { static thread_local uint32_t f = 0; return f; }
inline uint32_t add3(uint32_t a, uint32_t b, uint32_t cin, bool set) { uint64_t t = (uint64_t)a + b + cin; if (set) cc() = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t sub3(uint32_t a, uint32_t b, uint32_t bin, bool set) { uint64_t t = (uint64_t)a - b - bin; if (set) cc() = (uint32_t)((t >> 32) & 1); return (uint32_t)t; }
inline uint32_t add_cc(uint32_t a, uint32_t b) { return add3(a, b, 0, true); }
inline uint32_t addc_cc(uint32_t a, uint32_t b) { return add3(a, b, cc(), true); }
inline uint32_t addc(uint32_t a, uint32_t b) { return add3(a, b, cc(), false); }
inline uint32_t sub_cc(uint32_t a, uint32_t b) { return sub3(a, b, 0, true); }
inline uint32_t subc_cc(uint32_t a, uint32_t b) { return sub3(a, b, cc(), true); }
inline uint32_t subc(uint32_t a, uint32_t b) { return sub3(a, b, cc(), false); }
inline uint32_t mul_lo(uint32_t a, uint32_t b) { return a * b; }
inline uint32_t mul_hi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
inline uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return add3(a * b, c, 0, true); }
inline uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return add3(a * b, c, cc(), true); }
inline uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { return add3(mul_hi(a, b), c, cc(), true); }
#endif
}  // namespace ptx

// ------------------------------------------------------------------------------------------------ the element type
template <class P>
struct alignas(16) Fe {
    uint32_t l[8];

    FF_HD static Fe zero() { Fe r; for (int i = 0; i < 8; i++) r.l[i] = 0; return r; }
    FF_HD static Fe one() { Fe r; for (int i = 0; i < 8; i++) r.l[i] = P::ONE(i); return r; }
    FF_HD static Fe r2() { Fe r; for (int i = 0; i < 8; i++) r.l[i] = P::R2(i); return r; }
    FF_HD static Fe modulus() { Fe r; for (int i = 0; i < 8; i++) r.l[i] = P::M(i); return r; }
    FF_HD bool is_zero() const { uint32_t o = 0; for (int i = 0; i < 8; i++) o |= l[i]; return o == 0; }
    FF_HD bool operator==(const Fe &b) const { uint32_t o = 0; for (int i = 0; i < 8; i++) o |= l[i] ^ b.l[i]; return o == 0; }
    FF_HD bool operator!=(const Fe &b) const { return !(*this == b); }

    // x in [0, 2m) -> [0, m)
    FF_HD static Fe reduce_once(const Fe &x) {
        Fe d;
        d.l[0] = ptx::sub_cc(x.l[0], P::M(0));
#pragma unroll
        for (int i = 1; i < 8; i++) d.l[i] = ptx::subc_cc(x.l[i], P::M(i));
        uint32_t borrow = ptx::subc(0u, 0u);   // 0 or 0xffffffff
#pragma unroll
        for (int i = 0; i < 8; i++) d.l[i] = borrow ? x.l[i] : d.l[i];
        return d;
    }
    // a + b mod m (inputs < m < 2^254, so the sum never carries out of limb 7)
    FF_HD static Fe add(const Fe &a, const Fe &b) {
        Fe r;
        r.l[0] = ptx::add_cc(a.l[0], b.l[0]);
#pragma unroll
        for (int i = 1; i < 7; i++) r.l[i] = ptx::addc_cc(a.l[i], b.l[i]);
        r.l[7] = ptx::addc(a.l[7], b.l[7]);
        return reduce_once(r);
    }
    // a - b mod m
    FF_HD static Fe sub(const Fe &a, const Fe &b) {
        Fe r;
        r.l[0] = ptx::sub_cc(a.l[0], b.l[0]);
#pragma unroll
        for (int i = 1; i < 8; i++) r.l[i] = ptx::subc_cc(a.l[i], b.l[i]);
        uint32_t borrow = ptx::subc(0u, 0u);
        r.l[0] = ptx::add_cc(r.l[0], P::M(0) & borrow);
#pragma unroll
        for (int i = 1; i < 7; i++) r.l[i] = ptx::addc_cc(r.l[i], P::M(i) & borrow);
        r.l[7] = ptx::addc(r.l[7], P::M(7) & borrow);
        return r;
    }
    FF_HD static Fe neg(const Fe &a) { return a.is_zero() ? a : sub(modulus(), a); }
    FF_HD static Fe dbl(const Fe &a) { return add(a, a); }

    // ---- Karatsuba variant: 3 x (4x4) limb products + a sliding-window REDC = 112 IMAD.WIDE instead of 128 ----------
    // 4x4 limb product, 8 limbs out.  Partial products land on aligned register pairs of two accumulators: Pe (pairs at
    // even columns) for i+j even, Qo (pairs at odd columns; Qo[k] is column k+1) for i+j odd.
    FF_HD static void mul4x4(const uint32_t *x, const uint32_t *y, uint32_t *out) {
        using namespace ptx;
        uint32_t Pe[8], Qo[8];   // Qo[k] = column k+1, k = 0..6
        // row 0: Pe <- x0*y0 (cols 0,1), x2*y0 (cols 2,3); Qo <- x1*y0 (cols 1,2), x3*y0 (cols 3,4)
        Pe[0] = mul_lo(x[0], y[0]); Pe[1] = mul_hi(x[0], y[0]); Pe[2] = mul_lo(x[2], y[0]); Pe[3] = mul_hi(x[2], y[0]);
        Qo[0] = mul_lo(x[1], y[0]); Qo[1] = mul_hi(x[1], y[0]); Qo[2] = mul_lo(x[3], y[0]); Qo[3] = mul_hi(x[3], y[0]);
        // row 1: i+j even -> j = 1,3 -> Pe cols 2..5 ; i+j odd -> j = 0,2 -> Qo cols 1..4
        Pe[2] = mad_lo_cc(x[1], y[1], Pe[2]); Pe[3] = madc_hi_cc(x[1], y[1], Pe[3]);
        Pe[4] = madc_lo_cc(x[3], y[1], 0u); Pe[5] = madc_hi_cc(x[3], y[1], 0u); Pe[6] = addc(0u, 0u);
        Qo[0] = mad_lo_cc(x[0], y[1], Qo[0]); Qo[1] = madc_hi_cc(x[0], y[1], Qo[1]);
        Qo[2] = madc_lo_cc(x[2], y[1], Qo[2]); Qo[3] = madc_hi_cc(x[2], y[1], Qo[3]); Qo[4] = addc(0u, 0u);
        // row 2: even -> j = 0,2 -> Pe cols 2..5 ; odd -> j = 1,3 -> Qo cols 3..6
        Pe[2] = mad_lo_cc(x[0], y[2], Pe[2]); Pe[3] = madc_hi_cc(x[0], y[2], Pe[3]);
        Pe[4] = madc_lo_cc(x[2], y[2], Pe[4]); Pe[5] = madc_hi_cc(x[2], y[2], Pe[5]); Pe[6] = addc(Pe[6], 0u);
        Qo[2] = mad_lo_cc(x[1], y[2], Qo[2]); Qo[3] = madc_hi_cc(x[1], y[2], Qo[3]);
        Qo[4] = madc_lo_cc(x[3], y[2], Qo[4]); Qo[5] = madc_hi_cc(x[3], y[2], 0u); Qo[6] = addc(0u, 0u);
        // row 3: even -> j = 1,3 -> Pe cols 4..7 ; odd -> j = 0,2 -> Qo cols 3..6
        Pe[4] = mad_lo_cc(x[1], y[3], Pe[4]); Pe[5] = madc_hi_cc(x[1], y[3], Pe[5]);
        Pe[6] = madc_lo_cc(x[3], y[3], Pe[6]); Pe[7] = madc_hi_cc(x[3], y[3], 0u);
        Qo[2] = mad_lo_cc(x[0], y[3], Qo[2]); Qo[3] = madc_hi_cc(x[0], y[3], Qo[3]);
        Qo[4] = madc_lo_cc(x[2], y[3], Qo[4]); Qo[5] = madc_hi_cc(x[2], y[3], Qo[5]); Qo[6] = addc(Qo[6], 0u);
        // out = Pe + (Qo << 32)
        out[0] = Pe[0];
        out[1] = add_cc(Pe[1], Qo[0]);
#pragma unroll
        for (int k = 2; k < 7; k++) out[k] = addc_cc(Pe[k], Qo[k - 1]);
        out[7] = addc(Pe[7], Qo[6]);
    }
    // d = |x - y| over 4 limbs, returns 1 when x < y
    FF_HD static uint32_t absdiff4(const uint32_t *x, const uint32_t *y, uint32_t *d) {
        using namespace ptx;
        d[0] = sub_cc(x[0], y[0]); d[1] = subc_cc(x[1], y[1]); d[2] = subc_cc(x[2], y[2]); d[3] = subc_cc(x[3], y[3]);
        uint32_t m = subc(0u, 0u);   // 0 or 0xffffffff
        d[0] = add_cc(d[0] ^ m, m & 1u); d[1] = addc_cc(d[1] ^ m, 0u); d[2] = addc_cc(d[2] ^ m, 0u); d[3] = addc(d[3] ^ m, 0u);
        return m & 1u;
    }
    // REDC round whose O chain takes the pending carry (of the stray-limb fold) and that reports overflow of the top
    // window limb: `over` = carries out of column c+8 (0..2)
    template <bool CARRY_IN>
    FF_HD static uint32_t redc_round_win(uint32_t (&E)[8], uint32_t (&O)[8]) {
        using namespace ptx;
        uint32_t q = E[0] * P::INV;
        O[0] = CARRY_IN ? madc_lo_cc(q, P::M(1), O[0]) : mad_lo_cc(q, P::M(1), O[0]);
        O[1] = madc_hi_cc(q, P::M(1), O[1]);
        O[2] = madc_lo_cc(q, P::M(3), O[2]);
        O[3] = madc_hi_cc(q, P::M(3), O[3]);
        O[4] = madc_lo_cc(q, P::M(5), O[4]);
        O[5] = madc_hi_cc(q, P::M(5), O[5]);
        O[6] = madc_lo_cc(q, P::M(7), O[6]);
        O[7] = madc_hi_cc(q, P::M(7), O[7]);
        uint32_t over = addc(0u, 0u);
        E[0] = mad_lo_cc(q, P::M(0), E[0]);
        E[1] = madc_hi_cc(q, P::M(0), E[1]);
        E[2] = madc_lo_cc(q, P::M(2), E[2]);
        E[3] = madc_hi_cc(q, P::M(2), E[3]);
        E[4] = madc_lo_cc(q, P::M(4), E[4]);
        E[5] = madc_hi_cc(q, P::M(4), E[5]);
        E[6] = madc_lo_cc(q, P::M(6), E[6]);
        E[7] = madc_hi_cc(q, P::M(6), E[7]);
        O[7] = addc_cc(O[7], 0u);
        over = addc(over, 0u);
        return over;
    }
    // Montgomery reduction of a 512-bit T (16 limbs, T < m * 2^256): returns T / 2^256 mod m.  A 9-limb window slides
    // over T; limb T[i+8] enters the window at round i as the fresh top limb of O.
    FF_HD static Fe redc512(const uint32_t *T) {
        using namespace ptx;
        uint32_t E[8], O[8];
#pragma unroll
        for (int k = 0; k < 8; k++) { E[k] = T[k]; O[k] = 0; }
        O[7] = T[8];
        uint32_t pend = redc_round_win<false>(E, O);
#pragma unroll
        for (int i = 1; i < 8; i++) {
            uint32_t stray = E[1], nO[8];
#pragma unroll
            for (int k = 0; k < 6; k++) nO[k] = E[k + 2];
            nO[6] = 0;
            nO[7] = add_cc(T[i + 8], pend);          // column i+8 enters; overflow of this add joins the next pend
            uint32_t pend2 = addc(0u, 0u);
#pragma unroll
            for (int k = 0; k < 8; k++) { E[k] = O[k]; O[k] = nO[k]; }
            E[0] = add_cc(E[0], stray);              // carry -> column c+1 = O[0], consumed by the round's O chain
            pend = redc_round_win<true>(E, O) + pend2;
        }
        Fe r;
        r.l[0] = add_cc(O[0], E[1]);
#pragma unroll
        for (int k = 1; k < 7; k++) r.l[k] = addc_cc(O[k], E[k + 1]);
        r.l[7] = addc(O[7], 0u);
        return reduce_once(r);
    }
    FF_HD static Fe mul_kara(const Fe &a, const Fe &b) {
        using namespace ptx;
        uint32_t z0[8], z2[8], mm[8], da[4], db[4], T[16];
        mul4x4(a.l, b.l, z0);
        mul4x4(a.l + 4, b.l + 4, z2);
        uint32_t sa = absdiff4(a.l, a.l + 4, da);         // |a_lo - a_hi|, sa = (a_lo < a_hi)
        uint32_t sb = absdiff4(b.l + 4, b.l, db);         // |b_hi - b_lo|, sb = (b_hi < b_lo)
        mul4x4(da, db, mm);
        // z1 = z0 + z2 +/- mm   (a_lo*b_hi + a_hi*b_lo), 9 limbs: z1[8] in {0,1,2 - borrow...} kept as signed-safe carry
        uint32_t s[9];
        s[0] = add_cc(z0[0], z2[0]);
#pragma unroll
        for (int k = 1; k < 8; k++) s[k] = addc_cc(z0[k], z2[k]);
        s[8] = addc(0u, 0u);
        uint32_t neg = (sa ^ sb) ? 0xffffffffu : 0u;      // subtract mm when the signs differ
        s[0] = add_cc(s[0], (mm[0] ^ neg)); // two's complement: + (~mm) + 1 when neg
        // the +1 of the two's complement is injected below through a second chain start; do it explicitly:
#pragma unroll
        for (int k = 1; k < 8; k++) s[k] = addc_cc(s[k], mm[k] ^ neg);
        s[8] = addc(s[8], neg);                           // sign extension of (~mm)
        s[0] = add_cc(s[0], neg & 1u);
#pragma unroll
        for (int k = 1; k < 8; k++) s[k] = addc_cc(s[k], 0u);
        s[8] = addc(s[8], 0u);
        // T = z0 + z1 * 2^128 + z2 * 2^256
#pragma unroll
        for (int k = 0; k < 4; k++) T[k] = z0[k];
        T[4] = add_cc(z0[4], s[0]); T[5] = addc_cc(z0[5], s[1]); T[6] = addc_cc(z0[6], s[2]); T[7] = addc_cc(z0[7], s[3]);
        T[8] = addc_cc(z2[0], s[4]); T[9] = addc_cc(z2[1], s[5]); T[10] = addc_cc(z2[2], s[6]); T[11] = addc_cc(z2[3], s[7]);
        T[12] = addc_cc(z2[4], s[8]); T[13] = addc_cc(z2[5], 0u); T[14] = addc_cc(z2[6], 0u); T[15] = addc(z2[7], 0u);
        return redc512(T);
    }

    // textbook CIOS on 64-bit temporaries: the host-side cross-check of mul() (tests/test_host_ff.py)
    static inline Fe mul_ref(const Fe &a, const Fe &b) {
        Fe r;
        uint64_t t[18];
        for (int i = 0; i < 18; i++) t[i] = 0;
        for (int i = 0; i < 8; i++) {
            uint64_t c = 0;
            for (int j = 0; j < 8; j++) { c += (uint64_t)a.l[j] * b.l[i] + t[i + j]; t[i + j] = (uint32_t)c; c >>= 32; }
            t[i + 8] += c;
        }
        for (int i = 0; i < 8; i++) {
            uint32_t q = (uint32_t)t[i] * P::INV;
            uint64_t c = 0;
            for (int j = 0; j < 8; j++) { c += (uint64_t)q * P::M(j) + t[i + j]; t[i + j] = (uint32_t)c; c >>= 32; }
            for (int k = i + 8; c && k < 18; k++) { c += t[k]; t[k] = (uint32_t)c; c >>= 32; }
        }
        for (int i = 0; i < 8; i++) r.l[i] = (uint32_t)t[8 + i];
        return reduce_once(r);
    }
    // One CIOS reduction round on the two accumulators (see mul()): q kills column E[0], q*m is added in place.
    FF_HD static void redc_round(uint32_t (&E)[8], uint32_t (&O)[8]) {
        using namespace ptx;
        uint32_t q = E[0] * P::INV;
        O[0] = mad_lo_cc(q, P::M(1), O[0]);
        O[1] = madc_hi_cc(q, P::M(1), O[1]);
        O[2] = madc_lo_cc(q, P::M(3), O[2]);
        O[3] = madc_hi_cc(q, P::M(3), O[3]);
        O[4] = madc_lo_cc(q, P::M(5), O[4]);
        O[5] = madc_hi_cc(q, P::M(5), O[5]);
        O[6] = madc_lo_cc(q, P::M(7), O[6]);
        O[7] = madc_hi_cc(q, P::M(7), O[7]);
        E[0] = mad_lo_cc(q, P::M(0), E[0]);
        E[1] = madc_hi_cc(q, P::M(0), E[1]);
        E[2] = madc_lo_cc(q, P::M(2), E[2]);
        E[3] = madc_hi_cc(q, P::M(2), E[3]);
        E[4] = madc_lo_cc(q, P::M(4), E[4]);
        E[5] = madc_hi_cc(q, P::M(4), E[5]);
        E[6] = madc_lo_cc(q, P::M(6), E[6]);
        E[7] = madc_hi_cc(q, P::M(6), E[7]);
        O[7] = addc(O[7], 0u);
    }

    // Montgomery product a*b/R mod m -- interleaved CIOS on two 8-limb accumulators.
    //   E holds columns c..c+7, O holds columns c+1..c+8 (c = current row).  A partial product a[j]*b[i] is a 64-bit
    //   value landing on an (even, odd) register pair of E when j is even and of O when j is odd, so every lo/hi pair
    //   is one IMAD.WIDE.U32 with carry on a fixed, aligned register pair and every row is an uninterrupted carry
    //   chain.  After the reduction round E[0] = 0; the accumulators trade places (O becomes the new E, E shifted
    //   down by two limbs becomes the new O -- pure register renaming once unrolled) and the left-over limb E[1] is
    //   folded into the new E[0], its carry entering the new O chain.
    FF_HD static Fe mul(const Fe &a, const Fe &b) {
#if !defined(__CUDA_ARCH__) && !defined(FF_HOST_EMULATE_PTX)
        return mul_ref(a, b);   // host: plain CIOS; the emulated-flag build (hosttest.cpp) exercises the schedule below
#elif defined(FF_KARATSUBA) && defined(__CUDA_ARCH__)
        return mul_kara(a, b);
#else
        using namespace ptx;
        uint32_t E[8], O[8];
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
            E[j] = mul_lo(a.l[j], b.l[0]); E[j + 1] = mul_hi(a.l[j], b.l[0]);
            O[j] = mul_lo(a.l[j + 1], b.l[0]); O[j + 1] = mul_hi(a.l[j + 1], b.l[0]);
        }
        redc_round(E, O);
#pragma unroll
        for (int i = 1; i < 8; i++) {
            uint32_t stray = E[1], nO[8];
#pragma unroll
            for (int k = 0; k < 6; k++) nO[k] = E[k + 2];
            nO[6] = 0; nO[7] = 0;
#pragma unroll
            for (int k = 0; k < 8; k++) { E[k] = O[k]; O[k] = nO[k]; }
            E[0] = add_cc(E[0], stray);                       // carry belongs to column c+1 = O[0]
            O[0] = madc_lo_cc(a.l[1], b.l[i], O[0]);
            O[1] = madc_hi_cc(a.l[1], b.l[i], O[1]);
            O[2] = madc_lo_cc(a.l[3], b.l[i], O[2]);
            O[3] = madc_hi_cc(a.l[3], b.l[i], O[3]);
            O[4] = madc_lo_cc(a.l[5], b.l[i], O[4]);
            O[5] = madc_hi_cc(a.l[5], b.l[i], O[5]);
            O[6] = madc_lo_cc(a.l[7], b.l[i], O[6]);
            O[7] = madc_hi_cc(a.l[7], b.l[i], O[7]);
            E[0] = mad_lo_cc(a.l[0], b.l[i], E[0]);
            E[1] = madc_hi_cc(a.l[0], b.l[i], E[1]);
            E[2] = madc_lo_cc(a.l[2], b.l[i], E[2]);
            E[3] = madc_hi_cc(a.l[2], b.l[i], E[3]);
            E[4] = madc_lo_cc(a.l[4], b.l[i], E[4]);
            E[5] = madc_hi_cc(a.l[4], b.l[i], E[5]);
            E[6] = madc_lo_cc(a.l[6], b.l[i], E[6]);
            E[7] = madc_hi_cc(a.l[6], b.l[i], E[7]);
            O[7] = addc(O[7], 0u);
            redc_round(E, O);
        }
        // result = O + (E >> 32); it is < 2m < 2^255
        Fe r;
        r.l[0] = add_cc(O[0], E[1]);
#pragma unroll
        for (int k = 1; k < 7; k++) r.l[k] = addc_cc(O[k], E[k + 1]);
        r.l[7] = addc(O[7], 0u);
        return reduce_once(r);
#endif
    }
    FF_HD static Fe sqr(const Fe &a) { return mul(a, a); }
    FF_HD static Fe to_mont(const Fe &a) { return mul(a, r2()); }
    FF_HD static Fe from_mont(const Fe &a) { Fe o = zero(); o.l[0] = 1; return mul(a, o); }

    // a^e, e = 8 x u32 little-endian plain integer
    FF_HD static Fe pow(const Fe &a, const uint32_t e[8]) {
        Fe acc = one();
        for (int i = 255; i >= 0; i--) {
            acc = sqr(acc);
            if ((e[i >> 5] >> (i & 31)) & 1) acc = mul(acc, a);
        }
        return acc;
    }
    FF_HD static Fe inv_fermat(const Fe &a) {   // a^(m-2); inv(0) = 0.  Kept as the cross-check of inv()
        uint32_t e[8];
        for (int i = 0; i < 8; i++) e[i] = P::M(i);
        e[0] -= 2;   // both moduli are odd with low limb >= 2
        return pow(a, e);
    }

    // ---- inversion by divsteps (Bernstein-Yang "safegcd", half-delta variant) ------------------------------------
    // 20 rounds of 30 divsteps on the low words of (f, g) = (m, a), each round followed by the 2x2 transition matrix
    // applied to the full-width (f, g) and, modulo m, to (d, e) -- numbers held as 9 signed 30-bit limbs.  600 divsteps
    // cover every input below 2^256; control flow is data-independent, so the 32 lanes of a warp never diverge.
    // Cost: ~1800 32x32->64 multiply-adds + ~10^4 ALU-pipe instructions, i.e. the IMAD-pipe time of ~15 field
    // products (Fermat: ~320 products).  This is what makes batched-affine bucket accumulation pay (msm.cu).
    FF_HD static int32_t divsteps30(int32_t zeta, uint32_t f, uint32_t g, int32_t (&t)[4]) {
        uint32_t u = 1, v = 0, q = 0, r = 1;
#pragma unroll 6
        for (int i = 0; i < 30; i++) {
            uint32_t c1 = (uint32_t)(zeta >> 31), c2 = 0u - (g & 1u);
            uint32_t x = (f ^ c1) - c1, y = (u ^ c1) - c1, z = (v ^ c1) - c1;
            g += x & c2; q += y & c2; r += z & c2;
            c1 &= c2;
            zeta = (zeta ^ (int32_t)c1) - 1;
            f += g & c1; u += q & c1; v += r & c1;
            g >>= 1; u <<= 1; v <<= 1;
        }
        t[0] = (int32_t)u; t[1] = (int32_t)v; t[2] = (int32_t)q; t[3] = (int32_t)r;
        return zeta;
    }
    FF_HD static Fe inv(const Fe &a) {   // inv(0) = 0
        const int32_t M30 = 0x3fffffff;
        int32_t d[9], e[9], f[9], g[9];
        // a (8 x 32 bits) -> 9 x 30 bits
#pragma unroll
        for (int i = 0; i < 9; i++) {
            const int bit = 30 * i, w = bit >> 5, s = bit & 31;
            uint64_t lo = a.l[w], hi = (w + 1 < 8) ? a.l[w + 1] : 0u;
            g[i] = (int32_t)(((lo | (hi << 32)) >> s) & (uint64_t)M30);
            f[i] = P::M30(i); d[i] = 0; e[i] = 0;
        }
        e[0] = 1;
        int32_t zeta = -1;
        for (int it = 0; it < 20; it++) {
            int32_t t[4];
            zeta = divsteps30(zeta, (uint32_t)f[0], (uint32_t)g[0], t);
            const int32_t u = t[0], v = t[1], q = t[2], r = t[3];
            // every product below is int32 x int32 -> int64 (one IMAD.WIDE with the 64-bit accumulator as addend)
#define FF_MW(a, b) ((int64_t)(a) * (int64_t)(b))
            // (d, e) <- t * (d, e) / 2^30 mod m: multiples of m are added so that the low 30 bits cancel
            {
                const int32_t sd = d[8] >> 31, se = e[8] >> 31;
                int32_t md = (u & sd) + (v & se), me = (q & sd) + (r & se);
                int64_t cd = FF_MW(u, d[0]) + FF_MW(v, e[0]), ce = FF_MW(q, d[0]) + FF_MW(r, e[0]);
                md -= (int32_t)((P::INV30 * (uint32_t)cd + (uint32_t)md) & (uint32_t)M30);
                me -= (int32_t)((P::INV30 * (uint32_t)ce + (uint32_t)me) & (uint32_t)M30);
                cd += FF_MW(P::M30(0), md); ce += FF_MW(P::M30(0), me);
                cd >>= 30; ce >>= 30;
#pragma unroll
                for (int i = 1; i < 9; i++) {
                    cd += FF_MW(u, d[i]) + FF_MW(v, e[i]) + FF_MW(P::M30(i), md);
                    ce += FF_MW(q, d[i]) + FF_MW(r, e[i]) + FF_MW(P::M30(i), me);
                    d[i - 1] = (int32_t)cd & M30; cd >>= 30;
                    e[i - 1] = (int32_t)ce & M30; ce >>= 30;
                }
                d[8] = (int32_t)cd; e[8] = (int32_t)ce;
            }
            // (f, g) <- t * (f, g) / 2^30 (exact)
            {
                int64_t cf = FF_MW(u, f[0]) + FF_MW(v, g[0]), cg = FF_MW(q, f[0]) + FF_MW(r, g[0]);
                cf >>= 30; cg >>= 30;
#pragma unroll
                for (int i = 1; i < 9; i++) {
                    cf += FF_MW(u, f[i]) + FF_MW(v, g[i]);
                    cg += FF_MW(q, f[i]) + FF_MW(r, g[i]);
                    f[i - 1] = (int32_t)cf & M30; cf >>= 30;
                    g[i - 1] = (int32_t)cg & M30; cg >>= 30;
                }
                f[8] = (int32_t)cf; g[8] = (int32_t)cg;
            }
#undef FF_MW
        }
        // now g = 0 and f = +/- gcd = +/- 1 (or +/- m when a = 0): a^-1 = sign(f) * d, brought to [0, m)
        {
            int32_t add = d[8] >> 31;
#pragma unroll
            for (int i = 0; i < 9; i++) d[i] += P::M30(i) & add;
            const int32_t ng = f[8] >> 31;
#pragma unroll
            for (int i = 0; i < 9; i++) d[i] = (d[i] ^ ng) - ng;
#pragma unroll
            for (int i = 0; i < 8; i++) { d[i + 1] += d[i] >> 30; d[i] &= M30; }
            add = d[8] >> 31;
#pragma unroll
            for (int i = 0; i < 9; i++) d[i] += P::M30(i) & add;
#pragma unroll
            for (int i = 0; i < 8; i++) { d[i + 1] += d[i] >> 30; d[i] &= M30; }
        }
        // 9 x 30 bits -> 8 x 32 bits; the plain inverse of (a R) is a^-1 R^-1, one product by R^3 restores a^-1 R
        Fe o;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int bit = 32 * i, w = bit / 30, s = bit % 30;
            uint64_t acc = (uint64_t)(uint32_t)d[w] >> s;
            acc |= (uint64_t)(uint32_t)d[w + 1] << (30 - s);
            if (w + 2 < 9 && 60 - s < 32) acc |= (uint64_t)(uint32_t)d[w + 2] << (60 - s);
            o.l[i] = (uint32_t)acc;
        }
        Fe r3;
#pragma unroll
        for (int i = 0; i < 8; i++) r3.l[i] = P::R3(i);
        return mul(o, r3);
    }
    FF_HD static Fe from_u64(uint64_t v) { Fe o = zero(); o.l[0] = (uint32_t)v; o.l[1] = (uint32_t)(v >> 32); return to_mont(o); }
};

using Fp = Fe<FpParams>;
using Fr = Fe<FrParams>;

// ------------------------------------------------------------------------------------------------ Fp2 = Fp[u]/(u^2+1)
struct alignas(16) Fp2 {
    Fp a0, a1;
    FF_HD static Fp2 zero() { return Fp2{Fp::zero(), Fp::zero()}; }
    FF_HD static Fp2 one() { return Fp2{Fp::one(), Fp::zero()}; }
    FF_HD bool is_zero() const { return a0.is_zero() && a1.is_zero(); }
    FF_HD bool operator==(const Fp2 &b) const { return a0 == b.a0 && a1 == b.a1; }
    FF_HD bool operator!=(const Fp2 &b) const { return !(*this == b); }
    FF_HD static Fp2 add(const Fp2 &a, const Fp2 &b) { return Fp2{Fp::add(a.a0, b.a0), Fp::add(a.a1, b.a1)}; }
    FF_HD static Fp2 sub(const Fp2 &a, const Fp2 &b) { return Fp2{Fp::sub(a.a0, b.a0), Fp::sub(a.a1, b.a1)}; }
    FF_HD static Fp2 neg(const Fp2 &a) { return Fp2{Fp::neg(a.a0), Fp::neg(a.a1)}; }
    FF_HD static Fp2 dbl(const Fp2 &a) { return add(a, a); }
    FF_HD static Fp2 mul(const Fp2 &a, const Fp2 &b) {   // Karatsuba, 3 base-field products
        Fp t0 = Fp::mul(a.a0, b.a0), t1 = Fp::mul(a.a1, b.a1);
        Fp m = Fp::mul(Fp::add(a.a0, a.a1), Fp::add(b.a0, b.a1));
        return Fp2{Fp::sub(t0, t1), Fp::sub(Fp::sub(m, t0), t1)};
    }
    FF_HD static Fp2 sqr(const Fp2 &a) {                 // (a0+a1)(a0-a1) + 2 a0 a1 u
        Fp s = Fp::add(a.a0, a.a1), d = Fp::sub(a.a0, a.a1), p = Fp::mul(a.a0, a.a1);
        return Fp2{Fp::mul(s, d), Fp::dbl(p)};
    }
    FF_HD static Fp2 inv(const Fp2 &a) {
        Fp d = Fp::inv(Fp::add(Fp::sqr(a.a0), Fp::sqr(a.a1)));
        return Fp2{Fp::mul(a.a0, d), Fp::neg(Fp::mul(a.a1, d))};
    }
};

}  // namespace ff

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
};
struct FrParams {   // scalar field r
    FF_HD static constexpr uint32_t M(int i) { constexpr uint32_t t[8] = {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u}; return t[i]; }
    static constexpr uint32_t INV = 0xefffffffu;
    FF_HD static constexpr uint32_t R2(int i) { constexpr uint32_t t[8] = {0xae216da7u, 0x1bb8e645u, 0xe35c59e3u, 0x53fe3ab1u, 0x53bb8085u, 0x8c49833du, 0x7f4e44a5u, 0x0216d0b1u}; return t[i]; }
    FF_HD static constexpr uint32_t ONE(int i) { constexpr uint32_t t[8] = {0x4ffffffbu, 0xac96341cu, 0x9f60cd29u, 0x36fc7695u, 0x7879462eu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u}; return t[i]; }
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
    FF_HD static Fe inv(const Fe &a) {   // Fermat; inv(0) = 0
        uint32_t e[8];
        for (int i = 0; i < 8; i++) e[i] = P::M(i);
        e[0] -= 2;   // both moduli are odd with low limb >= 2
        return pow(a, e);
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

// BN254 optimal-ate pairing: tower Fp2 -> Fp6 -> Fp12, Miller loop with affine line functions, final exponentiation.
// Replaces gnark-crypto bn254.MillerLoop / FinalExponentiation / PairingCheck (ecc/bn254/pairing.go, out of tree) as used
// by groth16.Verify -- src/prover/prover/prover.go:276, src/verifier/main.go:284 -- and by pedersen.VerifyingKey.Verify.
//   Fp2 = Fp[u]/(u^2+1), Fp6 = Fp2[v]/(v^3 - xi), xi = 9+u, Fp12 = Fp6[w]/(w^2 - v); memory layout of Fp12 = gnark-crypto
//   E12{C0, C1 E6{B0, B1, B2 E2{A0, A1}}}, Montgomery limbs.
// The Miller loop keeps the running point T affine: with the divstep inversion of ff.cuh a slope costs about as much as
// the projective formulas' extra products, every line has the fixed sparse shape (-yP) + (lambda xP) w + (yT - lambda xT) w^3
// and the loop value equals the textbook one bit for bit (the oracle, oracle/py/pairing.py, computes exactly this).
// The final exponentiation uses the exact exponent (q^12-1)/r = (q^6-1) * (q^6+1)/r: conjugate/inverse, then one
// square-and-multiply over the 1268-bit cofactor -- O(1) host work per pairing PRODUCT (one per Verify, one per batch).
#pragma once
#include "ec.cuh"
#include "pairing_consts.h"

namespace pairing {
using ff::Fp;
using ff::Fp2;

FF_HD Fp2 fp2_mul_xi(const Fp2 &a) {   // (9 + u) a
    Fp2 t = Fp2::dbl(Fp2::dbl(Fp2::dbl(a)));
    t = Fp2::add(t, a);
    return Fp2{Fp::sub(t.a0, a.a1), Fp::add(t.a1, a.a0)};
}
FF_HD Fp2 fp2_mul_fp(const Fp2 &a, const Fp &k) { return Fp2{Fp::mul(a.a0, k), Fp::mul(a.a1, k)}; }
FF_HD Fp2 fp2_conj(const Fp2 &a) { return Fp2{a.a0, Fp::neg(a.a1)}; }

struct alignas(16) Fp6 {
    Fp2 c0, c1, c2;
    FF_HD static Fp6 zero() { return Fp6{Fp2::zero(), Fp2::zero(), Fp2::zero()}; }
    FF_HD static Fp6 one() { return Fp6{Fp2::one(), Fp2::zero(), Fp2::zero()}; }
    FF_HD bool operator==(const Fp6 &b) const { return c0 == b.c0 && c1 == b.c1 && c2 == b.c2; }
    FF_HD static Fp6 add(const Fp6 &a, const Fp6 &b) { return Fp6{Fp2::add(a.c0, b.c0), Fp2::add(a.c1, b.c1), Fp2::add(a.c2, b.c2)}; }
    FF_HD static Fp6 sub(const Fp6 &a, const Fp6 &b) { return Fp6{Fp2::sub(a.c0, b.c0), Fp2::sub(a.c1, b.c1), Fp2::sub(a.c2, b.c2)}; }
    FF_HD static Fp6 neg(const Fp6 &a) { return Fp6{Fp2::neg(a.c0), Fp2::neg(a.c1), Fp2::neg(a.c2)}; }
    FF_HD static Fp6 mul(const Fp6 &a, const Fp6 &b) {   // 6 Fp2 products
        Fp2 t0 = Fp2::mul(a.c0, b.c0), t1 = Fp2::mul(a.c1, b.c1), t2 = Fp2::mul(a.c2, b.c2);
        Fp2 m12 = Fp2::mul(Fp2::add(a.c1, a.c2), Fp2::add(b.c1, b.c2));
        Fp2 m01 = Fp2::mul(Fp2::add(a.c0, a.c1), Fp2::add(b.c0, b.c1));
        Fp2 m02 = Fp2::mul(Fp2::add(a.c0, a.c2), Fp2::add(b.c0, b.c2));
        Fp6 r;
        r.c0 = Fp2::add(t0, fp2_mul_xi(Fp2::sub(Fp2::sub(m12, t1), t2)));
        r.c1 = Fp2::add(Fp2::sub(Fp2::sub(m01, t0), t1), fp2_mul_xi(t2));
        r.c2 = Fp2::add(Fp2::sub(Fp2::sub(m02, t0), t2), t1);
        return r;
    }
    FF_HD static Fp6 mul_by_v(const Fp6 &a) { return Fp6{fp2_mul_xi(a.c2), a.c0, a.c1}; }
    FF_HD static Fp6 inv(const Fp6 &a) {
        Fp2 A = Fp2::sub(Fp2::sqr(a.c0), fp2_mul_xi(Fp2::mul(a.c1, a.c2)));
        Fp2 B = Fp2::sub(fp2_mul_xi(Fp2::sqr(a.c2)), Fp2::mul(a.c0, a.c1));
        Fp2 C = Fp2::sub(Fp2::sqr(a.c1), Fp2::mul(a.c0, a.c2));
        Fp2 F = Fp2::add(Fp2::mul(a.c0, A), fp2_mul_xi(Fp2::add(Fp2::mul(a.c2, B), Fp2::mul(a.c1, C))));
        Fp2 fi = Fp2::inv(F);
        return Fp6{Fp2::mul(A, fi), Fp2::mul(B, fi), Fp2::mul(C, fi)};
    }
};

struct alignas(16) Fp12 {
    Fp6 c0, c1;
    FF_HD static Fp12 one() { return Fp12{Fp6::one(), Fp6::zero()}; }
    FF_HD bool operator==(const Fp12 &b) const { return c0 == b.c0 && c1 == b.c1; }
    FF_HD static Fp12 mul(const Fp12 &a, const Fp12 &b) {   // 3 Fp6 products
        Fp6 t0 = Fp6::mul(a.c0, b.c0), t1 = Fp6::mul(a.c1, b.c1);
        Fp6 m = Fp6::mul(Fp6::add(a.c0, a.c1), Fp6::add(b.c0, b.c1));
        return Fp12{Fp6::add(t0, Fp6::mul_by_v(t1)), Fp6::sub(Fp6::sub(m, t0), t1)};
    }
    FF_HD static Fp12 sqr(const Fp12 &a) {                  // (c0+c1)(c0+v c1) - t - v t, t = c0 c1: 2 Fp6 products
        Fp6 t = Fp6::mul(a.c0, a.c1);
        Fp6 s = Fp6::mul(Fp6::add(a.c0, a.c1), Fp6::add(a.c0, Fp6::mul_by_v(a.c1)));
        return Fp12{Fp6::sub(Fp6::sub(s, t), Fp6::mul_by_v(t)), Fp6::add(t, t)};
    }
    FF_HD static Fp12 conj(const Fp12 &a) { return Fp12{a.c0, Fp6::neg(a.c1)}; }   // = a^(q^6)
    FF_HD static Fp12 inv(const Fp12 &a) {
        Fp6 t = Fp6::sub(Fp6::mul(a.c0, a.c0), Fp6::mul_by_v(Fp6::mul(a.c1, a.c1)));
        Fp6 ti = Fp6::inv(t);
        return Fp12{Fp6::mul(a.c0, ti), Fp6::neg(Fp6::mul(a.c1, ti))};
    }
    // f * ((a, 0, 0) + (b, c, 0) w): the shape of every line function (a in Fp, b, c in Fp2)
    FF_HD static Fp12 mul_line(const Fp12 &f, const Fp &a, const Fp2 &b, const Fp2 &c) {
        Fp12 l;
        l.c0 = Fp6{Fp2{a, Fp::zero()}, Fp2::zero(), Fp2::zero()};
        l.c1 = Fp6{b, c, Fp2::zero()};
        // t0 = f.c0 * (a,0,0) is a scaling by an Fp element; the rest is the generic schoolbook
        Fp6 t0 = Fp6{fp2_mul_fp(f.c0.c0, a), fp2_mul_fp(f.c0.c1, a), fp2_mul_fp(f.c0.c2, a)};
        Fp6 t1 = Fp6::mul(f.c1, l.c1);
        Fp6 m = Fp6::mul(Fp6::add(f.c0, f.c1), Fp6::add(l.c0, l.c1));
        return Fp12{Fp6::add(t0, Fp6::mul_by_v(t1)), Fp6::sub(Fp6::sub(m, t0), t1)};
    }
};

// Miller loop value f_{6x+2,Q}(P) * l_{[6x+2]Q, pi Q}(P) * l_{[6x+2]Q + pi Q, -pi^2 Q}(P); 1 when either point is infinity
FF_HD Fp12 miller_loop(const ec::G1Affine &P, const ec::G2Affine &Q) {
    Fp12 f = Fp12::one();
    if (P.is_inf() || Q.is_inf()) return f;
    const Fp nyp = Fp::neg(P.y);
    Fp2 tx = Q.x, ty = Q.y;
    // one step: multiply f by the line through T with slope lam, then move T to (x3, y3)
    auto line = [&](const Fp2 &lam) { f = Fp12::mul_line(f, nyp, fp2_mul_fp(lam, P.x), Fp2::sub(ty, Fp2::mul(lam, tx))); };
    auto dbl_step = [&]() {
        Fp2 xx = Fp2::sqr(tx);
        Fp2 lam = Fp2::mul(Fp2::add(Fp2::dbl(xx), xx), Fp2::inv(Fp2::dbl(ty)));
        line(lam);
        Fp2 x3 = Fp2::sub(Fp2::sqr(lam), Fp2::dbl(tx));
        ty = Fp2::sub(Fp2::mul(lam, Fp2::sub(tx, x3)), ty);
        tx = x3;
    };
    auto add_step = [&](const Fp2 &qx, const Fp2 &qy, bool move) {
        Fp2 lam = Fp2::mul(Fp2::sub(qy, ty), Fp2::inv(Fp2::sub(qx, tx)));
        line(lam);
        if (!move) return;
        Fp2 x3 = Fp2::sub(Fp2::sub(Fp2::sqr(lam), tx), qx);
        ty = Fp2::sub(Fp2::mul(lam, Fp2::sub(tx, x3)), ty);
        tx = x3;
    };
    for (int i = pairing_consts::ATE_LOOP_BITS - 2; i >= 0; i--) {
        f = Fp12::sqr(f);
        dbl_step();
        if ((pairing_consts::ATE_LOOP_LO >> i) & 1) add_step(Q.x, Q.y, true);
    }
    Fp2 g12{pairing_consts::G12_A0(), pairing_consts::G12_A1()};
    Fp2 g13{pairing_consts::G13_A0(), pairing_consts::G13_A1()};
    Fp2 q1x = Fp2::mul(fp2_conj(Q.x), g12), q1y = Fp2::mul(fp2_conj(Q.y), g13);   // pi(Q)
    add_step(q1x, q1y, true);
    add_step(fp2_mul_fp(Q.x, pairing_consts::G22()), Q.y, false);         // -pi^2(Q) = (x G22, y)
    return f;
}

// f^((q^12-1)/r)
inline Fp12 final_exponentiation(const Fp12 &f) {
    Fp12 g = Fp12::mul(Fp12::conj(f), Fp12::inv(f));   // f^(q^6 - 1)
    Fp12 acc = Fp12::one();
    for (int i = pairing_consts::FINAL_EXP_BITS - 1; i >= 0; i--) {
        acc = Fp12::sqr(acc);
        if ((pairing_consts::FINAL_EXP[i >> 5] >> (i & 31)) & 1) acc = Fp12::mul(acc, g);
    }
    return acc;
}

}  // namespace pairing

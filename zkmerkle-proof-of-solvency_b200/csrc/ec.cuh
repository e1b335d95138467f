// BN254 G1 / G2 group law for the MSM kernels: extended-Jacobian ("XYZZ") accumulators with affine inputs.
//   x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2;  ZZ = 0 encodes infinity.
// Formulas: EFD "madd-2008-s" (8M+2S), "add-2008-s" (12M+2S), "dbl-2008-s-1", "mdbl-2008-s".
// Replaces gnark-crypto's g1JacExtended / g2JacExtended bucket arithmetic inside G1Jac.MultiExp / G2Jac.MultiExp
// (ecc/bn254/multiexp*.go, out of tree), reached from src/prover/prover/prover.go:269.  Written once over the
// coordinate field F (ff::Fp for G1, ff::Fp2 for G2); memory layout of Affine<F> is gnark-crypto's G1Affine /
// G2Affine (X then Y, Montgomery limbs; (0,0) = infinity).
#pragma once
#include "ff.cuh"

namespace ec {

template <class F>
struct alignas(16) Affine {
    F x, y;
    FF_HD bool is_inf() const { return x.is_zero() && y.is_zero(); }
    FF_HD static Affine inf() { return Affine{F::zero(), F::zero()}; }
};

template <class F>
struct alignas(16) XYZZ {
    F X, Y, ZZ, ZZZ;

    FF_HD bool is_inf() const { return ZZ.is_zero(); }
    FF_HD static XYZZ inf() { return XYZZ{F::one(), F::one(), F::zero(), F::zero()}; }
    FF_HD static XYZZ from_affine(const Affine<F> &p) {
        if (p.is_inf()) return inf();
        return XYZZ{p.x, p.y, F::one(), F::one()};
    }

    // 2 * (affine p)
    FF_HD static XYZZ dbl_affine(const Affine<F> &p) {
        F U = F::dbl(p.y), V = F::sqr(U), W = F::mul(U, V), S = F::mul(p.x, V);
        F xx = F::sqr(p.x), M = F::add(F::dbl(xx), xx);
        XYZZ r;
        r.X = F::sub(F::sqr(M), F::dbl(S));
        r.Y = F::sub(F::mul(M, F::sub(S, r.X)), F::mul(W, p.y));
        r.ZZ = V; r.ZZZ = W;
        return r;
    }
    FF_HD XYZZ dbl() const {
        if (is_inf()) return *this;
        F U = F::dbl(Y), V = F::sqr(U), W = F::mul(U, V), S = F::mul(X, V);
        F xx = F::sqr(X), M = F::add(F::dbl(xx), xx);
        XYZZ r;
        r.X = F::sub(F::sqr(M), F::dbl(S));
        r.Y = F::sub(F::mul(M, F::sub(S, r.X)), F::mul(W, Y));
        r.ZZ = F::mul(V, ZZ); r.ZZZ = F::mul(W, ZZZ);
        return r;
    }
    // this += (neg ? -p : p), p affine -- the bucket-accumulation step
    FF_HD void add_affine(const Affine<F> &p, bool neg) {
        if (p.is_inf()) return;
        F py = neg ? F::neg(p.y) : p.y;
        if (is_inf()) { X = p.x; Y = py; ZZ = F::one(); ZZZ = F::one(); return; }
        F U2 = F::mul(p.x, ZZ), S2 = F::mul(py, ZZZ);
        F Pp = F::sub(U2, X), R = F::sub(S2, Y);
        if (Pp.is_zero()) {
            if (R.is_zero()) { *this = dbl_affine(Affine<F>{p.x, py}); } else { *this = inf(); }
            return;
        }
        F PP = F::sqr(Pp), PPP = F::mul(Pp, PP), Q = F::mul(X, PP);
        F X3 = F::sub(F::sub(F::sqr(R), PPP), F::dbl(Q));
        Y = F::sub(F::mul(R, F::sub(Q, X3)), F::mul(Y, PPP));
        X = X3;
        ZZ = F::mul(ZZ, PP); ZZZ = F::mul(ZZZ, PPP);
    }
    // this += q
    FF_HD void add(const XYZZ &q) {
        if (q.is_inf()) return;
        if (is_inf()) { *this = q; return; }
        F U1 = F::mul(X, q.ZZ), U2 = F::mul(q.X, ZZ), S1 = F::mul(Y, q.ZZZ), S2 = F::mul(q.Y, ZZZ);
        F Pp = F::sub(U2, U1), R = F::sub(S2, S1);
        if (Pp.is_zero()) {
            if (R.is_zero()) { *this = dbl(); } else { *this = inf(); }
            return;
        }
        F PP = F::sqr(Pp), PPP = F::mul(Pp, PP), Q = F::mul(U1, PP);
        F X3 = F::sub(F::sub(F::sqr(R), PPP), F::dbl(Q));
        Y = F::sub(F::mul(R, F::sub(Q, X3)), F::mul(S1, PPP));
        X = X3;
        ZZ = F::mul(F::mul(ZZ, q.ZZ), PP); ZZZ = F::mul(F::mul(ZZZ, q.ZZZ), PPP);
    }
    FF_HD XYZZ negated() const { XYZZ r = *this; r.Y = F::neg(Y); return r; }

    // k * this for a small plain integer k (double-and-add, MSB first)
    FF_HD XYZZ mul_u32(uint32_t k) const {
        XYZZ acc = inf();
        for (int i = 31; i >= 0; i--) {
            acc = acc.dbl();
            if ((k >> i) & 1) acc.add(*this);
        }
        return acc;
    }
    // k * this, k = 8 x u32 plain little-endian
    FF_HD XYZZ mul_256(const uint32_t k[8]) const {
        XYZZ acc = inf();
        for (int i = 255; i >= 0; i--) {
            acc = acc.dbl();
            if ((k[i >> 5] >> (i & 31)) & 1) acc.add(*this);
        }
        return acc;
    }
    FF_HD Affine<F> to_affine() const {
        if (is_inf()) return Affine<F>::inf();
        F zi = F::inv(ZZZ);              // 1/ZZZ
        F y = F::mul(Y, zi);
        F zz_inv = F::sqr(F::mul(zi, ZZ));   // (ZZ/ZZZ)^2 = ZZ^2/ZZ^3 = 1/ZZ   (invariant ZZ^3 = ZZZ^2)
        return Affine<F>{F::mul(X, zz_inv), y};
    }
};

using G1Affine = Affine<ff::Fp>;
using G2Affine = Affine<ff::Fp2>;
using G1XYZZ = XYZZ<ff::Fp>;
using G2XYZZ = XYZZ<ff::Fp2>;

}  // namespace ec

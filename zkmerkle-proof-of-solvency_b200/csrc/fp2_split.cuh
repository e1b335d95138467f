// Fp2 arithmetic on values SPLIT over a lane pair (msm_g2pair.cu): lane 2k holds the a0 component, lane 2k+1 the a1 component.
// Written once over a value type V and a lane context L:
//   device: V = ff::Fp, L = LanePair (role = lane & 1, xchg = 8 x shfl.xor with the partner lane)
//   host  : V = HostPair (both components side by side), L = HostLanes (xchg swaps the components, sel picks per component)
// so that the exact formula text the GPU runs is checked on the CPU against ec::XYZZ<Fp2> (hosttest.cpp, tests/test_host_ff.py).
// Every function is branch-free across the two roles: the even and odd lanes of a warp never diverge.
#pragma once
#include "ec.cuh"

namespace fp2split {
using ff::Fp;

// ---- host emulation of a lane pair -------------------------------------------------------------------------------
struct HostPair {
    Fp c[2];
    static HostPair add(const HostPair &a, const HostPair &b) { return HostPair{{Fp::add(a.c[0], b.c[0]), Fp::add(a.c[1], b.c[1])}}; }
    static HostPair sub(const HostPair &a, const HostPair &b) { return HostPair{{Fp::sub(a.c[0], b.c[0]), Fp::sub(a.c[1], b.c[1])}}; }
    static HostPair mul(const HostPair &a, const HostPair &b) { return HostPair{{Fp::mul(a.c[0], b.c[0]), Fp::mul(a.c[1], b.c[1])}}; }
    static HostPair dbl(const HostPair &a) { return add(a, a); }
    static HostPair neg(const HostPair &a) { return HostPair{{Fp::neg(a.c[0]), Fp::neg(a.c[1])}}; }
};
struct HostLanes {
    HostPair xchg(const HostPair &v) const { return HostPair{{v.c[1], v.c[0]}}; }
    // role ? a : b, per lane: component 0 is role 0 (takes b), component 1 is role 1 (takes a)
    HostPair sel_role(const HostPair &a, const HostPair &b) const { return HostPair{{b.c[0], a.c[1]}}; }
};

// ---- the shared formula text ----------------------------------------------------------------------------------------
// z = x*y, t = u*v: Karatsuba spread evenly -- lane0: x0y0, u0v0, (x0+x1)(y0+y1); lane1: x1y1, u1v1, (u0+u1)(v0+v1)
template <class V, class L>
FF_HD void mul2(const L &ln, const V &x, const V &y, const V &u, const V &v, V &z, V &t) {
    const V r1 = ln.xchg(ln.sel_role(x, u)), r2 = ln.xchg(ln.sel_role(y, v));   // lane0 receives (x1, y1), lane1 (u0, v0)
    const V own1 = V::mul(x, y);                                                 // lane0: A = x0 y0      lane1: C = x1 y1
    const V own2 = V::mul(u, v);                                                 // lane0: B = u0 v0      lane1: D = u1 v1
    const V m = V::mul(V::add(ln.sel_role(u, x), r1), V::add(ln.sel_role(v, y), r2));   // lane0: M1, lane1: M2
    const V e1 = ln.xchg(ln.sel_role(own1, V::sub(m, own1)));                    // lane0 receives C      lane1 receives M1 - A
    const V e2 = ln.xchg(own2);                                                  // lane0 receives D      lane1 receives B
    z = ln.sel_role(V::sub(e1, own1), V::sub(own1, e1));                         // lane0: A - C          lane1: (M1 - A) - C
    t = ln.sel_role(V::sub(V::sub(m, e2), own2), V::sub(own2, e2));              // lane0: B - D          lane1: M2 - B - D
}
// a^2: lane0 (a0+a1)(a0-a1), lane1 2 a0 a1
template <class V, class L>
FF_HD V sqr(const L &ln, const V &a) {
    const V ao = ln.xchg(a);
    const V r = V::mul(ln.sel_role(a, V::add(a, ao)), ln.sel_role(ao, V::sub(a, ao)));
    return ln.sel_role(V::dbl(r), r);
}

template <class V> struct Acc { V X, Y, ZZ, ZZZ; };

// madd-2008-s: acc + (px, py), both finite and with different x (the callers select the special cases from P and R:
// P = x ZZ - X = 0 flags equal x, then R = y ZZZ - Y = 0 a doubling and R != 0 a cancellation)
template <class V, class L>
FF_HD Acc<V> madd(const L &ln, const Acc<V> &acc, const V &px, const V &py, V &P, V &R) {
    V U2, S2;
    mul2(ln, px, acc.ZZ, py, acc.ZZZ, U2, S2);
    P = V::sub(U2, acc.X); R = V::sub(S2, acc.Y);
    const V PP = sqr(ln, P);
    V PPP, Q;
    mul2(ln, P, PP, acc.X, PP, PPP, Q);
    Acc<V> r;
    r.X = V::sub(V::sub(sqr(ln, R), PPP), V::dbl(Q));
    V t1, t2;
    mul2(ln, R, V::sub(Q, r.X), acc.Y, PPP, t1, t2);
    r.Y = V::sub(t1, t2);
    mul2(ln, acc.ZZ, PP, acc.ZZZ, PPP, r.ZZ, r.ZZZ);
    return r;
}
// mdbl-2008-s: 2 * (px, py)
template <class V, class L>
FF_HD Acc<V> dbl_affine(const L &ln, const V &px, const V &py) {
    const V U = V::dbl(py);
    const V Vv = sqr(ln, U);
    V W, S;
    mul2(ln, U, Vv, px, Vv, W, S);
    const V xx = sqr(ln, px);
    const V M = V::add(V::dbl(xx), xx);
    Acc<V> r;
    r.X = V::sub(sqr(ln, M), V::dbl(S));
    V t1, t2;
    mul2(ln, M, V::sub(S, r.X), W, py, t1, t2);
    r.Y = V::sub(t1, t2);
    r.ZZ = Vv; r.ZZZ = W;
    return r;
}

}  // namespace fp2split

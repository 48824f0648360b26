// Pallas / Vesta group arithmetic (y^2 = x^3 + 5, a = 0) in extended Jacobian "XYZZ" coordinates.
//
// GPU replacement for ark-ec 0.3 `short_weierstrass_jacobian` as used by `VariableBaseMSM`
// (SURVEY row a10).  The group law is unique, so after normalising to affine every result is
// bit-identical to the arkworks one regardless of the coordinate system used on the way.
//
// XYZZ: x = X/ZZ, y = Y/ZZZ with ZZ^3 = ZZZ^2; ZZ == 0 encodes the identity.
// Formulas: EFD "madd-2008-s", "mdbl-2008-s-1", "add-2008-s", "dbl-2008-s-1".
#pragma once
#include "fe.cuh"

namespace pasta {

// Cold group operations are real function calls (one copy per kernel) -- keeps ptxas time and the
// instruction-cache footprint down; the hot mixed addition stays inlined.
#ifdef __CUDACC__
#define PASTA_HD_COLD __host__ __device__ __noinline__
#else
#define PASTA_HD_COLD
#endif

struct alignas(16) affine {
    fe x, y;  // Montgomery form; (0, 0) encodes the identity (never on y^2 = x^3 + 5)
};
struct alignas(16) xyzz {
    fe x, y, zz, zzz;
};

// C tags the curve through its *base* field parameters.
template <class F>
struct Ec {
    using fd = Fd<F>;

    PASTA_HD static xyzz identity() {
        xyzz r;
        r.x = fe_zero();
        r.y = fe_zero();
        r.zz = fe_zero();
        r.zzz = fe_zero();
        return r;
    }
    PASTA_HD static bool is_identity(const xyzz &p) { return fe_is_zero(p.zz); }
    PASTA_HD static bool is_identity(const affine &p) { return fe_is_zero(p.x) && fe_is_zero(p.y); }

    PASTA_HD static xyzz from_affine(const affine &q) {
        xyzz r;
        if (is_identity(q)) return identity();
        r.x = q.x;
        r.y = q.y;
        r.zz = fd::one();
        r.zzz = fd::one();
        return r;
    }

    // 2 * (affine q), q != identity
    PASTA_HD_COLD static xyzz dbl_affine(const affine &q) {
        xyzz r;
        fe U = fd::dbl(q.y);
        fe V = fd::sqr(U);
        fe W = fd::mul(U, V);
        fe S = fd::mul(q.x, V);
        fe X2 = fd::sqr(q.x);
        fe M = fd::add(fd::dbl(X2), X2);
        r.x = fd::sub(fd::sqr(M), fd::dbl(S));
        r.y = fd::dot2(M, fd::sub(S, r.x), W, fd::neg(q.y));  // difference of two products, one reduction
        r.zz = V;
        r.zzz = W;
        return r;
    }

    PASTA_HD_COLD static xyzz dbl(const xyzz &p) {
        if (is_identity(p)) return p;
        xyzz r;
        fe U = fd::dbl(p.y);
        fe V = fd::sqr(U);
        fe W = fd::mul(U, V);
        fe S = fd::mul(p.x, V);
        fe X2 = fd::sqr(p.x);
        fe M = fd::add(fd::dbl(X2), X2);
        r.x = fd::sub(fd::sqr(M), fd::dbl(S));
        r.y = fd::dot2(M, fd::sub(S, r.x), W, fd::neg(p.y));
        r.zz = fd::mul(V, p.zz);
        r.zzz = fd::mul(W, p.zzz);
        return r;
    }

    // p += q (q affine).  Handles identity, doubling and cancellation.
    PASTA_HD static void add_mixed(xyzz &p, const affine &q) {
        if (is_identity(q)) return;
        if (is_identity(p)) {
            p = from_affine(q);
            return;
        }
        fe U2 = fd::mul(q.x, p.zz);
        fe S2 = fd::mul(q.y, p.zzz);
        fe P = fd::sub(U2, p.x);
        fe R = fd::sub(S2, p.y);
        if (fe_is_zero(P)) {
            if (fe_is_zero(R))
                p = dbl_affine(q);
            else
                p = identity();
            return;
        }
        fe PP = fd::sqr(P);
        fe PPP = fd::mul(P, PP);
        fe Q = fd::mul(p.x, PP);
        fe X3 = fd::sub(fd::sub(fd::sqr(R), PPP), fd::dbl(Q));
        fe Y3 = fd::dot2(R, fd::sub(Q, X3), fd::neg(p.y), PPP);
        p.x = X3;
        p.y = Y3;
        p.zz = fd::mul(p.zz, PP);
        p.zzz = fd::mul(p.zzz, PPP);
    }

    PASTA_HD_COLD static void add(xyzz &p, const xyzz &q) {
        if (is_identity(q)) return;
        if (is_identity(p)) {
            p = q;
            return;
        }
        fe U1 = fd::mul(p.x, q.zz);
        fe U2 = fd::mul(q.x, p.zz);
        fe S1 = fd::mul(p.y, q.zzz);
        fe S2 = fd::mul(q.y, p.zzz);
        fe P = fd::sub(U2, U1);
        fe R = fd::sub(S2, S1);
        if (fe_is_zero(P)) {
            if (fe_is_zero(R))
                p = dbl(p);
            else
                p = identity();
            return;
        }
        fe PP = fd::sqr(P);
        fe PPP = fd::mul(P, PP);
        fe Q = fd::mul(U1, PP);
        fe X3 = fd::sub(fd::sub(fd::sqr(R), PPP), fd::dbl(Q));
        fe Y3 = fd::dot2(R, fd::sub(Q, X3), fd::neg(S1), PPP);
        p.x = X3;
        p.y = Y3;
        p.zz = fd::mul(fd::mul(p.zz, q.zz), PP);
        p.zzz = fd::mul(fd::mul(p.zzz, q.zzz), PPP);
    }

    PASTA_HD static affine neg(const affine &q) {
        affine r;
        r.x = q.x;
        r.y = fd::neg(q.y);
        return r;
    }

    // Normalise (one field inversion).  Identity -> (0, 0).
    PASTA_HD_COLD static affine to_affine(const xyzz &p) {
        affine r;
        if (is_identity(p)) {
            r.x = fe_zero();
            r.y = fe_zero();
            return r;
        }
        // 1/ZZZ = i3; 1/ZZ = ZZZ * i3 * ... use: ZZ^3 = ZZZ^2  =>  1/ZZ = (ZZ * i3)^2
        fe i3 = fd::inv(p.zzz);
        fe t = fd::mul(p.zz, i3);  // ZZ/ZZZ = 1/Z
        fe i2 = fd::sqr(t);        // 1/ZZ
        r.x = fd::mul(p.x, i2);
        r.y = fd::mul(p.y, i3);
        return r;
    }
};

}  // namespace pasta

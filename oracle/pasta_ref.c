/*
 * CPU oracle (TEST INFRASTRUCTURE ONLY) -- plain-C restatement of the reference's CPU arithmetic.
 *
 * What it restates (all un-vendored, pinned in AL/operator/mina/lib/Cargo.lock):
 *   - ark-ff 0.3 Fp256 Montgomery arithmetic (4 x 64-bit limbs, R = 2^256)
 *   - ark-ec 0.3 `VariableBaseMSM::multi_scalar_mul` (bucket method, window c = ln(n)+2, one
 *     rayon task per window) -- lambdaclass/openmina_algebra @ 017531e
 *   - poly-commitment `b_poly_coefficients`, kimchi `ScalarChallenge::to_field`,
 *     `SRS::create` hash-to-curve (blake2b-512 + groupmap) -- openmina-proof-systems @ 44e0d3b
 *   - mina-poseidon `ArithmeticSponge` permutation (55 full rounds, x^7, MDS then round constant)
 * reference call sites: AL/operator/mina/lib/src/lib.rs:34,99-111; verifier_index.rs:169,204-208;
 * AL/operator/mina_account/lib/src/merkle_verifier.rs:27.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg load this library.  The
 * shipped verifier never links it.
 *
 * Parity: field/curve/MSM/b_poly/endo/SRS routines are PINNED against the accumulator KATs in
 * tests/golden/mina_state.proof and against srs/{vesta,pallas}.srs (tests/test_oracle_kats.py).
 * oracle_poseidon_permute is PARITY UNPINNED: it takes the MDS + round constants as data and no
 * table that passes the reference's known-answer test (merkle_verifier.rs:43-58) is available
 * here -- see DESIGN.md section 0.  The reference binary itself was never run (unbuildable here).
 *
 * All field elements cross the ABI as 32-byte little-endian canonical integers.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef struct { uint64_t l[4]; } fe;

typedef struct {
    fe p;        /* modulus */
    fe r;        /* R mod p  (Montgomery one) */
    fe r2;       /* R^2 mod p */
    uint64_t ninv; /* -p^-1 mod 2^64 */
    fe five;     /* curve coefficient b = 5 (Montgomery) */
    fe root_of_unity; /* 5^((p-1)/2^32), Montgomery */
    fe t_minus1_div2; /* ((p-1)/2^32 - 1)/2, plain integer */
    fe half_p;   /* (p-1)/2 plain */
} field_ctx;

static const fe MOD_P = {{0x992d30ed00000001ULL, 0x224698fc094cf91bULL, 0x0ULL, 0x4000000000000000ULL}};
static const fe MOD_Q = {{0x8c46eb2100000001ULL, 0x224698fc0994a8ddULL, 0x0ULL, 0x4000000000000000ULL}};

static field_ctx CTX[2]; /* 0 = Fp, 1 = Fq */
static int ctx_ready = 0;

static int fe_is_zero(const fe *a) { return (a->l[0] | a->l[1] | a->l[2] | a->l[3]) == 0; }
static int fe_eq(const fe *a, const fe *b) { return memcmp(a, b, sizeof(fe)) == 0; }
static int fe_geq(const fe *a, const fe *b) {
    for (int i = 3; i >= 0; i--) {
        if (a->l[i] > b->l[i]) return 1;
        if (a->l[i] < b->l[i]) return 0;
    }
    return 1;
}
static uint64_t fe_add_raw(fe *o, const fe *a, const fe *b) {
    u128 c = 0;
    for (int i = 0; i < 4; i++) { c += (u128)a->l[i] + b->l[i]; o->l[i] = (uint64_t)c; c >>= 64; }
    return (uint64_t)c;
}
static uint64_t fe_sub_raw(fe *o, const fe *a, const fe *b) {
    uint64_t borrow = 0;
    for (int i = 0; i < 4; i++) {
        u128 d = (u128)a->l[i] - b->l[i] - borrow;
        o->l[i] = (uint64_t)d; borrow = (uint64_t)(d >> 64) & 1;
    }
    return borrow;
}
static void f_add(const field_ctx *F, fe *o, const fe *a, const fe *b) {
    fe t; fe_add_raw(&t, a, b);           /* p < 2^255: no carry out */
    if (fe_geq(&t, &F->p)) fe_sub_raw(&t, &t, &F->p);
    *o = t;
}
static void f_sub(const field_ctx *F, fe *o, const fe *a, const fe *b) {
    fe t; if (fe_sub_raw(&t, a, b)) fe_add_raw(&t, &t, &F->p);
    *o = t;
}
static void f_neg(const field_ctx *F, fe *o, const fe *a) {
    if (fe_is_zero(a)) { *o = *a; return; }
    fe_sub_raw(o, &F->p, a);
}
static void f_dbl(const field_ctx *F, fe *o, const fe *a) { f_add(F, o, a, a); }

/* CIOS Montgomery product, fully unrolled over 4 x 64-bit limbs (what ark-ff's Fp256 does; the modulus
 * limb 2 is zero for both Pasta primes and is skipped).  -O3 -march=native turns the u128 products into
 * mulx/adc chains. */
#define MUL_ROUND(bi)                                                          \
    do {                                                                       \
        u128 c = (u128)a0 * (bi) + t0; t0 = (uint64_t)c; c >>= 64;             \
        c += (u128)a1 * (bi) + t1; t1 = (uint64_t)c; c >>= 64;                 \
        c += (u128)a2 * (bi) + t2; t2 = (uint64_t)c; c >>= 64;                 \
        c += (u128)a3 * (bi) + t3; t3 = (uint64_t)c; c >>= 64;                 \
        c += t4; t4 = (uint64_t)c; uint64_t t5 = (uint64_t)(c >> 64);          \
        uint64_t m = t0 * ninv;                                                \
        c = (u128)m * p0 + t0; c >>= 64;                                       \
        c += (u128)m * p1 + t1; t0 = (uint64_t)c; c >>= 64;                    \
        c += t2; t1 = (uint64_t)c; c >>= 64;                                   \
        c += (u128)m * p3 + t3; t2 = (uint64_t)c; c >>= 64;                    \
        c += t4; t3 = (uint64_t)c; t4 = t5 + (uint64_t)(c >> 64);              \
    } while (0)
static void f_mul(const field_ctx *F, fe *o, const fe *a, const fe *b) {
    const uint64_t a0 = a->l[0], a1 = a->l[1], a2 = a->l[2], a3 = a->l[3];
    const uint64_t p0 = F->p.l[0], p1 = F->p.l[1], p3 = F->p.l[3], ninv = F->ninv;
    uint64_t t0 = 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0;
    MUL_ROUND(b->l[0]);
    MUL_ROUND(b->l[1]);
    MUL_ROUND(b->l[2]);
    MUL_ROUND(b->l[3]);
    fe r = {{t0, t1, t2, t3}};
    if (t4 || fe_geq(&r, &F->p)) fe_sub_raw(&r, &r, &F->p);
    *o = r;
}
static void f_sqr(const field_ctx *F, fe *o, const fe *a) { f_mul(F, o, a, a); }
static void f_to_mont(const field_ctx *F, fe *o, const fe *a) { f_mul(F, o, a, &F->r2); }
static void f_from_mont(const field_ctx *F, fe *o, const fe *a) {
    fe one = {{1, 0, 0, 0}}; f_mul(F, o, a, &one);
}
/* a^e, a Montgomery, e plain 256-bit */
static void f_pow(const field_ctx *F, fe *o, const fe *a, const fe *e) {
    fe acc = F->r;
    for (int i = 255; i >= 0; i--) {
        f_sqr(F, &acc, &acc);
        if ((e->l[i / 64] >> (i % 64)) & 1) f_mul(F, &acc, &acc, a);
    }
    *o = acc;
}
static void f_inv(const field_ctx *F, fe *o, const fe *a) {
    fe e = F->p; e.l[0] -= 2; f_pow(F, o, a, &e);
}
static void fe_shr(fe *a, int k) { /* k < 64 */
    for (int i = 0; i < 4; i++) {
        uint64_t hi = (i < 3) ? a->l[i + 1] : 0;
        a->l[i] = k ? ((a->l[i] >> k) | (hi << (64 - k))) : a->l[i];
    }
}

static void ctx_init_one(field_ctx *F, const fe *p) {
    F->p = *p;
    uint64_t inv = 1;
    for (int i = 0; i < 6; i++) inv *= 2 - p->l[0] * inv;
    F->ninv = (uint64_t)0 - inv;
    /* R mod p by 256 doublings of 1; R^2 by 256 more */
    fe x = {{1, 0, 0, 0}};
    for (int i = 0; i < 256; i++) f_dbl(F, &x, &x);
    F->r = x;
    for (int i = 0; i < 256; i++) f_dbl(F, &x, &x);
    F->r2 = x;
    fe five = {{5, 0, 0, 0}}; f_to_mont(F, &F->five, &five);
    fe t = *p; t.l[0] -= 1;           /* p-1 */
    fe half = t; fe_shr(&half, 1); F->half_p = half;
    fe todd = t; fe_shr(&todd, 32);   /* (p-1)/2^32 */
    f_pow(F, &F->root_of_unity, &F->five, &todd);
    fe tm = todd; tm.l[0] -= 1; fe_shr(&tm, 1); F->t_minus1_div2 = tm;
}
static void ctx_init(void) {
    if (ctx_ready) return;
    ctx_init_one(&CTX[0], &MOD_P);
    ctx_init_one(&CTX[1], &MOD_Q);
    ctx_ready = 1;
}

static void fe_from_bytes(fe *o, const uint8_t *b) { memcpy(o->l, b, 32); }
static void fe_to_bytes(uint8_t *b, const fe *a) { memcpy(b, a->l, 32); }

/* ark-ff 0.3 sqrt (Tonelli-Shanks as in SURVEY Appendix B.8). returns 0 when non-residue. */
static int f_sqrt(const field_ctx *F, fe *o, const fe *a) {
    if (fe_is_zero(a)) { *o = *a; return 1; }
    fe leg; f_pow(F, &leg, a, &F->half_p);
    if (!fe_eq(&leg, &F->r)) return 0;
    fe z = F->root_of_unity, w, x, b;
    f_pow(F, &w, a, &F->t_minus1_div2);
    f_mul(F, &x, a, &w);
    f_mul(F, &b, &x, &w);
    int v = 32;
    while (!fe_eq(&b, &F->r)) {
        int k = 0; fe b2k = b;
        while (!fe_eq(&b2k, &F->r)) { f_sqr(F, &b2k, &b2k); k++; }
        int j = v - k - 1;
        w = z;
        for (int i = 0; i < j; i++) f_sqr(F, &w, &w);
        f_sqr(F, &z, &w);
        f_mul(F, &b, &b, &z);
        f_mul(F, &x, &x, &w);
        v = k;
    }
    *o = x; return 1;
}

/* ---------------------------------------------------------------- curve: Jacobian, a = 0 ----- */
typedef struct { fe x, y, z; } jac;   /* z == 0 -> identity */
typedef struct { fe x, y; int inf; } aff;

static void j_set_inf(const field_ctx *F, jac *o) { o->x = F->r; o->y = F->r; memset(&o->z, 0, sizeof(fe)); }
static int j_is_inf(const jac *a) { return fe_is_zero(&a->z); }

static void j_double(const field_ctx *F, jac *o, const jac *p) {
    if (j_is_inf(p)) { *o = *p; return; }
    fe A, B, C, D, E, Fv, t, X3, Y3, Z3;
    f_sqr(F, &A, &p->x); f_sqr(F, &B, &p->y); f_sqr(F, &C, &B);
    f_add(F, &t, &p->x, &B); f_sqr(F, &t, &t); f_sub(F, &t, &t, &A); f_sub(F, &t, &t, &C); f_dbl(F, &D, &t);
    f_dbl(F, &E, &A); f_add(F, &E, &E, &A);
    f_sqr(F, &Fv, &E);
    f_dbl(F, &t, &D); f_sub(F, &X3, &Fv, &t);
    f_sub(F, &t, &D, &X3); f_mul(F, &Y3, &E, &t);
    f_dbl(F, &t, &C); f_dbl(F, &t, &t); f_dbl(F, &t, &t); f_sub(F, &Y3, &Y3, &t);
    f_mul(F, &Z3, &p->y, &p->z); f_dbl(F, &Z3, &Z3);
    o->x = X3; o->y = Y3; o->z = Z3;
}
static void j_add_mixed(const field_ctx *F, jac *o, const jac *p, const aff *q) {
    if (q->inf) { *o = *p; return; }
    if (j_is_inf(p)) { o->x = q->x; o->y = q->y; o->z = F->r; return; }
    fe Z1Z1, U2, S2, H, HH, I, J, r, V, t, X3, Y3, Z3;
    f_sqr(F, &Z1Z1, &p->z);
    f_mul(F, &U2, &q->x, &Z1Z1);
    f_mul(F, &S2, &q->y, &p->z); f_mul(F, &S2, &S2, &Z1Z1);
    if (fe_eq(&U2, &p->x)) {
        if (fe_eq(&S2, &p->y)) { j_double(F, o, p); return; }
        j_set_inf(F, o); return;
    }
    f_sub(F, &H, &U2, &p->x);
    f_sqr(F, &HH, &H);
    f_dbl(F, &I, &HH); f_dbl(F, &I, &I);
    f_mul(F, &J, &H, &I);
    f_sub(F, &r, &S2, &p->y); f_dbl(F, &r, &r);
    f_mul(F, &V, &p->x, &I);
    f_sqr(F, &X3, &r); f_sub(F, &X3, &X3, &J); f_dbl(F, &t, &V); f_sub(F, &X3, &X3, &t);
    f_sub(F, &t, &V, &X3); f_mul(F, &Y3, &r, &t);
    f_mul(F, &t, &p->y, &J); f_dbl(F, &t, &t); f_sub(F, &Y3, &Y3, &t);
    f_add(F, &Z3, &p->z, &H); f_sqr(F, &Z3, &Z3); f_sub(F, &Z3, &Z3, &Z1Z1); f_sub(F, &Z3, &Z3, &HH);
    o->x = X3; o->y = Y3; o->z = Z3;
}
static void j_add(const field_ctx *F, jac *o, const jac *p, const jac *q) {
    if (j_is_inf(p)) { *o = *q; return; }
    if (j_is_inf(q)) { *o = *p; return; }
    fe Z1Z1, Z2Z2, U1, U2, S1, S2, H, I, J, r, V, t, X3, Y3, Z3;
    f_sqr(F, &Z1Z1, &p->z); f_sqr(F, &Z2Z2, &q->z);
    f_mul(F, &U1, &p->x, &Z2Z2); f_mul(F, &U2, &q->x, &Z1Z1);
    f_mul(F, &S1, &p->y, &q->z); f_mul(F, &S1, &S1, &Z2Z2);
    f_mul(F, &S2, &q->y, &p->z); f_mul(F, &S2, &S2, &Z1Z1);
    if (fe_eq(&U1, &U2)) {
        if (fe_eq(&S1, &S2)) { j_double(F, o, p); return; }
        j_set_inf(F, o); return;
    }
    f_sub(F, &H, &U2, &U1);
    f_dbl(F, &I, &H); f_sqr(F, &I, &I);
    f_mul(F, &J, &H, &I);
    f_sub(F, &r, &S2, &S1); f_dbl(F, &r, &r);
    f_mul(F, &V, &U1, &I);
    f_sqr(F, &X3, &r); f_sub(F, &X3, &X3, &J); f_dbl(F, &t, &V); f_sub(F, &X3, &X3, &t);
    f_sub(F, &t, &V, &X3); f_mul(F, &Y3, &r, &t);
    f_mul(F, &t, &S1, &J); f_dbl(F, &t, &t); f_sub(F, &Y3, &Y3, &t);
    f_add(F, &Z3, &p->z, &q->z); f_sqr(F, &Z3, &Z3); f_sub(F, &Z3, &Z3, &Z1Z1); f_sub(F, &Z3, &Z3, &Z2Z2);
    f_mul(F, &Z3, &Z3, &H);
    o->x = X3; o->y = Y3; o->z = Z3;
}
static void j_to_affine_bytes(const field_ctx *F, const jac *p, uint8_t *out64, int *is_inf) {
    if (j_is_inf(p)) { memset(out64, 0, 64); *is_inf = 1; return; }
    fe zi, zi2, x, y;
    f_inv(F, &zi, &p->z); f_sqr(F, &zi2, &zi);
    f_mul(F, &x, &p->x, &zi2); f_mul(F, &y, &p->y, &zi2); f_mul(F, &y, &y, &zi);
    f_from_mont(F, &x, &x); f_from_mont(F, &y, &y);
    fe_to_bytes(out64, &x); fe_to_bytes(out64 + 32, &y); *is_inf = 0;
}

/* ------------------------------------------------------------------------ MSM (ark-ec 0.3) ---- */
static int ark_ln_without_floats(size_t a) {
    /* ark_std::log2(a) = ceil(log2(a)); ln ~= log2 * 69 / 100 */
    int lg = 0; while (((size_t)1 << lg) < a) lg++;
    return lg * 69 / 100;
}
int oracle_ark_window_bits(size_t n) { return n < 32 ? 3 : ark_ln_without_floats(n) + 2; }

typedef struct {
    const field_ctx *F; size_t n; const fe *scalars; const aff *bases; int c; int w_start; jac result;
} window_job;

static void window_run(window_job *job) {
    const field_ctx *F = job->F; int c = job->c; int w_start = job->w_start;
    size_t nb = ((size_t)1 << c) - 1;
    jac res; j_set_inf(F, &res);
    jac *buckets = (jac *)malloc(nb * sizeof(jac));
    for (size_t i = 0; i < nb; i++) j_set_inf(F, &buckets[i]);
    fe one = {{1, 0, 0, 0}};
    for (size_t i = 0; i < job->n; i++) {
        const fe *s = &job->scalars[i];
        if (fe_is_zero(s)) continue;
        if (fe_eq(s, &one)) { if (w_start == 0) j_add_mixed(F, &res, &res, &job->bases[i]); continue; }
        /* scalar >> w_start, low c bits */
        int limb = w_start / 64, sh = w_start % 64;
        uint64_t v = s->l[limb] >> sh;
        if (sh && limb < 3) v |= s->l[limb + 1] << (64 - sh);
        v &= ((uint64_t)1 << c) - 1;
        if (v) j_add_mixed(F, &buckets[v - 1], &buckets[v - 1], &job->bases[i]);
    }
    jac running; j_set_inf(F, &running);
    for (size_t i = nb; i-- > 0;) {
        j_add(F, &running, &running, &buckets[i]);
        j_add(F, &res, &res, &running);
    }
    free(buckets);
    job->result = res;
}

typedef struct { window_job *jobs; int njobs; int next; pthread_mutex_t mu; } job_queue;
static void *worker(void *arg) {
    job_queue *q = (job_queue *)arg;
    for (;;) {
        pthread_mutex_lock(&q->mu);
        int k = q->next < q->njobs ? q->next++ : -1;
        pthread_mutex_unlock(&q->mu);
        if (k < 0) break;
        window_run(&q->jobs[k]);
    }
    return NULL;
}

/* field_id: 0 -> points over Fp (Pallas, scalars in Fq); 1 -> points over Fq (Vesta).
 * scalars: n x 32 bytes canonical; points: n x 64 bytes canonical affine (x||y), all-zero = infinity.
 * c_override: 0 = arkworks' choice.  Returns 1 if the result is the identity, else 0. */
int oracle_msm(int field_id, size_t n, const uint8_t *scalars, const uint8_t *points,
               uint8_t *out64, int nthreads, int c_override) {
    ctx_init();
    const field_ctx *F = &CTX[field_id];
    fe *sc = (fe *)malloc((n ? n : 1) * sizeof(fe));
    aff *bs = (aff *)malloc((n ? n : 1) * sizeof(aff));
    for (size_t i = 0; i < n; i++) {
        fe_from_bytes(&sc[i], scalars + 32 * i);
        fe x, y; fe_from_bytes(&x, points + 64 * i); fe_from_bytes(&y, points + 64 * i + 32);
        bs[i].inf = fe_is_zero(&x) && fe_is_zero(&y);
        f_to_mont(F, &bs[i].x, &x); f_to_mont(F, &bs[i].y, &y);
    }
    int c = c_override ? c_override : oracle_ark_window_bits(n);
    int nwin = (255 + c - 1) / c;
    window_job *jobs = (window_job *)calloc(nwin, sizeof(window_job));
    for (int w = 0; w < nwin; w++) {
        jobs[w].F = F; jobs[w].n = n; jobs[w].scalars = sc; jobs[w].bases = bs; jobs[w].c = c; jobs[w].w_start = w * c;
    }
    job_queue q; q.jobs = jobs; q.njobs = nwin; q.next = 0; pthread_mutex_init(&q.mu, NULL);
    if (nthreads <= 1) worker(&q);
    else {
        pthread_t th[256]; if (nthreads > 256) nthreads = 256;
        for (int t = 0; t < nthreads; t++) pthread_create(&th[t], NULL, worker, &q);
        for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    }
    jac total = jobs[nwin - 1].result;
    for (int w = nwin - 2; w >= 0; w--) {
        for (int k = 0; k < c; k++) j_double(F, &total, &total);
        j_add(F, &total, &total, &jobs[w].result);
    }
    int is_inf; j_to_affine_bytes(F, &total, out64, &is_inf);
    free(jobs); free(sc); free(bs);
    return is_inf;
}

/* ----------------------------------------------------- point decompression (ark-serialize 0.3) */
int oracle_decompress(int field_id, size_t n, const uint8_t *in33, uint8_t *out64) {
    ctx_init();
    const field_ctx *F = &CTX[field_id];
    for (size_t i = 0; i < n; i++) {
        const uint8_t *b = in33 + 33 * i; uint8_t *o = out64 + 64 * i;
        if (b[32] & 0x40) { memset(o, 0, 64); continue; }
        fe x, xm, y, t; fe_from_bytes(&x, b);
        if (fe_geq(&x, &F->p)) return -1;
        f_to_mont(F, &xm, &x);
        f_sqr(F, &t, &xm); f_mul(F, &t, &t, &xm); f_add(F, &t, &t, &F->five);
        if (!f_sqrt(F, &y, &t)) return -1;
        f_from_mont(F, &y, &y);
        /* flag 0x80 <=> y > (p-1)/2 */
        int big = !fe_geq(&F->half_p, &y);
        if (((b[32] & 0x80) != 0) != big) fe_sub_raw(&y, &F->p, &y);
        fe_to_bytes(o, &x); fe_to_bytes(o + 32, &y);
    }
    return 0;
}

/* -------------------------------------------------------------------------------- blake2b-512 */
static const uint64_t B2_IV[8] = {
    0x6a09e667f3bcc908ULL, 0xbb67ae8584caa73bULL, 0x3c6ef372fe94f82bULL, 0xa54ff53a5f1d36f1ULL,
    0x510e527fade682d1ULL, 0x9b05688c2b3e6c1fULL, 0x1f83d9abfb41bd6bULL, 0x5be0cd19137e2179ULL};
static const uint8_t B2_SIGMA[12][16] = {
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
    {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
    {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
    {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
    {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0},
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3}};
static uint64_t rotr64(uint64_t x, int r) { return (x >> r) | (x << (64 - r)); }
#define B2G(a, b, c, d, x, y) \
    v[a] = v[a] + v[b] + (x); v[d] = rotr64(v[d] ^ v[a], 32); v[c] = v[c] + v[d]; v[b] = rotr64(v[b] ^ v[c], 24); \
    v[a] = v[a] + v[b] + (y); v[d] = rotr64(v[d] ^ v[a], 16); v[c] = v[c] + v[d]; v[b] = rotr64(v[b] ^ v[c], 63);
/* single-call blake2b-512, no key, input <= 128 bytes is all the SRS derivation needs but any length works */
void oracle_blake2b512(const uint8_t *in, size_t len, uint8_t out[64]) {
    uint64_t h[8]; memcpy(h, B2_IV, sizeof h); h[0] ^= 0x01010000ULL ^ 64;
    size_t off = 0; uint64_t t = 0;
    for (;;) {
        uint8_t block[128]; memset(block, 0, 128);
        size_t take = len - off > 128 ? 128 : len - off;
        int last = (off + take == len);
        memcpy(block, in + off, take); off += take; t += take;
        uint64_t m[16], v[16]; memcpy(m, block, 128);
        for (int i = 0; i < 8; i++) { v[i] = h[i]; v[i + 8] = B2_IV[i]; }
        v[12] ^= t; if (last) v[14] = ~v[14];
        for (int r = 0; r < 12; r++) {
            const uint8_t *s = B2_SIGMA[r];
            B2G(0, 4, 8, 12, m[s[0]], m[s[1]]) B2G(1, 5, 9, 13, m[s[2]], m[s[3]])
            B2G(2, 6, 10, 14, m[s[4]], m[s[5]]) B2G(3, 7, 11, 15, m[s[6]], m[s[7]])
            B2G(0, 5, 10, 15, m[s[8]], m[s[9]]) B2G(1, 6, 11, 12, m[s[10]], m[s[11]])
            B2G(2, 7, 8, 13, m[s[12]], m[s[13]]) B2G(3, 4, 9, 14, m[s[14]], m[s[15]])
        }
        for (int i = 0; i < 8; i++) h[i] ^= v[i] ^ v[i + 8];
        if (last) break;
    }
    memcpy(out, h, 64);
}

/* ------------------------------------------------------- SRS derivation (SRS::create, groupmap) */
static void srs_hash_to_field(const field_ctx *F, const uint8_t *msg, size_t len, fe *out_mont) {
    uint8_t dg[64]; oracle_blake2b512(msg, len, dg);
    /* 248 bits, LSB-first inside each byte, as a big-endian bit string */
    fe v = {{0, 0, 0, 0}};
    for (int i = 0; i < 31; i++)
        for (int j = 0; j < 8; j++) {
            int bit = (dg[i] >> j) & 1;
            int pos = 247 - (i * 8 + j);
            if (bit) v.l[pos / 64] |= (uint64_t)1 << (pos % 64);
        }
    f_to_mont(F, out_mont, &v);
}
typedef struct { fe fu, inv3, sqrt_m3, sqrt_m3_minus1_half; int ready; } gm_params;
static gm_params GM[2];
static void gm_init(int field_id) {
    if (GM[field_id].ready) return;
    const field_ctx *F = &CTX[field_id]; gm_params *g = &GM[field_id];
    fe three = {{3, 0, 0, 0}}, six = {{6, 0, 0, 0}}, two = {{2, 0, 0, 0}}, t;
    f_to_mont(F, &g->fu, &six);
    f_to_mont(F, &three, &three); f_inv(F, &g->inv3, &three);
    f_neg(F, &t, &three); f_sqrt(F, &g->sqrt_m3, &t);
    f_to_mont(F, &two, &two); f_inv(F, &two, &two);
    f_sub(F, &t, &g->sqrt_m3, &F->r); f_mul(F, &g->sqrt_m3_minus1_half, &t, &two);
    g->ready = 1;
}
static void gm_to_group(int field_id, const fe *t_mont, fe *x_out, fe *y_out) {
    const field_ctx *F = &CTX[field_id]; const gm_params *g = &GM[field_id];
    fe t2, ainv, alpha, tmp, xs[3], t2fu, t2inv;
    f_sqr(F, &t2, t_mont);
    f_add(F, &t2fu, &t2, &g->fu);
    f_mul(F, &ainv, &t2fu, &t2);
    if (fe_is_zero(&ainv)) alpha = ainv; else f_inv(F, &alpha, &ainv);
    f_sqr(F, &tmp, &t2); f_mul(F, &tmp, &tmp, &alpha); f_mul(F, &tmp, &tmp, &g->sqrt_m3);
    f_sub(F, &xs[0], &g->sqrt_m3_minus1_half, &tmp);
    f_neg(F, &xs[1], &F->r); f_sub(F, &xs[1], &xs[1], &xs[0]);
    f_mul(F, &t2inv, &alpha, &t2fu);
    f_sqr(F, &tmp, &t2fu); f_mul(F, &tmp, &tmp, &t2inv); f_mul(F, &tmp, &tmp, &g->inv3);
    f_sub(F, &xs[2], &F->r, &tmp);
    for (int k = 0; k < 3; k++) {
        fe rhs, y; f_sqr(F, &rhs, &xs[k]); f_mul(F, &rhs, &rhs, &xs[k]); f_add(F, &rhs, &rhs, &F->five);
        if (f_sqrt(F, &y, &rhs)) { *x_out = xs[k]; *y_out = y; return; }
    }
    memset(x_out, 0, sizeof(fe)); memset(y_out, 0, sizeof(fe));
}
/* out64[i] = g[start+i] for i < count; if want_h, writes h to h64. */
void oracle_srs_derive(int field_id, uint32_t start, uint32_t count, uint8_t *out64, int want_h, uint8_t *h64) {
    ctx_init(); gm_init(field_id);
    const field_ctx *F = &CTX[field_id];
    for (uint32_t k = 0; k < count; k++) {
        uint32_t i = start + k; uint8_t msg[4] = {(uint8_t)(i >> 24), (uint8_t)(i >> 16), (uint8_t)(i >> 8), (uint8_t)i};
        fe t, x, y; srs_hash_to_field(F, msg, 4, &t); gm_to_group(field_id, &t, &x, &y);
        f_from_mont(F, &x, &x); f_from_mont(F, &y, &y);
        fe_to_bytes(out64 + 64 * k, &x); fe_to_bytes(out64 + 64 * k + 32, &y);
    }
    if (want_h) {
        uint8_t msg[12] = {'s', 'r', 's', '_', 'm', 'i', 's', 'c', 0, 0, 0, 0};
        fe t, x, y; srs_hash_to_field(F, msg, 12, &t); gm_to_group(field_id, &t, &x, &y);
        f_from_mont(F, &x, &x); f_from_mont(F, &y, &y);
        fe_to_bytes(h64, &x); fe_to_bytes(h64 + 32, &y);
    }
}
/* to_group on an arbitrary field element (used by the IPA check for U) */
void oracle_to_group(int field_id, const uint8_t t32[32], uint8_t out64[64]) {
    ctx_init(); gm_init(field_id);
    const field_ctx *F = &CTX[field_id];
    fe t, x, y; fe_from_bytes(&t, t32); f_to_mont(F, &t, &t); gm_to_group(field_id, &t, &x, &y);
    f_from_mont(F, &x, &x); f_from_mont(F, &y, &y);
    fe_to_bytes(out64, &x); fe_to_bytes(out64 + 32, &y);
}

/* ------------------------------------------------------- IPA scalar helpers (poly-commitment) */
/* scalar_field_id: the field the challenge lands in (0 = Fp, 1 = Fq). endo32 = endo_r canonical. */
void oracle_endo_to_field(int scalar_field_id, size_t n, const uint8_t *pre16, const uint8_t endo32[32], uint8_t *out32) {
    ctx_init();
    const field_ctx *F = &CTX[scalar_field_id];
    fe endo; fe_from_bytes(&endo, endo32); f_to_mont(F, &endo, &endo);
    fe two = {{2, 0, 0, 0}}, one_m = F->r, neg_one; f_to_mont(F, &two, &two); f_neg(F, &neg_one, &one_m);
    for (size_t k = 0; k < n; k++) {
        uint64_t lo, hi; memcpy(&lo, pre16 + 16 * k, 8); memcpy(&hi, pre16 + 16 * k + 8, 8);
        fe a = two, b = two;
        for (int i = 63; i >= 0; i--) {
            f_dbl(F, &a, &a); f_dbl(F, &b, &b);
            int b0 = 2 * i, b1 = 2 * i + 1;
            int r0 = (int)(((b0 < 64 ? lo >> b0 : hi >> (b0 - 64))) & 1);
            int r1 = (int)(((b1 < 64 ? lo >> b1 : hi >> (b1 - 64))) & 1);
            const fe *s = r0 ? &one_m : &neg_one;
            if (!r1) f_add(F, &b, &b, s); else f_add(F, &a, &a, s);
        }
        fe r; f_mul(F, &r, &a, &endo); f_add(F, &r, &r, &b); f_from_mont(F, &r, &r);
        fe_to_bytes(out32 + 32 * k, &r);
    }
}
/* s[i] = prod_j chals[k-1-j]^{bit_j(i)}, i < 2^k; chals canonical, out canonical */
void oracle_bpoly_coeffs(int scalar_field_id, int k, const uint8_t *chals32, uint8_t *out32) {
    ctx_init();
    const field_ctx *F = &CTX[scalar_field_id];
    size_t n = (size_t)1 << k;
    fe *s = (fe *)malloc(n * sizeof(fe)); fe *ch = (fe *)malloc(k * sizeof(fe));
    for (int i = 0; i < k; i++) { fe_from_bytes(&ch[i], chals32 + 32 * i); f_to_mont(F, &ch[i], &ch[i]); }
    s[0] = F->r;
    size_t pw = 1; int kk = 0;
    for (size_t i = 1; i < n; i++) {
        if (i == pw << 1) { pw <<= 1; kk++; }
        f_mul(F, &s[i], &s[i - pw], &ch[k - 1 - kk]);
    }
    for (size_t i = 0; i < n; i++) { fe t; f_from_mont(F, &t, &s[i]); fe_to_bytes(out32 + 32 * i, &t); }
    free(s); free(ch);
}

/* ------------------------------------------------------------ Poseidon (mina-poseidon kimchi) */
/* params: 9 MDS entries (row-major) then 55*3 round constants, canonical 32-byte LE. */
void oracle_poseidon_permute(int field_id, const uint8_t *params, size_t nperm, uint8_t *states96) {
    ctx_init();
    const field_ctx *F = &CTX[field_id];
    fe mds[9], rc[165];
    for (int i = 0; i < 9; i++) { fe_from_bytes(&mds[i], params + 32 * i); f_to_mont(F, &mds[i], &mds[i]); }
    for (int i = 0; i < 165; i++) { fe_from_bytes(&rc[i], params + 32 * (9 + i)); f_to_mont(F, &rc[i], &rc[i]); }
    for (size_t k = 0; k < nperm; k++) {
        fe st[3];
        for (int i = 0; i < 3; i++) { fe_from_bytes(&st[i], states96 + 96 * k + 32 * i); f_to_mont(F, &st[i], &st[i]); }
        for (int r = 0; r < 55; r++) {
            fe sb[3];
            for (int i = 0; i < 3; i++) {
                fe x2, x4, x6; f_sqr(F, &x2, &st[i]); f_sqr(F, &x4, &x2); f_mul(F, &x6, &x4, &x2); f_mul(F, &sb[i], &x6, &st[i]);
            }
            for (int i = 0; i < 3; i++) {
                fe acc, t; f_mul(F, &acc, &mds[3 * i], &sb[0]);
                f_mul(F, &t, &mds[3 * i + 1], &sb[1]); f_add(F, &acc, &acc, &t);
                f_mul(F, &t, &mds[3 * i + 2], &sb[2]); f_add(F, &acc, &acc, &t);
                f_add(F, &st[i], &acc, &rc[3 * r + i]);
            }
        }
        for (int i = 0; i < 3; i++) { fe t; f_from_mont(F, &t, &st[i]); fe_to_bytes(states96 + 96 * k + 32 * i, &t); }
    }
}

/* ------------------------------------------------------------------ self-timing (cpu_baseline honesty) */
/* ns per Montgomery multiplication on one core: a dependent chain of `iters` products (bench.py prints it
 * next to arkworks' expected 20-25 ns with the x86 asm backend, so nobody mistakes a slow checker for a
 * slow reference). */
#include <time.h>
double oracle_modmul_ns(int field_id, int iters) {
    ctx_init();
    const field_ctx *F = &CTX[field_id];
    fe a = F->r2, b = F->five;
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int i = 0; i < iters; i++) { f_mul(F, &a, &a, &b); f_mul(F, &b, &b, &a); }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    volatile uint64_t sink = a.l[0] ^ b.l[0]; (void)sink;
    return ((t1.tv_sec - t0.tv_sec) * 1e9 + (t1.tv_nsec - t0.tv_nsec)) / (2.0 * iters);
}

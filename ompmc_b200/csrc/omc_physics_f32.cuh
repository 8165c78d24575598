// omc_physics_f32.cuh -- single-precision elastic-scattering samplers for the wavefront kernels.
//
// The condensed-history step spends most of its instructions in msdist()/mscat()/spinRejection()
// (src/ompmc.c:3097-3199, 3606-3976): angle sampling, not bookkeeping.  Angles do not need 53-bit
// arithmetic (the reference's own tables are 16-bit (spin) / text (msnew.data) precision), whereas fp64
// logs/exps/divisions/square roots are software sequences on the GPU and double every register.  These
// versions compute the SAME algorithm in fp32 with the hardware MUFU approximations; positions, energies,
// path lengths and all CSDA bookkeeping stay fp64 in the callers.  Cancellation-prone forms are avoided:
// samplers work with xi = 1 - cos(theta) instead of cos(theta).
// The lock-step kernel never uses this header (it stays bit-faithful fp64).
#pragma once
#include "omc_physics.cuh"

namespace omc {

constexpr float RMf = (float)OMC_RM;

__device__ __forceinline__ float nextf(Rng &g) {           // 24-bit lattice in [0,1), like RANMAR's
    if (g.pos >= 4u) g.refill();
    const uint32_t w = g.pos == 0u ? g.b0 : (g.pos == 1u ? g.b1 : (g.pos == 2u ? g.b2 : g.b3));
    g.pos += 1;
    return (float)(w >> 8) * (1.0f / 16777216.0f);
}
__device__ __forceinline__ float fdiv(float a, float b) { return __fdividef(a, b); }
__device__ __forceinline__ float frcp(float a) { return __frcp_rn(a); }

// selectAzimuthalAngle(), src/ompmc.c:101-122
static __device__ __noinline__ void azimuth_f(Rng &g, float &cphi, float &sphi) {
    float x, x2, y, y2, r2;
    do {
        x = nextf(g); x = 2.0f * x - 1.0f; x2 = x * x;
        y = nextf(g); y2 = y * y;
        r2 = x2 + y2;
    } while (r2 > 1.0f || r2 == 0.0f);
    r2 = frcp(r2);
    cphi = (x2 - y2) * r2;
    sphi = 2.0f * x * y * r2;
}

// spinRejection(), src/ompmc.c:3097-3168, argument omc = 1 - cos(theta)
static __device__ __noinline__ float spin_rejection_f(const DevProblem &P, Rng &g, int imed, int qel, float elke, float beta2, float q1,
                                                      float omc_, bool &spin_index, bool is_single, SpinState &sr) {
    if (spin_index) {
        spin_index = false;
        float ai;
        const float b2min = (float)P.b2spin_min, espml = (float)P.espml;
        if (beta2 >= b2min) {
            ai = (beta2 - b2min) * (float)P.dbeta2i;
            sr.i = (int)ai; ai -= (float)sr.i; sr.i += 16;
            if (sr.i > 30) { sr.i = 30; ai = 1.0f; }               // beta2 -> 1 in fp32
        } else if (elke > espml) {
            ai = (elke - espml) * (float)P.dleneri;
            sr.i = (int)ai; ai -= (float)sr.i;
        } else {
            sr.i = 0; ai = -1.0f;
        }
        float r = nextf(g);
        if (r < ai) sr.i += 1;
        if (is_single) {
            sr.j = 0;
        } else {
            float qq1 = 2.0f * q1;
            qq1 = fdiv(qq1, 1.0f + qq1);
            float aj = qq1 * (float)P.dqq1i;
            sr.j = (int)aj;
            if (sr.j >= 15) {
                sr.j = 15;
            } else {
                aj -= (float)sr.j;
                r = nextf(g);
                if (r < aj) sr.j += 1;
            }
        }
    }
    const float xi = sqrtf(0.5f * omc_);
    float ak = xi * 31.0f;
    int k = (int)ak;
    if (k > 30) k = 30;
    ak -= (float)k;
    const float *row = P.spin_rej_f + (((size_t)(imed * 2 + qel) * OMC_SPIN_NE + sr.i) * OMC_SPIN_NQ + sr.j) * OMC_SPIN_NU;
    return (1.0f - ak) * __ldg(row + k) + ak * __ldg(row + k + 1);
}

// sscat(), src/ompmc.c:3170-3199
static __device__ __noinline__ void sscat_f(const DevProblem &P, Rng &g, int imed, int qel, float chia2, float elke, float beta2,
                                            float &cost, float &sint) {
    bool spin_index = true;
    SpinState sr;
    float xi, rejf, r;
    do {
        xi = nextf(g);
        xi = fdiv(2.0f * chia2 * xi, 1.0f - xi + chia2);
        rejf = spin_rejection_f(P, g, imed, qel, elke, beta2, 0.0f, xi, spin_index, true, sr);
        r = nextf(g);
    } while (r > rejf);
    cost = 1.0f - xi;
    sint = sqrtf(xi * (2.0f - xi));
}

// mscat(), src/ompmc.c:3606-3785 (Q1: u takes tabulated values only, its interpolation draw is still made)
static __device__ __noinline__ void mscat_f(const DevProblem &P, Rng &g, int imed, int qel, bool &spin_index, bool &find_index,
                                            float elke, float beta2, float q1, float lambda, float chia2, float &cost, float &sint,
                                            MsState &ms, SpinState &sr) {
    float xi, rejf, r;
    const float explambda = __expf(-lambda);
    if (lambda <= 13.8f) {
        const float sprob = nextf(g);
        if (sprob < explambda) { cost = 1.0f; sint = 0.0f; return; }
        float wsum = (1.0f + lambda) * explambda;
        if (sprob < wsum) {
            do {
                xi = nextf(g);
                xi = fdiv(2.0f * chia2 * xi, 1.0f - xi + chia2);
                rejf = spin_rejection_f(P, g, imed, qel, elke, beta2, q1, xi, spin_index, false, sr);
                r = nextf(g);
            } while (r > rejf);
            cost = 1.0f - xi;
            sint = sqrtf(xi * (2.0f - xi));
            return;
        }
        if (lambda <= 1.0f) {
            int icount = 0;
            float wprob = explambda, sinz, cosz, phi;
            wsum = explambda;
            cost = 1.0f; sint = 0.0f;
            do {
                icount += 1;
                if (icount > 20) break;
                wprob = wprob * lambda / (float)icount;
                wsum = wsum + wprob;
                do {
                    xi = nextf(g);
                    xi = fdiv(2.0f * chia2 * xi, 1.0f - xi + chia2);
                    rejf = spin_rejection_f(P, g, imed, qel, elke, beta2, q1, xi, spin_index, false, sr);
                    r = nextf(g);
                } while (r > rejf);
                cosz = 1.0f - xi;
                sinz = xi * (2.0f - xi);
                if (sinz > 1.0E-20f) {
                    sinz = sqrtf(sinz);
                    xi = nextf(g);
                    phi = xi * 6.2831853f;
                    cost = cost * cosz - sint * sinz * __cosf(phi);
                    sint = sqrtf(fmaxf(0.0f, (1.0f - cost) * (1.0f + cost)));
                }
            } while (wsum <= sprob);
            return;
        }
    }
    if (lambda <= 1.0E5f) {
        const float llmbda = __logf(lambda);
        if (find_index) {
            float ai = llmbda * (float)P.dllambi;
            ms.i = (int)ai; ai -= (float)ms.i;
            xi = nextf(g);
            if (xi < ai) ms.i += 1;
            if (ms.i > 63) ms.i = 63;
            if (q1 < 1.0E-3f) {
                ms.j = 0;
            } else if (q1 < 0.5f) {
                float aj = q1 * (float)P.dqmsi;
                ms.j = (int)aj; aj -= (float)ms.j;
                xi = nextf(g);
                if (xi < aj) ms.j += 1;
            } else {
                ms.j = 7;
            }
            float om;
            if (llmbda < 2.2299f)
                om = chia2 * (lambda + 4.0f) * (1.347006f + llmbda * (0.209364f - llmbda * (0.45525f - llmbda * (0.50142f - 0.081234f * llmbda))));
            else
                om = chia2 * (lambda + 4.0f) * (-2.77164f + llmbda * (2.94874f - llmbda * (0.1535754f - llmbda * 0.00552888f)));
            ms.omega2 = (double)om;
            find_index = false;
        }
        const float omega2 = (float)ms.omega2;
        const MsEntryF *tab = P.ms_f + (ms.i * OMC_MS_NQ + ms.j) * OMC_MS_NU;
        do {
            xi = nextf(g);
            float ak = xi * 31.0f;
            int k = (int)ak;
            ak -= (float)k;
            const float4 t0 = __ldg(reinterpret_cast<const float4 *>(tab + k));   // {ums, wms, ims, fms}
            if (ak > t0.y) k = __float_as_int(t0.z);
            const float u = __ldg(&tab[k].ums);
            xi = nextf(g);                                     // Q1: dead interpolation draw
            xi = fdiv(omega2 * u, 1.0f + 0.5f * omega2 - u);
            if (xi > 1.99999f) xi = 1.99999f;
            rejf = spin_rejection_f(P, g, imed, qel, elke, beta2, q1, xi, spin_index, false, sr);
            r = nextf(g);
        } while (r > rejf);
        cost = 1.0f - xi;
        sint = sqrtf(xi * (2.0f - xi));
    }
}

// msdist(), src/ompmc.c:3787-3976 (PRESTA-II) in fp32; end point and direction are returned in fp64
static __device__ __noinline__ double msdist_f(const DevProblem &P, Rng &g, const Part &p, int imed, int qel, double rhof_d, double de_d,
                                               double tustep_d, double eke_d, double &xf, double &yf, double &zf, double &uf,
                                               double &vf, double &wf) {
    const MedRec &M = P.med[imed];
    MsState ms;
    SpinState sr;
    const float rhof = (float)rhof_d, de = (float)de_d, tustep = (float)tustep_d, eke = (float)eke_d;
    const float xcc = (float)M.xcc, blcc = (float)M.blcc;
    float e = eke - 0.5f * de;
    const float tau = e * (1.0f / RMf), tau2 = tau * tau;
    const float epsilon = fdiv(de, eke), epsilonp = fdiv(de, e);
    e *= (1.0f - (epsilonp * epsilonp) * fdiv(6.0f + 10.0f * tau + 5.0f * tau2, 24.0f * tau2 + 72.0f * tau + 48.0f));
    const float p2 = e * (e + 2.0f * RMf);
    const float beta2 = fdiv(p2, p2 + (RMf * RMf));
    float chia2 = fdiv(xcc, 4.0f * p2 * blcc);
    float lambda = fdiv(0.5f * tustep * rhof * blcc, beta2);
    const float t12 = fdiv(epsilonp, (tau + 1.0f) * (tau + 2.0f));
    const float temp2 = 0.166666f * (4.0f + tau * (6.0f + tau * (7.0f + tau * (4.0f + tau)))) * t12 * t12;
    lambda *= (1.0f - temp2);
    float elke = __logf(e);
    int lelke = (int)(elke * (float)M.eke1 + (float)M.eke0) - 1;
    if (lelke < 0) { lelke = 0; elke = (float)((1.0 - M.eke0) / M.eke1); }
    const ElecBin *B = P.ebin + (size_t)qel * P.nmed * MXEKE + imed * MXEKE + lelke;
    const double2 c_eta = ldg2(&B->eta1), c_q1 = ldg2(&B->q1c1), c_q2 = ldg2(&B->q2c1), c_bl = ldg2(&B->blcce1);
    const float etap = elke * (float)c_eta.x + (float)c_eta.y;
    const float xi_corr = elke * (float)c_q1.x + (float)c_q1.y;
    float gamma = elke * (float)c_q2.x + (float)c_q2.y;
    const float ms_corr = elke * (float)c_bl.x + (float)c_bl.y;
    chia2 *= etap;
    lambda = fdiv(lambda, etap * (1.0f + chia2));
    lambda *= ms_corr;
    const float chilog = __logf(1.0f + frcp(chia2));
    const float q1 = 2.0f * chia2 * (chilog * (1.0f + chia2) - 1.0f);
    gamma = fdiv(6.0f * chia2 * (1.0f + chia2) * (chilog * (1.0f + 2.0f * chia2) - 2.0f), q1) * gamma;
    float xi = q1 * lambda;
    bool find_index = true, spin_index = true;
    float w1 = 1.0f, sint1 = 0.0f, cphi1, sphi1, w2 = 1.0f, sint2 = 0.0f, cphi2, sphi2;
    g.align();
    mscat_f(P, g, imed, qel, spin_index, find_index, elke, beta2, xi, lambda, chia2, w1, sint1, ms, sr);
    g.align();
    azimuth_f(g, cphi1, sphi1);
    g.align();
    mscat_f(P, g, imed, qel, spin_index, find_index, elke, beta2, xi, lambda, chia2, w2, sint2, ms, sr);
    g.align();
    azimuth_f(g, cphi2, sphi2);
    const float u2 = sint2 * cphi2, v2 = sint2 * sphi2;
    float u2p = w1 * u2 + sint1 * w2;
    float us = u2p * cphi1 - v2 * sphi1, vs = u2p * sphi1 + v2 * cphi1, ws = w1 * w2 - sint1 * u2;
    xi *= 2.0f * xi_corr;
    const float eta = nextf(g);
    const float eta1 = 0.5f * (1.0f - eta);
    float delta = 0.9082483f - (0.1020621f - 0.0263747f * gamma) * xi;
    float temp1 = 2.0f + tau;
    float temp = fdiv(2.0f + tau * temp1, (tau + 1.0f) * temp1);
    const float c1 = chilog * (1.0f + chia2) - 1.0f, c2 = chilog * (1.0f + 2.0f * chia2) - 2.0f;
    temp -= fdiv(tau + 1.0f, (tau + 2.0f) * c1);
    temp *= epsilonp;
    temp1 = 1.0f - temp;
    delta += 0.40824829f * (fdiv(epsilon * (tau + 1.0f), (tau + 2.0f) * c1 * c2) - 0.25f * (temp * temp));
    const float b = eta * delta, cc = eta * (1.0f - delta);
    const float w1v2 = w1 * v2;
    float ut = b * sint1 * cphi1 + cc * (cphi1 * u2 - sphi1 * w1v2) + eta1 * us * temp1;
    float vt = b * sint1 * sphi1 + cc * (sphi1 * u2 + cphi1 * w1v2) + eta1 * vs * temp1;
    float wt = eta1 * (1.0f + temp) + b * w1 + cc * w2 + eta1 * ws * temp1;
    const float ustep = tustep * sqrtf(ut * ut + vt * vt + wt * wt);
    const float u0 = (float)p.u, v0 = (float)p.v, w0 = (float)p.w;
    const float sint02 = u0 * u0 + v0 * v0;
    if (sint02 > 1.0E-20f) {
        const float sint0i = rsqrtf(sint02), sint0 = sint02 * sint0i;
        const float cphi0 = sint0i * u0, sphi0 = sint0i * v0;
        u2p = w0 * us + sint0 * ws;
        ws = w0 * ws - sint0 * us;
        us = u2p * cphi0 - vs * sphi0;
        vs = u2p * sphi0 + vs * cphi0;
        u2p = w0 * ut + sint0 * wt;
        wt = w0 * wt - sint0 * ut;
        ut = u2p * cphi0 - vt * sphi0;
        vt = u2p * sphi0 + vt * cphi0;
    } else {
        wt = w0 * wt; ws = w0 * ws;
    }
    // keep the direction a unit vector in fp32 (the fp64 path relies on 1e-16 rounding instead)
    const float nrm = rsqrtf(us * us + vs * vs + ws * ws);
    xf = p.x + tustep_d * (double)ut; yf = p.y + tustep_d * (double)vt; zf = p.z + tustep_d * (double)wt;
    uf = (double)(us * nrm); vf = (double)(vs * nrm); wf = (double)(ws * nrm);
    return (double)ustep;
}

// ---------------------------------------------------------------------------------------------
// Block-draw version of the condensed-history step for edo_kernel<CH>.  ncu on the versions above: the Rng object
// is passed by reference through __noinline__ samplers, so it lives in LOCAL memory (LDL + STL = 12.6 % of the
// executed instructions), its word-select chain is another 11 %, and the Philox block function runs at 14 of 32
// lanes because lanes reach their refills at different draw sites (8 partial executions per warp and step).
// Here every random number comes from a whole Philox block drawn at a fixed program point by all lanes together
// (Rng::block(), register-only); words that a lane does not need are dropped.  Same algorithm and distributions as
// msdist()/mscat()/spinRejection() (src/ompmc.c:3097-3168, 3606-3976); statistically neutral differences: the
// four table-index rounding decisions (mscat :3702-3724, spinRejection :3104-3150) are always made, up front, from
// 16-bit halves of two words; the dead interpolation draw of mscat (Q1) is not generated.
//   block G0 = {sprob pass 0, sprob pass 1, eta of msdist, rfict of the caller's sigma-ratio test}
//   block G1 = {ms i|j rounding, spin i|j rounding, first (u, r) pair}
//   block G2 = {second (u, r) pair, azimuth 1, azimuth 2}          -> three blocks for the common step (was ~5)
//   further (u, r) pairs, only after a rejection, two per block
// The azimuths are (cos, sin)(2 pi u) instead of selectAzimuthalAngle()'s box rejection (:101-122): same uniform distribution,
// no retry loop (it ran at 11 of 32 lanes and cost a block per round).  Measured on B200 (prostate6mv, 4e7-history call):
// box-rejection azimuths 1.106e8, sincos 1.155e8, packed plan + one trial loop for both polar angles 1.205e8 histories/s.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float u24(uint32_t w) { return (float)(w >> 8) * (1.0f / 16777216.0f); }
__device__ __forceinline__ float u16lo(uint32_t w) { return (float)(w & 0xffffu) * (1.0f / 65536.0f); }
__device__ __forceinline__ float u16hi(uint32_t w) { return (float)(w >> 16) * (1.0f / 65536.0f); }

// spinRejection() index part, :3104-3150: the (i, j) row of the Mott table for this step
__device__ __forceinline__ const float *spin_row(const DevProblem &P, int imed, int qel, float elke, float beta2, float q1, float ri, float rj,
                                                 bool is_single = false) {
    int i;
    float ai;
    const float b2min = (float)P.b2spin_min, espml = (float)P.espml;
    if (beta2 >= b2min) {
        ai = (beta2 - b2min) * (float)P.dbeta2i;
        i = (int)ai; ai -= (float)i; i += 16;
        if (i > 30) { i = 30; ai = 1.0f; }                     // beta2 -> 1 in fp32
    } else if (elke > espml) {
        ai = (elke - espml) * (float)P.dleneri;
        i = (int)ai; ai -= (float)i;
    } else {
        i = 0; ai = -1.0f;
    }
    if (ri < ai) i += 1;
    int j = 0;
    if (!is_single) {
        float qq1 = 2.0f * q1;
        qq1 = fdiv(qq1, 1.0f + qq1);
        float aj = qq1 * (float)P.dqq1i;
        j = (int)aj;
        if (j >= 15) {
            j = 15;
        } else {
            aj -= (float)j;
            if (rj < aj) j += 1;
        }
    }
    return P.spin_rej_f + (((size_t)(imed * 2 + qel) * OMC_SPIN_NE + i) * OMC_SPIN_NQ + j) * OMC_SPIN_NU;
}
// spinRejection() value part, :3152-3167, argument omc = 1 - cos(theta)
__device__ __forceinline__ float spin_rej_row(const float *row, float omc_) {
    const float xi = sqrtf(0.5f * omc_);
    float ak = xi * 31.0f;
    int k = (int)ak;
    if (k > 30) k = 30;
    ak -= (float)k;
    return (1.0f - ak) * __ldg(row + k) + ak * __ldg(row + k + 1);
}

// source of (u, r) word pairs: the pair left over in G1 first, then two pairs per fresh block
struct PairSrc {
    uint4 b;
    uint32_t a0, a1;
    int have, nb;
    __device__ __forceinline__ void next(Rng &g, uint32_t &u, uint32_t &r) {
        if (have) { u = a0; r = a1; have = 0; return; }
        if (nb >= 2) { b = g.block(); nb = 0; }
        u = nb ? b.z : b.x; r = nb ? b.w : b.y;
        nb += 1;
    }
};

// sscat() + the azimuth of uphi21(), src/ompmc.c:3170-3199 / :101-122, for the boundary-crossing step: (u, r)
// pairs two per block; `wi` = the word whose low half rounds the spin-table energy index
__device__ __forceinline__ void sscat_b(const DevProblem &P, Rng &g, int imed, int qel, float chia2, float elke, float beta2, uint32_t wi,
                                        float &cost, float &sint, float &cphi, float &sphi) {
    const float *row = spin_row(P, imed, qel, elke, beta2, 0.0f, u16lo(wi), 0.0f, true);
    PairSrc ps;
    ps.a0 = ps.a1 = 0u; ps.have = 0; ps.nb = 2;
    float x;
    for (;;) {
        uint32_t wu, wr;
        ps.next(g, wu, wr);
        const float u = u24(wu);
        x = fdiv(2.0f * chia2 * u, 1.0f - u + chia2);
        if (!(u24(wr) > spin_rej_row(row, x))) break;
    }
    cost = 1.0f - x;
    sint = sqrtf(x * (2.0f - x));
    {   // selectAzimuthalAngle() :101-122 draws a uniform azimuth by box rejection; (cos, sin)(2 pi u) is the same distribution
        uint32_t wu, wr;
        ps.next(g, wu, wr);
        __sincosf(6.2831853f * u24(wu), &sphi, &cphi);
    }
}

__device__ __forceinline__ double msdist_b(const DevProblem &P, Rng &g, const Part &p, int imed, int qel, double rhof_d, double de_d,
                                           double tustep_d, double eke_d, double &xf, double &yf, double &zf, double &uf, double &vf,
                                           double &wf, uint32_t &w_rfict) {
    const MedRec &M = P.med[imed];
    const float rhof = (float)rhof_d, de = (float)de_d, tustep = (float)tustep_d, eke = (float)eke_d;
    const float xcc = (float)M.xcc, blcc = (float)M.blcc;
    float e = eke - 0.5f * de;
    const float tau = e * (1.0f / RMf), tau2 = tau * tau;
    const float epsilon = fdiv(de, eke), epsilonp = fdiv(de, e);
    e *= (1.0f - (epsilonp * epsilonp) * fdiv(6.0f + 10.0f * tau + 5.0f * tau2, 24.0f * tau2 + 72.0f * tau + 48.0f));
    const float p2 = e * (e + 2.0f * RMf);
    const float beta2 = fdiv(p2, p2 + (RMf * RMf));
    float chia2 = fdiv(xcc, 4.0f * p2 * blcc);
    float lambda = fdiv(0.5f * tustep * rhof * blcc, beta2);
    const float t12 = fdiv(epsilonp, (tau + 1.0f) * (tau + 2.0f));
    const float temp2 = 0.166666f * (4.0f + tau * (6.0f + tau * (7.0f + tau * (4.0f + tau)))) * t12 * t12;
    lambda *= (1.0f - temp2);
    float elke = __logf(e);
    int lelke = (int)(elke * (float)M.eke1 + (float)M.eke0) - 1;
    if (lelke < 0) { lelke = 0; elke = (float)((1.0 - M.eke0) / M.eke1); }
    const ElecBin *B = P.ebin + (size_t)qel * P.nmed * MXEKE + imed * MXEKE + lelke;
    const double2 c_eta = ldg2(&B->eta1), c_q1 = ldg2(&B->q1c1), c_q2 = ldg2(&B->q2c1), c_bl = ldg2(&B->blcce1);
    const float etap = elke * (float)c_eta.x + (float)c_eta.y;
    const float xi_corr = elke * (float)c_q1.x + (float)c_q1.y;
    float gamma = elke * (float)c_q2.x + (float)c_q2.y;
    const float ms_corr = elke * (float)c_bl.x + (float)c_bl.y;
    chia2 *= etap;
    lambda = fdiv(lambda, etap * (1.0f + chia2));
    lambda *= ms_corr;
    const float chilog = __logf(1.0f + frcp(chia2));
    const float q1 = 2.0f * chia2 * (chilog * (1.0f + chia2) - 1.0f);
    gamma = fdiv(6.0f * chia2 * (1.0f + chia2) * (chilog * (1.0f + 2.0f * chia2) - 2.0f), q1) * gamma;
    float xi = q1 * lambda;

    const uint4 g0 = g.block();
    const uint4 g1 = g.block();
    w_rfict = g0.w;
    // table rows of this step: mscat :3694-3735 (find_index), spinRejection :3104-3150 (spin_index)
    const float explambda = __expf(-lambda);
    const float llmbda = __logf(lambda);
    int mi, mj;
    {
        float ai = llmbda * (float)P.dllambi;
        mi = (int)ai; ai -= (float)mi;
        if (u16lo(g1.x) < ai) mi += 1;
        mi = max(0, min(mi, 63));
        if (xi < 1.0E-3f) {
            mj = 0;
        } else if (xi < 0.5f) {
            float aj = xi * (float)P.dqmsi;
            mj = (int)aj; aj -= (float)mj;
            if (u16hi(g1.x) < aj) mj += 1;
        } else {
            mj = 7;
        }
    }
    float omega2;
    if (llmbda < 2.2299f)
        omega2 = chia2 * (lambda + 4.0f) * (1.347006f + llmbda * (0.209364f - llmbda * (0.45525f - llmbda * (0.50142f - 0.081234f * llmbda))));
    else
        omega2 = chia2 * (lambda + 4.0f) * (-2.77164f + llmbda * (2.94874f - llmbda * (0.1535754f - llmbda * 0.00552888f)));
    const MsEntryF *tab = P.ms_f + (mi * OMC_MS_NQ + mj) * OMC_MS_NU;
    const float *row = spin_row(P, imed, qel, elke, beta2, xi, u16lo(g1.y), u16hi(g1.y));

    // Packed draw plan (see the header comment of this section): G2 = {second (u, r) pair, azimuth 1, azimuth 2} is always drawn,
    // so a step whose two polar angles are both accepted at their first trial -- the common case -- costs three blocks in all.
    // The two polar angles are sampled by ONE loop over trials (trial t serves whichever pass the lane is at), not by one
    // rejection loop per pass: a warp iterates max_lanes(trials_0 + trials_1) times instead of max(trials_0) + max(trials_1),
    // and since every lane in the loop consumes exactly one pair per iteration, the lanes that need a fresh block need it in
    // the same iteration.
    const uint4 g2 = g.block();
    float w1 = 1.0f, sint1 = 0.0f, w2 = 1.0f, sint2 = 0.0f;
    int reg0 = 2, reg1 = 2;
#pragma unroll
    for (int pass = 0; pass < 2; pass++) {
        // regime of mscat(): 0 no scattering (or Q7), 1 single scattering, 2 table, 3 plural scattering (lambda <= 1)
        const float sprob = u24(pass ? g0.y : g0.x);
        int r = 2;
        if (lambda <= 13.8f) {
            if (sprob < explambda) r = 0;
            else if (sprob < (1.0f + lambda) * explambda) r = 1;
            else if (lambda <= 1.0f) r = 3;
        } else if (!(lambda <= 1.0E5f)) {
            r = 0;
        }
        if (pass == 0) reg0 = r; else reg1 = r;
    }
    if (reg0 == 3 || reg1 == 3) {                          // :3652-3682, rare in a condensed-history step: its own blocks
#pragma unroll 1
        for (int pass = 0; pass < 2; pass++) {
            if ((pass ? reg1 : reg0) != 3) continue;
            const float sprob = u24(pass ? g0.y : g0.x);
            float c = 1.0f, sv = 0.0f;
            int icount = 0;
            float wprob = explambda, wsum = explambda;
            do {
                icount += 1;
                if (icount > 20) break;
                wprob = wprob * lambda / (float)icount;
                wsum = wsum + wprob;
                float x;
                uint4 bb;
                for (;;) {
                    bb = g.block();                            // {u, r, phi, -}
                    const float u = u24(bb.x);
                    x = fdiv(2.0f * chia2 * u, 1.0f - u + chia2);
                    if (!(u24(bb.y) > spin_rej_row(row, x))) break;
                }
                const float cosz = 1.0f - x;
                float sinz = x * (2.0f - x);
                if (sinz > 1.0E-20f) {
                    sinz = sqrtf(sinz);
                    const float phi = u24(bb.z) * 6.2831853f;
                    c = c * cosz - sv * sinz * __cosf(phi);
                    sv = sqrtf(fmaxf(0.0f, (1.0f - c) * (1.0f + c)));
                }
            } while (wsum <= sprob);
            if (pass == 0) { w1 = c; sint1 = sv; } else { w2 = c; sint2 = sv; }
            if (pass == 0) reg0 = 0; else reg1 = 0;
        }
    }
    {
        int pass = (reg0 != 0) ? 0 : ((reg1 != 0) ? 1 : 2);
        uint4 bx = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll 1
        for (int t = 0; pass < 2; t++) {
            uint32_t wu, wr;
            if (t == 0) { wu = g1.z; wr = g1.w; }
            else if (t == 1) { wu = g2.x; wr = g2.y; }
            else {
                if (!(t & 1)) bx = g.block();
                wu = (t & 1) ? bx.z : bx.x; wr = (t & 1) ? bx.w : bx.y;
            }
            const float u = u24(wu);
            float x;
            if ((pass ? reg1 : reg0) == 1) {
                x = fdiv(2.0f * chia2 * u, 1.0f - u + chia2);
            } else {
                float ak = u * 31.0f;
                int k = (int)ak;
                ak -= (float)k;
                const float4 t0 = __ldg(reinterpret_cast<const float4 *>(tab + k));   // {ums, wms, ims, fms}
                if (ak > t0.y) k = __float_as_int(t0.z);
                const float um = __ldg(&tab[k].ums);
                x = fdiv(omega2 * um, 1.0f + 0.5f * omega2 - um);
                if (x > 1.99999f) x = 1.99999f;
            }
            if (!(u24(wr) > spin_rej_row(row, x))) {
                const float c = 1.0f - x, sv = sqrtf(x * (2.0f - x));
                if (pass == 0) { w1 = c; sint1 = sv; pass = (reg1 != 0) ? 1 : 2; }
                else { w2 = c; sint2 = sv; pass = 2; }
            }
        }
    }
    // both azimuths: selectAzimuthalAngle() :101-122 draws a uniform azimuth by box rejection; (cos, sin)(2 pi u) is the same
    // distribution without the loop
    float cphi1, sphi1, cphi2, sphi2;
    __sincosf(6.2831853f * u24(g2.z), &sphi1, &cphi1);
    __sincosf(6.2831853f * u24(g2.w), &sphi2, &cphi2);
    const float u2 = sint2 * cphi2, v2 = sint2 * sphi2;
    float u2p = w1 * u2 + sint1 * w2;
    float us = u2p * cphi1 - v2 * sphi1, vs = u2p * sphi1 + v2 * cphi1, ws = w1 * w2 - sint1 * u2;
    xi *= 2.0f * xi_corr;
    const float eta = u24(g0.z);
    const float eta1 = 0.5f * (1.0f - eta);
    float delta = 0.9082483f - (0.1020621f - 0.0263747f * gamma) * xi;
    float temp1 = 2.0f + tau;
    float temp = fdiv(2.0f + tau * temp1, (tau + 1.0f) * temp1);
    const float c1 = chilog * (1.0f + chia2) - 1.0f, c2 = chilog * (1.0f + 2.0f * chia2) - 2.0f;
    temp -= fdiv(tau + 1.0f, (tau + 2.0f) * c1);
    temp *= epsilonp;
    temp1 = 1.0f - temp;
    delta += 0.40824829f * (fdiv(epsilon * (tau + 1.0f), (tau + 2.0f) * c1 * c2) - 0.25f * (temp * temp));
    const float b = eta * delta, cc = eta * (1.0f - delta);
    const float w1v2 = w1 * v2;
    float ut = b * sint1 * cphi1 + cc * (cphi1 * u2 - sphi1 * w1v2) + eta1 * us * temp1;
    float vt = b * sint1 * sphi1 + cc * (sphi1 * u2 + cphi1 * w1v2) + eta1 * vs * temp1;
    float wt = eta1 * (1.0f + temp) + b * w1 + cc * w2 + eta1 * ws * temp1;
    const float ustep = tustep * sqrtf(ut * ut + vt * vt + wt * wt);
    const float u0 = (float)p.u, v0 = (float)p.v, w0 = (float)p.w;
    const float sint02 = u0 * u0 + v0 * v0;
    if (sint02 > 1.0E-20f) {
        const float sint0i = rsqrtf(sint02), sint0 = sint02 * sint0i;
        const float cphi0 = sint0i * u0, sphi0 = sint0i * v0;
        u2p = w0 * us + sint0 * ws;
        ws = w0 * ws - sint0 * us;
        us = u2p * cphi0 - vs * sphi0;
        vs = u2p * sphi0 + vs * cphi0;
        u2p = w0 * ut + sint0 * wt;
        wt = w0 * wt - sint0 * ut;
        ut = u2p * cphi0 - vt * sphi0;
        vt = u2p * sphi0 + vt * cphi0;
    } else {
        wt = w0 * wt; ws = w0 * ws;
    }
    const float nrm = rsqrtf(us * us + vs * vs + ws * ws);
    xf = p.x + tustep_d * (double)ut; yf = p.y + tustep_d * (double)vt; zf = p.z + tustep_d * (double)wt;
    uf = (double)(us * nrm); vf = (double)(vs * nrm); wf = (double)(ws * nrm);
    return (double)ustep;
}

// ---------------------------------------------------------------------------------------------
// Block-draw versions of the two interactions that make up 95 % of all discrete interactions at 6 MV (Compton 3.3
// and Moller 4.9 per history): same fp64 sampling code as compton() / moller() in omc_physics.cuh
// (src/ompmc.c:1670-1783, 4359-4435), random numbers from whole Philox blocks so that the generator state stays in
// registers (one block per Klein-Nishina try, two Moller tries or two azimuth tries per block).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double u32d(uint32_t w) { return (double)w * (1.0 / 4294967296.0); }

// selectAzimuthalAngle(), src/ompmc.c:101-122
__device__ __forceinline__ void azimuth_blk(Rng &g, double &cphi, double &sphi) {
    for (;;) {
        const uint4 b = g.block();
        double x = 2.0 * u32d(b.x) - 1.0, y = u32d(b.y), x2 = x * x, y2 = y * y, r2 = x2 + y2;
        if (r2 > 1.0 || r2 == 0.0) {
            x = 2.0 * u32d(b.z) - 1.0; y = u32d(b.w); x2 = x * x; y2 = y * y; r2 = x2 + y2;
            if (r2 > 1.0 || r2 == 0.0) continue;
        }
        r2 = 1.0 / r2;
        cphi = (x2 - y2) * r2;
        sphi = 2.0 * x * y * r2;
        return;
    }
}

__device__ __forceinline__ void compton_b(Rng &g, Part &p, Part &q) {
    const double eig = p.e, ko = p.e / RM;
    const double broi = 1.0 + 2.0 * ko, bro = 1.0 / broi;
    double sinthe = 0.0, costhe = 0.0, br, aux, rejf3, temp;
    const double alph1 = log(broi), alph2 = ko * (broi + 1.0) * (bro * bro), alpha = alph1 + alph2, rejmax = broi + bro;
    do {
        const uint4 b = g.block();
        const double r1 = u32d(b.x), r2 = u32d(b.y), r3 = u32d(b.z);
        if (ko > 2.0) {
            if (r1 * alpha < alph1) br = exp(alph1 * r2) * bro;
            else br = sqrt(r2 * (broi * broi) + (1.0 - r2)) * bro;
            temp = (1.0 - br) / (ko * br);
            sinthe = fmax(0.0, temp * (2.0 - temp));
            aux = 1.0 + (br * br);
            rejf3 = aux - br * sinthe;
            if (r3 * aux > rejf3) { br = -1.0; continue; }
        } else {
            br = bro + (1.0 - bro) * r1;
            temp = (1.0 - br) / (ko * br);
            sinthe = fmax(0.0, temp * (2.0 - temp));
            rejf3 = 1.0 + br * br - br * sinthe;
            if (r2 * br * rejmax > rejf3) { br = -1.0; continue; }
        }
    } while ((br < bro) || (br > 1));
    costhe = 1.0 - temp;
    sinthe = sqrt(sinthe);
    const double esg = br * eig, ese = eig - esg + RM;
    p.e = esg;
    Frame f;
    azimuth_blk(g, f.cphi, f.sphi);
    f.A = p.u; f.B = p.v; f.C = p.w;
    frame_apply(f, costhe, sinthe, p);
    aux = 1.0 + br * br - 2.0 * br * costhe;
    if (aux > 1.0E-8) {
        costhe = (1.0 - br * costhe) / sqrt(aux);
        sinthe = (1.0 - costhe) * (1.0 + costhe);
        sinthe = (sinthe > 0.0) ? -sqrt(sinthe) : 0.0;
    } else {
        costhe = 0.0; sinthe = -1.0;
    }
    uphi32(f, costhe, sinthe, q, p);
    q.e = ese;
    q.iq = -1;
}

__device__ __forceinline__ bool moller_b(const DevProblem &P, Rng &g, Part &p, Part &q, int imed) {
    const MedRec &M = P.med[imed];
    const double eie = p.e, ekin = eie - RM, te = M.te;
    if (ekin <= 2.0 * te) return false;
    const double t0 = ekin / RM, e0 = t0 + 1.0, extrae = eie - M.thmoll;
    const double g2 = (t0 * t0) / (e0 * e0), g3 = (2.0 * t0 + 1.0) / (e0 * e0);
    const double gmax = (1.0 + 1.25 * g2);
    double br;
    for (;;) {
        const uint4 b = g.block();
        br = te / (ekin - extrae * u32d(b.x));
        double r = br / (1.0 - br);
        if (!(u32d(b.y) * gmax > (1.0 + g2 * (br * br) + r * (r - g3)))) break;
        br = te / (ekin - extrae * u32d(b.z));
        r = br / (1.0 - br);
        if (!(u32d(b.w) * gmax > (1.0 + g2 * (br * br) + r * (r - g3)))) break;
    }
    const double ekse2 = br * ekin, ese1 = eie - ekse2, ese2 = ekse2 + RM;
    p.e = ese1;
    q.e = ese2;
    const double h1 = (eie + RM) / ekin;
    double costh = h1 * (ese1 - RM) / (ese1 + RM);
    double sinthe = sqrt(1.0 - costh), costhe = sqrt(costh);
    Frame f;
    azimuth_blk(g, f.cphi, f.sphi);
    f.A = p.u; f.B = p.v; f.C = p.w;
    frame_apply(f, costhe, sinthe, p);
    q.iq = -1;
    costh = h1 * (ese2 - RM) / (ese2 + RM);
    sinthe = -sqrt(1.0 - costh);
    costhe = sqrt(costh);
    uphi32(f, costhe, sinthe, q, p);
    return true;
}

// ---------------------------------------------------------------------------------------------
// CSDA helpers in mixed precision: energies and path lengths are fp64 quantities, but the series below only
// needs RATIOS to ~1e-7, so the logs / divisions (software sequences in fp64) are done in fp32.  Differences of
// nearly equal energies are taken in fp64 BEFORE the conversion (no cancellation in fp32).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double flog(double x) { return (double)__logf((float)x); }

// computeDrange(), src/ompmc.c:3979-4014
static __device__ __forceinline__ double drange_m(const ElecBin *B, double ekei, double ekef, double elkei, double elkef) {
    const float fedep = fdiv((float)(ekei - ekef), (float)ekei);
    const float elktmp = 0.5f * ((float)elkei + (float)elkef + 0.25f * fedep * fedep * (1.0f + fedep * (1.0f + 0.875f * fedep)));
    const double2 cd = ldg2(&B->dedx1);
    const float d1 = (float)cd.x;
    const float dedxmid = frcp(elktmp * d1 + (float)cd.y);
    float aux = d1 * dedxmid;
    const float tf = 2.0f - fedep;
    aux = fdiv(aux * (1.0f + 2.0f * aux) * fedep * fedep, 6.0f * tf * tf);
    return (double)(fedep * dedxmid * (1.0f + aux)) * ekei;
}

// computeEloss(), src/ompmc.c:4016-4108; rinv = 1 / rhof
static __device__ __noinline__ double eloss_m(const ElecBin *B0, const MedRec &M, double rhof, double rinv, double tustep, double range,
                                              double eke, double elke, int lelke) {
    double de;
    double tuss = range - __ldg(&B0[lelke].range_ep) * rinv;
    if (tuss >= tustep) {
        const double2 cd = ldg2(&B0[lelke].dedx1);
        const float d1 = (float)cd.x;
        const float dedxmid = (float)elke * d1 + (float)cd.y;
        const float aux = fdiv(d1, dedxmid);
        const float def = dedxmid * (float)(tustep * rhof);
        const float fedep = fdiv(def, (float)eke);
        de = (double)(def * (1.0f - 0.5f * fedep * aux * (1.0f - 0.333333f * fedep * (aux - 1.0f - 0.25f * fedep * (2.0f - aux * (4.0f - aux))))));
    } else {
        int lt = lelke;
        tuss = (range - tustep) * rhof;
        if (tuss <= 0) {
            de = eke - M.te * 0.99;
        } else {
            while (tuss < __ldg(&B0[lt].range_ep)) lt -= 1;
            const float elktmp = fdiv((float)(lt + 2) - (float)M.eke0, (float)M.eke1);
            const double eketmp = __ldg(&B0[lt + 1].e_array);
            tuss = (__ldg(&B0[lt + 1].range_ep) - tuss) * rinv;
            const double2 cd = ldg2(&B0[lt].dedx1);
            const float d1 = (float)cd.x;
            const float dedxmid = elktmp * d1 + (float)cd.y;
            const float aux = fdiv(d1, dedxmid);
            const float def = dedxmid * (float)(tuss * rhof);
            const float fedep = fdiv(def, (float)eketmp);
            de = (double)(def * (1.0f - 0.5f * fedep * aux * (1.0f - 0.333333f * fedep * (aux - 1.0f - 0.25f * fedep * (2.0f - aux * (4.0f - aux))))));
            de += eke - eketmp;
        }
    }
    return de;
}

}  // namespace omc

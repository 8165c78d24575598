// omc_physics.cuh -- device functions of the ompMC shower() hot path.
//
// Everything here works on particles held in REGISTERS (struct Part) and an explicit RNG object;
// how particles are scheduled (one history per thread with a LIFO stack, or wavefront queues) is
// the kernels' business (omc_lockstep.cu / omc_wavefront.cu).  Each function cites the reference
// lines whose behaviour it reproduces; reference quirks that change sampled distributions or the
// number of random draws are reproduced on purpose.
#pragma once
#include "omc_types.cuh"

// The transport kernels are instruction-cache bound when every sampler is inlined at every call site
// (the fused wave kernel was ~200 KB of SASS with stall_no_instruction as its top stall reason), so the
// samplers and the Philox refill are real functions: one copy each.
#define OMC_FN static __device__ __noinline__

namespace omc {

// ---------------------------------------------------------------------------------------------
// RNG: Philox4x32-10, one stream per (history, sub-stream).  key = (seed0, seed1),
// counter = (block, stream, hist_lo, hist_hi); draw k = word (k & 3) of block k >> 2, as u32 * 2^-32.
// Replaces RANMAR (src/omc_random.c:58-187); same [0,1) contract as setRandom() (:172-187).
// ---------------------------------------------------------------------------------------------
struct Rng {
    uint32_t k0, k1;        // key
    uint32_t blk, stream;   // counter words 0,1
    uint32_t h0, h1;        // counter words 2,3 = history id
    uint32_t b0, b1, b2, b3;
    uint32_t pos;           // next unread word of the current block (4 = none left)

    __device__ __forceinline__ void seed(uint32_t s0, uint32_t s1, unsigned long long hist, uint32_t strm, uint32_t ndrawn = 0) {
        k0 = s0; k1 = s1; stream = strm;
        h0 = (uint32_t)hist; h1 = (uint32_t)(hist >> 32);
        blk = ndrawn >> 2; pos = 4;
        if (ndrawn & 3u) { refill(); pos = ndrawn & 3u; }
    }
    __device__ __forceinline__ uint32_t ndraws() const { return blk * 4u - (4u - pos); }
    // drop the unread words of the current block, so that the lanes of a warp that are about to run the
    // same sampling routine refill together (wavefront kernels only; a pure function of this stream)
    __device__ __forceinline__ void align() { pos = 4u; }

    __device__ __noinline__ void refill() {
        uint32_t c0 = blk, c1 = stream, c2 = h0, c3 = h1, ka = k0, kb = k1;
#pragma unroll
        for (int r = 0; r < 10; r++) {
            uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
            uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
            c0 = hi1 ^ c1 ^ ka; c1 = lo1; c2 = hi0 ^ c3 ^ kb; c3 = lo0;
            ka += 0x9E3779B9u; kb += 0xBB67AE85u;
        }
        b0 = c0; b1 = c1; b2 = c2; b3 = c3;
        blk += 1; pos = 0;
    }
    // One whole Philox block as a pure function of the counter: register-only, no buffered words.  The wavefront
    // kernels' photon path consumes its streams in whole blocks at fixed program points (block()), so the state
    // never has to live in local memory and the lanes of a warp always generate together.
    static __device__ __forceinline__ uint4 philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t ka, uint32_t kb) {
#pragma unroll
        for (int r = 0; r < 10; r++) {
            const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
            const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
            c0 = hi1 ^ c1 ^ ka; c1 = lo1; c2 = hi0 ^ c3 ^ kb; c3 = lo0;
            ka += 0x9E3779B9u; kb += 0xBB67AE85u;
        }
        return make_uint4(c0, c1, c2, c3);
    }
    __device__ __forceinline__ uint4 block() {
        const uint4 r = philox(blk, stream, h0, h1, k0, k1);
        blk += 1u; pos = 4u;
        return r;
    }
    // seed at the first block boundary at or after draw `ndrawn` (block-wise consumers)
    __device__ __forceinline__ void seed_blocks(uint32_t s0, uint32_t s1, uint32_t hlo, uint32_t hhi, uint32_t strm, uint32_t ndrawn) {
        k0 = s0; k1 = s1; stream = strm; h0 = hlo; h1 = hhi;
        blk = (ndrawn + 3u) >> 2; pos = 4u;
    }
    __device__ __forceinline__ double next() {
        if (pos >= 4u) refill();
        uint32_t w = pos == 0u ? b0 : (pos == 1u ? b1 : (pos == 2u ? b2 : b3));
        pos += 1;
        return (double)w * (1.0 / 4294967296.0);
    }
};

// ---------------------------------------------------------------------------------------------
// table lookup helpers: pwlfInterval / pwlfEval, src/ompmc.c:203-211
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int elec_interval(const MedRec &m, double elke) { return (int)(elke * m.eke1 + m.eke0) - 1; }
__device__ __forceinline__ double pwl(double lvar, double c1, double c0) { return lvar * c1 + c0; }
// one 16-byte load for a {c1, c0} coefficient pair (the records of omc_types.cuh keep every pair adjacent and 16-byte
// aligned): a table gather costs L1 a wavefront per distinct sector and per INSTRUCTION, so halving the instructions
// halves that (ncu: l1tex throughput of esize_kernel 73 %)
__device__ __forceinline__ double2 ldg2(const double *pair) { return __ldg(reinterpret_cast<const double2 *>(pair)); }
__device__ __forceinline__ double pwl2(double lvar, const double *pair) { const double2 c = ldg2(pair); return lvar * c.x + c.y; }

// 32-byte sector loads through the read-only path
__device__ __forceinline__ RegionRec load_region(const DevProblem &P, int ir) {
    const double2 *q = reinterpret_cast<const double2 *>(P.reg + ir);
    const double2 a = __ldg(q), b = __ldg(q + 1);
    RegionRec r;
    r.rhof = a.x; r.ecut = a.y; r.pcut = b.x;
    r.med = (int)(__double_as_longlong(b.y) & 0xffffffffll);
    r.pad = 0;
    return r;
}
__device__ __forceinline__ int region_med(const DevProblem &P, int ir) { return __ldg(&P.reg[ir].med); }

// Wavefront kernels: when ecut/pcut depend on the medium only (always true for regions filled by the
// reference's initRegions(), omc_dosxyz.c:890-962) a voxel is an 8-byte {float rhof, int med} record and the
// cut-offs come from the per-medium record; 3M voxels then take 24 MB (L2-resident) instead of 96 MB.
__device__ __forceinline__ RegionRec load_region_w(const DevProblem &P, int ir) {
    if (P.reg8 == nullptr) return load_region(P, ir);
    const int2 a = __ldg(reinterpret_cast<const int2 *>(P.reg8) + ir);
    RegionRec r;
    r.rhof = (double)__int_as_float(a.x);
    r.med = a.y;
    r.pad = 0;
    if (a.y >= 0) { r.ecut = P.med[a.y].ecut; r.pcut = P.med[a.y].pcut; }
    else { r.ecut = 0.0; r.pcut = 0.0; }
    return r;
}

// region record without the cut-offs: {rhof, med} only
__device__ __forceinline__ void load_region_rm(const DevProblem &P, int ir, double &rhof, int &med) {
    if (P.reg8 == nullptr) {
        const RegionRec r = load_region(P, ir);
        rhof = r.rhof; med = r.med;
        return;
    }
    const int2 a = __ldg(reinterpret_cast<const int2 *>(P.reg8) + ir);
    rhof = (double)__int_as_float(a.x);
    med = a.y;
}

// total photon cross section per unit density ratio, 1 / (gmfp * cohfac) of photon() src/ompmc.c:1954-1966
__device__ __forceinline__ double phot_sig0(const DevProblem &P, int imed, double gle) {
    const MedRec &M = P.med[imed];
    const int lgle = (int)(gle * M.ge1 + M.ge0) - 1;
    const PhotBin *B = P.phot + imed * MXGE + lgle;
    const double2 a = __ldg(reinterpret_cast<const double2 *>(&B->gmfp1));
    const double2 b = __ldg(reinterpret_cast<const double2 *>(&B->cohe1));
    return 1.0 / (pwl(gle, a.x, a.y) * pwl(gle, b.x, b.y));
}

// interaction choice at a photon interaction site, photon() src/ompmc.c:2027-2067: 1 Compton, 2 pair, 3 photo, 4 Rayleigh
// (r1, r2: the Rayleigh test and the branching draw)
__device__ __forceinline__ int photon_interaction_type(const DevProblem &P, int imed, double gle, double eig, double r1, double r2) {
    const MedRec &M = P.med[imed];
    const int lgle = (int)(gle * M.ge1 + M.ge0) - 1;
    const PhotBin *B = P.phot + imed * MXGE + lgle;
    const double coh = pwl(gle, __ldg(&B->cohe1), __ldg(&B->cohe0));
    if (r1 <= 1.0 - coh) return 4;
    const double r = r2;
    const double gbr1 = pwl(gle, __ldg(&B->gbr11), __ldg(&B->gbr10));
    if (r <= gbr1 && eig > 2.0 * RM) return 2;
    const double gbr2 = pwl(gle, __ldg(&B->gbr21), __ldg(&B->gbr20));
    return (r < gbr2) ? 1 : 3;
}

// index i with b[i] <= x < b[i+1], clamped to the first / last bin for points outside [b[0], b[n]); uniform grids
// by division (corrected against the tabulated planes, so the answer is the same as a search), others by bisection
__device__ __forceinline__ int find_bin(const double *b, int n, double x, double inv, bool uniform) {
    int i;
    if (uniform) {
        i = (int)((x - __ldg(b)) * inv);
        i = max(0, min(i, n - 1));
        if (x < __ldg(b + i)) i -= 1;
        else if (x >= __ldg(b + i + 1)) i += 1;
        i = max(0, min(i, n - 1));
    } else {
        int lo = 0, hi = n;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (x >= __ldg(b + mid)) lo = mid; else hi = mid;
        }
        i = lo;
    }
    return i;
}

// ---------------------------------------------------------------------------------------------
// geometry: howfar()/hownear(), omc_dosxyz.c:187-334.  Axis test order z, x, y; strict '<'.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void decode_region(const DevProblem &P, int irl, int &irx, int &iry, int &irz) {
    irx = (irl - 1) % P.isize;
    irz = (irl - 1 - irx) / P.ijmax;
    iry = ((irl - 1 - irx) - irz * P.ijmax) / P.isize;
}

__device__ __forceinline__ void howfar_i(const DevProblem &P, const Part &p, int &idisc, int &irnew, double &ustep) {
    const int irl = p.ir;
    if (irl == 0) { idisc = 1; return; }
    int irx, iry, irz;
    decode_region(P, irl, irx, iry, irz);
    double dist;
    if (p.w > 0.0) {
        dist = (__ldg(P.zb + irz + 1) - p.z) / p.w;
        if (dist < ustep) { ustep = dist; irnew = (irz != P.ksize - 1) ? irl + P.ijmax : 0; }
    } else if (p.w < 0.0) {
        dist = -(p.z - __ldg(P.zb + irz)) / p.w;
        if (dist < ustep) { ustep = dist; irnew = (irz != 0) ? irl - P.ijmax : 0; }
    }
    if (p.u > 0.0) {
        dist = (__ldg(P.xb + irx + 1) - p.x) / p.u;
        if (dist < ustep) { ustep = dist; irnew = (irx != P.isize - 1) ? irl + 1 : 0; }
    } else if (p.u < 0.0) {
        dist = -(p.x - __ldg(P.xb + irx)) / p.u;
        if (dist < ustep) { ustep = dist; irnew = (irx != 0) ? irl - 1 : 0; }
    }
    if (p.v > 0.0) {
        dist = (__ldg(P.yb + iry + 1) - p.y) / p.v;
        if (dist < ustep) { ustep = dist; irnew = (iry != P.jsize - 1) ? irl + P.isize : 0; }
    } else if (p.v < 0.0) {
        dist = -(p.y - __ldg(P.yb + iry)) / p.v;
        if (dist < ustep) { ustep = dist; irnew = (iry != 0) ? irl - P.isize : 0; }
    }
}
OMC_FN void howfar(const DevProblem &P, const Part &p, int &idisc, int &irnew, double &ustep) { howfar_i(P, p, idisc, irnew, ustep); }

__device__ __forceinline__ double hownear_i(const DevProblem &P, const Part &p) {
    const int irl = p.ir;
    if (irl == 0) return 0.0;
    int irx, iry, irz;
    decode_region(P, irl, irx, iry, irz);
    double t = 1.0E10;
    t = fmin(t, __ldg(P.xb + irx + 1) - p.x); t = fmin(t, p.x - __ldg(P.xb + irx));
    t = fmin(t, __ldg(P.yb + iry + 1) - p.y); t = fmin(t, p.y - __ldg(P.yb + iry));
    t = fmin(t, __ldg(P.zb + irz + 1) - p.z); t = fmin(t, p.z - __ldg(P.zb + irz));
    return t;
}
OMC_FN double hownear(const DevProblem &P, const Part &p) { return hownear_i(P, p); }

// ---------------------------------------------------------------------------------------------
// azimuth + rotations: selectAzimuthalAngle / uphi21 / uphi32, src/ompmc.c:101-199
// ---------------------------------------------------------------------------------------------
OMC_FN void azimuth(Rng &g, double &cphi, double &sphi) {
    double x, x2, y, y2, r2;
    do {
        x = g.next(); x = 2.0 * x - 1.0; x2 = x * x;
        y = g.next(); y2 = y * y;
        r2 = x2 + y2;
    } while (r2 > 1.0);
    r2 = 1 / r2;
    cphi = (x2 - y2) * r2;
    sphi = 2.0 * x * y * r2;
}

struct Frame { double A, B, C, cphi, sphi; };

__device__ __forceinline__ void frame_apply(const Frame &f, double costhe, double sinthe, Part &p) {
    double sinps2 = f.A * f.A + f.B * f.B;
    if (sinps2 < 1.0E-20) {
        p.u = sinthe * f.cphi; p.v = sinthe * f.sphi; p.w = f.C * costhe;
    } else {
        double sinpsi = sqrt(sinps2);
        double us = sinthe * f.cphi, vs = sinthe * f.sphi;
        double sindel = f.B / sinpsi, cosdel = f.A / sinpsi;
        p.u = f.C * cosdel * us - sindel * vs + f.A * costhe;
        p.v = f.C * sindel * us + cosdel * vs + f.B * costhe;
        p.w = -sinpsi * us + f.C * costhe;
    }
}
// uphi21: draw azimuth, frame = direction of p, rotate p
__device__ __forceinline__ void uphi21(Rng &g, Frame &f, double costhe, double sinthe, Part &p) {
    azimuth(g, f.cphi, f.sphi);
    f.A = p.u; f.B = p.v; f.C = p.w;
    frame_apply(f, costhe, sinthe, p);
}
// uphi32: q inherits position / region / weight of prev (transferProperties), direction from frame
__device__ __forceinline__ void uphi32(const Frame &f, double costhe, double sinthe, Part &q, const Part &prev) {
    q.x = prev.x; q.y = prev.y; q.z = prev.z; q.ir = prev.ir; q.wt = prev.wt;
    frame_apply(f, costhe, sinthe, q);
}

// ---------------------------------------------------------------------------------------------
// photon interactions
// ---------------------------------------------------------------------------------------------
constexpr double HC_INVERSE = 80.65506856998;
constexpr double TWICE_HC2 = 0.000307444456;

// rayleigh(), src/ompmc.c:1102-1145.  Q2: ibin == 0 always; Q3: medium-0 form factor tables.
OMC_FN void rayleigh(const DevProblem &P, Rng &g, Part &p, double pmax, double eig) {
    const double xmax = HC_INVERSE * eig;
    const double dwi = (double)OMC_MXRAYFF - 1.0;
    double r0, r1, xv, costhe, csqthe;
    do {
        r1 = g.next();
        do {
            r0 = g.next(); r0 *= pmax;
            int ibin = (int)r0 * dwi;
            int ib = __ldg(P.ray_i + ibin) - 1;
            if ((__ldg(P.ray_i + ibin + 1) - 1) > ib) {
                // the reference scans `while (r0 >= fcum[ib+1]) ib++` -- up to 99 dependent loads because ibin == 0
                // always (Q2); fcum is a cumulative distribution, so bisection finds the same ib
                int lo = ib, hi = OMC_MXRAYFF - 2;
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (r0 < __ldg(P.ray_fcum + mid + 1)) hi = mid; else lo = mid + 1;
                }
                ib = lo;
            }
            r0 = (r0 - __ldg(P.ray_fcum + ib)) * __ldg(P.ray_c + ib);
            xv = __ldg(P.ray_xgrid + ib) * exp(log(1.0 + r0) * __ldg(P.ray_b + ib));
        } while (xv >= xmax);
        xv /= eig;
        costhe = 1.0 - TWICE_HC2 * (xv * xv);
        csqthe = costhe * costhe;
    } while (2.0 * r1 >= (1.0 + csqthe));
    double sinthe = sqrt(1.0 - csqthe);
    Frame f;
    uphi21(g, f, costhe, sinthe, p);
}

__device__ __forceinline__ double pair_rej(double zbrang, double xi, double esedei, double eseder, double tteig) {
    double a = (1.0 + eseder) * (1.0 + esedei) / (2.0 * tteig);
    double xh = xi - 0.5;
    return 2.0 + 3.0 * (esedei + eseder) -
           4.0 * (esedei + eseder + 1.0 - 4.0 * (xh * xh)) * (1.0 + 0.25 * log((a * a) + zbrang * (xi * xi)));
}

// pair(), src/ompmc.c:1429-1667: p (photon) -> first charged particle, q -> the lower-energy one.
OMC_FN void pair(const DevProblem &P, Rng &g, Part &p, Part &q, int imed) {
    const MedRec &M = P.med[imed];
    const double eig = p.e;
    double ese1, ese2;
    int iq1, iq2, l, l1;
    if (eig <= 2.1) {
        double r0 = g.next(), r1 = g.next();
        ese2 = RM + 0.5 * r0 * (eig - 2.0 * RM);
        ese1 = eig - ese2;
        if (r1 < 0.5) { iq1 = -1; iq2 = 1; } else { iq1 = 1; iq2 = -1; }
    } else {
        double Amax, Bmax, delta, aux;
        const double delcm = M.delcm;
        if (eig < 50.0) {
            l = 4; l1 = l + 1;
            delta = 4.0 * delcm / eig;
            if (delta < 1.0) {
                Amax = M.dl[0][l] + delta * (M.dl[1][l] + delta * M.dl[2][l]);
                Bmax = M.dl[0][l1] + delta * (M.dl[1][l1] + delta * M.dl[2][l1]);
            } else {
                aux = log(delta + M.dl[5][l]);
                Amax = M.dl[3][l] + M.dl[4][l] * aux;
                Bmax = M.dl[3][l1] + M.dl[4][l1] * aux;
            }
            aux = 1.0 - 2.0 * RM / eig;
            aux = aux * aux;
            aux *= Amax / 3.0;
            aux /= (Bmax + aux);
        } else {
            l = 6;
            Amax = M.dl[0][l];
            Bmax = M.dl[0][l + 1];
            aux = M.bpar1 * (1.0 - M.bpar0 * RM / eig);
        }
        const double eavail = eig - 2.0 * RM;
        double rejf, rejmax, r4;
        do {
            double br, r0 = g.next(), r1 = g.next();
            r4 = g.next();
            if (r0 > aux) {
                br = 0.5 * r1; rejmax = Bmax; l1 = l + 1;
            } else {
                double r2 = g.next(), r3 = g.next();
                br = 0.5 * (1.0 - fmax(fmax(r1, r2), r3)); rejmax = Amax; l1 = l;
            }
            ese2 = br * eavail + RM;
            ese1 = eig - ese2;
            delta = (eig * delcm) / (ese2 * ese1);
            if (delta < 1.0) rejf = M.dl[0][l1] + delta * (M.dl[1][l1] + delta * M.dl[2][l1]);
            else rejf = M.dl[3][l1] + M.dl[4][l1] * log(delta + M.dl[5][l1]);
        } while (r4 * rejmax > rejf);
        ese1 = eig - ese2;
        r4 = g.next();
        if (r4 < 0.5) { iq1 = -1; iq2 = 1; } else { iq1 = 1; iq2 = -1; }
    }
    p.e = ese1;
    q.e = ese2;
    // Motz-Olsen-Koch leading-term angles (iprdst = 2), :1573-1658.  Q4: the accept test is '&&'.
    const double zb = M.zbrang;
    Frame f;
#pragma unroll 1
    for (int j = 0; j < 2; j++) {
        double ese = (j == 0) ? ese1 : ese2;
        double tteig = eig / RM, ttese = ese / RM;
        double esedei = ttese / (tteig - ttese), eseder = 1.0 / esedei;
        double pt = 3.14159265358979323846 * ttese;
        double ximin = 1.0 / (1.0 + (pt * pt));
        double rejmin = pair_rej(zb, ximin, esedei, eseder, tteig);
        double y2 = 2.0 / tteig;
        double ya = y2 * y2;
        double xitry = fmax(0.01, fmax(ximin, fmin(0.5, sqrt(ya / zb))));
        double galpha = 1.0 + 0.25 * log(ya + zb * xitry * xitry);
        double gbeta = 0.5 * zb * xitry / (ya + zb * xitry * xitry);
        galpha -= gbeta * (xitry - 0.5);
        double ximid = galpha / (3.0 * gbeta);
        if (galpha >= 0.0) ximid = 0.5 - ximid + sqrt((ximid * ximid) + 0.25);
        else ximid = 0.5 - ximid - sqrt((ximid * ximid) + 0.25);
        ximid = fmax(0.01, fmax(ximin, fmin(0.5, ximid)));
        double rejmid = pair_rej(zb, ximid, esedei, eseder, tteig);
        double rejtop = 1.0 * fmax(rejmin, rejmid);
        double theta, rtest, rejfactor;
        do {
            double xitst = g.next();
            double rejtst = pair_rej(zb, xitst, esedei, eseder, tteig);
            rtest = g.next();
            theta = sqrt(1.0 / xitst - 1.0) / ttese;
            rejfactor = rejtst / rejtop;
        } while ((rtest > rejfactor) && (theta >= 3.14159265358979323846));
        double sinthe = sin(theta), costhe = cos(theta);
        if (j == 0) {
            uphi21(g, f, costhe, sinthe, p);
        } else {
            sinthe = -sinthe;
            uphi32(f, costhe, sinthe, q, p);
        }
    }
    q.iq = iq2;
    p.iq = iq1;
}

// compton(), src/ompmc.c:1670-1783: p -> scattered photon, q -> recoil electron
OMC_FN void compton(Rng &g, Part &p, Part &q) {
    const double eig = p.e, ko = p.e / RM;
    const double broi = 1.0 + 2.0 * ko, bro = 1.0 / broi;
    bool first_time = true;
    double sinthe = 0.0, costhe = 0.0, br, r1, r2, r3, aux, rejf3, temp;
    double alph1 = 0.0, alph2 = 0.0, alpha = 0.0, rejmax = 0.0;
    do {
        if (ko > 2.0) {
            if (first_time) {
                alph1 = log(broi);
                alph2 = ko * (broi + 1.0) * (bro * bro);
                alpha = alph1 + alph2;
            }
            do {
                r1 = g.next(); r2 = g.next();
                if (r1 * alpha < alph1) br = exp(alph1 * r2) * bro;
                else br = sqrt(r2 * (broi * broi) + (1.0 - r2)) * bro;
                temp = (1.0 - br) / (ko * br);
                sinthe = fmax(0.0, temp * (2.0 - temp));
                aux = 1.0 + (br * br);
                rejf3 = aux - br * sinthe;
                r3 = g.next();
            } while (r3 * aux > rejf3);
        } else {
            if (first_time) rejmax = broi + bro;
            do {
                r1 = g.next(); r2 = g.next();
                br = bro + (1.0 - bro) * r1;
                temp = (1.0 - br) / (ko * br);
                sinthe = fmax(0.0, temp * (2.0 - temp));
                rejf3 = 1.0 + br * br - br * sinthe;
            } while (r2 * br * rejmax > rejf3);
        }
        first_time = false;
    } while ((br < bro) || (br > 1));
    costhe = 1.0 - temp;
    sinthe = sqrt(sinthe);
    const double esg = br * eig, ese = eig - esg + RM;
    p.e = esg;
    Frame f;
    uphi21(g, f, costhe, sinthe, p);
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

// photo(), src/ompmc.c:1786-1846: photon becomes an electron of energy hv + RM, Sauter angle
OMC_FN void photo(Rng &g, Part &p, double ecut) {
    p.e += RM;
    p.iq = -1;
    const double eelec = p.e;
    if (eelec > ecut) {
        double beta = sqrt((eelec - RM) * (eelec + RM)) / eelec;
        double costhe, sinth2, gamma = eelec / RM;
        double alpha = 0.5 * gamma - 0.5 + 1.0 / gamma;
        double ratio = beta / alpha, rn2, xi;
        do {
            double rn = g.next();
            rn = 2.0 * rn - 1.0;
            if (ratio <= 0.2) {
                double fkappa = rn + 0.5 * ratio * (1.0 - rn) * (1.0 + rn);
                if (gamma < 100.0) costhe = (beta + fkappa) / (1.0 + beta * fkappa);
                else if (fkappa > 0.0) {
                    double gm1 = gamma - 1.0;
                    costhe = 1.0 - (1.0 - fkappa) * (gamma - 3.0) / (2.0 * (1.0 + fkappa) * (gm1 * gm1 * gm1));
                } else costhe = (beta + fkappa) / (1.0 + beta * fkappa);
                xi = (1.0 + beta * fkappa) * (gamma * gamma);
            } else {
                xi = (gamma * gamma) * (1.0 + alpha * (sqrt(1.0 + ratio * (2.0 * rn + ratio)) - 1.0));
                costhe = (1.0 - 1.0 / xi) / beta;
            }
            sinth2 = fmax((1.0 - costhe) * (1.0 + costhe), 0.0);
            rn2 = g.next();
        } while (rn2 > 0.5 * (1.0 + gamma) * sinth2 * xi / gamma);
        double sinthe = sqrt(sinth2);
        Frame f;
        uphi21(g, f, costhe, sinthe, p);
    }
}

// ---------------------------------------------------------------------------------------------
// elastic scattering of electrons
// ---------------------------------------------------------------------------------------------
struct SpinState { int i, j; };

// spinRejection(), src/ompmc.c:3097-3168
OMC_FN double spin_rejection(const DevProblem &P, Rng &g, int imed, int qel, double elke, double beta2, double q1,
                                        double cost, bool &spin_index, bool is_single, SpinState &sr) {
    if (spin_index) {
        spin_index = false;
        double ai;
        if (beta2 >= P.b2spin_min) {
            ai = (beta2 - P.b2spin_min) * P.dbeta2i;
            sr.i = (int)ai; ai -= (double)sr.i; sr.i += 16;
        } else if (elke > P.espml) {
            ai = (elke - P.espml) * P.dleneri;
            sr.i = (int)ai; ai -= sr.i;
        } else {
            sr.i = 0; ai = -1.0;
        }
        double r = g.next();
        if (r < ai) sr.i += 1;
        if (is_single) {
            sr.j = 0;
        } else {
            double qq1 = 2.0 * q1;
            qq1 = qq1 / (1.0 + qq1);
            double aj = qq1 * P.dqq1i;
            sr.j = (int)aj;
            if (sr.j >= 15) {
                sr.j = 15;
            } else {
                aj -= (double)sr.j;
                r = g.next();
                if (r < aj) sr.j += 1;
            }
        }
    }
    double xi = sqrt(0.5 * (1.0 - cost));
    double ak = xi * 31;
    int k = (int)ak;
    ak -= (double)k;
    const double *row = P.spin_rej + (((size_t)(imed * 2 + qel) * OMC_SPIN_NE + sr.i) * OMC_SPIN_NQ + sr.j) * OMC_SPIN_NU;
    return (1.0 - ak) * __ldg(row + k) + ak * __ldg(row + k + 1);
}

// sscat(), src/ompmc.c:3170-3199
OMC_FN void sscat(const DevProblem &P, Rng &g, int imed, int qel, double chia2, double elke, double beta2,
                             double &cost, double &sint) {
    bool spin_index = true;
    SpinState sr;
    double xi, rejf, r;
    do {
        xi = g.next();
        xi = 2.0 * chia2 * xi / (1.0 - xi + chia2);
        cost = 1.0 - xi;
        rejf = spin_rejection(P, g, imed, qel, elke, beta2, 0.0, cost, spin_index, true, sr);
        r = g.next();
    } while (r > rejf);
    sint = sqrt(xi * (2.0 - xi));
}

struct MsState { int i, j; double omega2; };

// mscat(), src/ompmc.c:3606-3785.  Q1: du == 0 (its draw is still consumed); Q7: lambda > 1e5 leaves
// cost/sint untouched.
OMC_FN void mscat(const DevProblem &P, Rng &g, int imed, int qel, bool &spin_index, bool &find_index, double elke,
                             double beta2, double q1, double lambda, double chia2, double &cost, double &sint, MsState &ms,
                             SpinState &sr) {
    double xi, rejf, r;
    const double explambda = exp(-lambda);
    if (lambda <= 13.8) {
        double sprob = g.next();
        if (sprob < explambda) { cost = 1.0; sint = 0.0; return; }
        double wsum = (1.0 + lambda) * explambda;
        if (sprob < wsum) {
            do {
                xi = g.next();
                xi = 2.0 * chia2 * xi / (1.0 - xi + chia2);
                cost = 1.0 - xi;
                rejf = spin_rejection(P, g, imed, qel, elke, beta2, q1, cost, spin_index, false, sr);
                r = g.next();
            } while (r > rejf);
            sint = sqrt(xi * (2.0 - xi));
            return;
        }
        if (lambda <= 1) {
            int icount = 0;
            double wprob = explambda, sinz, cosz, phi;
            wsum = explambda;
            cost = 1.0; sint = 0.0;
            do {
                icount += 1;
                if (icount > 20) break;
                wprob = wprob * lambda / icount;
                wsum = wsum + wprob;
                do {
                    xi = g.next();
                    xi = 2.0 * chia2 * xi / (1.0 - xi + chia2);
                    cosz = 1.0 - xi;
                    rejf = spin_rejection(P, g, imed, qel, elke, beta2, q1, cosz, spin_index, false, sr);
                    r = g.next();
                } while (r > rejf);
                sinz = xi * (2.0 - xi);
                if (sinz > 1.0E-20) {
                    sinz = sqrt(sinz);
                    xi = g.next();
                    phi = xi * 6.2831853;
                    cost = cost * cosz - sint * sinz * cos(phi);
                    sint = sqrt(fmax(0.0, (1.0 - cost) * (1.0 + cost)));
                }
            } while (wsum <= sprob);
            return;
        }
    }
    if (lambda <= 1.0E5) {
        const double llmbda = log(lambda);
        if (find_index) {
            double ai = llmbda * P.dllambi;
            ms.i = (int)ai; ai -= (double)ms.i;
            xi = g.next();
            if (xi < ai) ms.i += 1;
            if (q1 < 1.0E-3) {
                ms.j = 0;
            } else if (q1 < 0.5) {
                double aj = q1 * P.dqmsi;
                ms.j = (int)aj; aj -= (double)ms.j;
                xi = g.next();
                if (xi < aj) ms.j += 1;
            } else {
                ms.j = 7;
            }
            if (llmbda < 2.2299)
                ms.omega2 = chia2 * (lambda + 4.0) *
                            (1.347006 + llmbda * (0.209364 - llmbda * (0.45525 - llmbda * (0.50142 - 0.081234 * llmbda))));
            else
                ms.omega2 = chia2 * (lambda + 4.0) * (-2.77164 + llmbda * (2.94874 - llmbda * (0.1535754 - llmbda * 0.00552888)));
            find_index = false;
        }
        const MsEntry *tab = P.ms + (ms.i * OMC_MS_NQ + ms.j) * OMC_MS_NU;
        do {
            xi = g.next();
            double ak = xi * 31;
            int k = (int)ak;
            ak -= k;
            if (ak > __ldg(&tab[k].wms)) k = __ldg(&tab[k].ims);
            double u = __ldg(&tab[k].ums);
            xi = g.next();                 // Q1: feeds the dead in-bin interpolation (du == 0), u is unchanged
            xi = ms.omega2 * u / (1.0 + 0.5 * ms.omega2 - u);
            if (xi > 1.99999) xi = 1.99999;
            cost = 1.0 - xi;
            rejf = spin_rejection(P, g, imed, qel, elke, beta2, q1, cost, spin_index, false, sr);
            r = g.next();
        } while (r > rejf);
        sint = sqrt(xi * (2.0 - xi));
    }
}

// msdist(), src/ompmc.c:3787-3976 (PRESTA-II): returns the straight-line step; (xf..wf) = end point
template <bool ALIGN = false>
__device__ __noinline__ double msdist(const DevProblem &P, Rng &g, const Part &p, int imed, int qel, double rhof, double de,
                                double tustep, double eke, double &xf, double &yf, double &zf, double &uf, double &vf,
                                double &wf) {
    const MedRec &M = P.med[imed];
    MsState ms;
    SpinState sr;
    double e = eke - 0.5 * de;
    const double tau = e / RM, tau2 = tau * tau;
    const double epsilon = de / eke, epsilonp = de / e;
    e *= (1.0 - (epsilonp * epsilonp) * (6.0 + 10.0 * tau + 5.0 * tau2) / (24.0 * tau2 + 72.0 * tau + 48.0));
    const double p2 = e * (e + 2.0 * RM);
    const double beta2 = p2 / (p2 + (RM * RM));
    double chia2 = M.xcc / (4.0 * p2 * M.blcc);
    double lambda = 0.5 * tustep * rhof * M.blcc / beta2;
    const double temp2 = 0.166666 * (4.0 + tau * (6.0 + tau * (7.0 + tau * (4.0 + tau)))) *
                         (epsilonp / ((tau + 1.0) * (tau + 2.0))) * (epsilonp / ((tau + 1.0) * (tau + 2.0)));
    lambda *= (1.0 - temp2);
    double elke = log(e);
    int lelke = elec_interval(M, elke);
    if (lelke < 0) { lelke = 0; elke = (1.0 - M.eke0) / M.eke1; }
    const ElecBin *B = P.ebin + (size_t)qel * P.nmed * MXEKE + imed * MXEKE + lelke;
    const double etap = pwl(elke, __ldg(&B->eta1), __ldg(&B->eta0));
    const double xi_corr = pwl(elke, __ldg(&B->q1c1), __ldg(&B->q1c0));
    double gamma = pwl(elke, __ldg(&B->q2c1), __ldg(&B->q2c0));
    const double ms_corr = pwl(elke, __ldg(&B->blcce1), __ldg(&B->blcce0));
    chia2 *= etap;
    lambda /= (etap * (1.0 + chia2));
    lambda *= ms_corr;
    const double chilog = log(1.0 + 1.0 / chia2);
    const double q1 = 2.0 * chia2 * (chilog * (1.0 + chia2) - 1.0);
    gamma = 6.0 * chia2 * (1.0 + chia2) * (chilog * (1.0 + 2.0 * chia2) - 2.0) / q1 * gamma;
    double xi = q1 * lambda;
    bool find_index = true, spin_index = true;
    double w1 = 1.0, sint1 = 0.0, cphi1, sphi1, w2 = 1.0, sint2 = 0.0, cphi2, sphi2;
    if (ALIGN) g.align();
    mscat(P, g, imed, qel, spin_index, find_index, elke, beta2, xi, lambda, chia2, w1, sint1, ms, sr);
    if (ALIGN) g.align();
    azimuth(g, cphi1, sphi1);
    if (ALIGN) g.align();
    mscat(P, g, imed, qel, spin_index, find_index, elke, beta2, xi, lambda, chia2, w2, sint2, ms, sr);
    if (ALIGN) g.align();
    azimuth(g, cphi2, sphi2);
    const double u2 = sint2 * cphi2, v2 = sint2 * sphi2;
    double u2p = w1 * u2 + sint1 * w2;
    double us = u2p * cphi1 - v2 * sphi1, vs = u2p * sphi1 + v2 * cphi1, ws = w1 * w2 - sint1 * u2;
    xi *= 2 * xi_corr;
    const double eta = g.next();
    const double eta1 = 0.5 * (1.0 - eta);
    double delta = 0.9082483 - (0.1020621 - 0.0263747 * gamma) * xi;
    double temp1 = 2.0 + tau;
    double temp = (2.0 + tau * temp1) / ((tau + 1.0) * temp1);
    temp -= (tau + 1.0) / ((tau + 2.0) * (chilog * (1.0 + chia2) - 1.0));
    temp *= epsilonp;
    temp1 = 1.0 - temp;
    delta += 0.40824829 * (epsilon * (tau + 1.0) / ((tau + 2.0) * (chilog * (1.0 + chia2) - 1.0) * (chilog * (1.0 + 2.0 * chia2) - 2.0)) -
                           0.25 * (temp * temp));
    const double b = eta * delta, cc = eta * (1.0 - delta);
    const double w1v2 = w1 * v2;
    double ut = b * sint1 * cphi1 + cc * (cphi1 * u2 - sphi1 * w1v2) + eta1 * us * temp1;
    double vt = b * sint1 * sphi1 + cc * (sphi1 * u2 + cphi1 * w1v2) + eta1 * vs * temp1;
    double wt = eta1 * (1.0 + temp) + b * w1 + cc * w2 + eta1 * ws * temp1;
    const double ustep = tustep * sqrt(ut * ut + vt * vt + wt * wt);
    const double u0 = p.u, v0 = p.v, w0 = p.w;
    const double sint02 = u0 * u0 + v0 * v0;
    if (sint02 > 1.0E-20) {
        double sint0 = sqrt(sint02), sint0i = 1.0 / sint0;
        double cphi0 = sint0i * u0, sphi0 = sint0i * v0;
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
    xf = p.x + tustep * ut; yf = p.y + tustep * vt; zf = p.z + tustep * wt;
    uf = us; vf = vs; wf = ws;
    return ustep;
}

// ---------------------------------------------------------------------------------------------
// CSDA helpers
// ---------------------------------------------------------------------------------------------
// computeDrange(), src/ompmc.c:3979-4014; B = bin record of (qel, imed, lelke)
OMC_FN double drange(const ElecBin *B, double ekei, double ekef, double elkei, double elkef) {
    const double fedep = 1.0 - ekef / ekei;
    const double elktmp = 0.5 * (elkei + elkef + 0.25 * fedep * fedep * (1.0 + fedep * (1.0 + 0.875 * fedep)));
    const double d1 = __ldg(&B->dedx1);
    double dedxmid = pwl(elktmp, d1, __ldg(&B->dedx0));
    dedxmid = 1.0 / dedxmid;
    double aux = d1 * dedxmid;
    aux = aux * (1.0 + 2.0 * aux) * fedep * fedep / (6.0 * (2.0 - fedep) * (2.0 - fedep));
    return fedep * ekei * dedxmid * (1.0 + aux);
}

// computeEloss(), src/ompmc.c:4016-4108; B0 = first bin record of (qel, imed)
OMC_FN double eloss(const ElecBin *B0, const MedRec &M, double rhof, double tustep, double range, double eke,
                               double elke, int lelke) {
    double aux, dedxmid, de, fedep;
    double tuss = range - __ldg(&B0[lelke].range_ep) / rhof;
    if (tuss >= tustep) {
        const double d1 = __ldg(&B0[lelke].dedx1);
        dedxmid = pwl(elke, d1, __ldg(&B0[lelke].dedx0));
        aux = d1 / dedxmid;
        de = dedxmid * tustep * rhof;
        fedep = de / eke;
        de *= (1.0 - 0.5 * fedep * aux * (1.0 - 0.333333 * fedep * (aux - 1.0 - 0.25 * fedep * (2.0 - aux * (4.0 - aux)))));
    } else {
        int lt = lelke;
        tuss = (range - tustep) * rhof;
        if (tuss <= 0) {
            de = eke - M.te * 0.99;
        } else {
            while (tuss < __ldg(&B0[lt].range_ep)) lt -= 1;
            const double elktmp = (lt + 2 - M.eke0) / M.eke1;
            const double eketmp = __ldg(&B0[lt + 1].e_array);
            tuss = (__ldg(&B0[lt + 1].range_ep) - tuss) / rhof;
            const double d1 = __ldg(&B0[lt].dedx1);
            dedxmid = pwl(elktmp, d1, __ldg(&B0[lt].dedx0));
            aux = d1 / dedxmid;
            de = dedxmid * tuss * rhof;
            fedep = de / eketmp;
            de *= (1.0 - 0.5 * fedep * aux * (1.0 - 0.333333 * fedep * (aux - 1.0 - 0.25 * fedep * (2.0 - aux * (4.0 - aux)))));
            de += eke - eketmp;
        }
    }
    return de;
}

// ---------------------------------------------------------------------------------------------
// discrete e-/e+ interactions.  Convention: p = the particle in the old stack slot, q = the one
// pushed on top (SURVEY.md Appendix A "Secondary placement").
// ---------------------------------------------------------------------------------------------
// photon Russian roulette when nsplit > 1 (e.g. src/ompmc.c:4150-4165): one draw per photon
__device__ __forceinline__ void roulette(Rng &g, Part &ph, int nsplit) {
    double r = g.next();
    if (r * (double)nsplit > 1.0) { ph.wt = 0.0; ph.e = 0.0; }
    else ph.wt *= nsplit;
}

// rannih(), src/ompmc.c:4111-4167 (Q6: one unused draw): p, q = the two 511 keV photons
OMC_FN void rannih(Rng &g, Part &p, Part &q, int nsplit) {
    double r = g.next();
    const double costhe = 2.0 * r - 1;
    const double sinthe = sqrt(fmax(0.0, (1.0 - costhe) * (1.0 + costhe)));
    r = g.next();
    double cphi, sphi;
    azimuth(g, cphi, sphi);
    p.e = RM; p.iq = 0;
    p.u = sinthe * cphi; p.v = sinthe * sphi; p.w = costhe;
    q = p;
    q.u = -1.0 * p.u; q.v = -1.0 * p.v; q.w = -1.0 * p.w;
    if (nsplit > 1) { roulette(g, p, nsplit); roulette(g, q, nsplit); }
}

// brems(), src/ompmc.c:4170-4356: p = radiating e-/e+ (keeps its direction), q = photon
OMC_FN void brems(const DevProblem &P, Rng &g, Part &p, Part &q, int imed, int nsplit) {
    const MedRec &M = P.med[imed];
    const double eie = p.e;
    const int l = (eie < 50.0) ? 1 : 3, l1 = l + 1;
    const double ekin = eie - RM, brmin = M.ap / ekin, waux = -log(brmin);
    const double a = p.u, b = p.v, cz = p.w;
    double sinpsi = a * a + b * b, sindel = 0.0, cosdel = 0.0;
    if (sinpsi > 1.0E-20) { sinpsi = sqrt(sinpsi); sindel = b / sinpsi; cosdel = a / sinpsi; }
    const double ztarg = M.zbrang, tteie = eie / RM;
    const double beta = sqrt((tteie - 1.0) * (tteie + 1.0)) / tteie;
    const double y2max = 2.0 * beta * (1.0 + beta) * tteie * tteie, y2maxi = 1.0 / y2max;
    const double z2max = y2max + 1.0, z2maxi = sqrt(z2max);
    double aux, br, delta, r6, r7, rejf, ese, esg, phi1, phi2;
    do {
        r6 = g.next(); r7 = g.next();
        br = brmin * exp(r6 * waux);
        esg = ekin * br;
        ese = eie - esg;
        delta = esg / eie / ese * M.delcm;
        aux = ese / eie;
        if (delta < 1.0) {
            phi1 = M.dl[0][l - 1] + delta * (M.dl[1][l - 1] + delta * M.dl[2][l - 1]);
            phi2 = M.dl[0][l1 - 1] + delta * (M.dl[1][l1 - 1] + delta * M.dl[2][l1 - 1]);
        } else {
            phi1 = M.dl[3][l - 1] + M.dl[4][l - 1] * log(delta + M.dl[5][l - 1]);
            phi2 = phi1;
        }
        rejf = (1.0 + (aux * aux)) * phi1 - 2.0 * aux * phi2 / 3.0;
    } while (r7 >= rejf);
    q = p;
    q.e = esg; q.iq = 0;
    double y2tst = 0.0;
    const double ttese = ese / RM, esedei = ttese / tteie;
    const double rjarg1 = 1.0 + esedei * esedei, rjarg2 = rjarg1 + 2.0 * esedei;
    double rjarg3, rtest = 1.0, rejtst = 0.0;
    aux = 2.0 * ese * tteie / esg;
    aux = aux * aux;
    const double aux1 = aux * ztarg;
    if (aux1 > 10.0) rjarg3 = -log(ztarg) + (1.0 - aux1) / (aux1 * aux1);
    else rjarg3 = log(aux / (1.0 + aux1));
    const double rejmax = rjarg1 * rjarg3 - rjarg2;
    while (rtest >= rejtst) {
        y2tst = g.next();
        rtest = g.next();
        const double aux3 = z2maxi / (y2tst + (1.0 - y2tst) * z2maxi);
        rtest = rtest * aux3 * rejmax;
        y2tst = (aux3 * aux3) - 1.0;
        const double a34 = (aux3 * aux3) * (aux3 * aux3);
        const double y2tst1 = esedei * y2tst / a34;
        const double aux4 = 16.0 * y2tst1 - rjarg2, aux5 = rjarg1 - 4.0 * y2tst1;
        if (rtest < aux4 + aux5 * rjarg3) break;
        const double aux2 = log(aux / (1.0 + aux1 / a34));
        rejtst = aux4 + aux5 * aux2;
    }
    const double costhe = 1.0 - 2.0 * y2tst * y2maxi;
    const double sinthe = sqrt(fmax(0.0, (1.0 - (costhe * costhe))));
    double cphi, sphi;
    azimuth(g, cphi, sphi);
    if (sinpsi >= 1.0E-10) {
        double us = sinthe * cphi, vs = sinthe * sphi;
        q.u = cz * cosdel * us - sindel * vs + a * costhe;
        q.v = cz * sindel * us + cosdel * vs + b * costhe;
        q.w = cz * costhe - sinpsi * us;
    } else {
        q.u = sinthe * cphi; q.v = sinthe * sphi; q.w = cz * costhe;
    }
    p.e = ese;
    if (nsplit > 1) roulette(g, q, nsplit);
}

// moller(), src/ompmc.c:4359-4435: returns false (nothing happens) below the kinematic threshold
OMC_FN bool moller(const DevProblem &P, Rng &g, Part &p, Part &q, int imed) {
    const MedRec &M = P.med[imed];
    const double eie = p.e, ekin = eie - RM;
    if (ekin <= 2.0 * M.te) return false;
    const double t0 = ekin / RM, e0 = t0 + 1.0, extrae = eie - M.thmoll;
    const double g2 = (t0 * t0) / (e0 * e0), g3 = (2.0 * t0 + 1.0) / (e0 * e0);
    const double gmax = (1.0 + 1.25 * g2);
    double br, rejf4, r, r27, r28;
    do {
        r27 = g.next();
        br = M.te / (ekin - extrae * r27);
        r = br / (1.0 - br);
        r28 = g.next();
        rejf4 = (1.0 + g2 * (br * br) + r * (r - g3));
        r28 *= gmax;
    } while (r28 > rejf4);
    const double ekse2 = br * ekin, ese1 = eie - ekse2, ese2 = ekse2 + RM;
    p.e = ese1;
    q.e = ese2;
    const double h1 = (eie + RM) / ekin;
    double costh = h1 * (ese1 - RM) / (ese1 + RM);
    double sinthe = sqrt(1.0 - costh), costhe = sqrt(costh);
    Frame f;
    uphi21(g, f, costhe, sinthe, p);
    q.iq = -1;
    costh = h1 * (ese2 - RM) / (ese2 + RM);
    sinthe = -sqrt(1.0 - costh);
    costhe = sqrt(costh);
    uphi32(f, costhe, sinthe, q, p);
    return true;
}

// bhabha(), src/ompmc.c:4438-4525: the lower-energy particle always ends up in q
OMC_FN void bhabha(const DevProblem &P, Rng &g, Part &p, Part &q, int imed) {
    const MedRec &M = P.med[imed];
    const double eip = p.e, ekin = eip - RM, t0 = ekin / RM, e0 = t0 + 1.0;
    const double yy = 1.0 / (t0 + 2.0), beta2 = ((e0 * e0) - 1.0) / (e0 * e0);
    const double ep0 = M.te / ekin, ep0c = 1.0 - ep0, yp = 1.0 - 2.0 * yy;
    const double b4 = yp * yp * yp, b3 = b4 + (yp * yp), b2 = yp * (3.0 + (yy * yy)), b1 = 2.0 - (yy * yy);
    double r3, r4, br, rejf2;
    do {
        r3 = g.next();
        br = ep0 / (1.0 - ep0c * r3);
        r4 = g.next();
        rejf2 = (1.0 - beta2 * br * (b1 - br * (b2 - br * (b3 - br * b4))));
    } while (r4 > rejf2);
    if (br < 0.5) {
        q.iq = -1;
    } else {
        p.iq = -1; q.iq = 1;
        br = 1.0 - br;
    }
    br = fmax(br, 0.0);
    const double ekse2 = br * ekin, ese1 = eip - ekse2, ese2 = ekse2 + RM;
    p.e = ese1; q.e = ese2;
    const double h1 = (eip + RM) / ekin;
    double costh = fmin(1.0, h1 * (ese1 - RM) / (ese1 + RM));
    double sinthe = sqrt(1.0 - costh), costhe = sqrt(costh);
    Frame f;
    uphi21(g, f, costhe, sinthe, p);
    costh = h1 * (ese2 - RM) / (ese2 + RM);
    sinthe = -sqrt(1.0 - costh);
    costhe = sqrt(costh);
    uphi32(f, costhe, sinthe, q, p);
}

// annih(), src/ompmc.c:4528-4645: positron -> two photons p, q
OMC_FN void annih(Rng &g, Part &p, Part &q, int nsplit) {
    const double avip = p.e + RM, a = avip / RM, gg = a - 1.0, t = gg - 1.0, pp = sqrt(a * t);
    const double pot = pp / t, ep0 = 1.0 / (a + pp), wsamp = log((1.0 - ep0) / ep0);
    const double aa = p.u, bb = p.v, cc = p.w;
    double sinpsi = (aa * aa) + (bb * bb), sindel = 0.0, cosdel = 0.0;
    if (sinpsi > 1.0E-20) { sinpsi = sqrt(sinpsi); sindel = bb / sinpsi; cosdel = aa / sinpsi; }
    double ep, rejf, r1, r2;
    do {
        r1 = g.next();
        ep = ep0 * exp(r1 * wsamp);
        r2 = g.next();
        double qv = ep * a - 1.0;
        rejf = 1.0 - (qv * qv) / (ep * ((a * a) - 2.0));
    } while (r2 > rejf);
    const double esg1 = avip * ep;
    p.e = esg1; p.iq = 0;
    double costhe = fmin(1.0, (esg1 - RM) * pot / esg1);
    double sinthe = sqrt(1.0 - (costhe * costhe));
    double sphi, cphi, us, vs;
    azimuth(g, cphi, sphi);
    q = p;
    if (sinpsi >= 1.0E-10) {
        us = sinthe * cphi; vs = sinthe * sphi;
        p.u = cc * cosdel * us - sindel * vs + aa * costhe;
        p.v = cc * sindel * us + cosdel * vs + bb * costhe;
        p.w = cc * costhe - sinpsi * us;
    } else {
        p.u = sinthe * cphi; p.v = sinthe * sphi; p.w = cc * costhe;
    }
    const double esg2 = avip - esg1;
    q.e = esg2; q.iq = 0;
    costhe = fmin(1.0, (esg2 - RM) * pot / esg2);
    sinthe = -sqrt(1.0 - (costhe * costhe));
    if (sinpsi >= 1.0E-10) {
        us = sinthe * cphi; vs = sinthe * sphi;
        q.u = cc * cosdel * us - sindel * vs + aa * costhe;
        q.v = cc * sindel * us + cosdel * vs + bb * costhe;
        q.w = cc * costhe - sinpsi * us;
    } else {
        q.u = sinthe * cphi; q.v = sinthe * sphi; q.w = cc * costhe;
    }
    if (nsplit > 1) { roulette(g, p, nsplit); roulette(g, q, nsplit); }
}

// ---------------------------------------------------------------------------------------------
// source: initHistory(), omc_dosxyz.c:964-1068.  Returns the sampled kinetic energy (score.ensrc).
// ---------------------------------------------------------------------------------------------
OMC_FN double init_history_dosxyz(const DevProblem &P, Rng &g, Part &p) {
    const SourceDosxyz &S = P.src;
    p.iq = S.charge;
    double ein;
    if (S.spectrum) {
        double r1 = g.next(), r2 = g.next();
        int k = (int)fmin(S.deltak * r1, S.deltak - 1.0);
        ein = __ldg(S.cdfinv1 + k) + r2 * __ldg(S.cdfinv2 + k);
    } else {
        ein = S.energy;
    }
    p.e = (p.iq != 0) ? ein + RM : ein;
    double rxyz;
    if (S.xsize == 0.0 || S.ysize == 0.0) {
        p.x = S.xinl; p.y = S.yinl;
        rxyz = sqrt((S.ssd * S.ssd) + (p.x * p.x) + (p.y * p.y));
        p.w = S.ssd / rxyz;
    } else {
        double fw, r3;
        do {
            r3 = g.next(); p.x = r3 * S.xsize + S.xinl;
            r3 = g.next(); p.y = r3 * S.ysize + S.yinl;
            r3 = g.next();
            rxyz = sqrt(S.ssd * S.ssd + p.x * p.x + p.y * p.y);
            p.w = S.ssd / rxyz;
            fw = p.w * p.w * p.w;
        } while (r3 >= fw);
    }
    p.z = __ldg(P.zb);
    p.u = p.x / rxyz; p.v = p.y / rxyz;
    int ix, iy;
    if (S.xsize == 0.0) ix = S.ixinl;
    else { ix = S.ixinl - 1; while ((__ldg(P.xb + ix + 1) < p.x) && ix < P.isize - 1) ix++; }
    if (S.ysize == 0.0) iy = S.iyinl;
    else { iy = S.iyinl - 1; while ((__ldg(P.yb + iy + 1) < p.y) && iy < P.jsize - 1) iy++; }
    p.ir = 1 + ix + iy * P.isize;
    p.wt = 1.0;
    return ein;
}

// initHistory(ibeamlet), omc_matrad.c:1084-1254 (Q13: the z clamp uses ybounds[0]; the 2*DBL_MIN nudges are
// no-ops for non-zero bounds).  Returns the sampled kinetic energy.
OMC_FN double init_history_matrad(const DevProblem &P, Rng &g, Part &p, int ib) {
    const SourceDosxyz &S = P.src;
    const SourceMatrad &M = P.msrc;
    p.iq = S.charge;
    double ein;
    if (S.spectrum) {
        double r1 = g.next(), r2 = g.next();
        int k = (int)fmin(S.deltak * r1, S.deltak - 1.0);
        ein = __ldg(S.cdfinv1 + k) + r2 * __ldg(S.cdfinv2 + k);
    } else {
        ein = S.energy;
    }
    p.e = (p.iq != 0) ? ein + RM : ein;
    const double r1 = g.next(), r2 = g.next();
    const double xiso = r1 * __ldg(M.xside1 + ib) + r2 * __ldg(M.xside2 + ib) + __ldg(M.xcorner + ib);
    const double yiso = r1 * __ldg(M.yside1 + ib) + r2 * __ldg(M.yside2 + ib) + __ldg(M.ycorner + ib);
    const double ziso = r1 * __ldg(M.zside1 + ib) + r2 * __ldg(M.zside2 + ib) + __ldg(M.zcorner + ib);
    const int ibeam = __ldg(M.ibeam + ib);
    const double dx = xiso - __ldg(M.xsource + ibeam), dy = yiso - __ldg(M.ysource + ibeam), dz = ziso - __ldg(M.zsource + ibeam);
    const double vnorm = sqrt((dx * dx) + (dy * dy) + (dz * dz));
    const double u = -dx / vnorm, v = -dy / vnorm, w = -dz / vnorm;
    const double xlo = __ldg(P.xb), xhi = __ldg(P.xb + P.isize), ylo = __ldg(P.yb), yhi = __ldg(P.yb + P.jsize);
    const double zlo = __ldg(P.zb), zhi = __ldg(P.zb + P.ksize);
    double ustep = 1.0E5, dist;
    if (u > 0.0) { dist = (xhi - xiso) / u; if (dist < ustep) ustep = dist; }
    if (u < 0.0) { dist = -(xiso - xlo) / u; if (dist < ustep) ustep = dist; }
    if (v > 0.0) { dist = (yhi - yiso) / v; if (dist < ustep) ustep = dist; }
    if (v < 0.0) { dist = -(yiso - ylo) / v; if (dist < ustep) ustep = dist; }
    if (w > 0.0) { dist = (zhi - ziso) / w; if (dist < ustep) ustep = dist; }
    if (w < 0.0) { dist = -(ziso - zlo) / w; if (dist < ustep) ustep = dist; }
    p.x = xiso + ustep * u; p.y = yiso + ustep * v; p.z = ziso + ustep * w;
    p.u = -u; p.v = -v; p.w = -w;
    const double tiny = 2.0 * 2.2250738585072014e-308;
    if (p.x < xlo) p.x = xlo + tiny;
    if (p.x > xhi) p.x = xhi - tiny;
    if (p.y < ylo) p.y = ylo + tiny;
    if (p.y > yhi) p.y = yhi - tiny;
    if (p.z < zlo) p.z = ylo + tiny;                               // Q13
    if (p.z > zhi) p.z = zhi - tiny;
    int ix = 0, iy = 0, iz = 0;
    // (omc_matrad.c:1233-1246 searches without an upper bound; with the Q13 clamp a z above zbounds[ksize] would walk past the
    // array -- on the device that is an illegal address for the whole context -- so the searches stop at the last voxel;
    // identical for every particle inside the grid)
    while (ix < P.isize - 1 && __ldg(P.xb + ix + 1) < p.x) ix++;
    while (iy < P.jsize - 1 && __ldg(P.yb + iy + 1) < p.y) iy++;
    while (iz < P.ksize - 1 && __ldg(P.zb + iz + 1) < p.z) iz++;
    p.ir = 1 + ix + iy * P.isize + iz * P.ijmax;
    p.wt = 1.0;
    return ein;
}

// source dispatch: ibeamlet < 0 -> omc_dosxyz point source, else the matRad beamlet `ibeamlet`
__device__ __forceinline__ double init_history(const DevProblem &P, Rng &g, Part &p, int ibeamlet) {
    return (ibeamlet < 0) ? init_history_dosxyz(P, g, p) : init_history_matrad(P, g, p, ibeamlet);
}

}  // namespace omc

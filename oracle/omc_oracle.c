/*
 * omc_oracle.c -- plain-C restatement of the ompMC shower() hot path.
 *
 * TEST INFRASTRUCTURE.  Header comment of omc_oracle.h states who may use it and how it is pinned
 * against the reference.  Every function cites the reference lines it follows; reference quirks
 * that change sampled distributions are reproduced on purpose and marked.
 *
 * Structure differs from the reference on purpose: particles are structs on an explicit
 * per-thread stack inside a history context, the RNG is an object (Philox per history, or RANMAR),
 * geometry callbacks take the particle as an argument, tables come in through omc_media_tables.
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "omc_oracle.h"
#include "omc_philox.h"
#include "omc_blob.h"

#define RM OMC_RM
#define MXGE OMC_MXGE
#define MXEKE OMC_MXEKE
#define MXSTACK 10000            /* src/ompmc.h:49 */
#define SGMFP 1.0E-05            /* src/ompmc.h:103 */
#define EPSEMFP 1.0E-5           /* src/ompmc.h:204 */
#define SKIN_DEPTH_FOR_BCA 3     /* src/ompmc.h:205 */
#define HC_INVERSE 80.65506856998
#define TWICE_HC2 0.000307444456
#define MXE_SPIN 15
#define MXQ_SPIN 15
#define MXU_SPIN 31
#define MXQ_MS 7
#define MXU_MS 31
#define LAMBMAX_MS 1.0E5
#define QMIN_MS 1.0E-3
#define QMAX_MS 0.5

/* ------------------------------------------------------------------------------------------ */
/* problem (read-only during transport) + score                                                */
/* ------------------------------------------------------------------------------------------ */
typedef struct problem {
    omc_media_tables T;
    omc_geometry G;
    omc_source_dosxyz S;
    omc_source_matrad M;      /* used instead of S when is_matrad */
    int is_matrad, ibeamlet;
    int nsplit;
    int nreg;
    double *endep, *accum, *accum2;   /* struct Score, omc_dosxyz.c:636-645 */
    double ensrc;
} problem;
static problem PB;
static omc_blob g_blob;

/* a1: one stack slot, src/ompmc.h:51-70 (dnear is written but never read by the reference) */
typedef struct part {
    int iq, ir;
    double e, x, y, z, u, v, w, wt;
} part;

/* a16: RANMAR state, src/omc_random.h:40-50 */
typedef struct ranmar {
    int urndm[97], buf[128];
    int crndm, cdrndm, cmrndm, ixx, jxx, pos;
} ranmar;

typedef struct hist_ctx {
    part *stk;
    int np, npold;
    int rng_mode;
    omc_philox ph;
    ranmar rm;
    unsigned ndeposit, flags;
    double edep_sum;
} hist_ctx;

/* per-history operation counts entering SURVEY.md 8d's algorithmic-bytes formula */
typedef struct work_counts { unsigned long long ausgab, howfar, hownear, pwlf, mscat, spin, hist; } work_counts;
static work_counts g_work;
#ifdef _OPENMP
#pragma omp threadprivate(g_work)
#endif
static work_counts g_work_total;

static int g_rng_mode = 1;
static uint32_t g_seed0 = 97, g_seed1 = 33;
static hist_ctx *g_ctx;            /* one per OpenMP thread */
static int g_nctx;

/* ---- a16: RNG ----------------------------------------------------------------------------- */
/* initRandom(), src/omc_random.c:58-137 */
static void ranmar_init(ranmar *r, int ixx, int jxx) {
    if (ixx <= 0 || ixx > 31328) ixx = 1802;
    if (jxx <= 0 || jxx > 31328) jxx = 9373;
    int i = (ixx / 177 % 177) + 2, j = (ixx % 177) + 2, k = (jxx / 169 % 178) + 1, l = jxx % 169;
    for (int ii = 0; ii < 97; ii++) {
        int s = 0, t = 8388608;
        for (int jj = 0; jj < 24; jj++) {
            int m = ((i * j % 179) * k) % 179;
            i = j; j = k; k = m;
            l = (53 * l + 1) % 169;
            if (l * m % 64 >= 32) s += t;
            t /= 2;
        }
        r->urndm[ii] = s;
    }
    r->crndm = 362436; r->cdrndm = 7654321; r->cmrndm = 16777213;
    r->ixx = 97; r->jxx = 33; r->pos = 128;
}
/* getRandom(), src/omc_random.c:142-170, incl. the if/else-if pointer wrap (Q14) */
static void ranmar_fill(ranmar *r) {
    for (int i = 0; i < 128; i++) {
        int iopt = r->urndm[r->ixx - 1] - r->urndm[r->jxx - 1];
        if (iopt < 0) iopt += 16777216;
        r->urndm[r->ixx - 1] = iopt;
        r->ixx -= 1; r->jxx -= 1;
        if (r->ixx == 0) r->ixx = 97;
        else if (r->jxx == 0) r->jxx = 97;
        r->crndm -= r->cdrndm;
        if (r->crndm < 0) r->crndm += r->cmrndm;
        iopt -= r->crndm;
        if (iopt < 0) iopt += 16777216;
        r->buf[i] = iopt;
    }
    r->pos = 0;
}
/* setRandom(), src/omc_random.c:172-187 */
static inline double rnd(hist_ctx *c) {
    if (c->rng_mode == 1) return omc_philox_next(&c->ph);
    if (c->rm.pos >= 128) ranmar_fill(&c->rm);
    return c->rm.buf[c->rm.pos++] * (1.0 / 16777216.0);
}

/* ---- a15: table lookup, src/ompmc.c:203-211 ------------------------------------------------ */
static inline int pwlf_interval(int idx, double lvar, const double *c1, const double *c0) {
    return (int)(lvar * c1[idx] + c0[idx]);
}
static inline double pwlf_eval(int idx, double lvar, const double *c1, const double *c0) {
    g_work.pwlf++;
    return lvar * c1[idx] + c0[idx];
}

/* ---- a19: ausgab(), omc_dosxyz.c:683-694 ----------------------------------------------------- */
static inline void deposit(hist_ctx *c, const part *p, double edep) {
    double en = p->wt * edep;
    g_work.ausgab++;
    c->ndeposit++;
    c->edep_sum += en;
#ifdef _OPENMP
#pragma omp atomic
#endif
    PB.endep[p->ir] += en;
}

/* ---- a17: howfar(), omc_dosxyz.c:187-297.  Axis order z, x, y with strict '<'. ---------------- */
static void howfar(const part *p, int *idisc, int *irnew, double *ustep) {
    const omc_geometry *g = &PB.G;
    int irl = p->ir;
    g_work.howfar++;
    if (irl == 0) { *idisc = 1; return; }
    int imax = g->isize, ijmax = g->isize * g->jsize;
    int irx = (irl - 1) % imax;
    int irz = (irl - 1 - irx) / ijmax;
    int iry = ((irl - 1 - irx) - irz * ijmax) / imax;
    double dist;
    if (p->w > 0.0) {
        dist = (g->zbounds[irz + 1] - p->z) / p->w;
        if (dist < *ustep) { *ustep = dist; *irnew = (irz != g->ksize - 1) ? irl + ijmax : 0; }
    } else if (p->w < 0.0) {
        dist = -(p->z - g->zbounds[irz]) / p->w;
        if (dist < *ustep) { *ustep = dist; *irnew = (irz != 0) ? irl - ijmax : 0; }
    }
    if (p->u > 0.0) {
        dist = (g->xbounds[irx + 1] - p->x) / p->u;
        if (dist < *ustep) { *ustep = dist; *irnew = (irx != g->isize - 1) ? irl + 1 : 0; }
    } else if (p->u < 0.0) {
        dist = -(p->x - g->xbounds[irx]) / p->u;
        if (dist < *ustep) { *ustep = dist; *irnew = (irx != 0) ? irl - 1 : 0; }
    }
    if (p->v > 0.0) {
        dist = (g->ybounds[iry + 1] - p->y) / p->v;
        if (dist < *ustep) { *ustep = dist; *irnew = (iry != g->jsize - 1) ? irl + imax : 0; }
    } else if (p->v < 0.0) {
        dist = -(p->y - g->ybounds[iry]) / p->v;
        if (dist < *ustep) { *ustep = dist; *irnew = (iry != 0) ? irl - imax : 0; }
    }
}

/* ---- a18: hownear(), omc_dosxyz.c:299-334 ---------------------------------------------------- */
static double hownear(const part *p) {
    const omc_geometry *g = &PB.G;
    int irl = p->ir;
    g_work.hownear++;
    if (irl == 0) return 0.0;
    int imax = g->isize, ijmax = g->isize * g->jsize;
    int irx = (irl - 1) % imax;
    int irz = (irl - 1 - irx) / ijmax;
    int iry = ((irl - 1 - irx) - irz * ijmax) / imax;
    double t = 1.0E10;
    t = fmin(t, g->xbounds[irx + 1] - p->x); t = fmin(t, p->x - g->xbounds[irx]);
    t = fmin(t, g->ybounds[iry + 1] - p->y); t = fmin(t, p->y - g->ybounds[iry]);
    t = fmin(t, g->zbounds[irz + 1] - p->z); t = fmin(t, p->z - g->zbounds[irz]);
    return t;
}

/* ---- a15: azimuth + rotations, src/ompmc.c:101-199 ------------------------------------------ */
static void azimuth(hist_ctx *c, double *cphi, double *sphi) {
    double x, x2, y, y2, r2;
    do {
        x = rnd(c); x = 2.0 * x - 1.0; x2 = x * x;
        y = rnd(c); y2 = y * y;
        r2 = x2 + y2;
    } while (r2 > 1.0);
    r2 = 1 / r2;
    *cphi = (x2 - y2) * r2;
    *sphi = 2.0 * x * y * r2;
}
typedef struct frame { double A, B, C, cphi, sphi; } frame;
/* common tail of uphi21/uphi32: rotate (costhe,sinthe,phi) from the saved frame into the lab */
static void frame_apply(const frame *f, double costhe, double sinthe, part *p) {
    double sinps2 = f->A * f->A + f->B * f->B;
    if (sinps2 < 1.0E-20) {
        p->u = sinthe * f->cphi; p->v = sinthe * f->sphi; p->w = f->C * costhe;
    } else {
        double sinpsi = sqrt(sinps2);
        double us = sinthe * f->cphi, vs = sinthe * f->sphi;
        double sindel = f->B / sinpsi, cosdel = f->A / sinpsi;
        p->u = f->C * cosdel * us - sindel * vs + f->A * costhe;
        p->v = f->C * sindel * us + cosdel * vs + f->B * costhe;
        p->w = -sinpsi * us + f->C * costhe;
    }
}
/* uphi21(), src/ompmc.c:127-164: new azimuth, frame = current direction of p, rotate p */
static void uphi21(hist_ctx *c, frame *f, double costhe, double sinthe, part *p) {
    azimuth(c, &f->cphi, &f->sphi);
    f->A = p->u; f->B = p->v; f->C = p->w;
    frame_apply(f, costhe, sinthe, p);
}
/* uphi32(), src/ompmc.c:166-199: q inherits position/region/weight of prev, direction from frame */
static void uphi32(const frame *f, double costhe, double sinthe, part *q, const part *prev) {
    q->x = prev->x; q->y = prev->y; q->z = prev->z; q->ir = prev->ir; q->wt = prev->wt;
    frame_apply(f, costhe, sinthe, q);
}

/* ---- a4: rayleigh(), src/ompmc.c:1102-1145 (Q2: ibin == 0, Q3: medium-0 form factor) -------- */
static void rayleigh(hist_ctx *c, int imed, double eig, double gle, int lgle) {
    const omc_media_tables *T = &PB.T;
    part *p = &c->stk[c->np];
    double pmax = pwlf_eval(imed * MXGE + lgle, gle, T->ray_pmax1, T->ray_pmax0);
    double xmax = HC_INVERSE * eig;
    double dwi = (double)OMC_MXRAYFF - 1.0;
    double r0, r1, xv, costhe, csqthe;
    do {
        r1 = rnd(c);
        do {
            r0 = rnd(c); r0 *= pmax;
            int ibin = (int)r0 * dwi;                     /* Q2: cast binds before the product */
            int ib = T->ray_i_array[ibin] - 1;            /* Q3: no imed*MXRAYFF offset */
            if ((T->ray_i_array[ibin + 1] - 1) > ib)
                while (r0 >= T->ray_fcum[ib + 1]) ib++;
            r0 = (r0 - T->ray_fcum[ib]) * T->ray_c_array[ib];
            xv = T->ray_xgrid[ib] * exp(log(1.0 + r0) * T->ray_b_array[ib]);
        } while (xv >= xmax);
        xv /= eig;
        costhe = 1.0 - TWICE_HC2 * (xv * xv);
        csqthe = costhe * costhe;
    } while (2.0 * r1 >= (1.0 + csqthe));
    double sinthe = sqrt(1.0 - csqthe);
    frame f;
    uphi21(c, &f, costhe, sinthe, p);
}

/* ---- a5: pair(), src/ompmc.c:1418-1667 -------------------------------------------------------- */
static double pair_rej(int imed, double xi, double esedei, double eseder, double tteig) {
    double a = (1.0 + eseder) * (1.0 + esedei) / (2.0 * tteig);
    double xh = xi - 0.5;
    return 2.0 + 3.0 * (esedei + eseder) -
           4.0 * (esedei + eseder + 1.0 - 4.0 * (xh * xh)) * (1.0 + 0.25 * log((a * a) + PB.T.zbrang[imed] * (xi * xi)));
}
static void pair(hist_ctx *c, int imed) {
    const omc_media_tables *T = &PB.T;
    int np = c->np;
    part *s = c->stk;
    double eig = s[np].e, ese1, ese2;
    int iq1, iq2, l, l1;
    c->npold = np;
    if (eig <= 2.1) {
        double r0 = rnd(c), r1 = rnd(c);
        ese2 = RM + 0.5 * r0 * (eig - 2.0 * RM);
        ese1 = eig - ese2;
        if (r1 < 0.5) { iq1 = -1; iq2 = 1; } else { iq1 = 1; iq2 = -1; }
    } else {
        double Amax, Bmax, delta, aux;
        if (eig < 50.0) {
            l = 4; l1 = l + 1;
            delta = 4.0 * T->delcm[imed] / eig;
            if (delta < 1.0) {
                Amax = T->dl1[imed * 8 + l] + delta * (T->dl2[imed * 8 + l] + delta * T->dl3[imed * 8 + l]);
                Bmax = T->dl1[imed * 8 + l1] + delta * (T->dl2[imed * 8 + l1] + delta * T->dl3[imed * 8 + l1]);
            } else {
                aux = log(delta + T->dl6[imed * 8 + l]);
                Amax = T->dl4[imed * 8 + l] + T->dl5[imed * 8 + l] * aux;
                Bmax = T->dl4[imed * 8 + l1] + T->dl5[imed * 8 + l1] * aux;
            }
            aux = 1.0 - 2.0 * RM / eig;
            aux = aux * aux;
            aux *= Amax / 3.0;
            aux /= (Bmax + aux);
        } else {
            l = 6;
            Amax = T->dl1[imed * 8 + l];
            Bmax = T->dl1[imed * 8 + l + 1];
            aux = T->bpar1[imed] * (1.0 - T->bpar0[imed] * RM / eig);
        }
        double eavail = eig - 2.0 * RM, rejf, rejmax, r4;
        do {
            double br, r0 = rnd(c), r1 = rnd(c);
            r4 = rnd(c);
            if (r0 > aux) {
                br = 0.5 * r1; rejmax = Bmax; l1 = l + 1;
            } else {
                double r2 = rnd(c), r3 = rnd(c);
                br = 0.5 * (1.0 - fmax(fmax(r1, r2), r3)); rejmax = Amax; l1 = l;
            }
            ese2 = br * eavail + RM;
            ese1 = eig - ese2;
            delta = (eig * T->delcm[imed]) / (ese2 * ese1);
            if (delta < 1.0)
                rejf = T->dl1[imed * 8 + l1] + delta * (T->dl2[imed * 8 + l1] + delta * T->dl3[imed * 8 + l1]);
            else
                rejf = T->dl4[imed * 8 + l1] + T->dl5[imed * 8 + l1] * log(delta + T->dl6[imed * 8 + l1]);
        } while (r4 * rejmax > rejf);
        ese1 = eig - ese2;
        r4 = rnd(c);
        if (r4 < 0.5) { iq1 = -1; iq2 = 1; } else { iq1 = 1; iq2 = -1; }
    }
    s[np].e = ese1;
    s[np + 1].e = ese2;
    /* angles: Motz-Olsen-Koch leading term (iprdst = 2), src/ompmc.c:1573-1658 */
    frame f;
    for (int j = 0; j < 2; j++) {
        double ese = (j == 0) ? ese1 : ese2;
        double tteig = eig / RM, ttese = ese / RM;
        double esedei = ttese / (tteig - ttese), eseder = 1.0 / esedei;
        double pt = M_PI * ttese;
        double ximin = 1.0 / (1.0 + (pt * pt));
        double rejmin = pair_rej(imed, ximin, esedei, eseder, tteig);
        double y2 = 2.0 / tteig;
        double ya = y2 * y2;
        double zb = T->zbrang[imed];
        double xitry = fmax(0.01, fmax(ximin, fmin(0.5, sqrt(ya / zb))));
        double galpha = 1.0 + 0.25 * log(ya + zb * xitry * xitry);
        double gbeta = 0.5 * zb * xitry / (ya + zb * xitry * xitry);
        galpha -= gbeta * (xitry - 0.5);
        double ximid = galpha / (3.0 * gbeta);
        if (galpha >= 0.0) ximid = 0.5 - ximid + sqrt((ximid * ximid) + 0.25);
        else ximid = 0.5 - ximid - sqrt((ximid * ximid) + 0.25);
        ximid = fmax(0.01, fmax(ximin, fmin(0.5, ximid)));
        double rejmid = pair_rej(imed, ximid, esedei, eseder, tteig);
        double rejtop = 1.0 * fmax(rejmin, rejmid);
        double theta, rtest, rejfactor;
        do {
            double xitst = rnd(c);
            double rejtst = pair_rej(imed, xitst, esedei, eseder, tteig);
            rtest = rnd(c);
            theta = sqrt(1.0 / xitst - 1.0) / ttese;
            rejfactor = rejtst / rejtop;
        } while ((rtest > rejfactor) && (theta >= M_PI));      /* Q4: '&&' */
        double sinthe = sin(theta), costhe = cos(theta);
        if (j == 0) {
            uphi21(c, &f, costhe, sinthe, &s[np]);
        } else {
            sinthe = -sinthe;
            np += 1; c->np = np;
            uphi32(&f, costhe, sinthe, &s[np], &s[np - 1]);
        }
    }
    s[np].iq = iq2;
    s[np - 1].iq = iq1;
}

/* ---- a6: compton(), src/ompmc.c:1670-1783 ----------------------------------------------------- */
static void compton(hist_ctx *c) {
    int np = c->np;
    part *s = c->stk;
    double eig = s[np].e, ko = s[np].e / RM;
    double broi = 1.0 + 2.0 * ko, bro = 1.0 / broi;
    int first_time = 1;
    double sinthe = 0.0, costhe = 0.0, br, r1, r2, r3, aux, rejf3, temp;
    double alph1 = 0.0, alph2 = 0.0, alpha = 0.0, rejmax = 0.0;
    c->npold = np;
    do {
        if (ko > 2.0) {
            if (first_time) {
                alph1 = log(broi);
                alph2 = ko * (broi + 1.0) * (bro * bro);
                alpha = alph1 + alph2;
            }
            do {
                r1 = rnd(c); r2 = rnd(c);
                if (r1 * alpha < alph1) br = exp(alph1 * r2) * bro;
                else br = sqrt(r2 * (broi * broi) + (1.0 - r2)) * bro;
                temp = (1.0 - br) / (ko * br);
                sinthe = fmax(0.0, temp * (2.0 - temp));
                aux = 1.0 + (br * br);
                rejf3 = aux - br * sinthe;
                r3 = rnd(c);
            } while (r3 * aux > rejf3);
        } else {
            if (first_time) rejmax = broi + bro;
            do {
                r1 = rnd(c); r2 = rnd(c);
                br = bro + (1.0 - bro) * r1;
                temp = (1.0 - br) / (ko * br);
                sinthe = fmax(0.0, temp * (2.0 - temp));
                rejf3 = 1.0 + br * br - br * sinthe;
            } while (r2 * br * rejmax > rejf3);
        }
        first_time = 0;
    } while ((br < bro) || (br > 1));
    costhe = 1.0 - temp;
    sinthe = sqrt(sinthe);
    double esg = br * eig, ese = eig - esg + RM;
    s[np].e = esg;
    frame f;
    uphi21(c, &f, costhe, sinthe, &s[np]);
    np += 1; c->np = np;
    aux = 1.0 + br * br - 2.0 * br * costhe;
    if (aux > 1.0E-8) {
        costhe = (1.0 - br * costhe) / sqrt(aux);
        sinthe = (1.0 - costhe) * (1.0 + costhe);
        sinthe = (sinthe > 0.0) ? -sqrt(sinthe) : 0.0;
    } else {
        costhe = 0.0; sinthe = -1.0;
    }
    uphi32(&f, costhe, sinthe, &s[np], &s[np - 1]);
    s[np].e = ese;
    s[np].iq = -1;
}

/* ---- a7: photo(), src/ompmc.c:1786-1846 ------------------------------------------------------- */
static void photo(hist_ctx *c) {
    int np = c->np;
    part *p = &c->stk[np];
    c->npold = np;
    p->e += RM;
    p->iq = -1;
    double eelec = p->e;
    if (eelec > PB.G.ecut[p->ir]) {
        double beta = sqrt((eelec - RM) * (eelec + RM)) / eelec;
        double costhe, sinth2, gamma = eelec / RM;
        double alpha = 0.5 * gamma - 0.5 + 1.0 / gamma;
        double ratio = beta / alpha, rn2, xi;
        do {
            double rn = rnd(c);
            rn = 2.0 * rn - 1.0;
            if (ratio <= 0.2) {
                double fkappa = rn + 0.5 * ratio * (1.0 - rn) * (1.0 + rn);
                if (gamma < 100.0) costhe = (beta + fkappa) / (1.0 + beta * fkappa);
                else if (fkappa > 0.0)
                    costhe = 1.0 - (1.0 - fkappa) * (gamma - 3.0) / (2.0 * (1.0 + fkappa) * pow((gamma - 1.0), 3.0));
                else costhe = (beta + fkappa) / (1.0 + beta * fkappa);
                xi = (1.0 + beta * fkappa) * (gamma * gamma);
            } else {
                xi = (gamma * gamma) * (1.0 + alpha * (sqrt(1.0f + ratio * (2.0 * rn + ratio)) - 1.0));
                costhe = (1.0 - 1.0 / xi) / beta;
            }
            sinth2 = fmax((1.0 - costhe) * (1.0 + costhe), 0.0);
            rn2 = rnd(c);
        } while (rn2 > 0.5 * (1.0 + gamma) * sinth2 * xi / gamma);
        double sinthe = sqrt(sinth2);
        frame f;
        uphi21(c, &f, costhe, sinthe, p);
    }
}

/* ---- a3: photon(), src/ompmc.c:1849-2121 ------------------------------------------------------ */
static void photon(hist_ctx *c) {
    const omc_media_tables *T = &PB.T;
    const omc_geometry *G = &PB.G;
    part *s = c->stk;
    int np = c->np, irl = s[np].ir, irold, irnew, imed = G->med[irl], idisc;
    double rhof, tstep, ustep, edep, eig = s[np].e, dpmfp, dpmfp_old, gmfpr0 = 0.0, gmfp = 0.0, gbr1, gbr2, r;
    int nsplit = PB.nsplit, i_survive, ip;
    double d_eta, eta_prime;
    part save;

    if (eig <= G->pcut[irl] || s[np].wt == 0) {            /* :1884-1891 */
        deposit(c, &s[np], eig);
        c->np -= 1;
        return;
    }
    int lgle = 0;
    double gle = log(eig), cohfac = 0.0;

    for (;;) {                                                /* start_mfp_loop, :1903 */
        r = rnd(c);
        r /= (double)nsplit;
        d_eta = 1.0 / (double)nsplit;
        eta_prime = 1.0 - r + d_eta;
        save = s[np];
        save.wt = s[np].wt / (double)nsplit;
        save.iq = 0;
        np -= 1;
        r = rnd(c);
        i_survive = (int)(r * nsplit);
        dpmfp_old = 0.0;

        for (int isplit = 0; isplit < nsplit; isplit++) {   /* :1924 */
            int ptrans = 1;
            eta_prime -= d_eta;
            if (eta_prime <= 0.0) break;
            dpmfp = -log(eta_prime) - dpmfp_old;
            dpmfp_old += dpmfp;
            np += 1; c->np = np;
            if (np >= MXSTACK) { c->flags |= 1u; fprintf(stderr, "oracle: stack overflow\n"); exit(1); }
            s[np] = save;
            irl = s[np].ir; irold = irl; imed = G->med[irl];
            int left = 0;
            do {                                              /* voxel-to-voxel march, :1951-2019 */
                if (imed != -1) {
                    lgle = pwlf_interval(imed, gle, T->ge1, T->ge0) - 1;
                    gmfpr0 = pwlf_eval(imed * MXGE + lgle, gle, T->gmfp1, T->gmfp0);
                    rhof = G->rhof[irl];
                    gmfp = gmfpr0 / rhof;
                    cohfac = pwlf_eval(imed * MXGE + lgle, gle, T->cohe1, T->cohe0);
                    gmfp *= cohfac;
                    tstep = gmfp * dpmfp;
                } else {
                    tstep = 1.0E8;
                }
                irnew = irl; idisc = 0; ustep = tstep;
                howfar(&s[np], &idisc, &irnew, &ustep);
                edep = 0.0;
                s[np].x += ustep * s[np].u; s[np].y += ustep * s[np].v; s[np].z += ustep * s[np].w;
                if (idisc > 0) {
                    np -= 1; c->np = np;
                    if (np < 0) return;
                    left = 1;
                    break;
                }
                if (imed != -1) dpmfp = fmax(0.0, dpmfp - ustep / gmfp);
                if (irnew != irold) { s[np].ir = irnew; irl = irnew; irold = irnew; imed = G->med[irl]; }
                if (imed != -1 && dpmfp <= SGMFP) ptrans = 0;
            } while (ptrans);
            if (left) break;                                  /* goto end_mfp_loop */

            save.x = s[np].x; save.y = s[np].y; save.z = s[np].z; save.ir = s[np].ir;

            r = rnd(c);                                       /* Rayleigh? :2027-2040 */
            if (r <= 1.0 - cohfac) {
                if (isplit != i_survive) { np -= 1; c->np = np; continue; }
                s[np].wt *= nsplit;
                rayleigh(c, imed, eig, gle, lgle);
                continue;
            }
            r = rnd(c);
            gbr1 = pwlf_eval(imed * MXGE + lgle, gle, T->gbr11, T->gbr10);
            if (r <= gbr1 && eig > 2.0 * RM) {
                pair(c, imed); np = c->np;
            } else {
                gbr2 = pwlf_eval(imed * MXGE + lgle, gle, T->gbr21, T->gbr20);
                if (r < gbr2) { compton(c); np = c->np; }
                else { photo(c); np = c->np; }
            }
            /* keep scattered photons of the chosen copy only, :2072-2093 */
            ip = c->npold;
            do {
                if (s[ip].iq == 0) {
                    if (isplit != i_survive) {
                        if (ip < np) {
                            s[ip].e = s[np].e; s[ip].iq = s[np].iq; s[ip].u = s[np].u; s[ip].v = s[np].v;
                            s[ip].w = s[np].w; s[ip].wt = s[np].wt;
                        }
                        np -= 1;
                    } else {
                        s[ip].wt *= nsplit;
                        ip += 1;
                    }
                } else {
                    ip += 1;
                }
            } while (ip <= np);
            c->np = np;
        }
        /* end_mfp_loop, :2095-2118 */
        if (np < 0) return;
        if (s[np].iq != 0) return;
        eig = s[np].e; irl = s[np].ir; imed = G->med[irl];
        if (eig <= G->pcut[irl]) {
            deposit(c, &s[np], eig);
            np -= 1; c->np = np;
            return;
        }
        gle = log(eig);
    }
}

/* ---- a11: spinRejection(), src/ompmc.c:3097-3168 --------------------------------------------- */
typedef struct spin_state { int i, j; } spin_state;
static double spin_rejection(hist_ctx *c, int imed, int qel, double elke, double beta2, double q1, double cost,
                             int *spin_index, int is_single, spin_state *sr) {
    const omc_media_tables *T = &PB.T;
    double ai, aj, ak, qq1, xi, r;
    g_work.spin++;
    if (*spin_index) {
        *spin_index = 0;
        if (beta2 >= T->b2spin_min) {
            ai = (beta2 - T->b2spin_min) * T->dbeta2i;
            sr->i = (int)ai; ai -= (double)sr->i; sr->i += MXE_SPIN + 1;
        } else if (elke > T->espml) {
            ai = (elke - T->espml) * T->dleneri;
            sr->i = (int)ai; ai -= sr->i;
        } else {
            sr->i = 0; ai = -1.0f;
        }
        r = rnd(c);
        if (r < ai) sr->i += 1;
        if (is_single) {
            sr->j = 0;
        } else {
            qq1 = 2.0 * q1;
            qq1 = qq1 / (1.0 + qq1);
            aj = qq1 * T->dqq1i;
            sr->j = (int)aj;
            if (sr->j >= MXQ_SPIN) {
                sr->j = MXQ_SPIN;
            } else {
                aj -= (double)sr->j;
                r = rnd(c);
                if (r < aj) sr->j += 1;
            }
        }
    }
    xi = sqrt(0.5 * (1.0 - cost));
    ak = xi * MXU_SPIN;
    int k = (int)ak;
    ak -= (double)k;
    const double *row = T->spin_rej + (size_t)imed * 2 * OMC_SPIN_NE * OMC_SPIN_NQ * OMC_SPIN_NU +
                        (size_t)qel * OMC_SPIN_NE * OMC_SPIN_NQ * OMC_SPIN_NU + (size_t)sr->i * OMC_SPIN_NQ * OMC_SPIN_NU +
                        (size_t)sr->j * OMC_SPIN_NU;
    return (1.0 - ak) * row[k] + ak * row[k + 1];
}

/* ---- a12: sscat(), src/ompmc.c:3170-3199 ------------------------------------------------------ */
static void sscat(hist_ctx *c, int imed, int qel, double chia2, double elke, double beta2, double *cost, double *sint) {
    int spin_index = 1;
    spin_state sr;
    double xi, rejf, r;
    do {
        xi = rnd(c);
        xi = 2.0 * chia2 * xi / (1.0 - xi + chia2);
        *cost = 1.0 - xi;
        rejf = spin_rejection(c, imed, qel, elke, beta2, 0, *cost, &spin_index, 1, &sr);
        r = rnd(c);
    } while (r > rejf);
    *sint = sqrt(xi * (2.0 - xi));
}

/* ---- a10: mscat(), src/ompmc.c:3606-3785 (Q1: du == 0, its draw still consumed; Q7) ----------- */
typedef struct ms_state { int i, j; double omega2; } ms_state;
static void mscat(hist_ctx *c, int imed, int qel, int *spin_index, int *find_index, double elke, double beta2,
                  double q1, double lambda, double chia2, double *cost, double *sint, ms_state *ms, spin_state *sr) {
    const omc_media_tables *T = &PB.T;
    double xi, rejf, r;
    g_work.mscat++;
    double explambda = exp(-lambda);
    if (lambda <= 13.8) {
        double sprob = rnd(c);
        if (sprob < explambda) { *cost = 1.0; *sint = 0.0; return; }
        double wsum = (1.0 + lambda) * explambda;
        if (sprob < wsum) {
            do {
                xi = rnd(c);
                xi = 2.0 * chia2 * xi / (1.0 - xi + chia2);
                *cost = 1.0 - xi;
                rejf = spin_rejection(c, imed, qel, elke, beta2, q1, *cost, spin_index, 0, sr);
                r = rnd(c);
            } while (r > rejf);
            *sint = sqrt(xi * (2.0 - xi));
            return;
        }
        if (lambda <= 1) {
            int icount = 0;
            double wprob = explambda, sinz, cosz, phi;
            wsum = explambda;
            *cost = 1.0; *sint = 0.0;
            do {
                icount += 1;
                if (icount > 20) break;
                wprob = wprob * lambda / icount;
                wsum = wsum + wprob;
                do {
                    xi = rnd(c);
                    xi = 2.0 * chia2 * xi / (1.0 - xi + chia2);
                    cosz = 1.0 - xi;
                    rejf = spin_rejection(c, imed, qel, elke, beta2, q1, cosz, spin_index, 0, sr);
                    r = rnd(c);
                } while (r > rejf);
                sinz = xi * (2.0 - xi);
                if (sinz > 1.0E-20) {
                    sinz = sqrt(sinz);
                    xi = rnd(c);
                    phi = xi * 6.2831853;
                    *cost = (*cost) * cosz - *sint * sinz * cos(phi);
                    *sint = sqrt(fmax(0.0, (double)((1.0 - (*cost)) * (1.0 + (*cost)))));
                }
            } while (wsum <= sprob);
            return;
        }
    }
    if (lambda <= LAMBMAX_MS) {
        double ai, aj, llmbda = log(lambda);
        if (*find_index) {
            ai = llmbda * T->dllambi;
            ms->i = (int)ai; ai -= (double)ms->i;
            xi = rnd(c);
            if (xi < ai) ms->i += 1;
            if (q1 < QMIN_MS) {
                ms->j = 0;
            } else if (q1 < QMAX_MS) {
                aj = q1 * T->dqmsi;
                ms->j = (int)aj; aj -= (double)ms->j;
                xi = rnd(c);
                if (xi < aj) ms->j += 1;
            } else {
                ms->j = MXQ_MS;
            }
            if (llmbda < 2.2299)
                ms->omega2 = chia2 * (lambda + 4.0) *
                             (1.347006 + llmbda * (0.209364 - llmbda * (0.45525 - llmbda * (0.50142 - 0.081234 * llmbda))));
            else
                ms->omega2 = chia2 * (lambda + 4.0) * (-2.77164 + llmbda * (2.94874 - llmbda * (0.1535754 - llmbda * 0.00552888)));
            *find_index = 0;
        }
        int k;
        double a, ak, u, du, x1;
        const int base = ms->i * OMC_MS_NQ * OMC_MS_NU + ms->j * OMC_MS_NU;
        do {
            xi = rnd(c);
            ak = xi * MXU_MS;
            k = ak;
            ak -= k;
            if (ak > T->wms[base + k]) k = T->ims[base + k];
            a = T->fms[base + k];
            u = T->ums[base + k];
            du = T->ums[base + k] - u;                        /* Q1: identically zero */
            xi = rnd(c);
            if (fabs(a) < 0.2) {
                x1 = 0.5 * (1.0 - xi) * a;
                u += xi * du * (1.0 + x1 * (1.0 - xi * a));
            } else {
                u -= du / a * (1.0 - sqrt(1.0 + xi * a * (2.0 + a)));
            }
            xi = ms->omega2 * u / (1.0 + 0.5 * ms->omega2 - u);
            if (xi > 1.99999) xi = 1.99999;
            *cost = 1.0 - xi;
            rejf = spin_rejection(c, imed, qel, elke, beta2, q1, *cost, spin_index, 0, sr);
            r = rnd(c);
        } while (r > rejf);
        *sint = sqrt(xi * (2.0 - xi));
    }
}

/* ---- a9: msdist(), src/ompmc.c:3787-3976 (PRESTA-II) ------------------------------------------ */
static double msdist(hist_ctx *c, const part *p, int imed, int iq, double rhof, double de, double tustep, double eke,
                     double *xf, double *yf, double *zf, double *uf, double *vf, double *wf) {
    const omc_media_tables *T = &PB.T;
    int qel = (1 + iq) / 2;
    ms_state ms;
    spin_state sr;
    double blccc = T->blcc[imed], xcccc = T->xcc[imed];
    double e = eke - 0.5 * de;
    double tau = e / RM, tau2 = tau * tau;
    double epsilon = de / eke, epsilonp = de / e;
    e *= (1.0 - (epsilonp * epsilonp) * (6.0 + 10.0 * tau + 5.0 * tau2) / (24.0 * tau2 + 72.0 * tau + 48.0));
    double p2 = e * (e + 2.0 * RM);
    double beta2 = p2 / (p2 + (RM * RM));
    double chia2 = xcccc / (4.0 * p2 * blccc);
    double lambda = 0.5 * tustep * rhof * blccc / beta2;
    double temp2 = 0.166666 * (4.0 + tau * (6.0 + tau * (7.0 + tau * (4.0 + tau)))) * (epsilonp / ((tau + 1.0) * (tau + 2.0))) *
                   (epsilonp / ((tau + 1.0) * (tau + 2.0)));
    lambda *= (1.0 - temp2);
    double elke = log(e);
    int lelke = pwlf_interval(imed, elke, T->eke1, T->eke0) - 1;
    if (lelke < 0) { lelke = 0; elke = (1.0 - T->eke0[imed]) / T->eke1[imed]; }
    double etap, xi_corr, gamma;
    int ix = MXEKE * imed + lelke;
    if (qel == 0) {
        etap = pwlf_eval(ix, elke, T->etae_ms1, T->etae_ms0);
        xi_corr = pwlf_eval(ix, elke, T->q1ce_ms1, T->q1ce_ms0);
        gamma = pwlf_eval(ix, elke, T->q2ce_ms1, T->q2ce_ms0);
    } else {
        etap = pwlf_eval(ix, elke, T->etap_ms1, T->etap_ms0);
        xi_corr = pwlf_eval(ix, elke, T->q1cp_ms1, T->q1cp_ms0);
        gamma = pwlf_eval(ix, elke, T->q2cp_ms1, T->q2cp_ms0);
    }
    double ms_corr = pwlf_eval(ix, elke, T->blcce1, T->blcce0);
    chia2 *= etap;
    lambda /= (etap * (1.0 + chia2));
    lambda *= ms_corr;
    double chilog = log(1.0 + 1.0 / chia2);
    double q1 = 2.0 * chia2 * (chilog * (1.0 + chia2) - 1.0);
    gamma = 6.0 * chia2 * (1.0 + chia2) * (chilog * (1.0 + 2.0 * chia2) - 2.0) / q1 * gamma;
    double xi = q1 * lambda;
    int find_index = 1, spin_index = 1;
    double w1, sint1, cphi1, sphi1, w2, sint2, cphi2, sphi2;
    mscat(c, imed, qel, &spin_index, &find_index, elke, beta2, xi, lambda, chia2, &w1, &sint1, &ms, &sr);
    azimuth(c, &cphi1, &sphi1);
    mscat(c, imed, qel, &spin_index, &find_index, elke, beta2, xi, lambda, chia2, &w2, &sint2, &ms, &sr);
    azimuth(c, &cphi2, &sphi2);
    double u2 = sint2 * cphi2, v2 = sint2 * sphi2, u2p = w1 * u2 + sint1 * w2;
    double us = u2p * cphi1 - v2 * sphi1, vs = u2p * sphi1 + v2 * cphi1, ws = w1 * w2 - sint1 * u2;
    xi *= 2 * xi_corr;
    double eta = rnd(c);
    double eta1 = 0.5 * (1.0 - eta);
    double delta = 0.9082483 - (0.1020621 - 0.0263747 * gamma) * xi;
    double temp1 = 2.0 + tau;
    double temp = (2.0 + tau * temp1) / ((tau + 1.0) * temp1);
    temp -= (tau + 1.0) / ((tau + 2.0) * (chilog * (1.0 + chia2) - 1.0));
    temp *= epsilonp;
    temp1 = 1.0 - temp;
    delta += 0.40824829 * (epsilon * (tau + 1.0) / ((tau + 2.0) * (chilog * (1.0 + chia2) - 1.0) * (chilog * (1.0 + 2.0 * chia2) - 2.0)) -
                           0.25 * (temp * temp));
    double b = eta * delta, cc = eta * (1.0 - delta);
    double w1v2 = w1 * v2;
    double ut = b * sint1 * cphi1 + cc * (cphi1 * u2 - sphi1 * w1v2) + eta1 * us * temp1;
    double vt = b * sint1 * sphi1 + cc * (sphi1 * u2 + cphi1 * w1v2) + eta1 * vs * temp1;
    double wt = eta1 * (1.0 + temp) + b * w1 + cc * w2 + eta1 * ws * temp1;
    double ustep = tustep * sqrt(ut * ut + vt * vt + wt * wt);
    double u0 = p->u, v0 = p->v, w0 = p->w;
    double sint02 = u0 * u0 + v0 * v0;
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
    *xf = p->x + tustep * ut; *yf = p->y + tustep * vt; *zf = p->z + tustep * wt;
    *uf = us; *vf = vs; *wf = ws;
    return ustep;
}

/* ---- a13: computeDrange(), src/ompmc.c:3979-4014 ---------------------------------------------- */
static double drange(int imed, int iq, int lelke, double ekei, double ekef, double elkei, double elkef) {
    const omc_media_tables *T = &PB.T;
    double fedep = 1.0 - ekef / ekei;
    double elktmp = 0.5 * (elkei + elkef + 0.25 * fedep * fedep * (1.0 + fedep * (1.0 + 0.875 * fedep)));
    const double *d1 = (iq < 0) ? T->ededx1 : T->pdedx1, *d0 = (iq < 0) ? T->ededx0 : T->pdedx0;
    double dedxmid = pwlf_eval(MXEKE * imed + lelke, elktmp, d1, d0);
    dedxmid = 1.0 / dedxmid;
    double aux = d1[MXEKE * imed + lelke] * dedxmid;
    aux = aux * (1.0 + 2.0 * aux) * fedep * fedep / (6.0 * (2.0 - fedep) * (2.0 - fedep));
    return fedep * ekei * dedxmid * (1.0 + aux);
}

/* ---- a13: computeEloss(), src/ompmc.c:4016-4108 ----------------------------------------------- */
static double eloss(int imed, int iq, double rhof, double tustep, double range, double eke, double elke, int lelke) {
    const omc_media_tables *T = &PB.T;
    int qel = (1 + iq) / 2;
    const double *rep = T->range_ep + (size_t)qel * T->nmed * MXEKE + (size_t)imed * MXEKE;
    const double *d1 = (iq < 0) ? T->ededx1 : T->pdedx1, *d0 = (iq < 0) ? T->ededx0 : T->pdedx0;
    double aux, dedxmid, de, fedep;
    double tuss = range - rep[lelke] / rhof;
    if (tuss >= tustep) {
        dedxmid = pwlf_eval(imed * MXEKE + lelke, elke, d1, d0);
        aux = d1[imed * MXEKE + lelke] / dedxmid;
        de = dedxmid * tustep * rhof;
        fedep = de / eke;
        de *= (1.0 - 0.5 * fedep * aux * (1.0 - 0.333333 * fedep * (aux - 1.0 - 0.25 * fedep * (2.0 - aux * (4.0 - aux)))));
    } else {
        int lt = lelke;
        tuss = (range - tustep) * rhof;
        if (tuss <= 0) {
            de = eke - T->pegs_te[imed] * 0.99;
        } else {
            while (tuss < rep[lt]) lt -= 1;
            double elktmp = (lt + 2 - T->eke0[imed]) / T->eke1[imed];
            double eketmp = T->e_array[imed * MXEKE + lt + 1];
            tuss = (rep[lt + 1] - tuss) / rhof;
            dedxmid = pwlf_eval(MXEKE * imed + lt, elktmp, d1, d0);
            aux = d1[MXEKE * imed + lt] / dedxmid;
            de = dedxmid * tuss * rhof;
            fedep = de / eketmp;
            de *= (1.0 - 0.5 * fedep * aux * (1.0 - 0.333333 * fedep * (aux - 1.0 - 0.25 * fedep * (2.0 - aux * (4.0 - aux)))));
            de += eke - eketmp;
        }
    }
    return de;
}

/* photon Russian roulette after e+- interactions when nsplit > 1, e.g. src/ompmc.c:4150-4165 */
static void roulette_photons(hist_ctx *c) {
    int nsplit = PB.nsplit;
    if (nsplit > 1) {
        part *s = c->stk;
        for (int ip = c->npold; ip <= c->np; ip++) {
            if (s[ip].iq == 0) {
                double r = rnd(c);
                if (r * (double)nsplit > 1.0) { s[ip].wt = 0.0; s[ip].e = 0.0; }
                else s[ip].wt *= nsplit;
            }
        }
    }
}

/* ---- a14: rannih(), src/ompmc.c:4111-4167 (Q6: one unused draw) -------------------------------- */
static void rannih(hist_ctx *c) {
    part *s = c->stk;
    int np = c->np;
    c->npold = np;
    double r = rnd(c);
    double costhe = 2.0 * r - 1;
    double sinthe = sqrt(fmax(0.0, (1.0 - costhe) * (1.0 + costhe)));
    r = rnd(c);                                               /* Q6 */
    double cphi, sphi;
    azimuth(c, &cphi, &sphi);
    s[np].e = RM; s[np].iq = 0;
    s[np].u = sinthe * cphi; s[np].v = sinthe * sphi; s[np].w = costhe;
    np += 1;
    s[np] = s[np - 1];
    s[np].u = -1.0 * s[np - 1].u; s[np].v = -1.0 * s[np - 1].v; s[np].w = -1.0 * s[np - 1].w;
    c->np = np;
    roulette_photons(c);
}

/* ---- a14: brems(), src/ompmc.c:4170-4356 (ibr_nist = 0, KM-2BS angle) -------------------------- */
static void brems(hist_ctx *c) {
    const omc_media_tables *T = &PB.T;
    part *s = c->stk;
    int np = c->np, irl = s[np].ir, imed = PB.G.med[irl];
    double eie = s[np].e, phi1, phi2;
    c->npold = np;
    int l = (eie < 50.0) ? 1 : 3, l1 = l + 1;
    double ekin = eie - RM, brmin = T->pegs_ap[imed] / ekin, waux = -log(brmin);
    double a = s[np].u, b = s[np].v, cz = s[np].w, sinpsi, sindel = 0, cosdel = 0;
    sinpsi = a * a + b * b;
    if (sinpsi > 1.0E-20) { sinpsi = sqrt(sinpsi); sindel = b / sinpsi; cosdel = a / sinpsi; }
    double ztarg = T->zbrang[imed], tteie = eie / RM;
    double beta = sqrt((tteie - 1.0) * (tteie + 1.0)) / tteie;
    double y2max = 2.0 * beta * (1.0 + beta) * tteie * tteie, y2maxi = 1.0 / y2max;
    double z2max = y2max + 1.0, z2maxi = sqrt(z2max);
    double aux, br, delta, r6, r7, rejf, ese, esg;
    const double *dl1 = T->dl1 + imed * 8, *dl2 = T->dl2 + imed * 8, *dl3 = T->dl3 + imed * 8, *dl4 = T->dl4 + imed * 8,
                 *dl5 = T->dl5 + imed * 8, *dl6 = T->dl6 + imed * 8;
    do {
        r6 = rnd(c); r7 = rnd(c);
        br = brmin * exp(r6 * waux);
        esg = ekin * br;
        ese = eie - esg;
        delta = esg / eie / ese * T->delcm[imed];
        aux = ese / eie;
        if (delta < 1.0) {
            phi1 = dl1[l - 1] + delta * (dl2[l - 1] + delta * dl3[l - 1]);
            phi2 = dl1[l1 - 1] + delta * (dl2[l1 - 1] + delta * dl3[l1 - 1]);
        } else {
            phi1 = dl4[l - 1] + dl5[l - 1] * log(delta + dl6[l - 1]);
            phi2 = phi1;
        }
        rejf = (1.0 + (aux * aux)) * phi1 - 2.0 * aux * phi2 / 3.0;
    } while (r7 >= rejf);
    np += 1;
    s[np] = s[np - 1];
    s[np].e = esg; s[np].iq = 0;
    double y2tst, ttese = ese / RM, esedei = ttese / tteie, rejmax;
    double rjarg1 = 1.0 + esedei * esedei, rjarg2 = rjarg1 + 2.0 * esedei, rjarg3;
    double rtest = 1.0, rejtst = 0.0;
    aux = 2.0 * ese * tteie / esg;
    aux = aux * aux;
    double aux1 = aux * ztarg;
    if (aux1 > 10.0) rjarg3 = -log(T->zbrang[imed]) + (1.0 - aux1) / (aux1 * aux1);
    else rjarg3 = log(aux / (1.0 + aux1));
    rejmax = rjarg1 * rjarg3 - rjarg2;
    while (rtest >= rejtst) {
        y2tst = rnd(c);
        rtest = rnd(c);
        double aux3 = z2maxi / (y2tst + (1.0 - y2tst) * z2maxi);
        rtest = rtest * aux3 * rejmax;
        y2tst = (aux3 * aux3) - 1.0;
        double y2tst1 = esedei * y2tst / pow(aux3, 4.0);
        double aux4 = 16.0 * y2tst1 - rjarg2, aux5 = rjarg1 - 4.0 * y2tst1;
        if (rtest < aux4 + aux5 * rjarg3) break;
        double aux2 = log(aux / (1.0 + aux1 / pow(aux3, 4.0)));
        rejtst = aux4 + aux5 * aux2;
    }
    double costhe = 1.0 - 2.0 * y2tst * y2maxi;
    double sinthe = sqrt(fmax(0.0, (1.0 - (costhe * costhe))));
    double cphi, sphi;
    azimuth(c, &cphi, &sphi);
    if (sinpsi >= 1.0E-10) {
        double us = sinthe * cphi, vs = sinthe * sphi;
        s[np].u = cz * cosdel * us - sindel * vs + a * costhe;
        s[np].v = cz * sindel * us + cosdel * vs + b * costhe;
        s[np].w = cz * costhe - sinpsi * us;
    } else {
        s[np].u = sinthe * cphi; s[np].v = sinthe * sphi; s[np].w = cz * costhe;
    }
    s[np - 1].e = ese;
    c->np = np;
    roulette_photons(c);
}

/* ---- a14: moller(), src/ompmc.c:4359-4435 ------------------------------------------------------ */
static void moller(hist_ctx *c) {
    const omc_media_tables *T = &PB.T;
    part *s = c->stk;
    int np = c->np, imed = PB.G.med[s[np].ir];
    double eie = s[np].e, ekin = eie - RM;
    c->npold = np;
    if (ekin <= 2.0 * T->pegs_te[imed]) return;
    double t0 = ekin / RM, e0 = t0 + 1.0, extrae = eie - T->pegs_thmoll[imed];
    double g2 = (t0 * t0) / (e0 * e0), g3 = (2.0 * t0 + 1.0) / (e0 * e0);
    double br, gmax = (1.0 + 1.25 * g2), rejf4, r, r27, r28;
    do {
        r27 = rnd(c);
        br = T->pegs_te[imed] / (ekin - extrae * r27);
        r = br / (1.0 - br);
        r28 = rnd(c);
        rejf4 = (1.0 + g2 * (br * br) + r * (r - g3));
        r28 *= gmax;
    } while (r28 > rejf4);
    double ekse2 = br * ekin, ese1 = eie - ekse2, ese2 = ekse2 + RM;
    s[np].e = ese1;
    s[np + 1].e = ese2;
    double h1 = (eie + RM) / ekin;
    double costh = h1 * (ese1 - RM) / (ese1 + RM);
    double sinthe = sqrt(1.0 - costh), costhe = sqrt(costh);
    frame f;
    uphi21(c, &f, costhe, sinthe, &s[np]);
    np += 1; c->np = np;
    s[np].iq = -1;
    costh = h1 * (ese2 - RM) / (ese2 + RM);
    sinthe = -sqrt(1.0 - costh);
    costhe = sqrt(costh);
    uphi32(&f, costhe, sinthe, &s[np], &s[np - 1]);
}

/* ---- a14: bhabha(), src/ompmc.c:4438-4525 ------------------------------------------------------ */
static void bhabha(hist_ctx *c) {
    const omc_media_tables *T = &PB.T;
    part *s = c->stk;
    int np = c->np, imed = PB.G.med[s[np].ir];
    double eip = s[np].e, ekin = eip - RM, t0 = ekin / RM, e0 = t0 + 1.0;
    double yy = 1.0 / (t0 + 2.0), beta2 = ((e0 * e0) - 1.0) / (e0 * e0);
    double ep0 = T->pegs_te[imed] / ekin, ep0c = 1.0 - ep0, yp = 1.0 - 2.0 * yy;
    c->npold = np;
    double b4 = pow(yp, 3.0), b3 = b4 + (yp * yp), b2 = yp * (3.0 + (yy * yy)), b1 = 2.0 - (yy * yy);
    double r3, r4, br, rejf2;
    do {
        r3 = rnd(c);
        br = ep0 / (1.0 - ep0c * r3);
        r4 = rnd(c);
        rejf2 = (1.0 - beta2 * br * (b1 - br * (b2 - br * (b3 - br * b4))));
    } while (r4 > rejf2);
    if (br < 0.5) {
        s[np + 1].iq = -1;
    } else {
        s[np].iq = -1; s[np + 1].iq = 1;
        br = 1.0 - br;
    }
    br = fmax(br, 0.0);
    double ekse2 = br * ekin, ese1 = eip - ekse2, ese2 = ekse2 + RM;
    s[np].e = ese1; s[np + 1].e = ese2;
    double h1 = (eip + RM) / ekin;
    double costh = fmin(1.0, h1 * (ese1 - RM) / (ese1 + RM));
    double sinthe = sqrt(1.0 - costh), costhe = sqrt(costh);
    frame f;
    uphi21(c, &f, costhe, sinthe, &s[np]);
    np += 1; c->np = np;
    costh = h1 * (ese2 - RM) / (ese2 + RM);
    sinthe = -sqrt(1.0 - costh);
    costhe = sqrt(costh);
    uphi32(&f, costhe, sinthe, &s[np], &s[np - 1]);
}

/* ---- a14: annih(), src/ompmc.c:4528-4645 ------------------------------------------------------- */
static void annih(hist_ctx *c) {
    part *s = c->stk;
    int np = c->np;
    double avip = s[np].e + RM, a = avip / RM, g = a - 1.0, t = g - 1.0, p = sqrt(a * t);
    c->npold = np;
    double pot = p / t, ep0 = 1.0 / (a + p), wsamp = log((1.0 - ep0) / ep0);
    double aa = s[np].u, bb = s[np].v, cc = s[np].w;
    double sinpsi = (aa * aa) + (bb * bb), sindel = 0, cosdel = 0;
    if (sinpsi > 1.0E-20) { sinpsi = sqrt(sinpsi); sindel = bb / sinpsi; cosdel = aa / sinpsi; }
    double ep, rejf, r1, r2;
    do {
        r1 = rnd(c);
        ep = ep0 * exp(r1 * wsamp);
        r2 = rnd(c);
        double q = ep * a - 1.0;
        rejf = 1.0 - (q * q) / (ep * ((a * a) - 2.0));
    } while (r2 > rejf);
    double esg1 = avip * ep;
    s[np].e = esg1; s[np].iq = 0;
    double costhe = fmin(1.0, (esg1 - RM) * pot / esg1);
    double sinthe = sqrt(1.0 - (costhe * costhe));
    double sphi, cphi, us, vs;
    azimuth(c, &cphi, &sphi);
    if (sinpsi >= 1.0E-10) {
        us = sinthe * cphi; vs = sinthe * sphi;
        s[np].u = cc * cosdel * us - sindel * vs + aa * costhe;
        s[np].v = cc * sindel * us + cosdel * vs + bb * costhe;
        s[np].w = cc * costhe - sinpsi * us;
    } else {
        s[np].u = sinthe * cphi; s[np].v = sinthe * sphi; s[np].w = cc * costhe;
    }
    np += 1;
    double esg2 = avip - esg1;
    s[np] = s[np - 1];
    s[np].e = esg2; s[np].iq = 0;
    costhe = fmin(1.0, (esg2 - RM) * pot / esg2);
    sinthe = -sqrt(1.0 - (costhe * costhe));
    if (sinpsi >= 1.0E-10) {
        us = sinthe * cphi; vs = sinthe * sphi;
        s[np].u = cc * cosdel * us - sindel * vs + aa * costhe;
        s[np].v = cc * sindel * us + cosdel * vs + bb * costhe;
        s[np].w = cc * costhe - sinpsi * us;
    } else {
        s[np].u = sinthe * cphi; s[np].v = sinthe * sphi; s[np].w = cc * costhe;
    }
    c->np = np;
    roulette_photons(c);
}

/* cut-off exit shared by the four ECUT tests of electron(): deposit, e+ -> rannih(), else pop.
 * src/ompmc.c:4665-4687, 5061-5088, 5119-5142, 5258-5281, 5301-5324 */
static void electron_end(hist_ctx *c, double edep, double eie, int iq) {
    deposit(c, &c->stk[c->np], edep);
    if (iq > 0 && edep < eie) { rannih(c); return; }
    c->np -= 1;
}

/* ---- a8: electron(), src/ompmc.c:4648-5433 ---------------------------------------------------- */
static void electron(hist_ctx *c) {
    const omc_media_tables *T = &PB.T;
    const omc_geometry *G = &PB.G;
    part *s = c->stk;
    const int np = c->np;
    part *p = &s[np];
    int irl = p->ir, imed = G->med[irl];
    double rhof = G->rhof[irl], edep = 0.0;
    frame fr;
    double eie = p->e;
    const int iq = p->iq, qel = (1 + iq) / 2;
    int medold = imed;
    double r;
    const int nmed = T->nmed;

    if (eie <= G->ecut[irl]) { electron_end(c, p->e - RM, eie, iq); return; }

    double elke = 0.0;
    int lelke = 0;
    double sigratio = 0.0, rfict = 0.0;
    const double *sg1 = (iq < 0) ? T->esig1 : T->psig1, *sg0 = (iq < 0) ? T->esig0 : T->psig0;
    const double *dd1 = (iq < 0) ? T->ededx1 : T->pdedx1, *dd0 = (iq < 0) ? T->ededx0 : T->pdedx0;
    const double *et1 = (iq < 0) ? T->etae_ms1 : T->etap_ms1, *et0 = (iq < 0) ? T->etae_ms0 : T->etap_ms0;
    const double *rep = T->range_ep + (size_t)qel * nmed * MXEKE;

    do { /* tstep loop, :4694 */
        int compute_tstep = 1;
        double eke = eie - RM, demfp = 0.0, sigf, sig0 = 0.0, dedx0, ustep = 0.0;
        if (imed != -1) {
            r = rnd(c);
            if (r == 0.0) r = 1.0E-30;
            demfp = fmax(-log(r), EPSEMFP);
            elke = log(eke);
            lelke = pwlf_interval(imed, elke, T->eke1, T->eke0) - 1;
            if (T->sig_ismonotone[qel * nmed + imed]) {
                sig0 = pwlf_eval(imed * MXEKE + lelke, elke, sg1, sg0);
                dedx0 = pwlf_eval(imed * MXEKE + lelke, elke, dd1, dd0);
                sig0 /= dedx0;
            } else {
                sig0 = (iq < 0) ? T->esig_e[imed] : T->psig_e[imed];
            }
        }
        do { /* ustep loop, :4765 */
            int call_howfar = 0, do_single = 0, called_msdist = 0;
            double tstep = 0, tustep = 0, total_de = 0, ekef, ekei, elkei, tuss, range = 0, p2, beta2, etap, tvstep = 0, de = 0;
            double total_tstep = 0.0;   /* re-initialised every iteration, exactly as src/ompmc.c:4787 */
            double xf = 0, yf = 0, zf = 0, uf = 0, vf = 0, wf = 0;
            if (imed == -1) {
                tstep = 10.0E8; ustep = tstep; tustep = ustep; call_howfar = 1;
            } else {
                rhof = G->rhof[irl];
                if (sig0 <= 0.0) {
                    tstep = 10.0E8; sig0 = 1.0E-15;
                } else {
                    if (compute_tstep) {
                        total_de = demfp / sig0;
                        ekef = eke - total_de;
                        if (ekef <= T->e_array[imed * MXEKE + 0]) {
                            tstep = 10.0E8;
                        } else {
                            double elkef = log(ekef);
                            int lelkef = pwlf_interval(imed, elkef, T->eke1, T->eke0) - 1;
                            if (lelkef == lelke) {
                                tstep = drange(imed, iq, lelke, eke, ekef, elke, elkef);
                            } else {
                                ekei = T->e_array[imed * MXEKE + lelke];
                                elkei = (lelke + 1 - T->eke0[imed]) / T->eke1[imed];
                                tuss = drange(imed, iq, lelke, eke, ekei, elke, elkei);
                                ekei = T->e_array[imed * MXEKE + lelkef + 1];
                                elkei = ((lelkef + 2) - T->eke0[imed]) / T->eke1[imed];
                                tstep = drange(imed, iq, lelkef, ekei, ekef, elkei, elkef);
                                tstep += tuss + rep[imed * MXEKE + lelke] - rep[imed * MXEKE + lelkef + 1];
                            }
                        }
                        total_tstep = tstep;
                        compute_tstep = 0;
                    }
                    tstep = total_tstep / rhof;
                }
                dedx0 = pwlf_eval(imed * MXEKE + lelke, elke, dd1, dd0);
                double dedx = rhof * dedx0;
                double tmxs = pwlf_eval(imed * MXEKE + lelke, elke, T->tmxs1, T->tmxs0);
                tmxs /= rhof;
                ekei = T->e_array[imed * MXEKE + lelke];
                elkei = (lelke + 1 - T->eke0[imed]) / T->eke1[imed];
                range = drange(imed, iq, lelke, eke, ekei, elke, elkei);
                range += rep[imed * MXEKE + lelke];
                range /= rhof;
                tustep = fmin(fmin(tstep, tmxs), range);
                double tperp = hownear(p);
                double blccl = rhof * T->blcc[imed], xccl = rhof * T->xcc[imed];
                p2 = eke * (eke + 2.0 * RM);
                beta2 = p2 / (p2 + (RM * RM));
                etap = pwlf_eval(MXEKE * imed + lelke, elke, et1, et0);
                double ms_corr = pwlf_eval(MXEKE * imed + lelke, elke, T->blcce1, T->blcce0);
                blccl = blccl / etap / (1.0 + 0.25 * etap * xccl / blccl / p2) * ms_corr;
                double ssmfp = beta2 / blccl;
                double skindepth = SKIN_DEPTH_FOR_BCA * ssmfp;
                tustep = fmin(tustep, fmax(tperp, skindepth));
                if ((tustep <= tperp) && (tustep > skindepth)) {
                    call_howfar = 0; do_single = 0; called_msdist = 1;
                    de = eloss(imed, iq, rhof, tustep, range, eke, elke, lelke);
                    tvstep = tustep;
                    ustep = msdist(c, p, imed, iq, rhof, de, tustep, eke, &xf, &yf, &zf, &uf, &vf, &wf);
                } else {
                    called_msdist = 0;
                    r = rnd(c);
                    if (r < 1.0E-30) r = 1.0E-30;
                    double lambda = (-1.0) * log(1.0 - r);
                    double lambda_max = 0.5 * blccl * RM / dedx;
                    lambda_max *= (eke / RM + 1.0) * (eke / RM + 1.0) * (eke / RM + 1.0);
                    if (lambda >= 0.0 && lambda_max > 0.0) {
                        if (lambda < lambda_max) tuss = lambda * ssmfp * (1.0 - 0.5 * lambda / lambda_max);
                        else tuss = 0.5 * lambda * ssmfp;
                        if (tuss < tustep) { tustep = tuss; do_single = 1; }
                        else do_single = 0;
                    } else {                                  /* Q8: silent drop */
                        c->flags |= 2u;
                        c->np -= 1;
                        return;
                    }
                    ustep = tustep;
                    call_howfar = (ustep < tperp) ? 0 : 1;
                }
            }
            int irold = p->ir, irnew = p->ir, idisc = 0;
            if (call_howfar) howfar(p, &idisc, &irnew, &ustep);
            if (idisc > 0) {                                  /* :5061-5088 */
                edep = (iq > 0) ? p->e + RM : p->e - RM;
                electron_end(c, edep, eie, iq);
                return;
            }
            if (ustep < 0) ustep = 0.0;
            double vstep;
            if (ustep == 0.0f || imed == -1) {                /* :5097-5146 */
                if (ustep != 0.0f) {
                    vstep = ustep; tvstep = vstep;
                    p->x += p->u * vstep; p->y += p->v * vstep; p->z += p->w * vstep;
                }
                if (irnew != irold) { p->ir = irnew; irl = irnew; imed = G->med[irl]; }
                if (eie <= G->ecut[irl]) { electron_end(c, p->e - RM, eie, iq); return; }
                break;
            }
            vstep = ustep;
            if (call_howfar) {
                tvstep = vstep;
                if (tvstep != tustep) do_single = 0;
                de = eloss(imed, iq, rhof, tvstep, range, eke, elke, lelke);
            } else {
                tvstep = tustep;
                if (called_msdist == 0) de = eloss(imed, iq, rhof, tvstep, range, eke, elke, lelke);
            }
            edep = de;
            ekef = eke - de;
            double sinthe = 0, costhe = 0;
            if (called_msdist == 0) {
                if (do_single) {
                    double ekems = fmax(ekef, G->ecut[irl] - RM);
                    p2 = ekems * (ekems + 2.0 * RM);
                    beta2 = p2 / (p2 + (RM * RM));
                    double chia2 = T->xcc[imed] / (4.0 * T->blcc[imed] * p2);
                    double elkems = log(ekems);
                    int lelkems = pwlf_interval(imed, elkems, T->eke1, T->eke0) - 1;
                    etap = pwlf_eval(MXEKE * imed + lelkems, elkems, et1, et0);
                    chia2 *= etap;
                    sscat(c, imed, qel, chia2, elkems, beta2, &costhe, &sinthe);
                } else {
                    sinthe = 0.0f; costhe = 1.0f;
                }
                xf = p->x + p->u * vstep; yf = p->y + p->v * vstep; zf = p->z + p->w * vstep;
                if (do_single) uphi21(c, &fr, costhe, sinthe, p);
                uf = p->u; vf = p->v; wf = p->w;
            }
            deposit(c, p, edep);                              /* :5245 */
            p->x = xf; p->y = yf; p->z = zf; p->u = uf; p->v = vf; p->w = wf;
            irold = p->ir;
            eie -= edep;
            p->e = eie;
            if (irnew == irl && eie <= G->ecut[irl]) { electron_end(c, p->e - RM, eie, iq); return; }
            medold = imed;
            if (imed != -1) {
                eke = eie - RM;
                elke = log(eke);
                lelke = pwlf_interval(imed, elke, T->eke1, T->eke0) - 1;
            }
            if (irnew != irold) { p->ir = irnew; irl = irnew; imed = G->med[irl]; }
            if (eie <= G->ecut[irl]) { electron_end(c, p->e - RM, eie, iq); return; }
            if (imed != medold) break;
            demfp -= de * sig0;
            total_de -= de;
            total_tstep -= tvstep * rhof;
            if (total_tstep < 1.0E-9) demfp = 0.0;
        } while (demfp >= EPSEMFP);

        /* C 'continue' in a do-while jumps to the loop test, which then compares the stale
         * rfict/sigratio of the previous pass (initially 0 >= 0): src/ompmc.c:5347-5349, 5372 */
        if ((imed != medold) || (ustep == 0.0) || (imed == -1)) continue;
        sigf = pwlf_eval(imed * MXEKE + lelke, elke, sg1, sg0);
        dedx0 = pwlf_eval(imed * MXEKE + lelke, elke, dd1, dd0);
        sigf /= dedx0;
        sigratio = sigf / sig0;
        rfict = rnd(c);
    } while (rfict >= sigratio);

    /* discrete interaction, :5375-5429 */
    if (iq < 0) {
        double ebr1 = pwlf_eval(imed * MXEKE + lelke, elke, T->ebr11, T->ebr10);
        r = rnd(c);
        if (r <= ebr1) {
            brems(c);
        } else if (p->e <= T->pegs_thmoll[imed]) {
            if (ebr1 <= 0) return;
            brems(c);
        } else {
            moller(c);
        }
    } else {
        double pbr1 = pwlf_eval(imed * MXEKE + lelke, elke, T->pbr11, T->pbr10);
        r = rnd(c);
        if (r < pbr1) {
            brems(c);
        } else {
            double pbr2 = pwlf_eval(imed * MXEKE + lelke, elke, T->pbr21, T->pbr20);
            if (r < pbr2) bhabha(c);
            else annih(c);
        }
    }
}

/* ---- a2: shower(), src/ompmc.c:5436-5447 ------------------------------------------------------ */
static void shower(hist_ctx *c) {
    while (c->np >= 0) {
        if (c->stk[c->np].iq == 0) photon(c);
        else electron(c);
    }
}

/* ---- a21: initHistory(), omc_dosxyz.c:964-1068 ------------------------------------------------- */
static double init_history(hist_ctx *c) {
    const omc_source_dosxyz *S = &PB.S;
    const omc_geometry *G = &PB.G;
    part *p = &c->stk[0];
    c->np = 0;
    p->iq = S->charge;
    double ein;
    if (S->spectrum) {
        double r1 = rnd(c), r2 = rnd(c);
        int k = (int)fmin(S->deltak * r1, S->deltak - 1.0);
        ein = S->cdfinv1[k] + r2 * S->cdfinv2[k];
    } else {
        ein = S->energy;
    }
    p->e = (p->iq != 0) ? ein + RM : ein;
    double rxyz;
    if (S->xsize == 0.0 || S->ysize == 0.0) {
        p->x = S->xinl; p->y = S->yinl;
        rxyz = sqrt((S->ssd * S->ssd) + (p->x * p->x) + (p->y * p->y));
        p->w = S->ssd / rxyz;
    } else {
        double fw, r3;
        do {
            r3 = rnd(c); p->x = r3 * S->xsize + S->xinl;
            r3 = rnd(c); p->y = r3 * S->ysize + S->yinl;
            r3 = rnd(c);
            rxyz = sqrt(S->ssd * S->ssd + p->x * p->x + p->y * p->y);
            p->w = S->ssd / rxyz;
            fw = p->w * p->w * p->w;
        } while (r3 >= fw);
    }
    p->z = G->zbounds[0];
    p->u = p->x / rxyz; p->v = p->y / rxyz;
    int ix, iy;
    if (S->xsize == 0.0) ix = S->ixinl;
    else { ix = S->ixinl - 1; while ((G->xbounds[ix + 1] < p->x) && ix < G->isize - 1) ix++; }
    if (S->ysize == 0.0) iy = S->iyinl;
    else { iy = S->iyinl - 1; while ((G->ybounds[iy + 1] < p->y) && iy < G->jsize - 1) iy++; }
    p->ir = 1 + ix + iy * G->isize;
    p->wt = 1.0;
    return ein;
}

/* ---- a22: initHistory(ibeamlet), omc_matrad.c:1084-1254 (Q13: z clamp uses ybounds[0]; the 2*DBL_MIN
 * nudges are no-ops for non-zero bounds) ------------------------------------------------------------ */
static double init_history_matrad(hist_ctx *c, int ib) {
    const omc_source_matrad *S = &PB.M;
    const omc_geometry *G = &PB.G;
    part *p = &c->stk[0];
    const int imax = G->isize, ijmax = G->isize * G->jsize;
    c->np = 0;
    p->iq = S->charge;
    double ein;
    if (S->spectrum) {
        double r1 = rnd(c), r2 = rnd(c);
        int k = (int)fmin(S->deltak * r1, S->deltak - 1.0);
        ein = S->cdfinv1[k] + r2 * S->cdfinv2[k];
    } else {
        ein = S->energy;
    }
    p->e = (p->iq != 0) ? ein + RM : ein;
    double r1 = rnd(c), r2 = rnd(c);
    double xiso = r1 * S->xside1[ib] + r2 * S->xside2[ib] + S->xcorner[ib];
    double yiso = r1 * S->yside1[ib] + r2 * S->yside2[ib] + S->ycorner[ib];
    double ziso = r1 * S->zside1[ib] + r2 * S->zside2[ib] + S->zcorner[ib];
    int ibeam = S->ibeam[ib];
    double dx = xiso - S->xsource[ibeam], dy = yiso - S->ysource[ibeam], dz = ziso - S->zsource[ibeam];
    double vnorm = sqrt((dx * dx) + (dy * dy) + (dz * dz));
    double u = -(xiso - S->xsource[ibeam]) / vnorm, v = -(yiso - S->ysource[ibeam]) / vnorm, w = -(ziso - S->zsource[ibeam]) / vnorm;
    double ustep = 1.0E5, dist;
    if (u > 0.0) { dist = (G->xbounds[G->isize] - xiso) / u; if (dist < ustep) ustep = dist; }
    if (u < 0.0) { dist = -(xiso - G->xbounds[0]) / u; if (dist < ustep) ustep = dist; }
    if (v > 0.0) { dist = (G->ybounds[G->jsize] - yiso) / v; if (dist < ustep) ustep = dist; }
    if (v < 0.0) { dist = -(yiso - G->ybounds[0]) / v; if (dist < ustep) ustep = dist; }
    if (w > 0.0) { dist = (G->zbounds[G->ksize] - ziso) / w; if (dist < ustep) ustep = dist; }
    if (w < 0.0) { dist = -(ziso - G->zbounds[0]) / w; if (dist < ustep) ustep = dist; }
    p->x = xiso + ustep * u; p->y = yiso + ustep * v; p->z = ziso + ustep * w;
    p->u = -u; p->v = -v; p->w = -w;
    const double tiny = 2.0 * 2.2250738585072014e-308;
    if (p->x < G->xbounds[0]) p->x = G->xbounds[0] + tiny;
    if (p->x > G->xbounds[G->isize]) p->x = G->xbounds[G->isize] - tiny;
    if (p->y < G->ybounds[0]) p->y = G->ybounds[0] + tiny;
    if (p->y > G->ybounds[G->jsize]) p->y = G->ybounds[G->jsize] - tiny;
    if (p->z < G->zbounds[0]) p->z = G->ybounds[0] + tiny;                       /* Q13 */
    if (p->z > G->zbounds[G->ksize]) p->z = G->zbounds[G->ksize] - tiny;
    int ix = 0, iy = 0, iz = 0;
    while (G->xbounds[ix + 1] < p->x) ix++;
    while (G->ybounds[iy + 1] < p->y) iy++;
    while (G->zbounds[iz + 1] < p->z) iz++;
    p->ir = 1 + ix + iy * imax + iz * ijmax;
    p->wt = 1.0;
    return ein;
}

/* ------------------------------------------------------------------------------------------ */
/* API                                                                                         */
/* ------------------------------------------------------------------------------------------ */
static void ctx_setup(void) {
    int n = 1;
#ifdef _OPENMP
    n = omp_get_max_threads();
#endif
    if (g_ctx && g_nctx >= n) return;
    g_ctx = realloc(g_ctx, (size_t)n * sizeof(hist_ctx));
    for (int i = g_nctx; i < n; i++) {
        memset(&g_ctx[i], 0, sizeof(hist_ctx));
        g_ctx[i].stk = malloc((MXSTACK + 2) * sizeof(part));
        ranmar_init(&g_ctx[i].rm, (int)g_seed0, (int)g_seed1 + i);   /* jxx += thread id, src/omc_random.c:70-73 */
    }
    g_nctx = n;
}

int orc_load_problem(const char *path) {
    omc_blob *b = &g_blob;
    omc_blob_free(b);
    if (omc_blob_read(b, path) != 0) return -1;
    omc_media_tables *T = &PB.T;
    memset(&PB.T, 0, sizeof PB.T);
    T->nmed = omc_blob_i32(b, "nmed")[0];
#define F(field) T->field = omc_blob_f64(b, #field);
#define I(field) T->field = omc_blob_i32(b, #field);
    F(ge0) F(ge1) F(gmfp0) F(gmfp1) F(gbr10) F(gbr11) F(gbr20) F(gbr21) F(cohe0) F(cohe1)
    F(ray_xgrid) F(ray_fcum) F(ray_b_array) F(ray_c_array) I(ray_i_array) F(ray_pmax0) F(ray_pmax1)
    F(dl1) F(dl2) F(dl3) F(dl4) F(dl5) F(dl6) F(bpar0) F(bpar1) F(delcm) F(zbrang)
    F(esig0) F(esig1) F(psig0) F(psig1) F(ededx0) F(ededx1) F(pdedx0) F(pdedx1) F(ebr10) F(ebr11) F(pbr10) F(pbr11)
    F(pbr20) F(pbr21) F(tmxs0) F(tmxs1) F(blcce0) F(blcce1) F(etae_ms0) F(etae_ms1) F(etap_ms0) F(etap_ms1)
    F(q1ce_ms0) F(q1ce_ms1) F(q1cp_ms0) F(q1cp_ms1) F(q2ce_ms0) F(q2ce_ms1) F(q2cp_ms0) F(q2cp_ms1)
    F(range_ep) F(e_array) F(eke0) F(eke1) I(sig_ismonotone) F(esig_e) F(psig_e) F(xcc) F(blcc)
    F(spin_rej) F(ums) F(fms) F(wms) I(ims) F(pegs_ap) F(pegs_ae) F(pegs_te) F(pegs_thmoll) F(pegs_rho) I(pegs_meke)
#undef F
#undef I
    T->b2spin_min = omc_blob_f64(b, "b2spin_min")[0]; T->dbeta2i = omc_blob_f64(b, "dbeta2i")[0];
    T->espml = omc_blob_f64(b, "espml")[0];           T->dleneri = omc_blob_f64(b, "dleneri")[0];
    T->dqq1i = omc_blob_f64(b, "dqq1i")[0];
    T->dllambi = omc_blob_f64(b, "dllambi")[0];       T->dqmsi = omc_blob_f64(b, "dqmsi")[0];
    omc_geometry *G = &PB.G;
    G->isize = omc_blob_i32(b, "isize")[0]; G->jsize = omc_blob_i32(b, "jsize")[0]; G->ksize = omc_blob_i32(b, "ksize")[0];
    G->xbounds = omc_blob_f64(b, "xbounds"); G->ybounds = omc_blob_f64(b, "ybounds"); G->zbounds = omc_blob_f64(b, "zbounds");
    G->med = omc_blob_i32(b, "region_med"); G->rhof = omc_blob_f64(b, "region_rhof");
    G->pcut = omc_blob_f64(b, "region_pcut"); G->ecut = omc_blob_f64(b, "region_ecut");
    omc_source_dosxyz *S = &PB.S;
    S->spectrum = omc_blob_i32(b, "src_spectrum")[0]; S->charge = omc_blob_i32(b, "src_charge")[0];
    S->energy = omc_blob_f64(b, "src_energy")[0];     S->deltak = omc_blob_f64(b, "src_deltak")[0];
    S->cdfinv1 = omc_blob_f64(b, "src_cdfinv1");      S->cdfinv2 = omc_blob_f64(b, "src_cdfinv2");
    PB.is_matrad = 0; PB.ibeamlet = 0;
    for (int i = 0; i < b->n; i++) if (!strncmp(b->e[i].name, "mr_nbeamlets", 32)) PB.is_matrad = 1;
    if (PB.is_matrad) {
        omc_source_matrad *Mr = &PB.M;
        Mr->spectrum = S->spectrum; Mr->charge = S->charge; Mr->energy = S->energy; Mr->deltak = S->deltak;
        Mr->cdfinv1 = S->cdfinv1; Mr->cdfinv2 = S->cdfinv2;
        Mr->nbixels = omc_blob_i32(b, "mr_nbeamlets")[0];
        Mr->nbeams = (int)omc_blob_find(b, "mr_xsource")->count;
        Mr->ibeam = omc_blob_i32(b, "mr_ibeam");
        Mr->xsource = omc_blob_f64(b, "mr_xsource"); Mr->ysource = omc_blob_f64(b, "mr_ysource"); Mr->zsource = omc_blob_f64(b, "mr_zsource");
        Mr->xcorner = omc_blob_f64(b, "mr_xcorner"); Mr->ycorner = omc_blob_f64(b, "mr_ycorner"); Mr->zcorner = omc_blob_f64(b, "mr_zcorner");
        Mr->xside1 = omc_blob_f64(b, "mr_xside1"); Mr->yside1 = omc_blob_f64(b, "mr_yside1"); Mr->zside1 = omc_blob_f64(b, "mr_zside1");
        Mr->xside2 = omc_blob_f64(b, "mr_xside2"); Mr->yside2 = omc_blob_f64(b, "mr_yside2"); Mr->zside2 = omc_blob_f64(b, "mr_zside2");
    }
    if (!PB.is_matrad) {
    S->ssd = omc_blob_f64(b, "src_ssd")[0];
    S->xinl = omc_blob_f64(b, "src_xinl")[0]; S->xinu = omc_blob_f64(b, "src_xinu")[0];
    S->yinl = omc_blob_f64(b, "src_yinl")[0]; S->yinu = omc_blob_f64(b, "src_yinu")[0];
    S->xsize = omc_blob_f64(b, "src_xsize")[0]; S->ysize = omc_blob_f64(b, "src_ysize")[0];
    S->ixinl = omc_blob_i32(b, "src_ixinl")[0]; S->ixinu = omc_blob_i32(b, "src_ixinu")[0];
    S->iyinl = omc_blob_i32(b, "src_iyinl")[0]; S->iyinu = omc_blob_i32(b, "src_iyinu")[0];
    }
    PB.nsplit = omc_blob_i32(b, "nsplit")[0];
    PB.nreg = G->isize * G->jsize * G->ksize + 1;
    free(PB.endep); free(PB.accum); free(PB.accum2);
    PB.endep = calloc(PB.nreg, sizeof(double)); PB.accum = calloc(PB.nreg, sizeof(double)); PB.accum2 = calloc(PB.nreg, sizeof(double));
    PB.ensrc = 0.0;
    /* fresh RANMAR sequences for the new problem */
    for (int i = 0; i < g_nctx; i++) free(g_ctx[i].stk);
    free(g_ctx); g_ctx = NULL; g_nctx = 0;
    ctx_setup();
    return 0;
}

void orc_set_rng(int mode, int seed0, int seed1) { g_rng_mode = mode; g_seed0 = (uint32_t)seed0; g_seed1 = (uint32_t)seed1; }
void orc_set_nsplit(int nsplit) { PB.nsplit = nsplit; }
void orc_set_beamlet(int ibeamlet) { PB.ibeamlet = ibeamlet; }
int orc_nreg(void) { return PB.nreg; }
int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#endif
    ctx_setup();
}

static void begin_history(hist_ctx *c, long long id) {
    c->rng_mode = g_rng_mode;
    if (g_rng_mode == 1) omc_philox_seed(&c->ph, g_seed0, g_seed1, (uint64_t)id, 0);
    c->ndeposit = 0; c->flags = 0; c->edep_sum = 0.0;
}

/* {initHistory(); shower();} x n, omc_dosxyz.c:1252-1259 */
void orc_run_histories(long long first, long long n, omc_history_record *rec) {
    ctx_setup();
    double ensrc = 0.0;
    long long i;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : ensrc)
#endif
    for (i = 0; i < n; i++) {
        int tid = 0;
#ifdef _OPENMP
        tid = omp_get_thread_num();
#endif
        hist_ctx *c = &g_ctx[tid];
        begin_history(c, first + i);
        ensrc += PB.is_matrad ? init_history_matrad(c, PB.ibeamlet) : init_history(c);
        int ir0 = c->stk[0].ir;
        shower(c);
        if (rec) {
            rec[i].ndraws = (unsigned)c->ph.ndraws; rec[i].ir_start = ir0; rec[i].ndeposit = c->ndeposit;
            rec[i].flags = c->flags; rec[i].edep = c->edep_sum;
        }
    }
    PB.ensrc += ensrc;
#ifdef _OPENMP
#pragma omp parallel
#endif
    {
#ifdef _OPENMP
#pragma omp critical
#endif
        {
            g_work_total.ausgab += g_work.ausgab; g_work_total.howfar += g_work.howfar; g_work_total.hownear += g_work.hownear;
            g_work_total.pwlf += g_work.pwlf; g_work_total.mscat += g_work.mscat; g_work_total.spin += g_work.spin;
            memset(&g_work, 0, sizeof g_work);
        }
    }
    g_work_total.hist += (unsigned long long)n;
}

/* totals since the last orc_reset_score(): {ausgab, howfar, hownear, pwlfEval, mscat, spinRejection, histories} */
void orc_get_work(unsigned long long *out) {
    out[0] = g_work_total.ausgab; out[1] = g_work_total.howfar; out[2] = g_work_total.hownear; out[3] = g_work_total.pwlf;
    out[4] = g_work_total.mscat; out[5] = g_work_total.spin; out[6] = g_work_total.hist;
}
void orc_set_endep(const double *in) { memcpy(PB.endep, in, (size_t)PB.nreg * sizeof(double)); }

/* a20: accumEndep(), omc_dosxyz.c:696-717 */
void orc_accum_endep(void) {
    for (int i = 0; i < PB.nreg; i++) {
        double e = PB.endep[i];
        PB.accum[i] += e;
        PB.accum2[i] += e * e;
    }
    memset(PB.endep, 0, (size_t)PB.nreg * sizeof(double));
}
void orc_reset_score(void) {
    size_t n = (size_t)PB.nreg * sizeof(double);
    memset(PB.endep, 0, n); memset(PB.accum, 0, n); memset(PB.accum2, 0, n);
    PB.ensrc = 0.0;
    memset(&g_work_total, 0, sizeof g_work_total);
}
void orc_zero_accum(void) { memset(PB.accum, 0, (size_t)PB.nreg * sizeof(double)); }   /* omc_matrad.c:1482 */
void orc_get_endep(double *out) { memcpy(out, PB.endep, (size_t)PB.nreg * sizeof(double)); }
void orc_get_accum(double *a, double *a2, double *ensrc) {
    size_t n = (size_t)PB.nreg * sizeof(double);
    if (a) memcpy(a, PB.accum, n);
    if (a2) memcpy(a2, PB.accum2, n);
    if (ensrc) *ensrc = PB.ensrc;
}
double orc_time_batches(long long first, long long nperbatch, int nbatch) {
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int ib = 0; ib < nbatch; ib++) {
        orc_run_histories(first + ib * nperbatch, nperbatch, NULL);
        orc_accum_endep();
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

void orc_test_geometry(int n, const double *q, const int *ir, const double *ustep_in, int *idisc, int *irnew,
                       double *ustep_out, double *tperp) {
    for (int i = 0; i < n; i++) {
        part p = {0, ir[i], 1.0, q[6 * i], q[6 * i + 1], q[6 * i + 2], q[6 * i + 3], q[6 * i + 4], q[6 * i + 5], 1.0};
        int id = 0, irn = ir[i];
        double us = ustep_in[i];
        howfar(&p, &id, &irn, &us);
        idisc[i] = id; irnew[i] = irn; ustep_out[i] = us;
        tperp[i] = hownear(&p);
    }
}
void orc_test_rng(long long hist, int n, double *out) {
    omc_philox g;
    omc_philox_seed(&g, g_seed0, g_seed1, (uint64_t)hist, 0);
    for (int i = 0; i < n; i++) out[i] = omc_philox_next(&g);
}
void orc_test_ranmar(int ixx, int jxx, int n, double *out) {
    ranmar r;
    ranmar_init(&r, ixx, jxx);
    for (int i = 0; i < n; i++) {
        if (r.pos >= 128) ranmar_fill(&r);
        out[i] = r.buf[r.pos++] * (1.0 / 16777216.0);
    }
}
void orc_run_particle(long long hist, int iq, double e, const double *q, int ir, double wt, omc_history_record *rec) {
    ctx_setup();
    hist_ctx *c = &g_ctx[0];
    int mode = g_rng_mode;
    g_rng_mode = 1;
    begin_history(c, hist);
    g_rng_mode = mode;
    c->np = 0;
    part *p = &c->stk[0];
    p->iq = iq; p->e = e; p->ir = ir; p->wt = wt;
    p->x = q[0]; p->y = q[1]; p->z = q[2]; p->u = q[3]; p->v = q[4]; p->w = q[5];
    shower(c);
    if (rec) {
        rec->ndraws = (unsigned)c->ph.ndraws; rec->ir_start = ir; rec->ndeposit = c->ndeposit; rec->flags = c->flags;
        rec->edep = c->edep_sum;
    }
}

/* ---- unit hooks: the restated samplers on explicit inputs (same contract as ref_test_samplers() of ref_harness.c and
 * omc_gpu_test_samplers() of the C-ABI; draws one by one from the Philox stream of history first + i) -------------------- */
static double orc_range(int imed, int iq, double eke, double rhof, double *elke_out, int *lelke_out) {
    /* electron() src/ompmc.c:4905-4918 */
    const omc_media_tables *T = &PB.T;
    int qel = (1 + iq) / 2;
    double elke = log(eke);
    int lelke = pwlf_interval(imed, elke, T->eke1, T->eke0) - 1;
    double ekei = T->e_array[imed * MXEKE + lelke];
    double elkei = (lelke + 1 - T->eke0[imed]) / T->eke1[imed];
    double range = drange(imed, iq, lelke, eke, ekei, elke, elkei);
    range += T->range_ep[(size_t)qel * T->nmed * MXEKE + (size_t)imed * MXEKE + lelke];
    *elke_out = elke; *lelke_out = lelke;
    return range / rhof;
}

void orc_test_samplers(int which, int n, const double *in, long long first, double *out) {
    const omc_media_tables *T = &PB.T;
    ctx_setup();
    hist_ctx *c = &g_ctx[0];
    int mode = g_rng_mode;
    for (int i = 0; i < n; i++) {
        const double *a = in + 8 * (size_t)i;
        double *o = out + 8 * (size_t)i;
        for (int k = 0; k < 8; k++) o[k] = 0.0;
        g_rng_mode = 1;
        begin_history(c, first + i);
        g_rng_mode = mode;
        c->np = 0; c->npold = 0;
        part *p = &c->stk[0];
        p->x = p->y = p->z = 0.0; p->wt = 1.0; p->ir = 1; p->iq = 0; p->e = 1.0; p->u = p->v = 0.0; p->w = 1.0;
        if (which == OMC_SAMPLER_DRANGE) {
            int imed = (int)a[0], iq = (int)a[1];
            double elkei = log(a[2]), elkef = log(a[3]);
            int lelke = pwlf_interval(imed, elkei, T->eke1, T->eke0) - 1;
            o[0] = drange(imed, iq, lelke, a[2], a[3], elkei, elkef);
        } else if (which == OMC_SAMPLER_ELOSS || which == OMC_SAMPLER_MSDIST) {
            int imed = (int)a[0], iq = (int)a[1], lelke;
            double rhof = a[2], eke = a[3], elke;
            double range = orc_range(imed, iq, eke, rhof, &elke, &lelke);
            double tustep = a[4] * range;
            double de = eloss(imed, iq, rhof, tustep, range, eke, elke, lelke);
            if (which == OMC_SAMPLER_ELOSS) { o[0] = range; o[1] = de; continue; }
            p->iq = iq; p->e = eke + RM; p->u = a[5]; p->v = a[6]; p->w = a[7];
            o[0] = msdist(c, p, imed, iq, rhof, de, tustep, eke, &o[1], &o[2], &o[3], &o[4], &o[5], &o[6]);
            o[7] = de;
        } else if (which == OMC_SAMPLER_SSCAT) {
            sscat(c, (int)a[0], (int)a[1], a[2], a[3], a[4], &o[0], &o[1]);
            azimuth(c, &o[2], &o[3]);
        } else if (which == OMC_SAMPLER_COMPTON) {
            p->iq = 0; p->e = a[0]; p->u = a[1]; p->v = a[2]; p->w = a[3];
            compton(c);
            for (int k = 0; k <= c->np; k++) {
                int off = (c->stk[k].iq == 0) ? 0 : 4;
                o[off] = c->stk[k].e; o[off + 1] = c->stk[k].u; o[off + 2] = c->stk[k].v; o[off + 3] = c->stk[k].w;
            }
        } else if (which == OMC_SAMPLER_MOLLER) {
            int imed = (int)a[0], ir = -1;
            for (int r = 1; r < PB.nreg; r++) if (PB.G.med[r] == imed) { ir = r; break; }
            if (ir < 0) continue;
            p->ir = ir; p->iq = -1; p->e = a[1]; p->u = a[2]; p->v = a[3]; p->w = a[4];
            moller(c);
            int hi = 0, lo = -1;
            if (c->np == 1) { hi = (c->stk[0].e >= c->stk[1].e) ? 0 : 1; lo = 1 - hi; }
            o[0] = c->stk[hi].e; o[1] = c->stk[hi].u; o[2] = c->stk[hi].v; o[3] = c->stk[hi].w;
            if (lo >= 0) { o[4] = c->stk[lo].e; o[5] = c->stk[lo].u; o[6] = c->stk[lo].v; o[7] = c->stk[lo].w; }
        }
    }
}

/* optical depth of a straight photon path: the transport loop of photon() src/ompmc.c:1951-2019 without the mean-free-path
 * budget (see ref_test_photon_tau() in ref_harness.c).  in {e, x, y, z, u, v, w, s} -> out {tau, region at the end} */
void orc_test_photon_tau(int n, const double *in, double *out) {
    const omc_media_tables *T = &PB.T;
    const omc_geometry *G = &PB.G;
    for (int i = 0; i < n; i++) {
        const double *a = in + 8 * (size_t)i;
        double gle = log(a[0]), left = a[7], tau = 0.0;
        int ix = 0, iy = 0, iz = 0;
        while (ix < G->isize - 1 && G->xbounds[ix + 1] <= a[1]) ix++;
        while (iy < G->jsize - 1 && G->ybounds[iy + 1] <= a[2]) iy++;
        while (iz < G->ksize - 1 && G->zbounds[iz + 1] <= a[3]) iz++;
        part p = {0, 1 + ix + iy * G->isize + iz * G->isize * G->jsize, a[0], a[1], a[2], a[3], a[4], a[5], a[6], 1.0};
        while (left > 0.0) {
            int imed = G->med[p.ir];
            double sig = 0.0;
            if (imed != -1) {
                int lgle = pwlf_interval(imed, gle, T->ge1, T->ge0) - 1;
                double gmfp = pwlf_eval(imed * MXGE + lgle, gle, T->gmfp1, T->gmfp0) / G->rhof[p.ir];
                gmfp *= pwlf_eval(imed * MXGE + lgle, gle, T->cohe1, T->cohe0);
                sig = 1.0 / gmfp;
            }
            int idisc = 0, irnew = p.ir;
            double ustep = left;
            howfar(&p, &idisc, &irnew, &ustep);
            if (idisc > 0) { p.ir = 0; break; }
            p.x += ustep * p.u; p.y += ustep * p.v; p.z += ustep * p.w;
            tau += ustep * sig;
            left -= ustep;
            if (irnew != p.ir) { p.ir = irnew; if (irnew == 0) break; }
        }
        out[2 * i] = tau; out[2 * i + 1] = (double)p.ir;
    }
}

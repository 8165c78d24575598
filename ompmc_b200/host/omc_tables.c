/*
 * omc_tables.c -- physics tables of the ompMC hot path built from the raw data files (see omc_tables.h).
 *
 * Organisation (one builder per table family, all writing into one owner object):
 *   pegs_read()        PEGS4 media: scalars + the electron piece-wise-linear (PWL) tables          src/ompmc.c:5489-5943
 *   photon_tables()    XCOM photo/pair/triplet/Rayleigh + Klein-Nishina -> gmfp, gbr1, gbr2, cohe    :216-665
 *   rayleigh_tables()  form factors -> xgrid, fcum, b, c, i arrays, pmax                             :757-1007
 *   pair_tables()      Coulomb-corrected radiation-logarithm parameters dl1..dl6, bpar, zbrang       :1148-1318
 *   mscat_read()       msnew.data (Rutherford multiple-scattering alias tables)                     :3201-3280
 *   spin_tables()      spinms.data -> Mott rejection table, screening / first / second moment PWLs,
 *                      Moller-Bhabha scattering power correction                                    :2376-2947
 *   electron_derived() monotonicity flags, maximum cross sections, CSDA range table, tmxs           :3282-3511
 *
 * Every arithmetic expression keeps the reference's operand order and types (float where the reference uses float),
 * because the goal is the SAME doubles, not similar ones.
 */
#define _GNU_SOURCE
#include "omc_tables.h"

#include <math.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define RM OMC_RM
#define NGE OMC_MXGE
#define NEKE OMC_MXEKE
#define NFF OMC_MXRAYFF
#define NELEM 50                 /* elements read from the XCOM / form-factor files */
#define FSC 0.00729735255664
#define XIMAX 0.5
#define ESTEPE 0.25
#define NE_SPIN 15               /* MXE_SPIN */
#define NE_SPIN1 31              /* MXE_SPIN1 = 2*MXE_SPIN + 1 */
#define NQ_SPIN 15
#define NU_SPIN 31

typedef struct { double z, wa, pz, rhoz; } Elem;
typedef struct {
    char name[32];
    int ne, meke, found;
    Elem el[NELEM];
    double rho, rlc, ae, ap, ue, up, te, thmoll, delcm;
} Medium;

struct omc_tables {
    omc_media_tables v;
    int nmed;
    Medium med[OMC_MXMED];
    void *own[1024];
    int nown;
    char *err;
    int errlen;
};

static int failf(omc_tables *t, const char *fmt, ...) {
    if (t->err && t->errlen > 0) {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(t->err, (size_t)t->errlen, fmt, ap);
        va_end(ap);
    }
    return 1;
}

static void *own(omc_tables *t, size_t n, size_t sz) {
    void *p = calloc(n ? n : 1, sz);
    if (p && t->nown < (int)(sizeof t->own / sizeof t->own[0])) t->own[t->nown++] = p;
    return p;
}
#define DARR(n) ((double *)own(t, (size_t)(n), sizeof(double)))
#define IARR(n) ((int *)own(t, (size_t)(n), sizeof(int)))

/* coefficients of the line through (x_prev, v_prev) and (x, v) in the form v = c1*x + c0, with 1/(x - x_prev) = k1 */
static inline void pwl_pair(double v_prev, double v, double k1, double x, double *c1, double *c0) {
    *c1 = (v - v_prev) * k1;
    *c0 = v - (*c1) * x;
}

/* ------------------------------------------------------------------------------------------------------------------
 * PEGS4 file
 * ---------------------------------------------------------------------------------------------------------------- */
static void trim_key(char *s) {                     /* keep the first run of non-blank characters */
    char *p = s;
    while (*p == ' ' || *p == '\t') p++;
    char *q = p;
    while (*q && *q != ' ' && *q != '\t' && *q != '\n' && *q != '\r') q++;
    *q = '\0';
    if (p != s) memmove(s, p, (size_t)(q - p) + 1);
}

/* "KEY=VALUE, KEY=VALUE, ..." -> calls set(key, value) for every pair; returns 0 if a value does not parse */
typedef int (*kv_fn)(void *ctx, const char *key, const char *value);
static int parse_kv_line(const char *line, kv_fn set, void *ctx) {
    char buf[1024];
    strncpy(buf, line, sizeof buf - 1);
    buf[sizeof buf - 1] = '\0';
    char *save = NULL;
    for (char *tok = strtok_r(buf, ",", &save); tok; tok = strtok_r(NULL, ",", &save)) {
        char *eq = strchr(tok, '=');
        if (!eq) continue;
        *eq = '\0';
        char key[64];
        strncpy(key, tok, sizeof key - 1);
        key[sizeof key - 1] = '\0';
        trim_key(key);
        if (!set(ctx, key, eq + 1)) return 0;
    }
    return 1;
}
static int set_medium_kv(void *ctx, const char *key, const char *value) {
    Medium *m = (Medium *)ctx;
    if (!strcmp(key, "RHO")) return sscanf(value, "%lf", &m->rho) == 1;
    if (!strcmp(key, "NE")) return sscanf(value, "%d", &m->ne) == 1;
    return 1;                                       /* IUNRST, EPSTFL, IAPRIM ...: not used by the hot path */
}
static int set_element_kv(void *ctx, const char *key, const char *value) {
    Elem *e = (Elem *)ctx;
    if (!strcmp(key, "Z")) return sscanf(value, "%lf", &e->z) == 1;
    if (!strcmp(key, "A")) return sscanf(value, "%lf", &e->wa) == 1;
    if (!strcmp(key, "PZ")) return sscanf(value, "%lf", &e->pz) == 1;
    if (!strcmp(key, "RHOZ")) return sscanf(value, "%lf", &e->rhoz) == 1;
    return 1;
}

static int pegs_read(omc_tables *t, const char *pegs_file, const char *const *names) {
    omc_media_tables *v = &t->v;
    const int nmed = t->nmed;
    FILE *fp = fopen(pegs_file, "r");
    if (!fp) return failf(t, "Unable to open file: %s", pegs_file);
    double *blcc = DARR(nmed), *xcc = DARR(nmed), *eke0 = DARR(nmed), *eke1 = DARR(nmed);
    double *tab[16];
    for (int k = 0; k < 16; k++) tab[k] = DARR(nmed * NEKE);
    /* order of the 16 numbers of one energy bin in the file: two rows of eight */
    v->esig0 = tab[0]; v->esig1 = tab[1]; v->psig0 = tab[2]; v->psig1 = tab[3];
    v->ededx0 = tab[4]; v->ededx1 = tab[5]; v->pdedx0 = tab[6]; v->pdedx1 = tab[7];
    v->ebr10 = tab[8]; v->ebr11 = tab[9]; v->pbr10 = tab[10]; v->pbr11 = tab[11];
    v->pbr20 = tab[12]; v->pbr21 = tab[13]; v->tmxs0 = tab[14]; v->tmxs1 = tab[15];
    v->blcc = blcc; v->xcc = xcc; v->eke0 = eke0; v->eke1 = eke1;
    char line[1024];
    int nfound = 0;
    while (nfound < nmed && fgets(line, sizeof line, fp)) {
        if (strncmp(line, " MEDIUM=", 8) != 0) continue;
        char name[25];
        int n = 0;
        for (int c = 8; c < 32 && line[c] && line[c] != ' ' && line[c] != '\n'; c++) name[n++] = line[c];
        name[n] = '\0';
        int imed = -1;
        for (int i = 0; i < nmed; i++) {
            char cname[21];
            strncpy(cname, names[i], 20);
            cname[20] = '\0';
            if (!strcmp(name, cname)) { imed = i; break; }
        }
        if (imed < 0 || t->med[imed].found) continue;
        Medium *m = &t->med[imed];
        memset(m, 0, sizeof *m);
        strncpy(m->name, name, sizeof m->name - 1);
        if (!fgets(line, sizeof line, fp) || !parse_kv_line(line, set_medium_kv, m)) continue;
        if (m->ne < 0 || m->ne > NELEM) { fclose(fp); return failf(t, "medium %s: NE = %d out of range", name, m->ne); }
        int ok = 1;
        for (int e = 0; e < m->ne && ok; e++)
            ok = fgets(line, sizeof line, fp) && parse_kv_line(line, set_element_kv, &m->el[e]);
        if (!ok) continue;
        if (!fgets(line, sizeof line, fp) || sscanf(line, "%lf %lf %lf %lf %lf", &m->rlc, &m->ae, &m->ap, &m->ue, &m->up) != 5) continue;
        m->te = m->ae - RM;
        m->thmoll = (m->te) * 2 + RM;
        int msge, mge, mseke, mleke, mcmfp, mrange;
        if (!fgets(line, sizeof line, fp) ||
            sscanf(line, "%d %d %d %d %d %d %d", &msge, &mge, &mseke, &m->meke, &mleke, &mcmfp, &mrange) != 7) continue;
        if (m->meke > NEKE) continue;
        for (int i = 0; i < 7; i++)
            if (!fgets(line, sizeof line, fp)) ok = 0;
        double skip[5];
        if (!ok || sscanf(line, "%lf %lf %lf %lf %lf", &skip[0], &skip[1], &skip[2], &skip[3], &skip[4]) != 5) continue;
        /* free-format stream from here: dl6, delcm, alphi(2), bpar(2), delpos(2), xr0, teff0, blcc, xcc, eke0, eke1, then the bins */
        double unused[9];
        if (fscanf(fp, "%lf", &unused[0]) != 1 || fscanf(fp, "%lf %lf %lf %lf %lf", &m->delcm, &unused[1], &unused[2], &unused[3], &unused[4]) != 5 ||
            fscanf(fp, "%lf %lf", &unused[5], &unused[6]) != 2 ||
            fscanf(fp, "%lf %lf %lf %lf", &unused[7], &unused[8], &blcc[imed], &xcc[imed]) != 4 ||
            fscanf(fp, "%lf %lf", &eke0[imed], &eke1[imed]) != 2) {
            fclose(fp);
            return failf(t, "medium %s: truncated PEGS4 record", name);
        }
        for (int k = 0; k < m->meke; k++)
            for (int c = 0; c < 16; c++)
                if (fscanf(fp, "%lf", &tab[c][imed * NEKE + k]) != 1) { fclose(fp); return failf(t, "medium %s: truncated PWL table", name); }
        /* radiation lengths -> cm */
        const double dfacti = 1.0 / (m->rlc);
        blcc[imed] *= dfacti;
        for (int k = 0; k < m->meke; k++) {
            const int i = imed * NEKE + k;
            tab[0][i] *= dfacti; tab[2][i] *= dfacti; tab[4][i] *= dfacti; tab[6][i] *= dfacti;      /* esig0 psig0 ededx0 pdedx0 */
            tab[7][i] *= dfacti; tab[1][i] *= dfacti; tab[3][i] *= dfacti; tab[5][i] *= dfacti;      /* pdedx1 esig1 psig1 ededx1 */
        }
        xcc[imed] *= sqrt(dfacti);
        m->found = 1;
        nfound++;
    }
    fclose(fp);
    for (int i = 0; i < nmed; i++)
        if (!t->med[i].found) return failf(t, "Medium %s not found on pegs file %s", names[i], pegs_file);
    double *ap = DARR(nmed), *ae = DARR(nmed), *te = DARR(nmed), *th = DARR(nmed), *rho = DARR(nmed);
    int *meke = IARR(nmed);
    for (int i = 0; i < nmed; i++) {
        ap[i] = t->med[i].ap; ae[i] = t->med[i].ae; te[i] = t->med[i].te; th[i] = t->med[i].thmoll; rho[i] = t->med[i].rho;
        meke[i] = t->med[i].meke;
    }
    v->pegs_ap = ap; v->pegs_ae = ae; v->pegs_te = te; v->pegs_thmoll = th; v->pegs_rho = rho; v->pegs_meke = meke;
    return 0;
}

/* ------------------------------------------------------------------------------------------------------------------
 * photon cross sections
 * ---------------------------------------------------------------------------------------------------------------- */
typedef struct { int n[NELEM]; double *le[NELEM], *ls[NELEM]; } Xcom;       /* ln E, ln sigma per element */

static int xcom_read(omc_tables *t, const char *folder, const char *file, Xcom *x) {
    char path[512];
    snprintf(path, sizeof path, "%s%s", folder, file);
    FILE *fp = fopen(path, "r");
    if (!fp) return failf(t, "Unable to open file: %s", path);
    memset(x, 0, sizeof *x);
    for (int z = 0; z < NELEM; z++) {
        int n;
        if (fscanf(fp, "%d", &n) != 1 || n < 1) { fclose(fp); return failf(t, "Could not read the data file %s", path); }
        x->n[z] = n;
        x->le[z] = DARR(n + 1);
        x->ls[z] = DARR(n + 1);
        for (int j = 0; j < n; j++)
            if (fscanf(fp, "%lf %lf", &x->le[z][j], &x->ls[z][j]) != 2) { fclose(fp); return failf(t, "Could not read the data file %s", path); }
    }
    fclose(fp);
    return 0;
}

/* macroscopic-to-be sum over the elements of a medium of p_Z * sigma_Z(E) on the medium's photon energy grid.
 * kind 0: photo-absorption / Rayleigh (plain log-log interpolation); 1: pair, 2: triplet (threshold 2 / 4 m_e, interpolated
 * with the (1 - eth/E)^3 factor divided out) */
static void xsec_on_grid(int kind, int ne, const double *zs, const double *pzs, const Xcom *x, double ge0, double ge1, double *res) {
    for (int j = 0; j < NGE; j++) res[j] = 0.0;
    for (int i = 0; i < ne; i++) {
        const int z = (int)(zs[i] + 0.5) - 1;
        int n = x->n[z];
        double eth = 0.0;
        double *d0 = (double *)malloc(((size_t)n + 1) * sizeof(double)), *d1 = (double *)malloc(((size_t)n + 1) * sizeof(double));
        if (kind == 0) {
            for (int j = 0; j < n; j++) { d0[j] = x->le[z][j]; d1[j] = x->ls[z][j]; }
        } else {
            for (int j = 0; j < n; j++) { d0[j + 1] = x->le[z][j]; d1[j + 1] = x->ls[z][j]; }
            eth = (kind == 1) ? 2.0 * RM : 4.0 * RM;
            n++;
            for (int j = 1; j < n; j++) d1[j] -= 3.0 * log(1.0 - eth / exp(d0[j]));
            d0[0] = (double)log(eth);
            d1[0] = (double)d1[1];
        }
        for (int j = 0; j < NGE; j++) {
            const double gle = ((double)(j + 1) - ge0) / ge1;
            const double e = exp(gle);
            double sig = 0.0;
            if ((gle < d0[0]) || (gle >= d0[n - 1])) {
                if (kind != 0) sig = (gle < d0[0]) ? 0.0 : exp(d1[n - 1]);
            } else {
                int k;
                for (k = 0; k < n - 1; k++)
                    if ((gle >= d0[k]) && (gle < d0[k + 1])) break;
                const double p = (gle - d0[k]) / (d0[k + 1] - d0[k]);
                sig = exp(p * d1[k + 1] + (1.0 - p) * d1[k]);
            }
            if ((kind != 0) && (e > eth)) sig *= (1.0 - eth / e) * (1.0 - eth / e) * (1.0 - eth / e);
            res[j] += pzs[i] * sig;
        }
        free(d0);
        free(d1);
    }
}

/* total Klein-Nishina cross section per electron */
static double klein_nishina(double e) {
    const double con = 0.1274783851;
    const double ko = e / RM;
    if (ko < 0.01) return 8.0 * con / 3.0 * (1.0 - ko * (2.0 - ko * (5.2 - 13.3 * ko))) / RM;
    const double c1 = 1.0 / (ko * ko);
    const double c2 = 1.0 - 2.0 * (1.0 + ko) * c1;
    const double c3 = (1.0 + 2.0 * ko) * c1;
    const double eps2 = 1.0;
    const double eps1 = 1.0 / (1.0 + 2.0 * ko);
    return (c1 * (1.0 / eps1 - 1.0 / eps2) + c2 * log(eps2 / eps1) + eps2 * (c3 + 0.5 * eps2) - eps1 * (c3 + 0.5 * eps1)) / e * con;
}

static int photon_tables(omc_tables *t, const char *folder) {
    omc_media_tables *v = &t->v;
    const int nmed = t->nmed;
    Xcom photo, rayl, pair, trip;
    if (xcom_read(t, folder, "xcom_photo.data", &photo) || xcom_read(t, folder, "xcom_rayleigh.data", &rayl) ||
        xcom_read(t, folder, "xcom_pair.data", &pair) || xcom_read(t, folder, "xcom_triplet.data", &trip)) return 1;
    double *ge0 = DARR(nmed), *ge1 = DARR(nmed);
    double *gmfp0 = DARR(nmed * NGE), *gmfp1 = DARR(nmed * NGE), *gbr10 = DARR(nmed * NGE), *gbr11 = DARR(nmed * NGE);
    double *gbr20 = DARR(nmed * NGE), *gbr21 = DARR(nmed * NGE), *cohe0 = DARR(nmed * NGE), *cohe1 = DARR(nmed * NGE);
    double *s_photo = DARR(NGE), *s_rayl = DARR(NGE), *s_pair = DARR(NGE), *s_trip = DARR(NGE);
    for (int i = 0; i < nmed; i++) {
        const Medium *m = &t->med[i];
        ge1[i] = (double)(NGE - 1) / log(m->up / m->ap);
        ge0[i] = 1.0 - ge1[i] * log(m->ap);
        double sumA = 0.0, sumZ = 0.0;
        for (int j = 0; j < m->ne; j++) {
            sumA += m->el[j].pz * m->el[j].wa;
            sumZ += m->el[j].pz * m->el[j].z;
        }
        const double con2 = m->rho / (sumA * 1.6605655);
        /* elements in order of increasing Z (the sums below run in that order) */
        double zs[NELEM], pzs[NELEM];
        int order[NELEM];
        for (int j = 0; j < m->ne; j++) order[j] = j;
        for (int a = 1; a < m->ne; a++) {
            const int o = order[a];
            int b = a - 1;
            while (b >= 0 && m->el[order[b]].z > m->el[o].z) { order[b + 1] = order[b]; b--; }
            order[b + 1] = o;
        }
        for (int j = 0; j < m->ne; j++) { zs[j] = m->el[order[j]].z; pzs[j] = m->el[order[j]].pz; }
        xsec_on_grid(0, m->ne, zs, pzs, &photo, ge0[i], ge1[i], s_photo);
        xsec_on_grid(0, m->ne, zs, pzs, &rayl, ge0[i], ge1[i], s_rayl);
        xsec_on_grid(1, m->ne, zs, pzs, &pair, ge0[i], ge1[i], s_pair);
        xsec_on_grid(2, m->ne, zs, pzs, &trip, ge0[i], ge1[i], s_trip);
        double gle = 0.0, gmfp = 0.0, gbr1 = 0.0, gbr2 = 0.0, cohe = 0.0, gmfp_o = 0.0, gbr1_o = 0.0, gbr2_o = 0.0, cohe_o = 0.0;
        for (int j = 0; j < NGE; j++) {
            gle = ((double)(j + 1) - ge0[i]) / ge1[i];
            const double e = exp(gle);
            const double sig_kn = sumZ * klein_nishina(e);
            const double sig_p = s_pair[j] + s_trip[j];
            const double sigma = sig_kn + sig_p + s_photo[j];
            gmfp = 1.0 / (sigma * con2);
            gbr1 = sig_p / sigma;
            gbr2 = gbr1 + sig_kn / sigma;
            cohe = sigma / (s_rayl[j] + sigma);
            if (j > 0) {
                const int idx = i * NGE + (j - 1);
                pwl_pair(gmfp_o, gmfp, ge1[i], gle, &gmfp1[idx], &gmfp0[idx]);
                pwl_pair(gbr1_o, gbr1, ge1[i], gle, &gbr11[idx], &gbr10[idx]);
                pwl_pair(gbr2_o, gbr2, ge1[i], gle, &gbr21[idx], &gbr20[idx]);
                pwl_pair(cohe_o, cohe, ge1[i], gle, &cohe1[idx], &cohe0[idx]);
            }
            gmfp_o = gmfp; gbr1_o = gbr1; gbr2_o = gbr2; cohe_o = cohe;
        }
        const int last = i * NGE + NGE - 1;                /* last bin: slope of the one before, through the last node */
        gmfp1[last] = gmfp1[last - 1]; gmfp0[last] = gmfp - gmfp1[last] * gle;
        gbr11[last] = gbr11[last - 1]; gbr10[last] = gbr1 - gbr11[last] * gle;
        gbr21[last] = gbr21[last - 1]; gbr20[last] = gbr2 - gbr21[last] * gle;
        cohe1[last] = cohe1[last - 1]; cohe0[last] = cohe - cohe1[last] * gle;
    }
    v->ge0 = ge0; v->ge1 = ge1; v->gmfp0 = gmfp0; v->gmfp1 = gmfp1; v->gbr10 = gbr10; v->gbr11 = gbr11;
    v->gbr20 = gbr20; v->gbr21 = gbr21; v->cohe0 = cohe0; v->cohe1 = cohe1;
    return 0;
}

/* ------------------------------------------------------------------------------------------------------------------
 * Rayleigh scattering: form-factor sampling tables (EGSnrc prepare_rayleigh_data)
 * ---------------------------------------------------------------------------------------------------------------- */
static int rayleigh_tables(omc_tables *t, const char *pgs4form_file) {
    omc_media_tables *v = &t->v;
    const int nmed = t->nmed;
    FILE *fp = fopen(pgs4form_file, "r");
    if (!fp) return failf(t, "Unable to open file: %s", pgs4form_file);
    double *xval = DARR(NFF), *aff = DARR(NELEM * NFF);
    int ok = 1;
    for (int i = 0; i < NFF && ok; i++) ok = fscanf(fp, "%lf", &xval[i]) == 1;
    for (int i = 0; i < NELEM * NFF && ok; i++) ok = fscanf(fp, "%lf", &aff[i]) == 1;
    fclose(fp);
    if (!ok) return failf(t, "Could not read atomic form factors file %s", pgs4form_file);
    double *xgrid = DARR(nmed * NFF), *fcum = DARR(nmed * NFF), *barr = DARR(nmed * NFF), *carr = DARR(nmed * NFF);
    int *iarr = IARR(nmed * NFF);
    double *pmax0 = DARR(nmed * NGE), *pmax1 = DARR(nmed * NGE);
    double *ff = DARR(NFF), *pe = DARR(NGE);
    for (int i = 0; i < nmed; i++) {
        const Medium *m = &t->med[i];
        double *xg = xgrid + i * NFF, *fc = fcum + i * NFF, *ba = barr + i * NFF, *ca = carr + i * NFF;
        int *ia = iarr + i * NFF;
        for (int j = 0; j < NFF; j++) {                    /* independent-atom model */
            double ff2 = 0.0;
            xg[j] = xval[j];
            for (int k = 0; k < m->ne; k++) {
                const int z = (int)m->el[k].z - 1;
                ff2 += m->el[k].pz * pow(aff[z * NFF + j], 2);
            }
            ff[j] = sqrt(ff2);
        }
        if (xg[0] < 1.0E-6) xg[0] = 0.0001;
        const double emin = exp((1.0 - v->ge0[i]) / v->ge1[i]);
        const double emax = exp((NGE - v->ge0[i]) / v->ge1[i]);
        for (int j = 0; j < NFF; j++)                      /* no log(0) below: smallest denormal instead */
            if (ff[j] == 0.0 && !signbit(ff[j])) { const uint64_t one = 1; memcpy(&ff[j], &one, sizeof one); }
        double sum0 = 0.0;
        fc[0] = 0.0;
        for (int j = 0; j < NFF - 1; j++) {                /* power-law segments of F^2, cumulative integral */
            const double b = log(ff[j + 1] / ff[j]) / log(xg[j + 1] / xg[j]);
            ba[j] = b;
            const double x1 = xg[j], x2 = xg[j + 1];
            const double pow_x1 = pow(x1, 2.0 * b), pow_x2 = pow(x2, 2.0 * b);
            sum0 += pow(ff[j], 2) * (pow(x2, 2) * pow_x2 - pow(x1, 2) * pow_x1) / ((1.0 + b) * pow_x1);
            fc[j + 1] = sum0;
        }
        const double dle = log(emax / emin) / ((double)NGE - 1.0);
        int idx = 1;
        for (int j = 1; j <= NGE; j++) {                   /* cumulative up to the kinematic limit of each energy */
            const double e = emin * exp(dle * ((double)j - 1.0));
            const double xmax = 20.607544 * 2.0 * e / RM;
            int k;
            for (k = 1; k <= NFF - 1; k++)
                if ((xmax >= xg[k - 1]) && (xmax < xg[k])) break;
            idx = k;
            const double b = ba[idx - 1], x1 = xg[idx - 1], x2 = xmax;
            const double pow_x1 = pow(x1, 2.0 * b), pow_x2 = pow(x2, 2.0 * b);
            pe[j - 1] = fc[idx - 1] + pow(ff[idx - 1], 2) * (pow(x2, 2) * pow_x2 - pow(x1, 2) * pow_x1) / ((1.0 + b) * pow_x1);
        }
        ia[NFF - 1] = idx;
        const double anorm = 1.0 / sqrt(pe[NGE - 1]), anorm1 = 1.005 / pe[NGE - 1], anorm2 = 1.0 / pe[NGE - 1];
        for (int j = 0; j < NGE; j++) {
            pe[j] *= anorm1;
            if (pe[j] > 1.0) pe[j] = 1.0;
        }
        for (int j = 0; j < NFF; j++) {
            ff[j] *= anorm;
            fc[j] *= anorm2;
            ca[j] = (1.0 + ba[j]) / pow(xg[j] * ff[j], 2);
        }
        /* uniform cumulative bins -> starting segment of the search */
        const double dw = 1.0 / ((double)NFF - 1.0);
        double xold = xg[0];
        int ibin = 1;
        double b = ba[0];
        double pow_x1 = pow(xg[0], 2.0 * b);
        ia[0] = 1;
        for (int j = 2; j <= NFF - 1; j++) {
            double w = dw;
            for (;;) {
                const double x1 = xold, x2 = xg[ibin];
                const double tt = pow(x1, 2) * pow(x1, 2.0 * b);
                const double pow_x2 = pow(x2, 2.0 * b);
                const double aux = pow(ff[ibin - 1], 2) * (pow(x2, 2) * pow_x2 - tt) / ((1.0 + b) * pow_x1);
                if (aux > w) {
                    xold = exp(log(tt + w * (1.0 + b) * pow_x1 / pow(ff[ibin - 1], 2)) / (2.0 + 2.0 * b));
                    ia[j - 1] = ibin;
                    break;
                }
                w -= aux;
                xold = x2;
                ibin++;
                b = ba[ibin - 1];
                pow_x1 = pow(xold, 2.0 * b);
            }
        }
        for (int j = 0; j < NFF; j++) ba[j] = 0.5 / (1.0 + ba[j]);     /* the form the sampler wants */
        for (int j = 0; j < NGE - 1; j++) {
            const double gle = ((j + 1) - v->ge0[i]) / v->ge1[i];
            pmax1[i * NGE + j] = (pe[j + 1] - pe[j]) * v->ge1[i];
            pmax0[i * NGE + j] = pe[j] - pmax1[i * NGE + j] * gle;
        }
        pmax0[i * NGE + NGE - 1] = pmax0[i * NGE + NGE - 2];
        pmax1[i * NGE + NGE - 1] = pmax1[i * NGE + NGE - 2];
    }
    v->ray_xgrid = xgrid; v->ray_fcum = fcum; v->ray_b_array = barr; v->ray_c_array = carr; v->ray_i_array = iarr;
    v->ray_pmax0 = pmax0; v->ray_pmax1 = pmax1;
    return 0;
}

/* ------------------------------------------------------------------------------------------------------------------
 * pair production / bremsstrahlung screening parameters (EGSnrc fix_brems, $INITIALIZE-PAIR-ANGLE)
 * ---------------------------------------------------------------------------------------------------------------- */
static double coulomb_correction(double zi) {
    const double a = FSC * zi;
    double fc = 1 + pow(a, 2);
    fc = 1.0 / fc;
    fc = fc + 0.20206 - 0.0369 * pow(a, 2);
    fc = fc + 0.0083 * pow(a, 4);
    fc = fc - 0.002 * pow(a, 6);
    return fc * pow(a, 2);
}
static double atomic_electron_term(double zi, double fc) {
    if (zi == 4) return 5.924 / (4.710 - fc);
    if (zi == 3) return 5.805 / (4.740 - fc);
    if (zi == 2) return 5.621 / (4.790 - fc);
    if (zi == 1) return 6.144 / (5.310 - fc);
    return log(1194.0 * pow(zi, -2.0 / 3.0)) / log(184.15 * pow(zi, -1.0 / 3.0)) - fc;
}

static int pair_tables(omc_tables *t) {
    omc_media_tables *v = &t->v;
    const int nmed = t->nmed;
    double *dl[6];
    for (int k = 0; k < 6; k++) dl[k] = DARR(nmed * 8);
    double *bpar0 = DARR(nmed), *bpar1 = DARR(nmed), *delcm = DARR(nmed), *zbrang = DARR(nmed);
    /* rows of the screening-function fits: {constant, +4*Z term?, linear, quadratic} for the two delta regimes */
    for (int im = 0; im < nmed; im++) {
        const Medium *m = &t->med[im];
        double Zt = 0.0, Zb = 0.0, Zf = 0.0;
        for (int i = 0; i < m->ne; i++) {
            const double zi = m->el[i].z, pi = m->el[i].pz;
            const double fc = coulomb_correction(zi);
            const double xi = atomic_electron_term(zi, fc);
            const double aux = pi * zi * (zi + xi);
            Zt = Zt + aux;
            Zb = Zb - aux * log(zi) / 3.0;
            Zf = Zf + aux * fc;
        }
        const double Zv = (Zb - Zf) / Zt, Zg = Zb / Zt;
        const double fmax1 = 2.0 * (20.863 + 4.0 * Zg) - 2.0 * (20.029 + 4.0 * Zg) / 3.0;
        const double fmax2 = 2.0 * (20.863 + 4.0 * Zv) - 2.0 * (20.029 + 4.0 * Zv) / 3.0;
        double *d1 = dl[0] + im * 8, *d2 = dl[1] + im * 8, *d3 = dl[2] + im * 8, *d4 = dl[3] + im * 8, *d5 = dl[4] + im * 8, *d6 = dl[5] + im * 8;
        for (int k = 0; k < 8; k++) d6[k] = 0.952;
        for (int r = 0; r < 2; r++) {                      /* bremsstrahlung: rows 0,1 with Zg/fmax1, rows 2,3 with Zv/fmax2 */
            const double Z = r ? Zv : Zg, f = r ? fmax2 : fmax1;
            const int a = 2 * r, b = 2 * r + 1;
            d1[a] = (20.863 + 4.0 * Z) / f; d2[a] = -3.242 / f; d3[a] = 0.625 / f; d4[a] = (21.12 + 4.0 * Z) / f; d5[a] = -4.184 / f;
            d1[b] = (20.029 + 4.0 * Z) / f; d2[b] = -1.93 / f; d3[b] = -0.086 / f; d4[b] = (21.12 + 4.0 * Z) / f; d5[b] = -4.184 / f;
        }
        for (int r = 0; r < 2; r++) {                      /* pair production: rows 4,5 with Zg, rows 6,7 with Zv */
            const double Z = r ? Zv : Zg;
            const int a = 4 + 2 * r, b = 5 + 2 * r;
            d1[a] = (3.0 * (20.863 + 4.0 * Z) - (20.029 + 4.0 * Z)); d2[a] = (3.0 * (-3.242) - (-1.930)); d3[a] = (3.0 * (0.625) - (-0.086));
            d4[a] = (2.0 * 21.12 + 8.0 * Z); d5[a] = (2.0 * (-4.184));
            d1[b] = (3.0 * (20.863 + 4.0 * Z) + (20.029 + 4.0 * Z)); d2[b] = (3.0 * (-3.242) + (-1.930)); d3[b] = (3.0 * 0.625 + (-0.086));
            d4[b] = (4.0 * 21.12 + 16.0 * Z); d5[b] = (4.0 * (-4.184));
        }
        bpar1[im] = d1[6] / (3.0 * d1[7] + d1[6]);
        bpar0[im] = 12.0 * d1[7] / (3.0 * d1[7] + d1[6]);
        double zb = 0.0, pznorm = 0.0;
        for (int i = 0; i < m->ne; i++) {
            zb += (double)(m->el[i].pz) * (m->el[i].z) * ((m->el[i].z) + 1.0f);
            pznorm += m->el[i].pz;
        }
        zbrang[im] = (8.116224E-05) * pow(zb / pznorm, 1.0 / 3.0);
        delcm[im] = m->delcm;
    }
    v->dl1 = dl[0]; v->dl2 = dl[1]; v->dl3 = dl[2]; v->dl4 = dl[3]; v->dl5 = dl[4]; v->dl6 = dl[5];
    v->bpar0 = bpar0; v->bpar1 = bpar1; v->delcm = delcm; v->zbrang = zbrang;
    return 0;
}

/* ------------------------------------------------------------------------------------------------------------------
 * multiple scattering: Rutherford alias tables (msnew.data)
 * ---------------------------------------------------------------------------------------------------------------- */
static int mscat_read(omc_tables *t, const char *folder) {
    omc_media_tables *v = &t->v;
    char path[512];
    snprintf(path, sizeof path, "%smsnew.data", folder);
    FILE *fp = fopen(path, "r");
    if (!fp) return failf(t, "Unable to open file: %s", path);
    const int NL = OMC_MS_NL, NQ = OMC_MS_NQ, NU = OMC_MS_NU;
    double *ums = DARR(NL * NQ * NU), *fms = DARR(NL * NQ * NU), *wms = DARR(NL * NQ * NU);
    int *ims = IARR(NL * NQ * NU);
    int ok = 1;
    for (int b = 0; b < NL * NQ && ok; b++) {
        double *u = ums + b * NU, *f = fms + b * NU, *w = wms + b * NU;
        int *im = ims + b * NU;
        for (int k = 0; k < NU && ok; k++) ok = fscanf(fp, "%lf", &u[k]) == 1;
        for (int k = 0; k < NU && ok; k++) ok = fscanf(fp, "%lf", &f[k]) == 1;
        for (int k = 0; k < NU - 1 && ok; k++) ok = fscanf(fp, "%lf", &w[k]) == 1;
        for (int k = 0; k < NU - 1 && ok; k++) ok = fscanf(fp, "%d", &im[k]) == 1;
        for (int k = 0; k < NU - 1; k++) {
            f[k] = f[k + 1] / f[k] - 1.0;
            im[k] = im[k] - 1;
        }
        f[NU - 1] = f[NU - 2];
    }
    fclose(fp);
    if (!ok) return failf(t, "Could not read %s", path);
    const double llammin = log(1.0), llammax = log(1.0E5);
    const double dllamb = (llammax - llammin) / (NL - 1);
    const double dqms = 0.5 / (NQ - 1);
    v->ums = ums; v->fms = fms; v->wms = wms; v->ims = ims;
    v->dllambi = 1.0 / dllamb;
    v->dqmsi = 1.0 / dqms;
    return 0;
}

/* ------------------------------------------------------------------------------------------------------------------
 * cubic spline used by the spin corrections (natural end conditions, coefficients per interval)
 * ---------------------------------------------------------------------------------------------------------------- */
static void spline_set(const double *x, const double *f, double *a, double *b, double *c, double *d, int n) {
    double s, r;
    const int m1 = 2, m2 = n - 1;
    int m, mr;
    s = 0;
    for (m = 1; m <= m2; m++) {
        d[m - 1] = x[m] - x[m - 1];
        r = (f[m] - f[m - 1]) / d[m - 1];
        c[m - 1] = r - s;
        s = r;
    }
    s = 0; r = 0;
    c[0] = 0; c[n - 1] = 0;
    for (m = m1; m <= m2; m++) {
        c[m - 1] = c[m - 1] + r * c[m - 2];
        b[m - 1] = 2 * (x[m - 2] - x[m]) - r * s;
        s = d[m - 1];
        r = s / b[m - 1];
    }
    mr = m2;
    for (m = m1; m <= m2; m++) {
        c[mr - 1] = (d[mr - 1] * c[mr] - c[mr - 1]) / b[mr - 1];
        mr = mr - 1;
    }
    for (m = 1; m <= m2; m++) {
        s = d[m - 1];
        r = c[m] - c[m - 1];
        d[m - 1] = r / s;
        c[m - 1] = 3 * c[m - 1];
        b[m - 1] = (f[m] - f[m - 1]) / s - (c[m - 1] + r) * s;
        a[m - 1] = f[m - 1];
    }
}
static double spline_eval(double s, const double *x, const double *a, const double *b, const double *c, const double *d, int n) {
    int m_lower, m_upper, direction, m, ml, mu, mav;
    if (x[0] > x[n - 1]) { direction = 1; m_lower = n; m_upper = 0; }
    else { direction = 0; m_lower = 0; m_upper = n; }
    if (s >= x[m_upper + direction - 1]) {
        m = m_upper + 2 * direction - 1;
    } else if (s <= x[m_lower - direction]) {
        m = m_lower - 2 * direction + 1;
    } else {
        ml = m_lower; mu = m_upper;
        while (abs(mu - ml) > 1) {
            mav = (ml + mu) / 2;
            if (s < x[mav - 1]) mu = mav; else ml = mav;
        }
        m = mu + direction - 1;
    }
    const double q = s - x[m - 1];
    return a[m - 1] + q * (b[m - 1] + q * (c[m - 1] + q * d[m - 1]));
}

/* PWL table in ln E on the electron grid of medium im from a function evaluated at the grid nodes (node 1 .. meke);
 * last bin copies the one before */
typedef double (*node_fn)(void *ctx, double eil);
static void pwl_table_from_nodes(node_fn f, void *ctx, int meke, double eke0, double eke1, double *c1, double *c0) {
    double eil = (1.0 - eke0) / eke1;
    double v1 = f(ctx, eil);
    for (int i = 1; i <= meke - 1; i++) {
        eil = (i + 1 - eke0) / eke1;
        const double v2 = f(ctx, eil);
        c1[i - 1] = (v2 - v1) * eke1;
        c0[i - 1] = v2 - c1[i - 1] * eil;
        v1 = v2;
    }
    c1[meke - 1] = c1[meke - 2];
    c0[meke - 1] = c0[meke - 2];
}
typedef struct { const double *x, *a, *b, *c, *d; int n; } SplineCtx;
static double spline_node(void *ctx, double eil) {
    const SplineCtx *s = (const SplineCtx *)ctx;
    return spline_eval(eil, s->x, s->a, s->b, s->c, s->d, s->n);
}

/* ------------------------------------------------------------------------------------------------------------------
 * spin effects (Mott correction data of EGSnrc's spinms.data)
 * ---------------------------------------------------------------------------------------------------------------- */
static int spin_tables(omc_tables *t, const char *folder) {
    omc_media_tables *v = &t->v;
    const int nmed = t->nmed;
    char path[512];
    snprintf(path, sizeof path, "%sspinms.data", folder);
    FILE *fp = fopen(path, "rb");
    if (!fp) return failf(t, "Unable to open file: %s", path);
    fseek(fp, 0, SEEK_END);
    const long len = ftell(fp);
    rewind(fp);
    unsigned char *raw = (unsigned char *)own(t, (size_t)len + 8, 1);
    if (fread(raw, 1, (size_t)len, fp) != (size_t)len) { fclose(fp); return failf(t, "Could not read %s", path); }
    fclose(fp);
    const float *fbuf = (const float *)raw;                /* the file seen as float32 records of 276 ... */
    const short *sbuf = (const short *)raw;                /* ... and as int16 (552 per record) */
    float espin_min, espin_max, b2spin_max, b2spin_min;
    memcpy(&espin_min, raw + 36, 4); memcpy(&espin_max, raw + 40, 4); memcpy(&b2spin_min, raw + 44, 4); memcpy(&b2spin_max, raw + 48, 4);
    v->b2spin_min = (double)b2spin_min;
    const int nener = NE_SPIN, NE1 = NE_SPIN1 + 1;         /* 32 energies: 16 log-spaced, 16 linear in beta^2 */
    double dloge = log(espin_max / espin_min) / (double)nener;
    double eloge = log(espin_min);
    double earray[NE_SPIN1 + 2];
    earray[0] = espin_min;
    for (int i = 1; i <= nener; i++) {
        eloge += dloge;
        earray[i] = exp(eloge);
    }
    double dbeta2 = (b2spin_max - b2spin_min) / nener;
    double beta2 = b2spin_min;
    earray[nener + 1] = espin_max;
    for (int i = nener + 2; i <= 2 * nener + 1; i++) {
        beta2 += dbeta2;
        if (beta2 < 0.999) earray[i] = RM * 1000.0 * (1.0 / sqrt(1.0 - beta2) - 1);
        else earray[i] = 50585.1;
    }
    espin_min /= 1000.0;                                   /* keV -> MeV (kept in float, as the reference does) */
    espin_max /= 1000.0;
    const double dlener = log(espin_max / espin_min) / NE_SPIN;
    v->dleneri = 1.0 / dlener;
    v->espml = log(espin_min);
    dbeta2 = (b2spin_max - b2spin_min) / NE_SPIN;
    v->dbeta2i = 1.0 / dbeta2;
    const double dqq1 = 0.5 / NQ_SPIN;
    v->dqq1i = 1.0 / dqq1;

    const size_t per_med = (size_t)2 * NE1 * (NQ_SPIN + 1) * (NU_SPIN + 1);
    double *rej = DARR(nmed * per_med);
    double *etae0 = DARR(nmed * NEKE), *etae1 = DARR(nmed * NEKE), *etap0 = DARR(nmed * NEKE), *etap1 = DARR(nmed * NEKE);
    double *q1ce0 = DARR(nmed * NEKE), *q1ce1 = DARR(nmed * NEKE), *q1cp0 = DARR(nmed * NEKE), *q1cp1 = DARR(nmed * NEKE);
    double *q2ce0 = DARR(nmed * NEKE), *q2ce1 = DARR(nmed * NEKE), *q2cp0 = DARR(nmed * NEKE), *q2cp1 = DARR(nmed * NEKE);
    double *blcce0 = DARR(nmed * NEKE), *blcce1 = DARR(nmed * NEKE);
    /* +NE1 zeros behind the positron half: the reference reads eta_array[je + 1] one past its end at the top energies
     * (heap garbage there; fixtures pin those bins, see oracle/gen_fixtures.sanitize) */
    double *eta_arr = DARR(3 * NE1), *c_arr = DARR(2 * NE1), *g_arr = DARR(2 * NE1);
    double elarray[NE_SPIN1 + 1], farray[NE_SPIN1 + 1], af[NE_SPIN1 + 1], bf[NE_SPIN1 + 1], cf[NE_SPIN1 + 1], df[NE_SPIN1 + 1];
    for (int im = 0; im < nmed; im++) {
        const Medium *m = &t->med[im];
        double *R = rej + im * per_med;
        double sum_Z2 = 0.0, sum_A = 0.0, sum_pz = 0.0, sum_Z = 0.0;
        memset(eta_arr, 0, 3 * NE1 * sizeof(double));
        memset(c_arr, 0, 2 * NE1 * sizeof(double));
        memset(g_arr, 0, 2 * NE1 * sizeof(double));
        for (int ie = 0; ie < m->ne; ie++) {
            const double z = m->el[ie].z;
            const int iz = (int)(z + 0.5);
            const double pz = m->el[ie].pz;
            const double tmp = z * (z + 1.0) * pz;
            sum_Z2 += tmp;
            sum_Z += pz * z;
            sum_A += pz * m->el[ie].wa;
            sum_pz += pz;
            const double z23 = pow(z, 2.0 / 3.0);
            for (int iq = 0; iq < 2; iq++) {
                for (int i = 0; i <= NE_SPIN1; i++) {
                    const long irec = 1 + (long)(iz - 1) * 4 * (nener + 1) + 2 * iq * (nener + 1) + i + 1;
                    if ((irec) * 1104 > len) return failf(t, "%s has no record for Z = %d", path, iz);
                    const float *rec = fbuf + 276 * (irec - 1);
                    double dum1 = rec[0];
                    const double dum2 = rec[1], dum3 = rec[2], aux_o = rec[3];
                    const float *fmax = rec + 4;
                    const short *i2 = sbuf + 552 * (irec - 1) + 40;
                    eta_arr[iq * NE1 + i] += tmp * log(z23 * aux_o);
                    const double tau = earray[i] / (1000.0 * RM);      /* energies of the file are in keV */
                    beta2 = tau * (tau + 2) / ((tau + 1) * (tau + 1));
                    const double eta = z23 / ((137.03604 * 0.88534138) * (137.03604 * 0.88534138)) * aux_o / 4 / tau / (tau + 2);
                    c_arr[iq * NE1 + i] += tmp * (log(1.0 + 1.0 / eta) - 1.0 / (1.0 + eta)) * dum1 * dum3;
                    g_arr[iq * NE1 + i] += tmp * dum2;
                    double *Rq = R + ((size_t)iq * NE1 + i) * (NQ_SPIN + 1) * (NU_SPIN + 1);
                    for (int j = 0; j <= NQ_SPIN; j++)
                        for (int k = 0; k <= NU_SPIN; k++) {
                            int ii2 = (int)i2[(NU_SPIN + 1) * j + k];
                            if (ii2 < 0) ii2 += 65536;
                            dum1 = ii2;
                            dum1 = dum1 * fmax[j] / 65535;
                            Rq[j * (NU_SPIN + 1) + k] += tmp * dum1;
                        }
                }
            }
        }
        (void)sum_A; (void)sum_pz;
        /* rejection function: maximum of every (energy, q) row scaled to one */
        for (size_t row = 0; row < (size_t)2 * NE1 * (NQ_SPIN + 1); row++) {
            double *r = R + row * (NU_SPIN + 1);
            double flmax = 0.0;
            for (int k = 0; k <= NU_SPIN; k++)
                if (flmax < r[k]) flmax = r[k];
            for (int k = 0; k <= NU_SPIN; k++) r[k] = r[k] / flmax;
        }
        for (int i = 0; i <= NE_SPIN1; i++) {
            const double tau = (earray[i] / RM) * 0.001;
            beta2 = tau * (tau + 2.0) / pow(tau + 1.0, 2.0);
            for (int iq = 0; iq < 2; iq++) {
                const double aux_o = exp(eta_arr[iq * NE1 + i] / sum_Z2) / (pow(137.03604 * 0.88534138, 2.0));
                eta_arr[iq * NE1 + i] = 0.26112447 * aux_o * (v->blcc[im]) / (v->xcc[im]);
                const double eta = aux_o / 4.0 / tau / (tau + 2);
                const double gamma = 3.0 * (1.0 + eta) * (log(1.0 + 1.0 / eta) * (1.0 + 2.0 * eta) - 2.0) / (log(1.0 + 1.0 / eta) * (1.0 + eta) - 1.0);
                g_arr[iq * NE1 + i] = g_arr[iq * NE1 + i] / sum_Z2 / gamma;
                c_arr[iq * NE1 + i] = c_arr[iq * NE1 + i] / sum_Z2 / (log(1.0 + 1.0 / eta) - 1.0 / (1.0 + eta));
            }
        }
        /* screening parameter correction eta_ms(E): linear interpolation of eta_arr in ln E below espin_max, in beta^2 above */
        const double eke0 = v->eke0[im], eke1 = v->eke1[im];
        const int neke = m->meke;
        double si1e = 0, si1p = 0, si2e, si2p;
        for (int i = 0; i < neke; i++) {
            const double eil = (i == 0) ? (1.0 - eke0) / eke1 : (i + 1.0 - eke0) / eke1;
            const double e = exp(eil);
            double se, sp;
            if (e <= espin_min) {
                se = eta_arr[0]; sp = eta_arr[NE1];
            } else {
                double aae;
                int je;
                if (e <= espin_max) {
                    aae = (eil - v->espml) * v->dleneri;
                    je = (int)aae;
                    aae = aae - je;
                } else {
                    const double tau = e / RM;
                    beta2 = (i == 0) ? tau * (tau + 2.0) / pow(tau + 1.0, 2.0) : tau * (tau + 2.0) / ((tau + 1.0) * (tau + 1.0));
                    aae = (beta2 - v->b2spin_min) * v->dbeta2i;
                    je = (int)aae;
                    aae = aae - je;
                    je = je + NE_SPIN + 1;
                }
                se = (1 - aae) * eta_arr[je] + aae * eta_arr[je + 1];
                sp = (1 - aae) * eta_arr[NE1 + je] + aae * eta_arr[NE1 + je + 1];
            }
            if (i == 0) { si1e = se; si1p = sp; continue; }
            si2e = se; si2p = sp;
            etae1[NEKE * im + i - 1] = (si2e - si1e) * eke1;
            etae0[NEKE * im + i - 1] = (si2e - etae1[NEKE * im + i - 1] * eil);
            etap1[NEKE * im + i - 1] = (si2p - si1p) * eke1;
            etap0[NEKE * im + i - 1] = (si2p - etap1[NEKE * im + i - 1] * eil);
            si1e = si2e; si1p = si2p;
        }
        etae1[NEKE * im + neke - 1] = etae1[NEKE * im + neke - 2]; etae0[NEKE * im + neke - 1] = etae0[NEKE * im + neke - 2];
        etap1[NEKE * im + neke - 1] = etap1[NEKE * im + neke - 2]; etap0[NEKE * im + neke - 1] = etap0[NEKE * im + neke - 2];

        /* first (c_arr) and second (g_arr) moment corrections: splines over ln E, one node dropped at the junction of the
         * two energy grids, a closing node = 1 at max(UE, 1e5) */
        for (int i = 0; i <= NE_SPIN; i++) elarray[i] = log(earray[i] / 1000.0);
        for (int i = NE_SPIN + 1; i <= NE_SPIN1 - 1; i++) elarray[i] = log(earray[i + 1] / 1000.0);
        const int ndata = NE_SPIN1 + 1;
        elarray[ndata - 1] = (m->ue > 1.0E5) ? log(m->ue) : log(1.0E5);
        SplineCtx sc = {elarray, af, bf, cf, df, ndata};
        double *dst1[4] = {q1ce1, q1cp1, q2ce1, q2cp1}, *dst0[4] = {q1ce0, q1cp0, q2ce0, q2cp0};
        for (int which = 0; which < 4; which++) {
            const double *src = ((which < 2) ? c_arr : g_arr) + (which & 1) * NE1;
            for (int i = 0; i <= NE_SPIN; i++) farray[i] = src[i];
            for (int i = NE_SPIN + 1; i <= NE_SPIN1 - 1; i++) farray[i] = src[i + 1];
            farray[ndata - 1] = 1.0;
            spline_set(elarray, farray, af, bf, cf, df, ndata);
            pwl_table_from_nodes(spline_node, &sc, neke, eke0, eke1, dst1[which] + NEKE * im, dst0[which] + NEKE * im);
        }
        q1ce0[NEKE * im + neke - 1] = q1ce1[NEKE * im + neke - 2];        /* Q5: the reference copies the SLOPE here */

        /* scattering power already carried by discrete Moller / Bhabha events is taken out of blcc */
        const double tauc = m->te / RM;
        si1e = 1.0;
        for (int i = 1; i <= neke - 1; i++) {
            const double eil = ((double)(i + 1) - eke0) / eke1;
            const double e = exp(eil);
            const int leil = i;
            const double tau = e / RM;
            si2e = 1.0;
            if (tau > 2.0 * tauc) {
                double sig = v->esig1[NEKE * im + leil] * eil + v->esig0[NEKE * im + leil];
                const double dedx = v->ededx1[NEKE * im + leil] * eil + v->ededx0[NEKE * im + leil];
                sig /= dedx;
                if (sig > 1.0E-6) {
                    const double etap = etae1[NEKE * im + leil] * eil + etae0[NEKE * im + leil];
                    const double eta = 0.25 * etap * (v->xcc[im]) / (v->blcc[im]) / tau / (tau + 2);
                    const double g_r = (1.0 + 2.0 * eta) * log(1.0 + 1.0 / eta) - 2.0;
                    double g_m = log(0.5 * tau / tauc) + (1.0 + ((tau + 2.0) / (tau + 1.0)) * ((tau + 2.0) / (tau + 1.0))) *
                        log(2.0 * (tau - tauc + 2.0) / (tau + 4.0)) - 0.25 * (tau + 2.0) *
                        (tau + 2.0 + 2.0 * (2.0 * tau + 1.0) / ((tau + 1.0) * (tau + 1.0))) * log((tau + 4.0) * (tau - tauc) / tau / (tau - tauc + 2.0)) +
                        0.5 * (tau - 2.0 * tauc) * (tau + 2.0) * (1.0 / (tau - tauc) - 1.0 / ((tau + 1.0) * (tau + 1.0)));
                    if (g_m < g_r) g_m /= g_r; else g_m = 1.0;
                    si2e = 1.0 - g_m * sum_Z / sum_Z2;
                }
            }
            blcce1[NEKE * im + i - 1] = (si2e - si1e) * eke1;
            blcce0[NEKE * im + i - 1] = si2e - blcce1[NEKE * im + i - 1] * eil;
            si1e = si2e;
        }
        blcce1[NEKE * im + neke - 1] = blcce1[NEKE * im + neke - 2];
        blcce0[NEKE * im + neke - 1] = blcce0[NEKE * im + neke - 2];
    }
    v->spin_rej = rej;
    v->etae_ms0 = etae0; v->etae_ms1 = etae1; v->etap_ms0 = etap0; v->etap_ms1 = etap1;
    v->q1ce_ms0 = q1ce0; v->q1ce_ms1 = q1ce1; v->q1cp_ms0 = q1cp0; v->q1cp_ms1 = q1cp1;
    v->q2ce_ms0 = q2ce0; v->q2ce_ms1 = q2ce1; v->q2cp_ms0 = q2cp0; v->q2cp_ms1 = q2cp1;
    v->blcce0 = blcce0; v->blcce1 = blcce1;
    return 0;
}

/* ------------------------------------------------------------------------------------------------------------------
 * electron quantities derived from the PEGS tables: cross-section maxima, CSDA range table, maximum step tmxs
 * ---------------------------------------------------------------------------------------------------------------- */
static int electron_derived(omc_tables *t) {
    omc_media_tables *v = &t->v;
    const int nmed = t->nmed;
    int *mono = IARR(2 * nmed);
    double *esig_e = DARR(nmed), *psig_e = DARR(nmed), *e_array = DARR(nmed * NEKE), *range_ep = DARR(2 * nmed * NEKE);
    double *tmxs0 = (double *)v->tmxs0, *tmxs1 = (double *)v->tmxs1;
    for (int im = 0; im < nmed; im++) {
        double sigee = 1.0E-15, sigep = 1.0E-15, sige_old = -1.0, sigp_old = -1.0;
        int ise = 1, isp = 1;
        const int neke = t->med[im].meke;
        for (int i = 1; i <= neke; i++) {
            const double ei = exp(((double)i - v->eke0[im]) / v->eke1[im]);
            const double eil = log(ei);
            const int l = im * NEKE + i - 1;
            double ededx = v->ededx1[l] * eil + v->ededx0[l];
            double sig = v->esig1[l] * eil + v->esig0[l];
            sig /= ededx;
            if (sig > sigee) sigee = sig;
            if (sig < sige_old) ise = 0;
            sige_old = sig;
            ededx = v->pdedx1[l] * eil + v->pdedx0[l];
            sig = v->psig1[l] * eil + v->psig0[l];
            sig /= ededx;
            if (sig > sigep) sigep = sig;
            if (sig < sigp_old) isp = 0;
            sigp_old = sig;
        }
        mono[0 * nmed + im] = ise; mono[1 * nmed + im] = isp;
        esig_e[im] = sigee; psig_e[im] = sigep;
    }
    for (int im = 0; im < nmed; im++) {
        const double eke0 = v->eke0[im], eke1 = v->eke1[im];
        const int neke = t->med[im].meke;
        double *ea = e_array + im * NEKE, *rp = range_ep + nmed * NEKE + im * NEKE, *re = range_ep + im * NEKE;
        const double *pd1 = v->pdedx1 + im * NEKE, *pd0 = v->pdedx0 + im * NEKE, *ed1 = v->ededx1 + im * NEKE, *ed0 = v->ededx0 + im * NEKE;
        double ei = exp((1.0 - eke0) / eke1);
        ea[0] = ei;
        re[0] = 0.0; rp[0] = 0.0;
        for (int i = 1; i <= neke - 1; i++) {              /* range by log-interpolated stopping power, series-expanded integral */
            const double eip1 = exp(((double)(i + 1) - eke0) / eke1);
            ea[i] = eip1;
            const double eke = 0.5 * (eip1 + ei);
            const double elke = log(eke);
            const int lelke = (int)(eke1 * elke + eke0) - 1;
            double ededx = pd1[lelke] * elke + pd0[lelke];
            double aux = pd1[i - 1] / ededx;
            rp[i] = rp[i - 1] + (eip1 - ei) / ededx * (1.0 + aux * (1.0 + 2.0 * aux) * pow((eip1 - ei) / eke, 2.0) / 24.0);
            ededx = ed1[lelke] * elke + ed0[lelke];
            aux = ed1[i - 1] / ededx;
            re[i] = re[i - 1] + (eip1 - ei) / ededx * (1.0 + aux * (1.0 + 2.0 * aux) * pow(((eip1 - ei) / eke), 2.0) / 24.0);
            ei = eip1;
        }
        /* tmxs: path length over which the first MS moment grows by XIMAX, capped at an energy loss fraction ESTEPE */
        double eil = (1.0 - eke0) / eke1;
        ei = exp(eil);
        int leil = 1;
        double p2 = ei * (ei + 2.0 * RM);
        double beta2 = p2 / (p2 + pow(RM, 2.0));
        double chi_a2 = v->xcc[im] / (4.0 * p2 * v->blcc[im]);
        const double dedx0 = ed1[leil] * eil + ed0[leil];
        double estepx = 2.0 * p2 * beta2 * dedx0 / ei / v->xcc[im] / (log(1.0 + 1.0 / chi_a2) * (1.0 + chi_a2) - 1.0);
        estepx *= XIMAX;
        if (estepx > ESTEPE) estepx = ESTEPE;
        double si = estepx * ei / dedx0, sip1;
        for (int i = 1; i <= neke - 1; i++) {
            const double elke = ((double)(i + 1) - eke0) / eke1;
            const double eke = exp(elke);
            const int lelke = i;
            p2 = eke * (eke + 2.0 * RM);
            beta2 = p2 / (p2 + pow(RM, 2.0));
            chi_a2 = v->xcc[im] / (4.0 * p2 * v->blcc[im]);
            double ededx = ed1[lelke] * elke + ed0[lelke];
            estepx = 2.0 * p2 * beta2 * ededx / eke / (v->xcc[im]) / (log(1.0 + 1.0 / chi_a2) * (1.0 + chi_a2) - 1.0);
            estepx = estepx * XIMAX;
            if (estepx > ESTEPE) estepx = ESTEPE;
            double ekef = (1.0 - estepx) * eke;
            int lelkef;
            if (ekef <= ea[0]) {
                sip1 = (ea[0] - ekef) / dedx0;
                ekef = ea[0];
                lelkef = 0;
            } else {
                const double elkef = log(ekef);
                lelkef = eke1 * elkef + eke0 - 1;
                const int leip1l = lelkef + 1;
                const double eip1l = ((double)(leip1l + 1) - eke0) / eke1;
                const double eip1 = ea[leip1l];
                double aux = (eip1 - ekef) / eip1;
                const double elktmp = 0.5 * (elkef + eip1l + 0.25 * aux * aux * (1.0 + aux * (1.0 + 0.875 * aux)));
                const double ektmp = 0.5 * (ekef + eip1);
                ededx = ed1[lelkef] * elktmp + ed0[lelkef];
                aux = ed1[lelkef] / ededx;
                sip1 = (eip1 - ekef) / ededx * (1.0 + aux * (1.0 + 2.0 * aux) * (pow(((eip1 - ekef) / ektmp), 2.0) / 24.0));
            }
            sip1 += re[lelke] - re[lelkef + 1];
            tmxs1[im * NEKE + i - 1] = (sip1 - si) * eke1;
            tmxs0[im * NEKE + i - 1] = sip1 - tmxs1[im * NEKE + i - 1] * elke;
            si = sip1;
        }
        tmxs0[im * NEKE + neke - 1] = tmxs0[im * NEKE + neke - 2];
        tmxs1[im * NEKE + neke - 1] = tmxs1[im * NEKE + neke - 2];
    }
    v->sig_ismonotone = mono; v->esig_e = esig_e; v->psig_e = psig_e; v->e_array = e_array; v->range_ep = range_ep;
    return 0;
}

/* ------------------------------------------------------------------------------------------------------------------ */
omc_tables *omc_tables_build(const char *data_folder, const char *pegs_file, const char *pgs4form_file, int nmed, const char *const *names,
                             char *err, int errlen) {
    if (err && errlen > 0) err[0] = '\0';
    if (nmed < 1 || nmed > OMC_MXMED || !data_folder || !pegs_file || !pgs4form_file || !names) {
        if (err && errlen > 0) snprintf(err, (size_t)errlen, "bad arguments (nmed must be 1..%d)", OMC_MXMED);
        return NULL;
    }
    omc_tables *t = (omc_tables *)calloc(1, sizeof *t);
    if (!t) return NULL;
    t->nmed = nmed; t->v.nmed = nmed; t->err = err; t->errlen = errlen;
    int rc = pegs_read(t, pegs_file, names);
    if (!rc) rc = photon_tables(t, data_folder);
    if (!rc) rc = rayleigh_tables(t, pgs4form_file);
    if (!rc) rc = pair_tables(t);
    if (!rc) rc = mscat_read(t, data_folder);
    if (!rc) {
        /* Euler's constant absorbed into blcc, xcc squared: the forms used throughout (src/ompmc.c:3289-3295) */
        double *blcc = (double *)t->v.blcc, *xcc = (double *)t->v.xcc;
        for (int i = 0; i < nmed; i++) {
            blcc[i] = 1.16699413758864573 * blcc[i];
            xcc[i] = pow(xcc[i], 2.0);
        }
        rc = spin_tables(t, data_folder);
    }
    if (!rc) rc = electron_derived(t);
    t->err = NULL;
    if (rc) { omc_tables_free(t); return NULL; }
    return t;
}

const omc_media_tables *omc_tables_view(const omc_tables *t) { return t ? &t->v : NULL; }

void omc_tables_free(omc_tables *t) {
    if (!t) return;
    for (int i = 0; i < t->nown; i++) free(t->own[i]);
    free(t);
}

/* ------------------------------------------------------------------------------------------------------------------
 * source spectrum -> inverse CDF on a 1000-point grid
 * ---------------------------------------------------------------------------------------------------------------- */
int omc_spectrum_cdfinv(const char *spectrum_file, double *cdfinv1, double *cdfinv2, double *emax, char *err, int errlen) {
    FILE *fp = fopen(spectrum_file, "r");
    if (!fp) { if (err) snprintf(err, (size_t)errlen, "Unable to open file: %s", spectrum_file); return 1; }
    char line[1024];
    double enmin;
    int n, imode;
    if (!fgets(line, sizeof line, fp) || !fgets(line, sizeof line, fp) || sscanf(line, "%d %lf %d", &n, &enmin, &imode) != 3 || n < 1 ||
        n > 200) {
        fclose(fp);
        if (err) snprintf(err, (size_t)errlen, "bad spectrum header in %s", spectrum_file);
        return 1;
    }
    double *eup = (double *)malloc((size_t)n * sizeof(double)), *pdf = (double *)malloc((size_t)n * sizeof(double)), *cdf = (double *)malloc((size_t)n * sizeof(double));
    for (int i = 0; i < n; i++)
        if (!fgets(line, sizeof line, fp) || sscanf(line, "%lf %lf", &eup[i], &pdf[i]) != 2) {
            fclose(fp); free(eup); free(pdf); free(cdf);
            if (err) snprintf(err, (size_t)errlen, "bad spectrum bin %d in %s", i, spectrum_file);
            return 1;
        }
    fclose(fp);
    if (imode == 1) {                                       /* counts / MeV -> counts / bin */
        pdf[0] *= (eup[0] - enmin);
        for (int i = 1; i < n; i++) pdf[i] *= (eup[i] - eup[i - 1]);
    } else if (imode != 0) {
        free(eup); free(pdf); free(cdf);
        if (err) snprintf(err, (size_t)errlen, "Invalid mode number in spectrum file.");
        return 1;
    }
    cdf[0] = pdf[0];
    for (int i = 1; i < n; i++) cdf[i] = cdf[i - 1] + pdf[i];
    const double fnorm = 1.0 / cdf[n - 1];
    const double deltak = OMC_INVDIM;
    const double gridsz = 1.0f / deltak;
    for (int i = 0; i < n; i++) cdf[i] *= fnorm;
    for (int k = 0; k < OMC_INVDIM; k++) {
        const double ak = (double)k * gridsz;
        int i;
        for (i = 0; i < n; i++)
            if (ak <= cdf[i]) break;
        cdfinv1[k] = (i != 0) ? eup[i - 1] : enmin;
        cdfinv2[k] = eup[i] - cdfinv1[k];
    }
    if (emax) *emax = eup[n - 1];
    free(eup); free(pdf); free(cdf);
    return 0;
}

/*
 * omc_host_input.h -- everything an omc_dosxyz run needs before the batch loop, from the reference's own input files and
 * without the reference's sources: the key = value input file (parseInputFile/getInputValue,
 * src/omc_utilities.c:62-123), the .egsphant phantom (initPhantom, omc_dosxyz.c:62-175), the per-region transport data
 * (initRegions, omc_dosxyz.c:890-962), the collimated point source (initSource, omc_dosxyz.c:368-627) and nsplit (initVrt,
 * src/ompmc.c:5964-5983).  The physics tables come from omc_tables.c.  Reference behaviour that shapes the numbers is kept
 * (Q10 region.ecut stays 0 when AE > global ecut, Q15 atoi bookkeeping, Q16 lines with '#' are dropped and
 * keys match by substring, first match wins).
 */
#ifndef OMC_HOST_INPUT_H
#define OMC_HOST_INPUT_H
#include <ctype.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "omc_tables.h"

/* ---- key = value input file ---------------------------------------------------------------------------------------- */
typedef struct { char key[128], value[384]; } inp_item;
typedef struct { inp_item it[128]; int n; } inp_file;

static int inp_blank(const char *s) {
    for (; *s; s++)
        if (!isspace((unsigned char)*s)) return 0;
    return 1;
}
/* `stem` without the ".inp" extension, as the reference's -i option takes it */
static int inp_parse(inp_file *f, const char *stem) {
    char path[512], line[512];
    snprintf(path, sizeof path, "%s.inp", stem);
    FILE *fp = fopen(path, "r");
    if (!fp) { printf("Unable to open file: %s\n", path); return 1; }
    f->n = 0;
    while (fgets(line, sizeof line, fp) && f->n < 128) {
        if (strchr(line, '#') || inp_blank(line)) continue;         /* a '#' anywhere drops the line (Q16) */
        char *eq = strchr(line, '=');
        if (!eq) continue;
        *eq = '\0';
        char *val = eq + 1;
        val[strcspn(val, "\r\n")] = '\0';
        snprintf(f->it[f->n].key, sizeof f->it[f->n].key, "%s", line);
        snprintf(f->it[f->n].value, sizeof f->it[f->n].value, "%s", val);
        f->n++;
    }
    fclose(fp);
    return 0;
}
/* first item whose key CONTAINS `key`; the reference reports "nothing parsed" for a file with a single item (Q16) */
static int inp_get(const inp_file *f, const char *key, char *dest) {
    if (f->n - 1 == 0) return 0;
    for (int i = 0; i < f->n; i++)
        if (strstr(f->it[i].key, key)) { strcpy(dest, f->it[i].value); return 1; }
    return 0;
}
static void strip_spaces(char *dst, const char *src) {
    for (; *src; src++)
        if (!isspace((unsigned char)*src)) *dst++ = *src;
    *dst = '\0';
}
static int inp_path(const inp_file *f, const char *key, char *dest) {   /* value with every blank removed */
    char buf[384];
    if (!inp_get(f, key, buf)) { printf("Can not find '%s' key on input file.\n", key); return 0; }
    strip_spaces(dest, buf);
    return 1;
}

/* ---- phantom ------------------------------------------------------------------------------------------------------- */
typedef struct {
    int nmed, isize, jsize, ksize;
    char names[OMC_MXMED][64];
    double *xb, *yb, *zb, *dens;
    int *med;                       /* 1-based medium index per voxel, x fastest */
} host_phantom;

static int phantom_read(host_phantom *p, const char *path) {
    FILE *fp = fopen(path, "r");
    char line[8192];
    if (!fp) { printf("Unable to open file: %s\n", path); return 1; }
    if (!fgets(line, sizeof line, fp)) { fclose(fp); return 1; }
    p->nmed = atoi(line);
    if (p->nmed < 1 || p->nmed > OMC_MXMED) { printf("Number of media %d out of range.\n", p->nmed); fclose(fp); return 1; }
    for (int i = 0; i < p->nmed; i++) {
        if (!fgets(line, sizeof line, fp)) { fclose(fp); return 1; }
        strip_spaces(p->names[i], line);
    }
    if (!fgets(line, sizeof line, fp) || !fgets(line, sizeof line, fp) ||      /* one dummy line, then the voxel counts */
        sscanf(line, "%d %d %d", &p->isize, &p->jsize, &p->ksize) != 3 || p->isize < 1 || p->jsize < 1 || p->ksize < 1) {
        printf("Bad voxel counts in %s\n", path);
        fclose(fp);
        return 1;
    }
    const size_t nvox = (size_t)p->isize * p->jsize * p->ksize;
    p->xb = malloc(((size_t)p->isize + 1) * sizeof(double)); p->yb = malloc(((size_t)p->jsize + 1) * sizeof(double));
    p->zb = malloc(((size_t)p->ksize + 1) * sizeof(double));
    p->med = malloc(nvox * sizeof(int)); p->dens = malloc(nvox * sizeof(double));
    int ok = 1;
    for (int i = 0; i <= p->isize && ok; i++) ok = fscanf(fp, "%lf", &p->xb[i]) == 1;
    for (int i = 0; i <= p->jsize && ok; i++) ok = fscanf(fp, "%lf", &p->yb[i]) == 1;
    for (int i = 0; i <= p->ksize && ok; i++) ok = fscanf(fp, "%lf", &p->zb[i]) == 1;
    if (!ok || !fgets(line, sizeof line, fp)) { printf("Bad voxel boundaries in %s\n", path); fclose(fp); return 1; }
    for (int k = 0; k < p->ksize; k++) {                    /* one digit per voxel, one text line per row, blank line per slice */
        for (int j = 0; j < p->jsize; j++) {
            for (int i = 0; i < p->isize; i++) p->med[i + (size_t)j * p->isize + (size_t)k * p->jsize * p->isize] = fgetc(fp) - '0';
            if (!fgets(line, sizeof line, fp)) ok = 0;
        }
        if (!fgets(line, sizeof line, fp)) ok = (k == p->ksize - 1) ? ok : 0;
    }
    for (size_t v = 0; v < nvox && ok; v++) ok = fscanf(fp, "%lf", &p->dens[v]) == 1;
    fclose(fp);
    if (!ok) { printf("Truncated phantom file %s\n", path); return 1; }
    return 0;
}

/* ---- resampling (BASELINE config 5: the PROSTATE grid "up-sampled 3 mm -> 2 mm / 1 mm") ---------------------------------------
 * The reference has no resampler: its users prepare the finer .egsphant outside.  This one keeps the physical
 * extent of the phantom and cuts every axis into round(extent / size) equal voxels of (about) the requested size -- any
 * ratio, not only integer splits.  A new voxel takes the VOLUME-weighted mean density of the old voxels it overlaps and the
 * medium that fills most of its volume (ties: the lowest medium index; 0 = vacuum competes like a medium), so that mass is
 * conserved exactly and an integer split reproduces the material map voxel for voxel.  Same rule, same arithmetic order as
 * ompmc_b200.problem.resample_phantom_to(). */
typedef struct { int n; int *lo, *cnt; double *w; } axis_overlap;   /* new cell i overlaps old cells lo[i] .. lo[i]+cnt[i]-1 with lengths w[off..] */

static double *axis_resample(const double *b, int n, double size, int *n_out) {
    const double ext = b[n] - b[0];
    int m = (int)floor(ext / size + 0.5);
    if (m < 1) m = 1;
    double *out = malloc(((size_t)m + 1) * sizeof(double));
    for (int i = 0; i <= m; i++) out[i] = b[0] + ext * (double)i / (double)m;
    out[m] = b[n];
    *n_out = m;
    return out;
}

static int phantom_resample(const host_phantom *src, double sx, double sy, double sz, host_phantom *dst) {
    if (!(sx > 0.0 && sy > 0.0 && sz > 0.0)) { printf("Voxel sizes must be positive.\n"); return 1; }
    *dst = *src;
    dst->xb = axis_resample(src->xb, src->isize, sx, &dst->isize);
    dst->yb = axis_resample(src->yb, src->jsize, sy, &dst->jsize);
    dst->zb = axis_resample(src->zb, src->ksize, sz, &dst->ksize);
    const size_t nvox = (size_t)dst->isize * dst->jsize * dst->ksize;
    if (nvox + 1 > 2147483647u) { printf("Resampled grid has too many voxels.\n"); return 1; }
    dst->med = malloc(nvox * sizeof(int));
    dst->dens = malloc(nvox * sizeof(double));
    /* per axis: first old cell and overlap lengths of every new cell */
    const double *ob[3] = {src->xb, src->yb, src->zb}, *nb[3] = {dst->xb, dst->yb, dst->zb};
    const int on[3] = {src->isize, src->jsize, src->ksize}, nn[3] = {dst->isize, dst->jsize, dst->ksize};
    int *first[3], *count[3];
    double **len[3];
    for (int a = 0; a < 3; a++) {
        first[a] = malloc((size_t)nn[a] * sizeof(int)); count[a] = malloc((size_t)nn[a] * sizeof(int));
        len[a] = malloc((size_t)nn[a] * sizeof(double *));
        int j = 0;
        for (int i = 0; i < nn[a]; i++) {
            const double lo = nb[a][i], hi = nb[a][i + 1];
            while (j < on[a] - 1 && ob[a][j + 1] <= lo) j++;
            int k = j, c = 0;
            double *w = malloc((size_t)on[a] * sizeof(double));
            while (k < on[a] && ob[a][k] < hi) {
                const double l = (ob[a][k + 1] < hi ? ob[a][k + 1] : hi) - (ob[a][k] > lo ? ob[a][k] : lo);
                w[c++] = l > 0.0 ? l : 0.0;
                k++;
            }
            first[a][i] = j; count[a][i] = c;
            len[a][i] = realloc(w, (size_t)(c > 0 ? c : 1) * sizeof(double));
        }
    }
    for (int kz = 0; kz < nn[2]; kz++)
        for (int jy = 0; jy < nn[1]; jy++)
            for (int ix = 0; ix < nn[0]; ix++) {
                double vol[OMC_MXMED + 1] = {0.0}, mass = 0.0, vtot = 0.0;
                for (int c = 0; c < count[2][kz]; c++)
                    for (int b = 0; b < count[1][jy]; b++)
                        for (int a = 0; a < count[0][ix]; a++) {
                            const size_t o = (size_t)(first[0][ix] + a) + (size_t)(first[1][jy] + b) * on[0] + (size_t)(first[2][kz] + c) * on[0] * on[1];
                            const double v = len[0][ix][a] * len[1][jy][b] * len[2][kz][c];
                            vol[src->med[o]] += v;
                            mass += v * src->dens[o];
                            vtot += v;
                        }
                int best = 0;
                for (int m = 1; m <= src->nmed; m++) if (vol[m] > vol[best]) best = m;
                const size_t d = (size_t)ix + (size_t)jy * nn[0] + (size_t)kz * nn[0] * nn[1];
                dst->med[d] = best;
                dst->dens[d] = vtot > 0.0 ? mass / vtot : 0.0;
            }
    for (int a = 0; a < 3; a++) {
        for (int i = 0; i < nn[a]; i++) free(len[a][i]);
        free(len[a]); free(first[a]); free(count[a]);
    }
    return 0;
}

/* ---- regions ------------------------------------------------------------------------------------------------------- */
typedef struct { int *med; double *rhof, *pcut, *ecut; } host_regions;

static void regions_init(host_regions *r, const host_phantom *p, const omc_media_tables *t, double ecut, double pcut) {
    const size_t nreg = (size_t)p->isize * p->jsize * p->ksize + 1;
    r->med = malloc(nreg * sizeof(int));
    r->rhof = calloc(nreg, sizeof(double)); r->pcut = calloc(nreg, sizeof(double)); r->ecut = calloc(nreg, sizeof(double));
    r->med[0] = -1;                                          /* region 0 = outside: vacuum */
    for (size_t i = 1; i < nreg; i++) {
        const int imed = p->med[i - 1] - 1;
        r->med[i] = imed;
        if (imed == -1) continue;                            /* vacuum voxel: rhof = cuts = 0 */
        r->rhof[i] = (p->dens[i - 1] == 0.0F) ? 1.0 : p->dens[i - 1] / t->pegs_rho[imed];
        if (t->pegs_ap[imed] <= pcut) {
            r->pcut[i] = pcut;
        } else {
            printf("Warning!, global pcut value is below PEGS's pcut value %f for medium %d, using PEGS value.\n", t->pegs_ap[imed], imed);
            r->pcut[i] = t->pegs_ap[imed];
        }
        if (t->pegs_ae[imed] <= ecut) {
            r->ecut[i] = ecut;
        } else {                                             /* Q10: the reference warns and leaves ecut unset; 0 here */
            printf("Warning!, global pcut value is below PEGS's ecut value %f for medium %d, using PEGS value.\n", t->pegs_ae[imed], imed);
        }
    }
}

/* ---- source -------------------------------------------------------------------------------------------------------- */
static int first_index_reaching(const double *b, int start, double x) {
    int i = start;
    while ((b[i] <= x) && (b[i + 1] < x)) i++;
    return i;
}
/* cdfinv1/cdfinv2: caller-provided [OMC_INVDIM] */
static int source_init(omc_source_dosxyz *s, const inp_file *f, const host_phantom *p, double *cdfinv1, double *cdfinv2) {
    char buf[384], path[384], err[256];
    memset(s, 0, sizeof *s);
    s->spectrum = 1;
    if (!inp_get(f, "spectrum file", buf)) {
        printf("Can not find 'spectrum file' key on input file.\nSwitch to monoenergetic case.\n");
        s->spectrum = 0;
    }
    if (s->spectrum) {
        strip_spaces(path, buf);
        if (omc_spectrum_cdfinv(path, cdfinv1, cdfinv2, NULL, err, sizeof err)) { printf("%s\n", err); return 1; }
        printf("Path to spectrum file : %s\n", path);
        s->deltak = OMC_INVDIM; s->cdfinv1 = cdfinv1; s->cdfinv2 = cdfinv2;
    } else {
        if (!inp_get(f, "mono energy", buf)) { printf("Can not find 'mono energy' key on input file.\n"); return 1; }
        s->energy = atof(buf);
        printf("%f monoenergetic source\n", s->energy);
    }
    if (!inp_get(f, "collimator bounds", buf)) { printf("Can not find 'collimator bounds' key on input file.\n"); return 1; }
    sscanf(buf, "%lf %lf %lf %lf", &s->xinl, &s->xinu, &s->yinl, &s->yinu);
    /* the field clipped to the phantom surface, x then y */
    if (s->xinl < p->xb[0]) s->xinl = p->xb[0];
    if (s->xinu <= s->xinl) s->xinu = s->xinl;
    if (s->xinu > p->xb[p->isize]) s->xinu = p->xb[p->isize];
    if (s->xinl > p->xb[p->isize]) s->xinl = p->xb[p->isize];
    s->ixinl = first_index_reaching(p->xb, 0, s->xinl);
    /* (the reference starts the upper search at ixinl - 1, i.e. reads xbounds[-1] when the field starts in the first voxel;
     * ixinu / iyinu are only printed, never used by the transport: started at ixinl here) */
    s->ixinu = first_index_reaching(p->xb, s->ixinl > 0 ? s->ixinl - 1 : 0, s->xinu);
    if (s->yinl < p->yb[0]) s->yinl = p->yb[0];
    if (s->yinu <= s->yinl) s->yinu = s->yinl;
    if (s->yinu > p->yb[p->jsize]) s->yinu = p->yb[p->jsize];
    if (s->yinl > p->yb[p->jsize]) s->yinl = p->yb[p->jsize];
    s->iyinl = first_index_reaching(p->yb, 0, s->yinl);
    s->iyinu = first_index_reaching(p->yb, s->iyinl > 0 ? s->iyinl - 1 : 0, s->yinu);
    printf("Index ranges for radiation field:\ni index ranges over i = %d to %d\nj index ranges over i = %d to %d\n", s->ixinl, s->ixinu,
           s->iyinl, s->iyinu);
    s->xsize = s->xinu - s->xinl;
    s->ysize = s->yinu - s->yinl;
    if (!inp_get(f, "charge", buf)) { printf("Can not find 'charge' key on input file.\n"); return 1; }
    s->charge = atoi(buf);
    if (s->charge < -1 || s->charge > 1) { printf("Particle kind not recognized.\n"); return 1; }
    if (!inp_get(f, "ssd", buf)) { printf("Can not find 'ssd' key on input file.\n"); return 1; }
    s->ssd = atof(buf);
    if (s->ssd < 0) { printf("SSD must be greater than zero.\n"); return 1; }
    return 0;
}
#endif

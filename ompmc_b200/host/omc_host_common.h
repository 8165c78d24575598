/*
 * omc_host_common.h -- shared by the plain-C host drivers (omc_dosxyz_b200.c, omc_matrad_b200.c): the problem blob
 * ("OMCBLOB1", written by ompmc_b200/problem.py save_blob: everything a reference user code holds in its globals just
 * before the batch loop -- the output of initMediaData(), the phantom, the regions, the source) and its mapping onto the
 * C-ABI structs of include/ompmc_b200.h.
 */
#ifndef OMC_HOST_COMMON_H
#define OMC_HOST_COMMON_H
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ompmc_b200.h"

/* ---- problem blob ------------------------------------------------------------------------- */
typedef struct { char name[33]; uint32_t dtype; uint64_t count; void *data; } blob_entry;
typedef struct { int n; blob_entry *e; } blob;

static int blob_read(blob *b, const char *path) {
    FILE *fp = fopen(path, "rb");
    if (!fp) return -1;
    char magic[8];
    uint32_t n = 0;
    if (fread(magic, 1, 8, fp) != 8 || memcmp(magic, "OMCBLOB1", 8) || fread(&n, 4, 1, fp) != 1) { fclose(fp); return -2; }
    b->n = (int)n;
    b->e = calloc(n, sizeof(blob_entry));
    for (uint32_t i = 0; i < n; i++) {
        blob_entry *e = &b->e[i];
        uint32_t pad;
        if (fread(e->name, 1, 32, fp) != 32 || fread(&e->dtype, 4, 1, fp) != 1 || fread(&pad, 4, 1, fp) != 1 ||
            fread(&e->count, 8, 1, fp) != 1) { fclose(fp); return -3; }
        size_t sz = (size_t)e->count * (e->dtype == 0 ? 8 : 4), psz = (sz + 7) & ~(size_t)7;
        e->data = malloc(psz ? psz : 8);
        if (fread(e->data, 1, psz, fp) != psz) { fclose(fp); return -4; }
    }
    fclose(fp);
    return 0;
}
static const blob_entry *blob_find(const blob *b, const char *name) {
    for (int i = 0; i < b->n; i++) if (!strcmp(b->e[i].name, name)) return &b->e[i];
    printf("Can not find '%s' in the problem file.\n", name);
    exit(EXIT_FAILURE);
}
static const double *F(const blob *b, const char *n) { return (const double *)blob_find(b, n)->data; }
static const int *I(const blob *b, const char *n) { return (const int *)blob_find(b, n)->data; }


/* media tables + geometry of the blob -> C-ABI structs (borrowed pointers into the blob) */
static void host_load_media_geometry(const blob *b, omc_media_tables *t, omc_geometry *g, const double **dens) {
    memset(t, 0, sizeof *t);
    t->nmed = I(b, "nmed")[0];
#define TF(f) t->f = F(b, #f);
#define TI(f) t->f = I(b, #f);
    TF(ge0) TF(ge1) TF(gmfp0) TF(gmfp1) TF(gbr10) TF(gbr11) TF(gbr20) TF(gbr21) TF(cohe0) TF(cohe1)
    TF(ray_xgrid) TF(ray_fcum) TF(ray_b_array) TF(ray_c_array) TI(ray_i_array) TF(ray_pmax0) TF(ray_pmax1)
    TF(dl1) TF(dl2) TF(dl3) TF(dl4) TF(dl5) TF(dl6) TF(bpar0) TF(bpar1) TF(delcm) TF(zbrang)
    TF(esig0) TF(esig1) TF(psig0) TF(psig1) TF(ededx0) TF(ededx1) TF(pdedx0) TF(pdedx1) TF(ebr10) TF(ebr11) TF(pbr10) TF(pbr11)
    TF(pbr20) TF(pbr21) TF(tmxs0) TF(tmxs1) TF(blcce0) TF(blcce1) TF(etae_ms0) TF(etae_ms1) TF(etap_ms0) TF(etap_ms1)
    TF(q1ce_ms0) TF(q1ce_ms1) TF(q1cp_ms0) TF(q1cp_ms1) TF(q2ce_ms0) TF(q2ce_ms1) TF(q2cp_ms0) TF(q2cp_ms1)
    TF(range_ep) TF(e_array) TF(eke0) TF(eke1) TI(sig_ismonotone) TF(esig_e) TF(psig_e) TF(xcc) TF(blcc)
    TF(spin_rej) TF(ums) TF(fms) TF(wms) TI(ims) TF(pegs_ap) TF(pegs_ae) TF(pegs_te) TF(pegs_thmoll) TF(pegs_rho) TI(pegs_meke)
#undef TF
#undef TI
    t->b2spin_min = F(b, "b2spin_min")[0]; t->dbeta2i = F(b, "dbeta2i")[0]; t->espml = F(b, "espml")[0];
    t->dleneri = F(b, "dleneri")[0]; t->dqq1i = F(b, "dqq1i")[0]; t->dllambi = F(b, "dllambi")[0]; t->dqmsi = F(b, "dqmsi")[0];
    g->isize = I(b, "isize")[0]; g->jsize = I(b, "jsize")[0]; g->ksize = I(b, "ksize")[0];
    g->xbounds = F(b, "xbounds"); g->ybounds = F(b, "ybounds"); g->zbounds = F(b, "zbounds");
    g->med = I(b, "region_med"); g->rhof = F(b, "region_rhof"); g->pcut = F(b, "region_pcut"); g->ecut = F(b, "region_ecut");
    *dens = F(b, "med_densities");
}
/* ---- the inverse: C-ABI structs -> problem blob (same entry names as ompmc_b200/problem.py), used by --dump-problem so that
 * what an input-file run hands to the GPU library can be compared entry by entry with a dump of the reference's globals */
static void blob_put(FILE *fp, const char *name, int is_int, const void *data, uint64_t count) {
    char nm[32];
    memset(nm, 0, sizeof nm);
    strncpy(nm, name, 31);
    const uint32_t dtype = is_int ? 1u : 0u, pad = 0;
    fwrite(nm, 1, 32, fp); fwrite(&dtype, 4, 1, fp); fwrite(&pad, 4, 1, fp); fwrite(&count, 8, 1, fp);
    const size_t sz = (size_t)count * (is_int ? 4 : 8), psz = (sz + 7) & ~(size_t)7;
    fwrite(data, 1, sz, fp);
    const char zeros[8] = {0};
    if (psz > sz) fwrite(zeros, 1, psz - sz, fp);
}
static int host_dump_problem(const char *stem, const omc_media_tables *t, const omc_geometry *g, const double *dens, const omc_source_dosxyz *s,
                             int nsplit) {
    char path[512];
    snprintf(path, sizeof path, "%s.problem", stem);
    FILE *fp = fopen(path, "wb");
    if (!fp) { printf("Unable to open file: %s\n", path); return 1; }
    const int nmed = t->nmed;
    const uint64_t nge = (uint64_t)nmed * OMC_MXGE, nek = (uint64_t)nmed * OMC_MXEKE, nff = (uint64_t)nmed * OMC_MXRAYFF;
    const uint64_t nsp = (uint64_t)nmed * 2 * OMC_SPIN_NE * OMC_SPIN_NQ * OMC_SPIN_NU, nms = (uint64_t)OMC_MS_NL * OMC_MS_NQ * OMC_MS_NU;
    const uint64_t nvox = (uint64_t)g->isize * g->jsize * g->ksize, nreg = nvox + 1;
    struct { const char *n; int is_int; const void *p; uint64_t c; } e[] = {
        {"nmed", 1, &t->nmed, 1},
        {"ge0", 0, t->ge0, (uint64_t)nmed}, {"ge1", 0, t->ge1, (uint64_t)nmed}, {"gmfp0", 0, t->gmfp0, nge}, {"gmfp1", 0, t->gmfp1, nge},
        {"gbr10", 0, t->gbr10, nge}, {"gbr11", 0, t->gbr11, nge}, {"gbr20", 0, t->gbr20, nge}, {"gbr21", 0, t->gbr21, nge},
        {"cohe0", 0, t->cohe0, nge}, {"cohe1", 0, t->cohe1, nge},
        {"ray_xgrid", 0, t->ray_xgrid, nff}, {"ray_fcum", 0, t->ray_fcum, nff}, {"ray_b_array", 0, t->ray_b_array, nff},
        {"ray_c_array", 0, t->ray_c_array, nff}, {"ray_i_array", 1, t->ray_i_array, nff}, {"ray_pmax0", 0, t->ray_pmax0, nge},
        {"ray_pmax1", 0, t->ray_pmax1, nge},
        {"dl1", 0, t->dl1, (uint64_t)nmed * 8}, {"dl2", 0, t->dl2, (uint64_t)nmed * 8}, {"dl3", 0, t->dl3, (uint64_t)nmed * 8},
        {"dl4", 0, t->dl4, (uint64_t)nmed * 8}, {"dl5", 0, t->dl5, (uint64_t)nmed * 8}, {"dl6", 0, t->dl6, (uint64_t)nmed * 8},
        {"bpar0", 0, t->bpar0, (uint64_t)nmed}, {"bpar1", 0, t->bpar1, (uint64_t)nmed}, {"delcm", 0, t->delcm, (uint64_t)nmed},
        {"zbrang", 0, t->zbrang, (uint64_t)nmed},
        {"esig0", 0, t->esig0, nek}, {"esig1", 0, t->esig1, nek}, {"psig0", 0, t->psig0, nek}, {"psig1", 0, t->psig1, nek},
        {"ededx0", 0, t->ededx0, nek}, {"ededx1", 0, t->ededx1, nek}, {"pdedx0", 0, t->pdedx0, nek}, {"pdedx1", 0, t->pdedx1, nek},
        {"ebr10", 0, t->ebr10, nek}, {"ebr11", 0, t->ebr11, nek}, {"pbr10", 0, t->pbr10, nek}, {"pbr11", 0, t->pbr11, nek},
        {"pbr20", 0, t->pbr20, nek}, {"pbr21", 0, t->pbr21, nek}, {"tmxs0", 0, t->tmxs0, nek}, {"tmxs1", 0, t->tmxs1, nek},
        {"blcce0", 0, t->blcce0, nek}, {"blcce1", 0, t->blcce1, nek}, {"etae_ms0", 0, t->etae_ms0, nek}, {"etae_ms1", 0, t->etae_ms1, nek},
        {"etap_ms0", 0, t->etap_ms0, nek}, {"etap_ms1", 0, t->etap_ms1, nek}, {"q1ce_ms0", 0, t->q1ce_ms0, nek}, {"q1ce_ms1", 0, t->q1ce_ms1, nek},
        {"q1cp_ms0", 0, t->q1cp_ms0, nek}, {"q1cp_ms1", 0, t->q1cp_ms1, nek}, {"q2ce_ms0", 0, t->q2ce_ms0, nek}, {"q2ce_ms1", 0, t->q2ce_ms1, nek},
        {"q2cp_ms0", 0, t->q2cp_ms0, nek}, {"q2cp_ms1", 0, t->q2cp_ms1, nek},
        {"range_ep", 0, t->range_ep, 2 * nek}, {"e_array", 0, t->e_array, nek}, {"eke0", 0, t->eke0, (uint64_t)nmed}, {"eke1", 0, t->eke1, (uint64_t)nmed},
        {"sig_ismonotone", 1, t->sig_ismonotone, (uint64_t)2 * nmed}, {"esig_e", 0, t->esig_e, (uint64_t)nmed}, {"psig_e", 0, t->psig_e, (uint64_t)nmed},
        {"xcc", 0, t->xcc, (uint64_t)nmed}, {"blcc", 0, t->blcc, (uint64_t)nmed},
        {"b2spin_min", 0, &t->b2spin_min, 1}, {"dbeta2i", 0, &t->dbeta2i, 1}, {"espml", 0, &t->espml, 1}, {"dleneri", 0, &t->dleneri, 1},
        {"dqq1i", 0, &t->dqq1i, 1}, {"spin_rej", 0, t->spin_rej, nsp},
        {"ums", 0, t->ums, nms}, {"fms", 0, t->fms, nms}, {"wms", 0, t->wms, nms}, {"ims", 1, t->ims, nms},
        {"dllambi", 0, &t->dllambi, 1}, {"dqmsi", 0, &t->dqmsi, 1},
        {"pegs_ap", 0, t->pegs_ap, (uint64_t)nmed}, {"pegs_ae", 0, t->pegs_ae, (uint64_t)nmed}, {"pegs_te", 0, t->pegs_te, (uint64_t)nmed},
        {"pegs_thmoll", 0, t->pegs_thmoll, (uint64_t)nmed}, {"pegs_rho", 0, t->pegs_rho, (uint64_t)nmed}, {"pegs_meke", 1, t->pegs_meke, (uint64_t)nmed},
        {"isize", 1, &g->isize, 1}, {"jsize", 1, &g->jsize, 1}, {"ksize", 1, &g->ksize, 1},
        {"xbounds", 0, g->xbounds, (uint64_t)g->isize + 1}, {"ybounds", 0, g->ybounds, (uint64_t)g->jsize + 1}, {"zbounds", 0, g->zbounds, (uint64_t)g->ksize + 1},
        {"med_densities", 0, dens, nvox},
        {"region_med", 1, g->med, nreg}, {"region_rhof", 0, g->rhof, nreg}, {"region_pcut", 0, g->pcut, nreg}, {"region_ecut", 0, g->ecut, nreg},
        {"src_spectrum", 1, &s->spectrum, 1}, {"src_charge", 1, &s->charge, 1}, {"src_energy", 0, &s->energy, 1}, {"src_deltak", 0, &s->deltak, 1},
        {"src_ssd", 0, &s->ssd, 1}, {"src_xinl", 0, &s->xinl, 1}, {"src_xinu", 0, &s->xinu, 1}, {"src_yinl", 0, &s->yinl, 1}, {"src_yinu", 0, &s->yinu, 1},
        {"src_xsize", 0, &s->xsize, 1}, {"src_ysize", 0, &s->ysize, 1}, {"src_ixinl", 1, &s->ixinl, 1}, {"src_ixinu", 1, &s->ixinu, 1},
        {"src_iyinl", 1, &s->iyinl, 1}, {"src_iyinu", 1, &s->iyinu, 1}, {"nsplit", 1, &nsplit, 1},
    };
    const uint32_t n = (uint32_t)(sizeof e / sizeof e[0]) + (s->spectrum ? 2u : 0u);
    fwrite("OMCBLOB1", 1, 8, fp);
    fwrite(&n, 4, 1, fp);
    for (size_t i = 0; i < sizeof e / sizeof e[0]; i++) blob_put(fp, e[i].n, e[i].is_int, e[i].p, e[i].c);
    if (s->spectrum) {
        blob_put(fp, "src_cdfinv1", 0, s->cdfinv1, (uint64_t)s->deltak);
        blob_put(fp, "src_cdfinv2", 0, s->cdfinv2, (uint64_t)s->deltak);
    }
    fclose(fp);
    printf("Problem written to %s\n", path);
    return 0;
}
#endif

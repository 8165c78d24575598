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
#endif

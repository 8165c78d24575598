/*
 * omc_tables.h -- host-side physics table initialisation for the B200 hot path, restated from scratch:
 * what the reference's initMediaData() computes (src/ompmc.c:5450-5487 -> readPegsFile :5489, initPhotonData :471,
 * initRayleighData :820, initPairData :1189, initMscatData :3282 incl. initSpinData :2376) from the PEGS4 / XCOM /
 * form-factor / msnew / spinms data files, delivered directly in the layout omc_gpu_set_media() takes
 * (include/ompmc_b200.h: omc_media_tables).  With it a host program needs no reference source to feed the GPU library.
 * Results are bit-identical to the reference's tables (tests/test_tables.py compares against blobs dumped from the
 * reference itself); reference quirks that reach the tables are reproduced on purpose (Q5, Q9).
 */
#ifndef OMC_TABLES_H
#define OMC_TABLES_H
#include "ompmc_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct omc_tables omc_tables;

/* data_folder must end with '/' and hold xcom_{photo,rayleigh,pair,triplet}.data, msnew.data and spinms.data (the reference's
 * "data folder" key); pegs_file / pgs4form_file as the reference's keys of the same names; names[nmed] = medium names as in
 * the phantom.  Returns NULL on failure with a message in err. */
omc_tables *omc_tables_build(const char *data_folder, const char *pegs_file, const char *pgs4form_file, int nmed,
                             const char *const *names, char *err, int errlen);
/* borrowed view, valid until omc_tables_free() */
const omc_media_tables *omc_tables_view(const omc_tables *t);
void omc_tables_free(omc_tables *t);

/* initSource() spectrum part, omc_dosxyz.c:383-507: inverse-CDF tables cdfinv1/cdfinv2[OMC_INVDIM]; returns 0 on success */
#define OMC_INVDIM 1000
int omc_spectrum_cdfinv(const char *spectrum_file, double *cdfinv1, double *cdfinv2, double *emax, char *err, int errlen);

#ifdef __cplusplus
}
#endif
#endif

/*
 * ref_harness.c -- drives the UNMODIFIED reference (edoerner/ompMC) as a shared library.
 * TEST INFRASTRUCTURE (oracle/): only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load the library built from this file.
 *
 * The reference's translation units are compiled from where they lie under /root/reference
 * (oracle/Makefile; nothing is copied into this repository):
 *     src/ompmc.c  src/omc_utilities.c                       -- as is
 *     src/omc_random.c                                       -- as is, its entry points renamed
 *                                                               ranmar_* by -D on the command line
 *     ucodes/omc_dosxyz/omc_dosxyz.c                         -- #included below, main() renamed
 * This file adds what the reference lacks to be usable as an oracle:
 *   (1) setRandom() dispatching between the reference's RANMAR and the per-history Philox stream
 *       the GPU library uses (oracle/omc_philox.h)  -> lock-step comparison history by history;
 *   (2) an ausgab() wrapper that counts deposits per history;
 *   (3) dump / load of the fully initialised problem (tables, geometry, regions, source) to a
 *       blob (oracle/omc_blob.h), because /root/reference and its data files do not exist on the
 *       GPU box;
 *   (4) the batch loop of main() (omc_dosxyz.c:1237-1263) as callable functions.
 */
#define _GNU_SOURCE
#include <time.h>
#ifdef OMC_REF_MATRAD
/* matRad user code: compiled against oracle/mexshim/mex.h (no MATLAB here); only its initHistory(ibeamlet),
 * callbacks and scoring are used, mexFunction()/parseInput() are never called. */
#define ausgab omc_dosxyz_reference_ausgab
#include "ucodes/omc_matrad/omc_matrad.c"
#undef ausgab
#undef exit
static int g_ibeamlet_shared = 0;      /* beamlet of the running batch (the reference's outer loop variable) */
#define INIT_HISTORY() initHistory(g_ibeamlet_shared)
#else
#define main omc_dosxyz_reference_main
#define ausgab omc_dosxyz_reference_ausgab
#include "ucodes/omc_dosxyz/omc_dosxyz.c"
#undef main
#undef ausgab
#define INIT_HISTORY() initHistory()
#endif

#include <stdint.h>
#include "omc_philox.h"
#include "omc_blob.h"
#include "../include/ompmc_b200.h"

/* renamed entry points of src/omc_random.c */
extern void ranmar_initRandom(void);
extern double ranmar_setRandom(void);
extern void ranmar_cleanRandom(void);

/* main() / mexFunction() of the user code (never called) still reference these two */
void initRandom(void) { ranmar_initRandom(); }
void cleanRandom(void) { ranmar_cleanRandom(); }
#ifdef OMC_REF_MATRAD
void ref_set_beamlet(int ibeamlet) { g_ibeamlet_shared = ibeamlet; }
#else
void ref_set_beamlet(int ibeamlet) { (void)ibeamlet; }
#endif

/* ---- (1) RNG dispatch -------------------------------------------------------------------- */
static int g_rng_mode = 0;              /* 0 = RANMAR (reference), 1 = Philox per history */
static uint32_t g_seed0 = 97, g_seed1 = 33;
static omc_philox g_philox;
static unsigned int g_ndeposit;
static double g_edep_sum;
#ifdef _OPENMP
#pragma omp threadprivate(g_philox, g_ndeposit, g_edep_sum)
#endif

double setRandom(void) {
    if (g_rng_mode == 0) return ranmar_setRandom();
    return omc_philox_next(&g_philox);
}

/* ---- (2) scoring wrapper ----------------------------------------------------------------- */
void ausgab(double edep) {
    g_ndeposit++;
    g_edep_sum += stack.wt[stack.np] * edep;
    omc_dosxyz_reference_ausgab(edep);
}

/* ---- life cycle -------------------------------------------------------------------------- */
static int g_ready = 0;

static void per_thread_init(void) {
#ifdef _OPENMP
#pragma omp parallel
#endif
    {
        ranmar_initRandom();
        initStack();
    }
}

int ref_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void ref_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

#ifndef OMC_REF_MATRAD
/* init chain of main(), omc_dosxyz.c:1155-1191 */
int ref_init_from_inp(const char *inp_stem) {
    char *stem = strdup(inp_stem);
    input_idx = 0;
    parseInputFile(stem);
    free(stem);
    initPhantom();
    initMediaData();
    initSource();
    initRegions();
    initVrt();
    initScore();
    per_thread_init();
    g_ready = 1;
    return 0;
}

#endif

void ref_set_rng(int mode, int seed0, int seed1) {
    g_rng_mode = mode;
    g_seed0 = (uint32_t)seed0;
    g_seed1 = (uint32_t)seed1;
}

void ref_set_nsplit(int nsplit) { vrt.nsplit = nsplit; }
int ref_nreg(void) { return geometry.isize * geometry.jsize * geometry.ksize + 1; }

/* ---- (3) problem dump / load -------------------------------------------------------------- */
#define NG (media.nmed * MXGE)
#define NE (media.nmed * MXEKE)
#define ELECTRON_ARRAYS(X)                                                                        \
    X(esig0) X(esig1) X(psig0) X(psig1) X(ededx0) X(ededx1) X(pdedx0) X(pdedx1) X(ebr10) X(ebr11)  \
    X(pbr10) X(pbr11) X(pbr20) X(pbr21) X(tmxs0) X(tmxs1) X(blcce0) X(blcce1) X(etae_ms0)          \
    X(etae_ms1) X(etap_ms0) X(etap_ms1) X(q1ce_ms0) X(q1ce_ms1) X(q1cp_ms0) X(q1cp_ms1)            \
    X(q2ce_ms0) X(q2ce_ms1) X(q2cp_ms0) X(q2cp_ms1)

#ifndef OMC_REF_MATRAD
int ref_dump_problem(const char *path) {
    static omc_blob b;
    b.n = 0;
    int nmed = media.nmed, nreg = ref_nreg();
    omc_blob_add_iscalar(&b, "nmed", nmed);
    /* photon */
    omc_blob_add_f64(&b, "ge0", nmed, photon_data.ge0);     omc_blob_add_f64(&b, "ge1", nmed, photon_data.ge1);
    omc_blob_add_f64(&b, "gmfp0", NG, photon_data.gmfp0);   omc_blob_add_f64(&b, "gmfp1", NG, photon_data.gmfp1);
    omc_blob_add_f64(&b, "gbr10", NG, photon_data.gbr10);   omc_blob_add_f64(&b, "gbr11", NG, photon_data.gbr11);
    omc_blob_add_f64(&b, "gbr20", NG, photon_data.gbr20);   omc_blob_add_f64(&b, "gbr21", NG, photon_data.gbr21);
    omc_blob_add_f64(&b, "cohe0", NG, photon_data.cohe0);   omc_blob_add_f64(&b, "cohe1", NG, photon_data.cohe1);
    /* rayleigh */
    omc_blob_add_f64(&b, "ray_xgrid", nmed * MXRAYFF, rayleigh_data.xgrid);
    omc_blob_add_f64(&b, "ray_fcum", nmed * MXRAYFF, rayleigh_data.fcum);
    omc_blob_add_f64(&b, "ray_b_array", nmed * MXRAYFF, rayleigh_data.b_array);
    omc_blob_add_f64(&b, "ray_c_array", nmed * MXRAYFF, rayleigh_data.c_array);
    omc_blob_add_i32(&b, "ray_i_array", nmed * RAYCDFSIZE, rayleigh_data.i_array);
    omc_blob_add_f64(&b, "ray_pmax0", NG, rayleigh_data.pmax0);
    omc_blob_add_f64(&b, "ray_pmax1", NG, rayleigh_data.pmax1);
    /* pair */
    omc_blob_add_f64(&b, "dl1", nmed * 8, pair_data.dl1); omc_blob_add_f64(&b, "dl2", nmed * 8, pair_data.dl2);
    omc_blob_add_f64(&b, "dl3", nmed * 8, pair_data.dl3); omc_blob_add_f64(&b, "dl4", nmed * 8, pair_data.dl4);
    omc_blob_add_f64(&b, "dl5", nmed * 8, pair_data.dl5); omc_blob_add_f64(&b, "dl6", nmed * 8, pair_data.dl6);
    omc_blob_add_f64(&b, "bpar0", nmed, pair_data.bpar0); omc_blob_add_f64(&b, "bpar1", nmed, pair_data.bpar1);
    omc_blob_add_f64(&b, "delcm", nmed, pair_data.delcm); omc_blob_add_f64(&b, "zbrang", nmed, pair_data.zbrang);
    /* electron */
#define X(a) omc_blob_add_f64(&b, #a, NE, electron_data.a);
    ELECTRON_ARRAYS(X)
#undef X
    omc_blob_add_f64(&b, "range_ep", 2 * NE, electron_data.range_ep);
    omc_blob_add_f64(&b, "e_array", NE, electron_data.e_array);
    omc_blob_add_f64(&b, "eke0", nmed, electron_data.eke0);   omc_blob_add_f64(&b, "eke1", nmed, electron_data.eke1);
    omc_blob_add_i32(&b, "sig_ismonotone", 2 * nmed, electron_data.sig_ismonotone);
    omc_blob_add_f64(&b, "esig_e", nmed, electron_data.esig_e); omc_blob_add_f64(&b, "psig_e", nmed, electron_data.psig_e);
    omc_blob_add_f64(&b, "xcc", nmed, electron_data.xcc);     omc_blob_add_f64(&b, "blcc", nmed, electron_data.blcc);
    /* spin */
    omc_blob_add_scalar(&b, "b2spin_min", spin_data.b2spin_min); omc_blob_add_scalar(&b, "dbeta2i", spin_data.dbeta2i);
    omc_blob_add_scalar(&b, "espml", spin_data.espml);           omc_blob_add_scalar(&b, "dleneri", spin_data.dleneri);
    omc_blob_add_scalar(&b, "dqq1i", spin_data.dqq1i);
    omc_blob_add_f64(&b, "spin_rej", (uint64_t)nmed * 2 * (MXE_SPIN1 + 1) * (MXQ_SPIN + 1) * (MXU_SPIN + 1), spin_data.spin_rej);
    /* mscat */
    int nms = (MXL_MS + 1) * (MXQ_MS + 1) * (MXU_MS + 1);
    omc_blob_add_f64(&b, "ums", nms, mscat_data.ums_array); omc_blob_add_f64(&b, "fms", nms, mscat_data.fms_array);
    omc_blob_add_f64(&b, "wms", nms, mscat_data.wms_array); omc_blob_add_i32(&b, "ims", nms, mscat_data.ims_array);
    omc_blob_add_scalar(&b, "dllambi", mscat_data.dllambi); omc_blob_add_scalar(&b, "dqmsi", mscat_data.dqmsi);
    /* pegs */
    omc_blob_add_f64(&b, "pegs_ap", nmed, pegs_data.ap); omc_blob_add_f64(&b, "pegs_ae", nmed, pegs_data.ae);
    omc_blob_add_f64(&b, "pegs_te", nmed, pegs_data.te); omc_blob_add_f64(&b, "pegs_thmoll", nmed, pegs_data.thmoll);
    omc_blob_add_f64(&b, "pegs_rho", nmed, pegs_data.rho); omc_blob_add_i32(&b, "pegs_meke", nmed, pegs_data.meke);
    /* geometry + regions */
    int nvox = nreg - 1;
    omc_blob_add_iscalar(&b, "isize", geometry.isize); omc_blob_add_iscalar(&b, "jsize", geometry.jsize);
    omc_blob_add_iscalar(&b, "ksize", geometry.ksize);
    omc_blob_add_f64(&b, "xbounds", geometry.isize + 1, geometry.xbounds);
    omc_blob_add_f64(&b, "ybounds", geometry.jsize + 1, geometry.ybounds);
    omc_blob_add_f64(&b, "zbounds", geometry.ksize + 1, geometry.zbounds);
    omc_blob_add_i32(&b, "med_indices", nvox, geometry.med_indices);
    omc_blob_add_f64(&b, "med_densities", nvox, geometry.med_densities);
    omc_blob_add_i32(&b, "region_med", nreg, region.med);   omc_blob_add_f64(&b, "region_rhof", nreg, region.rhof);
    omc_blob_add_f64(&b, "region_pcut", nreg, region.pcut); omc_blob_add_f64(&b, "region_ecut", nreg, region.ecut);
    /* source */
    omc_blob_add_iscalar(&b, "src_spectrum", source.spectrum); omc_blob_add_iscalar(&b, "src_charge", source.charge);
    omc_blob_add_scalar(&b, "src_energy", source.spectrum ? 0.0 : source.energy);
    omc_blob_add_scalar(&b, "src_deltak", source.spectrum ? source.deltak : 0.0);
    if (source.spectrum) {
        omc_blob_add_f64(&b, "src_cdfinv1", (uint64_t)source.deltak, source.cdfinv1);
        omc_blob_add_f64(&b, "src_cdfinv2", (uint64_t)source.deltak, source.cdfinv2);
    } else {
        double z = 0.0;
        omc_blob_add_f64(&b, "src_cdfinv1", 1, &z); omc_blob_add_f64(&b, "src_cdfinv2", 1, &z);
    }
    omc_blob_add_scalar(&b, "src_ssd", source.ssd);
    omc_blob_add_scalar(&b, "src_xinl", source.xinl); omc_blob_add_scalar(&b, "src_xinu", source.xinu);
    omc_blob_add_scalar(&b, "src_yinl", source.yinl); omc_blob_add_scalar(&b, "src_yinu", source.yinu);
    omc_blob_add_scalar(&b, "src_xsize", source.xsize); omc_blob_add_scalar(&b, "src_ysize", source.ysize);
    omc_blob_add_iscalar(&b, "src_ixinl", source.ixinl); omc_blob_add_iscalar(&b, "src_ixinu", source.ixinu);
    omc_blob_add_iscalar(&b, "src_iyinl", source.iyinl); omc_blob_add_iscalar(&b, "src_iyinu", source.iyinu);
    omc_blob_add_iscalar(&b, "nsplit", vrt.nsplit);
    int rc = omc_blob_write(&b, path);
    omc_blob_free(&b);
    return rc;
}

#endif

static double *dup_f64(const omc_blob *b, const char *name) {
    const omc_blob_entry *e = omc_blob_find(b, name);
    double *p = malloc((e->count ? e->count : 1) * sizeof(double));
    memcpy(p, e->data, e->count * sizeof(double));
    return p;
}
static int *dup_i32(const omc_blob *b, const char *name) {
    const omc_blob_entry *e = omc_blob_find(b, name);
    int *p = malloc((e->count ? e->count : 1) * sizeof(int));
    memcpy(p, e->data, e->count * sizeof(int));
    return p;
}
static double sc(const omc_blob *b, const char *name) { return omc_blob_f64(b, name)[0]; }
static int isc(const omc_blob *b, const char *name) { return omc_blob_i32(b, name)[0]; }

/* Fill the reference's globals from a blob instead of running its file-reading init chain. */
int ref_load_problem(const char *path) {
    static omc_blob b;
    if (omc_blob_read(&b, path) != 0) return -1;
    int nmed = isc(&b, "nmed");
    media.nmed = nmed;
    photon_data.ge0 = dup_f64(&b, "ge0");     photon_data.ge1 = dup_f64(&b, "ge1");
    photon_data.gmfp0 = dup_f64(&b, "gmfp0"); photon_data.gmfp1 = dup_f64(&b, "gmfp1");
    photon_data.gbr10 = dup_f64(&b, "gbr10"); photon_data.gbr11 = dup_f64(&b, "gbr11");
    photon_data.gbr20 = dup_f64(&b, "gbr20"); photon_data.gbr21 = dup_f64(&b, "gbr21");
    photon_data.cohe0 = dup_f64(&b, "cohe0"); photon_data.cohe1 = dup_f64(&b, "cohe1");
    rayleigh_data.xgrid = dup_f64(&b, "ray_xgrid");     rayleigh_data.fcum = dup_f64(&b, "ray_fcum");
    rayleigh_data.b_array = dup_f64(&b, "ray_b_array"); rayleigh_data.c_array = dup_f64(&b, "ray_c_array");
    rayleigh_data.i_array = dup_i32(&b, "ray_i_array");
    rayleigh_data.pmax0 = dup_f64(&b, "ray_pmax0");     rayleigh_data.pmax1 = dup_f64(&b, "ray_pmax1");
    pair_data.dl1 = dup_f64(&b, "dl1"); pair_data.dl2 = dup_f64(&b, "dl2"); pair_data.dl3 = dup_f64(&b, "dl3");
    pair_data.dl4 = dup_f64(&b, "dl4"); pair_data.dl5 = dup_f64(&b, "dl5"); pair_data.dl6 = dup_f64(&b, "dl6");
    pair_data.bpar0 = dup_f64(&b, "bpar0"); pair_data.bpar1 = dup_f64(&b, "bpar1");
    pair_data.delcm = dup_f64(&b, "delcm"); pair_data.zbrang = dup_f64(&b, "zbrang");
#define X(a) electron_data.a = dup_f64(&b, #a);
    ELECTRON_ARRAYS(X)
#undef X
    electron_data.range_ep = dup_f64(&b, "range_ep"); electron_data.e_array = dup_f64(&b, "e_array");
    electron_data.eke0 = dup_f64(&b, "eke0");         electron_data.eke1 = dup_f64(&b, "eke1");
    electron_data.sig_ismonotone = dup_i32(&b, "sig_ismonotone");
    electron_data.esig_e = dup_f64(&b, "esig_e");     electron_data.psig_e = dup_f64(&b, "psig_e");
    electron_data.xcc = dup_f64(&b, "xcc");           electron_data.blcc = dup_f64(&b, "blcc");
    electron_data.expeke1 = calloc(nmed, sizeof(double));
    spin_data.b2spin_min = sc(&b, "b2spin_min"); spin_data.dbeta2i = sc(&b, "dbeta2i");
    spin_data.espml = sc(&b, "espml");           spin_data.dleneri = sc(&b, "dleneri");
    spin_data.dqq1i = sc(&b, "dqq1i");           spin_data.spin_rej = dup_f64(&b, "spin_rej");
    mscat_data.ums_array = dup_f64(&b, "ums"); mscat_data.fms_array = dup_f64(&b, "fms");
    mscat_data.wms_array = dup_f64(&b, "wms"); mscat_data.ims_array = dup_i32(&b, "ims");
    mscat_data.dllambi = sc(&b, "dllambi");    mscat_data.dqmsi = sc(&b, "dqmsi");
    for (int i = 0; i < nmed; i++) {
        pegs_data.ap[i] = omc_blob_f64(&b, "pegs_ap")[i]; pegs_data.ae[i] = omc_blob_f64(&b, "pegs_ae")[i];
        pegs_data.te[i] = omc_blob_f64(&b, "pegs_te")[i]; pegs_data.thmoll[i] = omc_blob_f64(&b, "pegs_thmoll")[i];
        pegs_data.rho[i] = omc_blob_f64(&b, "pegs_rho")[i]; pegs_data.meke[i] = omc_blob_i32(&b, "pegs_meke")[i];
    }
    geometry.isize = isc(&b, "isize"); geometry.jsize = isc(&b, "jsize"); geometry.ksize = isc(&b, "ksize");
    geometry.xbounds = dup_f64(&b, "xbounds"); geometry.ybounds = dup_f64(&b, "ybounds");
    geometry.zbounds = dup_f64(&b, "zbounds");
    geometry.med_indices = dup_i32(&b, "med_indices"); geometry.med_densities = dup_f64(&b, "med_densities");
    region.med = dup_i32(&b, "region_med");   region.rhof = dup_f64(&b, "region_rhof");
    region.pcut = dup_f64(&b, "region_pcut"); region.ecut = dup_f64(&b, "region_ecut");
    source.spectrum = isc(&b, "src_spectrum"); source.charge = isc(&b, "src_charge");
    source.energy = sc(&b, "src_energy");      source.deltak = sc(&b, "src_deltak");
    source.cdfinv1 = dup_f64(&b, "src_cdfinv1"); source.cdfinv2 = dup_f64(&b, "src_cdfinv2");
#ifdef OMC_REF_MATRAD
    source.nbeamlets = isc(&b, "mr_nbeamlets");
    source.ibeam = dup_i32(&b, "mr_ibeam");
    source.xsource = dup_f64(&b, "mr_xsource"); source.ysource = dup_f64(&b, "mr_ysource"); source.zsource = dup_f64(&b, "mr_zsource");
    source.xcorner = dup_f64(&b, "mr_xcorner"); source.ycorner = dup_f64(&b, "mr_ycorner"); source.zcorner = dup_f64(&b, "mr_zcorner");
    source.xside1 = dup_f64(&b, "mr_xside1"); source.yside1 = dup_f64(&b, "mr_yside1"); source.zside1 = dup_f64(&b, "mr_zside1");
    source.xside2 = dup_f64(&b, "mr_xside2"); source.yside2 = dup_f64(&b, "mr_yside2"); source.zside2 = dup_f64(&b, "mr_zside2");
#else
    source.ssd = sc(&b, "src_ssd");
    source.xinl = sc(&b, "src_xinl"); source.xinu = sc(&b, "src_xinu");
    source.yinl = sc(&b, "src_yinl"); source.yinu = sc(&b, "src_yinu");
    source.xsize = sc(&b, "src_xsize"); source.ysize = sc(&b, "src_ysize");
    source.ixinl = isc(&b, "src_ixinl"); source.ixinu = isc(&b, "src_ixinu");
    source.iyinl = isc(&b, "src_iyinl"); source.iyinu = isc(&b, "src_iyinu");
#endif
    vrt.nsplit = isc(&b, "nsplit");
    omc_blob_free(&b);
    initScore();
    /* ranmar_initRandom() reads "rng seeds" through getInputValue(): provide it */
    input_idx = 1;
    strcpy(input_items[0].key, "rng seeds ");
    snprintf(input_items[0].value, BUFFER_SIZE, " %u %u", g_seed0, g_seed1);
    strcpy(input_items[1].key, "unused ");
    strcpy(input_items[1].value, " 0");
    per_thread_init();
    g_ready = 1;
    return 0;
}

/* ---- (4) the batch loop ------------------------------------------------------------------- */
/*
 * {initHistory(); shower();} for history ids [first, first+n) -- omc_dosxyz.c:1252-1259.
 * Philox mode re-keys the stream per history id; RANMAR mode continues the thread's sequence
 * exactly as the reference does.  rec (nullable) receives one record per history.
 */
void ref_run_histories(long long first, long long n, omc_history_record *rec) {
    long long i;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic)
#endif
    for (i = 0; i < n; i++) {
        if (g_rng_mode == 1) omc_philox_seed(&g_philox, g_seed0, g_seed1, (uint64_t)(first + i), 0);
        g_ndeposit = 0;
        g_edep_sum = 0.0;
        INIT_HISTORY();
        int ir0 = stack.ir[0];
        shower();
        if (rec) {
            rec[i].ndraws = (unsigned int)g_philox.ndraws;
            rec[i].ir_start = ir0;
            rec[i].ndeposit = g_ndeposit;
            rec[i].flags = 0;
            rec[i].edep = g_edep_sum;
        }
    }
}

void ref_accum_endep(void) { accumEndep(); }

void ref_reset_score(void) {
    size_t n = (size_t)ref_nreg() * sizeof(double);
    memset(score.endep, 0, n); memset(score.accum_endep, 0, n); memset(score.accum_endep2, 0, n);
    score.ensrc = 0.0;
}

void ref_set_endep(const double *in) { memcpy(score.endep, in, (size_t)ref_nreg() * sizeof(double)); }
void ref_get_endep(double *out) { memcpy(out, score.endep, (size_t)ref_nreg() * sizeof(double)); }
void ref_get_accum(double *a, double *a2, double *ensrc) {
    size_t n = (size_t)ref_nreg() * sizeof(double);
    if (a) memcpy(a, score.accum_endep, n);
    if (a2) memcpy(a2, score.accum_endep2, n);
    if (ensrc) *ensrc = score.ensrc;
}
/* accumulateResults() in place (omc_dosxyz.c:719-799): accum -> dose or energy, accum2 -> rel. sigma */
void ref_accumulate_results(int iout, int nhist, int nbatch) { accumulateResults(iout, nhist, nbatch); }
#ifndef OMC_REF_MATRAD
int ref_write_3ddose(const char *stem, int nperbatch, int nbatch) {
    /* outputResults() (omc_dosxyz.c:801-886) needs "output folder"; stem is taken relative to cwd */
    strcpy(input_items[1].key, "output folder ");
    strcpy(input_items[1].value, " ./");
    if (input_idx < 1) input_idx = 1;
    char *s = strdup(stem);
    outputResults(s, 1, nperbatch, nbatch);
    free(s);
    return 0;
}

#endif

/* timing helper: the whole batch loop, returns wall seconds */
double ref_time_batches(long long first, long long nperbatch, int nbatch) {
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int ib = 0; ib < nbatch; ib++) {
        ref_run_histories(first + ib * nperbatch, nperbatch, NULL);
        accumEndep();
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

/* ---- unit hooks: reference geometry callbacks on explicit inputs -------------------------- */
void ref_test_geometry(int n, const double *xyzuvw, const int *ir, const double *ustep_in, int *idisc,
                       int *irnew, double *ustep_out, double *tperp) {
    for (int i = 0; i < n; i++) {
        stack.np = 0;
        stack.x[0] = xyzuvw[6 * i + 0]; stack.y[0] = xyzuvw[6 * i + 1]; stack.z[0] = xyzuvw[6 * i + 2];
        stack.u[0] = xyzuvw[6 * i + 3]; stack.v[0] = xyzuvw[6 * i + 4]; stack.w[0] = xyzuvw[6 * i + 5];
        stack.ir[0] = ir[i];
        int id = 0, irn = ir[i];
        double us = ustep_in[i];
        howfar(&id, &irn, &us);
        idisc[i] = id; irnew[i] = irn; ustep_out[i] = us;
        tperp[i] = hownear();
    }
}

/* n raw draws of history `hist` in Philox mode */
void ref_test_rng(long long hist, int n, double *out) {
    omc_philox g;
    omc_philox_seed(&g, g_seed0, g_seed1, (uint64_t)hist, 0);
    for (int i = 0; i < n; i++) out[i] = omc_philox_next(&g);
}

/*
 * Run one top-of-stack particle through shower() with the Philox stream of `hist` and report
 * what came out: used by sampler-level unit tests (a single photon / electron of chosen energy in
 * a chosen region).  Returns deposits through the normal score grid.
 */
void ref_run_particle(long long hist, int iq, double e, const double *xyzuvw, int ir, double wt,
                      omc_history_record *rec) {
    omc_philox_seed(&g_philox, g_seed0, g_seed1, (uint64_t)hist, 0);
    int mode = g_rng_mode;
    g_rng_mode = 1;
    g_ndeposit = 0; g_edep_sum = 0.0;
    stack.np = 0;
    stack.iq[0] = iq; stack.e[0] = e; stack.ir[0] = ir; stack.wt[0] = wt; stack.dnear[0] = 0.0;
    stack.x[0] = xyzuvw[0]; stack.y[0] = xyzuvw[1]; stack.z[0] = xyzuvw[2];
    stack.u[0] = xyzuvw[3]; stack.v[0] = xyzuvw[4]; stack.w[0] = xyzuvw[5];
    shower();
    if (rec) {
        rec->ndraws = (unsigned int)g_philox.ndraws; rec->ir_start = ir; rec->ndeposit = g_ndeposit;
        rec->flags = 0; rec->edep = g_edep_sum;
    }
    g_rng_mode = mode;
}

/* ---- unit hooks: the reference's OWN sampler functions on explicit inputs ------------------ */
/*
 * Counterpart of omc_gpu_test_samplers() (include/ompmc_b200.h): record i = 8 doubles in, 8 doubles out, random
 * numbers from the Philox stream of history first + i drawn one by one as the reference draws them.  Every number
 * comes out of the unmodified functions of src/ompmc.c; this wrapper only sets the top of the stack.
 */
static double ref_range(int imed, int iq, double eke, double rhof, double *elke_out, int *lelke_out) {
    /* electron() src/ompmc.c:4905-4918 */
    int qel = (1 + iq) / 2;
    double elke = log(eke);
    int lelke = pwlfInterval(imed, elke, electron_data.eke1, electron_data.eke0) - 1;
    double ekei = electron_data.e_array[imed * MXEKE + lelke];
    double elkei = (lelke + 1 - electron_data.eke0[imed]) / electron_data.eke1[imed];
    double range = computeDrange(imed, iq, lelke, eke, ekei, elke, elkei);
    range += electron_data.range_ep[qel * media.nmed * MXEKE + imed * MXEKE + lelke];
    *elke_out = elke; *lelke_out = lelke;
    return range / rhof;
}

void ref_test_samplers(int which, int n, const double *in, long long first, double *out) {
    int mode = g_rng_mode;
    g_rng_mode = 1;
    for (int i = 0; i < n; i++) {
        const double *a = in + 8 * (size_t)i;
        double *o = out + 8 * (size_t)i;
        for (int k = 0; k < 8; k++) o[k] = 0.0;
        omc_philox_seed(&g_philox, g_seed0, g_seed1, (uint64_t)(first + i), 0);
        stack.np = 0; stack.npold = 0;
        stack.x[0] = stack.y[0] = stack.z[0] = 0.0; stack.wt[0] = 1.0; stack.ir[0] = 1; stack.dnear[0] = 0.0;
        if (which == OMC_SAMPLER_DRANGE) {
            int imed = (int)a[0], iq = (int)a[1];
            double elkei = log(a[2]), elkef = log(a[3]);
            int lelke = pwlfInterval(imed, elkei, electron_data.eke1, electron_data.eke0) - 1;
            o[0] = computeDrange(imed, iq, lelke, a[2], a[3], elkei, elkef);
        } else if (which == OMC_SAMPLER_ELOSS || which == OMC_SAMPLER_MSDIST) {
            int imed = (int)a[0], iq = (int)a[1], lelke;
            double rhof = a[2], eke = a[3], elke;
            double range = ref_range(imed, iq, eke, rhof, &elke, &lelke);
            double tustep = a[4] * range;
            double de = computeEloss(imed, iq, 1, rhof, tustep, range, eke, elke, lelke);
            if (which == OMC_SAMPLER_ELOSS) { o[0] = range; o[1] = de; continue; }
            stack.iq[0] = iq; stack.e[0] = eke + RM;
            stack.u[0] = a[5]; stack.v[0] = a[6]; stack.w[0] = a[7];
            o[0] = msdist(imed, iq, rhof, de, tustep, eke, &o[1], &o[2], &o[3], &o[4], &o[5], &o[6]);
            o[7] = de;
        } else if (which == OMC_SAMPLER_SSCAT) {
            double cphi, sphi;
            sscat((int)a[0], (int)a[1], a[2], a[3], a[4], &o[0], &o[1]);
            selectAzimuthalAngle(&cphi, &sphi);
            o[2] = cphi; o[3] = sphi;
        } else if (which == OMC_SAMPLER_COMPTON) {
            stack.iq[0] = 0; stack.e[0] = a[0]; stack.u[0] = a[1]; stack.v[0] = a[2]; stack.w[0] = a[3];
            compton();
            /* the reference leaves the scattered photon and the electron on the stack (order not relied upon) */
            for (int k = 0; k <= stack.np; k++) {
                int off = (stack.iq[k] == 0) ? 0 : 4;
                o[off] = stack.e[k]; o[off + 1] = stack.u[k]; o[off + 2] = stack.v[k]; o[off + 3] = stack.w[k];
            }
        } else if (which == OMC_SAMPLER_MOLLER) {
            /* moller() reads the medium from region.med[stack.ir]: find a region filled with it */
            int imed = (int)a[0], ir = -1, nreg = ref_nreg();
            for (int r = 1; r < nreg; r++) if (region.med[r] == imed) { ir = r; break; }
            if (ir < 0) continue;
            stack.ir[0] = ir; stack.iq[0] = -1; stack.e[0] = a[1]; stack.u[0] = a[2]; stack.v[0] = a[3]; stack.w[0] = a[4];
            moller();
            /* higher-energy electron first, as the production sampler reports them */
            int hi = 0, lo = -1;
            if (stack.np == 1) { hi = (stack.e[0] >= stack.e[1]) ? 0 : 1; lo = 1 - hi; }
            o[0] = stack.e[hi]; o[1] = stack.u[hi]; o[2] = stack.v[hi]; o[3] = stack.w[hi];
            if (lo >= 0) { o[4] = stack.e[lo]; o[5] = stack.u[lo]; o[6] = stack.v[lo]; o[7] = stack.w[lo]; }
        }
    }
    g_rng_mode = mode;
}

#ifndef OMC_REF_MATRAD
/*
 * Optical depth of the straight photon path of length s from (x,y,z) along (u,v,w): the transport loop of photon()
 * src/ompmc.c:1951-2019 (the reference's howfar(), its gmfp / Rayleigh-correction tables and density scaling) with the
 * mean-free-path budget left out.  in: n records {e, x, y, z, u, v, w, s}; out: n records {tau, region at the end (0 = outside)}.
 * Checker of the Woodcock photon flight of the production kernels: the interaction depth must be exponential in tau.
 */
void ref_test_photon_tau(int n, const double *in, double *out) {
    for (int i = 0; i < n; i++) {
        const double *a = in + 8 * (size_t)i;
        double gle = log(a[0]), left = a[7], tau = 0.0;
        int ix = 0, iy = 0, iz = 0;
        while (ix < geometry.isize - 1 && geometry.xbounds[ix + 1] <= a[1]) ix++;
        while (iy < geometry.jsize - 1 && geometry.ybounds[iy + 1] <= a[2]) iy++;
        while (iz < geometry.ksize - 1 && geometry.zbounds[iz + 1] <= a[3]) iz++;
        int irl = 1 + ix + iy * geometry.isize + iz * geometry.isize * geometry.jsize;
        stack.np = 0;
        stack.x[0] = a[1]; stack.y[0] = a[2]; stack.z[0] = a[3]; stack.u[0] = a[4]; stack.v[0] = a[5]; stack.w[0] = a[6];
        stack.ir[0] = irl; stack.iq[0] = 0; stack.e[0] = a[0]; stack.wt[0] = 1.0;
        while (left > 0.0) {
            int imed = region.med[irl];
            double sig = 0.0;
            if (imed != -1) {
                int lgle = pwlfInterval(imed, gle, photon_data.ge1, photon_data.ge0) - 1;
                double gmfp = pwlfEval(imed * MXGE + lgle, gle, photon_data.gmfp1, photon_data.gmfp0) / region.rhof[irl];
                gmfp *= pwlfEval(imed * MXGE + lgle, gle, photon_data.cohe1, photon_data.cohe0);
                sig = 1.0 / gmfp;
            }
            int idisc = 0, irnew = irl;
            double ustep = left;
            howfar(&idisc, &irnew, &ustep);
            if (idisc > 0) { irl = 0; break; }
            stack.x[0] += ustep * stack.u[0]; stack.y[0] += ustep * stack.v[0]; stack.z[0] += ustep * stack.w[0];
            tau += ustep * sig;
            left -= ustep;
            if (irnew != irl) { irl = irnew; stack.ir[0] = irl; if (irl == 0) break; }
        }
        out[2 * i] = tau; out[2 * i + 1] = (double)irl;
    }
}
#endif

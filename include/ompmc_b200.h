/*
 * ompmc_b200.h -- C-ABI of the B200-native replacement for ompMC's shower() hot path.
 *
 * The reference (edoerner/ompMC) has no plugin/FFI layer: its "API" is link-time C symbols between
 * a user code (ucodes/omc_dosxyz/omc_dosxyz.c, ucodes/omc_matrad/omc_matrad.c) and src/ompmc.c plus
 * shared global structs.  Per-step host callbacks (howfar/hownear/ausgab, src/ompmc.h:37-39) cannot
 * cross a device boundary, so the drop-in boundary sits at the *batch loop*
 * (omc_dosxyz.c:1237-1263, omc_matrad.c:1389-1414):
 *
 *     for ibatch { omp parallel for ihist { initHistory(); shower(); }  accumEndep(); }
 *
 * Everything above that loop (ini parsing, .egsphant reader, table initialisation, statistics,
 * .3ddose writer) stays host C and hands its already-built global structs to this library as
 * plain pointers.  No torch / C++ types appear in any signature.  All functions return 0 on
 * success and a non-zero code otherwise; omc_gpu_last_error() returns the message the reference
 * would have printf'ed before exit(EXIT_FAILURE).
 *
 * Array layouts are exactly the reference's (same strides, same 0-based C indexing), so a user
 * code can pass e.g. photon_data.gmfp0 directly.
 */
#ifndef OMPMC_B200_H
#define OMPMC_B200_H

#ifdef __cplusplus
extern "C" {
#endif

/* compile-time table dimensions, src/ompmc.h:104,129-130,256-259,283-285,325-326 */
#define OMC_MXGE       2000   /* MXGE    */
#define OMC_MXEKE      500    /* MXEKE   */
#define OMC_MXRAYFF    100    /* MXRAYFF == RAYCDFSIZE */
#define OMC_MXMED      9      /* MXMED   */
#define OMC_SPIN_NE    32     /* MXE_SPIN1+1 */
#define OMC_SPIN_NQ    16     /* MXQ_SPIN+1  */
#define OMC_SPIN_NU    32     /* MXU_SPIN+1  */
#define OMC_MS_NL      64     /* MXL_MS+1 */
#define OMC_MS_NQ      8      /* MXQ_MS+1 */
#define OMC_MS_NU      32     /* MXU_MS+1 */
#define OMC_RM         0.5109989461   /* src/ompmc.h:46 */

/*
 * Physics tables == the output of initMediaData() (src/ompmc.c:5450-5482).
 * Field <-> reference global:
 *   ge*..cohe*     struct Photon   photon_data    (src/ompmc.h:105-112)
 *   ray_*          struct Rayleigh rayleigh_data  (src/ompmc.h:134-143)
 *   dl1..zbrang    struct Pair     pair_data      (src/ompmc.h:155-167)
 *   esig0..blcc    struct Electron electron_data  (src/ompmc.h:207-249)
 *   spin_*         struct Spin     spin_data      (src/ompmc.h:261-269)
 *   ums..dqmsi     struct Mscat    mscat_data     (src/ompmc.h:292-300)
 *   pegs_*         struct Pegs     pegs_data      (src/ompmc.h:379-404)
 */
typedef struct omc_media_tables {
    int nmed;
    /* photon, [nmed] and [nmed*OMC_MXGE] */
    const double *ge0, *ge1;
    const double *gmfp0, *gmfp1, *gbr10, *gbr11, *gbr20, *gbr21, *cohe0, *cohe1;
    /* Rayleigh, [nmed*OMC_MXRAYFF] / [nmed*OMC_MXGE] */
    const double *ray_xgrid, *ray_fcum, *ray_b_array, *ray_c_array;
    const int    *ray_i_array;
    const double *ray_pmax0, *ray_pmax1;
    /* pair + brems screening, [nmed*8] and [nmed] */
    const double *dl1, *dl2, *dl3, *dl4, *dl5, *dl6;
    const double *bpar0, *bpar1, *delcm, *zbrang;
    /* electron PWL tables, [nmed*OMC_MXEKE] */
    const double *esig0, *esig1, *psig0, *psig1, *ededx0, *ededx1, *pdedx0, *pdedx1;
    const double *ebr10, *ebr11, *pbr10, *pbr11, *pbr20, *pbr21, *tmxs0, *tmxs1;
    const double *blcce0, *blcce1, *etae_ms0, *etae_ms1, *etap_ms0, *etap_ms1;
    const double *q1ce_ms0, *q1ce_ms1, *q1cp_ms0, *q1cp_ms1;
    const double *q2ce_ms0, *q2ce_ms1, *q2cp_ms0, *q2cp_ms1;
    const double *range_ep;        /* [2*nmed*OMC_MXEKE], qel-major */
    const double *e_array;         /* [nmed*OMC_MXEKE] */
    const double *eke0, *eke1;     /* [nmed] */
    const int    *sig_ismonotone;  /* [2*nmed], qel-major */
    const double *esig_e, *psig_e, *xcc, *blcc;   /* [nmed] */
    /* spin (Mott) rejection, spin_rej[nmed][2][32][16][32] */
    double b2spin_min, dbeta2i, espml, dleneri, dqq1i;
    const double *spin_rej;
    /* screened-Rutherford MS alias tables [64][8][32] */
    const double *ums, *fms, *wms;
    const int    *ims;
    double dllambi, dqmsi;
    /* PEGS4 scalars, [nmed] */
    const double *pegs_ap, *pegs_ae, *pegs_te, *pegs_thmoll, *pegs_rho;
    const int    *pegs_meke;
} omc_media_tables;

/*
 * Voxel geometry + per-region transport data: struct Geom (omc_dosxyz.c:46-59) and
 * struct Region (src/ompmc.h:412-418).  nreg = isize*jsize*ksize + 1; region 0 = outside.
 */
typedef struct omc_geometry {
    int isize, jsize, ksize;
    const double *xbounds, *ybounds, *zbounds;   /* [isize+1],[jsize+1],[ksize+1] */
    const int    *med;                           /* [nreg] 0-based medium, -1 = vacuum */
    const double *rhof, *pcut, *ecut;            /* [nreg] */
} omc_geometry;

/* struct Source of omc_dosxyz.c:342-366 as filled by initSource() (:368-632). */
typedef struct omc_source_dosxyz {
    int spectrum;            /* 0 mono-energetic, 1 spectrum */
    int charge;              /* 0 photon, -1 e-, +1 e+ */
    double energy;           /* mono energy */
    double deltak;           /* number of inverse-CDF bins (INVDIM = 1000) */
    const double *cdfinv1, *cdfinv2;   /* [(int)deltak] */
    double ssd;
    double xinl, xinu, yinl, yinu, xsize, ysize;
    int ixinl, ixinu, iyinl, iyinu;
} omc_source_dosxyz;

/* struct Source of omc_matrad.c:514-541 (bixel arrays), as filled by initSource() (:543-751). */
typedef struct omc_source_matrad {
    int spectrum, charge;
    double energy, deltak;
    const double *cdfinv1, *cdfinv2;
    int nbeams, nbixels;
    const int    *ibeam;                         /* [nbixels] beam of each bixel */
    const double *xsource, *ysource, *zsource;   /* [nbeams]  */
    const double *xcorner, *ycorner, *zcorner;   /* [nbixels] */
    const double *xside1, *yside1, *zside1;      /* [nbixels] */
    const double *xside2, *yside2, *zside2;      /* [nbixels] */
} omc_source_matrad;

/* Per-history debug record (lock-step parity with the instrumented reference). */
typedef struct omc_history_record {
    unsigned int ndraws;     /* random numbers consumed by initHistory()+shower() */
    int          ir_start;   /* region index chosen by initHistory() */
    unsigned int ndeposit;   /* number of ausgab() calls */
    unsigned int flags;      /* bit0: stack overflow, bit1: invalid lambda drop (Q8) */
    double       edep;       /* sum of wt*edep over all ausgab() calls (incl. region 0) */
} omc_history_record;

/* Work counters of the last run (filled on the device). */
typedef struct omc_gpu_counters {
    unsigned long long histories;      /* histories started                           */
    unsigned long long kernel_launches;/* kernels of this library launched            */
    unsigned long long photon_steps;   /* howfar() calls from photon()                */
    unsigned long long electron_steps; /* hownear() calls (ustep-loop iterations)     */
    unsigned long long deposits;       /* ausgab() calls                              */
    unsigned long long rng_draws;      /* random numbers consumed                     */
    unsigned long long errors;         /* stack/queue overflows, invalid-lambda drops */
    unsigned long long reserved[9];    /* [0..3] drain diagnostics */
} omc_gpu_counters;

typedef struct omc_gpu_ctx *omc_gpu_handle;

/* kernel selection for omc_gpu_set_option("kernel", ...) */
#define OMC_KERNEL_LOCKSTEP  0   /* one history per thread, reference draw order (parity anchor) */
#define OMC_KERNEL_WAVEFRONT 1   /* particle-queue production kernels (nsplit <= 255)            */

/* ---- life cycle ------------------------------------------------------------------------- */
/* replaces initStack()/initRandom() per thread (omc_dosxyz.c:1184-1191) */
int  omc_gpu_create(omc_gpu_handle *h, int device_id);
/* replaces clean*() (omc_dosxyz.c:1285-1300) */
void omc_gpu_destroy(omc_gpu_handle h);
const char *omc_gpu_last_error(omc_gpu_handle h);

/* ---- problem upload (arrays are borrowed for the duration of the call and copied) -------- */
int omc_gpu_set_media(omc_gpu_handle h, const omc_media_tables *t);        /* initMediaData() output, src/ompmc.c:5450 */
int omc_gpu_set_geometry(omc_gpu_handle h, const omc_geometry *g);        /* initPhantom()+initRegions(), omc_dosxyz.c:62,890 */
int omc_gpu_set_source_dosxyz(omc_gpu_handle h, const omc_source_dosxyz *s); /* initSource(), omc_dosxyz.c:368 */
int omc_gpu_set_source_matrad(omc_gpu_handle h, const omc_source_matrad *s); /* initSource(), omc_matrad.c:543 */
int omc_gpu_set_vrt(omc_gpu_handle h, int nsplit);                         /* initVrt(), src/ompmc.c:5964 */
int omc_gpu_set_seed(omc_gpu_handle h, int ixx, int jxx);                  /* "rng seeds", src/omc_random.c:58-82 */
/* tuning / debug knobs: "kernel", "threads_per_block", "stack_depth", "pool_size", "record_histories", "drain_threshold"
 * (0: no drain kernel, bit-reproducible whatever the schedule) */
int omc_gpu_set_option(omc_gpu_handle h, const char *key, long long value);

/* ---- the hot path ------------------------------------------------------------------------ */
/*
 * {initHistory(); shower();} for history ids [first_history, first_history+nhist)
 * (omc_dosxyz.c:1252-1259; omc_matrad.c:1393-1400 when ibeamlet >= 0), scoring into the batch
 * grid score.endep.  History id -> RNG stream, so results do not depend on scheduling or on how
 * ids are split over GPUs.  Asynchronous: returns after enqueueing on the context's stream.
 */
int omc_gpu_run_histories(omc_gpu_handle h, long long first_history, long long nhist, int ibeamlet);
/* accumEndep() (omc_dosxyz.c:696-717): accum += e, accum2 += e*e, zero the batch grid. */
int omc_gpu_accum_batch(omc_gpu_handle h);
/* One iteration of the reference batch loop ({initHistory(); shower();} x nhist + accumEndep(), omc_dosxyz.c:1237-1263).
 * With the wavefront kernels consecutive calls are PIPELINED: the call returns when all its histories have been
 * started and every earlier batch is complete and accumulated; the tail of this batch keeps running underneath the
 * next call (each particle scores into the grid of the batch its history id belongs to) and is completed and
 * accumulated by the next call or by whichever call reads results (get_tallies, accumulate_results, synchronize...).
 * Pipelining attributes a particle in flight to its batch by its history id, so consecutive batches overlap only when
 * their id ranges ASCEND (first_history >= the end of the previous range, as in the reference's loop, where batch k owns
 * ids [k * nperbatch, (k+1) * nperbatch)); a call whose range starts lower first completes the batch in flight. */
int omc_gpu_run_batch(omc_gpu_handle h, long long first_history, long long nhist, int ibeamlet);
/* The same pipelining with the accumulation left to the caller (multi-GPU: the batch grid is summed over ranks before
 * accumEndep() squares it): start_batch returns when its histories are all started and the PREVIOUS started batch is
 * complete; omc_gpu_completed_batches() = how many complete batch grids wait for omc_gpu_accum_batch(), which takes
 * them oldest first (omc_gpu_device_ptrs / omc_gpu_get_batch_grid address the oldest one); omc_gpu_finish_batches
 * completes the batch in flight.  At most two batches may be waiting / in flight together. */
int omc_gpu_start_batch(omc_gpu_handle h, long long first_history, long long nhist, int ibeamlet);
int omc_gpu_finish_batches(omc_gpu_handle h);
int omc_gpu_completed_batches(omc_gpu_handle h);
int omc_gpu_synchronize(omc_gpu_handle h);

/* ---- results ----------------------------------------------------------------------------- */
/* score.accum_endep / score.accum_endep2 / score.ensrc (omc_dosxyz.c:636-645); each [nreg] fp64; any may be NULL */
int omc_gpu_get_tallies(omc_gpu_handle h, double *accum_endep, double *accum_endep2, double *ensrc);
/* accumulateResults(iout, nhist, nbatch) (omc_dosxyz.c:719-799; omc_matrad.c passes the total history count,
 * omc_dosxyz.c:1282 the per-batch one) evaluated ON THE DEVICE from the resident tallies: batch mean, batch-method
 * relative uncertainty, MeV -> Gy with the voxel mass (iout != 0), dose 0 / uncertainty 0.9999999 where the density
 * is below 0.044 g/cm3 or nothing was scored.  med_densities = geometry.med_densities [nvox] (g/cm3).  dose and
 * unc are indexed like score.accum_endep (element irl = 1 + ix + iy*isize + iz*isize*jsize; element 0 is copied
 * through), so a user code passes score.accum_endep / score.accum_endep2 and calls its outputResults() unchanged.
 * The device tallies are not modified. */
int omc_gpu_accumulate_results(omc_gpu_handle h, int iout, int nhist, int nbatch, const double *med_densities, double *dose,
                               double *unc);
/* outputResults(output_file, iout, nhist, nbatch) (omc_dosxyz.c:801-886) with BOTH halves on the device: accumulateResults()
 * as above, then the text of the .3ddose file -- the reference's fprintf("%e ") per dose value and fprintf("%f ") per
 * uncertainty (:862-877), byte for byte what glibc prints -- is produced by a formatting kernel and streamed through pinned
 * buffers to `path` (the complete file name; the reference builds it from "output folder" + output_file + ".3ddose",
 * :822-833) while the next chunk is being formatted.  At 1 mm voxels (8e7 values per block) the per-value fprintf of the
 * reference is the wall-clock bottleneck of the output phase.  The header lines (dimensions, voxel
 * boundaries) are written by the host with the reference's formats.  The device tallies are not modified. */
int omc_gpu_write_3ddose(omc_gpu_handle h, const char *path, int iout, int nhist, int nbatch, const double *med_densities);
/* The beamlet loop of omc_matrad.c:1389-1493 for beamlets [ib0, ib0+nb) in ONE pass of the wavefront kernels: all their
 * histories run concurrently (beamlet ib0+k owns history ids [first_history + k*nhist, +nhist) and its own fp32 dose
 * grid), then accumulateResults(1, nhist, nbatch) + the relDoseThreshold test + the sparse column assembly
 * (omc_matrad.c:1416-1477) run on the device.  nbatch only enters as the reference's normalisation: the
 * batch-method uncertainty is never exported by omc_matrad (Q12), so no batch structure is kept.  jc[nb+1] receives
 * the column starts (jc[0] = 0), *nnz_total their sum; rows (irl-1, ascending) and values of all columns stay on the
 * device until omc_gpu_fetch_columns(ir[nnz_total], val[nnz_total]).  Needs nb * nreg * 4 bytes of device memory.  A pass pays
 * the ramp-up and the tail of its longest particle lineages once whatever its size, but its dose atomics spread over nb grids:
 * measured on B200 (PROSTATE at 3 mm, 12 MB per grid, 320 beamlets x 1e6 histories) 64 beamlets per pass run 7.3-7.9e7
 * histories/s, all 320 in one pass 6.5e7, one per pass 1.2e7 -- hence OMC_BEAMLETS_PER_PASS. */
int omc_gpu_run_beamlets(omc_gpu_handle h, long long first_history, int nhist, int nbatch, int ib0, int nb, double rel_threshold,
                         const double *med_densities, long long *jc, long long *nnz_total);
/* default pass size of the drivers, and the HBM budget that caps it for large grids (1 mm voxels: 0.33 GB per beamlet) */
#define OMC_BEAMLETS_PER_PASS 64
#define OMC_BEAMLET_GRID_BUDGET (64.0 * 1073741824.0)
int omc_gpu_fetch_columns(omc_gpu_handle h, long long *ir, double *val);
/* score.endep of the running batch, [nreg] fp64 (before accum_batch) */
int omc_gpu_get_batch_grid(omc_gpu_handle h, double *endep);
/* memset of the three grids (initScore(), omc_dosxyz.c:647-665; omc_matrad.c:1482 zeroes accum only: which = 1) */
int omc_gpu_reset_tallies(omc_gpu_handle h, int which /* 0 = all, 1 = accum_endep only */);
/* device addresses + element counts, for an NCCL reduce driven by the host plumbing (multi-GPU) */
int omc_gpu_device_ptrs(omc_gpu_handle h, void **endep, void **accum_endep, void **accum_endep2, long long *nreg);
/* cudaStream_t the context launches on (so callers can time with events on the right stream) */
void *omc_gpu_stream(omc_gpu_handle h);
int omc_gpu_get_counters(omc_gpu_handle h, omc_gpu_counters *c);
/* debug: copy per-history records of the last run_histories (needs option record_histories=1) */
int omc_gpu_get_history_records(omc_gpu_handle h, omc_history_record *out, long long n);

/* ---- multi-GPU: histories shard, the only exchange is the sum of a completed batch grid ---------------- */
/*
 * The reference parallelises the history loop of ONE batch over an OpenMP team that shares score.endep
 * (omc_dosxyz.c:1184-1191, :1252-1259, :690-691).  Here the team is a set of GPUs joined by an NCCL communicator that lives
 * INSIDE the library (NCCL is resolved with dlopen at the first call; single-GPU use needs none):
 *
 *   one process per GPU (MPI-style launchers, torchrun):   rank 0 calls omc_gpu_comm_unique_id(id) and hands the 128 bytes
 *       to the other ranks by whatever means the launcher offers; every rank calls omc_gpu_comm_init(h, rank, world, id).
 *   one process, several GPUs (a plain C user code):        omc_gpu_multi_* below does exactly that with one host thread
 *       per device.
 *
 * With a communicator, omc_gpu_run_batch() / omc_gpu_start_batch() take the batch's WHOLE id range on every rank and
 * transport this rank's contiguous slice of it (history id -> RNG stream: the result does not depend on the rank count
 * beyond fp64 summation order); omc_gpu_accum_batch() -- explicit, or owed by omc_gpu_run_batch() -- sums the completed batch
 * grid over the ranks (ncclAllReduce, fp64) BEFORE accumEndep() squares it, so accum / accum2 on every rank are those of a
 * single-GPU run of the same batches.  Both run on a side stream behind an event: the transport stream starts the next
 * batch at once and waits for them only when that dose grid is needed again, one batch later.  Every rank must issue the
 * same sequence of batch calls.  score.ensrc and the work counters stay per rank (omc_gpu_comm_sum adds host values up).
 */
/* the sharding rule itself (pure host arithmetic, no device needed): rank r of `world` owns [*lo, *lo + *count) of the batch
 * [first, first + nhist); contiguous slices in rank order, the first nhist % world ranks get one history more */
int omc_gpu_shard_range(long long first, long long nhist, int rank, int world, long long *lo, long long *count);
int omc_gpu_comm_unique_id(char *id128 /* [128] out */);
int omc_gpu_comm_init(omc_gpu_handle h, int rank, int world, const char *id128);   /* world == 1: drops the communicator */
int omc_gpu_comm_rank(omc_gpu_handle h);
int omc_gpu_comm_size(omc_gpu_handle h);
int omc_gpu_comm_sum(omc_gpu_handle h, double *values, int n);   /* in-place sum of n host doubles over the ranks (collective) */
/* omc_matrad with one process per GPU: the columns each rank computed (omc_gpu_run_beamlets / omc_gpu_fetch_columns) -> the
 * complete sparse matrix on every rank, in beamlet order as omc_matrad.c:1416-1477 appends them.  jc[nb_total+1] = global
 * column starts (per-beamlet counts summed with omc_gpu_comm_sum), mine[b] != 0 marks this rank's beamlets, ir_mine / val_mine
 * hold their rows and values concatenated in beamlet order.  Collective. */
int omc_gpu_comm_gather_columns(omc_gpu_handle h, int nb_total, const long long *jc, const unsigned char *mine, const long long *ir_mine,
                                const double *val_mine, long long *ir_out /* [jc[nb_total]] */, double *val_out);

/* One handle over several GPUs of this node for a single-process user code: the batch loop of omc_dosxyz.c:1237-1263 with
 * `omc_gpu_multi_run_batch(m, ibatch*nperbatch, nperbatch, -1)` in place of the OpenMP loop + accumEndep(), and the beamlet
 * loop of omc_matrad.c:1389-1493 with omc_gpu_multi_run_beamlets().  Setters broadcast the problem to every device; results
 * come from device 0 (every device holds the same statistics), ensrc and the counters are summed. */
typedef struct omc_gpu_multi_ctx *omc_gpu_multi;
int  omc_gpu_multi_create(omc_gpu_multi *m, int ndev /* <= 0: all visible */, const int *device_ids /* NULL: 0..ndev-1 */);
void omc_gpu_multi_destroy(omc_gpu_multi m);
int  omc_gpu_multi_size(omc_gpu_multi m);
omc_gpu_handle omc_gpu_multi_device(omc_gpu_multi m, int i);
const char *omc_gpu_multi_last_error(omc_gpu_multi m);
int omc_gpu_multi_set_media(omc_gpu_multi m, const omc_media_tables *t);
int omc_gpu_multi_set_geometry(omc_gpu_multi m, const omc_geometry *g);
int omc_gpu_multi_set_source_dosxyz(omc_gpu_multi m, const omc_source_dosxyz *s);
int omc_gpu_multi_set_source_matrad(omc_gpu_multi m, const omc_source_matrad *s);
int omc_gpu_multi_set_vrt(omc_gpu_multi m, int nsplit);
int omc_gpu_multi_set_seed(omc_gpu_multi m, int ixx, int jxx);
int omc_gpu_multi_set_option(omc_gpu_multi m, const char *key, long long value);
int omc_gpu_multi_reset_tallies(omc_gpu_multi m, int which);
int omc_gpu_multi_run_batch(omc_gpu_multi m, long long first_history, long long nhist, int ibeamlet);
int omc_gpu_multi_synchronize(omc_gpu_multi m);
int omc_gpu_multi_get_tallies(omc_gpu_multi m, double *accum_endep, double *accum_endep2, double *ensrc);
int omc_gpu_multi_accumulate_results(omc_gpu_multi m, int iout, int nhist, int nbatch, const double *med_densities, double *dose,
                                     double *unc);
int omc_gpu_multi_write_3ddose(omc_gpu_multi m, const char *path, int iout, int nhist, int nbatch, const double *med_densities);
int omc_gpu_multi_get_counters(omc_gpu_multi m, omc_gpu_counters *c);
/* omc_gpu_run_beamlets() over the devices: passes of `per_pass` consecutive beamlets (<= 0: OMC_BEAMLETS_PER_PASS), pass p on
 * device p % ndev, the column slices gathered in beamlet order as the reference's loop appends them (omc_matrad.c:1416-1477);
 * beamlet ib0 + k owns history ids [first_history + k*nhist, +nhist) whatever the device count. */
int omc_gpu_multi_run_beamlets(omc_gpu_multi m, long long first_history, int nhist, int nbatch, int ib0, int nb, int per_pass,
                               double rel_threshold, const double *med_densities, long long *jc /* [nb+1] */, long long *nnz_total);
int omc_gpu_multi_fetch_columns(omc_gpu_multi m, long long *ir, double *val);

/* sizeof() of the structs above as this library was compiled (0 media, 1 geometry, 2 source_dosxyz,
 * 3 source_matrad, 4 history_record, 5 counters): lets a foreign-language binding check its layout */
int omc_gpu_abi_sizeof(int what);

/* ---- unit-test hooks: run single device functions on explicit inputs --------------------- */
/*
 * howfar()/hownear() (omc_dosxyz.c:187-334) for n particles: in x,y,z,u,v,w,ustep_in,ir -> out
 * idisc, irnew, ustep, tperp.
 */
int omc_gpu_test_geometry(omc_gpu_handle h, int n, const double *xyzuvw /* [6*n] */, const int *ir,
                          const double *ustep_in, int *idisc, int *irnew, double *ustep_out,
                          double *tperp);
/*
 * shower() (src/ompmc.c:5436-5447) for n explicit top-of-stack particles -- charge, TOTAL energy, position,
 * direction, region, weight -- particle i using the Philox stream of history first_history + i, through the
 * lock-step kernel.  Deposits go to the batch grid; one record per particle.  Sampler-level parity tests.
 */
int omc_gpu_test_particles(omc_gpu_handle h, int n, const int *iq, const double *e, const double *xyzuvw, const int *ir,
                           const double *wt, long long first_history, omc_history_record *records);
/*
 * The samplers of the PRODUCTION (wavefront) kernels on explicit inputs, one thread per record, record i drawing from the
 * Philox stream of history first_history + i.  `in` and `out` hold n records of 8 doubles.  These are the mixed-precision /
 * block-draw re-implementations (csrc/omc_physics_f32.cuh, omc_wavefront.cu) of the reference functions cited per sampler;
 * tests/test_gpu_production_samplers.py compares them with the reference's own functions on the same inputs.
 *   OMC_SAMPLER_DRANGE   computeDrange()  src/ompmc.c:3979-4014   in {imed, iq, ekei, ekef (same PWL bin)}   out {path length}
 *   OMC_SAMPLER_ELOSS    computeEloss()   :4016-4108 + the range of electron() :4905-4918
 *                                         in {imed, iq, rhof, eke, tustep / range}                            out {range, de}
 *   OMC_SAMPLER_MSDIST   msdist() + mscat() + spinRejection()  :3787-3976, :3606-3785, :3097-3168
 *                                         in {imed, iq, rhof, eke, tustep / range, u, v, w}   out {ustep, dx, dy, dz, uf, vf, wf, de}
 *   OMC_SAMPLER_SSCAT    sscat() + selectAzimuthalAngle()  :3170-3199, :101-122
 *                                         in {imed, qel, chia2, elke, beta2}                 out {cos, sin, cos(phi), sin(phi)}
 *   OMC_SAMPLER_COMPTON  compton()        :1670-1783   in {e, u, v, w}        out {e_photon, u, v, w, e_electron (total), u, v, w}
 *   OMC_SAMPLER_MOLLER   moller()         :4359-4435   in {imed, e, u, v, w}  out {e1, u1, v1, w1, e2 (0: none created), u2, v2, w2}
 *   OMC_SAMPLER_WOODCOCK photon free flight :1951-2019 (Woodcock tracking instead of the voxel march)
 *                                         in {e, x, y, z, u, v, w}  out {1 = interaction site / 0 = left the phantom, x, y, z, ir, tentative collisions}
 *   OMC_SAMPLER_ESTEP    step-size phase of electron() :4694-4967
 *                                         in {iq, e (total), x, y, z, ir}  out {class (1 CH, 2 BCA, 0 below cut-off), tustep, tperp, range,
 *                                                                               total_tstep, demfp, blccl, ssmfp}
 */
enum { OMC_SAMPLER_DRANGE = 0, OMC_SAMPLER_ELOSS = 1, OMC_SAMPLER_MSDIST = 2, OMC_SAMPLER_SSCAT = 3, OMC_SAMPLER_COMPTON = 4,
       OMC_SAMPLER_MOLLER = 5, OMC_SAMPLER_WOODCOCK = 6, OMC_SAMPLER_ESTEP = 7 };
int omc_gpu_test_samplers(omc_gpu_handle h, int which, int n, const double *in /* [8*n] */, long long first_history,
                          double *out /* [8*n] */);
/* n raw Philox draws of history `hist` as the transport sees them (double in [0,1)) */
int omc_gpu_test_rng(omc_gpu_handle h, long long hist, int n, double *out);
/* n host doubles through the formatting kernel of omc_gpu_write_3ddose into `path`: one block of a .3ddose file, i.e. what
 * `for (i < n) fprintf(fp, mode ? "%f " : "%e ", values[i]); fprintf(fp, "\n");` writes (omc_dosxyz.c:859-877), byte for byte. */
int omc_gpu_test_format(omc_gpu_handle h, int mode, long long n, const double *values, const char *path);

#ifdef __cplusplus
}
#endif
#endif /* OMPMC_B200_H */

/*
 * omc_dosxyz_dropin.c -- the reference's OWN user code (ucodes/omc_dosxyz/omc_dosxyz.c) with its batch loop handed to
 * libompmc_b200.so: the patch of INTEGRATION.md as a compilable program.
 *
 * Nothing of the reference is copied here.  Its user code is #included from where it lies (OMC_REF_DOSXYZ_C, given on the
 * compiler command line by oracle/Makefile) with main() renamed, and linked with the reference's src/ompmc.c,
 * src/omc_utilities.c and src/omc_random.c exactly as ucodes/omc_dosxyz/Makefile:18-34 does.  So parseInputFile(),
 * initPhantom(), initMediaData(), initSource(), initRegions(), initVrt(), initScore(), accumulateResults() and
 * outputResults() below ARE the reference's functions filling and reading the reference's globals; the only thing replaced
 * is what SURVEY.md 8b names as the boundary:
 *
 *     omc_dosxyz.c:1252-1262   #pragma omp parallel for ... { initHistory(); shower(); }  accumEndep();
 *  -> omc_gpu_multi_run_batch(gpu, first_history, nperbatch, -1)
 *
 * The main() below restates the control flow of omc_dosxyz.c:1072-1307 (options -i/-o, init order, batch plan by atoi,
 * printed lines) around that call.  There is no CPU fallback: without a CUDA device the program stops with the library's
 * message and EXIT_FAILURE, the reference's error behaviour.  Extra option (not in the reference): -k 0|1 selects the
 * lock-step parity kernel or the wavefront kernels (default: wavefront, lock-step when nsplit > 255); -g N (or OMC_GPUS=N) the number of GPUs.
 */
#define main omc_dosxyz_reference_main
#include OMC_REF_DOSXYZ_C
#undef main

#include "ompmc_b200.h"

static omc_gpu_multi gpu;                     /* one handle over `ngpu` devices (-g N / OMC_GPUS; default 1) */

static void die(const char *what) {           /* the reference's error behaviour: printf + exit */
    printf("%s: %s\n", what, gpu ? omc_gpu_multi_last_error(gpu) : "no CUDA device / library not usable");
    exit(EXIT_FAILURE);
}

/* hand the reference's initialised globals (borrowed pointers) to the device; INTEGRATION.md, gpu_upload() */
static void gpu_upload(int kernel, int ngpu) {
    if (omc_gpu_multi_create(&gpu, ngpu, NULL)) { gpu = NULL; die("omc_gpu_multi_create"); }

    omc_media_tables t;
    memset(&t, 0, sizeof t);
    t.nmed = media.nmed;
    t.ge0 = photon_data.ge0;     t.ge1 = photon_data.ge1;
    t.gmfp0 = photon_data.gmfp0; t.gmfp1 = photon_data.gmfp1;
    t.gbr10 = photon_data.gbr10; t.gbr11 = photon_data.gbr11;
    t.gbr20 = photon_data.gbr20; t.gbr21 = photon_data.gbr21;
    t.cohe0 = photon_data.cohe0; t.cohe1 = photon_data.cohe1;
    t.ray_xgrid = rayleigh_data.xgrid;     t.ray_fcum = rayleigh_data.fcum;
    t.ray_b_array = rayleigh_data.b_array; t.ray_c_array = rayleigh_data.c_array;
    t.ray_i_array = rayleigh_data.i_array;
    t.ray_pmax0 = rayleigh_data.pmax0;     t.ray_pmax1 = rayleigh_data.pmax1;
    t.dl1 = pair_data.dl1; t.dl2 = pair_data.dl2; t.dl3 = pair_data.dl3;
    t.dl4 = pair_data.dl4; t.dl5 = pair_data.dl5; t.dl6 = pair_data.dl6;
    t.bpar0 = pair_data.bpar0; t.bpar1 = pair_data.bpar1; t.delcm = pair_data.delcm; t.zbrang = pair_data.zbrang;
#define E(f) t.f = electron_data.f;
    E(esig0) E(esig1) E(psig0) E(psig1) E(ededx0) E(ededx1) E(pdedx0) E(pdedx1) E(ebr10) E(ebr11)
    E(pbr10) E(pbr11) E(pbr20) E(pbr21) E(tmxs0) E(tmxs1) E(blcce0) E(blcce1) E(etae_ms0) E(etae_ms1)
    E(etap_ms0) E(etap_ms1) E(q1ce_ms0) E(q1ce_ms1) E(q1cp_ms0) E(q1cp_ms1) E(q2ce_ms0) E(q2ce_ms1)
    E(q2cp_ms0) E(q2cp_ms1) E(range_ep) E(e_array) E(eke0) E(eke1) E(sig_ismonotone) E(esig_e) E(psig_e)
    E(xcc) E(blcc)
#undef E
    t.b2spin_min = spin_data.b2spin_min; t.dbeta2i = spin_data.dbeta2i; t.espml = spin_data.espml;
    t.dleneri = spin_data.dleneri; t.dqq1i = spin_data.dqq1i; t.spin_rej = spin_data.spin_rej;
    t.ums = mscat_data.ums_array; t.fms = mscat_data.fms_array; t.wms = mscat_data.wms_array;
    t.ims = mscat_data.ims_array; t.dllambi = mscat_data.dllambi; t.dqmsi = mscat_data.dqmsi;
    t.pegs_ap = pegs_data.ap; t.pegs_ae = pegs_data.ae; t.pegs_te = pegs_data.te;
    t.pegs_thmoll = pegs_data.thmoll; t.pegs_rho = pegs_data.rho; t.pegs_meke = pegs_data.meke;
    if (omc_gpu_multi_set_media(gpu, &t)) die("omc_gpu_multi_set_media");

    omc_geometry g = { geometry.isize, geometry.jsize, geometry.ksize,
                       geometry.xbounds, geometry.ybounds, geometry.zbounds,
                       region.med, region.rhof, region.pcut, region.ecut };
    if (omc_gpu_multi_set_geometry(gpu, &g)) die("omc_gpu_multi_set_geometry");

    omc_source_dosxyz s = { source.spectrum, source.charge, source.energy, source.deltak,
                            source.cdfinv1, source.cdfinv2, source.ssd,
                            source.xinl, source.xinu, source.yinl, source.yinu, source.xsize, source.ysize,
                            source.ixinl, source.ixinu, source.iyinl, source.iyinu };
    if (omc_gpu_multi_set_source_dosxyz(gpu, &s)) die("omc_gpu_multi_set_source_dosxyz");
    if (omc_gpu_multi_set_vrt(gpu, vrt.nsplit)) die("omc_gpu_multi_set_vrt");

    char buffer[BUFFER_SIZE];
    int ixx = 1802, jxx = 9373;               /* defaults of initRandom(), src/omc_random.c:64-82 */
    if (getInputValue(buffer, "rng seeds") == 1) sscanf(buffer, "%d %d", &ixx, &jxx);
    if (omc_gpu_multi_set_seed(gpu, ixx, jxx)) die("omc_gpu_multi_set_seed");
    if (kernel < 0) kernel = vrt.nsplit > 255 ? OMC_KERNEL_LOCKSTEP : OMC_KERNEL_WAVEFRONT;
    if (omc_gpu_multi_set_option(gpu, "kernel", kernel)) die("omc_gpu_multi_set_option");
}

int main(int argc, char **argv) {
    double tbegin = omc_get_time();
    char *input_file = NULL, *output_file = NULL;
    int kernel = -1, ngpu = getenv("OMC_GPUS") ? atoi(getenv("OMC_GPUS")) : 1;
    for (int i = 1; i < argc; i++) {
        if ((!strcmp(argv[i], "-i") || !strcmp(argv[i], "--input")) && i + 1 < argc) input_file = argv[++i];
        else if (!strcmp(argv[i], "-g") && i + 1 < argc) ngpu = atoi(argv[++i]);
        else if ((!strcmp(argv[i], "-o") || !strcmp(argv[i], "--output")) && i + 1 < argc) output_file = argv[++i];
        else if (!strcmp(argv[i], "-k") && i + 1 < argc) kernel = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--verbose")) verbose_flag = 1;
        else if (!strcmp(argv[i], "--brief")) verbose_flag = 0;
        else { printf("usage: %s -i <input stem> -o <output stem> [-k 0|1] [-g ngpu] [--verbose]\n", argv[0]); exit(EXIT_FAILURE); }
    }
    if (!input_file || !output_file) { printf("usage: %s -i <input stem> -o <output stem>\n", argv[0]); exit(EXIT_FAILURE); }

    /* the reference's own initialisation, in the reference's order (omc_dosxyz.c:1157-1182) */
    parseInputFile(input_file);
    initPhantom();
    initMediaData();
    initSource();
    initRegions();
    initVrt();
    initScore();
    if (verbose_flag) { listRayleigh(); listPair(); listPhoton(); listElectron(); listMscat(); listSpin(); }

    /* batch plan with the reference's atoi/int bookkeeping (omc_dosxyz.c:1207-1225) */
    char buffer[BUFFER_SIZE];
    if (getInputValue(buffer, "ncase") != 1) { printf("Can not find 'ncase' key on input file.\n"); exit(EXIT_FAILURE); }
    int nhist = atoi(buffer);
    if (getInputValue(buffer, "nbatch") != 1) { printf("Can not find 'nbatch' key on input file.\n"); exit(EXIT_FAILURE); }
    int nbatch = atoi(buffer);
    if (nhist / nbatch == 0) nhist = nbatch;
    int nperbatch = nhist / nbatch;
    nhist = nperbatch * nbatch;
    printf("Total number of particle histories: %d\n", nhist);
    printf("Number of statistical batches: %d\n", nbatch);
    printf("Histories per batch: %d\n", nperbatch);

    gpu_upload(kernel, ngpu);
    printf("GPUs: %d\n", omc_gpu_multi_size(gpu));
    printf("Execution time up to this point : %8.2f seconds\n", omc_get_time() - tbegin);

    double tloop = omc_get_time();
    for (int ibatch = 0; ibatch < nbatch; ibatch++) {
        printf("%-10d\t%-15.2f\n", ibatch, omc_get_time() - tbegin);
        /* == { initHistory(); shower(); } x nperbatch + accumEndep()   (omc_dosxyz.c:1252-1262) */
        if (omc_gpu_multi_run_batch(gpu, (long long)ibatch * nperbatch, nperbatch, -1)) die("omc_gpu_multi_run_batch");
    }
    /* the tallies accumEndep() would have left in the reference's struct Score (omc_dosxyz.c:696-717) */
    if (omc_gpu_multi_get_tallies(gpu, score.accum_endep, score.accum_endep2, &score.ensrc)) die("omc_gpu_multi_get_tallies");
    printf("Simulation finished\n");
    printf("Execution time up to this point : %8.2f seconds\n", omc_get_time() - tbegin);
    printf("Batch loop: %.6f seconds, %.4e histories/s\n", omc_get_time() - tloop, nhist / (omc_get_time() - tloop));

    if (verbose_flag) {
        int gridsize = geometry.isize * geometry.jsize * geometry.ksize;
        double etot = 0.0;
        for (int irl = 1; irl < gridsize + 1; irl++) etot += score.accum_endep[irl];
        printf("Fraction of incident energy deposited in the phantom: %5.4f\n", etot / score.ensrc);
    }

    /* unchanged reference code from here: accumulateResults() + the .3ddose writer (omc_dosxyz.c:719-886, :1281-1282) */
    outputResults(output_file, 1, nperbatch, nbatch);

    omc_gpu_multi_destroy(gpu);
    cleanPhantom(); cleanPhoton(); cleanRayleigh(); cleanPair(); cleanElectron(); cleanMscat(); cleanSpin();
    cleanRegions(); cleanScore(); cleanSource();
    printf("Total execution time : %8.5f seconds\n", omc_get_time() - tbegin);
    return EXIT_SUCCESS;
}

/*
 * omc_matrad_dropin.c -- the reference's OWN matRad user code (ucodes/omc_matrad/omc_matrad.c) with its beamlet loop handed to
 * libompmc_b200.so: the mexFunction() a maintainer builds with `mex omc_matrad_dropin.c src/ompmc.c src/omc_utilities.c
 * src/omc_random.c -lompmc_b200` (the matRad part of INTEGRATION.md as a compilable file).
 *
 * Nothing of the reference is copied here.  Its user code is #included from where it lies (OMC_REF_MATRAD_C, given on the
 * compiler command line) with mexFunction() renamed, so parseInput(), initPhantom(), initMediaData(), initSource(),
 * initRegions(), initVrt() below ARE the reference's functions filling the reference's globals from the five MATLAB inputs
 * (cubeRho, cubeMatIx, ompMCgeo, ompMCsource, ompMCoptions).  What is replaced is what SURVEY.md 8b / 8f-1 name:
 *
 *     omc_matrad.c:1389-1493   for ibeamlet { for ibatch { omp for { initHistory(ibeamlet); shower(); } accumEndep(); }
 *                                             accumulateResults(); threshold; append the column to the sparse matrix }
 *  -> omc_gpu_run_beamlets() per group of OMC_BEAMLETS_PER_PASS beamlets (all of them in ONE pass of the wavefront kernels,
 *     statistics + threshold + column assembly on the device) + omc_gpu_fetch_columns() straight into the mxArray.
 *
 * The mexFunction() below restates the control flow of omc_matrad.c:1258-1543 around those calls (same printed lines, same
 * sparse output: rows irl-1 ascending per column, values = accumulateResults(1, nhist, nbatch)).  There is no CPU fallback:
 * without a CUDA device it stops through mexErrMsgIdAndTxt() with the library's message, the reference's error behaviour.
 *
 * Without MATLAB (this image) the file is compiled against the stand-in oracle/mexshim/mex.h; with -DOMC_DROPIN_MAIN it also
 * gets a main() that builds the five inputs from a problem blob (ompmc_b200/problem.py: phantom cubes, bounds, beamlets) and
 * the paths of the reference's data files, calls mexFunction() and writes the matrix as the CSC file of omc_matrad_b200
 * (--dump-problem: stop after the reference's initialisation and dump what would go to the device -- no GPU needed; the CPU
 * test-suite compares that with the problem the Python host builds).
 */
#define mexFunction omc_matrad_reference_mexFunction
#include OMC_REF_MATRAD_C
#undef mexFunction
#undef exit                      /* the reference's macro (omc_matrad.c:41) */

#include "ompmc_b200.h"

static omc_gpu_handle gpu;
static int omc_dropin_dump_only = 0;           /* (main() below, --dump-problem) */
static const char *omc_dropin_dump_stem = NULL;

static void gpu_die(const char *what) {        /* the reference's error behaviour in a mex: mexErrMsgIdAndTxt */
    static char msg[600];
    snprintf(msg, sizeof msg, "%s: %s", what, gpu ? omc_gpu_last_error(gpu) : "no CUDA device / library not usable");
    mexErrMsgIdAndTxt("matRad:matRad_ompInterface:invalid", msg);
}

static void fill_media(omc_media_tables *t) {  /* borrowed pointers into the reference's globals */
    memset(t, 0, sizeof *t);
    t->nmed = media.nmed;
    t->ge0 = photon_data.ge0;     t->ge1 = photon_data.ge1;
    t->gmfp0 = photon_data.gmfp0; t->gmfp1 = photon_data.gmfp1;
    t->gbr10 = photon_data.gbr10; t->gbr11 = photon_data.gbr11;
    t->gbr20 = photon_data.gbr20; t->gbr21 = photon_data.gbr21;
    t->cohe0 = photon_data.cohe0; t->cohe1 = photon_data.cohe1;
    t->ray_xgrid = rayleigh_data.xgrid;     t->ray_fcum = rayleigh_data.fcum;
    t->ray_b_array = rayleigh_data.b_array; t->ray_c_array = rayleigh_data.c_array;
    t->ray_i_array = rayleigh_data.i_array;
    t->ray_pmax0 = rayleigh_data.pmax0;     t->ray_pmax1 = rayleigh_data.pmax1;
    t->dl1 = pair_data.dl1; t->dl2 = pair_data.dl2; t->dl3 = pair_data.dl3;
    t->dl4 = pair_data.dl4; t->dl5 = pair_data.dl5; t->dl6 = pair_data.dl6;
    t->bpar0 = pair_data.bpar0; t->bpar1 = pair_data.bpar1; t->delcm = pair_data.delcm; t->zbrang = pair_data.zbrang;
#define E(f) t->f = electron_data.f;
    E(esig0) E(esig1) E(psig0) E(psig1) E(ededx0) E(ededx1) E(pdedx0) E(pdedx1) E(ebr10) E(ebr11)
    E(pbr10) E(pbr11) E(pbr20) E(pbr21) E(tmxs0) E(tmxs1) E(blcce0) E(blcce1) E(etae_ms0) E(etae_ms1)
    E(etap_ms0) E(etap_ms1) E(q1ce_ms0) E(q1ce_ms1) E(q1cp_ms0) E(q1cp_ms1) E(q2ce_ms0) E(q2ce_ms1)
    E(q2cp_ms0) E(q2cp_ms1) E(range_ep) E(e_array) E(eke0) E(eke1) E(sig_ismonotone) E(esig_e) E(psig_e)
    E(xcc) E(blcc)
#undef E
    t->b2spin_min = spin_data.b2spin_min; t->dbeta2i = spin_data.dbeta2i; t->espml = spin_data.espml;
    t->dleneri = spin_data.dleneri; t->dqq1i = spin_data.dqq1i; t->spin_rej = spin_data.spin_rej;
    t->ums = mscat_data.ums_array; t->fms = mscat_data.fms_array; t->wms = mscat_data.wms_array;
    t->ims = mscat_data.ims_array; t->dllambi = mscat_data.dllambi; t->dqmsi = mscat_data.dqmsi;
    t->pegs_ap = pegs_data.ap; t->pegs_ae = pegs_data.ae; t->pegs_te = pegs_data.te;
    t->pegs_thmoll = pegs_data.thmoll; t->pegs_rho = pegs_data.rho; t->pegs_meke = pegs_data.meke;
}

static void fill_geometry(omc_geometry *g) {
    g->isize = geometry.isize; g->jsize = geometry.jsize; g->ksize = geometry.ksize;
    g->xbounds = geometry.xbounds; g->ybounds = geometry.ybounds; g->zbounds = geometry.zbounds;
    g->med = region.med; g->rhof = region.rhof; g->pcut = region.pcut; g->ecut = region.ecut;
}

static void fill_source(omc_source_matrad *s) {            /* struct Source of omc_matrad.c:507-541 */
    memset(s, 0, sizeof *s);
    s->spectrum = source.spectrum; s->charge = source.charge; s->energy = source.energy; s->deltak = source.deltak;
    s->cdfinv1 = source.cdfinv1; s->cdfinv2 = source.cdfinv2;
    s->nbixels = source.nbeamlets; s->ibeam = source.ibeam;
    s->nbeams = 0;                                          /* the reference keeps no beam count: highest beam index + 1 */
    for (int i = 0; i < source.nbeamlets; i++)
        if (source.ibeam[i] + 1 > s->nbeams) s->nbeams = source.ibeam[i] + 1;
    s->xsource = source.xsource; s->ysource = source.ysource; s->zsource = source.zsource;
    s->xcorner = source.xcorner; s->ycorner = source.ycorner; s->zcorner = source.zcorner;
    s->xside1 = source.xside1; s->yside1 = source.yside1; s->zside1 = source.zside1;
    s->xside2 = source.xside2; s->yside2 = source.yside2; s->zside2 = source.zside2;
}

#ifdef OMC_DROPIN_MAIN
static int dump_problem(const char *stem);
#endif

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) {
    double tbegin = omc_get_time();
    if (nrhs != 5)
        mexErrMsgIdAndTxt("matRad:matRad_ompInterface:invalidNumInputs", "Two or three input arguments required.");
    if (nlhs > 1)
        mexErrMsgIdAndTxt("matRad:matRad_ompInterface:invalidNumOutputs", "Too many output arguments.");

    /* the reference's own initialisation, in the reference's order (omc_matrad.c:1279-1303) */
    parseInput(nrhs, prhs);
    initPhantom();
    initMediaData();
    initSource();
    initRegions();
    initVrt();
#ifdef OMC_DROPIN_MAIN
    if (omc_dropin_dump_only) { if (dump_problem(omc_dropin_dump_stem)) gpu_die("dump"); return; }
#endif

    /* history bookkeeping with the reference's atoi/int arithmetic (omc_matrad.c:1318-1350) */
    char buffer[BUFFER_SIZE];
    if (getInputValue(buffer, "ncase") != 1) mexErrMsgIdAndTxt("matRad:matRad_ompInterface:invalid", "Can not find 'ncase' key on input file.");
    int nhist = atoi(buffer);
    if (getInputValue(buffer, "nbatch") != 1) mexErrMsgIdAndTxt("matRad:matRad_ompInterface:invalid", "Can not find 'nbatch' key on input file.");
    int nbatch = atoi(buffer);
    if (nbatch < 1) mexErrMsgIdAndTxt("matRad:matRad_ompInterface:invalid", "nbatch must be positive.");
    if (nhist / nbatch == 0) nhist = nbatch;
    const int nperbatch = nhist / nbatch;
    nhist = nperbatch * nbatch;
    mexPrintf("Total number of particle histories: %d\n", nhist);
    mexPrintf("Number of statistical batches: %d\n", nbatch);
    mexPrintf("Histories per batch: %d\n", nperbatch);
    if (getInputValue(buffer, "relative dose threshold") != 1)
        mexErrMsgIdAndTxt("matRad:matRad_ompInterface:invalid", "Can not find 'relative dose threshold' key on input file.");
    const double relDoseThreshold = atof(buffer);
    mexPrintf("Using a relative dose cut-off of %f\n", relDoseThreshold);

    /* hand the initialised globals to the device */
    if (omc_gpu_create(&gpu, 0)) { gpu = NULL; gpu_die("omc_gpu_create"); }
    omc_media_tables t; omc_geometry g; omc_source_matrad s;
    fill_media(&t); fill_geometry(&g); fill_source(&s);
    if (omc_gpu_set_media(gpu, &t)) gpu_die("omc_gpu_set_media");
    if (omc_gpu_set_geometry(gpu, &g)) gpu_die("omc_gpu_set_geometry");
    if (omc_gpu_set_source_matrad(gpu, &s)) gpu_die("omc_gpu_set_source_matrad");
    if (omc_gpu_set_vrt(gpu, vrt.nsplit)) gpu_die("omc_gpu_set_vrt");
    int ixx = 1802, jxx = 9373;                             /* defaults of initRandom(), src/omc_random.c:64-82 */
    if (getInputValue(buffer, "rng seeds") == 1) sscanf(buffer, "%d %d", &ixx, &jxx);
    if (omc_gpu_set_seed(gpu, ixx, jxx)) gpu_die("omc_gpu_set_seed");
    if (omc_gpu_set_option(gpu, "kernel", OMC_KERNEL_WAVEFRONT)) gpu_die("omc_gpu_set_option");

    /* output matrix as the reference creates and grows it (omc_matrad.c:1368-1380, :1434-1462) */
    const mwSize nCubeElements = (mwSize)geometry.isize * geometry.jsize * geometry.ksize;
    const int nbeamlets = source.nbeamlets;
    mwSize nzmax = (mwSize)ceil((double)nCubeElements * (double)nbeamlets * 0.01);
    plhs[0] = mxCreateSparse(nCubeElements, (mwSize)nbeamlets, nzmax, mxREAL);
    double *sr = mxGetPr(plhs[0]);
    mwIndex *irs = mxGetIr(plhs[0]);
    mwIndex *jcs = mxGetJc(plhs[0]);
    mwIndex linIx = 0;
    jcs[0] = 0;
    mexPrintf("Execution time up to this point : %8.2f seconds\n", omc_get_time() - tbegin);

    int group = (int)(OMC_BEAMLET_GRID_BUDGET / ((double)(nCubeElements + 1) * 4.0));
    if (group > OMC_BEAMLETS_PER_PASS) group = OMC_BEAMLETS_PER_PASS;
    if (group < 1) group = 1;
    long long *jc = malloc(((size_t)group + 1) * sizeof(long long));
    for (int ib0 = 0; ib0 < nbeamlets; ib0 += group) {
        const int nb = nbeamlets - ib0 < group ? nbeamlets - ib0 : group;
        long long nnz = 0;
        /* == omc_matrad.c:1389-1432 for beamlets [ib0, ib0+nb): beamlet b owns the history ids [b*nhist, (b+1)*nhist) */
        if (omc_gpu_run_beamlets(gpu, (long long)ib0 * nhist, nhist, nbatch, ib0, nb, relDoseThreshold, geometry.med_densities, jc, &nnz))
            gpu_die("omc_gpu_run_beamlets");
        if (linIx + (mwIndex)nnz > nzmax) {                 /* grow the sparse matrix, :1434-1462 */
            nzmax = linIx + (mwIndex)nnz + (mwIndex)ceil((double)nCubeElements * (double)nbeamlets * 0.01);
            mxSetNzmax(plhs[0], nzmax);
            mxSetPr(plhs[0], (double *)mxRealloc(sr, nzmax * sizeof(double)));
            mxSetIr(plhs[0], (mwIndex *)mxRealloc(irs, nzmax * sizeof(mwIndex)));
            sr = mxGetPr(plhs[0]);
            irs = mxGetIr(plhs[0]);
        }
        /* rows (irl - 1, ascending) and values of the nb columns, straight into the matrix; mwIndex is 64-bit */
        if (omc_gpu_fetch_columns(gpu, (long long *)(irs + linIx), sr + linIx)) gpu_die("omc_gpu_fetch_columns");
        for (int k = 0; k < nb; k++) jcs[ib0 + k + 1] = linIx + (mwIndex)jc[k + 1];
        linIx += (mwIndex)nnz;
    }
    free(jc);
    mexPrintf("Sparse MC Dij has %d (%f percent) elements!\n", (int)linIx, (double)linIx / ((double)nCubeElements * (double)nbeamlets));
    /* truncate to the exact size, :1500-1506 */
    mxSetNzmax(plhs[0], linIx);
    mxSetPr(plhs[0], (double *)mxRealloc(sr, (linIx ? linIx : 1) * sizeof(double)));
    mxSetIr(plhs[0], (mwIndex *)mxRealloc(irs, (linIx ? linIx : 1) * sizeof(mwIndex)));
    mexPrintf("Simulation finished\n");
    mexPrintf("Execution time up to this point : %8.2f seconds\n", omc_get_time() - tbegin);

    omc_gpu_destroy(gpu);
    gpu = NULL;
    cleanPhantom(); cleanPhoton(); cleanRayleigh(); cleanPair(); cleanElectron(); cleanMscat(); cleanSpin();
    cleanRegions(); cleanSource();
    mexPrintf("Total execution time : %8.5f seconds\n", omc_get_time() - tbegin);
}

#ifdef OMC_DROPIN_MAIN
/* ---- stand-alone harness (no MATLAB): the five mex inputs from a problem blob -------------------------------------------------- */
#include "omc_host_common.h"

static int dump_problem(const char *stem) {
    omc_media_tables t; omc_geometry g; omc_source_matrad s;
    fill_media(&t); fill_geometry(&g); fill_source(&s);
    omc_source_dosxyz d;                                    /* (host_dump_problem writes the energy part of a source) */
    memset(&d, 0, sizeof d);
    d.spectrum = s.spectrum; d.charge = s.charge; d.energy = s.energy; d.deltak = s.deltak; d.cdfinv1 = s.cdfinv1; d.cdfinv2 = s.cdfinv2;
    if (host_dump_problem(stem, &t, &g, geometry.med_densities, &d, vrt.nsplit)) return 1;
    char path[512];
    snprintf(path, sizeof path, "%s.beamlets", stem);
    FILE *fp = fopen(path, "wb");
    if (!fp) { printf("Unable to open file: %s\n", path); return 1; }
    const uint32_t n = 15;
    const uint64_t nb = (uint64_t)s.nbixels, nm = (uint64_t)s.nbeams;
    fwrite("OMCBLOB1", 1, 8, fp);
    fwrite(&n, 4, 1, fp);
    blob_put(fp, "mr_nbeamlets", 1, &s.nbixels, 1); blob_put(fp, "mr_nbeams", 1, &s.nbeams, 1); blob_put(fp, "mr_ibeam", 1, s.ibeam, nb);
    blob_put(fp, "mr_xsource", 0, s.xsource, nm); blob_put(fp, "mr_ysource", 0, s.ysource, nm); blob_put(fp, "mr_zsource", 0, s.zsource, nm);
    blob_put(fp, "mr_xcorner", 0, s.xcorner, nb); blob_put(fp, "mr_ycorner", 0, s.ycorner, nb); blob_put(fp, "mr_zcorner", 0, s.zcorner, nb);
    blob_put(fp, "mr_xside1", 0, s.xside1, nb); blob_put(fp, "mr_yside1", 0, s.yside1, nb); blob_put(fp, "mr_zside1", 0, s.zside1, nb);
    blob_put(fp, "mr_xside2", 0, s.xside2, nb); blob_put(fp, "mr_yside2", 0, s.yside2, nb); blob_put(fp, "mr_zside2", 0, s.zside2, nb);
    fclose(fp);
    return 0;
}

static mxArray *dbl_from(const double *v, mwSize n) { return mx_new_double_row(n, v); }
static mxArray *dbl_scalar(double v) { return mx_new_double_row(1, &v); }

int main(int argc, char **argv) {
    const char *pfile = NULL, *stem = "dij", *names = NULL, *datadir = NULL, *pegs = NULL, *form = NULL, *spectrum = NULL;
    double nhist = 100000, nbatch = 10, rel = 0.001, ecut = 0.7, pcut = 0.01, mono = 0.0;
    for (int i = 1; i < argc; i++) {
        if (!strcmp(argv[i], "-p") && i + 1 < argc) pfile = argv[++i];
        else if (!strcmp(argv[i], "-o") && i + 1 < argc) stem = argv[++i];
        else if (!strcmp(argv[i], "-m") && i + 1 < argc) names = argv[++i];
        else if (!strcmp(argv[i], "--data") && i + 1 < argc) datadir = argv[++i];
        else if (!strcmp(argv[i], "--pegs") && i + 1 < argc) pegs = argv[++i];
        else if (!strcmp(argv[i], "--pgs4form") && i + 1 < argc) form = argv[++i];
        else if (!strcmp(argv[i], "--spectrum") && i + 1 < argc) spectrum = argv[++i];
        else if (!strcmp(argv[i], "--mono") && i + 1 < argc) mono = atof(argv[++i]);
        else if (!strcmp(argv[i], "-n") && i + 1 < argc) nhist = atof(argv[++i]);
        else if (!strcmp(argv[i], "-b") && i + 1 < argc) nbatch = atof(argv[++i]);
        else if (!strcmp(argv[i], "-t") && i + 1 < argc) rel = atof(argv[++i]);
        else if (!strcmp(argv[i], "--ecut") && i + 1 < argc) ecut = atof(argv[++i]);
        else if (!strcmp(argv[i], "--pcut") && i + 1 < argc) pcut = atof(argv[++i]);
        else if (!strcmp(argv[i], "--dump-problem")) omc_dropin_dump_only = 1;
        else {
            printf("usage: %s -p problem.blob -m MEDIUM1,MEDIUM2,... --data <folder/> --pegs <file> --pgs4form <file> [--spectrum <file> | --mono E]\n"
                   "       -n nHistories -b nBatches -t relDoseThreshold [--ecut E --pcut E] -o out_stem [--dump-problem]\n", argv[0]);
            return 2;
        }
    }
    if (!pfile || !names || !datadir || !pegs || !form) { printf("missing -p / -m / --data / --pegs / --pgs4form\n"); return 2; }
    omc_dropin_dump_stem = stem;
    blob b;
    if (blob_read(&b, pfile) != 0) { printf("Unable to open file: %s\n", pfile); return EXIT_FAILURE; }
    const int isize = I(&b, "isize")[0], jsize = I(&b, "jsize")[0], ksize = I(&b, "ksize")[0];
    const mwSize dims[3] = {(mwSize)isize, (mwSize)jsize, (mwSize)ksize}, nvox = dims[0] * dims[1] * dims[2];

    /* prhs[0] cubeRho (double, 3-D), prhs[1] cubeMatIx (int32, 1-based medium index), x fastest as in MATLAB's column-major order */
    mxArray *cubeRhoA = mx_new_double(3, dims), *cubeMatA = mx_new_int32(3, dims);
    memcpy(cubeRhoA->data, F(&b, "med_densities"), nvox * sizeof(double));
    memcpy(cubeMatA->data, I(&b, "med_indices"), nvox * sizeof(int));
    /* prhs[2] ompMCgeo: material (n x 1 cell of strings), xBounds / yBounds / zBounds */
    int nmat = 1;
    for (const char *c = names; *c; c++) nmat += (*c == ',');
    mxArray *geo = mx_new_struct(), *mat = mx_new_cell((mwSize)nmat, 1);
    {
        char *copy = strdup(names), *save = NULL;
        int k = 0;
        for (char *tok = strtok_r(copy, ",", &save); tok && k < nmat; tok = strtok_r(NULL, ",", &save)) mat->fields[k++] = mxCreateString(tok);
        free(copy);
    }
    mx_set_field(geo, "material", mat);
    mx_set_field(geo, "xBounds", dbl_from(F(&b, "xbounds"), (mwSize)isize + 1));
    mx_set_field(geo, "yBounds", dbl_from(F(&b, "ybounds"), (mwSize)jsize + 1));
    mx_set_field(geo, "zBounds", dbl_from(F(&b, "zbounds"), (mwSize)ksize + 1));
    /* prhs[3] ompMCsource: nBixels, iBeam (1-based, double), beam sources, bixel corners and sides */
    const int nbix = I(&b, "mr_nbeamlets")[0];
    const mwSize nbeams = (mwSize)blob_find(&b, "mr_xsource")->count;
    mxArray *src = mx_new_struct(), *ibeam = mx_new_double_row((mwSize)nbix, NULL);
    for (int i = 0; i < nbix; i++) ((double *)ibeam->data)[i] = (double)(I(&b, "mr_ibeam")[i] + 1);
    mx_set_field(src, "nBixels", dbl_scalar((double)nbix));
    mx_set_field(src, "iBeam", ibeam);
    mx_set_field(src, "xSource", dbl_from(F(&b, "mr_xsource"), nbeams)); mx_set_field(src, "ySource", dbl_from(F(&b, "mr_ysource"), nbeams));
    mx_set_field(src, "zSource", dbl_from(F(&b, "mr_zsource"), nbeams));
    mx_set_field(src, "xCorner", dbl_from(F(&b, "mr_xcorner"), (mwSize)nbix)); mx_set_field(src, "yCorner", dbl_from(F(&b, "mr_ycorner"), (mwSize)nbix));
    mx_set_field(src, "zCorner", dbl_from(F(&b, "mr_zcorner"), (mwSize)nbix));
    mx_set_field(src, "xSide1", dbl_from(F(&b, "mr_xside1"), (mwSize)nbix)); mx_set_field(src, "ySide1", dbl_from(F(&b, "mr_yside1"), (mwSize)nbix));
    mx_set_field(src, "zSide1", dbl_from(F(&b, "mr_zside1"), (mwSize)nbix));
    mx_set_field(src, "xSide2", dbl_from(F(&b, "mr_xside2"), (mwSize)nbix)); mx_set_field(src, "ySide2", dbl_from(F(&b, "mr_yside2"), (mwSize)nbix));
    mx_set_field(src, "zSide2", dbl_from(F(&b, "mr_zside2"), (mwSize)nbix));
    /* prhs[4] ompMCoptions (omc_matrad.c:93-245) */
    mxArray *opt = mx_new_struct();
    const double seeds[2] = {97, 33};
    mx_set_field(opt, "verbose", mx_new_logical_scalar(0));
    mx_set_field(opt, "nHistories", dbl_scalar(nhist));
    mx_set_field(opt, "nBatches", dbl_scalar(nbatch));
    mx_set_field(opt, "nSplit", dbl_scalar((double)I(&b, "nsplit")[0]));
    mx_set_field(opt, "spectrumFile", mxCreateString(spectrum ? spectrum : ""));
    mx_set_field(opt, "monoEnergy", dbl_scalar(mono));
    mx_set_field(opt, "charge", dbl_scalar((double)I(&b, "src_charge")[0]));
    mx_set_field(opt, "global_ecut", dbl_scalar(ecut));
    mx_set_field(opt, "global_pcut", dbl_scalar(pcut));
    mx_set_field(opt, "randomSeeds", dbl_from(seeds, 2));
    mx_set_field(opt, "pegsFile", mxCreateString(pegs));
    mx_set_field(opt, "pgs4formFile", mxCreateString(form));
    mx_set_field(opt, "dataFolder", mxCreateString(datadir));
    mx_set_field(opt, "outputFolder", mxCreateString("./"));
    mx_set_field(opt, "relDoseThreshold", dbl_scalar(rel));

    const mxArray *prhs[5] = {cubeRhoA, cubeMatA, geo, src, opt};
    mxArray *plhs[1] = {NULL};
    mexFunction(1, plhs, 5, prhs);
    if (omc_dropin_dump_only) return EXIT_SUCCESS;

    /* the sparse matrix as the CSC file of omc_matrad_b200 */
    const mxArray *D = plhs[0];
    const long long hdr[3] = {(long long)nvox, nbix, (long long)mxGetJc(D)[nbix]};
    char *fn = malloc(strlen(stem) + 16);
    sprintf(fn, "%s.csc", stem);
    FILE *fp = fopen(fn, "wb");
    if (!fp) { printf("Unable to open file: %s\n", fn); return EXIT_FAILURE; }
    fwrite("OMCCSC1", 1, 8, fp);
    fwrite(hdr, sizeof(long long), 3, fp);
    fwrite(mxGetJc(D), sizeof(mwIndex), (size_t)nbix + 1, fp);
    fwrite(mxGetIr(D), sizeof(mwIndex), (size_t)hdr[2], fp);
    fwrite(mxGetPr(D), sizeof(double), (size_t)hdr[2], fp);
    fclose(fp);
    free(fn);
    return EXIT_SUCCESS;
}
#endif

/*
 * omc_matrad_b200.c -- plain-C entry point of the matRad user code on the B200 hot path: what is left of
 * ucodes/omc_matrad/omc_matrad.c's mexFunction (:1258-1543) when MATLAB is taken away.
 *
 *   history bookkeeping                 omc_matrad.c:1371-1383   (nhist, nbatch, nperbatch; int semantics kept)
 *   beamlet loop + accumulateResults +  omc_matrad.c:1389-1493   -> omc_gpu_run_beamlets(): `group` beamlets per pass of the
 *   threshold + sparse column assembly                              wavefront kernels, columns assembled on the device
 *   mxCreateSparse(nvox, nbeamlets)     omc_matrad.c:1339-1350   -> a CSC file: the Jc / Ir / Pr arrays of the sparse matrix
 *
 * Output file <stem>.csc (little endian): "OMCCSC1\0", int64 nrows (= voxels), int64 ncols (= beamlets), int64 nnz,
 * int64 jc[ncols+1], int64 ir[nnz] (voxel index irl-1, ascending inside a column), double pr[nnz] (Gy per history as
 * accumulateResults(1, nhist, nbatch) normalises it).
 *
 * usage: omc_matrad_b200 -p problem.blob -n nHistories -b nbatch -t relDoseThreshold -o out_stem [-g beamlets per pass, default 64] [-d device]
 *        [-r rank -w world]   (one process per GPU: this rank's groups of consecutive beamlets only, columns of the others left empty)
 *        [-G ngpu]            (ONE process over ngpu GPUs of the node, or OMC_GPUS: omc_gpu_multi_run_beamlets() deals whole passes to the
 *                              devices and gathers the column slices in beamlet order -- the complete matrix in one file)
 * There is no CPU transport here: without a CUDA device the program exits with the library's error.
 */
#include <math.h>
#include <time.h>

#include "omc_host_common.h"

static omc_gpu_handle gpu;
static void die(const char *what) {
    printf("%s: %s\n", what, omc_gpu_last_error(gpu));
    exit(EXIT_FAILURE);
}
static double now_s(void) {
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec;
}
static uint64_t count_of(const blob *b, const char *name) { return blob_find(b, name)->count; }

int main(int argc, char **argv) {
    const char *pfile = NULL, *ncase = "100000", *nbatch_s = "10", *stem = "omc_matrad_b200", *seeds = "97 33";
    double rel = 1.0e-3;
    int device = 0, group = 0, rank = 0, world = 1, ngpu = getenv("OMC_GPUS") ? atoi(getenv("OMC_GPUS")) : 1;
    for (int i = 1; i < argc; i++) {
        if (!strcmp(argv[i], "-p") && i + 1 < argc) pfile = argv[++i];
        else if (!strcmp(argv[i], "-n") && i + 1 < argc) ncase = argv[++i];
        else if (!strcmp(argv[i], "-b") && i + 1 < argc) nbatch_s = argv[++i];
        else if (!strcmp(argv[i], "-t") && i + 1 < argc) rel = atof(argv[++i]);
        else if (!strcmp(argv[i], "-o") && i + 1 < argc) stem = argv[++i];
        else if (!strcmp(argv[i], "-g") && i + 1 < argc) group = atoi(argv[++i]);
        else if (!strcmp(argv[i], "-d") && i + 1 < argc) device = atoi(argv[++i]);
        else if (!strcmp(argv[i], "-r") && i + 1 < argc) rank = atoi(argv[++i]);
        else if (!strcmp(argv[i], "-w") && i + 1 < argc) world = atoi(argv[++i]);
        else if (!strcmp(argv[i], "-s") && i + 1 < argc) seeds = argv[++i];
        else if (!strcmp(argv[i], "-G") && i + 1 < argc) ngpu = atoi(argv[++i]);
        else {
            printf("usage: %s -p problem.blob -n nHistories -b nbatch -t relDoseThreshold -o out_stem [-g beamlets per pass, default 64] [-d device] [-r rank -w world | -G ngpu]\n",
                   argv[0]);
            return 2;
        }
    }
    if (!pfile) { printf("Can not find the problem file (-p).\n"); return 2; }
    if (world < 1 || rank < 0 || rank >= world) { printf("rank/world out of range.\n"); return 2; }
    const double tbegin = now_s();
    blob b;
    if (blob_read(&b, pfile) != 0) { printf("Unable to open file: %s\n", pfile); return EXIT_FAILURE; }
    omc_media_tables t;
    omc_geometry g;
    const double *dens;
    host_load_media_geometry(&b, &t, &g, &dens);
    const long long nvox = (long long)g.isize * g.jsize * g.ksize;

    omc_source_matrad s;
    memset(&s, 0, sizeof s);
    s.spectrum = I(&b, "src_spectrum")[0]; s.charge = I(&b, "src_charge")[0]; s.energy = F(&b, "src_energy")[0];
    s.deltak = F(&b, "src_deltak")[0]; s.cdfinv1 = F(&b, "src_cdfinv1"); s.cdfinv2 = F(&b, "src_cdfinv2");
    s.nbixels = I(&b, "mr_nbeamlets")[0]; s.nbeams = (int)count_of(&b, "mr_xsource");
    s.ibeam = I(&b, "mr_ibeam");
    s.xsource = F(&b, "mr_xsource"); s.ysource = F(&b, "mr_ysource"); s.zsource = F(&b, "mr_zsource");
    s.xcorner = F(&b, "mr_xcorner"); s.ycorner = F(&b, "mr_ycorner"); s.zcorner = F(&b, "mr_zcorner");
    s.xside1 = F(&b, "mr_xside1"); s.yside1 = F(&b, "mr_yside1"); s.zside1 = F(&b, "mr_zside1");
    s.xside2 = F(&b, "mr_xside2"); s.yside2 = F(&b, "mr_yside2"); s.zside2 = F(&b, "mr_zside2");
    const int nsplit = I(&b, "nsplit")[0], nbeamlets = s.nbixels;

    printf("Number of voxels on each direction (X,Y,Z) : (%d, %d, %d)\n", g.isize, g.jsize, g.ksize);
    printf("Number of beamlets : %d (%d beams)\n", nbeamlets, s.nbeams);
    if (ngpu > 1) {                                             /* the whole node behind one handle */
        omc_gpu_multi m;
        if (omc_gpu_multi_create(&m, ngpu, NULL)) { printf("No %d CUDA devices / no NCCL: this program has no CPU transport path.\n", ngpu); return EXIT_FAILURE; }
#define MD(call) do { if (call) { printf("%s: %s\n", #call, omc_gpu_multi_last_error(m)); exit(EXIT_FAILURE); } } while (0)
        MD(omc_gpu_multi_set_media(m, &t)); MD(omc_gpu_multi_set_geometry(m, &g)); MD(omc_gpu_multi_set_source_matrad(m, &s));
        MD(omc_gpu_multi_set_vrt(m, nsplit));
        int ixx = 97, jxx = 33;
        sscanf(seeds, "%d %d", &ixx, &jxx);
        MD(omc_gpu_multi_set_seed(m, ixx, jxx)); MD(omc_gpu_multi_set_option(m, "kernel", OMC_KERNEL_WAVEFRONT));
        int nhist = atoi(ncase), nbatch = atoi(nbatch_s);
        if (nbatch <= 0) { printf("Can not find 'nbatch' key on input file.\n"); return EXIT_FAILURE; }
        if (nhist / nbatch == 0) nhist = nbatch;
        nhist = (nhist / nbatch) * nbatch;
        printf("Total number of particle histories: %d\nGPUs: %d\n", nhist, omc_gpu_multi_size(m));
        if (group < 1) {
            long long capn = (long long)(OMC_BEAMLET_GRID_BUDGET / ((double)(nvox + 1) * 4.0));
            group = capn > OMC_BEAMLETS_PER_PASS ? OMC_BEAMLETS_PER_PASS : (capn < 1 ? 1 : (int)capn);
        }
        long long *jc = calloc((size_t)nbeamlets + 1, sizeof(long long)), tot = 0;
        const double t0 = now_s();
        printf("Execution time up to this point : %8.2f seconds\n", t0 - tbegin);
        MD(omc_gpu_multi_run_beamlets(m, 0, nhist, nbatch, 0, nbeamlets, group, rel, dens, jc, &tot));
        long long *ir = malloc((size_t)(tot ? tot : 1) * sizeof(long long));
        double *val = malloc((size_t)(tot ? tot : 1) * sizeof(double));
        MD(omc_gpu_multi_fetch_columns(m, ir, val));
        const double t1 = now_s();
        printf("Simulation finished\n");
        printf("Beamlets computed: %d on %d GPUs, histories per second: %.4g, non-zeros: %lld (%.3f %% of the matrix)\n", nbeamlets,
               omc_gpu_multi_size(m), (double)nbeamlets * nhist / (t1 - t0), tot, 100.0 * (double)tot / ((double)nvox * nbeamlets));
        char *fn = malloc(strlen(stem) + 16);
        sprintf(fn, "%s.csc", stem);
        FILE *fp = fopen(fn, "wb");
        if (!fp) { printf("Unable to open file: %s\n", fn); return EXIT_FAILURE; }
        const long long hdr[3] = {nvox, nbeamlets, tot};
        fwrite("OMCCSC1", 1, 8, fp);
        fwrite(hdr, sizeof(long long), 3, fp);
        fwrite(jc, sizeof(long long), (size_t)nbeamlets + 1, fp);
        fwrite(ir, sizeof(long long), (size_t)tot, fp);
        fwrite(val, sizeof(double), (size_t)tot, fp);
        fclose(fp);
        omc_gpu_multi_destroy(m);
        printf("Total execution time : %8.5f seconds\n", now_s() - tbegin);
        return EXIT_SUCCESS;
    }
    if (omc_gpu_create(&gpu, device)) { printf("No CUDA device: this program has no CPU transport path.\n"); return EXIT_FAILURE; }
    if (omc_gpu_set_media(gpu, &t)) die("omc_gpu_set_media");
    if (omc_gpu_set_geometry(gpu, &g)) die("omc_gpu_set_geometry");
    if (omc_gpu_set_source_matrad(gpu, &s)) die("omc_gpu_set_source_matrad");
    if (omc_gpu_set_vrt(gpu, nsplit)) die("omc_gpu_set_vrt");
    int ixx = 97, jxx = 33;
    sscanf(seeds, "%d %d", &ixx, &jxx);
    omc_gpu_set_seed(gpu, ixx, jxx);
    if (omc_gpu_set_option(gpu, "kernel", OMC_KERNEL_WAVEFRONT)) die("omc_gpu_set_option");

    /* history bookkeeping as omc_matrad.c:1371-1383 (same as omc_dosxyz.c:1207-1225) */
    int nhist = atoi(ncase), nbatch = atoi(nbatch_s);
    if (nbatch <= 0) { printf("Can not find 'nbatch' key on input file.\n"); return EXIT_FAILURE; }
    if (nhist / nbatch == 0) nhist = nbatch;
    const int nperbatch = nhist / nbatch;
    nhist = nperbatch * nbatch;
    printf("Total number of particle histories: %d\n", nhist);
    printf("Number of statistical batches: %d\n", nbatch);
    printf("Histories per batch: %d\n", nperbatch);

    /* columns of this rank, in beamlet order; beamlet b always owns history ids [b*nhist, (b+1)*nhist) */
    long long *jc = calloc((size_t)nbeamlets + 1, sizeof(long long));
    long long *ncol = calloc((size_t)nbeamlets, sizeof(long long));
    long long **cir = calloc((size_t)nbeamlets, sizeof(long long *));
    double **cval = calloc((size_t)nbeamlets, sizeof(double *));
    if (group < 1) {            /* default: 64 beamlets per pass, fewer when their dose grids would not fit the HBM budget */
        long long capn = (long long)(OMC_BEAMLET_GRID_BUDGET / ((double)(nvox + 1) * 4.0));
        group = capn > OMC_BEAMLETS_PER_PASS ? OMC_BEAMLETS_PER_PASS : (int)capn;
        if (group < 1) group = 1;
    }
    if (group > nbeamlets) group = nbeamlets > 0 ? nbeamlets : 1;
    printf("Beamlets per pass: up to %d\n", group);
    long long *gjc = malloc(((size_t)group + 1) * sizeof(long long));
    const double t0 = now_s();
    printf("Execution time up to this point : %8.2f seconds\n", t0 - tbegin);
    long long done = 0;
    /* Sharding plan (same as ompmc_b200.matrad.beamlet_groups): contiguous groups of at most `group` beamlets -- one pass of the
     * wavefront kernels each, which pays the tail of its longest particle lineages once per GROUP, so a rank owns whole groups --
     * as many groups as a multiple of `world`, dealt round-robin. */
    const int ngroups = world * ((nbeamlets + world * group - 1) / (world * group));
    const int gsize = ngroups > 0 ? (nbeamlets + ngroups - 1) / ngroups : 1;
    for (int b0 = 0; b0 < nbeamlets; b0 += gsize) {
        if ((b0 / gsize) % world != rank) continue;
        const int nb = (nbeamlets - b0 < gsize) ? nbeamlets - b0 : gsize;
        long long tot = 0;
        if (omc_gpu_run_beamlets(gpu, (long long)b0 * nhist, nhist, nbatch, b0, nb, rel, dens, gjc, &tot)) die("omc_gpu_run_beamlets");
        long long *ir = malloc((size_t)(tot ? tot : 1) * sizeof(long long));
        double *val = malloc((size_t)(tot ? tot : 1) * sizeof(double));
        if (omc_gpu_fetch_columns(gpu, ir, val)) die("omc_gpu_fetch_columns");
        for (int k = 0; k < nb; k++) {
            const long long n = gjc[k + 1] - gjc[k];
            ncol[b0 + k] = n;
            cir[b0 + k] = malloc((size_t)(n ? n : 1) * sizeof(long long));
            cval[b0 + k] = malloc((size_t)(n ? n : 1) * sizeof(double));
            memcpy(cir[b0 + k], ir + gjc[k], (size_t)n * sizeof(long long));
            memcpy(cval[b0 + k], val + gjc[k], (size_t)n * sizeof(double));
        }
        free(ir); free(val);
        done += nb;
    }
    const double t1 = now_s();
    for (int b2 = 0; b2 < nbeamlets; b2++) jc[b2 + 1] = jc[b2] + ncol[b2];
    printf("Simulation finished\n");
    printf("Beamlets computed by this rank: %lld, histories per second: %.4g, non-zeros: %lld (%.3f %% of the matrix)\n", done,
           (double)done * nhist / (t1 - t0), jc[nbeamlets], 100.0 * (double)jc[nbeamlets] / ((double)nvox * nbeamlets));

    char *fn = malloc(strlen(stem) + 16);
    sprintf(fn, "%s.csc", stem);
    FILE *fp = fopen(fn, "wb");
    if (!fp) { printf("Unable to open file: %s\n", fn); return EXIT_FAILURE; }
    const long long hdr[3] = {nvox, nbeamlets, jc[nbeamlets]};
    fwrite("OMCCSC1", 1, 8, fp);
    fwrite(hdr, sizeof(long long), 3, fp);
    fwrite(jc, sizeof(long long), (size_t)nbeamlets + 1, fp);
    for (int b2 = 0; b2 < nbeamlets; b2++) if (ncol[b2]) fwrite(cir[b2], sizeof(long long), (size_t)ncol[b2], fp);
    for (int b2 = 0; b2 < nbeamlets; b2++) if (ncol[b2]) fwrite(cval[b2], sizeof(double), (size_t)ncol[b2], fp);
    fclose(fp);
    omc_gpu_destroy(gpu);
    printf("Total execution time : %8.5f seconds\n", now_s() - tbegin);
    return EXIT_SUCCESS;
}

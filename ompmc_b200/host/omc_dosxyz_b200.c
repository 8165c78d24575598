/*
 * omc_dosxyz_b200.c -- plain-C host driver of the B200 hot path: the part of ompMC's omc_dosxyz user code that
 * sits around the batch loop, talking to the GPU only through the C-ABI of include/ompmc_b200.h.
 *
 *   main() batch bookkeeping + batch loop      ucodes/omc_dosxyz/omc_dosxyz.c:1207-1263   (atoi/int semantics kept)
 *   accumulateResults()                        omc_dosxyz.c:719-799
 *   outputResults() -> <stem>.3ddose           omc_dosxyz.c:801-886   (same printf formats, byte-compatible)
 *
 * What it does NOT contain: the reference's table builders (initMediaData & co).  Their output -- plus phantom,
 * regions and source, i.e. everything a reference user code holds in its globals just before the batch loop --
 * is read from a problem blob (format: "OMCBLOB1", see ompmc_b200/problem.py save_blob/load_blob).  A maintainer
 * who links the reference's own init code instead follows INTEGRATION.md; the calls below are the same.
 *
 * usage: omc_dosxyz_b200 -p problem.blob -n ncase -b nbatch -o out_stem [-k 0|1] [-d device | -g ngpu] [-s "ixx jxx"]
 * There is no CPU transport here: without a CUDA device the program exits with the library's error.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "ompmc_b200.h"

#include "omc_host_common.h"
#include "omc_host_input.h"

/* ---- accumulateResults(), omc_dosxyz.c:719-799 ---------------------------------------------- */
static void accumulate_results(const omc_geometry *g, const double *dens, double *accum, double *accum2, int iout, int nhist, int nbatch) {
    const int imax = g->isize, ijmax = g->isize * g->jsize;
    const double inc_fluence = (double)nhist;
    for (int iz = 0; iz < g->ksize; iz++)
        for (int iy = 0; iy < g->jsize; iy++)
            for (int ix = 0; ix < g->isize; ix++) {
                const int irl = 1 + ix + iy * imax + iz * ijmax;
                double endep = accum[irl] / (double)nbatch, endep2 = accum2[irl] / (double)nbatch, unc;
                if (endep != 0.0) {
                    unc = endep2 - endep * endep;
                    unc /= (double)(nbatch - 1);
                    unc = sqrt(unc) / endep;
                } else {
                    endep = 0.0; unc = 0.9999999;
                }
                if (iout) {
                    double mass = (g->xbounds[ix + 1] - g->xbounds[ix]) * (g->ybounds[iy + 1] - g->ybounds[iy]) *
                                  (g->zbounds[iz + 1] - g->zbounds[iz]);
                    mass *= dens[irl - 1];
                    endep *= 1.602E-10 / (mass * inc_fluence);
                } else {
                    endep /= inc_fluence;
                }
                if (dens[irl - 1] < 0.044) { endep = 0.0; unc = 0.9999999; }     /* "zero dose in air" */
                accum[irl] = endep;
                accum2[irl] = unc;
            }
}

/* ---- outputResults(), omc_dosxyz.c:841-879 -------------------------------------------------- */
static int write_3ddose(const char *stem, const omc_geometry *g, const double *dose, const double *unc) {
    char *fn = malloc(strlen(stem) + 16);
    sprintf(fn, "%s.3ddose", stem);
    FILE *fp = fopen(fn, "w");
    if (!fp) { printf("Unable to open file: %s\n", fn); free(fn); return 1; }
    const int imax = g->isize, ijmax = g->isize * g->jsize;
    fprintf(fp, "%5d%5d%5d\n", g->isize, g->jsize, g->ksize);
    for (int i = 0; i <= g->isize; i++) fprintf(fp, "%f ", g->xbounds[i]);
    fprintf(fp, "\n");
    for (int i = 0; i <= g->jsize; i++) fprintf(fp, "%f ", g->ybounds[i]);
    fprintf(fp, "\n");
    for (int i = 0; i <= g->ksize; i++) fprintf(fp, "%f ", g->zbounds[i]);
    fprintf(fp, "\n");
    for (int iz = 0; iz < g->ksize; iz++)
        for (int iy = 0; iy < g->jsize; iy++)
            for (int ix = 0; ix < g->isize; ix++) fprintf(fp, "%e ", dose[1 + ix + iy * imax + iz * ijmax]);
    fprintf(fp, "\n");
    for (int iz = 0; iz < g->ksize; iz++)
        for (int iy = 0; iy < g->jsize; iy++)
            for (int ix = 0; ix < g->isize; ix++) fprintf(fp, "%f ", unc[1 + ix + iy * imax + iz * ijmax]);
    fprintf(fp, "\n");
    fclose(fp);
    free(fn);
    return 0;
}

static double now_s(void) {
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec;
}

static omc_gpu_multi gpu;       /* one handle over -g N devices (default 1) */
static void die(const char *what) {
    printf("%s: %s\n", what, omc_gpu_multi_last_error(gpu));
    exit(EXIT_FAILURE);
}

int main(int argc, char **argv) {
    const char *pfile = NULL, *ifile = NULL, *ncase = NULL, *nbatch_s = NULL, *stem = "omc_b200", *seeds = NULL, *voxel = NULL;
    int kernel = -1, device = 0, dump_only = 0, ngpu = getenv("OMC_GPUS") ? atoi(getenv("OMC_GPUS")) : 1;
    for (int i = 1; i < argc; i++) {
        if (!strcmp(argv[i], "-p") && i + 1 < argc) pfile = argv[++i];
        else if (!strcmp(argv[i], "-i") && i + 1 < argc) ifile = argv[++i];
        else if (!strcmp(argv[i], "-n") && i + 1 < argc) ncase = argv[++i];
        else if (!strcmp(argv[i], "-b") && i + 1 < argc) nbatch_s = argv[++i];
        else if (!strcmp(argv[i], "-o") && i + 1 < argc) stem = argv[++i];
        else if (!strcmp(argv[i], "-k") && i + 1 < argc) kernel = atoi(argv[++i]);
        else if (!strcmp(argv[i], "-d") && i + 1 < argc) device = atoi(argv[++i]);
        else if (!strcmp(argv[i], "-g") && i + 1 < argc) ngpu = atoi(argv[++i]);
        else if (!strcmp(argv[i], "-s") && i + 1 < argc) seeds = argv[++i];
        else if (!strcmp(argv[i], "-v") && i + 1 < argc) voxel = argv[++i];
        else if (!strcmp(argv[i], "--dump-problem")) dump_only = 1;
        else {
            printf("usage: %s (-i input_stem | -p problem.blob) [-n ncase] [-b nbatch] -o out_stem [-k 0|1] [-d device | -g ngpu] [-s \"ixx jxx\"]\n"
                   "  -g ngpu         GPUs of this node to spread every batch over (default 1, or OMC_GPUS); -d picks the device when ngpu = 1\n"
                   "  -i input_stem   the reference's own input file <input_stem>.inp (phantom, PEGS4, XCOM, spectrum ... read and\n"
                   "                  initialised here, no reference code involved), exactly like `omc_dosxyz -i input_stem -o out_stem`\n"
                   "  -v \"dx dy dz\"   with -i: resample the phantom to voxels of about this size in cm (any ratio; mass-conserving), BASELINE config 5\n"
                   "  -p problem.blob tables + phantom + source dumped from an initialised user code\n"
                   "  --dump-problem  with -i: write <out_stem>.problem (every array handed to the GPU library) and exit, no GPU needed\n",
                   argv[0]);
            return 2;
        }
    }
    if (!pfile && !ifile) { printf("Can not find the input (-i) or problem (-p) file.\n"); return 2; }
    const double tbegin = now_s();
    omc_media_tables t;
    omc_geometry g;
    const double *dens;
    omc_source_dosxyz s;
    int nsplit = 1;
    char seed_buf[384], ncase_buf[384], nbatch_buf[384];
    if (ifile) {
        /* the reference's main() up to the batch loop, omc_dosxyz.c:1155-1205, on our own initialisers */
        static inp_file inp;
        static host_phantom ph;
        static host_regions reg;
        static double cdfinv1[OMC_INVDIM], cdfinv2[OMC_INVDIM];
        char path[384], folder[384], pegs[384], ffile[384], buf[384], err[512];
        if (inp_parse(&inp, ifile)) return EXIT_FAILURE;
        if (!inp_path(&inp, "phantom file", path) || phantom_read(&ph, path)) return EXIT_FAILURE;
        printf("Path to phantom file : %s\n", path);
        if (voxel) {                                             /* config 5: the same phantom on a finer (or coarser) grid */
            double vx = 0, vy = 0, vz = 0;
            static host_phantom fine;
            if (sscanf(voxel, "%lf %lf %lf", &vx, &vy, &vz) != 3 || phantom_resample(&ph, vx, vy, vz, &fine)) { printf("Bad -v voxel sizes.\n"); return EXIT_FAILURE; }
            printf("Phantom resampled from (%d, %d, %d) to (%d, %d, %d) voxels\n", ph.isize, ph.jsize, ph.ksize, fine.isize, fine.jsize, fine.ksize);
            ph = fine;
        }
        if (!inp_path(&inp, "pegs file", pegs) || !inp_path(&inp, "data folder", folder) || !inp_path(&inp, "pgs4form file", ffile)) return EXIT_FAILURE;
        const char *names[OMC_MXMED];
        for (int i = 0; i < ph.nmed; i++) names[i] = ph.names[i];
        omc_tables *tb = omc_tables_build(folder, pegs, ffile, ph.nmed, names, err, sizeof err);
        if (!tb) { printf("%s\n", err); return EXIT_FAILURE; }
        t = *omc_tables_view(tb);
        if (source_init(&s, &inp, &ph, cdfinv1, cdfinv2)) return EXIT_FAILURE;
        if (!inp_get(&inp, "global ecut", buf)) { printf("Can not find 'global ecut' key on input file.\n"); return EXIT_FAILURE; }
        const double ecut = atof(buf);
        if (!inp_get(&inp, "global pcut", buf)) { printf("Can not find 'global pcut' key on input file.\n"); return EXIT_FAILURE; }
        const double pcut = atof(buf);
        regions_init(&reg, &ph, &t, ecut, pcut);
        if (!inp_get(&inp, "nsplit", buf)) { printf("Can not find 'nsplit' key on input file.\n"); return EXIT_FAILURE; }
        nsplit = atoi(buf);
        printf(nsplit > 1 ? "Photon splitting enabled, nsplit = %d\n" : "Photon splitting disabled\n", nsplit);
        g.isize = ph.isize; g.jsize = ph.jsize; g.ksize = ph.ksize; g.xbounds = ph.xb; g.ybounds = ph.yb; g.zbounds = ph.zb;
        g.med = reg.med; g.rhof = reg.rhof; g.pcut = reg.pcut; g.ecut = reg.ecut;
        dens = ph.dens;
        if (!seeds && inp_get(&inp, "rng seeds", seed_buf)) seeds = seed_buf;
        if (!ncase) { if (!inp_get(&inp, "ncase", ncase_buf)) { printf("Can not find 'ncase' key on input file.\n"); return EXIT_FAILURE; } ncase = ncase_buf; }
        if (!nbatch_s) { if (!inp_get(&inp, "nbatch", nbatch_buf)) { printf("Can not find 'nbatch' key on input file.\n"); return EXIT_FAILURE; } nbatch_s = nbatch_buf; }
        if (dump_only) return host_dump_problem(stem, &t, &g, dens, &s, nsplit) ? EXIT_FAILURE : EXIT_SUCCESS;
        /* outputResults() writes "output folder" + output_file + ".3ddose" (omc_dosxyz.c:822-833): a relative -o name goes there */
        static char out_stem[800];
        if (stem[0] != '/' && inp_path(&inp, "output folder", folder)) {
            snprintf(out_stem, sizeof out_stem, "%s%s", folder, stem);
            stem = out_stem;
        }
    } else {
        static blob b;
        if (blob_read(&b, pfile) != 0) { printf("Unable to open file: %s\n", pfile); return EXIT_FAILURE; }
        host_load_media_geometry(&b, &t, &g, &dens);
        memset(&s, 0, sizeof s);
        s.spectrum = I(&b, "src_spectrum")[0]; s.charge = I(&b, "src_charge")[0]; s.energy = F(&b, "src_energy")[0];
        s.deltak = F(&b, "src_deltak")[0]; s.cdfinv1 = F(&b, "src_cdfinv1"); s.cdfinv2 = F(&b, "src_cdfinv2"); s.ssd = F(&b, "src_ssd")[0];
        s.xinl = F(&b, "src_xinl")[0]; s.xinu = F(&b, "src_xinu")[0]; s.yinl = F(&b, "src_yinl")[0]; s.yinu = F(&b, "src_yinu")[0];
        s.xsize = F(&b, "src_xsize")[0]; s.ysize = F(&b, "src_ysize")[0];
        s.ixinl = I(&b, "src_ixinl")[0]; s.ixinu = I(&b, "src_ixinu")[0]; s.iyinl = I(&b, "src_iyinl")[0]; s.iyinu = I(&b, "src_iyinu")[0];
        nsplit = I(&b, "nsplit")[0];
    }
    if (!ncase) ncase = "100000";
    if (!nbatch_s) nbatch_s = "10";
    if (!seeds) seeds = "97 33";
    const int gridsize = g.isize * g.jsize * g.ksize;

    printf("Number of media in phantom : %d\n", t.nmed);
    printf("Number of voxels on each direction (X,Y,Z) : (%d, %d, %d)\n", g.isize, g.jsize, g.ksize);
    if (omc_gpu_multi_create(&gpu, ngpu > 1 ? ngpu : 1, ngpu > 1 ? NULL : &device)) {
        printf("No CUDA device (or no NCCL for %d of them): this program has no CPU transport path.\n", ngpu);
        return EXIT_FAILURE;
    }
    printf("GPUs: %d\n", omc_gpu_multi_size(gpu));
    if (omc_gpu_multi_set_media(gpu, &t)) die("omc_gpu_multi_set_media");
    if (omc_gpu_multi_set_geometry(gpu, &g)) die("omc_gpu_multi_set_geometry");
    if (omc_gpu_multi_set_source_dosxyz(gpu, &s)) die("omc_gpu_multi_set_source_dosxyz");
    if (omc_gpu_multi_set_vrt(gpu, nsplit)) die("omc_gpu_multi_set_vrt");
    int ixx = 97, jxx = 33;
    sscanf(seeds, "%d %d", &ixx, &jxx);
    omc_gpu_multi_set_seed(gpu, ixx, jxx);
    if (kernel < 0) kernel = nsplit > 255 ? OMC_KERNEL_LOCKSTEP : OMC_KERNEL_WAVEFRONT;
    if (omc_gpu_multi_set_option(gpu, "kernel", kernel)) die("omc_gpu_multi_set_option");

    /* batch bookkeeping exactly as omc_dosxyz.c:1207-1225 */
    int nhist = atoi(ncase), nbatch = atoi(nbatch_s);
    if (nbatch <= 0) { printf("Can not find 'nbatch' key on input file.\n"); return EXIT_FAILURE; }
    if (nhist / nbatch == 0) nhist = nbatch;
    const int nperbatch = nhist / nbatch;
    nhist = nperbatch * nbatch;
    printf("Total number of particle histories: %d\n", nhist);
    printf("Number of statistical batches: %d\n", nbatch);
    printf("Histories per batch: %d\n", nperbatch);
    printf("Execution time up to this point : %8.2f seconds\n", now_s() - tbegin);

    const double t0 = now_s();
    for (int ibatch = 0; ibatch < nbatch; ibatch++) {
        if (ibatch == 0) printf("%-10s\t%-15s\n", "Batch #", "Elapsed time");
        printf("%-10d\t%-15.2f\n", ibatch, now_s() - tbegin);
        if (omc_gpu_multi_run_batch(gpu, (long long)ibatch * nperbatch, nperbatch, -1)) die("omc_gpu_multi_run_batch");
    }
    double *accum = malloc(((size_t)gridsize + 1) * sizeof(double)), *accum2 = malloc(((size_t)gridsize + 1) * sizeof(double)), ensrc = 0.0;
    if (omc_gpu_multi_get_tallies(gpu, accum, accum2, &ensrc)) die("omc_gpu_multi_get_tallies");
    const double t1 = now_s();
    printf("Simulation finished\n");
    printf("Execution time up to this point : %8.2f seconds\n", t1 - tbegin);
    printf("Histories per second (batch loop): %.4g\n", (double)nhist / (t1 - t0));
    double etot = 0.0;
    for (int irl = 1; irl < gridsize + 1; irl++) etot += accum[irl];
    printf("Fraction of incident energy deposited in the phantom: %5.4f\n", etot / ensrc);
    /* accumulateResults(iout = 1, nperbatch, nbatch), omc_dosxyz.c:1281-1282: on the device by default (the tallies are
     * resident there); OMC_HOST_RESULTS=1 runs the host restatement above instead -- same numbers, bit for bit */
    const int host_results = getenv("OMC_HOST_RESULTS") ? atoi(getenv("OMC_HOST_RESULTS")) : 0;
    const double t2 = now_s();
    if (host_results == 1) {                /* host restatement of both halves of outputResults() */
        accumulate_results(&g, dens, accum, accum2, 1, nperbatch, nbatch);
        if (write_3ddose(stem, &g, accum, accum2)) return EXIT_FAILURE;
    } else if (host_results == 2) {         /* statistics on the device, the reference's fprintf loop on the host */
        if (omc_gpu_multi_accumulate_results(gpu, 1, nperbatch, nbatch, dens, accum, accum2)) die("omc_gpu_multi_accumulate_results");
        if (write_3ddose(stem, &g, accum, accum2)) return EXIT_FAILURE;
    } else if (host_results == 3) {         /* probe: both writers on the SAME tallies -> <stem>.3ddose (device) and <stem>_host.3ddose */
        char *fn = malloc(strlen(stem) + 32);
        sprintf(fn, "%s.3ddose", stem);
        double ta = now_s();
        if (omc_gpu_multi_write_3ddose(gpu, fn, 1, nperbatch, nbatch, dens)) die("omc_gpu_multi_write_3ddose");
        printf("Device writer: %8.3f seconds\n", now_s() - ta);
        ta = now_s();
        if (omc_gpu_multi_accumulate_results(gpu, 1, nperbatch, nbatch, dens, accum, accum2)) die("omc_gpu_multi_accumulate_results");
        sprintf(fn, "%s_host", stem);
        if (write_3ddose(fn, &g, accum, accum2)) return EXIT_FAILURE;
        printf("Host writer: %8.3f seconds\n", now_s() - ta);
        free(fn);
    } else {                                /* default: statistics AND the text of the file on the device */
        char *fn = malloc(strlen(stem) + 16);
        sprintf(fn, "%s.3ddose", stem);
        if (omc_gpu_multi_write_3ddose(gpu, fn, 1, nperbatch, nbatch, dens)) die("omc_gpu_multi_write_3ddose");
        free(fn);
    }
    printf("Output written in %8.3f seconds\n", now_s() - t2);
    omc_gpu_multi_destroy(gpu);
    printf("Total execution time : %8.5f seconds\n", now_s() - tbegin);
    return EXIT_SUCCESS;
}

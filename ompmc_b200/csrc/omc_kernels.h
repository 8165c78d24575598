// omc_kernels.h -- host-callable launchers of the CUDA kernels (internal to libompmc_b200.so).
#pragma once
#include "omc_types.cuh"

namespace omc {

// omc_lockstep.cu
void launch_lockstep(const DevProblem &P, Part *stack, int depth, int blocks, int threads, long long first, long long nhist,
                     int ibeamlet, const Part *inject, cudaStream_t stream);
int lockstep_blocks_per_sm(int threads);
void launch_test_geometry(const DevProblem &P, int n, const double *xyzuvw, const int *ir, const double *ustep_in, int *idisc,
                          int *irnew, double *ustep_out, double *tperp, cudaStream_t stream);
void launch_test_rng(uint32_t s0, uint32_t s1, unsigned long long hist, int n, double *out, cudaStream_t stream);
void launch_accum(double *endep, double *accum, double *accum2, long long n, cudaStream_t stream);
void launch_results(const DevProblem &P, const double *accum, const double *accum2, const double *dens, int iout, int nhist, int nbatch,
                    double *dose, double *unc, cudaStream_t stream);


// ---- omc_wavefront.cu ---------------------------------------------------------------------------
// One particle queue in HBM, structure-of-arrays with 16-byte lanes (one 128-bit access per warp and lane):
//   xy = {x, y}   ze = {z, e}   (fp64: positions and energies)        dw = {u, v, w, wt} (fp32: direction cosines and weight)
//   irq = {ir, iq | tag << 16}   rng = {hist_lo, hist_hi, stream, draws consumed}
// 72 bytes per particle (the first layout kept direction and weight in fp64: 88).  Directions leave the samplers in fp32
// precision anyway (omc_physics_f32.cuh) or are rounded once per interaction; weights are 1, 1/nsplit or nsplit/nsplit.
// Photon queues only (aux != nullptr): aux = {mfp left (-1: not sampled yet; Woodcock flight: -1 / -2 = not yet / already
// inside the phantom box), eta' of the running split copy}; for photons in flight tag = isplit | i_survive << 8.
struct PartQueue {
    double2 *xy, *ze;
    float4 *dw;
    double2 *aux;
    int2 *rm;        // electron queues only (else nullptr): voxel record {float rhof, int med} of region ir, med = -2: not known
    int2 *irq;
    uint4 *rng;
    unsigned cap;
};
constexpr size_t PART_QUEUE_BYTES_PER_SLOT = 2 * sizeof(double2) + sizeof(float4) + sizeof(int2) + sizeof(uint4);   // + 16 photons, + 8 electrons

constexpr int WAVE_THREADS = 128;   // threads per block == particles per chunk

// A counter alone in its 256-byte line: every hot counter of a wave is hit by one atomic per warp and iteration from
// the whole machine, and the L2 atomic unit serialises per line (all counters in ONE line made the reservation
// atomics the top stall of esize_kernel).
struct alignas(256) PadU {
    unsigned v;
};

struct WaveCtl {
    PadU n_p[2], n_e[2], n_ip[2], n_ie[2];       // queue fill counts, [parity]: cur = parity, next = parity ^ 1
    PadU n_ch, n_bca;                            // step-class queues, filled and drained inside one wave
    PadU tk[5];                                  // chunk tickets per class (misc_kernel)
    PadU overflow, drain_ticket;
    PadU old_seen;                               // particles of the PREVIOUS batch met by the consumers of this wave
    unsigned n_src;                              // histories injected by the current wave
    unsigned parity, target, live, waves;
    // batch pipelining: histories with id < hist_split belong to the previous batch, whose tail is still in flight
    // while this batch is injected; they score into dose grid (grid_new ^ 1), everything else into grid_new
    unsigned has_old, old_done, grid_new;
    unsigned old_last;                           // old_seen of the last completed wave (how many stragglers the old batch has left)
    unsigned long long hist_split;
    unsigned long long hist_next, hist_end;
};

// electrons between "step size known" and "step taken", 120 bytes (seven 16-byte lanes + one 8-byte lane):
//   v[0..2] = {x,y} {z,e} {tustep,range}                       fp64
//   d = {u, v, w, wt}   f = {demfp, sig0, rhof, dedx}          fp32 (fp32-born, or only enter ratios)
//   m = {float blccl, float ssmfp, ir, iq+1 | (imed+1) << 2 | flags << 6 | lelke << 16}
//   t = {float tperp (rounded down), float elke}               rng as in PartQueue
// flags replace the fp64 total_tstep of electron() :4787-4830, which the step only needs for its test "was this step the
// whole distance to the next interaction" (:5290): bit 0 = yes if the step is taken in full, bit 1 = yes whatever the step.
// (The first layouts: 21 doubles + 2 x 16 B = 200 B, then nine 16-byte lanes = 144 B.)
// ONE set of arrays of 2*cap slots holds both step classes: condensed-history steps fill slots 0, 1, 2, ... and
// boundary-crossing steps 2*cap-1, 2*cap-2, ..., so esize_kernel stores with a single converged code path.
struct EStepQueue {
    double2 *v[3];
    float4 *d, *f;
    uint4 *m;
    float2 *t;
    uint4 *rng;
    unsigned cap;                                // per class
};
constexpr size_t ESTEP_QUEUE_BYTES_PER_SLOT = 3 * sizeof(double2) + 2 * sizeof(float4) + 2 * sizeof(uint4) + sizeof(float2);

struct WaveQueues {
    PartQueue p[2], e[2], ip[2], ie[2];
    EStepQueue es;
};

struct WaveLaunch {
    int blocks[4], max_cross, ibeamlet, woodcock, max_virtual;
    float *mb_grid;                 // multi-beamlet pass (see WaveArgs), nullptr: off
    unsigned long long mb_first;
    unsigned mb_per, mb_n;
    int mb_ib0;
};

void wave_blocks_per_sm(int out[4]);

// omc_lockstep.cu: finish the last particles of a wavefront run one per thread (see drain_kernel);
// queues = {P, E, IP, IE} of the current wave, tags as in omc_wavefront.cu (TAG_*)
struct DrainArgs {
    PartQueue q[4];
    const unsigned *count[4];
    unsigned *ticket;
};
void launch_drain(const DevProblem &P, const DrainArgs &D, Part *stack, int depth, int blocks, cudaStream_t stream);
// streams / events of one wave: s = electron chain; s2 = misc_kernel; s3 = the boundary-crossing step kernel, which runs
// beside the condensed-history one (both only read the step queue); s2 == nullptr: everything in order on s
struct WaveStreams {
    cudaStream_t s, s2, s3;
    cudaEvent_t fork, join, fork3, join3;
};
void launch_wave(const DevProblem &P, WaveCtl *ctl, const WaveQueues &Q, const WaveLaunch &L, const WaveStreams &W);
// unit-test hook: one production sampler on explicit inputs (omc_gpu_test_samplers)
void launch_test_samplers(const DevProblem &P, int which, int n, const double *in, unsigned long long first, double *out, cudaStream_t s);
void launch_flush(float *g32, double *g64, long long n, cudaStream_t s);
// multi-beamlet pass: per-beamlet maximum (pass 0) / count above threshold (pass 1), then ordered compaction into CSC
void launch_mb_scan(const float *grids, long long nreg, int nb, const DevProblem &P, const double *dens, int nhist, int nbatch, double rel,
                    double *dmax, unsigned long long *nnz, int pass, cudaStream_t s);
void launch_mb_fill(const float *grids, long long nreg, int nb, const DevProblem &P, const double *dens, int nhist, int nbatch, double rel,
                    const double *dmax, const long long *jc, long long *ir, double *val, cudaStream_t s);
// .3ddose text on the device (omc_format.cuh): mode 0 = "%e " (13 bytes per value), 1 = "%f " (9 bytes); values the device does
// not certify go to fb[0..*nfb) (capped at fb_cap, *nfb keeps counting) for the host to format
struct Pow10;
struct FormatFallback {
    unsigned long long index;
    double value;
};
void launch_format(int mode, const double *src, long long n, const Pow10 *tab, char *out, FormatFallback *fb, unsigned *nfb, unsigned fb_cap,
                   cudaStream_t stream);
// start the next batch while the tail of the previous one is still in the queues (see WaveCtl::hist_split)
void launch_rearm(WaveCtl *ctl, unsigned long long first, unsigned long long nhist, unsigned nsplit, cudaStream_t s);

}  // namespace omc

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


// ---- omc_wavefront.cu ---------------------------------------------------------------------------
// One particle queue in HBM, structure-of-arrays (coalesced 8-byte lanes); irq = {ir, iq | tag << 16},
// rng = {hist_lo, hist_hi, stream, draws consumed}; aux = photon mfp left (-1: not sampled yet), aux2 = eta' of the
// running split copy (photon splitting); for photons in flight tag = isplit | i_survive << 8.
struct PartQueue {
    double *x, *y, *z, *u, *v, *w, *e, *wt, *aux, *aux2;
    int2 *irq;
    uint4 *rng;
    unsigned cap;
};
constexpr size_t PART_QUEUE_BYTES_PER_SLOT = 10 * sizeof(double) + sizeof(int2) + sizeof(uint4);

constexpr int WAVE_THREADS = 128;   // threads per block == particles per chunk

struct WaveCtl {
    unsigned n_p[2], n_e[2], n_ip[2], n_ie[2];   // queue fill counts, [parity]: cur = parity, next = parity ^ 1
    unsigned n_ch[2], n_bca[2];                  // step-class queues, [parity]: filled by esize_kernel (cur) and the edo kernels (next)
    unsigned tk[5];                              // chunk tickets per class (misc_kernel)
    unsigned n_src;                              // histories injected by the current wave
    unsigned parity, target, overflow, live, waves, drain_ticket;
    unsigned long long hist_next, hist_end;
};

// electrons between "step size known" and "step taken": Part + EStep (21 doubles) + {ir, iq, lelke, imed} + rng
struct EStepQueue {
    double *d[21];
    uint4 *w[2];
    unsigned cap;
};

struct WaveQueues {
    PartQueue p[2], e[2], ip[2], ie[2];
    EStepQueue ch[2], bca[2];
};

struct WaveLaunch {
    int blocks[4], max_cross, electron_iters, ibeamlet, woodcock, max_virtual;
};

void wave_blocks_per_sm(int out[4]);

// omc_lockstep.cu: finish the last particles of a wavefront run one per thread (see drain_kernel);
// queues = {P, E, IP, IE} of the current wave, tags as in omc_wavefront.cu (TAG_*)
struct DrainArgs {
    PartQueue q[4];
    const unsigned *count[4];
    EStepQueue sq[2];               // electrons waiting in the step-class queues (their step state is dropped: resampled)
    const unsigned *scount[2];
    unsigned *ticket;
};
void launch_drain(const DevProblem &P, const DrainArgs &D, Part *stack, int depth, int blocks, cudaStream_t stream);
void launch_wave(const DevProblem &P, WaveCtl *ctl, const WaveQueues &Q, const WaveLaunch &L, cudaStream_t s, cudaStream_t s2,
                 cudaEvent_t fork, cudaEvent_t join);
void launch_flush(float *g32, double *g64, long long n, cudaStream_t s);

}  // namespace omc

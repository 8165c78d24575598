// omc_capi.cu -- C-ABI of libompmc_b200.so (declared in include/ompmc_b200.h).
//
// Host side only: packs the reference-layout tables into the device records of omc_types.cuh,
// owns device memory / the stream, launches the kernels.  There is NO CPU transport path in this
// library: every entry point that does physics launches a CUDA kernel or fails.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "omc_kernels.h"
#include "omc_format.cuh"
#include "omc_nccl.h"

using namespace omc;

// Device buffers of one set_*() call are kept in a pool and REUSED by the next call when the sizes repeat
// (a user code that re-uploads the same-shaped problem pays no cudaMalloc/cudaFree, only the copies).
struct BufPool {
    std::vector<void *> ptr;
    std::vector<size_t> bytes;
    size_t next = 0;
    void begin() { next = 0; }
    void clear() {
        for (void *p : ptr) cudaFree(p);
        ptr.clear(); bytes.clear(); next = 0;
    }
};

struct omc_gpu_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    std::string err;
    DevProblem P{};
    BufPool media_bufs, geom_bufs, source_bufs;
    bool have_media = false, have_geom = false, have_source = false;
    double *accum = nullptr, *accum2 = nullptr;
    int tally_nreg = -1;
    const Part *inject = nullptr;   // unit-test hook (omc_gpu_test_particles)
    // options
    int kernel = OMC_KERNEL_LOCKSTEP;
    int threads_per_block = 128;
    int stack_depth = 0;          // 0 = auto from nsplit
    int max_blocks = 0;           // 0 = auto (SMs x occupancy)
    int record = 0;
    // scratch
    Part *stack = nullptr;
    size_t stack_bytes = 0;
    omc_history_record *records = nullptr;
    long long records_cap = 0, last_nhist = 0;
    unsigned long long launches = 0;
    // wavefront state
    std::vector<MedRec> med_host;
    double cut_e[OMC_MXMED] = {0}, cut_p[OMC_MXMED] = {0}, rho_max[OMC_MXMED] = {0};
    bool cuts_uniform = false, med_dirty = false;
    WaveQueues wq{};
    std::vector<void *> wave_bufs;
    WaveCtl *ctl = nullptr;        // device
    WaveCtl *ctl_host = nullptr;   // pinned: the latest status examined by the host loop
    WaveCtl *ctl_slot[2] = {nullptr, nullptr};   // pinned: status read-backs in flight (the host loop looks one group ahead)
    cudaEvent_t ev_stat[2] = {nullptr, nullptr};
    // the `check_every` waves captured as a CUDA graph, kept across calls while the launch parameters do not change
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t gexec = nullptr;
    std::vector<char> graph_key;
    // multi-GPU (omc_gpu_comm_init): completed batch grids are summed over the ranks on a side stream, off the transport stream
    nccl_comm_t comm = nullptr;
    int rank = 0, world = 1;
    cudaStream_t side = nullptr;
    cudaEvent_t ev_done[2] = {nullptr, nullptr}, ev_free[2] = {nullptr, nullptr};
    bool grid_wait[2] = {false, false};          // grid g may be scored into only after ev_free[g]
    unsigned long long reduces = 0;
    unsigned pool_target = 1u << 23;   // particles kept in flight (B200: 2 Mi 1.03e8, 4 Mi 1.07e8, 8 Mi 1.09e8 histories/s)
    unsigned pool_cap = 0, pool_cap_opt = 0;
    int max_cross = 16, check_every = 16;
    int photon_tracking = 1;      // 0: voxel-to-voxel march as in photon(); 1: Woodcock flight when nsplit == 1
    int max_virtual = 8;          // Woodcock: tentative collisions per photon per wave
    unsigned long long waves = 0;
    int trace = 0, use_graph = 1, overlap = 1, source_kind = 0, lookahead = 1;
    cudaStream_t stream2 = nullptr, stream3 = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_fork3 = nullptr, ev_join3 = nullptr;
    unsigned drain_threshold = 8192;     // measured on B200 (16M-history call): 0 -> 183.5 ms, 8 Ki -> 182.4, 32 Ki -> 193.9, 128 Ki -> 223.5
    // omc_gpu_accumulate_results scratch
    double *res_dens = nullptr, *res_dose = nullptr, *res_unc = nullptr;
    int res_nreg = -1;
    // omc_gpu_write_3ddose: power-of-ten table, two text buffers (device + pinned host), fallback lists
    Pow10 *fmt_tab = nullptr;
    bool fmt_ready = false;
    char *fmt_dev[2] = {nullptr, nullptr}, *fmt_host[2] = {nullptr, nullptr};
    FormatFallback *fmt_fb_dev[2] = {nullptr, nullptr}, *fmt_fb_host[2] = {nullptr, nullptr};
    unsigned *fmt_nfb_dev = nullptr, *fmt_nfb_host = nullptr;
    cudaEvent_t fmt_ev[2] = {nullptr, nullptr};
    // batch pipelining (see wave_run)
    int run_grid = -1, last_ibeamlet = -1;
    long long hist_hi = 0;         // end of the history-id range of the batch in flight (pipelining needs ascending ids)
    std::vector<int> done_q;
    bool auto_acc[2] = {false, false};
    bool pipeline_next = false, pipeline_auto = false;   // how the next omc_gpu_run_histories() body is to run (set by the callers below)
    // multi-beamlet pass (omc_gpu_run_beamlets)
    float *mb_grid = nullptr;
    size_t mb_grid_elems = 0;
    bool mb_active = false;
    unsigned long long mb_first = 0;
    unsigned mb_per = 0, mb_n = 0;
    int mb_ib0 = 0;
    double *mb_dmax = nullptr;
    unsigned long long *mb_nnz = nullptr;
    long long *mb_jc = nullptr, *mb_ir = nullptr;
    double *mb_val = nullptr;
    int mb_meta_cap = 0;
    long long mb_col_cap = 0, mb_total = 0;
};

#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) {                                                                         \
            h->err = std::string(#call) + ": " + cudaGetErrorString(e_);                                 \
            return 1;                                                                                    \
        }                                                                                                \
    } while (0)

static int fail(omc_gpu_handle h, const char *msg) {
    h->err = msg;
    return 2;
}

template <typename T>
static int upload(omc_gpu_handle h, BufPool &pool, const T *host, size_t n, const T **dev) {
    const size_t nb = (n ? n : 1) * sizeof(T);
    void *d = nullptr;
    if (pool.next < pool.ptr.size() && pool.bytes[pool.next] == nb) {
        d = pool.ptr[pool.next];
    } else {
        if (pool.next < pool.ptr.size()) {          // shape changed: drop this and all later buffers
            for (size_t i = pool.next; i < pool.ptr.size(); i++) cudaFree(pool.ptr[i]);
            pool.ptr.resize(pool.next); pool.bytes.resize(pool.next);
        }
        CK(cudaMalloc(&d, nb));
        pool.ptr.push_back(d); pool.bytes.push_back(nb);
    }
    pool.next++;
    if (n) CK(cudaMemcpyAsync(d, host, n * sizeof(T), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    *dev = static_cast<const T *>(d);
    return 0;
}

static void free_pool(std::vector<void *> &pool) {
    for (void *p : pool) cudaFree(p);
    pool.clear();
}

static int alloc_queue(omc_gpu_handle h, PartQueue &q, unsigned cap, bool photons, bool electrons = false) {
    q.cap = cap;
    double2 **f[2] = {&q.xy, &q.ze};
    for (auto pp : f) {
        CK(cudaMalloc((void **)pp, (size_t)cap * sizeof(double2)));
        h->wave_bufs.push_back(*pp);
    }
    CK(cudaMalloc((void **)&q.dw, (size_t)cap * sizeof(float4)));
    h->wave_bufs.push_back(q.dw);
    q.aux = nullptr;
    if (photons) {
        CK(cudaMalloc((void **)&q.aux, (size_t)cap * sizeof(double2)));
        h->wave_bufs.push_back(q.aux);
    }
    q.rm = nullptr;
    if (electrons) {
        CK(cudaMalloc((void **)&q.rm, (size_t)cap * sizeof(int2)));
        h->wave_bufs.push_back(q.rm);
    }
    CK(cudaMalloc((void **)&q.irq, (size_t)cap * sizeof(int2)));
    h->wave_bufs.push_back(q.irq);
    CK(cudaMalloc((void **)&q.rng, (size_t)cap * sizeof(uint4)));
    h->wave_bufs.push_back(q.rng);
    return 0;
}

static int alloc_estep_queue(omc_gpu_handle h, EStepQueue &q, unsigned cap) {
    q.cap = cap;                                  // per step class; both classes share arrays of 2 * cap slots
    for (int i = 0; i < 3; i++) {
        CK(cudaMalloc((void **)&q.v[i], (size_t)2 * cap * sizeof(double2)));
        h->wave_bufs.push_back(q.v[i]);
    }
    CK(cudaMalloc((void **)&q.d, (size_t)2 * cap * sizeof(float4)));
    h->wave_bufs.push_back(q.d);
    CK(cudaMalloc((void **)&q.f, (size_t)2 * cap * sizeof(float4)));
    h->wave_bufs.push_back(q.f);
    CK(cudaMalloc((void **)&q.t, (size_t)2 * cap * sizeof(float2)));
    h->wave_bufs.push_back(q.t);
    CK(cudaMalloc((void **)&q.m, (size_t)2 * cap * sizeof(uint4)));
    h->wave_bufs.push_back(q.m);
    CK(cudaMalloc((void **)&q.rng, (size_t)2 * cap * sizeof(uint4)));
    h->wave_bufs.push_back(q.rng);
    return 0;
}

// ---- wavefront driver -------------------------------------------------------------------------------
// Two dose grids (fp32 chunk grid + fp64 batch grid each) let the NEXT batch start while the tail of the previous
// one is still in the queues: a particle scores into the grid of the batch its history id belongs to
// (WaveCtl::hist_split).  h->run_grid = grid of the batch whose tail is in flight (-1: queues empty);
// h->done_q = grids of completed batches that have not been accumulated yet (accumEndep), oldest first.
static int wave_prepare(omc_gpu_handle h) {
    const unsigned target = h->pool_target;
    const unsigned cap = h->pool_cap_opt ? h->pool_cap_opt : 2u * target + 65536u;
    if (h->pool_cap != cap) {
        free_pool(h->wave_bufs);
        h->pool_cap = 0;
        h->run_grid = -1;                         // (whatever was in flight is gone with the queues)
        for (int i = 0; i < 2; i++) {
            if (alloc_queue(h, h->wq.p[i], cap, true)) return 1;
            if (alloc_queue(h, h->wq.e[i], cap, false, true)) return 1;
            if (alloc_queue(h, h->wq.ip[i], cap, false)) return 1;
            if (alloc_queue(h, h->wq.ie[i], cap, false)) return 1;
        }
        if (alloc_estep_queue(h, h->wq.es, cap)) return 1;
        h->pool_cap = cap;
    }
    if (!h->stream2) {
        CK(cudaStreamCreateWithFlags(&h->stream2, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&h->stream3, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&h->ev_fork3, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&h->ev_join3, cudaEventDisableTiming));
    }
    if (!h->ctl) {
        CK(cudaMalloc((void **)&h->ctl, sizeof(WaveCtl)));
        CK(cudaMallocHost((void **)&h->ctl_host, sizeof(WaveCtl)));
        memset(h->ctl_host, 0, sizeof(WaveCtl));
        for (int i = 0; i < 2; i++) {
            CK(cudaMallocHost((void **)&h->ctl_slot[i], sizeof(WaveCtl)));
            CK(cudaEventCreateWithFlags(&h->ev_stat[i], cudaEventDisableTiming));
        }
    }
    return 0;
}

static int grid_done(omc_gpu_handle h, int g) {      // fold the fp32 chunk grid of a completed batch into its fp64 batch grid
    const size_t off = (size_t)g * h->P.nreg;
    launch_flush(h->P.endep32 + off, h->P.endep + off, h->P.nreg, h->stream);
    h->launches += 1;
    h->done_q.push_back(g);
    return 0;
}

static bool in_done(omc_gpu_handle h, int g) {
    for (int d : h->done_q) if (d == g) return true;
    return false;
}

// drain_kernel (omc_lockstep.cu) over four queues: one thread follows each queued particle and its descendants to the end,
// scoring straight into the fp64 grid `g` of the batch they belong to
static int drain_queues(omc_gpu_handle h, const PartQueue *const q[4], const unsigned *const cnt[4], int g) {
    const int tpb = 128;
    const int blocks = h->sm_count * lockstep_blocks_per_sm(tpb);
    const int depth = 48;
    const size_t need = (size_t)depth * blocks * tpb * sizeof(Part);
    if (need > h->stack_bytes) {
        cudaFree(h->stack);
        h->stack = nullptr; h->stack_bytes = 0;
        CK(cudaMalloc((void **)&h->stack, need));
        h->stack_bytes = need;
    }
    DrainArgs D;
    for (int k = 0; k < 4; k++) { D.q[k] = *q[k]; D.count[k] = cnt[k]; }
    D.ticket = &h->ctl->drain_ticket.v;
    CK(cudaMemsetAsync(&h->ctl->drain_ticket.v, 0, sizeof(unsigned), h->stream));
    DevProblem Pd = h->P;
    Pd.endep = h->P.endep + (size_t)g * h->P.nreg;
    launch_drain(Pd, D, h->stack, depth, blocks, h->stream);
    h->launches += 1;
    return 0;
}

static void drop_graph(omc_gpu_handle h) {
    if (h->gexec) { cudaGraphExecDestroy(h->gexec); h->gexec = nullptr; }
    if (h->graph) { cudaGraphDestroy(h->graph); h->graph = nullptr; }
    h->graph_key.clear();
}

// start == true : inject histories [first, first+nhist) and return once all of them are started AND the previous batch
//                 (if its tail was in flight) is complete; the tail of the new batch stays in the queues.
// start == false: run what is in the queues to the end (waves, then drain_kernel for the last few particles).
// g_new: dose grid of the new batch (start only).
static int wave_run(omc_gpu_handle h, bool start, long long first, long long nhist, int ibeamlet, int g_new) {
    DevProblem &P = h->P;
    if (!start && h->run_grid < 0) return 0;
    if (wave_prepare(h)) return 1;
    const unsigned target = h->pool_target;
    int g_old = h->run_grid;
    bool old_pending = false;
    if (start) {
        if (h->grid_wait[g_new]) {                              // the grid's previous batch is still being summed over the ranks
            CK(cudaStreamWaitEvent(h->stream, h->ev_free[g_new], 0));
            h->grid_wait[g_new] = false;
        }
        if (g_old < 0) {                                        // empty pipeline
            WaveCtl c;
            memset(&c, 0, sizeof c);
            c.target = target;
            c.hist_next = (unsigned long long)first; c.hist_end = (unsigned long long)(first + nhist);
            c.hist_split = (unsigned long long)first; c.grid_new = (unsigned)g_new; c.old_done = 1;
            const unsigned first_room = target / (unsigned)(P.nsplit > 1 ? 16 * P.nsplit : 1);  // (splitting: see advance_kernel)
            c.n_src = (unsigned)((unsigned long long)nhist < (unsigned long long)first_room ? nhist : first_room);
            CK(cudaMemcpyAsync(h->ctl, &c, sizeof c, cudaMemcpyHostToDevice, h->stream));
            CK(cudaStreamSynchronize(h->stream));               // (c is a stack object)
        } else {                                                // previous batch still in flight: it becomes "old"
            launch_rearm(h->ctl, (unsigned long long)first, (unsigned long long)nhist, (unsigned)P.nsplit, h->stream);
            old_pending = true;
        }
        h->last_ibeamlet = ibeamlet;
    } else {
        const WaveCtl &s0 = *h->ctl_host;                       // status as of the last check
        if (s0.live == 0 && s0.n_src == 0 && s0.hist_next >= s0.hist_end) {
            h->run_grid = -1;
            return grid_done(h, g_old);
        }
        ibeamlet = h->last_ibeamlet;
    }
    WaveLaunch L;
    memset(&L, 0, sizeof L);
    int occ[4];
    wave_blocks_per_sm(occ);
    for (int i = 0; i < 4; i++) L.blocks[i] = h->max_blocks > 0 ? h->max_blocks : h->sm_count * occ[i];
    L.max_cross = h->max_cross; L.ibeamlet = ibeamlet;
    L.woodcock = (h->photon_tracking == 1 && P.nsplit == 1) ? 1 : 0;
    L.max_virtual = h->max_virtual > 0 ? ((h->max_virtual + 1) & ~1) : 8;   // even: whole Philox blocks, so results do not depend on it
    L.mb_grid = h->mb_active ? h->mb_grid : nullptr;
    L.mb_first = h->mb_first; L.mb_per = h->mb_per; L.mb_n = h->mb_n; L.mb_ib0 = h->mb_ib0;
    WaveStreams W;
    W.s = h->stream; W.s2 = h->overlap ? h->stream2 : nullptr; W.s3 = h->stream3;
    W.fork = h->ev_fork; W.join = h->ev_join; W.fork3 = h->ev_fork3; W.join3 = h->ev_join3;
    const int every = h->check_every > 0 ? h->check_every : 1;
    // `every` waves are captured once into a CUDA graph (all launch parameters are wave-invariant: cur/next parity lives
    // in WaveCtl on the device), so the host issues one graph launch per `every` waves.  The instantiated graph is kept
    // across calls for as long as nothing that the kernels receive by value changes (problem, queues, launch shape).
    if (h->use_graph) {
        std::vector<char> key(sizeof(DevProblem) + sizeof(WaveLaunch) + sizeof(WaveQueues) + 2 * sizeof(int));
        char *k = key.data();
        memcpy(k, &P, sizeof(DevProblem)); k += sizeof(DevProblem);
        memcpy(k, &L, sizeof(WaveLaunch)); k += sizeof(WaveLaunch);
        memcpy(k, &h->wq, sizeof(WaveQueues)); k += sizeof(WaveQueues);
        memcpy(k, &every, sizeof(int)); k += sizeof(int);
        memcpy(k, &h->overlap, sizeof(int));
        if (!h->gexec || key != h->graph_key) {
            drop_graph(h);
            CK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
            for (int i = 0; i < every; i++) launch_wave(P, h->ctl, h->wq, L, W);
            CK(cudaStreamEndCapture(h->stream, &h->graph));
            CK(cudaGraphInstantiate(&h->gexec, h->graph, 0));
            h->graph_key = key;
        }
    } else {
        drop_graph(h);
    }
    int rc = 0;
    bool drained = false, done = false;
    // The host looks ONE group of waves ahead while the pool is busy: group k+1 is enqueued before the status read-back of
    // group k is examined, so the device never idles for the host's wake-up + launch latency (it did, once per 16 waves,
    // and with 8 ranks on one socket that showed as rank skew).  Every decision below is monotone in the status (a batch
    // that was complete stays complete), so acting one group late only appends waves over queues that are emptier.
    // Near the end (few particles alive) the look-ahead is dropped: the extra group would be pure latency there.
    int issued = 0, examined = 0;
    bool ahead = false;
    for (unsigned long long wave = 0; !done; wave += every) {
        const int slot = issued & 1;
        if (h->gexec) {
            CK(cudaGraphLaunch(h->gexec, h->stream));
        } else {
            for (int i = 0; i < every; i++) launch_wave(P, h->ctl, h->wq, L, W);
        }
        h->launches += 5ull * every;
        h->waves += every;
        // Fold the fp32 chunk grids into the fp64 batch grids every few checks, not only when a batch completes: with 1e8
        // histories per batch a hot voxel's fp32 sum reaches 1e4 MeV, where one ulp is 1e-3 MeV.  (Between two graph launches
        // nothing else runs on this stream, so the plain read-add-zero of flush_kernel is safe.)
        if (!h->mb_active && ((wave / every) & 3ull) == 3ull) {
            const int gs[2] = {start ? g_new : g_old, (start && old_pending) ? g_old : -1};
            for (int g : gs)
                if (g >= 0) {
                    const size_t off = (size_t)g * P.nreg;
                    launch_flush(P.endep32 + off, P.endep + off, P.nreg, h->stream);
                    h->launches += 1;
                }
        }
        CK(cudaMemcpyAsync(h->ctl_slot[slot], h->ctl, sizeof(WaveCtl), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaEventRecord(h->ev_stat[slot], h->stream));
        issued += 1;
        while (!done && issued - examined > (ahead ? 1 : 0)) {
            const int es = examined & 1;
            CK(cudaEventSynchronize(h->ev_stat[es]));
            memcpy(h->ctl_host, h->ctl_slot[es], sizeof(WaveCtl));
            examined += 1;
            const WaveCtl &s = *h->ctl_host;
            if (h->trace)
                fprintf(stderr, "wave %llu live %u n_src %u P %u E %u IP %u IE %u hist_next %llu old %u/%u\n",
                        (unsigned long long)examined * every, s.live, s.n_src, s.n_p[s.parity].v, s.n_e[s.parity].v, s.n_ip[s.parity].v,
                        s.n_ie[s.parity].v, s.hist_next, s.has_old, s.old_done);
            if (s.overflow.v) {
                h->err = "particle queue overflow on the device: increase option pool_size";
                rc = 7; done = true;
                break;
            }
            const bool exhausted = s.hist_next >= s.hist_end && s.n_src == 0;
            ahead = h->lookahead && !exhausted && s.live > (1u << 18);
            if (old_pending && s.old_done) {                    // the previous batch has left the queues: its grid is final
                grid_done(h, g_old);
                old_pending = false;
            }
            if (start) {
                // the tail of the new batch stays in flight; it is always the NEXT call that completes a batch, even an
                // already empty one, so that every rank of a multi-GPU run sees the same sequence of completed batches
                if (exhausted && !old_pending) done = true;
            } else if (exhausted && s.live == 0) {
                done = true;
            } else if (exhausted && s.live <= h->drain_threshold && P.nsplit == 1 && !h->mb_active) {
                // (split photons in flight cannot be handed over; the drain scores into ONE fp64 grid, not per beamlet)
                drained = true; done = true;
            }
        }
        if (!done && wave > 50000000ull) { rc = fail(h, "wavefront did not terminate"); break; }
    }
    if (issued > examined) {                                    // a group was still in flight: its status is the current one
        CK(cudaEventSynchronize(h->ev_stat[(issued - 1) & 1]));
        memcpy(h->ctl_host, h->ctl_slot[(issued - 1) & 1], sizeof(WaveCtl));
    }
    if (rc) { h->run_grid = -1; return rc; }
    if (start) {
        h->run_grid = g_new;
        CK(cudaGetLastError());
        return 0;
    }
    if (drained && h->ctl_host->live > 0) {
        // few particles left: one thread follows each to the end (omc_lockstep.cu: drain_kernel)
        const int par = (int)h->ctl_host->parity;
        const PartQueue *q[4] = {&h->wq.p[par], &h->wq.e[par], &h->wq.ip[par], &h->wq.ie[par]};
        const unsigned *cnt[4] = {&h->ctl->n_p[par].v, &h->ctl->n_e[par].v, &h->ctl->n_ip[par].v, &h->ctl->n_ie[par].v};
        if (drain_queues(h, q, cnt, g_old)) return 1;
    }
    memset(h->ctl_host, 0, sizeof(WaveCtl));                    // (status: nothing alive any more)
    h->run_grid = -1;
    grid_done(h, g_old);
    CK(cudaGetLastError());
    return 0;
}

static int free_grid(omc_gpu_handle h) {
    const int busy = h->run_grid;
    for (int g = 0; g < 2; g++)
        if (g != busy && !in_done(h, g)) return g;
    return -1;
}

// accumEndep() of every completed batch that is waiting for it (oldest first); `only_auto`: just those started by
// omc_gpu_run_batch(), which owes them an accumulation.  With a communicator (omc_gpu_comm_init) the completed grid is first
// summed over the ranks -- before accumEndep() squares it, so the statistics are those of a single-GPU run -- and both steps
// run on the SIDE stream, ordered after the grid's final fold by an event: the transport stream goes on with the next batch
// and only waits (ev_free) when this grid is needed again, one whole batch later.  A rank that finishes its slice early no
// longer holds the others' waves behind its collective.
static int accum_done(omc_gpu_handle h, bool only_auto) {
    std::vector<int> keep;
    bool took = false;
    for (int g : h->done_q) {
        if ((only_auto && !h->auto_acc[g]) || (!only_auto && took)) { keep.push_back(g); continue; }   // explicit call: the oldest one only
        took = true;
        const size_t off = (size_t)g * h->P.nreg;
        if (h->comm) {
            CK(cudaEventRecord(h->ev_done[g], h->stream));
            CK(cudaStreamWaitEvent(h->side, h->ev_done[g], 0));
            const int e = nccl_api().AllReduce(h->P.endep + off, h->P.endep + off, (size_t)h->P.nreg, NCCL_FLOAT64, NCCL_SUM, h->comm, h->side);
            if (e) { h->err = std::string("ncclAllReduce: ") + nccl_api().GetErrorString(e); return 1; }
            launch_accum(h->P.endep + off, h->accum, h->accum2, h->P.nreg, h->side);
            CK(cudaEventRecord(h->ev_free[g], h->side));
            h->grid_wait[g] = true;
            h->reduces += 1;
        } else {
            launch_accum(h->P.endep + off, h->accum, h->accum2, h->P.nreg, h->stream);
        }
        h->launches += 1;
        h->auto_acc[g] = false;
    }
    h->done_q = keep;
    if (cudaGetLastError() != cudaSuccess) { h->err = "accum kernel launch failed"; return 1; }
    return 0;
}

// complete whatever is in flight and settle the accumulations owed by omc_gpu_run_batch()
static int flush_all(omc_gpu_handle h) {
    if (h->run_grid >= 0) {
        int rc = wave_run(h, false, 0, 0, -1, -1);
        if (rc) return rc;
    }
    return accum_done(h, true);
}

// histories [first, first+nhist) through the wavefront kernels.  pipelined: leave the tail of this batch in flight
// (omc_gpu_start_batch / omc_gpu_run_batch); otherwise run it to the end (omc_gpu_run_histories).
static int run_wavefront(omc_gpu_handle h, long long first, long long nhist, int ibeamlet, bool pipelined, bool auto_acc) {
    int g;
    if (!pipelined) {
        int rc = flush_all(h);
        if (rc) return rc;
        // several un-accumulated calls in a row keep adding into the same batch grid, as they always did
        if (!h->done_q.empty()) { g = h->done_q.back(); h->done_q.pop_back(); }
        else g = 0;
    } else {
        // A particle in flight is attributed to the previous or to the new batch by `history id < first` (WaveCtl::hist_split):
        // a batch that re-uses or lowers ids cannot share the queues with the tail of the previous one -- complete that first.
        if (h->run_grid >= 0 && first < h->hist_hi) {
            int rc = wave_run(h, false, 0, 0, -1, -1);
            if (rc) return rc;
            if (auto_acc && (rc = accum_done(h, true))) return rc;
        }
        g = free_grid(h);
        if (g < 0) return fail(h, "two batches are waiting for omc_gpu_accum_batch(): accumulate before starting another one");
    }
    h->auto_acc[g] = auto_acc;
    int rc = wave_run(h, true, first, nhist, ibeamlet, g);
    if (rc) return rc;
    h->hist_hi = first + nhist;
    if (!pipelined) rc = wave_run(h, false, 0, 0, -1, -1);
    return rc;
}

extern "C" {

int omc_gpu_create(omc_gpu_handle *out, int device_id) {
    if (!out) return 2;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
        fprintf(stderr, "ompmc_b200: no CUDA device available; this library has no CPU fallback\n");
        return 3;
    }
    if (device_id < 0 || device_id >= ndev) return 4;
    omc_gpu_handle h = new omc_gpu_ctx();
    h->device = device_id;
    if (cudaSetDevice(device_id) != cudaSuccess || cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete h;
        return 5;
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device_id) == cudaSuccess) h->sm_count = prop.multiProcessorCount;
    h->P.nsplit = 1;
    h->P.seed0 = 97; h->P.seed1 = 33;
    void *c = nullptr, *e = nullptr;
    if (cudaMalloc(&c, sizeof(Counters)) != cudaSuccess || cudaMalloc(&e, sizeof(double)) != cudaSuccess) {
        delete h;
        return 6;
    }
    cudaMemset(c, 0, sizeof(Counters));
    cudaMemset(e, 0, sizeof(double));
    h->P.counters = static_cast<Counters *>(c);
    h->P.ensrc = static_cast<double *>(e);
    *out = h;
    return 0;
}

void omc_gpu_destroy(omc_gpu_handle h) {
    if (!h) return;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    h->media_bufs.clear(); h->geom_bufs.clear(); h->source_bufs.clear(); free_pool(h->wave_bufs);
    drop_graph(h);
    if (h->comm) { nccl_api().CommDestroy(h->comm); h->comm = nullptr; }
    if (h->side) {
        cudaStreamSynchronize(h->side); cudaStreamDestroy(h->side);
        for (int i = 0; i < 2; i++) { cudaEventDestroy(h->ev_done[i]); cudaEventDestroy(h->ev_free[i]); }
    }
    cudaFree(h->ctl);
    if (h->ctl_host) cudaFreeHost(h->ctl_host);
    for (int i = 0; i < 2; i++) {
        if (h->ctl_slot[i]) cudaFreeHost(h->ctl_slot[i]);
        if (h->ev_stat[i]) cudaEventDestroy(h->ev_stat[i]);
    }
    cudaFree(h->P.endep); cudaFree(h->P.endep32); cudaFree(h->accum); cudaFree(h->accum2);
    cudaFree(h->P.counters); cudaFree(h->P.ensrc); cudaFree(h->stack); cudaFree(h->records);
    cudaFree(h->res_dens); cudaFree(h->res_dose); cudaFree(h->res_unc);
    cudaFree(h->fmt_tab); cudaFree(h->fmt_nfb_dev);
    if (h->fmt_nfb_host) cudaFreeHost(h->fmt_nfb_host);
    for (int i = 0; i < 2; i++) {
        cudaFree(h->fmt_dev[i]); cudaFree(h->fmt_fb_dev[i]);
        if (h->fmt_host[i]) cudaFreeHost(h->fmt_host[i]);
        if (h->fmt_fb_host[i]) cudaFreeHost(h->fmt_fb_host[i]);
        if (h->fmt_ev[i]) cudaEventDestroy(h->fmt_ev[i]);
    }
    cudaFree(h->mb_grid); cudaFree(h->mb_dmax); cudaFree(h->mb_nnz); cudaFree(h->mb_jc); cudaFree(h->mb_ir); cudaFree(h->mb_val);
    if (h->stream2) {
        cudaStreamDestroy(h->stream2); cudaStreamDestroy(h->stream3);
        cudaEventDestroy(h->ev_fork); cudaEventDestroy(h->ev_join); cudaEventDestroy(h->ev_fork3); cudaEventDestroy(h->ev_join3);
    }
    cudaStreamDestroy(h->stream);
    delete h;
}

const char *omc_gpu_last_error(omc_gpu_handle h) { return h ? h->err.c_str() : "null handle"; }

int omc_gpu_set_media(omc_gpu_handle h, const omc_media_tables *t) {
    if (!h || !t) return 2;
    if (t->nmed < 1 || t->nmed > OMC_MXMED) return fail(h, "nmed out of range (MXMED = 9, src/ompmc.h:356)");
    CK(cudaSetDevice(h->device));
    h->media_bufs.begin();
    const int nmed = t->nmed;
    DevProblem &P = h->P;
    P.nmed = nmed;
    // per-medium scalars
    std::vector<MedRec> med(nmed);
    for (int m = 0; m < nmed; m++) {
        MedRec &r = med[m];
        memset(&r, 0, sizeof r);
        r.ge1 = t->ge1[m]; r.ge0 = t->ge0[m]; r.eke1 = t->eke1[m]; r.eke0 = t->eke0[m];
        r.xcc = t->xcc[m]; r.blcc = t->blcc[m]; r.esig_e = t->esig_e[m]; r.psig_e = t->psig_e[m];
        r.te = t->pegs_te[m]; r.thmoll = t->pegs_thmoll[m]; r.ap = t->pegs_ap[m];
        r.delcm = t->delcm[m]; r.zbrang = t->zbrang[m]; r.bpar0 = t->bpar0[m]; r.bpar1 = t->bpar1[m];
        const double *dl[6] = {t->dl1, t->dl2, t->dl3, t->dl4, t->dl5, t->dl6};
        for (int a = 0; a < 6; a++)
            for (int k = 0; k < 8; k++) r.dl[a][k] = dl[a][m * 8 + k];
        r.sig_ismonotone[0] = t->sig_ismonotone[0 * nmed + m];
        r.sig_ismonotone[1] = t->sig_ismonotone[1 * nmed + m];
    }
    if (upload(h, h->media_bufs, med.data(), med.size(), &P.med)) return 1;
    h->med_host = med;
    h->med_dirty = true;
    // photon bins
    std::vector<PhotBin> pb((size_t)nmed * MXGE);
    for (size_t i = 0; i < pb.size(); i++) {
        PhotBin &b = pb[i];
        b.gmfp1 = t->gmfp1[i]; b.gmfp0 = t->gmfp0[i]; b.cohe1 = t->cohe1[i]; b.cohe0 = t->cohe0[i];
        b.gbr11 = t->gbr11[i]; b.gbr10 = t->gbr10[i]; b.gbr21 = t->gbr21[i]; b.gbr20 = t->gbr20[i];
        b.pmax1 = t->ray_pmax1[i]; b.pmax0 = t->ray_pmax0[i];
    }
    if (upload(h, h->media_bufs, pb.data(), pb.size(), &P.phot)) return 1;
    // electron bins, [qel][imed][lelke]
    std::vector<ElecBin> eb((size_t)2 * nmed * MXEKE);
    for (int q = 0; q < 2; q++)
        for (size_t i = 0; i < (size_t)nmed * MXEKE; i++) {
            ElecBin &b = eb[(size_t)q * nmed * MXEKE + i];
            memset(&b, 0, sizeof b);
            if (q == 0) {
                b.sig1 = t->esig1[i]; b.sig0 = t->esig0[i]; b.dedx1 = t->ededx1[i]; b.dedx0 = t->ededx0[i];
                b.eta1 = t->etae_ms1[i]; b.eta0 = t->etae_ms0[i]; b.q1c1 = t->q1ce_ms1[i]; b.q1c0 = t->q1ce_ms0[i];
                b.q2c1 = t->q2ce_ms1[i]; b.q2c0 = t->q2ce_ms0[i]; b.bra1 = t->ebr11[i]; b.bra0 = t->ebr10[i];
            } else {
                b.sig1 = t->psig1[i]; b.sig0 = t->psig0[i]; b.dedx1 = t->pdedx1[i]; b.dedx0 = t->pdedx0[i];
                b.eta1 = t->etap_ms1[i]; b.eta0 = t->etap_ms0[i]; b.q1c1 = t->q1cp_ms1[i]; b.q1c0 = t->q1cp_ms0[i];
                b.q2c1 = t->q2cp_ms1[i]; b.q2c0 = t->q2cp_ms0[i]; b.bra1 = t->pbr11[i]; b.bra0 = t->pbr10[i];
                b.brb1 = t->pbr21[i]; b.brb0 = t->pbr20[i];
            }
            b.tmxs1 = t->tmxs1[i]; b.tmxs0 = t->tmxs0[i]; b.blcce1 = t->blcce1[i]; b.blcce0 = t->blcce0[i];
            b.range_ep = t->range_ep[(size_t)q * nmed * MXEKE + i];
            b.e_array = t->e_array[i];
        }
    if (upload(h, h->media_bufs, eb.data(), eb.size(), &P.ebin)) return 1;
    // Rayleigh form-factor tables (only medium 0 is ever indexed by the reference, Q3; all are uploaded)
    if (upload(h, h->media_bufs, t->ray_xgrid, (size_t)nmed * OMC_MXRAYFF, &P.ray_xgrid)) return 1;
    if (upload(h, h->media_bufs, t->ray_fcum, (size_t)nmed * OMC_MXRAYFF, &P.ray_fcum)) return 1;
    if (upload(h, h->media_bufs, t->ray_b_array, (size_t)nmed * OMC_MXRAYFF, &P.ray_b)) return 1;
    if (upload(h, h->media_bufs, t->ray_c_array, (size_t)nmed * OMC_MXRAYFF, &P.ray_c)) return 1;
    if (upload(h, h->media_bufs, t->ray_i_array, (size_t)nmed * OMC_MXRAYFF, &P.ray_i)) return 1;
    // spin
    P.b2spin_min = t->b2spin_min; P.dbeta2i = t->dbeta2i; P.espml = t->espml; P.dleneri = t->dleneri; P.dqq1i = t->dqq1i;
    if (upload(h, h->media_bufs, t->spin_rej, (size_t)nmed * 2 * OMC_SPIN_NE * OMC_SPIN_NQ * OMC_SPIN_NU, &P.spin_rej)) return 1;
    // mscat
    const size_t nms = (size_t)OMC_MS_NL * OMC_MS_NQ * OMC_MS_NU;
    std::vector<MsEntry> ms(nms);
    for (size_t i = 0; i < nms; i++) {
        ms[i].ums = t->ums[i]; ms[i].fms = t->fms[i]; ms[i].wms = t->wms[i]; ms[i].ims = t->ims[i]; ms[i].pad = 0;
    }
    if (upload(h, h->media_bufs, ms.data(), nms, &P.ms)) return 1;
    {   // fp32 copies used by the wavefront kernels' single-precision angle samplers
        std::vector<MsEntryF> msf(nms);
        for (size_t i = 0; i < nms; i++) {
            msf[i].ums = (float)t->ums[i]; msf[i].wms = (float)t->wms[i]; msf[i].ims = t->ims[i]; msf[i].fms = (float)t->fms[i];
        }
        if (upload(h, h->media_bufs, msf.data(), nms, &P.ms_f)) return 1;
        const size_t nsp = (size_t)nmed * 2 * OMC_SPIN_NE * OMC_SPIN_NQ * OMC_SPIN_NU;
        std::vector<float> spf(nsp);
        for (size_t i = 0; i < nsp; i++) spf[i] = (float)t->spin_rej[i];
        if (upload(h, h->media_bufs, spf.data(), nsp, &P.spin_rej_f)) return 1;
    }
    P.dllambi = t->dllambi; P.dqmsi = t->dqmsi;
    h->have_media = true;
    return 0;
}

int omc_gpu_set_geometry(omc_gpu_handle h, const omc_geometry *g) {
    if (!h || !g) return 2;
    if (g->isize < 1 || g->jsize < 1 || g->ksize < 1) return fail(h, "empty voxel grid");
    const long long nvox = (long long)g->isize * g->jsize * g->ksize;
    if (nvox + 1 > 2147483647LL) return fail(h, "region index must fit the reference's 32-bit int");
    CK(cudaSetDevice(h->device));
    h->geom_bufs.begin();
    DevProblem &P = h->P;
    P.isize = g->isize; P.jsize = g->jsize; P.ksize = g->ksize;
    P.ijmax = g->isize * g->jsize;
    P.nreg = (int)(nvox + 1);
    if (upload(h, h->geom_bufs, g->xbounds, (size_t)g->isize + 1, &P.xb)) return 1;
    if (upload(h, h->geom_bufs, g->ybounds, (size_t)g->jsize + 1, &P.yb)) return 1;
    if (upload(h, h->geom_bufs, g->zbounds, (size_t)g->ksize + 1, &P.zb)) return 1;
    std::vector<RegionRec> reg((size_t)P.nreg);
    for (int i = 0; i < P.nreg; i++) {
        reg[i].rhof = g->rhof[i]; reg[i].ecut = g->ecut[i]; reg[i].pcut = g->pcut[i]; reg[i].med = g->med[i]; reg[i].pad = 0;
        if (g->med[i] < -1 || g->med[i] >= OMC_MXMED) return fail(h, "region medium index out of range");
    }
    if (upload(h, h->geom_bufs, reg.data(), reg.size(), &P.reg)) return 1;
    // compact 8-byte records for the wavefront kernels when the cut-offs depend on the medium only
    {
        bool uniform = true;
        bool seen[OMC_MXMED] = {false};
        for (int i = 1; i < P.nreg && uniform; i++) {
            const int m = g->med[i];
            if (m < 0) continue;
            if (!seen[m]) { seen[m] = true; h->cut_e[m] = g->ecut[i]; h->cut_p[m] = g->pcut[i]; }
            else if (h->cut_e[m] != g->ecut[i] || h->cut_p[m] != g->pcut[i]) uniform = false;
        }
        // largest density ratio per medium (as the kernels will read it): majorant of the Woodcock photon flight
        for (int m = 0; m < OMC_MXMED; m++) h->rho_max[m] = 0.0;
        for (int i = 1; i < P.nreg; i++) {
            const int m = g->med[i];
            if (m < 0) continue;
            const double r = g->rhof[i], rf = (double)(float)g->rhof[i];
            if (r > h->rho_max[m]) h->rho_max[m] = r;
            if (rf > h->rho_max[m]) h->rho_max[m] = rf;
        }
        // uniformly spaced axes: voxel lookup by division (find_bin)
        {
            const double *b[3] = {g->xbounds, g->ybounds, g->zbounds};
            const int nb[3] = {g->isize, g->jsize, g->ksize};
            double inv[3]; int uni[3];
            for (int a = 0; a < 3; a++) {
                const double d = (b[a][nb[a]] - b[a][0]) / nb[a];
                uni[a] = d > 0.0;
                for (int i = 0; i < nb[a] && uni[a]; i++)
                    if (fabs((b[a][i + 1] - b[a][i]) - d) > 1.0e-9 * d) uni[a] = 0;
                inv[a] = d > 0.0 ? 1.0 / d : 0.0;
            }
            P.inv_dx = inv[0]; P.inv_dy = inv[1]; P.inv_dz = inv[2];
            P.uniform_x = uni[0]; P.uniform_y = uni[1]; P.uniform_z = uni[2];
        }
        P.reg8 = nullptr;
        h->cuts_uniform = uniform;
        if (uniform) {
            struct R8 { float rhof; int med; };
            std::vector<R8> r8((size_t)P.nreg);
            for (int i = 0; i < P.nreg; i++) { r8[i].rhof = (float)g->rhof[i]; r8[i].med = g->med[i]; }
            const R8 *d = nullptr;
            if (upload(h, h->geom_bufs, r8.data(), r8.size(), &d)) return 1;
            P.reg8 = d;
        }
        h->med_dirty = true;
    }
    if (h->tally_nreg != P.nreg) {
        cudaFree(P.endep); cudaFree(P.endep32); cudaFree(h->accum); cudaFree(h->accum2);
        P.endep = nullptr; P.endep32 = nullptr; h->accum = h->accum2 = nullptr;
        CK(cudaMalloc((void **)&P.endep, (size_t)2 * P.nreg * sizeof(double)));      // two batch grids (pipelining), grid g at g * nreg
        CK(cudaMalloc((void **)&P.endep32, (size_t)2 * P.nreg * sizeof(float)));
        CK(cudaMalloc((void **)&h->accum, (size_t)P.nreg * sizeof(double)));
        CK(cudaMalloc((void **)&h->accum2, (size_t)P.nreg * sizeof(double)));
        h->tally_nreg = P.nreg;
    }
    h->have_geom = true;
    return omc_gpu_reset_tallies(h, 0);
}

int omc_gpu_set_source_dosxyz(omc_gpu_handle h, const omc_source_dosxyz *s) {
    if (!h || !s) return 2;
    if (s->charge < -1 || s->charge > 1) return fail(h, "Particle kind not recognized.");          // omc_dosxyz.c:602-605
    if (s->ssd < 0) return fail(h, "SSD must be greater than zero.");                              // omc_dosxyz.c:614-617
    CK(cudaSetDevice(h->device));
    h->source_bufs.begin();
    SourceDosxyz &S = h->P.src;
    S.spectrum = s->spectrum; S.charge = s->charge; S.energy = s->energy; S.deltak = s->deltak;
    S.ssd = s->ssd; S.xinl = s->xinl; S.yinl = s->yinl; S.xsize = s->xsize; S.ysize = s->ysize;
    S.ixinl = s->ixinl; S.iyinl = s->iyinl;
    S.cdfinv1 = S.cdfinv2 = nullptr;
    if (s->spectrum) {
        const size_t n = (size_t)s->deltak;
        if (n < 1 || !s->cdfinv1 || !s->cdfinv2) return fail(h, "spectrum source without inverse-CDF tables");
        if (upload(h, h->source_bufs, s->cdfinv1, n, &S.cdfinv1)) return 1;
        if (upload(h, h->source_bufs, s->cdfinv2, n, &S.cdfinv2)) return 1;
    }
    h->have_source = true;
    h->source_kind = 0;
    return 0;
}

int omc_gpu_set_source_matrad(omc_gpu_handle h, const omc_source_matrad *s) {
    if (!h || !s) return 2;
    if (s->charge < -1 || s->charge > 1) return fail(h, "Particle kind not recognized.");
    if (s->nbixels < 1 || s->nbeams < 1) return fail(h, "beamlet source without beamlets");
    CK(cudaSetDevice(h->device));
    h->source_bufs.begin();
    SourceDosxyz &S = h->P.src;
    memset(&S, 0, sizeof S);
    S.spectrum = s->spectrum; S.charge = s->charge; S.energy = s->energy; S.deltak = s->deltak;
    if (s->spectrum) {
        const size_t n = (size_t)s->deltak;
        if (n < 1 || !s->cdfinv1 || !s->cdfinv2) return fail(h, "spectrum source without inverse-CDF tables");
        if (upload(h, h->source_bufs, s->cdfinv1, n, &S.cdfinv1)) return 1;
        if (upload(h, h->source_bufs, s->cdfinv2, n, &S.cdfinv2)) return 1;
    }
    SourceMatrad &M = h->P.msrc;
    M.nbixels = s->nbixels; M.nbeams = s->nbeams;
    for (int i = 0; i < s->nbixels; i++)
        if (s->ibeam[i] < 0 || s->ibeam[i] >= s->nbeams) return fail(h, "beamlet refers to a beam that does not exist");
    const size_t nb = (size_t)s->nbixels, nm = (size_t)s->nbeams;
    if (upload(h, h->source_bufs, s->ibeam, nb, &M.ibeam)) return 1;
    if (upload(h, h->source_bufs, s->xsource, nm, &M.xsource) || upload(h, h->source_bufs, s->ysource, nm, &M.ysource) ||
        upload(h, h->source_bufs, s->zsource, nm, &M.zsource)) return 1;
    if (upload(h, h->source_bufs, s->xcorner, nb, &M.xcorner) || upload(h, h->source_bufs, s->ycorner, nb, &M.ycorner) ||
        upload(h, h->source_bufs, s->zcorner, nb, &M.zcorner)) return 1;
    if (upload(h, h->source_bufs, s->xside1, nb, &M.xside1) || upload(h, h->source_bufs, s->yside1, nb, &M.yside1) ||
        upload(h, h->source_bufs, s->zside1, nb, &M.zside1)) return 1;
    if (upload(h, h->source_bufs, s->xside2, nb, &M.xside2) || upload(h, h->source_bufs, s->yside2, nb, &M.yside2) ||
        upload(h, h->source_bufs, s->zside2, nb, &M.zside2)) return 1;
    h->have_source = true;
    h->source_kind = 1;
    return 0;
}

int omc_gpu_set_vrt(omc_gpu_handle h, int nsplit) {
    if (!h) return 2;
    if (nsplit < 1) nsplit = 1;   // "Photon splitting disabled" (src/ompmc.c:5973-5978); nsplit <= 0 would divide by zero in photon()
    h->P.nsplit = nsplit;
    return 0;
}

int omc_gpu_set_seed(omc_gpu_handle h, int ixx, int jxx) {
    if (!h) return 2;
    h->P.seed0 = (uint32_t)ixx; h->P.seed1 = (uint32_t)jxx;
    return 0;
}

int omc_gpu_set_option(omc_gpu_handle h, const char *key, long long value) {
    if (!h || !key) return 2;
    std::string k(key);
    if (k == "kernel") h->kernel = (int)value;
    else if (k == "threads_per_block") h->threads_per_block = (int)value;
    else if (k == "stack_depth") h->stack_depth = (int)value;
    else if (k == "max_blocks") h->max_blocks = (int)value;
    else if (k == "record_histories") h->record = (int)value;
    else if (k == "pool_size") h->pool_target = (unsigned)value;
    else if (k == "trace") h->trace = (int)value;
    else if (k == "use_graph") h->use_graph = (int)value;
    else if (k == "overlap") h->overlap = (int)value;
    else if (k == "drain_threshold") h->drain_threshold = (unsigned)value;
    else if (k == "pool_cap") h->pool_cap_opt = (unsigned)value;
    else if (k == "max_cross") h->max_cross = (int)value;
    else if (k == "photon_tracking") h->photon_tracking = (int)value;
    else if (k == "max_virtual") h->max_virtual = (int)value;
    else if (k == "check_every") h->check_every = (int)value;
    else if (k == "lookahead") h->lookahead = (int)value;
    else return fail(h, "unknown option");
    return 0;
}

int omc_gpu_run_histories(omc_gpu_handle h, long long first, long long nhist, int ibeamlet) {
    if (!h) return 2;
    if (!h->have_media || !h->have_geom || !h->have_source) return fail(h, "media, geometry and source must be set first");
    if (h->source_kind == 1) {
        if (ibeamlet < 0 || ibeamlet >= h->P.msrc.nbixels) return fail(h, "beamlet index out of range for the matRad source");
    } else {
        ibeamlet = -1;                  // omc_dosxyz source: the argument is ignored, as the reference has none
    }
    if (nhist < 0) return fail(h, "negative history count");
    if (nhist == 0 && !h->pipeline_next) return 0;   // (a pipelined start registers even an empty batch: its grid is still owed)
    CK(cudaSetDevice(h->device));
    DevProblem &P = h->P;
    if (h->med_dirty) {                 // per-medium cut-offs (from the geometry) into the per-medium records
        for (size_t m = 0; m < h->med_host.size(); m++) {
            h->med_host[m].ecut = h->cut_e[m]; h->med_host[m].pcut = h->cut_p[m]; h->med_host[m].rhomax = h->rho_max[m];
        }
        CK(cudaMemcpyAsync((void *)P.med, h->med_host.data(), h->med_host.size() * sizeof(MedRec), cudaMemcpyHostToDevice, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        h->med_dirty = false;
    }
    if (h->record) {
        if (h->records_cap < nhist) {
            cudaFree(h->records);
            h->records = nullptr;
            CK(cudaMalloc((void **)&h->records, (size_t)nhist * sizeof(omc_history_record)));
            h->records_cap = nhist;
        }
        P.records = h->records;
    } else {
        P.records = nullptr;
    }
    h->last_nhist = nhist;
    if (h->kernel == OMC_KERNEL_LOCKSTEP) {
        {                                                      // the lock-step kernel always scores into batch grid 0
            int rc = flush_all(h);
            if (rc) return rc;
            bool has0 = false;
            for (int d : h->done_q) has0 |= (d == 0);
            if (!has0) h->done_q.push_back(0);
            h->auto_acc[0] = h->pipeline_auto;
        }
        const int tpb = h->threads_per_block;
        int blocks_cap = h->max_blocks > 0 ? h->max_blocks : h->sm_count * lockstep_blocks_per_sm(tpb);
        long long want = (nhist + tpb - 1) / tpb;
        const int blocks = (int)(want < blocks_cap ? want : blocks_cap);
        const int depth = h->stack_depth > 0 ? h->stack_depth : (P.nsplit <= 1 ? 32 : 64 + 24 * P.nsplit);
        const size_t need = (size_t)depth * blocks * tpb * sizeof(Part);
        if (need > h->stack_bytes) {
            cudaFree(h->stack);
            h->stack = nullptr; h->stack_bytes = 0;
            CK(cudaMalloc((void **)&h->stack, need));
            h->stack_bytes = need;
        }
        if (h->grid_wait[0]) { CK(cudaStreamWaitEvent(h->stream, h->ev_free[0], 0)); h->grid_wait[0] = false; }
        if (nhist > 0) {
            launch_lockstep(P, h->stack, depth, blocks, tpb, first, nhist, ibeamlet, h->inject, h->stream);
            h->launches += 1;
            CK(cudaGetLastError());
        }
    } else if (h->kernel == OMC_KERNEL_WAVEFRONT) {
        if (P.nsplit > 255) return fail(h, "wavefront kernels support nsplit <= 255; use the lock-step kernel beyond");
        if (h->record) return fail(h, "per-history records are a lock-step kernel feature");
        int rc = run_wavefront(h, first, nhist, ibeamlet, h->pipeline_next, h->pipeline_auto);
        if (rc) return rc;
    } else {
        return fail(h, "unknown kernel");
    }
    return 0;
}

int omc_gpu_accum_batch(omc_gpu_handle h) {
    if (!h || !h->have_geom) return 2;
    CK(cudaSetDevice(h->device));
    if (h->done_q.empty() && h->run_grid >= 0) {               // the batch that was started is still in flight: complete it
        int rc = wave_run(h, false, 0, 0, -1, -1);
        if (rc) return rc;
    }
    if (h->done_q.empty()) {                                   // nothing was run: accumEndep() of an empty batch grid
        launch_accum(h->P.endep, h->accum, h->accum2, h->P.nreg, h->stream);
        h->launches += 1;
        CK(cudaGetLastError());
        return 0;
    }
    return accum_done(h, false);
}

// this rank's contiguous slice of the batch's history ids (the first nhist % world ranks get one more)
static void shard_range(omc_gpu_handle h, long long &first, long long &nhist) {
    if (h->world <= 1) return;
    omc_gpu_shard_range(first, nhist, h->rank, h->world, &first, &nhist);
}

int omc_gpu_start_batch(omc_gpu_handle h, long long first, long long nhist, int ibeamlet) {
    if (!h) return 2;
    shard_range(h, first, nhist);
    const bool wave = (h->kernel == OMC_KERNEL_WAVEFRONT) && !h->record;
    h->pipeline_next = wave; h->pipeline_auto = false;
    int rc = omc_gpu_run_histories(h, first, nhist, ibeamlet);
    h->pipeline_next = false;
    return rc;
}

int omc_gpu_finish_batches(omc_gpu_handle h) {
    if (!h || !h->have_geom) return 2;
    CK(cudaSetDevice(h->device));
    if (h->run_grid >= 0) return wave_run(h, false, 0, 0, -1, -1);
    return 0;
}

int omc_gpu_completed_batches(omc_gpu_handle h) { return h ? (int)h->done_q.size() : -1; }

int omc_gpu_run_batch(omc_gpu_handle h, long long first, long long nhist, int ibeamlet) {
    if (!h) return 2;
    // pipelined with the wavefront kernels: on return the histories are all started and every EARLIER batch is
    // accumulated; this batch is completed and accumulated by the next call or by whatever reads results
    const bool wave = (h->kernel == OMC_KERNEL_WAVEFRONT) && !h->record;
    shard_range(h, first, nhist);
    h->pipeline_next = wave; h->pipeline_auto = true;
    int rc = omc_gpu_run_histories(h, first, nhist, ibeamlet);
    h->pipeline_next = false; h->pipeline_auto = false;
    if (rc) return rc;
    if (!wave) return omc_gpu_accum_batch(h);
    return accum_done(h, true);
}

int omc_gpu_synchronize(omc_gpu_handle h) {
    if (!h) return 2;
    CK(cudaSetDevice(h->device));
    if (h->have_geom) {
        int rc = flush_all(h);
        if (rc) return rc;
    }
    CK(cudaStreamSynchronize(h->stream));
    if (h->side) CK(cudaStreamSynchronize(h->side));
    Counters c;
    CK(cudaMemcpy(&c, h->P.counters, sizeof c, cudaMemcpyDeviceToHost));
    if (c.errors) {
        // the reference aborts here: "Stack overflow with np = %d. Increase MXSTACK!" (src/ompmc.c:1937-1940)
        h->err = "particle stack / queue overflow on the device: increase option stack_depth (or pool_size)";
        return 7;
    }
    return 0;
}

int omc_gpu_get_tallies(omc_gpu_handle h, double *accum, double *accum2, double *ensrc) {
    if (!h || !h->have_geom) return 2;
    int rc = omc_gpu_synchronize(h);
    if (rc) return rc;
    const size_t n = (size_t)h->P.nreg * sizeof(double);
    if (accum) CK(cudaMemcpy(accum, h->accum, n, cudaMemcpyDeviceToHost));
    if (accum2) CK(cudaMemcpy(accum2, h->accum2, n, cudaMemcpyDeviceToHost));
    if (ensrc) CK(cudaMemcpy(ensrc, h->P.ensrc, sizeof(double), cudaMemcpyDeviceToHost));
    return 0;
}

// accumulateResults() on the device: leaves dose / relative uncertainty in h->res_dose / h->res_unc (indexed like the tallies)
static int results_on_device(omc_gpu_handle h, int iout, int nhist, int nbatch, const double *med_densities) {
    // (nbatch == 1 is accepted as the reference accepts it: its division by nbatch - 1 leaves NaN uncertainties in the file)
    if (nbatch < 1) return fail(h, "accumulateResults: batch count must be positive");
    if (nhist < 1) return fail(h, "accumulateResults: history count must be positive");
    int rc = omc_gpu_synchronize(h);
    if (rc) return rc;
    const size_t nreg = (size_t)h->P.nreg, nvox = nreg - 1;
    if (h->res_nreg != h->P.nreg) {
        cudaFree(h->res_dens); cudaFree(h->res_dose); cudaFree(h->res_unc);
        h->res_dens = h->res_dose = h->res_unc = nullptr; h->res_nreg = -1;
        CK(cudaMalloc((void **)&h->res_dens, (nvox ? nvox : 1) * sizeof(double)));
        CK(cudaMalloc((void **)&h->res_dose, nreg * sizeof(double)));
        CK(cudaMalloc((void **)&h->res_unc, nreg * sizeof(double)));
        h->res_nreg = h->P.nreg;
    }
    CK(cudaMemcpyAsync(h->res_dens, med_densities, nvox * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    launch_results(h->P, h->accum, h->accum2, h->res_dens, iout, nhist, nbatch, h->res_dose, h->res_unc, h->stream);
    h->launches += 1;
    CK(cudaGetLastError());
    return 0;
}

int omc_gpu_accumulate_results(omc_gpu_handle h, int iout, int nhist, int nbatch, const double *med_densities, double *dose,
                               double *unc) {
    if (!h || !h->have_geom || !med_densities || !dose || !unc) return 2;
    int rc = results_on_device(h, iout, nhist, nbatch, med_densities);
    if (rc) return rc;
    const size_t nreg = (size_t)h->P.nreg;
    CK(cudaMemcpyAsync(dose, h->res_dose, nreg * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(unc, h->res_unc, nreg * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

// ---- outputResults() (omc_dosxyz.c:801-886): accumulateResults() + the .3ddose text, both on the device ------------------------
static const long long FMT_CHUNK = 4ll << 20;          // values per chunk: 52 MiB of "%e " text
static const unsigned FMT_FB_CAP = 1u << 16;           // host-formatted values per chunk before the whole chunk goes to snprintf

static int format_setup(omc_gpu_handle h) {
    if (h->fmt_ready) return 0;
    std::vector<Pow10> tab(kPow10N);
    build_pow10_table(tab.data());
    if (!h->fmt_tab) CK(cudaMalloc((void **)&h->fmt_tab, sizeof(Pow10) * kPow10N));
    CK(cudaMemcpyAsync(h->fmt_tab, tab.data(), sizeof(Pow10) * kPow10N, cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    // (a failed allocation leaves fmt_ready false: the next call retries only what is still missing)
    if (!h->fmt_nfb_dev) CK(cudaMalloc((void **)&h->fmt_nfb_dev, 2 * sizeof(unsigned)));
    if (!h->fmt_nfb_host) CK(cudaMallocHost((void **)&h->fmt_nfb_host, 2 * sizeof(unsigned)));
    for (int i = 0; i < 2; i++) {
        if (!h->fmt_dev[i]) CK(cudaMalloc((void **)&h->fmt_dev[i], (size_t)FMT_CHUNK * kFmtEWidth));
        if (!h->fmt_host[i]) CK(cudaMallocHost((void **)&h->fmt_host[i], (size_t)FMT_CHUNK * kFmtEWidth));
        if (!h->fmt_fb_dev[i]) CK(cudaMalloc((void **)&h->fmt_fb_dev[i], sizeof(FormatFallback) * FMT_FB_CAP));
        if (!h->fmt_fb_host[i]) CK(cudaMallocHost((void **)&h->fmt_fb_host[i], sizeof(FormatFallback) * FMT_FB_CAP));
        if (!h->fmt_ev[i]) CK(cudaEventCreateWithFlags(&h->fmt_ev[i], cudaEventDisableTiming));
    }
    h->fmt_ready = true;
    return 0;
}

// one block of the file: n values of src (device) as text; chunk c+1 is formatted and copied while chunk c is written
static int format_block(omc_gpu_handle h, FILE *fp, const double *src, long long n, int mode) {
    const int W = mode == 0 ? kFmtEWidth : kFmtFWidth;
    const char *spec = mode == 0 ? "%e " : "%f ";
    const long long nchunk = (n + FMT_CHUNK - 1) / FMT_CHUNK;
    auto issue = [&](long long c) -> int {
        const int b = (int)(c & 1);
        const long long lo = c * FMT_CHUNK, cnt = std::min(FMT_CHUNK, n - lo);
        CK(cudaMemsetAsync(h->fmt_nfb_dev + b, 0, sizeof(unsigned), h->stream));
        launch_format(mode, src + lo, cnt, h->fmt_tab, h->fmt_dev[b], h->fmt_fb_dev[b], h->fmt_nfb_dev + b, FMT_FB_CAP, h->stream);
        h->launches += 1;
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(h->fmt_host[b], h->fmt_dev[b], (size_t)cnt * W, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaMemcpyAsync(h->fmt_nfb_host + b, h->fmt_nfb_dev + b, sizeof(unsigned), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaMemcpyAsync(h->fmt_fb_host[b], h->fmt_fb_dev[b], sizeof(FormatFallback) * FMT_FB_CAP, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaEventRecord(h->fmt_ev[b], h->stream));
        return 0;
    };
    if (nchunk > 0 && issue(0)) return 1;
    std::vector<double> raw;
    char one[512];
    for (long long c = 0; c < nchunk; c++) {
        const int b = (int)(c & 1);
        const long long lo = c * FMT_CHUNK, cnt = std::min(FMT_CHUNK, n - lo);
        CK(cudaEventSynchronize(h->fmt_ev[b]));
        if (c + 1 < nchunk && issue(c + 1)) return 1;          // the other buffer: formatted + copied while this one is written
        const unsigned nfb = h->fmt_nfb_host[b];
        const char *text = h->fmt_host[b];
        bool ok = true;
        if (nfb == 0) {
            ok = fwrite(text, 1, (size_t)cnt * W, fp) == (size_t)cnt * W;
        } else if (nfb <= FMT_FB_CAP) {                         // splice the host-formatted values in, in index order
            FormatFallback *fb = h->fmt_fb_host[b];
            std::sort(fb, fb + nfb, [](const FormatFallback &x, const FormatFallback &y) { return x.index < y.index; });
            long long at = 0;
            for (unsigned k = 0; k < nfb && ok; k++) {
                const long long i = (long long)fb[k].index;
                ok = fwrite(text + at * W, 1, (size_t)(i - at) * W, fp) == (size_t)(i - at) * W;
                const int len = snprintf(one, sizeof one, spec, fb[k].value);
                ok = ok && fwrite(one, 1, (size_t)len, fp) == (size_t)len;
                at = i + 1;
            }
            ok = ok && fwrite(text + at * W, 1, (size_t)(cnt - at) * W, fp) == (size_t)(cnt - at) * W;
        } else {                                                // (pathological grid, e.g. all NaN) the reference's loop for this chunk
            raw.resize((size_t)cnt);
            CK(cudaMemcpy(raw.data(), src + lo, (size_t)cnt * sizeof(double), cudaMemcpyDeviceToHost));
            for (long long i = 0; i < cnt && ok; i++) ok = fprintf(fp, spec, raw[(size_t)i]) > 0;
        }
        if (!ok) return fail(h, "omc_gpu_write_3ddose: write failed");
    }
    return fputc('\n', fp) == EOF ? fail(h, "omc_gpu_write_3ddose: write failed") : 0;
}

int omc_gpu_write_3ddose(omc_gpu_handle h, const char *path, int iout, int nhist, int nbatch, const double *med_densities) {
    if (!h || !path || !h->have_geom || !med_densities) return 2;
    CK(cudaSetDevice(h->device));
    int rc = results_on_device(h, iout, nhist, nbatch, med_densities);
    if (rc) return rc;
    if (format_setup(h)) return 1;
    const DevProblem &P = h->P;
    std::vector<double> xb(P.isize + 1), yb(P.jsize + 1), zb(P.ksize + 1);
    CK(cudaMemcpyAsync(xb.data(), P.xb, xb.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(yb.data(), P.yb, yb.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(zb.data(), P.zb, zb.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    FILE *fp = fopen(path, "w");
    if (!fp) { h->err = std::string("Unable to open file: ") + path; return 2; }
    static const size_t IOBUF = 8u << 20;
    std::vector<char> iobuf(IOBUF);
    setvbuf(fp, iobuf.data(), _IOFBF, IOBUF);
    fprintf(fp, "%5d%5d%5d\n", P.isize, P.jsize, P.ksize);
    for (double v : xb) fprintf(fp, "%f ", v);
    fprintf(fp, "\n");
    for (double v : yb) fprintf(fp, "%f ", v);
    fprintf(fp, "\n");
    for (double v : zb) fprintf(fp, "%f ", v);
    fprintf(fp, "\n");
    const long long nvox = (long long)P.nreg - 1;
    rc = format_block(h, fp, h->res_dose + 1, nvox, 0);
    if (!rc) rc = format_block(h, fp, h->res_unc + 1, nvox, 1);
    if (fclose(fp) != 0 && !rc) rc = fail(h, "omc_gpu_write_3ddose: write failed");
    return rc;
}

int omc_gpu_run_beamlets(omc_gpu_handle h, long long first, int nhist, int nbatch, int ib0, int nb, double rel_threshold,
                         const double *med_densities, long long *jc, long long *nnz_total) {
    if (!h || !jc || !nnz_total || !med_densities) return 2;
    if (!h->have_media || !h->have_geom || !h->have_source) return fail(h, "media, geometry and source must be set first");
    if (h->source_kind != 1) return fail(h, "omc_gpu_run_beamlets needs the matRad beamlet source (omc_gpu_set_source_matrad)");
    if (nb < 1 || ib0 < 0 || ib0 + nb > h->P.msrc.nbixels) return fail(h, "beamlet range out of bounds");
    if (nhist < 1 || nbatch < 1) return fail(h, "history / batch counts must be positive");
    if (h->kernel != OMC_KERNEL_WAVEFRONT) return fail(h, "omc_gpu_run_beamlets runs on the wavefront kernels (option kernel = 1)");
    if (h->P.nsplit > 255) return fail(h, "wavefront kernels support nsplit <= 255");
    CK(cudaSetDevice(h->device));
    int rc = flush_all(h);
    if (rc) return rc;
    DevProblem &P = h->P;
    const size_t nreg = (size_t)P.nreg, nvox = nreg - 1;
    const size_t need = (size_t)nb * nreg;
    if (need > h->mb_grid_elems) {
        cudaFree(h->mb_grid);
        h->mb_grid = nullptr; h->mb_grid_elems = 0;
        CK(cudaMalloc((void **)&h->mb_grid, need * sizeof(float)));
        h->mb_grid_elems = need;
    }
    if (nb > h->mb_meta_cap) {
        cudaFree(h->mb_dmax); cudaFree(h->mb_nnz); cudaFree(h->mb_jc);
        h->mb_dmax = nullptr; h->mb_nnz = nullptr; h->mb_jc = nullptr; h->mb_meta_cap = 0;
        CK(cudaMalloc((void **)&h->mb_dmax, (size_t)nb * sizeof(double)));
        CK(cudaMalloc((void **)&h->mb_nnz, (size_t)nb * sizeof(unsigned long long)));
        CK(cudaMalloc((void **)&h->mb_jc, ((size_t)nb + 1) * sizeof(long long)));
        h->mb_meta_cap = nb;
    }
    if (h->res_nreg != P.nreg) {                               // (density scratch shared with omc_gpu_accumulate_results)
        cudaFree(h->res_dens); cudaFree(h->res_dose); cudaFree(h->res_unc);
        h->res_dens = h->res_dose = h->res_unc = nullptr; h->res_nreg = -1;
        CK(cudaMalloc((void **)&h->res_dens, (nvox ? nvox : 1) * sizeof(double)));
        CK(cudaMalloc((void **)&h->res_dose, nreg * sizeof(double)));
        CK(cudaMalloc((void **)&h->res_unc, nreg * sizeof(double)));
        h->res_nreg = P.nreg;
    }
    CK(cudaMemcpyAsync(h->res_dens, med_densities, nvox * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemsetAsync(h->mb_grid, 0, need * sizeof(float), h->stream));
    if (h->med_dirty) {
        for (size_t m = 0; m < h->med_host.size(); m++) {
            h->med_host[m].ecut = h->cut_e[m]; h->med_host[m].pcut = h->cut_p[m]; h->med_host[m].rhomax = h->rho_max[m];
        }
        CK(cudaMemcpyAsync((void *)P.med, h->med_host.data(), h->med_host.size() * sizeof(MedRec), cudaMemcpyHostToDevice, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        h->med_dirty = false;
    }
    P.records = nullptr;
    // all beamlets of the group in one pass of the wavefront kernels
    h->mb_active = true;
    h->mb_first = (unsigned long long)first; h->mb_per = (unsigned)nhist; h->mb_n = (unsigned)nb; h->mb_ib0 = ib0;
    h->auto_acc[0] = false;
    rc = wave_run(h, true, first, (long long)nhist * nb, ib0, 0);
    if (!rc) rc = wave_run(h, false, 0, 0, -1, -1);
    h->mb_active = false;
    h->done_q.clear();                                          // (the two batch grids were not used)
    if (rc) return rc;
    // accumulateResults + threshold + column assembly, omc_matrad.c:1416-1477
    launch_mb_scan(h->mb_grid, (long long)nreg, nb, P, h->res_dens, nhist, nbatch, rel_threshold, h->mb_dmax, h->mb_nnz, 0, h->stream);
    launch_mb_scan(h->mb_grid, (long long)nreg, nb, P, h->res_dens, nhist, nbatch, rel_threshold, h->mb_dmax, h->mb_nnz, 1, h->stream);
    h->launches += 2;
    CK(cudaGetLastError());
    std::vector<unsigned long long> nnz((size_t)nb);
    CK(cudaMemcpyAsync(nnz.data(), h->mb_nnz, (size_t)nb * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    jc[0] = 0;
    for (int b = 0; b < nb; b++) jc[b + 1] = jc[b] + (long long)nnz[b];
    const long long total = jc[nb];
    if (total > h->mb_col_cap) {
        cudaFree(h->mb_ir); cudaFree(h->mb_val);
        h->mb_ir = nullptr; h->mb_val = nullptr; h->mb_col_cap = 0;
        CK(cudaMalloc((void **)&h->mb_ir, (size_t)(total ? total : 1) * sizeof(long long)));
        CK(cudaMalloc((void **)&h->mb_val, (size_t)(total ? total : 1) * sizeof(double)));
        h->mb_col_cap = total;
    }
    CK(cudaMemcpyAsync(h->mb_jc, jc, ((size_t)nb + 1) * sizeof(long long), cudaMemcpyHostToDevice, h->stream));
    if (total > 0) {
        launch_mb_fill(h->mb_grid, (long long)nreg, nb, P, h->res_dens, nhist, nbatch, rel_threshold, h->mb_dmax, h->mb_jc, h->mb_ir,
                       h->mb_val, h->stream);
        h->launches += 1;
        CK(cudaGetLastError());
    }
    CK(cudaStreamSynchronize(h->stream));
    h->mb_total = total;
    *nnz_total = total;
    return 0;
}

int omc_gpu_fetch_columns(omc_gpu_handle h, long long *ir, double *val) {
    if (!h || !ir || !val) return 2;
    CK(cudaSetDevice(h->device));
    if (h->mb_total > 0) {
        CK(cudaMemcpy(ir, h->mb_ir, (size_t)h->mb_total * sizeof(long long), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(val, h->mb_val, (size_t)h->mb_total * sizeof(double), cudaMemcpyDeviceToHost));
    }
    return 0;
}

int omc_gpu_get_batch_grid(omc_gpu_handle h, double *endep) {
    if (!h || !h->have_geom || !endep) return 2;
    int rc = omc_gpu_synchronize(h);
    if (rc) return rc;
    const int g = h->done_q.empty() ? 0 : h->done_q.front();   // the completed batch waiting for accumEndep(), else grid 0
    CK(cudaMemcpy(endep, h->P.endep + (size_t)g * h->P.nreg, (size_t)h->P.nreg * sizeof(double), cudaMemcpyDeviceToHost));
    return 0;
}

int omc_gpu_reset_tallies(omc_gpu_handle h, int which) {
    if (!h || !h->have_geom) return 2;
    CK(cudaSetDevice(h->device));
    const size_t n = (size_t)h->P.nreg;
    if (h->side) CK(cudaStreamSynchronize(h->side));
    if (which == 0) {                       // everything starts over: particles still in flight are dropped with their grids
        h->run_grid = -1; h->done_q.clear(); h->auto_acc[0] = h->auto_acc[1] = false;
        h->grid_wait[0] = h->grid_wait[1] = false;
        if (h->ctl_host) memset(h->ctl_host, 0, sizeof(WaveCtl));
    } else {                                // omc_matrad.c:1482 zeroes accum_endep between beamlets: settle what is owed first
        int rc = flush_all(h);
        if (rc) return rc;
    }
    CK(cudaMemsetAsync(h->accum, 0, n * sizeof(double), h->stream));
    if (which == 0) {
        CK(cudaMemsetAsync(h->accum2, 0, n * sizeof(double), h->stream));
        CK(cudaMemsetAsync(h->P.endep, 0, 2 * n * sizeof(double), h->stream));
        CK(cudaMemsetAsync(h->P.endep32, 0, 2 * n * sizeof(float), h->stream));
        CK(cudaMemsetAsync(h->P.ensrc, 0, sizeof(double), h->stream));
        CK(cudaMemsetAsync(h->P.counters, 0, sizeof(Counters), h->stream));
        h->launches = 0;
    }
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

int omc_gpu_device_ptrs(omc_gpu_handle h, void **endep, void **accum, void **accum2, long long *nreg) {
    if (!h || !h->have_geom) return 2;
    if (endep) *endep = h->P.endep + (size_t)(h->done_q.empty() ? 0 : h->done_q.front()) * h->P.nreg;
    if (accum) *accum = h->accum;
    if (accum2) *accum2 = h->accum2;
    if (nreg) *nreg = h->P.nreg;
    return 0;
}

// ---- multi-GPU: NCCL inside the library ------------------------------------------------------------------------
int omc_gpu_shard_range(long long first, long long nhist, int rank, int world, long long *lo, long long *count) {
    if (!lo || !count || world < 1 || rank < 0 || rank >= world || nhist < 0) return 2;
    const long long base = nhist / world, extra = nhist % world, r = rank;
    *lo = first + r * base + (r < extra ? r : extra);
    *count = base + (r < extra ? 1 : 0);
    return 0;
}

int omc_gpu_comm_unique_id(char *id128) {
    if (!id128) return 2;
    NcclApi &N = nccl_api();
    if (!N.ok) { fprintf(stderr, "ompmc_b200: %s\n", N.err.c_str()); return 8; }
    nccl_uid id;
    if (N.GetUniqueId(&id)) return 8;
    memcpy(id128, id.internal, 128);
    return 0;
}

int omc_gpu_comm_init(omc_gpu_handle h, int rank, int world, const char *id128) {
    if (!h || !id128) return 2;
    if (world < 1 || rank < 0 || rank >= world) return fail(h, "omc_gpu_comm_init: rank / world out of range");
    if (h->run_grid >= 0 || !h->done_q.empty()) return fail(h, "omc_gpu_comm_init: batches are in flight");
    CK(cudaSetDevice(h->device));
    if (h->comm) { nccl_api().CommDestroy(h->comm); h->comm = nullptr; }
    h->rank = rank; h->world = world;
    if (world == 1) return 0;
    NcclApi &N = nccl_api();
    if (!N.ok) { h->err = N.err; return 8; }
    nccl_uid id;
    memcpy(id.internal, id128, 128);
    const int e = N.CommInitRank(&h->comm, world, id, rank);
    if (e) { h->comm = nullptr; h->world = 1; h->rank = 0; h->err = std::string("ncclCommInitRank: ") + N.GetErrorString(e); return 8; }
    if (!h->side) {
        CK(cudaStreamCreateWithFlags(&h->side, cudaStreamNonBlocking));
        for (int i = 0; i < 2; i++) {
            CK(cudaEventCreateWithFlags(&h->ev_done[i], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&h->ev_free[i], cudaEventDisableTiming));
        }
    }
    return 0;
}

int omc_gpu_comm_rank(omc_gpu_handle h) { return h ? h->rank : -1; }
int omc_gpu_comm_size(omc_gpu_handle h) { return h ? h->world : -1; }

int omc_gpu_comm_sum(omc_gpu_handle h, double *values, int n) {
    if (!h || !values || n < 1) return 2;
    if (!h->comm) return 0;
    CK(cudaSetDevice(h->device));
    double *d = nullptr;
    CK(cudaMalloc((void **)&d, (size_t)n * sizeof(double)));
    cudaError_t ce = cudaMemcpyAsync(d, values, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, h->stream);
    int e = 0;
    if (ce == cudaSuccess) e = nccl_api().AllReduce(d, d, (size_t)n, NCCL_FLOAT64, NCCL_SUM, h->comm, h->stream);
    if (ce == cudaSuccess && !e) ce = cudaMemcpyAsync(values, d, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
    if (ce == cudaSuccess && !e) ce = cudaStreamSynchronize(h->stream);
    cudaFree(d);
    if (e) { h->err = std::string("ncclAllReduce: ") + nccl_api().GetErrorString(e); return 8; }
    if (ce != cudaSuccess) { h->err = std::string("omc_gpu_comm_sum: ") + cudaGetErrorString(ce); return 1; }
    return 0;
}

// Sparse columns of the beamlets this rank computed -> the complete matrix on every rank, in beamlet order (what the
// reference's loop builds column by column, omc_matrad.c:1416-1477).  jc[nb_total + 1] = global column starts (from the
// per-beamlet counts, summed over the ranks with omc_gpu_comm_sum); mine[b] != 0 marks this rank's beamlets, whose rows and
// values are passed concatenated in beamlet order.  One all-reduce of zero-padded arrays over NVLink (rows travel as exact
// doubles): the matrix is a few tens of MB, the exchange is noise next to the transport.
int omc_gpu_comm_gather_columns(omc_gpu_handle h, int nb_total, const long long *jc, const unsigned char *mine, const long long *ir_mine,
                                const double *val_mine, long long *ir_out, double *val_out) {
    if (!h || !jc || !mine || !ir_out || !val_out || nb_total < 1) return 2;
    const long long total = jc[nb_total];
    if (total <= 0) return 0;
    if (2 * total > 2147483647LL) return fail(h, "omc_gpu_comm_gather_columns: more than 2^30 non-zeros");
    std::vector<double> buf((size_t)2 * total, 0.0);           // [rows as doubles | values]
    long long at = 0;
    for (int b = 0; b < nb_total; b++) {
        if (!mine[b]) continue;
        const long long n = jc[b + 1] - jc[b];
        for (long long k = 0; k < n; k++) {
            buf[(size_t)(jc[b] + k)] = (double)ir_mine[at + k];
            buf[(size_t)(total + jc[b] + k)] = val_mine[at + k];
        }
        at += n;
    }
    const int rc = omc_gpu_comm_sum(h, buf.data(), (int)(2 * total));
    if (rc) return rc;
    for (long long k = 0; k < total; k++) { ir_out[k] = (long long)buf[(size_t)k]; val_out[k] = buf[(size_t)(total + k)]; }
    return 0;
}

int omc_gpu_abi_sizeof(int what) {
    switch (what) {
        case 0: return (int)sizeof(omc_media_tables);
        case 1: return (int)sizeof(omc_geometry);
        case 2: return (int)sizeof(omc_source_dosxyz);
        case 3: return (int)sizeof(omc_source_matrad);
        case 4: return (int)sizeof(omc_history_record);
        case 5: return (int)sizeof(omc_gpu_counters);
        default: return -1;
    }
}

void *omc_gpu_stream(omc_gpu_handle h) { return h ? (void *)h->stream : nullptr; }

int omc_gpu_get_counters(omc_gpu_handle h, omc_gpu_counters *out) {
    if (!h || !out) return 2;
    CK(cudaSetDevice(h->device));
    if (h->have_geom) {
        int rc = flush_all(h);
        if (rc) return rc;
    }
    CK(cudaStreamSynchronize(h->stream));
    static_assert(sizeof(Counters) == sizeof(omc_gpu_counters), "counter layouts must match");
    CK(cudaMemcpy(out, h->P.counters, sizeof(Counters), cudaMemcpyDeviceToHost));
    out->kernel_launches = h->launches;
    return 0;
}

int omc_gpu_get_history_records(omc_gpu_handle h, omc_history_record *out, long long n) {
    if (!h || !out) return 2;
    if (!h->records || n > h->last_nhist) return fail(h, "no history records (set option record_histories=1 before running)");
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaMemcpy(out, h->records, (size_t)n * sizeof(omc_history_record), cudaMemcpyDeviceToHost));
    return 0;
}

int omc_gpu_test_geometry(omc_gpu_handle h, int n, const double *xyzuvw, const int *ir, const double *ustep_in, int *idisc,
                          int *irnew, double *ustep_out, double *tperp) {
    if (!h || !h->have_geom) return 2;
    if (n <= 0) return 0;
    CK(cudaSetDevice(h->device));
    double *dq = nullptr, *dus = nullptr, *duo = nullptr, *dtp = nullptr;
    int *dir = nullptr, *did = nullptr, *dirn = nullptr;
    CK(cudaMalloc((void **)&dq, (size_t)6 * n * sizeof(double))); CK(cudaMalloc((void **)&dus, (size_t)n * sizeof(double)));
    CK(cudaMalloc((void **)&duo, (size_t)n * sizeof(double)));    CK(cudaMalloc((void **)&dtp, (size_t)n * sizeof(double)));
    CK(cudaMalloc((void **)&dir, (size_t)n * sizeof(int)));       CK(cudaMalloc((void **)&did, (size_t)n * sizeof(int)));
    CK(cudaMalloc((void **)&dirn, (size_t)n * sizeof(int)));
    CK(cudaMemcpyAsync(dq, xyzuvw, (size_t)6 * n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(dus, ustep_in, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(dir, ir, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    launch_test_geometry(h->P, n, dq, dir, dus, did, dirn, duo, dtp, h->stream);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaMemcpy(idisc, did, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(irnew, dirn, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(ustep_out, duo, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(tperp, dtp, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost));
    cudaFree(dq); cudaFree(dus); cudaFree(duo); cudaFree(dtp); cudaFree(dir); cudaFree(did); cudaFree(dirn);
    return 0;
}

int omc_gpu_test_particles(omc_gpu_handle h, int n, const int *iq, const double *e, const double *xyzuvw, const int *ir,
                           const double *wt, long long first_history, omc_history_record *records) {
    if (!h || n <= 0 || !records) return 2;
    if (!h->have_media || !h->have_geom || !h->have_source) return fail(h, "media, geometry and source must be set first");
    CK(cudaSetDevice(h->device));
    std::vector<Part> host((size_t)n);
    for (int i = 0; i < n; i++) {
        Part &p = host[i];
        p.x = xyzuvw[6 * i]; p.y = xyzuvw[6 * i + 1]; p.z = xyzuvw[6 * i + 2];
        p.u = xyzuvw[6 * i + 3]; p.v = xyzuvw[6 * i + 4]; p.w = xyzuvw[6 * i + 5];
        p.e = e[i]; p.wt = wt[i]; p.ir = ir[i]; p.iq = iq[i];
        if (ir[i] < 0 || ir[i] >= h->P.nreg) return fail(h, "region index out of range");
    }
    Part *dev = nullptr;
    CK(cudaMalloc((void **)&dev, (size_t)n * sizeof(Part)));
    CK(cudaMemcpyAsync(dev, host.data(), (size_t)n * sizeof(Part), cudaMemcpyHostToDevice, h->stream));
    const int kernel = h->kernel, record = h->record;
    h->kernel = OMC_KERNEL_LOCKSTEP; h->record = 1; h->inject = dev;
    int rc = omc_gpu_run_histories(h, first_history, n, -1);
    if (!rc) rc = omc_gpu_get_history_records(h, records, n);
    h->kernel = kernel; h->record = record; h->inject = nullptr;
    cudaFree(dev);
    return rc;
}

int omc_gpu_test_samplers(omc_gpu_handle h, int which, int n, const double *in, long long first_history, double *out) {
    if (!h || !in || !out || n <= 0) return 2;
    if (which < 0 || (which & 0xff) > OMC_SAMPLER_ESTEP || (which >> 8) > 2) return fail(h, "unknown sampler");
    if (!h->have_media || !h->have_geom) return fail(h, "media and geometry must be set first");
    CK(cudaSetDevice(h->device));
    if (h->med_dirty) {
        for (size_t m = 0; m < h->med_host.size(); m++) {
            h->med_host[m].ecut = h->cut_e[m]; h->med_host[m].pcut = h->cut_p[m]; h->med_host[m].rhomax = h->rho_max[m];
        }
        CK(cudaMemcpyAsync((void *)h->P.med, h->med_host.data(), h->med_host.size() * sizeof(MedRec), cudaMemcpyHostToDevice, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        h->med_dirty = false;
    }
    double *din = nullptr, *dout = nullptr;
    CK(cudaMalloc((void **)&din, (size_t)8 * n * sizeof(double)));
    CK(cudaMalloc((void **)&dout, (size_t)8 * n * sizeof(double)));
    // (on the context's stream: a cudaMemcpy on the legacy stream may return while its DMA from the staging buffer is still
    // running, and h->stream is a non-blocking stream that would not wait for it)
    CK(cudaMemcpyAsync(din, in, (size_t)8 * n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    launch_test_samplers(h->P, which, n, din, (unsigned long long)first_history, dout, h->stream);
    h->launches += 1;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, dout, (size_t)8 * n * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(din); cudaFree(dout);
    if (e != cudaSuccess) { h->err = std::string("omc_gpu_test_samplers: ") + cudaGetErrorString(e); return 1; }
    return 0;
}

int omc_gpu_test_rng(omc_gpu_handle h, long long hist, int n, double *out) {
    if (!h || !out || n <= 0) return 2;
    CK(cudaSetDevice(h->device));
    double *d = nullptr;
    CK(cudaMalloc((void **)&d, (size_t)n * sizeof(double)));
    launch_test_rng(h->P.seed0, h->P.seed1, (unsigned long long)hist, n, d, h->stream);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaMemcpy(out, d, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost));
    cudaFree(d);
    return 0;
}

int omc_gpu_test_format(omc_gpu_handle h, int mode, long long n, const double *values, const char *path) {
    if (!h || !values || !path || n < 0 || (mode != 0 && mode != 1)) return 2;
    CK(cudaSetDevice(h->device));
    if (format_setup(h)) return 1;
    double *d = nullptr;
    CK(cudaMalloc((void **)&d, (size_t)(n ? n : 1) * sizeof(double)));
    CK(cudaMemcpyAsync(d, values, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    FILE *fp = fopen(path, "w");
    if (!fp) { cudaFree(d); h->err = std::string("Unable to open file: ") + path; return 2; }
    int rc = format_block(h, fp, d, n, mode);
    if (fclose(fp) != 0 && !rc) rc = fail(h, "omc_gpu_test_format: write failed");
    cudaFree(d);
    return rc;
}

}  // extern "C"

// omc_wavefront.cu -- production path: event-based particle queues (BASELINE.json north_star (a)).
//
// Particles live in HBM as structure-of-arrays queues.  One "wave" moves every live particle through one
// event of its class, with one kernel per event class so that ALL warps on the machine run the same,
// instruction-cache-sized code at the same time:
//
//   misc_kernel       persistent; every WARP pulls typed 32-particle chunks from per-class tickets, classes in a fixed order:
//        P   photon flight       P[cur]  -> P[next] (flight unfinished) | IP[next] (at an interaction site)
//        IP  photon interaction  IP[cur] -> P[next], E[next]            Compton / pair / photo / Rayleigh
//        IE  e+- interaction     IE[cur] -> E[next], P[next]            brems / Moller / Bhabha / annihilation
//        S   source              initHistory() for new history ids -> P[next] or E[next]
//   esize_kernel      E[cur] -> step queue, class CH | BCA   cut-off test, distance to the next interaction, step-size limits
//   edo_kernel<CH>    CH  -> E[next] | IE[next]   condensed-history step (PRESTA-II msdist)
//   edo_kernel<BCA>   BCA -> E[next] | IE[next]   boundary-crossing / single-scattering step
//   advance_kernel    swaps cur/next, sizes the next injection so that `pool_size` particles stay in flight
//
// (interaction queues are consumed one wave after they were filled).  Sorting electrons into the CH and
// BCA queues after the step size is known matters because those two long paths otherwise split every warp.
// An earlier version fused everything into one persistent kernel with an in-block shared-memory re-sort;
// ncu showed it stalled on instruction fetch (stall_no_instruction on top, ~100 KB hot path, every block at
// a different pc) and on the re-sort barriers, see profiles/r01_wave_kernel_summary.md.
//
// Physics = the same device functions as the lock-step kernel (omc_physics.cuh) except that the elastic
// scattering angles are sampled in fp32 (omc_physics_f32.cuh).  Differences that are statistically neutral:
// every particle owns a Philox sub-stream derived from its parent's (results do not depend on scheduling);
// unread words of a Philox block are dropped at fixed physics points so that warps refill together; the
// reference's zero-length "second ustep iteration" (src/ompmc.c:4787 re-initialises total_tstep, see
// DESIGN.md) is not executed, its only effect being one wasted draw; with nsplit == 1 the unused
// survivor-index draw of photon() (:1916) is skipped.  Dose is scored with fp32 atomics into a chunk grid
// folded into the fp64 batch grid at the end of every omc_gpu_run_histories() call (north_star (d)).
#include "omc_physics.cuh"
#include "omc_physics_f32.cuh"
#include "omc_kernels.h"

#ifndef OMC_WAVE_F32
#define OMC_WAVE_F32 1     // 1: fp32 angle samplers (omc_physics_f32.cuh); 0: the fp64 ones of the lock-step kernel
#endif

#ifndef OMC_CH_BLOCK_RNG
#define OMC_CH_BLOCK_RNG 1   // 1: condensed-history step with whole-block draws, Rng in registers (msdist_b); 0: msdist_f
#endif
#ifndef OMC_PREFETCH
#define OMC_PREFETCH 1       // 1: the electron kernels prefetch the queue records of their NEXT loop iteration into L1
#endif
#ifndef OMC_WARP_AGGREGATE_DOSE
#define OMC_WARP_AGGREGATE_DOSE 0
#endif
// minimum resident blocks per SM asked of ptxas for each kernel (register cap 65536 / (128 * n)); measured on
// B200, v3 kernels: (5,6,6,6) 4.93e7 hist/s, (4,4,3,4) 4.30e7, (8,6,5,8) 4.90e7 -- latency hiding needs >= 20 warps/SM;
// round-2 kernels (slim records, packed draw plan): (5,5,5,6) 1.243e8, (6,6,6,6) 1.303e8 (80 registers everywhere)
#ifndef OMC_MB_MISC
#define OMC_MB_MISC 6
#endif
#ifndef OMC_MB_ESIZE
#define OMC_MB_ESIZE 6
#endif
#ifndef OMC_MB_ECH
#define OMC_MB_ECH 6
#endif
#ifndef OMC_MB_EBCA
#define OMC_MB_EBCA 6
#endif

namespace omc {

constexpr int NT = WAVE_THREADS;   // threads per block == particles per chunk

// ---- queues ----------------------------------------------------------------------------------
__device__ __forceinline__ void q_load_part(const PartQueue &q, unsigned i, Part &p, int &tag) {
    const int2 a = q.irq[i];                                   // first: a voxel-record load usually depends on it
    const double2 xy = q.xy[i], ze = q.ze[i];
    const float4 dw = q.dw[i];
    p.ir = a.x; p.iq = (int)(short)(a.y & 0xffff); tag = a.y >> 16;
    p.x = xy.x; p.y = xy.y; p.z = ze.x; p.e = ze.y; p.u = (double)dw.x; p.v = (double)dw.y; p.w = (double)dw.z; p.wt = (double)dw.w;
}

__device__ __forceinline__ void q_load(const PartQueue &q, unsigned i, Part &p, Rng &g, const DevProblem &P, double &aux, int &tag) {
    q_load_part(q, i, p, tag);
    aux = q.aux ? q.aux[i].x : 0.0;
    const uint4 r = q.rng[i];
    g.seed(P.seed0, P.seed1, ((unsigned long long)r.y << 32) | r.x, r.z, r.w);
}

__device__ __forceinline__ void q_store(const PartQueue &q, unsigned i, const Part &p, const Rng &g, double aux, int tag,
                                        double aux2 = 0.0) {
    q.xy[i] = make_double2(p.x, p.y); q.ze[i] = make_double2(p.z, p.e);
    q.dw[i] = make_float4((float)p.u, (float)p.v, (float)p.w, (float)p.wt);
    if (q.aux) q.aux[i] = make_double2(aux, aux2);             // photon queues only
    q.irq[i] = make_int2(p.ir, (p.iq & 0xffff) | (tag << 16));
    q.rng[i] = make_uint4(g.h0, g.h1, g.stream, g.ndraws());
}

__device__ __forceinline__ void prefetch_l1(const void *p) {
#if OMC_PREFETCH
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#endif
}
// the record of slot i of a particle queue / of the step queue (what the next grid-stride iteration will load)
__device__ __forceinline__ void q_prefetch_e(const PartQueue &q, unsigned i) {
    prefetch_l1(q.irq + i); prefetch_l1(q.rm + i); prefetch_l1(q.xy + i); prefetch_l1(q.ze + i); prefetch_l1(q.dw + i);
    prefetch_l1(q.rng + i);
}
__device__ __forceinline__ void es_prefetch(const EStepQueue &S, unsigned s) {
#pragma unroll
    for (int k = 0; k < 3; k++) prefetch_l1(S.v[k] + s);
    prefetch_l1(S.d + s); prefetch_l1(S.f + s); prefetch_l1(S.m + s); prefetch_l1(S.t + s); prefetch_l1(S.rng + s);
}

// warp-aggregated slot reservation: one atomic per (warp, queue) instead of one per lane
__device__ __forceinline__ unsigned q_reserve(unsigned *count) {
    const unsigned m = __match_any_sync(__activemask(), (unsigned long long)count);
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(m) - 1;
    unsigned base = 0;
    if (lane == leader) base = atomicAdd(count, (unsigned)__popc(m));
    base = __shfl_sync(m, base, leader);
    return base + __popc(m & ((1u << lane) - 1u));
}

__device__ __forceinline__ void q_push(const PartQueue &q, unsigned *count, WaveCtl *ctl, const Part &p, const Rng &g, double aux,
                                       int tag, double aux2 = 0.0) {
    const unsigned slot = q_reserve(count);
    if (slot >= q.cap) { atomicAdd(&ctl->overflow.v, 1u); return; }
    q_store(q, slot, p, g, aux, tag, aux2);
}

// push into an electron queue together with the voxel record of p.ir as the producer knows it (saves esize_kernel the
// dependent random load ir -> voxel record, its top stall); med = -2: not known
__device__ __forceinline__ void q_push_e(const PartQueue &q, unsigned *count, WaveCtl *ctl, const Part &p, const Rng &g, int tag,
                                         float rhof, int med) {
    const unsigned slot = q_reserve(count);
    if (slot >= q.cap) { atomicAdd(&ctl->overflow.v, 1u); return; }
    q_store(q, slot, p, g, 0.0, tag);
    q.rm[slot] = make_int2(__float_as_int(rhof), med);
}

// sub-stream of a particle created by `parent` (a function of the parent's stream position only)
__device__ __forceinline__ void child_rng(const Rng &parent, Rng &c, unsigned k) {
    unsigned s = parent.stream * 0x9E3779B1u + parent.ndraws() * 0x85EBCA77u + (k + 1u) * 0xC2B2AE3Du;
    s ^= s >> 15; s *= 0x2C1B3C6Du; s ^= s >> 12; s *= 0x297A2D39u; s ^= s >> 15;
    c = parent;
    c.stream = s | 1u;          // never 0: stream 0 is the primary's
    c.blk = 0; c.pos = 4;
}

struct Tally {
    unsigned ndep, nestep, npstep;
};

struct WaveArgs {
    WaveCtl *ctl;
    WaveQueues Q;
    int max_cross, ibeamlet, woodcock, max_virtual;
    // multi-beamlet pass (omc_gpu_run_beamlets): history id -> beamlet -> its own fp32 dose grid
    float *mb_grid;                 // [mb_n][nreg], nullptr: off
    unsigned long long mb_first;    // beamlet (mb_ib0 + k) owns history ids [mb_first + k * mb_per, + mb_per)
    unsigned mb_per, mb_n;
    int mb_ib0;
};

// Which of the two dose grids a particle scores into (batch pipelining, WaveCtl::hist_split), wave-uniform part
struct BatchSel {
    unsigned long long split;
    float *g_new, *g_old;
    unsigned has_old;
    // multi-beamlet pass: one grid per beamlet, beamlet = (history id - mb_first) / mb_per
    float *mb_grid;
    unsigned long long mb_first;
    unsigned mb_per;
    double mb_inv;
    size_t nreg;
    __device__ __forceinline__ void init(const DevProblem &P, const WaveCtl *c, const WaveArgs &A) {
        split = c->hist_split; has_old = c->has_old;
        g_new = P.endep32 + (size_t)c->grid_new * P.nreg;
        g_old = P.endep32 + (size_t)(c->grid_new ^ 1u) * P.nreg;
        mb_grid = A.mb_grid; mb_first = A.mb_first; mb_per = A.mb_per; nreg = (size_t)P.nreg;
        mb_inv = A.mb_per ? 1.0 / (double)A.mb_per : 0.0;
    }
    __device__ __forceinline__ bool is_old(uint32_t h0, uint32_t h1) const {
        return has_old && ((((unsigned long long)h1) << 32) | h0) < split;
    }
    // index of the beamlet that owns a history id (multi-beamlet pass)
    __device__ __forceinline__ unsigned beamlet_of(uint32_t h0, uint32_t h1) const {
        const unsigned long long x = ((((unsigned long long)h1) << 32) | h0) - mb_first;
        unsigned q = (unsigned)(__ull2double_rz(x) * mb_inv);
        if ((unsigned long long)(q + 1u) * mb_per <= x) q += 1u;
        else if ((unsigned long long)q * mb_per > x) q -= 1u;
        return q;
    }
    __device__ __forceinline__ float *grid(bool old, uint32_t h0, uint32_t h1) const {
        if (mb_grid != nullptr) return mb_grid + (size_t)beamlet_of(h0, h1) * nreg;
        return old ? g_old : g_new;
    }
    // consumers count the previous batch's particles they meet; zero in a whole wave = that batch is complete
    __device__ __forceinline__ void count(WaveCtl *c, unsigned mask, bool old) const {
        if (!has_old) return;
        const unsigned m = __ballot_sync(mask, old);
        if (m && (threadIdx.x & 31) == __ffs(mask) - 1) atomicAdd(&c->old_seen.v, (unsigned)__popc(m));
    }
};

// ausgab(): fp32 chunk grid (north_star (d)); optionally warp-aggregated when lanes hit the same voxel
__device__ __forceinline__ void deposit32(float *grid32, Tally &t, int ir, double en) {
    t.ndep++;
#if OMC_WARP_AGGREGATE_DOSE
    const unsigned m = __activemask();
    const unsigned peers = __match_any_sync(m, ir);
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(peers) - 1;
    float v = (float)en;
    for (unsigned rest = peers & ~(1u << leader); rest; rest &= rest - 1) {
        const int src = __ffs(rest) - 1;
        const float o = __shfl_sync(peers, v, src);
        if (lane == leader) v += o;
    }
    if (lane == leader) atomicAdd(grid32 + ir, v);
#else
    atomicAdd(grid32 + ir, (float)en);
#endif
}

enum { TAG_NONE = 0, TAG_COMPTON = 1, TAG_PAIR = 2, TAG_PHOTO = 3, TAG_RAYLEIGH = 4, TAG_BREMS = 5, TAG_MOLLER = 6, TAG_BHABHA = 7,
       TAG_ANNIH = 8, TAG_RANNIH = 9 };


// ---------------------------------------------------------------------------------------------
// chunk P: photon free flight, photon() src/ompmc.c:1884-2067 for nsplit == 1
// ---------------------------------------------------------------------------------------------
__device__ void photon_chunk(const DevProblem &P, const WaveArgs &A, const BatchSel &BS, int par, unsigned i, unsigned n, Tally &t) {
    if (i >= n) return;
    WaveCtl *ctl = A.ctl;
    Part p; Rng g; double dpmfp; int tag;
    q_load(A.Q.p[par], i, p, g, P, dpmfp, tag);
    const bool old = BS.is_old(g.h0, g.h1);
    BS.count(ctl, __activemask(), old);
    float *dg = BS.grid(old, g.h0, g.h1);
    RegionRec R = load_region_w(P, p.ir);
    // uniform photon splitting, :1903-1945: the record is the ray of nsplit copies of weight wt/nsplit whose
    // interaction depths are stratified (eta'_k = eta'_0 - k/nsplit); copy `isplit` is the one in flight.
    const int nsplit = P.nsplit;
    const double d_eta = 1.0 / (double)nsplit;
    int isplit = tag & 0xff, isurv = (tag >> 8) & 0xff;
    double eta = A.Q.p[par].aux[i].y;
    if (dpmfp < 0.0) {                                         // fresh photon: cut-off test + number of mfp
        if (p.e <= R.pcut || p.wt == 0) { deposit32(dg, t, p.ir, p.wt * p.e); return; }
        g.align();
        const double r = g.next();
        eta = 1.0 - r / (double)nsplit;                        // eta' of the first copy
        isplit = 0; isurv = 0;
        if (nsplit > 1) {                                      // (with nsplit == 1 the survivor draw :1916 is unused)
            p.wt /= (double)nsplit;
            isurv = (int)(g.next() * nsplit);
        }
        dpmfp = -log(eta);
    }
    const double gle = log(p.e);
    if (p.ir == 0) return;                                     // outside the phantom: howfar() discards (idisc)
    // Voxel march, :1951-2019 + howfar() omc_dosxyz.c:187-297, restated for speed (statistically neutral):
    // voxel indices are tracked incrementally instead of being decoded from the region number at every
    // crossing (three integer divisions), plane distances use the reciprocal direction cosines (three fp64
    // divisions per FLIGHT instead of per crossing), and the axis tests are selects, not branches.  Axis order
    // z, x, y with strict '<' as in the reference.
    int ix, iy, iz;
    decode_region(P, p.ir, ix, iy, iz);
    const double iu = (p.u != 0.0) ? 1.0 / p.u : 0.0, iv = (p.v != 0.0) ? 1.0 / p.v : 0.0, iw = (p.w != 0.0) ? 1.0 / p.w : 0.0;
    const int sx = p.u > 0.0, sy = p.v > 0.0, sz = p.w > 0.0;
    int imed = R.med, medc = -2;
    double sig0 = 0.0, sig = 0.0;                              // sig = 1 / gmfp
    bool hit = false;
    for (int k = 0; k < A.max_cross; k++) {
        double tstep = 1.0E8;
        if (imed != -1) {
            if (imed != medc) {                                // (imed, gle) -> table values, kept while the medium stays
                sig0 = phot_sig0(P, imed, gle);
                medc = imed;
            }
            sig = sig0 * R.rhof;
            tstep = dpmfp * (double)__frcp_rn((float)sig);
        }
        const double dz = (p.w != 0.0) ? (__ldg(P.zb + iz + sz) - p.z) * iw : 1.0E30;
        const double dx = (p.u != 0.0) ? (__ldg(P.xb + ix + sx) - p.x) * iu : 1.0E30;
        const double dy = (p.v != 0.0) ? (__ldg(P.yb + iy + sy) - p.y) * iv : 1.0E30;
        double ustep = tstep;
        int axis = -1;
        if (dz < ustep) { ustep = dz; axis = 2; }
        if (dx < ustep) { ustep = dx; axis = 0; }
        if (dy < ustep) { ustep = dy; axis = 1; }
        t.npstep++;
        p.x += ustep * p.u; p.y += ustep * p.v; p.z += ustep * p.w;
        if (imed != -1) dpmfp = fmax(0.0, dpmfp - ustep * sig);
        if (axis >= 0) {
            int ir = p.ir;
            bool out;
            if (axis == 2) { iz += sz ? 1 : -1; ir += sz ? P.ijmax : -P.ijmax; out = (iz < 0) | (iz >= P.ksize); }
            else if (axis == 0) { ix += sx ? 1 : -1; ir += sx ? 1 : -1; out = (ix < 0) | (ix >= P.isize); }
            else { iy += sy ? 1 : -1; ir += sy ? P.isize : -P.isize; out = (iy < 0) | (iy >= P.jsize); }
            if (out) return;                                   // left the phantom
            p.ir = ir;
            R = load_region_w(P, ir);
            imed = R.med;
        }
        if (imed != -1 && dpmfp <= 1.0E-05) {                  // interaction site of copy `isplit`, :2027-2067
            // The interaction type is sampled by the interaction chunk of the next wave (p_interact_chunk), where
            // all lanes of a warp do it together, from the copy's own sub-stream.
            const bool surv = (isplit == isurv);
            if (nsplit == 1) { hit = true; break; }            // the photon itself goes on, pushed below
            {
                Rng gq;
                child_rng(g, gq, (unsigned)isplit);
                q_push(A.Q.ip[par ^ 1], &ctl->n_ip[par ^ 1].v, ctl, p, gq, -1.0, TAG_NONE | (surv ? 16 : 0));
            }
            isplit += 1;
            const double eta_new = eta - d_eta;
            if (isplit >= nsplit || eta_new <= 0.0) return;    // all copies done
            dpmfp = log(eta / eta_new);                        // = -log(eta'_k) - sum of the previous copies' mfp
            eta = eta_new;
        }
    }
    if (hit) q_push(A.Q.ip[par ^ 1], &ctl->n_ip[par ^ 1].v, ctl, p, g, -1.0, TAG_NONE | 16);
    else q_push(A.Q.p[par ^ 1], &ctl->n_p[par ^ 1].v, ctl, p, g, dpmfp, isplit | (isurv << 8), eta);
}

// chunk P, Woodcock flight (nsplit == 1): the distance to the next TENTATIVE collision is sampled with the
// majorant cross section of the whole phantom, max over media of sigma_m(E) * rhomax_m; the collision is real
// with probability sigma(voxel) / majorant.  Interaction sites have exactly the distribution of the reference's
// voxel-to-voxel march (photon() src/ompmc.c:1951-2019: mean free paths accumulated over the voxels crossed),
// but a flight costs a few voxel look-ups instead of one per voxel crossed, and there is no dependent chain of
// voxel-record loads.  Draws are consumed in whole Philox blocks: {distance, accept} x 2 per block.
// One Woodcock flight of at most max_virtual tentative collisions.  Returns 1 = real collision at p (p.ir = its voxel),
// 0 = flight unfinished (p = last tentative site), -1 = the photon left the phantom.  `entered`: has been seen inside the box.
__device__ __forceinline__ int woodcock_flight(const DevProblem &P, Rng &g, Part &p, bool &entered, int max_virtual, unsigned &npstep) {
    const double gle = log(p.e);
    double smaj = 0.0;
    for (int m = 0; m < P.nmed; m++) {
        const double rmax = P.med[m].rhomax;
        if (rmax > 0.0) smaj = fmax(smaj, rmax * phot_sig0(P, m, gle));
    }
    if (!(smaj > 0.0)) return -1;                              // vacuum everywhere: the photon leaves
    const double imaj = 1.0 / smaj;
    const double x0 = __ldg(P.xb), x1 = __ldg(P.xb + P.isize), y0 = __ldg(P.yb), y1 = __ldg(P.yb + P.jsize),
                 z0 = __ldg(P.zb), z1 = __ldg(P.zb + P.ksize);
    int medc = -2;
    double sigc = 0.0;
    uint4 b = make_uint4(0u, 0u, 0u, 0u);
    for (int k = 0; k < max_virtual; k++) {
        if (!(k & 1)) b = g.block();
        const uint32_t w0 = (k & 1) ? b.z : b.x, w1 = (k & 1) ? b.w : b.y;
        const float r = ((float)(w0 >> 8) + 0.5f) * (1.0f / 16777216.0f);
        const double s = (double)(-__logf(r)) * imaj;
        p.x += s * p.u; p.y += s * p.v; p.z += s * p.w;
        npstep++;
        if (p.x >= x0 && p.x < x1 && p.y >= y0 && p.y < y1 && p.z >= z0 && p.z < z1) {
            entered = true;
        } else {
            if (entered) return -1;                            // left the phantom
            // Not inside yet: a source particle sitting exactly on a max face, or one that the matRad source
            // clamped with the wrong bound (omc_matrad.c:1225, SURVEY Q13: p.z = ybounds[0]).  The reference keeps
            // such a particle in the voxel layer it was assigned to until it crosses a plane; the clamped
            // look-up below does the same.  It is discarded if it moves away from the box.
            if ((p.x < x0 && p.u <= 0.0) || (p.x >= x1 && p.u >= 0.0) || (p.y < y0 && p.v <= 0.0) || (p.y >= y1 && p.v >= 0.0) ||
                (p.z < z0 && p.w <= 0.0) || (p.z >= z1 && p.w >= 0.0)) return -1;
        }
        const int ix = find_bin(P.xb, P.isize, p.x, P.inv_dx, P.uniform_x != 0);
        const int iy = find_bin(P.yb, P.jsize, p.y, P.inv_dy, P.uniform_y != 0);
        const int iz = find_bin(P.zb, P.ksize, p.z, P.inv_dz, P.uniform_z != 0);
        p.ir = 1 + ix + iy * P.isize + iz * P.ijmax;
        double rhof; int med;
        load_region_rm(P, p.ir, rhof, med);
        if (med < 0) continue;
        if (med != medc) { sigc = phot_sig0(P, med, gle); medc = med; }
        if ((double)w1 * (1.0 / 4294967296.0) * smaj < sigc * rhof) return 1;
    }
    return 0;
}

// (Measured and rejected, round 2: LANE REFILL -- a warp owning 256 photons, a lane taking the next one as soon as its flight
// ends, two tentative collisions per Philox block.  The flight loop then runs with nearly all lanes instead of 10 of 32, but the
// per-lane record loads are scattered and the loop carries the whole photon state: 1.105e8 against 1.299e8 histories/s.)
__device__ void photon_chunk_wc(const DevProblem &P, const WaveArgs &A, const BatchSel &BS, int par, unsigned i, unsigned n, Tally &t) {
    if (i >= n) return;
    WaveCtl *ctl = A.ctl;
    const PartQueue &q = A.Q.p[par];
    Part p; Rng g;
    bool entered;                                              // has been seen inside the phantom box (see woodcock_flight)
    {
        int tag;
        q_load_part(q, i, p, tag);
            entered = q.aux[i].x <= -2.0;
        const uint4 r = q.rng[i];
        g.seed_blocks(P.seed0, P.seed1, r.x, r.y, r.z, r.w);
    }
    const bool old = BS.is_old(g.h0, g.h1);
    BS.count(ctl, __activemask(), old);
    float *dg = BS.grid(old, g.h0, g.h1);
    if (p.ir == 0) return;                                     // outside the phantom: howfar() discards (idisc)
    {   // cut-off test of photon() :1884 (a flight that spans several waves repeats it, harmlessly)
        double rhof; int med;
        load_region_rm(P, p.ir, rhof, med);
        const double pcut = (P.reg8 != nullptr) ? (med >= 0 ? P.med[med].pcut : 0.0) : load_region(P, p.ir).pcut;
        if (p.e <= pcut || p.wt == 0) { deposit32(dg, t, p.ir, p.wt * p.e); return; }
    }
    const int st = woodcock_flight(P, g, p, entered, A.max_virtual, t.npstep);
    if (st < 0) return;
    if (st > 0) q_push(A.Q.ip[par ^ 1], &ctl->n_ip[par ^ 1].v, ctl, p, g, -1.0, TAG_NONE | 16);
    else q_push(A.Q.p[par ^ 1], &ctl->n_p[par ^ 1].v, ctl, p, g, entered ? -2.0 : -1.0, TAG_NONE);
}

// chunk IP: photon interactions
__device__ void p_interact_chunk(const DevProblem &P, const WaveArgs &A, const BatchSel &BS, int par, unsigned i, unsigned n) {
    if (i >= n) return;
    WaveCtl *ctl = A.ctl;
    const PartQueue &pn = A.Q.p[par ^ 1], &en = A.Q.e[par ^ 1];
    Part p, q; Rng g, gq; int tag;
    q_load_part(A.Q.ip[par], i, p, tag);
    {
        const uint4 r = A.Q.ip[par].rng[i];
        g.seed_blocks(P.seed0, P.seed1, r.x, r.y, r.z, r.w);
    }
    BS.count(ctl, __activemask(), BS.is_old(g.h0, g.h1));
    const RegionRec R = load_region_w(P, p.ir);
    const int imed = R.med;
    const float rho_f = (float)R.rhof;
    // photon splitting: scattered photons are kept for the surviving copy only and get the full weight back;
    // charged secondaries of every copy are kept with the copy's weight wt/nsplit (:2072-2093)
    const bool surv = (tag & 16) != 0;
    int type = tag & 15;
    const double back = (double)P.nsplit;
    if (type == TAG_NONE) {                                    // interaction choice, photon() :2027-2067
        const uint4 b = g.block();
        type = photon_interaction_type(P, imed, log(p.e), p.e, u32d(b.x), u32d(b.y));
        if (!surv && type == TAG_RAYLEIGH) return;             // a Rayleigh-scattered non-survivor is simply dropped, :2030-2034
    }
    if (type == TAG_COMPTON) {                                 // the common one: block draws, generator in registers
        compton_b(g, p, q);
        child_rng(g, gq, 0);
        if (surv) { p.wt *= back; q_push(pn, &ctl->n_p[par ^ 1].v, ctl, p, g, -1.0, TAG_NONE); }
        q_push_e(en, &ctl->n_e[par ^ 1].v, ctl, q, gq, TAG_NONE, rho_f, imed);
        return;
    }
    Rng g2 = g;                                                // rare ones: the word-by-word samplers shared with the lock-step kernel
    if (type == TAG_PAIR) {
        pair(P, g2, p, q, imed);
        child_rng(g2, gq, 0);
        q_push_e(en, &ctl->n_e[par ^ 1].v, ctl, p, g2, TAG_NONE, rho_f, imed);
        q_push_e(en, &ctl->n_e[par ^ 1].v, ctl, q, gq, TAG_NONE, rho_f, imed);
    } else if (type == TAG_PHOTO) {
        photo(g2, p, R.ecut);
        q_push_e(en, &ctl->n_e[par ^ 1].v, ctl, p, g2, TAG_NONE, rho_f, imed);
    } else {                                                   // Rayleigh (surviving copy only): direction change
        const MedRec &M = P.med[imed];
        const double gle = log(p.e);
        const int lgle = (int)(gle * M.ge1 + M.ge0) - 1;
        const PhotBin *B = P.phot + imed * MXGE + lgle;
        p.wt *= back;
        rayleigh(P, g2, p, pwl2(gle, &B->pmax1), p.e);
        q_push(pn, &ctl->n_p[par ^ 1].v, ctl, p, g2, -1.0, TAG_NONE);
    }
}

// chunk IE: discrete electron / positron interactions
__device__ void e_interact_chunk(const DevProblem &P, const WaveArgs &A, const BatchSel &BS, int par, unsigned i, unsigned n) {
    if (i >= n) return;
    WaveCtl *ctl = A.ctl;
    const PartQueue &pn = A.Q.p[par ^ 1], &en = A.Q.e[par ^ 1];
    Part p, q; Rng g, gq; int tag;
    q_load_part(A.Q.ie[par], i, p, tag);
    {
        const uint4 r = A.Q.ie[par].rng[i];
        g.seed_blocks(P.seed0, P.seed1, r.x, r.y, r.z, r.w);
    }
    BS.count(ctl, __activemask(), BS.is_old(g.h0, g.h1));
    double rho_d; int imed;
    load_region_rm(P, p.ir, rho_d, imed);
    const float rho_f = (float)rho_d;
    if (tag == TAG_MOLLER) {                                   // the common one: block draws, generator in registers
        const bool created = moller_b(P, g, p, q, imed);
        q_push_e(en, &ctl->n_e[par ^ 1].v, ctl, p, g, TAG_NONE, rho_f, imed);
        if (created) {
            child_rng(g, gq, 0);
            q_push_e(en, &ctl->n_e[par ^ 1].v, ctl, q, gq, TAG_NONE, rho_f, imed);
        }
        return;
    }
    Rng g2 = g;
    if (tag == TAG_BREMS) {
        brems(P, g2, p, q, imed, P.nsplit);                    // incl. Russian roulette of the photon when nsplit > 1
        child_rng(g2, gq, 0);
        q_push_e(en, &ctl->n_e[par ^ 1].v, ctl, p, g2, TAG_NONE, rho_f, imed);
        if (q.wt != 0.0) q_push(pn, &ctl->n_p[par ^ 1].v, ctl, q, gq, -1.0, TAG_NONE);
    } else if (tag == TAG_BHABHA) {
        bhabha(P, g2, p, q, imed);
        child_rng(g2, gq, 0);
        q_push_e(en, &ctl->n_e[par ^ 1].v, ctl, p, g2, TAG_NONE, rho_f, imed);
        q_push_e(en, &ctl->n_e[par ^ 1].v, ctl, q, gq, TAG_NONE, rho_f, imed);
    } else {                                                   // annihilation in flight / at rest
        if (tag == TAG_ANNIH) annih(g2, p, q, P.nsplit);
        else rannih(g2, p, q, P.nsplit);
        child_rng(g2, gq, 0);
        if (p.wt != 0.0) q_push(pn, &ctl->n_p[par ^ 1].v, ctl, p, g2, -1.0, TAG_NONE);
        if (q.wt != 0.0) q_push(pn, &ctl->n_p[par ^ 1].v, ctl, q, gq, -1.0, TAG_NONE);
    }
}

// chunk S: initHistory() for history ids hist_next + [i0, i0 + NT)
__device__ void source_chunk(const DevProblem &P, const WaveArgs &A, const BatchSel &BS, int par, unsigned i, unsigned n, double &ensrc) {
    if (i >= n) return;
    WaveCtl *ctl = A.ctl;
    Rng g;
    const unsigned long long hist = ctl->hist_next + i;
    g.seed(P.seed0, P.seed1, hist, 0u);
    Part p;
    const int ibeamlet = (A.mb_grid != nullptr) ? A.mb_ib0 + (int)BS.beamlet_of((uint32_t)hist, (uint32_t)(hist >> 32)) : A.ibeamlet;
    ensrc += init_history(P, g, p, ibeamlet);
    if (p.iq == 0) q_push(A.Q.p[par ^ 1], &ctl->n_p[par ^ 1].v, ctl, p, g, -1.0, TAG_NONE);
    else q_push_e(A.Q.e[par ^ 1], &ctl->n_e[par ^ 1].v, ctl, p, g, TAG_NONE, 0.0f, -2);
}

// ---------------------------------------------------------------------------------------------
// chunk E: condensed-history / boundary-crossing electron steps, the tstep/ustep loops of electron()
// src/ompmc.c:4694-5372, `iters` real steps per chunk.
// ---------------------------------------------------------------------------------------------
enum { CLS_NONE = 0, CLS_CH = 1, CLS_BCA = 2 };

// state handed from the "step size" phase to the "do the step" phase through shared memory
struct EStep {
    double eke, elke, demfp, sig0, total_tstep, range, tustep, tperp, rhof, ecut, dedx, blccl, ssmfp;
    int lelke, imed;
};
// (EStepQueue in omc_kernels.h carries Part + Rng + EStep between esize_kernel and the step kernels; eke = e - RM and
// the cut-off are recomputed / re-read instead of stored)
__device__ __forceinline__ double region_ecut(const DevProblem &P, int ir, int imed) {
    if (P.reg8 != nullptr) return imed >= 0 ? P.med[imed].ecut : 0.0;
    return load_region(P, ir).ecut;
}
__device__ __forceinline__ void es_put(const EStepQueue &S, unsigned s, const Part &p, const Rng &g, const EStep &e) {
    // "the step reaches the interaction point", electron() :5290 (total_tstep - tvstep * rhof < 1e-9), decided here where
    // total_tstep lives: for the full step tustep, and for any step at all (total_tstep itself below the tolerance)
    const unsigned flags = ((e.total_tstep - e.tustep * e.rhof < 1.0E-9) ? 1u : 0u) | ((e.total_tstep < 1.0E-9) ? 2u : 0u);
    S.v[0][s] = make_double2(p.x, p.y); S.v[1][s] = make_double2(p.z, p.e); S.v[2][s] = make_double2(e.tustep, e.range);
    S.d[s] = make_float4((float)p.u, (float)p.v, (float)p.w, (float)p.wt);
    S.f[s] = make_float4((float)e.demfp, (float)e.sig0, (float)e.rhof, (float)e.dedx);
    S.m[s] = make_uint4((unsigned)__float_as_int((float)e.blccl), (unsigned)__float_as_int((float)e.ssmfp), (unsigned)p.ir,
                        (unsigned)(p.iq + 1) | ((unsigned)(e.imed + 1) << 2) | (flags << 6) | ((unsigned)e.lelke << 16));
    S.t[s] = make_float2(__double2float_rd(e.tperp), (float)e.elke);
    S.rng[s] = make_uint4(g.h0, g.h1, g.stream, g.ndraws());
}
template <bool BLOCKS>
__device__ __forceinline__ void es_get(const EStepQueue &S, unsigned s, Part &p, Rng &g, EStep &e, const DevProblem &P) {
    const uint4 m = S.m[s];
    const double2 v0 = S.v[0][s], v1 = S.v[1][s], v2 = S.v[2][s];
    const float4 d = S.d[s], f = S.f[s];
    const float2 t = S.t[s];
    const uint4 r = S.rng[s];
    p.ir = (int)m.z; p.iq = (int)(m.w & 3u) - 1; e.imed = (int)((m.w >> 2) & 15u) - 1; e.lelke = (int)m.w >> 16;
    p.x = v0.x; p.y = v0.y; p.z = v1.x; p.e = v1.y; e.tustep = v2.x; e.range = v2.y;
    p.u = (double)d.x; p.v = (double)d.y; p.w = (double)d.z; p.wt = (double)d.w;
    e.tperp = (double)t.x; e.elke = (double)t.y;
    e.demfp = (double)f.x; e.sig0 = (double)f.y; e.rhof = (double)f.z; e.dedx = (double)f.w;
    e.blccl = (double)__int_as_float((int)m.x); e.ssmfp = (double)__int_as_float((int)m.y);
    // a stand-in that makes the :5290 test of estep_do() come out as the flags say (see es_put)
    const unsigned flags = (m.w >> 6) & 3u;
    e.total_tstep = (flags & 2u) ? 0.0 : ((flags & 1u) ? e.tustep * e.rhof : 1.0E30);
    e.eke = p.e - RM;
    e.ecut = region_ecut(P, p.ir, e.imed);
    if (BLOCKS) g.seed_blocks(P.seed0, P.seed1, r.x, r.y, r.z, r.w);
    else g.seed(P.seed0, P.seed1, ((unsigned long long)r.y << 32) | r.x, r.z, r.w);
}

// Phase A: cut-off test, distance to the next discrete interaction, step-size restrictions.
// Returns the step class, or CLS_NONE with `st` = -1 (finished) / TAG_RANNIH.
__device__ __forceinline__ int estep_size(const DevProblem &P, float *dg, Rng &g, Part &p, EStep &e, Tally &t, int &st, int2 rm) {
    RegionRec R;
    if (rm.y > -2 && P.reg8 != nullptr) {                      // voxel record handed over by the producer
        R.rhof = (double)__int_as_float(rm.x); R.med = rm.y; R.pcut = 0.0; R.pad = 0;
        R.ecut = (rm.y >= 0) ? P.med[rm.y].ecut : 0.0;
    } else {
        R = load_region_w(P, p.ir);
    }
    const int imed = R.med, iq = p.iq, qel = (1 + iq) / 2;
    const double eie = p.e;
    t.nestep++;
    st = 0;
    if (eie <= R.ecut) {                                       // :4665-4687
        deposit32(dg, t, p.ir, p.wt * (eie - RM));
        st = (iq > 0) ? TAG_RANNIH : -1;
        return CLS_NONE;
    }
    e.imed = imed; e.rhof = R.rhof; e.ecut = R.ecut;
    e.eke = eie - RM;
    if (imed == -1) {                                          // vacuum / outside: handled by the BCA group
        e.tustep = 10.0E8; e.tperp = 0.0; e.demfp = 0.0; e.sig0 = 0.0; e.total_tstep = 0.0; e.range = 0.0; e.elke = 0.0; e.lelke = 0;
        e.dedx = 0.0; e.blccl = 0.0; e.ssmfp = 0.0;
        return CLS_BCA;
    }
    const MedRec &M = P.med[imed];
    const ElecBin *B0 = P.ebin + (size_t)qel * P.nmed * MXEKE + imed * MXEKE;
    const double rhof = R.rhof, eke = e.eke;
    const double rinv = 1.0 / rhof;
    const uint4 blk = g.block();                               // one draw is used; block-wise so the state stays in registers
#if OMC_WAVE_F32
    // mixed precision (see omc_physics_f32.cuh): logs and ratios in fp32, energy / length sums in fp64
    float rf = (float)(blk.x >> 8) * (1.0f / 16777216.0f);
    if (rf == 0.0f) rf = 1.0E-30f;
    e.demfp = fmax((double)(-__logf(rf)), 1.0E-5);
    const double elke = flog(eke);
    const int lelke = elec_interval(M, elke);
    e.elke = elke; e.lelke = lelke;
    const ElecBin *B = B0 + lelke;
    const double2 re = ldg2(&B->range_ep);                     // {range_ep, e_array} of this bin
    const double dedx0 = pwl2(elke, &B->dedx1);
    double sig0;
    if (M.sig_ismonotone[qel]) sig0 = (double)fdiv((float)pwl2(elke, &B->sig1), (float)dedx0);
    else sig0 = (iq < 0) ? M.esig_e : M.psig_e;
    double tstep;
    e.total_tstep = 0.0;
    if (sig0 <= 0.0) {
        tstep = 10.0E8; sig0 = 1.0E-15;
    } else {
        const double ekef = eke - (double)fdiv((float)e.demfp, (float)sig0);
        if (ekef <= __ldg(&B0[0].e_array)) {
            tstep = 10.0E8;
        } else {
            const double elkef = flog(ekef);
            const int lelkef = elec_interval(M, elkef);
            if (lelkef == lelke) {
                tstep = drange_m(B, eke, ekef, elke, elkef);
            } else {
                const float ieke1 = frcp((float)M.eke1);
                double ekei = re.y, elkei = (double)(((float)(lelke + 1) - (float)M.eke0) * ieke1);
                const double tuss = drange_m(B, eke, ekei, elke, elkei);
                ekei = __ldg(&B0[lelkef + 1].e_array);
                elkei = (double)(((float)(lelkef + 2) - (float)M.eke0) * ieke1);
                tstep = drange_m(B0 + lelkef, ekei, ekef, elkei, elkef);
                tstep += tuss + re.x - __ldg(&B0[lelkef + 1].range_ep);
            }
        }
        e.total_tstep = tstep;
        tstep = tstep * rinv;
    }
    e.sig0 = sig0;
    e.dedx = rhof * dedx0;
    const double tmxs = pwl2(elke, &B->tmxs1) * rinv;
    {
        const double ekei = re.y, elkei = (double)fdiv((float)(lelke + 1) - (float)M.eke0, (float)M.eke1);
        e.range = (drange_m(B, eke, ekei, elke, elkei) + re.x) * rinv;
    }
    double tustep = fmin(fmin(tstep, tmxs), e.range);
    const double tperp = hownear_i(P, p);
    const float xccl = (float)(rhof * M.xcc);
    const float p2 = (float)(eke * (eke + 2.0 * RM));
    const float beta2 = fdiv(p2, p2 + RMf * RMf);
    const float etap = (float)pwl2(elke, &B->eta1);
    const float ms_corr = (float)pwl2(elke, &B->blcce1);
    float blcclf = (float)(rhof * M.blcc);
    blcclf = fdiv(fdiv(blcclf, etap), 1.0f + fdiv(0.25f * etap * xccl, blcclf * p2)) * ms_corr;
    const double blccl = (double)blcclf;
    const double ssmfp = (double)fdiv(beta2, blcclf);
#else
    double r = (double)blk.x * (1.0 / 4294967296.0);
    if (r == 0.0) r = 1.0E-30;
    e.demfp = fmax(-log(r), 1.0E-5);
    const double elke = log(eke);
    const int lelke = elec_interval(M, elke);
    e.elke = elke; e.lelke = lelke;
    const ElecBin *B = B0 + lelke;
    const double dedx0 = pwl2(elke, &B->dedx1);
    double sig0;
    if (M.sig_ismonotone[qel]) sig0 = pwl2(elke, &B->sig1) / dedx0;
    else sig0 = (iq < 0) ? M.esig_e : M.psig_e;
    double tstep;
    e.total_tstep = 0.0;
    if (sig0 <= 0.0) {
        tstep = 10.0E8; sig0 = 1.0E-15;
    } else {
        const double ekef = eke - e.demfp / sig0;
        if (ekef <= __ldg(&B0[0].e_array)) {
            tstep = 10.0E8;
        } else {
            const double elkef = log(ekef);
            const int lelkef = elec_interval(M, elkef);
            if (lelkef == lelke) {
                tstep = drange(B, eke, ekef, elke, elkef);
            } else {
                double ekei = __ldg(&B->e_array), elkei = (lelke + 1 - M.eke0) / M.eke1;
                const double tuss = drange(B, eke, ekei, elke, elkei);
                ekei = __ldg(&B0[lelkef + 1].e_array);
                elkei = ((lelkef + 2) - M.eke0) / M.eke1;
                tstep = drange(B0 + lelkef, ekei, ekef, elkei, elkef);
                tstep += tuss + __ldg(&B->range_ep) - __ldg(&B0[lelkef + 1].range_ep);
            }
        }
        e.total_tstep = tstep;
        tstep = tstep * rinv;
    }
    e.sig0 = sig0;
    e.dedx = rhof * dedx0;
    const double tmxs = pwl2(elke, &B->tmxs1) * rinv;
    {
        const double ekei = __ldg(&B->e_array), elkei = (lelke + 1 - M.eke0) / M.eke1;
        e.range = (drange(B, eke, ekei, elke, elkei) + __ldg(&B->range_ep)) * rinv;
    }
    double tustep = fmin(fmin(tstep, tmxs), e.range);
    const double tperp = hownear_i(P, p);
    double blccl = rhof * M.blcc;
    const double xccl = rhof * M.xcc;
    const double p2 = eke * (eke + 2.0 * RM);
    const double beta2 = p2 / (p2 + (RM * RM));
    const double etap = pwl2(elke, &B->eta1);
    const double ms_corr = pwl2(elke, &B->blcce1);
    blccl = blccl / etap / (1.0 + 0.25 * etap * xccl / blccl / p2) * ms_corr;
    const double ssmfp = beta2 / blccl;
#endif
    const double skindepth = 3 * ssmfp;
    tustep = fmin(tustep, fmax(tperp, skindepth));
    e.tustep = tustep; e.tperp = tperp; e.blccl = blccl; e.ssmfp = ssmfp;
    return ((tustep <= tperp) && (tustep > skindepth)) ? CLS_CH : CLS_BCA;
}

// Phase B: take the step.  Returns 0 = keep travelling, -1 = finished, TAG_* = interaction due.
// (rho_out, med_out) = voxel record of the region the electron ends in, when known (med_out = -2 otherwise)
__device__ __forceinline__ int estep_do(const DevProblem &P, float *dg, Rng &g, Part &p, const EStep &e, int cls, Tally &t, float &rho_out,
                                        int &med_out) {
    rho_out = (float)e.rhof; med_out = e.imed;
    const int iq = p.iq, qel = (1 + iq) / 2, imed = e.imed;
    double eie = p.e, ustep, tustep = e.tustep, tvstep, de = 0.0;
    const double rhof = e.rhof, eke0 = e.eke;
    bool call_howfar, do_single = false;
    double xf = 0, yf = 0, zf = 0, uf = 0, vf = 0, wf = 0;
    uint32_t w_rfict = 0;                                      // (block-draw CH step: the word kept for the sigma-ratio test)
    uint4 g0 = make_uint4(0u, 0u, 0u, 0u);                     // (block-draw BCA step: its first block)
    const ElecBin *B0 = (imed >= 0) ? P.ebin + (size_t)qel * P.nmed * MXEKE + imed * MXEKE : nullptr;
    if (cls == CLS_CH) {                                       // condensed-history step, :4973-4996
        call_howfar = false;
#if OMC_WAVE_F32
        de = eloss_m(B0, P.med[imed], rhof, 1.0 / rhof, tustep, e.range, eke0, e.elke, e.lelke);
#if OMC_CH_BLOCK_RNG
        ustep = msdist_b(P, g, p, imed, qel, rhof, de, tustep, eke0, xf, yf, zf, uf, vf, wf, w_rfict);
#else
        ustep = msdist_f(P, g, p, imed, qel, rhof, de, tustep, eke0, xf, yf, zf, uf, vf, wf);
#endif
#else
        de = eloss(B0, P.med[imed], rhof, tustep, e.range, eke0, e.elke, e.lelke);
        ustep = msdist<true>(P, g, p, imed, qel, rhof, de, tustep, eke0, xf, yf, zf, uf, vf, wf);
#endif
    } else if (imed == -1) {                                   // :4815-4821
        ustep = tustep; call_howfar = true;
    } else {                                                   // exact boundary crossing, :4997-5057
#if OMC_WAVE_F32 && OMC_CH_BLOCK_RNG
        g0 = g.block();                                        // {lambda, spin index rounding, rfict, branch}
        double r = (double)g0.x * (1.0 / 4294967296.0);
#else
        g.align();
        double r = g.next();
#endif
        if (r < 1.0E-30) r = 1.0E-30;
        const double lambda = (-1.0) * log(1.0 - r);
        double lambda_max = 0.5 * e.blccl * RM / e.dedx;
        lambda_max *= (eke0 / RM + 1.0) * (eke0 / RM + 1.0) * (eke0 / RM + 1.0);
        if (!(lambda >= 0.0 && lambda_max > 0.0)) return -1;   // Q8: dropped without deposit
        const double tuss = (lambda < lambda_max) ? lambda * e.ssmfp * (1.0 - 0.5 * lambda / lambda_max) : 0.5 * lambda * e.ssmfp;
        if (tuss < tustep) { tustep = tuss; do_single = true; }
        ustep = tustep;
        call_howfar = !(ustep < e.tperp);
    }
    const int irl = p.ir;
    int irnew = irl, idisc = 0;
    if (call_howfar) howfar_i(P, p, idisc, irnew, ustep);
    if (idisc > 0) {                                           // :5061-5088 (no annihilation quanta: edep > eie)
        deposit32(dg, t, p.ir, p.wt * ((iq > 0) ? p.e + RM : p.e - RM));
        return -1;
    }
    if (ustep < 0) ustep = 0.0;
    double ecut = e.ecut;
    if (ustep == 0.0 || imed == -1) {                          // :5097-5146
        if (ustep != 0.0) { p.x += p.u * ustep; p.y += p.v * ustep; p.z += p.w * ustep; }
        if (irnew != irl) {
            p.ir = irnew;
            const RegionRec R = load_region_w(P, irnew);
            ecut = R.ecut; rho_out = (float)R.rhof; med_out = R.med;
        }
        if (eie <= ecut) {
            deposit32(dg, t, p.ir, p.wt * (eie - RM));
            return (iq > 0) ? TAG_RANNIH : -1;
        }
        return 0;
    }
    const MedRec &M = P.med[imed];
    if (cls != CLS_CH) {
        tvstep = call_howfar ? ustep : tustep;
        if (call_howfar && tvstep != tustep) do_single = false;
#if OMC_WAVE_F32
        de = eloss_m(B0, M, rhof, 1.0 / rhof, tvstep, e.range, eke0, e.elke, e.lelke);
#else
        de = eloss(B0, M, rhof, tvstep, e.range, eke0, e.elke, e.lelke);
#endif
        xf = p.x + p.u * ustep; yf = p.y + p.v * ustep; zf = p.z + p.w * ustep;
        if (do_single) {                                       // :5180-5207
            const double ekems = fmax(eke0 - de, ecut - RM);
            const double p2 = ekems * (ekems + 2.0 * RM);
            const double beta2 = p2 / (p2 + (RM * RM));
            double chia2 = M.xcc / (4.0 * M.blcc * p2);
            const double elkems = log(ekems);
            const int lelkems = elec_interval(M, elkems);
            chia2 *= pwl2(elkems, &B0[lelkems].eta1);
            double costhe, sinthe;
            Frame fr;
#if OMC_WAVE_F32 && OMC_CH_BLOCK_RNG
            float cf, sf, cphi, sphi;
            sscat_b(P, g, imed, qel, (float)chia2, (float)elkems, (float)beta2, g0.y, cf, sf, cphi, sphi);
            costhe = (double)cf; sinthe = (double)sf;
            fr.cphi = (double)cphi; fr.sphi = (double)sphi;
            fr.A = p.u; fr.B = p.v; fr.C = p.w;
            frame_apply(fr, costhe, sinthe, p);
#else
            g.align();
#if OMC_WAVE_F32
            float cf, sf;
            sscat_f(P, g, imed, qel, (float)chia2, (float)elkems, (float)beta2, cf, sf);
            costhe = (double)cf; sinthe = (double)sf;
#else
            sscat(P, g, imed, qel, chia2, elkems, beta2, costhe, sinthe);
#endif
            g.align();
            uphi21(g, fr, costhe, sinthe, p);
#endif
        }
        uf = p.u; vf = p.v; wf = p.w;
    } else {
        tvstep = tustep;
    }
    deposit32(dg, t, p.ir, p.wt * de);                          // :5245
    p.x = xf; p.y = yf; p.z = zf; p.u = uf; p.v = vf; p.w = wf;
    eie -= de;
    p.e = eie;
    if (irnew == irl && eie <= ecut) {
        deposit32(dg, t, p.ir, p.wt * (eie - RM));
        return (iq > 0) ? TAG_RANNIH : -1;
    }
#if OMC_WAVE_F32
    const double eke = eie - RM, elke = flog(eke);
#else
    const double eke = eie - RM, elke = log(eke);
#endif
    const int lelke = elec_interval(M, elke);
    int imed_new = imed;
    if (irnew != irl) {
        p.ir = irnew;
        const RegionRec R = load_region_w(P, irnew);
        imed_new = R.med; ecut = R.ecut;
        rho_out = (float)R.rhof; med_out = R.med;
    }
    if (eie <= ecut) {
        deposit32(dg, t, p.ir, p.wt * (eie - RM));
        return (iq > 0) ? TAG_RANNIH : -1;
    }
    if (imed_new != imed) return 0;                            // new medium: resample from the top
    double demfp = e.demfp - de * e.sig0;
    if (e.total_tstep - tvstep * rhof < 1.0E-9) demfp = 0.0;
    if (demfp >= 1.0E-5) return 0;   // interaction point not reached (the reference burns a zero step, then resamples)
    // fictitious cross-section rejection, :5354-5372
    const ElecBin *B = B0 + lelke;
    const double sigf = pwl2(elke, &B->sig1) / pwl2(elke, &B->dedx1);
    const double br1 = pwl2(elke, &B->bra1);   // :5375-5429
    double r;
#if OMC_WAVE_F32 && OMC_CH_BLOCK_RNG
    if (cls == CLS_CH) {
        if ((double)w_rfict * (1.0 / 4294967296.0) >= sigf / e.sig0) return 0;
        r = (double)g.block().x * (1.0 / 4294967296.0);
    } else {
        if ((double)g0.z * (1.0 / 4294967296.0) >= sigf / e.sig0) return 0;
        r = (double)g0.w * (1.0 / 4294967296.0);
    }
#else
    {
        g.align();
        const double rfict = g.next();
        if (rfict >= sigf / e.sig0) return 0;
        r = g.next();
    }
#endif
    if (iq < 0) {
        if (r <= br1) return TAG_BREMS;
        if (p.e <= M.thmoll) return (br1 <= 0) ? 0 : TAG_BREMS;
        return TAG_MOLLER;
    }
    if (r < br1) return TAG_BREMS;
    const double pbr2 = pwl2(elke, &B->brb1);
    return (r < pbr2) ? TAG_BHABHA : TAG_ANNIH;
}

__device__ __forceinline__ void flush_tally(const DevProblem &P, Tally &t, double ensrc) {
    for (int o = 16; o > 0; o >>= 1) {
        ensrc += __shfl_xor_sync(0xffffffffu, ensrc, o);
        t.ndep += __shfl_xor_sync(0xffffffffu, t.ndep, o);
        t.nestep += __shfl_xor_sync(0xffffffffu, t.nestep, o);
        t.npstep += __shfl_xor_sync(0xffffffffu, t.npstep, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (ensrc != 0.0) atomicAdd(P.ensrc, ensrc);
        if (t.ndep) atomicAdd(&P.counters->deposits, (unsigned long long)t.ndep);
        if (t.nestep) atomicAdd(&P.counters->electron_steps, (unsigned long long)t.nestep);
        if (t.npstep) atomicAdd(&P.counters->photon_steps, (unsigned long long)t.npstep);
    }
}

// ---------------------------------------------------------------------------------------------
// electron kernels.  esize_kernel sizes the next step of every electron in E[cur] and sorts it into the CH or BCA
// class of the step queue; the step kernels push survivors back to E[next].
// (Measured alternative, not kept: sizing the next step of a surviving electron inside the step kernels and
// pushing it straight into the next wave's class queue removes a third of the queue traffic but ran 7 % SLOWER
// on B200 -- 5.28e7 vs 5.67e7 histories/s: the kernels are latency bound, not bandwidth bound, and the longer
// per-thread dependent chain plus finished lanes idling through the sizing code cost more.)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT, OMC_MB_ESIZE) esize_kernel(const __grid_constant__ DevProblem P, const __grid_constant__ WaveArgs A) {
    WaveCtl *ctl = A.ctl;
    const int par = (int)ctl->parity;
    const PartQueue &q = A.Q.e[par];
    const EStepQueue &S = A.Q.es;
    const unsigned n = min(ctl->n_e[par].v, q.cap);
    const unsigned lane = threadIdx.x & 31u, stride = gridDim.x * blockDim.x;
    const unsigned lt = (1u << lane) - 1u;
    Tally t = {0, 0, 0};
    BatchSel BS;
    BS.init(P, ctl, A);
    // warp-uniform loop: all 32 lanes stay converged through the slot reservation and the store
    for (unsigned base = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < n; base += stride) {
        const unsigned i = base + lane;
        Part p; Rng g; EStep e;
        int cls = CLS_NONE, st = 0;
        bool old = false;
        if (i < n) {
            int tag;
            q_load_part(q, i, p, tag);
            if (i + stride < n) q_prefetch_e(q, i + stride);
            {
                const int2 rm = q.rm[i];
                const uint4 r = q.rng[i];
                g.seed_blocks(P.seed0, P.seed1, r.x, r.y, r.z, r.w);
                old = BS.is_old(r.x, r.y);
                cls = estep_size(P, BS.grid(old, r.x, r.y), g, p, e, t, st, rm);
            }
        }
        BS.count(ctl, 0xffffffffu, old);
        // one reservation per (warp, class): lane 0 asks for the CH slots, lane 1 for the BCA slots, together.
        // (The return of these two atomics is the kernel's top stall site -- 24 % of the samples in the round-2 capture.  Measured
        // and rejected: warp-private regions of the step queue, no atomics at all, the step kernels reading region by region:
        // 1.234e8 against 1.299e8 histories/s -- thousands of separate write streams and the ragged region ends cost more than
        // the round trip they save.)
        const unsigned m_ch = __ballot_sync(0xffffffffu, cls == CLS_CH), m_bca = __ballot_sync(0xffffffffu, cls == CLS_BCA);
        unsigned b = 0;
        if (lane == 0 && m_ch) b = atomicAdd(&ctl->n_ch.v, (unsigned)__popc(m_ch));
        if (lane == 1 && m_bca) b = atomicAdd(&ctl->n_bca.v, (unsigned)__popc(m_bca));
        const unsigned b_ch = __shfl_sync(0xffffffffu, b, 0), b_bca = __shfl_sync(0xffffffffu, b, 1);
        if (cls != CLS_NONE) {
            const unsigned k = (cls == CLS_CH) ? b_ch + __popc(m_ch & lt) : b_bca + __popc(m_bca & lt);
            if (k < S.cap) es_put(S, (cls == CLS_CH) ? k : 2u * S.cap - 1u - k, p, g, e);
            else atomicAdd(&ctl->overflow.v, 1u);
        } else if (st > 0) {
            q_push(A.Q.ie[par ^ 1], &ctl->n_ie[par ^ 1].v, ctl, p, g, 0.0, st);
        }
    }
    flush_tally(P, t, 0.0);
}

template <int CLS>
__global__ void __launch_bounds__(NT, (CLS == 1 ? OMC_MB_ECH : OMC_MB_EBCA)) edo_kernel(const __grid_constant__ DevProblem P, const __grid_constant__ WaveArgs A) {
    WaveCtl *ctl = A.ctl;
    const int par = (int)ctl->parity;
    const EStepQueue &S = A.Q.es;
    const unsigned n = min(CLS == CLS_CH ? ctl->n_ch.v : ctl->n_bca.v, S.cap);
    const PartQueue &qe = A.Q.e[par ^ 1], &qi = A.Q.ie[par ^ 1];
    const unsigned lane = threadIdx.x & 31u, stride = gridDim.x * blockDim.x, lt = (1u << lane) - 1u;
    Tally t = {0, 0, 0};
    BatchSel BS;
    BS.init(P, ctl, A);
    // warp-uniform loop; one converged reservation per warp for both output queues (lane 0: E, lane 1: IE)
    for (unsigned base = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < n; base += stride) {
        const unsigned i = base + lane;
        Part p; Rng g; EStep e;
        float rho_new = 0.0f; int med_new = -2, st = -1;
        if (i < n) {
            es_get<OMC_WAVE_F32 && OMC_CH_BLOCK_RNG>(S, (CLS == CLS_CH) ? i : 2u * S.cap - 1u - i, p, g, e, P);
            if (i + stride < n) es_prefetch(S, (CLS == CLS_CH) ? i + stride : 2u * S.cap - 1u - (i + stride));
            st = estep_do(P, BS.grid(BS.is_old(g.h0, g.h1), g.h0, g.h1), g, p, e, CLS, t, rho_new, med_new);
        }
        const unsigned m_e = __ballot_sync(0xffffffffu, st == 0), m_i = __ballot_sync(0xffffffffu, st > 0);
        unsigned b = 0;
        if (lane == 0 && m_e) b = atomicAdd(&ctl->n_e[par ^ 1].v, (unsigned)__popc(m_e));
        if (lane == 1 && m_i) b = atomicAdd(&ctl->n_ie[par ^ 1].v, (unsigned)__popc(m_i));
        const unsigned b_e = __shfl_sync(0xffffffffu, b, 0), b_i = __shfl_sync(0xffffffffu, b, 1);
        if (st == 0) {
            const unsigned slot = b_e + __popc(m_e & lt);
            if (slot < qe.cap) { q_store(qe, slot, p, g, 0.0, TAG_NONE); qe.rm[slot] = make_int2(__float_as_int(rho_new), med_new); }
            else atomicAdd(&ctl->overflow.v, 1u);
        } else if (st > 0) {
            const unsigned slot = b_i + __popc(m_i & lt);
            if (slot < qi.cap) q_store(qi, slot, p, g, 0.0, st);
            else atomicAdd(&ctl->overflow.v, 1u);
        }
    }
    flush_tally(P, t, 0.0);
}

// ---------------------------------------------------------------------------------------------
// photons, interactions, source: persistent kernel, typed chunks
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT, OMC_MB_MISC) misc_kernel(const __grid_constant__ DevProblem P, const __grid_constant__ WaveArgs A) {
    // Every WARP pulls its own typed 32-particle chunks (no block-level barrier: an earlier block-granular version
    // spent 14 % of its stall samples on the per-chunk __syncthreads(), waiting for the slowest warp).
    constexpr unsigned CH = 32u;
    WaveCtl *ctl = A.ctl;
    const int par = (int)ctl->parity;
    const int lane = threadIdx.x & 31;
    // chunk classes: 0 electron interactions, 1 photon interactions, 2 source, 3 photon flights
    // (counts are clamped to the queue capacity: after an overflow the counters run past it)
    const unsigned cap = A.Q.p[0].cap;
    const unsigned cnt0 = min(ctl->n_ie[par].v, cap), cnt1 = min(ctl->n_ip[par].v, cap), cnt2 = ctl->n_src, cnt3 = min(ctl->n_p[par].v, cap);
    Tally t = {0, 0, 0};
    double ensrc = 0.0;
    BatchSel BS;
    BS.init(P, ctl, A);
    // All warps of the machine work through the classes in the same order (heaviest first), so that at any time
    // nearly every warp of an SM runs the same code: with each warp picking classes round-robin the kernel's top
    // stall was instruction fetch (ncu: stall_no_instruction 5.6 warps per issue).
#pragma unroll 1
    for (int k = 0; k < 4; k++) {
        const unsigned c = (k == 0) ? 3u : (k == 1) ? 1u : (k == 2) ? 0u : 2u;
        const unsigned cnt = c == 0 ? cnt0 : (c == 1 ? cnt1 : (c == 2 ? cnt2 : cnt3));
        for (;;) {
            unsigned chunk = 0;
            if (lane == 0) chunk = atomicAdd(&ctl->tk[c].v, 1u);
            chunk = __shfl_sync(0xffffffffu, chunk, 0);
            if ((unsigned long long)chunk * CH >= cnt) break;
            const unsigned i = chunk * CH + lane;
            if (c == 3) {
                if (A.woodcock) photon_chunk_wc(P, A, BS, par, i, cnt3, t);
                else photon_chunk(P, A, BS, par, i, cnt3, t);
            } else if (c == 2) source_chunk(P, A, BS, par, i, cnt2, ensrc);
            else if (c == 1) p_interact_chunk(P, A, BS, par, i, cnt1);
            else e_interact_chunk(P, A, BS, par, i, cnt0);
            __syncwarp();
        }
    }
    flush_tally(P, t, ensrc);
}

// swap cur/next and size the next injection (one thread, after all kernels of the wave)
__global__ void advance_kernel(const __grid_constant__ DevProblem P, WaveCtl *c) {
    if (blockIdx.x || threadIdx.x) return;
    const int par = (int)c->parity, nxt = par ^ 1;
    c->hist_next += c->n_src;
    P.counters->histories += c->n_src;
    c->n_p[par].v = 0; c->n_e[par].v = 0; c->n_ip[par].v = 0; c->n_ie[par].v = 0; c->n_ch.v = 0; c->n_bca.v = 0;
    const unsigned live = c->n_p[nxt].v + c->n_e[nxt].v + c->n_ip[nxt].v + c->n_ie[nxt].v;
    const unsigned long long left = c->hist_end - c->hist_next;
    // photon splitting multiplies the particles a history puts into the queues: a photon record in flight is a ray that will
    // still release up to nsplit interaction sites over its next waves, so it counts as nsplit particles when the next
    // injection is sized (counting it as one let the population overshoot the queue capacity at nsplit = 20)
    // (2 nsplit: the surviving scattered photon is split again at its next flight, so a history at nsplit = 20 has ~40 charged
    // particles alive within a few waves of each other -- measured 836 electron steps per history against 42 without splitting)
    const unsigned ns = (unsigned)(P.nsplit > 1 ? 2 * P.nsplit : 1);
    const unsigned long long load = (unsigned long long)live + (unsigned long long)(ns - 1u) * c->n_p[nxt].v;
    // With splitting the secondaries of an injection show up over the following waves (and their number per history depends on
    // the cut-offs: delta rays down to AE), so the pool is filled by an eighth of the estimated room per wave -- the controller
    // sees what a history really costs before it has committed the whole pool (the queues hold twice the target).
    const unsigned room = (unsigned)(((load < c->target) ? c->target - load : 0ull) / ns) / (P.nsplit > 1 ? 8u : 1u);
    c->n_src = (unsigned)(left < (unsigned long long)room ? left : (unsigned long long)room);
    c->live = live;
    if (c->has_old) {                                          // nothing of the previous batch was met in this wave: it is complete
        if (c->old_seen.v == 0) { c->has_old = 0; c->old_done = 1; }
        c->old_last = c->old_seen.v;
        c->old_seen.v = 0;
    }
    c->tk[0].v = c->tk[1].v = c->tk[2].v = c->tk[3].v = c->tk[4].v = 0;
    c->parity = (unsigned)nxt;
    c->waves += 1;
    if (c->overflow.v) P.counters->errors = c->overflow.v;
}

// fold the fp32 chunk grid into the fp64 batch grid
__global__ void flush_kernel(float *__restrict__ g32, double *__restrict__ g64, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float v = g32[i];
        if (v != 0.0f) { g64[i] += (double)v; g32[i] = 0.0f; }
    }
}

// ---- host-side launchers ----------------------------------------------------------------------
void wave_blocks_per_sm(int out[4]) {
    const void *k[4] = {(const void *)misc_kernel, (const void *)esize_kernel, (const void *)edo_kernel<CLS_CH>,
                        (const void *)edo_kernel<CLS_BCA>};
    for (int i = 0; i < 4; i++) {
        int n = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k[i], NT, 0) != cudaSuccess || n < 1) n = 1;
        out[i] = n;
    }
}

// One wave.  misc_kernel (photons / interactions / source) and the electron chain touch disjoint inputs and
// append to the same output queues through atomic counters, so they run as parallel branches; so do the two step
// kernels after esize_kernel (forked and joined with events; under stream capture this becomes a fork/join graph).
void launch_wave(const DevProblem &P, WaveCtl *ctl, const WaveQueues &Q, const WaveLaunch &L, const WaveStreams &W) {
    WaveArgs A;
    A.ctl = ctl; A.Q = Q; A.max_cross = L.max_cross; A.ibeamlet = L.ibeamlet;
    A.woodcock = L.woodcock; A.max_virtual = L.max_virtual;
    A.mb_grid = L.mb_grid; A.mb_first = L.mb_first; A.mb_per = L.mb_per; A.mb_n = L.mb_n; A.mb_ib0 = L.mb_ib0;
    const bool par = (W.s2 != nullptr);
    cudaStream_t s = W.s, sm = par ? W.s2 : s, sb = par ? W.s3 : s;
    if (par) { cudaEventRecord(W.fork, s); cudaStreamWaitEvent(W.s2, W.fork, 0); }
    misc_kernel<<<L.blocks[0], NT, 0, sm>>>(P, A);
    if (par) cudaEventRecord(W.join, W.s2);
    esize_kernel<<<L.blocks[1], NT, 0, s>>>(P, A);
    if (par) { cudaEventRecord(W.fork3, s); cudaStreamWaitEvent(W.s3, W.fork3, 0); }
    edo_kernel<CLS_BCA><<<L.blocks[3], NT, 0, sb>>>(P, A);
    if (par) cudaEventRecord(W.join3, W.s3);
    edo_kernel<CLS_CH><<<L.blocks[2], NT, 0, s>>>(P, A);
    if (par) { cudaStreamWaitEvent(s, W.join, 0); cudaStreamWaitEvent(s, W.join3, 0); }
    advance_kernel<<<1, 32, 0, s>>>(P, ctl);
}

// Next batch into the running pipeline: new history range, the dose grids swap roles, whatever is still alive
// belongs to the previous batch (ids below `first`).
__global__ void rearm_kernel(WaveCtl *c, unsigned long long first, unsigned long long nhist, unsigned nsplit) {
    if (blockIdx.x || threadIdx.x) return;
    c->hist_next = first; c->hist_end = first + nhist; c->hist_split = first;
    c->grid_new ^= 1u;
    c->has_old = (c->live > 0) ? 1u : 0u;
    c->old_done = c->has_old ? 0u : 1u;
    c->old_seen.v = 0;
    c->old_last = c->live;                                     // (not counted yet: the first wave of the new batch will)
    const unsigned ns = nsplit > 1u ? 2u * nsplit : 1u;           // (a split ray in flight counts as 2 nsplit particles, see advance_kernel)
    const unsigned long long load = (unsigned long long)c->live + (unsigned long long)(ns - 1u) * c->n_p[c->parity].v;
    const unsigned room = (unsigned)(((load < c->target) ? c->target - load : 0ull) / ns) / (nsplit > 1u ? 8u : 1u);
    c->n_src = (unsigned)(nhist < (unsigned long long)room ? nhist : (unsigned long long)room);
}
void launch_rearm(WaveCtl *ctl, unsigned long long first, unsigned long long nhist, unsigned nsplit, cudaStream_t s) {
    rearm_kernel<<<1, 32, 0, s>>>(ctl, first, nhist, nsplit);
}

// ---------------------------------------------------------------------------------------------
// Multi-beamlet pass: accumulateResults(1, nhist, nbatch) + threshold + sparse column assembly of
// omc_matrad.c:1416-1477 on the device, one block per beamlet.  dose() below = accumulateResults() for one voxel.
// ---------------------------------------------------------------------------------------------
struct MbGeom {
    const double *xb, *yb, *zb, *dens;
    int isize, jsize;
    double inc_fluence, nbatch;
};
__device__ __forceinline__ double mb_dose(const float *grid, const MbGeom &G, long long v) {
    double endep = (double)grid[v + 1];
    endep /= G.nbatch;
    if (endep == 0.0) return 0.0;
    const double dens = G.dens[v];
    if (dens < 0.044) return 0.0;
    const int ix = (int)(v % G.isize), iy = (int)((v / G.isize) % G.jsize), iz = (int)(v / ((long long)G.isize * G.jsize));
    double mass = (G.xb[ix + 1] - G.xb[ix]) * (G.yb[iy + 1] - G.yb[iy]) * (G.zb[iz + 1] - G.zb[iz]);
    mass *= dens;
    return endep * (1.602E-10 / (mass * G.inc_fluence));
}
// pass 1: dmax[b]; pass 2 (thresh != nullptr): nnz[b] = #{dose > rel * dmax[b]}
__global__ void mb_scan_kernel(const float *__restrict__ grids, long long nreg, int nb, MbGeom G, double rel, double *dmax,
                               unsigned long long *nnz, int pass) {
    __shared__ double s_max[32];
    __shared__ unsigned long long s_cnt[32];
    for (int b = blockIdx.x; b < nb; b += gridDim.x) {
        const float *grid = grids + (size_t)b * nreg;
        const double thresh = pass ? dmax[b] * rel : 0.0;
        double m = 0.0;
        unsigned long long cnt = 0;
        for (long long v = threadIdx.x; v < nreg - 1; v += blockDim.x) {
            const double d = mb_dose(grid, G, v);
            if (pass) cnt += (d > thresh) ? 1ull : 0ull;
            else m = fmax(m, d);
        }
        for (int o = 16; o > 0; o >>= 1) {
            m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
            cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        }
        if ((threadIdx.x & 31) == 0) { s_max[threadIdx.x >> 5] = m; s_cnt[threadIdx.x >> 5] = cnt; }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (unsigned w = 1; w < blockDim.x / 32; w++) { m = fmax(m, s_max[w]); cnt += s_cnt[w]; }
            if (pass) nnz[b] = cnt; else dmax[b] = m;
        }
        __syncthreads();
    }
}
// pass 3: ordered compaction (rows ascending, as the reference's irl loop writes them)
__global__ void mb_fill_kernel(const float *__restrict__ grids, long long nreg, int nb, MbGeom G, double rel, const double *dmax,
                               const long long *jc, long long *ir, double *val) {
    __shared__ unsigned s_warp[32];
    __shared__ unsigned long long s_base;
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, nwarp = blockDim.x / 32;
    for (int b = blockIdx.x; b < nb; b += gridDim.x) {
        const float *grid = grids + (size_t)b * nreg;
        const double thresh = dmax[b] * rel;
        if (threadIdx.x == 0) s_base = (unsigned long long)jc[b];
        __syncthreads();
        for (long long v0 = 0; v0 < nreg - 1; v0 += blockDim.x) {
            const long long v = v0 + threadIdx.x;
            double d = 0.0;
            bool keep = false;
            if (v < nreg - 1) { d = mb_dose(grid, G, v); keep = d > thresh; }
            const unsigned m = __ballot_sync(0xffffffffu, keep);
            if (lane == 0) s_warp[warp] = (unsigned)__popc(m);
            __syncthreads();
            unsigned before = 0, total = 0;
            for (unsigned w = 0; w < nwarp; w++) { const unsigned c = s_warp[w]; if (w < warp) before += c; total += c; }
            if (keep) {
                const unsigned long long at = s_base + before + (unsigned)__popc(m & ((1u << lane) - 1u));
                ir[at] = v; val[at] = d;
            }
            __syncthreads();
            if (threadIdx.x == 0) s_base += total;
            __syncthreads();
        }
    }
}
void launch_mb_scan(const float *grids, long long nreg, int nb, const DevProblem &P, const double *dens, int nhist, int nbatch, double rel,
                    double *dmax, unsigned long long *nnz, int pass, cudaStream_t s) {
    MbGeom G{P.xb, P.yb, P.zb, dens, P.isize, P.jsize, (double)nhist, (double)nbatch};
    mb_scan_kernel<<<nb < 148 * 4 ? nb : 148 * 4, 512, 0, s>>>(grids, nreg, nb, G, rel, dmax, nnz, pass);
}
void launch_mb_fill(const float *grids, long long nreg, int nb, const DevProblem &P, const double *dens, int nhist, int nbatch, double rel,
                    const double *dmax, const long long *jc, long long *ir, double *val, cudaStream_t s) {
    MbGeom G{P.xb, P.yb, P.zb, dens, P.isize, P.jsize, (double)nhist, (double)nbatch};
    mb_fill_kernel<<<nb < 148 * 4 ? nb : 148 * 4, 512, 0, s>>>(grids, nreg, nb, G, rel, dmax, jc, ir, val);
}

// ---------------------------------------------------------------------------------------------
// Unit-test hook (omc_gpu_test_samplers, include/ompmc_b200.h): ONE production sampler per launch on explicit inputs,
// one thread per record of 8 doubles, random numbers from the Philox stream of history first + i consumed exactly as the
// transport kernels consume them.  Nothing here is used by the transport path.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double test_range(const DevProblem &P, const ElecBin *B0, const MedRec &M, double eke, double rinv, double &elke, int &lelke) {
    elke = flog(eke);                                          // as estep_size()
    lelke = elec_interval(M, elke);
    const ElecBin *B = B0 + lelke;
    const double2 re = ldg2(&B->range_ep);
    const double elkei = (double)fdiv((float)(lelke + 1) - (float)M.eke0, (float)M.eke1);
    return (drange_m(B, eke, re.y, elke, elkei) + re.x) * rinv;
}

__global__ void test_sampler_kernel(const __grid_constant__ DevProblem P, int which_v, int n, const double *__restrict__ in,
                                    unsigned long long first, double *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // bits 8.. of the sampler id select a variant of the condensed-history step: 0 = the production one (msdist_b, block draws),
    // 1 = msdist_f (fp32, word-by-word draws), 2 = the fp64 msdist of the lock-step kernel -- to tell precision effects from
    // restructuring effects when a distribution test fails
    const int which = which_v & 0xff, variant = which_v >> 8;
    const double *a = in + 8 * (size_t)i;
    double *o = out + 8 * (size_t)i;
    for (int k = 0; k < 8; k++) o[k] = 0.0;
    Rng g;
    const unsigned long long hist = first + (unsigned long long)i;
    g.seed_blocks(P.seed0, P.seed1, (uint32_t)hist, (uint32_t)(hist >> 32), 0u, 0u);
    if (which == OMC_SAMPLER_DRANGE) {
        const int imed = (int)a[0], qel = (1 + (int)a[1]) / 2;
        const MedRec &M = P.med[imed];
        const ElecBin *B0 = P.ebin + (size_t)qel * P.nmed * MXEKE + imed * MXEKE;
        const double elkei = flog(a[2]), elkef = flog(a[3]);
        o[0] = drange_m(B0 + elec_interval(M, elkei), a[2], a[3], elkei, elkef);
    } else if (which == OMC_SAMPLER_ELOSS || which == OMC_SAMPLER_MSDIST) {
        const int imed = (int)a[0], iq = (int)a[1], qel = (1 + iq) / 2;
        const MedRec &M = P.med[imed];
        const ElecBin *B0 = P.ebin + (size_t)qel * P.nmed * MXEKE + imed * MXEKE;
        const double rhof = a[2], eke = a[3], rinv = 1.0 / rhof;
        double elke; int lelke;
        const double range = test_range(P, B0, M, eke, rinv, elke, lelke);
        const double tustep = a[4] * range;
        const double de = eloss_m(B0, M, rhof, rinv, tustep, range, eke, elke, lelke);
        if (which == OMC_SAMPLER_ELOSS) { o[0] = range; o[1] = de; return; }
        Part p;
        p.x = p.y = p.z = 0.0; p.u = a[5]; p.v = a[6]; p.w = a[7]; p.e = eke + RM; p.wt = 1.0; p.ir = 1; p.iq = iq;
        double xf, yf, zf, uf, vf, wf;
        uint32_t w_rfict;
        if (variant == 0) {
            o[0] = msdist_b(P, g, p, imed, qel, rhof, de, tustep, eke, xf, yf, zf, uf, vf, wf, w_rfict);
        } else {
            Rng g2;
            g2.seed(P.seed0, P.seed1, hist, 0u);
            o[0] = (variant == 1) ? msdist_f(P, g2, p, imed, qel, rhof, de, tustep, eke, xf, yf, zf, uf, vf, wf)
                                  : msdist<true>(P, g2, p, imed, qel, rhof, de, tustep, eke, xf, yf, zf, uf, vf, wf);
        }
        o[1] = xf; o[2] = yf; o[3] = zf; o[4] = uf; o[5] = vf; o[6] = wf; o[7] = de;
    } else if (which == OMC_SAMPLER_SSCAT) {
        const uint4 g0 = g.block();
        float cf, sf, cphi, sphi;
        sscat_b(P, g, (int)a[0], (int)a[1], (float)a[2], (float)a[3], (float)a[4], g0.y, cf, sf, cphi, sphi);
        o[0] = (double)cf; o[1] = (double)sf; o[2] = (double)cphi; o[3] = (double)sphi;
    } else if (which == OMC_SAMPLER_COMPTON || which == OMC_SAMPLER_MOLLER) {
        const int off = (which == OMC_SAMPLER_MOLLER) ? 1 : 0;
        Part p, q;
        p.x = p.y = p.z = 0.0; p.e = a[off]; p.u = a[off + 1]; p.v = a[off + 2]; p.w = a[off + 3]; p.wt = 1.0; p.ir = 1;
        p.iq = off ? -1 : 0;
        q = p; q.e = 0.0;
        if (which == OMC_SAMPLER_COMPTON) compton_b(g, p, q);
        else if (!moller_b(P, g, p, q, (int)a[0])) q.e = 0.0;
        o[0] = p.e; o[1] = p.u; o[2] = p.v; o[3] = p.w; o[4] = q.e; o[5] = q.u; o[6] = q.v; o[7] = q.w;
    } else if (which == OMC_SAMPLER_WOODCOCK) {
        Part p;
        p.e = a[0]; p.x = a[1]; p.y = a[2]; p.z = a[3]; p.u = a[4]; p.v = a[5]; p.w = a[6]; p.wt = 1.0; p.iq = 0;
        p.ir = 1 + find_bin(P.xb, P.isize, p.x, P.inv_dx, P.uniform_x != 0) + find_bin(P.yb, P.jsize, p.y, P.inv_dy, P.uniform_y != 0) * P.isize +
               find_bin(P.zb, P.ksize, p.z, P.inv_dz, P.uniform_z != 0) * P.ijmax;
        bool entered = false;
        unsigned nv = 0;
        int st = 0;
        for (int k = 0; k < 100000 && st == 0; k++) st = woodcock_flight(P, g, p, entered, 8, nv);
        o[0] = (st > 0) ? 1.0 : 0.0; o[1] = p.x; o[2] = p.y; o[3] = p.z; o[4] = (double)p.ir; o[5] = (double)nv;
    } else if (which == OMC_SAMPLER_ESTEP) {
        Part p;
        p.iq = (int)a[0]; p.e = a[1]; p.x = a[2]; p.y = a[3]; p.z = a[4]; p.ir = (int)a[5]; p.u = 0.0; p.v = 0.0; p.w = 1.0; p.wt = 0.0;
        EStep e;
        e.total_tstep = e.range = e.tustep = e.tperp = e.demfp = e.blccl = e.ssmfp = 0.0;
        Tally t = {0, 0, 0};
        int st = 0;
        const int cls = estep_size(P, P.endep32, g, p, e, t, st, make_int2(0, -2));    // (weight 0: a cut-off deposit adds nothing)
        o[0] = (double)cls; o[1] = e.tustep; o[2] = e.tperp; o[3] = e.range; o[4] = e.total_tstep; o[5] = e.demfp; o[6] = e.blccl; o[7] = e.ssmfp;
    }
}
void launch_test_samplers(const DevProblem &P, int which, int n, const double *in, unsigned long long first, double *out, cudaStream_t s) {
    test_sampler_kernel<<<(n + 127) / 128, 128, 0, s>>>(P, which, n, in, first, out);
}

void launch_flush(float *g32, double *g64, long long n, cudaStream_t s) {
    int blocks = (int)((n + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    flush_kernel<<<blocks, 256, 0, s>>>(g32, g64, n);
}

}  // namespace omc

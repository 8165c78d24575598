// omc_wavefront.cu -- production path: event-based particle queues (BASELINE.json north_star (a)).
//
// Particles live in HBM as structure-of-arrays queues.  One "wave" runs every live particle through
// exactly one event of its kind, with one kernel per event class so that a warp executes one code
// path:
//
//     photon_flight   P_cur  -> P_next (flight not finished) | IQ_phot (at an interaction site)
//     electron_step   E_cur  -> E_next (still travelling)    | IQ_elec (discrete interaction due)
//     photon_interact IQ_phot-> P_next, E_next               (Compton / pair / photo / Rayleigh)
//     electron_interact IQ_elec -> E_next, P_next            (brems / Moller / Bhabha / annihilation)
//     source          tops P_next (or E_next) up with new histories while any remain
//     advance         next -> cur bookkeeping (single thread)
//
// Physics = the same device functions as the lock-step kernel (omc_physics.cuh).  Differences that
// are statistically neutral: every particle owns a Philox sub-stream derived from its parent's, so
// results do not depend on scheduling; the reference's zero-length "second ustep iteration"
// (src/ompmc.c:4787 re-initialises total_tstep) is not executed, its only effect being
// one wasted random draw; with nsplit == 1 the unused survivor-index draw of photon() (:1916) is
// skipped.  Dose is scored with fp32 atomics into a chunk grid that is folded into the fp64 batch grid
// at the end of every omc_gpu_run_histories() call (north_star (d)).
#include "omc_physics.cuh"
#include "omc_kernels.h"

#ifndef OMC_WARP_AGGREGATE_DOSE
#define OMC_WARP_AGGREGATE_DOSE 0
#endif

namespace omc {

// ---- queues ----------------------------------------------------------------------------------
__device__ __forceinline__ void q_load(const PartQueue &q, unsigned i, Part &p, Rng &g, const DevProblem &P, double &aux, int &tag) {
    p.x = q.x[i]; p.y = q.y[i]; p.z = q.z[i]; p.u = q.u[i]; p.v = q.v[i]; p.w = q.w[i];
    p.e = q.e[i]; p.wt = q.wt[i]; aux = q.aux[i];
    const int2 a = q.irq[i];
    p.ir = a.x; p.iq = (int)(short)(a.y & 0xffff); tag = a.y >> 16;
    const uint4 r = q.rng[i];
    g.seed(P.seed0, P.seed1, ((unsigned long long)r.y << 32) | r.x, r.z, r.w);
}

__device__ __forceinline__ void q_store(const PartQueue &q, unsigned i, const Part &p, const Rng &g, double aux, int tag) {
    q.x[i] = p.x; q.y[i] = p.y; q.z[i] = p.z; q.u[i] = p.u; q.v[i] = p.v; q.w[i] = p.w;
    q.e[i] = p.e; q.wt[i] = p.wt; q.aux[i] = aux;
    q.irq[i] = make_int2(p.ir, (p.iq & 0xffff) | (tag << 16));
    q.rng[i] = make_uint4(g.h0, g.h1, g.stream, g.ndraws());
}

// warp-aggregated slot reservation: one atomic per (warp, queue) instead of one per lane
__device__ __forceinline__ unsigned q_reserve(unsigned *count) {
    const unsigned m = __activemask();
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(m) - 1;
    unsigned base = 0;
    if (lane == leader) base = atomicAdd(count, (unsigned)__popc(m));
    base = __shfl_sync(m, base, leader);
    return base + __popc(m & ((1u << lane) - 1u));
}

__device__ __forceinline__ void q_push(const PartQueue &q, unsigned *count, WaveCtl *ctl, const Part &p, const Rng &g, double aux,
                                       int tag) {
    const unsigned slot = q_reserve(count);
    if (slot >= q.cap) { atomicAdd(&ctl->overflow, 1u); return; }
    q_store(q, slot, p, g, aux, tag);
}

// sub-stream of a particle created by `parent` (a function of the parent's stream position only)
__device__ __forceinline__ void child_rng(const Rng &parent, Rng &c, unsigned k) {
    unsigned s = parent.stream * 0x9E3779B1u + parent.ndraws() * 0x85EBCA77u + (k + 1u) * 0xC2B2AE3Du;
    s ^= s >> 15; s *= 0x2C1B3C6Du; s ^= s >> 12; s *= 0x297A2D39u; s ^= s >> 15;
    c = parent;
    c.stream = s | 1u;          // never 0: stream 0 is the primary's
    c.blk = 0; c.pos = 4;
}

// ausgab(): fp32 chunk grid (north_star (d)); warp-aggregated when lanes hit the same voxel
struct Tally {
    unsigned ndep, nstep;
};
__device__ __forceinline__ void deposit32(const DevProblem &P, Tally &t, int ir, double en) {
    t.ndep++;
#if OMC_WARP_AGGREGATE_DOSE
    const unsigned m = __activemask();
    const unsigned peers = __match_any_sync(m, ir);
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(peers) - 1;
    float v = (float)en;
    // butterfly over the peer set
    for (unsigned rest = peers & ~(1u << leader); rest; rest &= rest - 1) {
        const int src = __ffs(rest) - 1;
        const float o = __shfl_sync(peers, v, src);
        if (lane == leader) v += o;
    }
    if (lane == leader) atomicAdd(P.endep32 + ir, v);
#else
    atomicAdd(P.endep32 + ir, (float)en);
#endif
}

enum { TAG_NONE = 0, TAG_COMPTON = 1, TAG_PAIR = 2, TAG_PHOTO = 3, TAG_RAYLEIGH = 4, TAG_BREMS = 5, TAG_MOLLER = 6, TAG_BHABHA = 7,
       TAG_ANNIH = 8, TAG_RANNIH = 9 };

// ---------------------------------------------------------------------------------------------
// photon free flight: photon(), src/ompmc.c:1884-2067 for nsplit == 1
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) photon_flight_kernel(const __grid_constant__ DevProblem P, WaveCtl *ctl, PartQueue cur,
                                                            PartQueue next, PartQueue iq, int max_cross) {
    const unsigned n = ctl->n_p_cur;
    Tally t = {0, 0};
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        Part p; Rng g; double dpmfp; int tag;
        q_load(cur, i, p, g, P, dpmfp, tag);
        RegionRec R = load_region(P, p.ir);
        if (dpmfp < 0.0) {                                     // fresh photon: cut-off test + number of mfp
            if (p.e <= R.pcut || p.wt == 0) { deposit32(P, t, p.ir, p.wt * p.e); continue; }
            const double r = g.next();
            dpmfp = -log(1.0 - r);                             // eta' = 1 - r  (:1905-1932 with nsplit = 1)
        }
        const double gle = log(p.e);
        int imed = R.med, medc = -2;
        double gmfpr0 = 0.0, cohfac = 0.0, gmfp = 0.0;
        int irl = p.ir;
        bool at_site = false, gone = false;
        for (int k = 0; k < max_cross; k++) {                  // voxel-to-voxel march, :1951-2019
            double tstep;
            if (imed != -1) {
                if (imed != medc) {                            // (imed, gle) -> table values, reused while the medium stays
                    const MedRec &M = P.med[imed];
                    const int lgle = (int)(gle * M.ge1 + M.ge0) - 1;
                    const PhotBin *B = P.phot + imed * MXGE + lgle;
                    const double2 a = __ldg(reinterpret_cast<const double2 *>(&B->gmfp1));
                    const double2 b = __ldg(reinterpret_cast<const double2 *>(&B->cohe1));
                    gmfpr0 = pwl(gle, a.x, a.y);
                    cohfac = pwl(gle, b.x, b.y);
                    medc = imed;
                }
                gmfp = gmfpr0 / R.rhof;
                gmfp *= cohfac;
                tstep = gmfp * dpmfp;
            } else {
                tstep = 1.0E8;
            }
            int irnew = irl, idisc = 0;
            double ustep = tstep;
            howfar(P, p, idisc, irnew, ustep);
            t.nstep++;
            p.x += ustep * p.u; p.y += ustep * p.v; p.z += ustep * p.w;
            if (idisc > 0) { gone = true; break; }
            if (imed != -1) dpmfp = fmax(0.0, dpmfp - ustep / gmfp);
            if (irnew != irl) {
                p.ir = irnew; irl = irnew;
                R = load_region(P, irl);
                imed = R.med;
            }
            if (imed != -1 && dpmfp <= 1.0E-05) { at_site = true; break; }
        }
        if (gone) continue;
        if (!at_site) { q_push(next, &ctl->n_p_next, ctl, p, g, dpmfp, TAG_NONE); continue; }
        if (imed != medc) {                                    // site reached right after a medium change
            const MedRec &M = P.med[imed];
            const int lgle = (int)(gle * M.ge1 + M.ge0) - 1;
            const PhotBin *B = P.phot + imed * MXGE + lgle;
            cohfac = pwl(gle, __ldg(&B->cohe1), __ldg(&B->cohe0));
        }
        double r = g.next();                                   // :2027-2067
        if (r <= 1.0 - cohfac) {
            tag = TAG_RAYLEIGH;
        } else {
            const MedRec &M = P.med[imed];
            const int lgle = (int)(gle * M.ge1 + M.ge0) - 1;
            const PhotBin *B = P.phot + imed * MXGE + lgle;
            r = g.next();
            const double gbr1 = pwl(gle, __ldg(&B->gbr11), __ldg(&B->gbr10));
            if (r <= gbr1 && p.e > 2.0 * RM) tag = TAG_PAIR;
            else {
                const double gbr2 = pwl(gle, __ldg(&B->gbr21), __ldg(&B->gbr20));
                tag = (r < gbr2) ? TAG_COMPTON : TAG_PHOTO;
            }
        }
        q_push(iq, &ctl->n_iq_phot, ctl, p, g, -1.0, tag);
    }
    if (t.nstep) atomicAdd(&P.counters->photon_steps, (unsigned long long)t.nstep);
    if (t.ndep) atomicAdd(&P.counters->deposits, (unsigned long long)t.ndep);
}

__global__ void __launch_bounds__(256) photon_interact_kernel(const __grid_constant__ DevProblem P, WaveCtl *ctl, PartQueue iq,
                                                              PartQueue pnext, PartQueue enext) {
    const unsigned n = ctl->n_iq_phot;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        Part p, q; Rng g, gq; double aux; int tag;
        q_load(iq, i, p, g, P, aux, tag);
        const RegionRec R = load_region(P, p.ir);
        const int imed = R.med;
        if (tag == TAG_COMPTON) {
            compton(g, p, q);
            child_rng(g, gq, 0);
            q_push(pnext, &ctl->n_p_next, ctl, p, g, -1.0, TAG_NONE);
            q_push(enext, &ctl->n_e_next, ctl, q, gq, 0.0, TAG_NONE);
        } else if (tag == TAG_PAIR) {
            pair(P, g, p, q, imed);
            child_rng(g, gq, 0);
            q_push(enext, &ctl->n_e_next, ctl, p, g, 0.0, TAG_NONE);
            q_push(enext, &ctl->n_e_next, ctl, q, gq, 0.0, TAG_NONE);
        } else if (tag == TAG_PHOTO) {
            photo(g, p, R.ecut);
            q_push(enext, &ctl->n_e_next, ctl, p, g, 0.0, TAG_NONE);
        } else {                                               // Rayleigh: direction change only
            const MedRec &M = P.med[imed];
            const double gle = log(p.e);
            const int lgle = (int)(gle * M.ge1 + M.ge0) - 1;
            const PhotBin *B = P.phot + imed * MXGE + lgle;
            rayleigh(P, g, p, pwl(gle, __ldg(&B->pmax1), __ldg(&B->pmax0)), p.e);
            q_push(pnext, &ctl->n_p_next, ctl, p, g, -1.0, TAG_NONE);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// one condensed-history / boundary-crossing electron step: the tstep/ustep loops of electron(),
// src/ompmc.c:4694-5372, one real step per call.  Returns a TAG_* (interaction due), 0 = keep
// travelling, -1 = particle finished.
// ---------------------------------------------------------------------------------------------
__device__ int electron_iter(const DevProblem &P, Rng &g, Part &p, Tally &t) {
    RegionRec R = load_region(P, p.ir);
    int imed = R.med;
    const int iq = p.iq, qel = (1 + iq) / 2;
    double eie = p.e;
    t.nstep++;
    if (eie <= R.ecut) {                                       // :4665-4687
        deposit32(P, t, p.ir, p.wt * (eie - RM));
        return (iq > 0) ? TAG_RANNIH : -1;
    }
    double ustep, tustep = 0.0, tvstep, de = 0.0, range = 0.0, sig0 = 0.0, demfp = 0.0, total_tstep = 0.0, rhof = R.rhof;
    double eke = eie - RM, elke = 0.0;
    int lelke = 0;
    bool call_howfar, do_single = false, called_msdist = false;
    double xf = 0, yf = 0, zf = 0, uf = 0, vf = 0, wf = 0;
    const ElecBin *B0 = nullptr;
    if (imed == -1) {                                          // vacuum (region 0 = outside): :4815-4821
        ustep = 10.0E8; tustep = ustep; call_howfar = true;
    } else {
        const MedRec &M = P.med[imed];
        B0 = P.ebin + (size_t)qel * P.nmed * MXEKE + imed * MXEKE;
        double r = g.next();
        if (r == 0.0) r = 1.0E-30;
        demfp = fmax(-log(r), 1.0E-5);
        elke = log(eke);
        lelke = elec_interval(M, elke);
        const ElecBin *B = B0 + lelke;
        const double dedx0 = pwl(elke, __ldg(&B->dedx1), __ldg(&B->dedx0));
        if (M.sig_ismonotone[qel]) sig0 = pwl(elke, __ldg(&B->sig1), __ldg(&B->sig0)) / dedx0;
        else sig0 = (iq < 0) ? M.esig_e : M.psig_e;
        double tstep;
        if (sig0 <= 0.0) {
            tstep = 10.0E8; sig0 = 1.0E-15;
        } else {
            const double ekef = eke - demfp / sig0;
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
            total_tstep = tstep;
            tstep = total_tstep / rhof;
        }
        const double dedx = rhof * dedx0;
        const double tmxs = pwl(elke, __ldg(&B->tmxs1), __ldg(&B->tmxs0)) / rhof;
        {
            const double ekei = __ldg(&B->e_array), elkei = (lelke + 1 - M.eke0) / M.eke1;
            range = (drange(B, eke, ekei, elke, elkei) + __ldg(&B->range_ep)) / rhof;
        }
        tustep = fmin(fmin(tstep, tmxs), range);
        const double tperp = hownear(P, p);
        double blccl = rhof * M.blcc;
        const double xccl = rhof * M.xcc;
        const double p2 = eke * (eke + 2.0 * RM);
        const double beta2 = p2 / (p2 + (RM * RM));
        const double etap = pwl(elke, __ldg(&B->eta1), __ldg(&B->eta0));
        const double ms_corr = pwl(elke, __ldg(&B->blcce1), __ldg(&B->blcce0));
        blccl = blccl / etap / (1.0 + 0.25 * etap * xccl / blccl / p2) * ms_corr;
        const double ssmfp = beta2 / blccl;
        const double skindepth = 3 * ssmfp;
        tustep = fmin(tustep, fmax(tperp, skindepth));
        if ((tustep <= tperp) && (tustep > skindepth)) {       // condensed-history step, :4973-4996
            call_howfar = false; called_msdist = true;
            de = eloss(B0, M, rhof, tustep, range, eke, elke, lelke);
            ustep = msdist(P, g, p, imed, qel, rhof, de, tustep, eke, xf, yf, zf, uf, vf, wf);
        } else {                                               // exact boundary crossing, :4997-5057
            r = g.next();
            if (r < 1.0E-30) r = 1.0E-30;
            const double lambda = (-1.0) * log(1.0 - r);
            double lambda_max = 0.5 * blccl * RM / dedx;
            lambda_max *= (eke / RM + 1.0) * (eke / RM + 1.0) * (eke / RM + 1.0);
            if (!(lambda >= 0.0 && lambda_max > 0.0)) return -1;   // Q8: dropped without deposit
            const double tuss = (lambda < lambda_max) ? lambda * ssmfp * (1.0 - 0.5 * lambda / lambda_max) : 0.5 * lambda * ssmfp;
            if (tuss < tustep) { tustep = tuss; do_single = true; }
            ustep = tustep;
            call_howfar = !(ustep < tperp);
        }
    }
    const int irl = p.ir;
    int irnew = irl, idisc = 0;
    if (call_howfar) howfar(P, p, idisc, irnew, ustep);
    if (idisc > 0) {                                           // :5061-5088 (no annihilation quanta: edep > eie)
        deposit32(P, t, p.ir, p.wt * ((iq > 0) ? p.e + RM : p.e - RM));
        return -1;
    }
    if (ustep < 0) ustep = 0.0;
    if (ustep == 0.0 || imed == -1) {                          // :5097-5146
        if (ustep != 0.0) { p.x += p.u * ustep; p.y += p.v * ustep; p.z += p.w * ustep; }
        if (irnew != irl) {
            p.ir = irnew;
            R = load_region(P, irnew);
        }
        if (eie <= R.ecut) {
            deposit32(P, t, p.ir, p.wt * (eie - RM));
            return (iq > 0) ? TAG_RANNIH : -1;
        }
        return 0;
    }
    const MedRec &M = P.med[imed];
    if (call_howfar) {
        tvstep = ustep;
        if (tvstep != tustep) do_single = false;
        de = eloss(B0, M, rhof, tvstep, range, eke, elke, lelke);
    } else {
        tvstep = tustep;
        if (!called_msdist) de = eloss(B0, M, rhof, tvstep, range, eke, elke, lelke);
    }
    if (!called_msdist) {
        xf = p.x + p.u * ustep; yf = p.y + p.v * ustep; zf = p.z + p.w * ustep;
        if (do_single) {                                       // :5180-5207
            const double ekems = fmax(eke - de, R.ecut - RM);
            const double p2 = ekems * (ekems + 2.0 * RM);
            const double beta2 = p2 / (p2 + (RM * RM));
            double chia2 = M.xcc / (4.0 * M.blcc * p2);
            const double elkems = log(ekems);
            const int lelkems = elec_interval(M, elkems);
            chia2 *= pwl(elkems, __ldg(&B0[lelkems].eta1), __ldg(&B0[lelkems].eta0));
            double costhe, sinthe;
            sscat(P, g, imed, qel, chia2, elkems, beta2, costhe, sinthe);
            Frame fr;
            uphi21(g, fr, costhe, sinthe, p);
        }
        uf = p.u; vf = p.v; wf = p.w;
    }
    deposit32(P, t, p.ir, p.wt * de);                          // :5245
    p.x = xf; p.y = yf; p.z = zf; p.u = uf; p.v = vf; p.w = wf;
    eie -= de;
    p.e = eie;
    if (irnew == irl && eie <= R.ecut) {
        deposit32(P, t, p.ir, p.wt * (eie - RM));
        return (iq > 0) ? TAG_RANNIH : -1;
    }
    eke = eie - RM;
    elke = log(eke);
    lelke = elec_interval(M, elke);
    int imed_new = imed;
    if (irnew != irl) {
        p.ir = irnew;
        R = load_region(P, irnew);
        imed_new = R.med;
    }
    if (eie <= R.ecut) {
        deposit32(P, t, p.ir, p.wt * (eie - RM));
        return (iq > 0) ? TAG_RANNIH : -1;
    }
    if (imed_new != imed) return 0;                            // new medium: resample from the top
    demfp -= de * sig0;
    total_tstep -= tvstep * rhof;
    if (total_tstep < 1.0E-9) demfp = 0.0;
    if (demfp >= 1.0E-5) return 0;   // interaction point not reached (the reference then burns a zero step and resamples)
    // fictitious cross-section rejection, :5354-5372
    const ElecBin *B = B0 + lelke;
    const double sigf = pwl(elke, __ldg(&B->sig1), __ldg(&B->sig0)) / pwl(elke, __ldg(&B->dedx1), __ldg(&B->dedx0));
    const double rfict = g.next();
    if (rfict >= sigf / sig0) return 0;
    const double br1 = pwl(elke, __ldg(&B->bra1), __ldg(&B->bra0));   // :5375-5429
    const double r = g.next();
    if (iq < 0) {
        if (r <= br1) return TAG_BREMS;
        if (p.e <= M.thmoll) return (br1 <= 0) ? 0 : TAG_BREMS;
        return TAG_MOLLER;
    }
    if (r < br1) return TAG_BREMS;
    const double pbr2 = pwl(elke, __ldg(&B->brb1), __ldg(&B->brb0));
    return (r < pbr2) ? TAG_BHABHA : TAG_ANNIH;
}

__global__ void __launch_bounds__(128) electron_step_kernel(const __grid_constant__ DevProblem P, WaveCtl *ctl, PartQueue cur,
                                                            PartQueue next, PartQueue iq, int iters) {
    const unsigned n = ctl->n_e_cur;
    Tally t = {0, 0};
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        Part p; Rng g; double aux; int tag;
        q_load(cur, i, p, g, P, aux, tag);
        int st = 0;
        for (int k = 0; k < iters && st == 0; k++) st = electron_iter(P, g, p, t);
        if (st == 0) q_push(next, &ctl->n_e_next, ctl, p, g, 0.0, TAG_NONE);
        else if (st > 0) q_push(iq, &ctl->n_iq_elec, ctl, p, g, 0.0, st);
    }
    if (t.nstep) atomicAdd(&P.counters->electron_steps, (unsigned long long)t.nstep);
    if (t.ndep) atomicAdd(&P.counters->deposits, (unsigned long long)t.ndep);
}

__global__ void __launch_bounds__(256) electron_interact_kernel(const __grid_constant__ DevProblem P, WaveCtl *ctl, PartQueue iq,
                                                                PartQueue enext, PartQueue pnext) {
    const unsigned n = ctl->n_iq_elec;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        Part p, q; Rng g, gq; double aux; int tag;
        q_load(iq, i, p, g, P, aux, tag);
        const int imed = region_med(P, p.ir);
        if (tag == TAG_MOLLER) {
            const bool created = moller(P, g, p, q, imed);
            q_push(enext, &ctl->n_e_next, ctl, p, g, 0.0, TAG_NONE);
            if (created) {
                child_rng(g, gq, 0);
                q_push(enext, &ctl->n_e_next, ctl, q, gq, 0.0, TAG_NONE);
            }
        } else if (tag == TAG_BREMS) {
            brems(P, g, p, q, imed, 1);
            child_rng(g, gq, 0);
            q_push(enext, &ctl->n_e_next, ctl, p, g, 0.0, TAG_NONE);
            q_push(pnext, &ctl->n_p_next, ctl, q, gq, -1.0, TAG_NONE);
        } else if (tag == TAG_BHABHA) {
            bhabha(P, g, p, q, imed);
            child_rng(g, gq, 0);
            q_push(enext, &ctl->n_e_next, ctl, p, g, 0.0, TAG_NONE);
            q_push(enext, &ctl->n_e_next, ctl, q, gq, 0.0, TAG_NONE);
        } else {                                               // annihilation in flight / at rest
            if (tag == TAG_ANNIH) annih(g, p, q, 1);
            else rannih(g, p, q, 1);
            child_rng(g, gq, 0);
            q_push(pnext, &ctl->n_p_next, ctl, p, g, -1.0, TAG_NONE);
            q_push(pnext, &ctl->n_p_next, ctl, q, gq, -1.0, TAG_NONE);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// source + bookkeeping
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned n_inject(const WaveCtl *c) {
    const unsigned live = c->n_p_next + c->n_e_next;
    const unsigned long long left = c->hist_end - c->hist_next;
    unsigned room = (live < c->target) ? c->target - live : 0u;
    return (unsigned)(left < (unsigned long long)room ? left : (unsigned long long)room);
}

// initHistory() for the next n_inject(ctl) history ids, appended to the photon or electron queue
__global__ void __launch_bounds__(256) source_kernel(const __grid_constant__ DevProblem P, const WaveCtl *ctl, PartQueue pnext,
                                                     PartQueue enext) {
    const unsigned n = n_inject(ctl);
    const bool photons = (P.src.charge == 0);
    const PartQueue &q = photons ? pnext : enext;
    const unsigned base = photons ? ctl->n_p_next : ctl->n_e_next;
    double ensrc = 0.0;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        Rng g;
        g.seed(P.seed0, P.seed1, ctl->hist_next + i, 0u);
        Part p;
        ensrc += init_history_dosxyz(P, g, p);
        if (base + i < q.cap) q_store(q, base + i, p, g, -1.0, TAG_NONE);
    }
    for (int o = 16; o > 0; o >>= 1) ensrc += __shfl_xor_sync(0xffffffffu, ensrc, o);
    if ((threadIdx.x & 31) == 0 && ensrc != 0.0) atomicAdd(P.ensrc, ensrc);
}

__global__ void advance_kernel(const __grid_constant__ DevProblem P, WaveCtl *ctl) {
    if (blockIdx.x || threadIdx.x) return;
    const unsigned n = n_inject(ctl);
    unsigned np = ctl->n_p_next, ne = ctl->n_e_next;
    if (P.src.charge == 0) np += n; else ne += n;
    if (np > ctl->cap_p || ne > ctl->cap_e) { ctl->overflow += 1; np = min(np, ctl->cap_p); ne = min(ne, ctl->cap_e); }
    ctl->hist_next += n;
    P.counters->histories += n;
    ctl->n_p_cur = np; ctl->n_e_cur = ne;
    ctl->n_p_next = 0; ctl->n_e_next = 0; ctl->n_iq_phot = 0; ctl->n_iq_elec = 0;
    ctl->waves += 1;
    ctl->live = np + ne;
    if (P.counters && ctl->overflow) P.counters->errors = ctl->overflow;
}

// fold the fp32 chunk grid into the fp64 batch grid
__global__ void flush_kernel(float *__restrict__ g32, double *__restrict__ g64, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float v = g32[i];
        if (v != 0.0f) { g64[i] += (double)v; g32[i] = 0.0f; }
    }
}

// ---- host-side launchers ----------------------------------------------------------------------
int wave_blocks_per_sm(int which) {
    int n = 0;
    cudaError_t e;
    if (which == 0) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, electron_step_kernel, 128, 0);
    else e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, photon_flight_kernel, 256, 0);
    return (e == cudaSuccess && n > 0) ? n : 1;
}

// One wave.  `parity` selects which half of the double-buffered queues is "cur".
void launch_wave(const DevProblem &P, WaveCtl *ctl, const WaveQueues &Q, int parity, const WaveLaunch &L, cudaStream_t s) {
    const PartQueue &pc = Q.p[parity], &pn = Q.p[parity ^ 1], &ec = Q.e[parity], &en = Q.e[parity ^ 1];
    photon_flight_kernel<<<L.blocks_phot, 256, 0, s>>>(P, ctl, pc, pn, Q.iq_phot, L.max_cross);
    electron_step_kernel<<<L.blocks_elec, 128, 0, s>>>(P, ctl, ec, en, Q.iq_elec, L.electron_iters);
    photon_interact_kernel<<<L.blocks_int, 256, 0, s>>>(P, ctl, Q.iq_phot, pn, en);
    electron_interact_kernel<<<L.blocks_int, 256, 0, s>>>(P, ctl, Q.iq_elec, en, pn);
    source_kernel<<<L.blocks_int, 256, 0, s>>>(P, ctl, pn, en);
    advance_kernel<<<1, 32, 0, s>>>(P, ctl);
}

void launch_flush(float *g32, double *g64, long long n, cudaStream_t s) {
    int blocks = (int)((n + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    flush_kernel<<<blocks, 256, 0, s>>>(g32, g64, n);
}

}  // namespace omc

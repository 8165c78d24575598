// omc_lockstep.cu -- "lock-step" kernel: one history per thread, LIFO particle stack, and exactly
// the reference's order of random draws, so that a history run here consumes the same Philox
// stream the instrumented reference / oracle consume on the CPU (oracle/ref_harness.c).  This is
// the on-device correctness anchor; the production path is the wavefront
// kernels in omc_wavefront.cu, validated statistically against this one.
//
// Compiled with -fmad=false so that a*b+c rounds twice as it does in the reference's x86-64 build.
#include "omc_physics.cuh"
#include "omc_kernels.h"
#include "omc_format.cuh"

namespace omc {

// LIFO stack of one thread, slot-major in HBM: slot i of thread t lives at base[i*stride + t], so the
// threads of a warp that are at the same depth touch neighbouring records.
struct StackRef {
    Part *base;
    size_t stride;
    __device__ __forceinline__ Part &operator[](int i) const { return base[(size_t)i * stride]; }
};

struct HistCtx {
    StackRef s;
    int np, npold, depth;
    Rng g;
    unsigned ndeposit, flags;
    unsigned nphot_steps, nelec_steps;
    double edep_sum;
};

// ausgab(), omc_dosxyz.c:683-694: fp64 atomic on the batch grid
__device__ __forceinline__ void deposit(const DevProblem &P, HistCtx &c, const Part &p, double edep) {
    const double en = p.wt * edep;
    c.ndeposit++;
    c.edep_sum += en;
    atomicAdd(P.endep + p.ir, en);
}

__device__ __forceinline__ bool push_ok(HistCtx &c, int np_new) {
    if (np_new >= c.depth) { c.flags |= 1u; return false; }
    return true;
}

// photon(), src/ompmc.c:1849-2121, incl. uniform splitting (nsplit) and roulette of scattered photons
__device__ void photon_ls(const DevProblem &P, HistCtx &c) {
    StackRef s = c.s;
    Rng &g = c.g;
    int np = c.np;
    Part p = s[np];
    const int nsplit = P.nsplit;
    {
        const RegionRec R = load_region(P, p.ir);
        if (p.e <= R.pcut || p.wt == 0) {
            deposit(P, c, p, p.e);
            c.np -= 1;
            return;
        }
    }
    double eig = p.e, gle = log(eig), cohfac = 0.0, gmfp = 0.0;
    int imed = 0, lgle_last = 0;

    for (;;) {                                                  // start_mfp_loop
        double r = g.next();
        r /= (double)nsplit;
        const double d_eta = 1.0 / (double)nsplit;
        double eta_prime = 1.0 - r + d_eta;
        Part save = p;
        save.wt = p.wt / (double)nsplit;
        save.iq = 0;
        np -= 1;
        r = g.next();
        const int i_survive = (int)(r * nsplit);
        double dpmfp_old = 0.0;

        for (int isplit = 0; isplit < nsplit; isplit++) {
            eta_prime -= d_eta;
            if (eta_prime <= 0.0) break;
            double dpmfp = -log(eta_prime) - dpmfp_old;
            dpmfp_old += dpmfp;
            np += 1; c.np = np;
            if (!push_ok(c, np)) { c.np = -1; return; }         // stack overflow: abandon the history, flagged
            p = save;
            int irl = p.ir, irold = irl;
            RegionRec R = load_region(P, irl);
            imed = R.med;
            bool left = false, ptrans = true;
            do {                                                // voxel-to-voxel march, :1951-2019
                double tstep;
                if (imed != -1) {
                    const MedRec &M = P.med[imed];
                    const int lgle = (int)(gle * M.ge1 + M.ge0) - 1;
                    const PhotBin *B = P.phot + imed * MXGE + lgle;
                    const double2 a = __ldg(reinterpret_cast<const double2 *>(&B->gmfp1));
                    const double2 b = __ldg(reinterpret_cast<const double2 *>(&B->cohe1));
                    const double gmfpr0 = pwl(gle, a.x, a.y);
                    gmfp = gmfpr0 / R.rhof;
                    cohfac = pwl(gle, b.x, b.y);
                    gmfp *= cohfac;
                    tstep = gmfp * dpmfp;
                    lgle_last = lgle;
                } else {
                    tstep = 1.0E8;
                }
                int irnew = irl, idisc = 0;
                double ustep = tstep;
                howfar(P, p, idisc, irnew, ustep);
                c.nphot_steps++;
                p.x += ustep * p.u; p.y += ustep * p.v; p.z += ustep * p.w;
                if (idisc > 0) {
                    np -= 1; c.np = np;
                    if (np < 0) return;
                    left = true;
                    break;
                }
                if (imed != -1) dpmfp = fmax(0.0, dpmfp - ustep / gmfp);
                if (irnew != irold) {
                    p.ir = irnew; irl = irnew; irold = irnew;
                    R = load_region(P, irl);
                    imed = R.med;
                }
                if (imed != -1 && dpmfp <= 1.0E-05) ptrans = false;
            } while (ptrans);
            if (left) break;

            save.x = p.x; save.y = p.y; save.z = p.z; save.ir = p.ir;

            // table values at the site: index = (medium of the CURRENT region, lgle of the last march
            // iteration) exactly as the reference's pwlfEval(imed*MXGE + lgle, ...) calls (:1107, :2046, :2054)
            const PhotBin *Bs = P.phot + imed * MXGE + lgle_last;
            r = g.next();                                       // Rayleigh? :2027-2040
            if (r <= 1.0 - cohfac) {
                if (isplit != i_survive) { np -= 1; c.np = np; continue; }
                p.wt *= nsplit;
                rayleigh(P, g, p, pwl(gle, __ldg(&Bs->pmax1), __ldg(&Bs->pmax0)), eig);
                s[np] = p;
                continue;
            }
            r = g.next();
            Part q;
            bool created = false;
            c.npold = np;
            const double gbr1 = pwl(gle, __ldg(&Bs->gbr11), __ldg(&Bs->gbr10));
            if (r <= gbr1 && eig > 2.0 * RM) {
                pair(P, g, p, q, imed); created = true;
            } else {
                const double gbr2 = pwl(gle, __ldg(&Bs->gbr21), __ldg(&Bs->gbr20));
                if (r < gbr2) { compton(g, p, q); created = true; }
                else photo(g, p, R.ecut);
            }
            s[np] = p;
            if (created) {
                if (!push_ok(c, np + 1)) { c.np = -1; return; }
                np += 1;
                s[np] = q;
            }
            // keep the scattered photons of the chosen copy only, :2072-2093
            int ip = c.npold;
            do {
                if (s[ip].iq == 0) {
                    if (isplit != i_survive) {
                        if (ip < np) {
                            Part t = s[ip];
                            const Part top = s[np];
                            t.e = top.e; t.iq = top.iq; t.u = top.u; t.v = top.v; t.w = top.w; t.wt = top.wt;
                            s[ip] = t;
                        }
                        np -= 1;
                    } else {
                        s[ip].wt *= nsplit;
                        ip += 1;
                    }
                } else {
                    ip += 1;
                }
            } while (ip <= np);
            c.np = np;
        }
        // end_mfp_loop, :2095-2118
        if (np < 0) return;
        p = s[np];
        if (p.iq != 0) return;
        eig = p.e;
        const RegionRec R = load_region(P, p.ir);
        if (eig <= R.pcut) {
            deposit(P, c, p, eig);
            np -= 1; c.np = np;
            return;
        }
        gle = log(eig);
    }
}

// cut-off exits of electron(): deposit, e+ -> rannih(), else pop (src/ompmc.c:4665-4687 and 4 more)
__device__ __forceinline__ void electron_end(const DevProblem &P, HistCtx &c, Part &p, double edep, double eie) {
    deposit(P, c, p, edep);
    if (p.iq > 0 && edep < eie) {
        Part q;
        rannih(c.g, p, q, P.nsplit);
        const int np = c.np;
        c.s[np] = p;
        if (!push_ok(c, np + 1)) { c.np = -1; return; }
        c.s[np + 1] = q;
        c.np = np + 1;
        return;
    }
    c.np -= 1;
}

// electron(), src/ompmc.c:4648-5433
__device__ void electron_ls(const DevProblem &P, HistCtx &c) {
    Rng &g = c.g;
    const int np = c.np;
    Part p = c.s[np];
    int irl = p.ir;
    RegionRec R = load_region(P, irl);
    int imed = R.med;
    double rhof = R.rhof, edep = 0.0;
    double eie = p.e;
    const int iq = p.iq, qel = (1 + iq) / 2;
    int medold = imed;
    double r;

    if (eie <= R.ecut) { electron_end(P, c, p, p.e - RM, eie); return; }

    double elke = 0.0;
    int lelke = 0;
    double sigratio = 0.0, rfict = 0.0;
    const ElecBin *Bq = P.ebin + (size_t)qel * P.nmed * MXEKE;

    do {  // tstep loop, :4694
        bool compute_tstep = true;
        double eke = eie - RM, demfp = 0.0, sig0 = 0.0, ustep = 0.0;
        if (imed != -1) {
            const MedRec &M = P.med[imed];
            r = g.next();
            if (r == 0.0) r = 1.0E-30;
            demfp = fmax(-log(r), 1.0E-5);
            elke = log(eke);
            lelke = elec_interval(M, elke);
            if (M.sig_ismonotone[qel]) {
                const ElecBin *B = Bq + imed * MXEKE + lelke;
                sig0 = pwl(elke, __ldg(&B->sig1), __ldg(&B->sig0));
                const double dedx0 = pwl(elke, __ldg(&B->dedx1), __ldg(&B->dedx0));
                sig0 /= dedx0;
            } else {
                sig0 = (iq < 0) ? M.esig_e : M.psig_e;
            }
        }
        do {  // ustep loop, :4765
            bool call_howfar = false, do_single = false, called_msdist = false;
            double tstep = 0, tustep = 0, ekef, ekei, elkei, tuss, range = 0, p2, beta2, etap, tvstep = 0, de = 0;
            double total_tstep = 0.0;   // re-initialised every iteration, as in src/ompmc.c:4787
            double xf = 0, yf = 0, zf = 0, uf = 0, vf = 0, wf = 0;
            c.nelec_steps++;
            if (imed == -1) {
                tstep = 10.0E8; ustep = tstep; tustep = ustep; call_howfar = true;
            } else {
                const MedRec &M = P.med[imed];
                const ElecBin *B0 = Bq + imed * MXEKE;
                const ElecBin *B = B0 + lelke;
                rhof = R.rhof;
                if (sig0 <= 0.0) {
                    tstep = 10.0E8; sig0 = 1.0E-15;
                } else {
                    if (compute_tstep) {
                        const double total_de = demfp / sig0;
                        ekef = eke - total_de;
                        if (ekef <= __ldg(&B0[0].e_array)) {
                            tstep = 10.0E8;
                        } else {
                            const double elkef = log(ekef);
                            const int lelkef = elec_interval(M, elkef);
                            if (lelkef == lelke) {
                                tstep = drange(B, eke, ekef, elke, elkef);
                            } else {
                                ekei = __ldg(&B->e_array);
                                elkei = (lelke + 1 - M.eke0) / M.eke1;
                                tuss = drange(B, eke, ekei, elke, elkei);
                                ekei = __ldg(&B0[lelkef + 1].e_array);
                                elkei = ((lelkef + 2) - M.eke0) / M.eke1;
                                tstep = drange(B0 + lelkef, ekei, ekef, elkei, elkef);
                                tstep += tuss + __ldg(&B->range_ep) - __ldg(&B0[lelkef + 1].range_ep);
                            }
                        }
                        total_tstep = tstep;
                        compute_tstep = false;
                    }
                    tstep = total_tstep / rhof;
                }
                const double dedx0 = pwl(elke, __ldg(&B->dedx1), __ldg(&B->dedx0));
                const double dedx = rhof * dedx0;
                double tmxs = pwl(elke, __ldg(&B->tmxs1), __ldg(&B->tmxs0));
                tmxs /= rhof;
                ekei = __ldg(&B->e_array);
                elkei = (lelke + 1 - M.eke0) / M.eke1;
                range = drange(B, eke, ekei, elke, elkei);
                range += __ldg(&B->range_ep);
                range /= rhof;
                tustep = fmin(fmin(tstep, tmxs), range);
                const double tperp = hownear(P, p);
                double blccl = rhof * M.blcc;
                const double xccl = rhof * M.xcc;
                p2 = eke * (eke + 2.0 * RM);
                beta2 = p2 / (p2 + (RM * RM));
                etap = pwl(elke, __ldg(&B->eta1), __ldg(&B->eta0));
                const double ms_corr = pwl(elke, __ldg(&B->blcce1), __ldg(&B->blcce0));
                blccl = blccl / etap / (1.0 + 0.25 * etap * xccl / blccl / p2) * ms_corr;
                const double ssmfp = beta2 / blccl;
                const double skindepth = 3 * ssmfp;
                tustep = fmin(tustep, fmax(tperp, skindepth));
                if ((tustep <= tperp) && (tustep > skindepth)) {
                    call_howfar = false; do_single = false; called_msdist = true;
                    de = eloss(B0, M, rhof, tustep, range, eke, elke, lelke);
                    tvstep = tustep;
                    ustep = msdist(P, g, p, imed, qel, rhof, de, tustep, eke, xf, yf, zf, uf, vf, wf);
                } else {
                    called_msdist = false;
                    r = g.next();
                    if (r < 1.0E-30) r = 1.0E-30;
                    const double lambda = (-1.0) * log(1.0 - r);
                    double lambda_max = 0.5 * blccl * RM / dedx;
                    lambda_max *= (eke / RM + 1.0) * (eke / RM + 1.0) * (eke / RM + 1.0);
                    if (lambda >= 0.0 && lambda_max > 0.0) {
                        if (lambda < lambda_max) tuss = lambda * ssmfp * (1.0 - 0.5 * lambda / lambda_max);
                        else tuss = 0.5 * lambda * ssmfp;
                        if (tuss < tustep) { tustep = tuss; do_single = true; }
                        else do_single = false;
                    } else {                                    // Q8: particle dropped without deposit
                        c.flags |= 2u;
                        c.np -= 1;
                        return;
                    }
                    ustep = tustep;
                    call_howfar = !(ustep < tperp);
                }
            }
            int irold = p.ir, irnew = p.ir, idisc = 0;
            if (call_howfar) howfar(P, p, idisc, irnew, ustep);
            if (idisc > 0) {                                    // :5061-5088
                edep = (iq > 0) ? p.e + RM : p.e - RM;
                electron_end(P, c, p, edep, eie);
                return;
            }
            if (ustep < 0) ustep = 0.0;
            double vstep;
            if (ustep == 0.0 || imed == -1) {                   // :5097-5146
                if (ustep != 0.0) {
                    vstep = ustep; tvstep = vstep;
                    p.x += p.u * vstep; p.y += p.v * vstep; p.z += p.w * vstep;
                }
                if (irnew != irold) {
                    p.ir = irnew; irl = irnew;
                    R = load_region(P, irl);
                    imed = R.med;
                }
                if (eie <= R.ecut) { electron_end(P, c, p, p.e - RM, eie); return; }
                break;
            }
            vstep = ustep;
            const MedRec &M = P.med[imed];
            const ElecBin *B0 = Bq + imed * MXEKE;
            if (call_howfar) {
                tvstep = vstep;
                if (tvstep != tustep) do_single = false;
                de = eloss(B0, M, rhof, tvstep, range, eke, elke, lelke);
            } else {
                tvstep = tustep;
                if (!called_msdist) de = eloss(B0, M, rhof, tvstep, range, eke, elke, lelke);
            }
            edep = de;
            ekef = eke - de;
            if (!called_msdist) {
                double sinthe = 0.0, costhe = 1.0;
                if (do_single) {
                    const double ekems = fmax(ekef, R.ecut - RM);
                    p2 = ekems * (ekems + 2.0 * RM);
                    beta2 = p2 / (p2 + (RM * RM));
                    double chia2 = M.xcc / (4.0 * M.blcc * p2);
                    const double elkems = log(ekems);
                    const int lelkems = elec_interval(M, elkems);
                    etap = pwl(elkems, __ldg(&B0[lelkems].eta1), __ldg(&B0[lelkems].eta0));
                    chia2 *= etap;
                    sscat(P, g, imed, qel, chia2, elkems, beta2, costhe, sinthe);
                }
                xf = p.x + p.u * vstep; yf = p.y + p.v * vstep; zf = p.z + p.w * vstep;
                if (do_single) {
                    Frame fr;
                    uphi21(g, fr, costhe, sinthe, p);
                }
                uf = p.u; vf = p.v; wf = p.w;
            }
            deposit(P, c, p, edep);                             // :5245
            p.x = xf; p.y = yf; p.z = zf; p.u = uf; p.v = vf; p.w = wf;
            irold = p.ir;
            eie -= edep;
            p.e = eie;
            if (irnew == irl && eie <= R.ecut) { electron_end(P, c, p, p.e - RM, eie); return; }
            medold = imed;
            if (imed != -1) {
                eke = eie - RM;
                elke = log(eke);
                lelke = elec_interval(M, elke);
            }
            if (irnew != irold) {
                p.ir = irnew; irl = irnew;
                R = load_region(P, irl);
                imed = R.med;
            }
            if (eie <= R.ecut) { electron_end(P, c, p, p.e - RM, eie); return; }
            if (imed != medold) break;
            demfp -= de * sig0;
            total_tstep -= tvstep * rhof;
            if (total_tstep < 1.0E-9) demfp = 0.0;
        } while (demfp >= 1.0E-5);

        // 'continue' in the reference's do-while re-tests the STALE rfict >= sigratio (always true)
        if ((imed != medold) || (ustep == 0.0) || (imed == -1)) continue;
        const ElecBin *B = Bq + imed * MXEKE + lelke;
        double sigf = pwl(elke, __ldg(&B->sig1), __ldg(&B->sig0));
        const double dedx0 = pwl(elke, __ldg(&B->dedx1), __ldg(&B->dedx0));
        sigf /= dedx0;
        sigratio = sigf / sig0;
        rfict = g.next();
    } while (rfict >= sigratio);

    // discrete interaction, :5375-5429
    const MedRec &M = P.med[imed];
    const ElecBin *B = Bq + imed * MXEKE + lelke;
    Part q;
    bool created = false;
    c.npold = np;
    if (iq < 0) {
        const double ebr1 = pwl(elke, __ldg(&B->bra1), __ldg(&B->bra0));
        r = g.next();
        if (r <= ebr1) {
            brems(P, g, p, q, imed, P.nsplit); created = true;
        } else if (p.e <= M.thmoll) {
            if (ebr1 <= 0) { c.s[np] = p; return; }
            brems(P, g, p, q, imed, P.nsplit); created = true;
        } else {
            created = moller(P, g, p, q, imed);
        }
    } else {
        const double pbr1 = pwl(elke, __ldg(&B->bra1), __ldg(&B->bra0));
        r = g.next();
        if (r < pbr1) {
            brems(P, g, p, q, imed, P.nsplit); created = true;
        } else {
            const double pbr2 = pwl(elke, __ldg(&B->brb1), __ldg(&B->brb0));
            if (r < pbr2) bhabha(P, g, p, q, imed);
            else annih(g, p, q, P.nsplit);
            created = true;
        }
    }
    c.s[np] = p;
    if (created) {
        if (!push_ok(c, np + 1)) { c.np = -1; return; }
        c.s[np + 1] = q;
        c.np = np + 1;
    }
}

__global__ void __launch_bounds__(128) lockstep_kernel(const __grid_constant__ DevProblem P, Part *stack, int depth,
                                                       long long first, long long nhist, int ibeamlet, const Part *inject) {
    const size_t nthreads = (size_t)gridDim.x * blockDim.x;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    HistCtx c;
    c.s.base = stack + tid;
    c.s.stride = nthreads;
    c.depth = depth;
    c.nphot_steps = c.nelec_steps = 0;
    double ensrc = 0.0;
    unsigned long long ndraws = 0, ndep = 0, nerr = 0, nh = 0;
    for (long long i = (long long)tid; i < nhist; i += (long long)nthreads) {
        const unsigned long long hist = (unsigned long long)(first + i);
        c.g.seed(P.seed0, P.seed1, hist, 0u);
        c.ndeposit = 0; c.flags = 0; c.edep_sum = 0.0;
        Part p;
        if (inject) p = inject[i];                              // unit-test hook: explicit top-of-stack particle
        else ensrc += init_history(P, c.g, p, ibeamlet);        // omc_dosxyz.c:1254 / omc_matrad.c:1396
        const int ir0 = p.ir;
        c.np = 0;
        c.s[0] = p;
        while (c.np >= 0) {                                     // shower(), src/ompmc.c:5436-5447
            if (c.s[c.np].iq == 0) photon_ls(P, c);
            else electron_ls(P, c);
        }
        nh++;
        ndraws += c.g.ndraws(); ndep += c.ndeposit; nerr += (c.flags & 1u);
        if (P.records) {
            omc_history_record rec;
            rec.ndraws = c.g.ndraws(); rec.ir_start = ir0; rec.ndeposit = c.ndeposit; rec.flags = c.flags; rec.edep = c.edep_sum;
            P.records[i] = rec;
        }
    }
    atomicAdd(P.ensrc, ensrc);
    atomicAdd(&P.counters->histories, nh);
    atomicAdd(&P.counters->photon_steps, (unsigned long long)c.nphot_steps);
    atomicAdd(&P.counters->electron_steps, (unsigned long long)c.nelec_steps);
    atomicAdd(&P.counters->deposits, ndep);
    atomicAdd(&P.counters->rng_draws, ndraws);
    if (nerr) atomicAdd(&P.counters->errors, nerr);
}

// Drain kernel for the wavefront path.  When only a few thousand particles are left in flight a wave costs
// its latency, not its work, and the stragglers (e.g. electrons crossing hundreds of air voxels) need
// thousands more waves.  Each thread instead takes one queued particle -- from the photon, electron or
// pending-interaction queues of the current wave -- and follows it and everything it creates to the end
// with the per-thread LIFO shower above, continuing the particle's own Philox stream.
__global__ void __launch_bounds__(128) drain_kernel(const __grid_constant__ DevProblem P, const __grid_constant__ DrainArgs D,
                                                    Part *stack, int depth) {
    const size_t nthreads = (size_t)gridDim.x * blockDim.x;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned n[4], tot = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) { n[k] = min(*D.count[k], D.q[k].cap); tot += n[k]; }
    HistCtx c;
    c.s.base = stack + tid;
    c.s.stride = nthreads;
    c.depth = depth;
    c.nphot_steps = c.nelec_steps = 0;
    unsigned long long ndep = 0, nerr = 0;
    // With fewer particles than threads only every spread-th lane takes tickets, so that the particles are dealt over as many
    // warps as possible: the per-thread shower diverges completely (lanes serialise), and the run time of the drain is the
    // longest serialised chain of a warp.
    unsigned spread = 1;
    while (spread < 32u && (size_t)tot * (spread * 2u) <= nthreads) spread *= 2u;
    const bool takes = (threadIdx.x & (spread - 1u)) == 0u;
    for (; takes;) {
        unsigned j = atomicAdd(D.ticket, 1u);
        if (j >= tot) break;
        int k = 0;
        while (j >= n[k]) { j -= n[k]; k++; }
        const PartQueue &q = D.q[k];
        Part p, s2;
        {
            const double2 xy = q.xy[j], ze = q.ze[j];
            const float4 dw = q.dw[j];
            p.x = xy.x; p.y = xy.y; p.z = ze.x; p.e = ze.y; p.u = (double)dw.x; p.v = (double)dw.y; p.w = (double)dw.z; p.wt = (double)dw.w;
        }
        const int2 a = q.irq[j];
        p.ir = a.x; p.iq = (int)(short)(a.y & 0xffff);
        const int tag = a.y >> 16;
        const uint4 r = q.rng[j];
        c.g.seed(P.seed0, P.seed1, ((unsigned long long)r.y << 32) | r.x, r.z, r.w);
        c.ndeposit = 0; c.flags = 0; c.edep_sum = 0.0;
        bool two = false;
        if (tag != 0) {                                       // interaction that was due in the next wave
            const RegionRec R = load_region(P, p.ir);
            const int imed = R.med;
            int type = tag & 15;                              // (bit 4 = survivor flag of photon splitting, always set here)
            if (type == 0) {                                  // choice deferred by the flight chunk
                const double r1 = c.g.next(), r2 = c.g.next();
                type = photon_interaction_type(P, imed, log(p.e), p.e, r1, r2);
            }
            switch (type) {
                case 1: compton(c.g, p, s2); two = true; break;
                case 2: pair(P, c.g, p, s2, imed); two = true; break;
                case 3: photo(c.g, p, R.ecut); break;
                case 4: {
                    const MedRec &M = P.med[imed];
                    const double gle = log(p.e);
                    const PhotBin *B = P.phot + imed * MXGE + ((int)(gle * M.ge1 + M.ge0) - 1);
                    rayleigh(P, c.g, p, pwl(gle, __ldg(&B->pmax1), __ldg(&B->pmax0)), p.e);
                } break;
                case 5: brems(P, c.g, p, s2, imed, 1); two = true; break;
                case 6: two = moller(P, c.g, p, s2, imed); break;
                case 7: bhabha(P, c.g, p, s2, imed); two = true; break;
                case 8: annih(c.g, p, s2, 1); two = true; break;
                default: rannih(c.g, p, s2, 1); two = true; break;
            }
        }
        c.np = 0;
        c.s[0] = p;
        if (two) { c.s[1] = s2; c.np = 1; }
        const unsigned steps0 = c.nelec_steps;
        while (c.np >= 0) {
            if (c.s[c.np].iq == 0) photon_ls(P, c);          // a photon caught in mid flight resamples its path (memoryless)
            else electron_ls(P, c);
        }
        ndep += c.ndeposit; nerr += (c.flags & 1u);
        atomicMax(&P.counters->reserved[0], (unsigned long long)(c.nelec_steps - steps0));   // longest drained chain
        if (c.nelec_steps - steps0 > 20000u) {
            atomicAdd(&P.counters->reserved[1], 1ull);
            P.counters->reserved[2] = (unsigned long long)__double_as_longlong(p.e);
            P.counters->reserved[3] = (unsigned long long)(unsigned)p.ir | ((unsigned long long)(unsigned)(p.iq + 2) << 32);
        }
    }
    atomicAdd(&P.counters->photon_steps, (unsigned long long)c.nphot_steps);
    atomicAdd(&P.counters->electron_steps, (unsigned long long)c.nelec_steps);
    atomicAdd(&P.counters->deposits, ndep);
    if (nerr) atomicAdd(&P.counters->errors, nerr);
}

void launch_drain(const DevProblem &P, const DrainArgs &D, Part *stack, int depth, int blocks, cudaStream_t stream) {
    drain_kernel<<<blocks, 128, 0, stream>>>(P, D, stack, depth);
}

int lockstep_blocks_per_sm(int threads) {
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, lockstep_kernel, threads, 0) != cudaSuccess) n = 1;
    return n > 0 ? n : 1;
}

void launch_lockstep(const DevProblem &P, Part *stack, int depth, int blocks, int threads, long long first, long long nhist,
                     int ibeamlet, const Part *inject, cudaStream_t stream) {
    lockstep_kernel<<<blocks, threads, 0, stream>>>(P, stack, depth, first, nhist, ibeamlet, inject);
}

// ---- unit-test kernels ------------------------------------------------------------------------
__global__ void test_geometry_kernel(const __grid_constant__ DevProblem P, int n, const double *xyzuvw, const int *ir,
                                     const double *ustep_in, int *idisc, int *irnew, double *ustep_out, double *tperp) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Part p;
    p.x = xyzuvw[6 * i]; p.y = xyzuvw[6 * i + 1]; p.z = xyzuvw[6 * i + 2];
    p.u = xyzuvw[6 * i + 3]; p.v = xyzuvw[6 * i + 4]; p.w = xyzuvw[6 * i + 5];
    p.ir = ir[i]; p.iq = 0; p.e = 1.0; p.wt = 1.0;
    int id = 0, irn = ir[i];
    double us = ustep_in[i];
    howfar(P, p, id, irn, us);
    idisc[i] = id; irnew[i] = irn; ustep_out[i] = us;
    tperp[i] = hownear(P, p);
}
void launch_test_geometry(const DevProblem &P, int n, const double *xyzuvw, const int *ir, const double *ustep_in, int *idisc,
                          int *irnew, double *ustep_out, double *tperp, cudaStream_t stream) {
    test_geometry_kernel<<<(n + 127) / 128, 128, 0, stream>>>(P, n, xyzuvw, ir, ustep_in, idisc, irnew, ustep_out, tperp);
}

__global__ void test_rng_kernel(uint32_t s0, uint32_t s1, unsigned long long hist, int n, double *out) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    Rng g;
    g.seed(s0, s1, hist, 0u);
    for (int i = 0; i < n; i++) out[i] = g.next();
}
void launch_test_rng(uint32_t s0, uint32_t s1, unsigned long long hist, int n, double *out, cudaStream_t stream) {
    test_rng_kernel<<<1, 32, 0, stream>>>(s0, s1, hist, n, out);
}

// accumEndep(), omc_dosxyz.c:696-717
__global__ void accum_kernel(double *__restrict__ endep, double *__restrict__ accum, double *__restrict__ accum2, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double e = endep[i];
        accum[i] += e;
        accum2[i] += e * e;
        endep[i] = 0.0;
    }
}
void launch_accum(double *endep, double *accum, double *accum2, long long n, cudaStream_t stream) {
    int blocks = (int)((n + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    accum_kernel<<<blocks, 256, 0, stream>>>(endep, accum, accum2, n);
}

// accumulateResults(), omc_dosxyz.c:719-799, same operation order (this file is compiled with -fmad=false)
__global__ void results_kernel(const double *__restrict__ accum, const double *__restrict__ accum2, const double *__restrict__ dens,
                               const double *__restrict__ xb, const double *__restrict__ yb, const double *__restrict__ zb, int isize,
                               int jsize, int ksize, int iout, double inc_fluence, double nbatch, double *__restrict__ dose,
                               double *__restrict__ unc) {
    const long long nvox = (long long)isize * jsize * ksize;
    for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < nvox; v += (long long)gridDim.x * blockDim.x) {
        const int ix = (int)(v % isize), iy = (int)((v / isize) % jsize), iz = (int)(v / ((long long)isize * jsize));
        const long long irl = v + 1;
        double endep = accum[irl], endep2 = accum2[irl], unc_endep;
        endep /= nbatch;
        endep2 /= nbatch;
        if (endep != 0.0) {
            unc_endep = endep2 - endep * endep;
            unc_endep /= (nbatch - 1.0);
            unc_endep = sqrt(unc_endep) / endep;
        } else {
            endep = 0.0;
            unc_endep = 0.9999999;
        }
        if (iout) {
            double mass = (xb[ix + 1] - xb[ix]) * (yb[iy + 1] - yb[iy]) * (zb[iz + 1] - zb[iz]);
            mass *= dens[v];
            endep *= 1.602E-10 / (mass * inc_fluence);
        } else {
            endep /= inc_fluence;
        }
        if (dens[v] < 0.044) { endep = 0.0; unc_endep = 0.9999999; }   // "zero dose in air", :782-796
        dose[irl] = endep;
        unc[irl] = unc_endep;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { dose[0] = accum[0]; unc[0] = accum2[0]; }
}
void launch_results(const DevProblem &P, const double *accum, const double *accum2, const double *dens, int iout, int nhist, int nbatch,
                    double *dose, double *unc, cudaStream_t stream) {
    const long long nvox = (long long)P.nreg - 1;
    int blocks = (int)((nvox + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    results_kernel<<<blocks, 256, 0, stream>>>(accum, accum2, dens, P.xb, P.yb, P.zb, P.isize, P.jsize, P.ksize, iout, (double)nhist,
                                               (double)nbatch, dose, unc);
}

// ---- .3ddose text on the device (outputResults(), omc_dosxyz.c:841-879) -------------------------------------------
// One thread formats one value with omc_format.cuh ("%e " = 13 bytes, "%f " = 9 bytes, glibc-exact); a block stages its 256
// records in shared memory and stores them as 16-byte words (256 * 13 and 256 * 9 are multiples of 16, the output buffer is
// 256-byte aligned).  Values the device does not certify (rounding ties, negative / non-finite numbers, other widths) are
// appended to `fb` for the host to format with snprintf; their slot in the text is left as written here (spaces).
template <int MODE>
__global__ void __launch_bounds__(256) format_kernel(const double *__restrict__ src, long long n, const Pow10 *__restrict__ tab,
                                                     char *__restrict__ out, FormatFallback *__restrict__ fb, unsigned *__restrict__ nfb,
                                                     unsigned fb_cap) {
    constexpr int W = MODE == 0 ? kFmtEWidth : kFmtFWidth;
    __shared__ __align__(16) char sh[256 * W];
    const long long nblk = (n + 255) / 256;
    for (long long blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
        const long long i = blk * 256 + threadIdx.x;
        if (i < n) {
            const double v = src[i];
            char *o = sh + threadIdx.x * W;
            const int bad = MODE == 0 ? fmt_e(v, tab, o) : fmt_f(v, o);
            if (bad) {
                for (int c = 0; c < W; c++) o[c] = ' ';
                const unsigned slot = atomicAdd(nfb, 1u);
                if (slot < fb_cap) { fb[slot].index = (unsigned long long)i; fb[slot].value = v; }
            }
        }
        __syncthreads();
        const long long base = blk * 256 * W;
        const long long cnt = (n - blk * 256 < 256 ? n - blk * 256 : 256) * W;
        if (cnt == 256 * W) {
            const uint4 *s4 = reinterpret_cast<const uint4 *>(sh);
            uint4 *o4 = reinterpret_cast<uint4 *>(out + base);
            for (int w = threadIdx.x; w < 256 * W / 16; w += 256) o4[w] = s4[w];
        } else {
            for (long long c = threadIdx.x; c < cnt; c += 256) out[base + c] = sh[c];
        }
        __syncthreads();
    }
}
void launch_format(int mode, const double *src, long long n, const Pow10 *tab, char *out, FormatFallback *fb, unsigned *nfb, unsigned fb_cap,
                   cudaStream_t stream) {
    long long blocks = (n + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    if (mode == 0) format_kernel<0><<<(int)blocks, 256, 0, stream>>>(src, n, tab, out, fb, nfb, fb_cap);
    else format_kernel<1><<<(int)blocks, 256, 0, stream>>>(src, n, tab, out, fb, nfb, fb_cap);
}

}  // namespace omc

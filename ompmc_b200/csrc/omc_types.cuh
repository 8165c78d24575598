// omc_types.cuh -- device-side data layout of one ompMC problem (HBM-resident, read-only during
// transport) and the particle record.  See DESIGN.md "Data layout in HBM".
//
// The reference keeps every PWL table as two separate double arrays (c1[], c0[]) per quantity
// (struct Photon/Electron, src/ompmc.h:105-249), so one electron sub-step gathers ~12 scattered
// 8-byte words.  Here all quantities that are looked up with the SAME (medium, energy-bin) index
// are interleaved into one record, so a sub-step touches one contiguous 160-byte record (five 32-byte
// sectors) and a photon free flight one 80-byte record.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/ompmc_b200.h"

namespace omc {

constexpr double RM = OMC_RM;
constexpr int MXGE = OMC_MXGE;
constexpr int MXEKE = OMC_MXEKE;

// Particle: one slot of struct Stack (src/ompmc.h:51-70) without dnear (never read).
struct Part {
    double x, y, z, u, v, w, e, wt;
    int ir, iq;
};

// Per-region transport data: struct Region (src/ompmc.h:412-418), one 32-byte sector per voxel.
struct __align__(32) RegionRec {
    double rhof, ecut, pcut;
    int med;      // 0-based medium, -1 vacuum
    int pad;
};

// photon_data {gmfp,cohe,gbr1,gbr2} + rayleigh_data.pmax at index imed*MXGE + lgle
struct __align__(16) PhotBin {
    double gmfp1, gmfp0, cohe1, cohe0, gbr11, gbr10, gbr21, gbr20, pmax1, pmax0;
};

// electron_data at index imed*MXEKE + lelke, one copy per charge state qel (0: e-, 1: e+)
struct __align__(32) ElecBin {
    double sig1, sig0;       // esig / psig
    double dedx1, dedx0;     // ededx / pdedx
    double tmxs1, tmxs0;
    double eta1, eta0;       // etae_ms / etap_ms
    double blcce1, blcce0;
    double q1c1, q1c0;       // q1ce_ms / q1cp_ms
    double q2c1, q2c0;       // q2ce_ms / q2cp_ms
    double bra1, bra0;       // ebr1 / pbr1
    double brb1, brb0;       // (unused) / pbr2
    double range_ep;         // range_ep[qel][imed][lelke]
    double e_array;          // e_array[imed][lelke]
};

// per-medium scalars: pegs_data, pair_data, eke0/1, ge0/1 ...
struct MedRec {
    double ge1, ge0;         // photon_data.ge1/ge0
    double eke1, eke0;       // electron_data.eke1/eke0
    double xcc, blcc, esig_e, psig_e;
    double te, thmoll, ap;
    double delcm, zbrang, bpar0, bpar1;
    double dl[6][8];         // pair_data.dl1..dl6 [8]
    double ecut, pcut;       // region.ecut/pcut when they depend on the medium only (see DevProblem::reg8)
    double rhomax;           // largest region.rhof of the voxels filled with this medium (0: medium absent); Woodcock majorant
    int sig_ismonotone[2];   // [qel]
};

// mscat_data alias tables: one 32-byte sector per (i, j, k)
struct __align__(32) MsEntry {
    double ums, fms, wms;
    int ims, pad;
};

// fp32 copy of MsEntry for the single-precision samplers (omc_physics_f32.cuh): one 16-byte load
struct __align__(16) MsEntryF {
    float ums, wms;
    int ims;
    float fms;
};

struct SourceDosxyz {       // struct Source, omc_dosxyz.c:342-366
    int spectrum, charge;
    double energy, deltak;
    const double *cdfinv1, *cdfinv2;
    double ssd, xinl, yinl, xsize, ysize;
    int ixinl, iyinl;
};

struct SourceMatrad {       // struct Source of omc_matrad.c:507-541 (bixel arrays), device pointers
    int nbixels, nbeams;
    const int *ibeam;
    const double *xsource, *ysource, *zsource;
    const double *xcorner, *ycorner, *zcorner, *xside1, *yside1, *zside1, *xside2, *yside2, *zside2;
};

struct Counters {            // mirrors omc_gpu_counters (include/ompmc_b200.h)
    unsigned long long histories, kernel_launches, photon_steps, electron_steps, deposits, rng_draws, errors;
    unsigned long long reserved[9];
};

// Everything a kernel needs, passed by value as a __grid_constant__ parameter (constant bank).
struct DevProblem {
    // geometry: struct Geom (omc_dosxyz.c:46-59)
    int isize, jsize, ksize, ijmax, nreg, nmed;
    const double *xb, *yb, *zb;
    const RegionRec *reg;
    const void *reg8;        // compact {float rhof; int med} records, or nullptr (see load_region_w)
    double inv_dx, inv_dy, inv_dz;   // 1 / voxel size per axis when that axis is uniformly spaced (find_bin)
    int uniform_x, uniform_y, uniform_z;
    // media
    const MedRec *med;
    const PhotBin *phot;     // [nmed*MXGE]
    const ElecBin *ebin;     // [2][nmed*MXEKE]
    const double *ray_xgrid, *ray_fcum, *ray_b, *ray_c;   // rayleigh_data, medium 0 only is ever read (Q3)
    const int *ray_i;
    double b2spin_min, dbeta2i, espml, dleneri, dqq1i;    // spin_data
    const double *spin_rej;  // [nmed][2][32][16][32]
    const MsEntry *ms;       // [64][8][32]
    const float *spin_rej_f; // fp32 copies for the wavefront kernels
    const MsEntryF *ms_f;
    double dllambi, dqmsi;
    SourceDosxyz src;        // spectrum / charge / energy are shared by both source kinds
    SourceMatrad msrc;
    int nsplit;
    uint32_t seed0, seed1;
    // scoring
    double *endep;           // fp64 batch grid [nreg]  (struct Score.endep, omc_dosxyz.c:638)
    float *endep32;          // fp32 chunk grid [nreg]  (wavefront kernels)
    double *ensrc;           // score.ensrc
    Counters *counters;
    omc_history_record *records;   // nullable
};

}  // namespace omc

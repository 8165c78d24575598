// omc_multi.cu -- several GPUs of one node behind ONE handle, for single-process C user codes (include/ompmc_b200.h,
// "multi-GPU" section).  Built on the public C-ABI only: one omc_gpu_handle per device, one host thread per device for
// every call that runs transport (each device's wave loop blocks its own thread), an NCCL communicator per device
// (omc_gpu_comm_init) so that omc_gpu_run_batch() shards the history ids of a batch over the devices and sums the
// completed batch grids on side streams.  Replaces the OpenMP team of omc_dosxyz.c:1184-1263 / omc_matrad.c:1389-1493
// at node scale.  No transport code here.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <functional>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/ompmc_b200.h"

struct omc_gpu_multi_ctx {
    std::vector<omc_gpu_handle> h;
    std::vector<int> dev;
    std::string err;
    // gathered sparse columns of the last omc_gpu_multi_run_beamlets()
    std::vector<long long> col_ir;
    std::vector<double> col_val;
};

namespace {

// f(i) for every device on its own thread; first non-zero return code wins
int on_all(omc_gpu_multi m, const std::function<int(int)> &f) {
    const int n = (int)m->h.size();
    std::vector<int> rc((size_t)n, 0);
    if (n == 1) {
        rc[0] = f(0);
    } else {
        std::vector<std::thread> th;
        th.reserve((size_t)n);
        for (int i = 0; i < n; i++) th.emplace_back([&, i] { rc[(size_t)i] = f(i); });
        for (auto &t : th) t.join();
    }
    for (int i = 0; i < n; i++)
        if (rc[(size_t)i]) {
            m->err = "device " + std::to_string(m->dev[(size_t)i]) + ": " + omc_gpu_last_error(m->h[(size_t)i]);
            return rc[(size_t)i];
        }
    return 0;
}

}  // namespace

extern "C" {

int omc_gpu_multi_create(omc_gpu_multi *out, int ndev, const int *device_ids) {
    if (!out) return 2;
    *out = nullptr;
    int have = 0;
    if (cudaGetDeviceCount(&have) != cudaSuccess || have <= 0) {
        fprintf(stderr, "ompmc_b200: no CUDA device available; this library has no CPU fallback\n");
        return 3;
    }
    if (ndev <= 0) ndev = have;
    omc_gpu_multi m = new omc_gpu_multi_ctx();
    m->h.assign((size_t)ndev, nullptr);
    for (int i = 0; i < ndev; i++) m->dev.push_back(device_ids ? device_ids[i] : i);
    // every device's context comes up on its own thread (a CUDA context costs ~0.5 s; eight in a row was most of the start-up)
    std::vector<int> crc((size_t)ndev, 0);
    {
        std::vector<std::thread> th;
        for (int i = 0; i < ndev; i++) th.emplace_back([&, i] { crc[(size_t)i] = omc_gpu_create(&m->h[(size_t)i], m->dev[(size_t)i]); });
        for (auto &t : th) t.join();
    }
    for (int i = 0; i < ndev; i++)
        if (crc[(size_t)i]) {
            const int rc = crc[(size_t)i];
            for (omc_gpu_handle x : m->h) if (x) omc_gpu_destroy(x);
            delete m;
            return rc;
        }
    if (ndev > 1) {
        char id[128];
        int rc = omc_gpu_comm_unique_id(id);
        if (!rc) rc = on_all(m, [&](int i) { return omc_gpu_comm_init(m->h[(size_t)i], i, ndev, id); });
        if (rc) {
            fprintf(stderr, "ompmc_b200: NCCL communicator over %d devices failed: %s\n", ndev, m->err.c_str());
            for (omc_gpu_handle x : m->h) omc_gpu_destroy(x);
            delete m;
            return rc;
        }
    }
    *out = m;
    return 0;
}

void omc_gpu_multi_destroy(omc_gpu_multi m) {
    if (!m) return;
    on_all(m, [&](int i) { omc_gpu_destroy(m->h[(size_t)i]); return 0; });
    delete m;
}

int omc_gpu_multi_size(omc_gpu_multi m) { return m ? (int)m->h.size() : -1; }
omc_gpu_handle omc_gpu_multi_device(omc_gpu_multi m, int i) { return (m && i >= 0 && i < (int)m->h.size()) ? m->h[(size_t)i] : nullptr; }
const char *omc_gpu_multi_last_error(omc_gpu_multi m) { return m ? m->err.c_str() : "null handle"; }

// ---- the problem goes to every device --------------------------------------------------------------------------------------
int omc_gpu_multi_set_media(omc_gpu_multi m, const omc_media_tables *t) {
    return m ? on_all(m, [&](int i) { return omc_gpu_set_media(m->h[(size_t)i], t); }) : 2;
}
int omc_gpu_multi_set_geometry(omc_gpu_multi m, const omc_geometry *g) {
    return m ? on_all(m, [&](int i) { return omc_gpu_set_geometry(m->h[(size_t)i], g); }) : 2;
}
int omc_gpu_multi_set_source_dosxyz(omc_gpu_multi m, const omc_source_dosxyz *s) {
    return m ? on_all(m, [&](int i) { return omc_gpu_set_source_dosxyz(m->h[(size_t)i], s); }) : 2;
}
int omc_gpu_multi_set_source_matrad(omc_gpu_multi m, const omc_source_matrad *s) {
    return m ? on_all(m, [&](int i) { return omc_gpu_set_source_matrad(m->h[(size_t)i], s); }) : 2;
}
int omc_gpu_multi_set_vrt(omc_gpu_multi m, int nsplit) {
    return m ? on_all(m, [&](int i) { return omc_gpu_set_vrt(m->h[(size_t)i], nsplit); }) : 2;
}
int omc_gpu_multi_set_seed(omc_gpu_multi m, int ixx, int jxx) {
    return m ? on_all(m, [&](int i) { return omc_gpu_set_seed(m->h[(size_t)i], ixx, jxx); }) : 2;
}
int omc_gpu_multi_set_option(omc_gpu_multi m, const char *key, long long value) {
    return m ? on_all(m, [&](int i) { return omc_gpu_set_option(m->h[(size_t)i], key, value); }) : 2;
}
int omc_gpu_multi_reset_tallies(omc_gpu_multi m, int which) {
    return m ? on_all(m, [&](int i) { return omc_gpu_reset_tallies(m->h[(size_t)i], which); }) : 2;
}

// ---- the batch loop ----------------------------------------------------------------------------------------------------------
// one iteration of omc_dosxyz.c:1237-1263 on all devices: every device takes its slice of [first, first + nhist)
int omc_gpu_multi_run_batch(omc_gpu_multi m, long long first_history, long long nhist, int ibeamlet) {
    return m ? on_all(m, [&](int i) { return omc_gpu_run_batch(m->h[(size_t)i], first_history, nhist, ibeamlet); }) : 2;
}

// completes what is in flight everywhere (the collectives of the last batches need every device to take part)
int omc_gpu_multi_synchronize(omc_gpu_multi m) {
    return m ? on_all(m, [&](int i) { return omc_gpu_synchronize(m->h[(size_t)i]); }) : 2;
}

int omc_gpu_multi_get_tallies(omc_gpu_multi m, double *accum, double *accum2, double *ensrc) {
    if (!m) return 2;
    int rc = omc_gpu_multi_synchronize(m);
    if (rc) return rc;
    // the batch grids were summed over the devices before accumEndep(): every device holds the same accum / accum2
    double e0 = 0.0;
    rc = omc_gpu_get_tallies(m->h[0], accum, accum2, &e0);
    if (rc) { m->err = omc_gpu_last_error(m->h[0]); return rc; }
    if (ensrc) {                                               // score.ensrc is per device: sum
        double tot = e0;
        for (size_t i = 1; i < m->h.size(); i++) {
            double e = 0.0;
            rc = omc_gpu_get_tallies(m->h[i], nullptr, nullptr, &e);
            if (rc) { m->err = omc_gpu_last_error(m->h[i]); return rc; }
            tot += e;
        }
        *ensrc = tot;
    }
    return 0;
}

int omc_gpu_multi_accumulate_results(omc_gpu_multi m, int iout, int nhist, int nbatch, const double *med_densities, double *dose, double *unc) {
    if (!m) return 2;
    int rc = omc_gpu_multi_synchronize(m);
    if (rc) return rc;
    rc = omc_gpu_accumulate_results(m->h[0], iout, nhist, nbatch, med_densities, dose, unc);
    if (rc) m->err = omc_gpu_last_error(m->h[0]);
    return rc;
}

int omc_gpu_multi_write_3ddose(omc_gpu_multi m, const char *path, int iout, int nhist, int nbatch, const double *med_densities) {
    if (!m) return 2;
    int rc = omc_gpu_multi_synchronize(m);
    if (rc) return rc;
    rc = omc_gpu_write_3ddose(m->h[0], path, iout, nhist, nbatch, med_densities);
    if (rc) m->err = omc_gpu_last_error(m->h[0]);
    return rc;
}

int omc_gpu_multi_get_counters(omc_gpu_multi m, omc_gpu_counters *out) {
    if (!m || !out) return 2;
    int rc = omc_gpu_multi_synchronize(m);
    if (rc) return rc;
    memset(out, 0, sizeof *out);
    for (size_t i = 0; i < m->h.size(); i++) {
        omc_gpu_counters c;
        rc = omc_gpu_get_counters(m->h[i], &c);
        if (rc) { m->err = omc_gpu_last_error(m->h[i]); return rc; }
        out->histories += c.histories; out->kernel_launches += c.kernel_launches; out->photon_steps += c.photon_steps;
        out->electron_steps += c.electron_steps; out->deposits += c.deposits; out->rng_draws += c.rng_draws; out->errors += c.errors;
    }
    return 0;
}

// ---- the beamlet loop of omc_matrad.c:1389-1493 over the devices ---------------------------------------------------------------
// Beamlets [ib0, ib0 + nb) are cut into passes of `per_pass` consecutive beamlets (<= 0: OMC_BEAMLETS_PER_PASS); pass p goes to
// device p % ndev (whole passes: a device that owned single beamlets would pay the tail of the longest lineages once per
// beamlet); beamlet ib0 + k owns history ids [first + k * nhist, + nhist) whatever the device count, so the matrix does not
// depend on it beyond the summation order of the fp32 dose atomics.  The column slices are gathered here in beamlet order
// (the order in which the reference's loop appends them, :1416-1477): jc[nb + 1], then omc_gpu_multi_fetch_columns().
int omc_gpu_multi_run_beamlets(omc_gpu_multi m, long long first_history, int nhist, int nbatch, int ib0, int nb, int per_pass,
                               double rel_threshold, const double *med_densities, long long *jc, long long *nnz_total) {
    if (!m || !jc || !nnz_total || !med_densities) return 2;
    if (nb < 1) { m->err = "beamlet range out of bounds"; return 2; }
    if (per_pass <= 0) per_pass = OMC_BEAMLETS_PER_PASS;
    const int ndev = (int)m->h.size();
    if ((nb + per_pass - 1) / per_pass < ndev) per_pass = (nb + ndev - 1) / ndev;     // fewer passes than devices: smaller passes, every device busy
    const int npass = (nb + per_pass - 1) / per_pass;
    struct Pass {
        std::vector<long long> jc, ir;
        std::vector<double> val;
    };
    std::vector<Pass> pass((size_t)npass);
    int rc = on_all(m, [&](int d) {
        for (int p = d; p < npass; p += ndev) {
            const int b0 = p * per_pass, cnt = std::min(per_pass, nb - b0);
            Pass &P = pass[(size_t)p];
            P.jc.assign((size_t)cnt + 1, 0);
            long long nnz = 0;
            int r = omc_gpu_run_beamlets(m->h[(size_t)d], first_history + (long long)b0 * nhist, nhist, nbatch, ib0 + b0, cnt, rel_threshold,
                                         med_densities, P.jc.data(), &nnz);
            if (r) return r;
            P.ir.resize((size_t)nnz); P.val.resize((size_t)nnz);
            if (nnz > 0 && (r = omc_gpu_fetch_columns(m->h[(size_t)d], P.ir.data(), P.val.data()))) return r;
        }
        return 0;
    });
    if (rc) return rc;
    long long total = 0;
    jc[0] = 0;
    for (int p = 0; p < npass; p++) {
        const Pass &P = pass[(size_t)p];
        const int b0 = p * per_pass, cnt = (int)P.jc.size() - 1;
        for (int k = 0; k < cnt; k++) jc[b0 + k + 1] = total + P.jc[(size_t)k + 1];
        total += P.jc[(size_t)cnt];
    }
    m->col_ir.resize((size_t)total); m->col_val.resize((size_t)total);
    long long at = 0;
    for (const Pass &P : pass) {
        if (!P.ir.empty()) {
            memcpy(m->col_ir.data() + at, P.ir.data(), P.ir.size() * sizeof(long long));
            memcpy(m->col_val.data() + at, P.val.data(), P.val.size() * sizeof(double));
        }
        at += (long long)P.ir.size();
    }
    *nnz_total = total;
    return 0;
}

int omc_gpu_multi_fetch_columns(omc_gpu_multi m, long long *ir, double *val) {
    if (!m || !ir || !val) return 2;
    if (!m->col_ir.empty()) {
        memcpy(ir, m->col_ir.data(), m->col_ir.size() * sizeof(long long));
        memcpy(val, m->col_val.data(), m->col_val.size() * sizeof(double));
    }
    return 0;
}

}  // extern "C"

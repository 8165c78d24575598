/*
 * omc_oracle.h -- CPU restatement of ompMC's shower() hot path.  TEST INFRASTRUCTURE (oracle/).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load the library built
 * from omc_oracle.c; the product (ompmc_b200/libompmc_b200.so) never links or calls it.
 *
 * Parity pin: the reference ships no golden vectors, so the oracle is pinned against
 * outputs of the reference itself: tests/golden/golden_*.npz were produced by running the
 * unmodified reference (oracle/_ref, oracle/ref_harness.c) with the per-history Philox stream,
 * and tests/test_oracle.py requires this restatement to reproduce them history by history
 * (draw counts and start regions bit-exact, deposited energy to 1e-9 relative).
 */
#ifndef OMC_ORACLE_H
#define OMC_ORACLE_H
#include "../include/ompmc_b200.h"

int    orc_load_problem(const char *blob_path);
void   orc_set_rng(int mode /*0 RANMAR, 1 Philox*/, int seed0, int seed1);
void   orc_set_nsplit(int nsplit);
void   orc_set_beamlet(int ibeamlet);   /* matRad source: beamlet of the following run_histories() calls */
int    orc_nreg(void);
void   orc_run_histories(long long first, long long n, omc_history_record *rec);
void   orc_accum_endep(void);
void   orc_reset_score(void);
void   orc_zero_accum(void);
void   orc_get_endep(double *out);
void   orc_get_accum(double *a, double *a2, double *ensrc);
double orc_time_batches(long long first, long long nperbatch, int nbatch);
int    orc_num_threads(void);
void   orc_set_num_threads(int n);
void   orc_test_geometry(int n, const double *xyzuvw, const int *ir, const double *ustep_in, int *idisc,
                         int *irnew, double *ustep_out, double *tperp);
void   orc_test_rng(long long hist, int n, double *out);
void   orc_run_particle(long long hist, int iq, double e, const double *xyzuvw, int ir, double wt,
                        omc_history_record *rec);
void   orc_get_work(unsigned long long *out7);
void   orc_set_endep(const double *in);
/* n RANMAR draws for seeds (ixx, jxx), src/omc_random.c:58-187 */
void   orc_test_ranmar(int ixx, int jxx, int n, double *out);
#endif

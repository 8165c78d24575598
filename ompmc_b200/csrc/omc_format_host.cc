// omc_format_host.cc -- omc_format.cuh compiled for the HOST by g++ (libomc_format_host.so): lets the CPU test-suite check the
// very code the formatting kernel runs (fmt_e / fmt_f are __host__ __device__ there) against glibc's snprintf.  Test hook only;
// the product formats on the device (omc_gpu_write_3ddose).
#include "omc_format.cuh"

extern "C" {
// text[n * width] (13 for mode 0 "%e ", 9 for mode 1 "%f "), flags[n] = 1 where the value is left to snprintf
long long omc_format_host(int mode, long long n, const double *values, char *text, unsigned char *flags) {
    static omc::Pow10 tab[omc::kPow10N];
    static bool built = false;
    if (!built) { omc::build_pow10_table(tab); built = true; }
    const int w = mode == 0 ? omc::kFmtEWidth : omc::kFmtFWidth;
    long long nflag = 0;
    for (long long i = 0; i < n; i++) {
        char *o = text + i * w;
        const int bad = mode == 0 ? omc::fmt_e(values[i], tab, o) : omc::fmt_f(values[i], o);
        if (bad) for (int c = 0; c < w; c++) o[c] = ' ';
        flags[i] = (unsigned char)bad;
        nflag += bad;
    }
    return nflag;
}
int omc_format_width(int mode) { return mode == 0 ? omc::kFmtEWidth : omc::kFmtFWidth; }
}

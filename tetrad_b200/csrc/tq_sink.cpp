// Host-side result sink of the hot path (no CUDA): the text the reference produces with pandas
// and per-line Python, formatted in C.
//   tq_format_tsv    rows of {name}.quartets_{rep}.tsv   (tetrad/src/run_inference.py:232-234)
//   tq_round_scores  the value each score has after a trip through that text (the reference computes
//                    its wQMC weights from the parsed TSV, run_inference.py:262-297)
//   tq_format_qmc    lines "a,b|c,d:weight" of the wQMC input (run_inference.py:305)
#include "../../include/tetrad_b200.h"
#include "tq_error.h"

#include <cmath>
#include <cstdio>
#include <cstring>

namespace {

// x >= 0 rounded to `digits` decimals exactly as printf("%.*f") does, returned as the integer
// N = round(x * 10^digits).  Fast path: scale and round; when the scaled value is too close to a
// rounding boundary for double arithmetic to be trusted (or too large), let snprintf decide.
bool scaled_round(double x, int digits, unsigned long long *out)
{
    static const double pow10[] = {1, 10, 100, 1000, 10000, 100000, 1000000};
    if (!(x >= 0.0) || x >= 4.0e9) return false;
    const double ip = std::floor(x);
    const double t = (x - ip) * pow10[digits];            // x - ip is exact; the product has one rounding
    const double ft = std::floor(t);
    const double d = t - ft;
    if (std::fabs(d - 0.5) < 1e-5) return false;          // near a tie: exact decimal expansion needed
    unsigned long long frac = (unsigned long long)ft + (d > 0.5 ? 1ull : 0ull);
    unsigned long long whole = (unsigned long long)ip;
    const unsigned long long scale = (unsigned long long)pow10[digits];
    if (frac >= scale) { frac -= scale; whole += 1; }
    *out = whole * scale + frac;
    return true;
}

unsigned long long exact_round(double x, int digits)
{
    unsigned long long n;
    if (scaled_round(x, digits, &n)) return n;
    char buf[512];
    snprintf(buf, sizeof(buf), "%.*f", digits, x);        // correctly rounded by the C library
    unsigned long long whole = 0, frac = 0;
    const char *p = buf;
    while (*p >= '0' && *p <= '9') whole = whole * 10 + (unsigned)(*p++ - '0');
    if (*p == '.') ++p;
    while (*p >= '0' && *p <= '9') frac = frac * 10 + (unsigned)(*p++ - '0');
    unsigned long long scale = 1;
    for (int i = 0; i < digits; ++i) scale *= 10;
    return whole * scale + frac;
}

inline char *put_uint(char *p, unsigned long long v)
{
    char tmp[24];
    int n = 0;
    do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    while (n) *p++ = tmp[--n];
    return p;
}

inline char *put_fixed(char *p, unsigned long long scaled, int digits)
{
    unsigned long long scale = 1;
    for (int i = 0; i < digits; ++i) scale *= 10;
    p = put_uint(p, scaled / scale);
    *p++ = '.';
    unsigned long long frac = scaled % scale;
    for (unsigned long long s = scale / 10; s; s /= 10) { *p++ = (char)('0' + frac / s); frac %= s; }
    return p;
}

bool printable(double x) { return x >= 0.0 && x < 4.0e9; }   // NaN / inf / negative: not a score

// A value the fast path does not take (NaN, inf, negative, >= 4e9) is written the way the reference's own formatting
// would write it: printf's full decimal expansion, up to 316 characters for a finite double.  It goes through a
// bounded temporary and is copied only if the caller's buffer still has room for it AND the rest of the row
// (`tail` bytes); returns nullptr otherwise.
char *put_odd(char *p, const char *end, double x, int digits, int tail)
{
    char tmp[400];
    int n = snprintf(tmp, sizeof(tmp), "%.*f", digits, x);
    if (n < 0) n = 0;
    if (n > (int)sizeof(tmp) - 1) n = (int)sizeof(tmp) - 1;
    if (end - p < n + tail) return nullptr;
    memcpy(p, tmp, (size_t)n);
    return p + n;
}

}  // namespace

extern "C" {

int64_t tq_format_tsv(const uint32_t *quartets, const uint32_t *rstat, const double *rscor, int64_t m,
                      char *out, int64_t cap)
{
    if (m < 0 || (m > 0 && (!quartets || !rstat || !rscor || !out))) return tq::set_error(TQ_EINVAL, "tq_format_tsv: bad arguments");
    char *p = out;
    const char *end = out + cap;
    for (int64_t i = 0; i < m; ++i) {
        // longest row of printable scores is < 160 bytes (4 x 10 + 3 x 17 + 2 x 10 digits, 9 separators)
        if (end - p < 160) return tq::set_error(TQ_ENOMEM, "tq_format_tsv: output buffer too small at row %lld", (long long)i);
        for (int k = 0; k < 4; ++k) { p = put_uint(p, quartets[4 * i + k]); *p++ = '\t'; }
        for (int k = 0; k < 3; ++k) {
            const double x = rscor[3 * i + k];
            if (printable(x)) p = put_fixed(p, exact_round(x, 6), 6);
            else if (!(p = put_odd(p, end, x, 6, 64)))      // 64: the rest of the row if every later score is printable
                return tq::set_error(TQ_ENOMEM, "tq_format_tsv: row %lld holds a score of more than 17 characters (%g); buffer too small", (long long)i, x);
            *p++ = '\t';
        }
        p = put_uint(p, rstat[2 * i]); *p++ = '\t';
        p = put_uint(p, rstat[2 * i + 1]); *p++ = '\n';
    }
    return (int64_t)(p - out);
}

int tq_round_scores(const double *rscor, int64_t n, double *out)
{
    if (n < 0 || (n > 0 && (!rscor || !out))) return tq::set_error(TQ_EINVAL, "tq_round_scores: bad arguments");
    for (int64_t i = 0; i < n; ++i) {
        const double x = rscor[i];
        // N / 10^6 with N < 2^53 is the correctly rounded value of the decimal text, i.e. what strtod returns
        out[i] = printable(x) ? (double)exact_round(x, 6) / 1.0e6 : x;
    }
    return TQ_OK;
}

int64_t tq_qmc_weights(const uint32_t *quartets, const uint32_t *rstat, const double *rscor, int64_t m, int weights,
                       int64_t min_snps, double min_ratio, uint32_t *tips_out, double *weight_out)
{
    if (m < 0 || weights < 0 || weights > 3 || (m > 0 && (!quartets || !rstat || !rscor || !tips_out || !weight_out)))
        return tq::set_error(TQ_EINVAL, "tq_qmc_weights: bad arguments");
    static const int order[3][4] = {{0, 1, 2, 3}, {0, 2, 1, 3}, {0, 3, 1, 2}};   // split -> a,b|c,d (run_inference.py:41-45)
    if (min_snps < 1) min_snps = 1;
    int64_t n = 0;
    for (int64_t i = 0; i < m; ++i) {
        double sc[3];
        for (int k = 0; k < 3; ++k) {
            const double x = rscor[3 * i + k];
            if (!printable(x)) return TQ_EINVAL;                       // the caller's numpy path handles odd values
            sc[k] = (double)exact_round(x, 6) / 1.0e6;                 // the score as the reference reads it back from the TSV
        }
        double weight = 1.0, ratio = 1.0;
        if (weights) {
            const double lo_ab = sc[0] < sc[1] ? sc[0] : sc[1], hi_ab = sc[0] < sc[1] ? sc[1] : sc[0];
            const double smin = lo_ab < sc[2] ? lo_ab : sc[2];
            const double smax = hi_ab > sc[2] ? hi_ab : sc[2];
            const double m2 = hi_ab < sc[2] ? hi_ab : sc[2];
            const double smid = lo_ab > m2 ? lo_ab : m2;
            const double smean = (smid + smax) / 2.0;                  // mean of the two largest (:281-283)
            ratio = (smin == 0.0) ? 1.0 : smean / smin;
            const double third = 1.0 - smin / ((smin + smid) + smax);
            weight = weights == 1 ? smean : weights == 2 ? ratio : third;
        }
        if ((int64_t)rstat[2 * i + 1] < min_snps || ratio < min_ratio) continue;   // filters :258, :275, :300
        const uint32_t best = rstat[2 * i] < 2u ? rstat[2 * i] : 2u;
        for (int k = 0; k < 4; ++k) tips_out[4 * n + k] = quartets[4 * i + order[best][k]];
        weight_out[n] = weight;
        ++n;
    }
    return n;
}

static int64_t format_qmc_rows(const uint32_t *tips, const double *weight, const int64_t *rows, int64_t n, char *out, int64_t cap)
{
    if (n < 0 || (n > 0 && (!tips || !weight || !out))) return tq::set_error(TQ_EINVAL, "tq_format_qmc: bad arguments");
    char *p = out;
    const char *end = out + cap;
    for (int64_t k = 0; k < n; ++k) {
        if (end - p < 128) return tq::set_error(TQ_ENOMEM, "tq_format_qmc: output buffer too small at line %lld", (long long)k);
        const int64_t i = rows ? rows[k] : k;
        p = put_uint(p, tips[4 * i]); *p++ = ',';
        p = put_uint(p, tips[4 * i + 1]); *p++ = '|';
        p = put_uint(p, tips[4 * i + 2]); *p++ = ',';
        p = put_uint(p, tips[4 * i + 3]); *p++ = ':';
        const double w = weight[i];
        if (printable(w)) p = put_fixed(p, exact_round(w, 5), 5);
        else if (!(p = put_odd(p, end, w, 5, 1)))
            return tq::set_error(TQ_ENOMEM, "tq_format_qmc: line %lld holds a weight of more than 16 characters (%g); buffer too small", (long long)k, w);
        *p++ = '\n';
    }
    return (int64_t)(p - out);
}

int64_t tq_format_qmc(const uint32_t *tips, const double *weight, int64_t n, char *out, int64_t cap)
{
    return format_qmc_rows(tips, weight, nullptr, n, out, cap);
}

int64_t tq_format_qmc_rows(const uint32_t *tips, const double *weight, const int64_t *rows, int64_t n, char *out, int64_t cap)
{
    if (n > 0 && !rows) return tq::set_error(TQ_EINVAL, "tq_format_qmc_rows: rows is NULL");
    return format_qmc_rows(tips, weight, rows, n, out, cap);
}

}  // extern "C"

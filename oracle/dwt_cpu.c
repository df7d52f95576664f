/* Plain-C restatement of the separable periodized DWT / IDWT (forward rows-then-columns, inverse
 * columns-then-rows) -- TEST / BASELINE INFRASTRUCTURE ONLY, never linked into the product.
 *
 * Follows pdwt/src/separable.cu:91-176 (analysis, odd sizes extended by one repeated sample) and
 * :246-328 (polyphase synthesis), i.e. what PyWavelets' wavedec2/waverec2(mode="periodization")
 * compute -- the CPU path the reference's tests compare against (test/test_wavelets.py:230,276).
 * fp32 data, fp32 accumulation (like DTYPE float, filters.h:18).  OpenMP over rows when available.
 *
 *   gcc -O3 -march=native -fopenmp -shared -fPIC -o oracle/_build/libdwt_cpu.so oracle/dwt_cpu.c
 */
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static inline int wrap_dwt(int i, int N) {
    const int Ne = N + (N & 1);
    i %= Ne;
    if (i < 0) i += Ne;
    return i >= N ? N - 1 : i;
}
static inline int wrap_per(int i, int N) {
    i %= N;
    return i < 0 ? i + N : i;
}

int dwt_cpu_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* torchrun exports OMP_NUM_THREADS=1 to its workers: the baseline sets its thread count itself */
void dwt_cpu_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* analysis of `rows` signals of length N with stride (rs = row stride, es = element stride) */
static void analysis(const float* in, float* lo, float* hi, int rows, int N, long rs_in, long es_in,
                     long rs_out, long es_out, const float* L, const float* H, int F) {
    const int N2 = (N + 1) / 2, c = (F - 1) / 2;
#pragma omp parallel for schedule(static)
    for (int r = 0; r < rows; r++) {
        const float* x = in + r * rs_in;
        for (int k = 0; k < N2; k++) {
            float a = 0.f, d = 0.f;
            if (2 * k - c >= 0 && 2 * k - c + F <= N) {
                const float* p = x + (long)(2 * k - c) * es_in;
                for (int j = 0; j < F; j++) {
                    a += p[j * es_in] * L[F - 1 - j];
                    d += p[j * es_in] * H[F - 1 - j];
                }
            } else {
                for (int j = 0; j < F; j++) {
                    const float v = x[(long)wrap_dwt(2 * k - c + j, N) * es_in];
                    a += v * L[F - 1 - j];
                    d += v * H[F - 1 - j];
                }
            }
            lo[r * rs_out + k * es_out] = a;
            hi[r * rs_out + k * es_out] = d;
        }
    }
}

/* synthesis: out[n] = sum_j lo[k]*IL[t] + hi[k]*IH[t], t = 2j + ((b+p)&1), k = (n>>1) + ((b+p)>>1) - j */
static void synthesis(const float* lo, const float* hi, float* out, int rows, int n2, int N, long rs_in,
                      long es_in, long rs_out, long es_out, const float* IL, const float* IH, int F) {
    const int p = F / 2 - 1, half = F / 2;
#pragma omp parallel for schedule(static)
    for (int r = 0; r < rows; r++) {
        const float* a = lo + r * rs_in;
        const float* d = hi + r * rs_in;
        for (int n = 0; n < N; n++) {
            const int b = n & 1, t0 = (b + p) & 1, k0 = (n >> 1) + ((b + p) >> 1);
            float s = 0.f;
            for (int j = 0; j < half; j++) {
                const int k = wrap_per(k0 - j, n2);
                s += a[k * es_in] * IL[2 * j + t0] + d[k * es_in] * IH[2 * j + t0];
            }
            out[r * rs_out + n * es_out] = s;
        }
    }
}

/* Multi-level 2D forward.  bands: caller-allocated [A_L, H1, V1, D1, H2, ...] like the reference's
 * d_coeffs; tmp: 2 * Nr * Nc floats.  Returns 0. */
int dwt_cpu_forward2d(const float* img, float** bands, float* tmp, int Nr, int Nc, int levels,
                      const float* L, const float* H, int F) {
    float* cur = (float*)malloc(sizeof(float) * (size_t)Nr * Nc);
    memcpy(cur, img, sizeof(float) * (size_t)Nr * Nc);
    int nr = Nr, nc = Nc;
    for (int l = 0; l < levels; l++) {
        const int nr2 = (nr + 1) / 2, nc2 = (nc + 1) / 2;
        float* lo = tmp;
        float* hi = tmp + (size_t)nr * nc2;
        analysis(cur, lo, hi, nr, nc, nc, 1, nc2, 1, L, H, F);                      /* rows   */
        float* A = (float*)malloc(sizeof(float) * (size_t)nr2 * nc2);
        analysis(lo, A, bands[3 * l + 1], nc2, nr, 1, nc2, 1, nc2, L, H, F);         /* columns of lo: A, H */
        analysis(hi, bands[3 * l + 2], bands[3 * l + 3], nc2, nr, 1, nc2, 1, nc2, L, H, F); /* of hi: V, D */
        free(cur);
        cur = A;
        nr = nr2;
        nc = nc2;
    }
    memcpy(bands[0], cur, sizeof(float) * (size_t)nr * nc);
    free(cur);
    return 0;
}

int dwt_cpu_inverse2d(float* img, float** bands, float* tmp, int Nr, int Nc, int levels, const float* IL,
                      const float* IH, int F) {
    int tnr[64], tnc[64];
    tnr[0] = Nr;
    tnc[0] = Nc;
    for (int l = 1; l <= levels; l++) {
        tnr[l] = (tnr[l - 1] + 1) / 2;
        tnc[l] = (tnc[l - 1] + 1) / 2;
    }
    float* cur = (float*)malloc(sizeof(float) * (size_t)tnr[levels] * tnc[levels]);
    memcpy(cur, bands[0], sizeof(float) * (size_t)tnr[levels] * tnc[levels]);
    for (int l = levels; l >= 1; l--) {
        const int nr = tnr[l], nc = tnc[l], Nro = tnr[l - 1], Nco = tnc[l - 1];
        float* t1 = tmp;
        float* t2 = tmp + (size_t)Nro * nc;
        synthesis(cur, bands[3 * (l - 1) + 1], t1, nc, nr, Nro, 1, nc, 1, nc, IL, IH, F);      /* columns: A,H */
        synthesis(bands[3 * (l - 1) + 2], bands[3 * (l - 1) + 3], t2, nc, nr, Nro, 1, nc, 1, nc, IL, IH, F);
        float* out = (float*)malloc(sizeof(float) * (size_t)Nro * Nco);
        synthesis(t1, t2, out, Nro, nc, Nco, nc, 1, Nco, 1, IL, IH, F);                         /* rows */
        free(cur);
        cur = out;
    }
    memcpy(img, cur, sizeof(float) * (size_t)Nr * Nc);
    free(cur);
    return 0;
}

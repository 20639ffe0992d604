/*
 * svad_oracle.c -- CPU restatement of the Silero-VAD v6.2.1 per-chunk forward pass.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity oracle and the CPU baseline of
 * bench.py; nothing under silero_vad_b200/ may import, link or call it.  It follows the
 * reference's DENSE algorithm (conv-basis STFT, all conv taps, fp32 everywhere), i.e. what
 * the shipped TorchScript graph computes, not the FFT/dead-tap-skipping form the CUDA
 * kernels use -- so agreement between the two is evidence, not tautology.
 *
 * Pinning: the reference's own tests assert only `is not None` (tests/test_basic.py:10-22,
 * "parity unpinned" by the reference); this restatement is pinned instead against outputs
 * of the reference TorchScript model itself, generated in the build container by
 * oracle/gen_golden.py and committed under tests/golden/ (see tests/test_oracle.py).
 *
 * Reference citations (paths relative to /root/reference):
 *   wrapper / context / state protocol ... src/silero_vad/utils_vad.py:57-92 (OnnxWrapper.__call__)
 *                                          and silero_vad.jit::forward
 *   audio_forward (reset, zero-pad tail) . src/silero_vad/utils_vad.py:94-110
 *   STFT (reflect pad, conv basis, |.|) .. silero_vad.jit::_model.stft.transform_ ;
 *                                          examples/onnx_sequence/export.py:22-44
 *   encoder 4 x (Conv1d k3 + ReLU) ....... silero_vad.jit::_model.encoder.{0..3} ;
 *                                          examples/onnx_sequence/export.py:52-59,70-71 ;
 *                                          src/silero_vad/tinygrad_model.py:19-22,61-64
 *   LSTM cell (gate order i,f,g,o) ....... silero_vad.jit::_model.decoder.forward -> torch.lstm_cell ;
 *                                          tuning/utils.py:163-184
 *   head ReLU -> Conv1d(128,1,1) -> sigmoid -> mean over length-1 axis
 *                                          silero_vad.jit::_model.decoder.decoder.{0..3}, _model.forward
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define HID 128
#define GATES 512
#define SB 8 /* streams processed together per weight pass (cache blocking only) */

typedef struct {
    int sr, n, ctx, nfft, hop, pad, F; /* 16k: 512,64,256,128,64,129 ; 8k: 256,32,128,64,32,65 */
    float *basisT;                     /* [nfft][2F]      (k-major transpose of basis[2F][1][nfft]) */
    float *w0T, *b0;                   /* [3][F][128]     tap-major: w0T[j][c][o] = W0[o][c][j]     */
    float *w1T, *b1;                   /* [3][128][64] */
    float *w2T, *b2;                   /* [3][64][64]  */
    float *w3T, *b3;                   /* [3][64][128] */
    float *wihT, *whhT, *bih, *bhh;    /* [128][512] each ; [512] */
    float *wout, bout;                 /* [128] */
} branch_t;

typedef struct svad_oracle {
    branch_t br[2]; /* 0: 16 kHz, 1: 8 kHz */
} svad_oracle_t;

/* ---------------------------------------------------------------- container reader */
typedef struct { char name[128]; int ndim; uint32_t dims[4]; float *data; size_t numel; } tensor_t;

static int read_container(const char *path, tensor_t **out, int *n_out) {
    FILE *f = fopen(path, "rb");
    if (!f) return -1;
    char magic[8];
    uint32_t n;
    if (fread(magic, 1, 8, f) != 8 || memcmp(magic, "SVADW001", 8) || fread(&n, 4, 1, f) != 1) { fclose(f); return -2; }
    tensor_t *ts = (tensor_t *)calloc(n, sizeof(tensor_t));
    for (uint32_t i = 0; i < n; i++) {
        uint32_t nl, nd;
        if (fread(&nl, 4, 1, f) != 1 || nl >= sizeof(ts[i].name)) { fclose(f); return -3; }
        if (fread(ts[i].name, 1, nl, f) != nl) { fclose(f); return -3; }
        ts[i].name[nl] = 0;
        if (fread(&nd, 4, 1, f) != 1 || nd > 4) { fclose(f); return -3; }
        ts[i].ndim = (int)nd;
        ts[i].numel = 1;
        for (uint32_t d = 0; d < nd; d++) {
            if (fread(&ts[i].dims[d], 4, 1, f) != 1) { fclose(f); return -3; }
            ts[i].numel *= ts[i].dims[d];
        }
        ts[i].data = (float *)malloc(ts[i].numel * sizeof(float));
        if (fread(ts[i].data, sizeof(float), ts[i].numel, f) != ts[i].numel) { fclose(f); return -4; }
    }
    fclose(f);
    *out = ts;
    *n_out = (int)n;
    return 0;
}

static const tensor_t *find(const tensor_t *ts, int n, const char *prefix, const char *suffix) {
    char full[192];
    snprintf(full, sizeof full, "%s%s", prefix, suffix);
    for (int i = 0; i < n; i++)
        if (!strcmp(ts[i].name, full)) return &ts[i];
    return NULL;
}

/* conv weight [O][C][3] -> tap-major transposed [3][C][O] */
static float *conv_T(const tensor_t *t) {
    int O = (int)t->dims[0], C = (int)t->dims[1];
    float *r = (float *)malloc(sizeof(float) * 3 * C * O);
    for (int o = 0; o < O; o++)
        for (int c = 0; c < C; c++)
            for (int j = 0; j < 3; j++) r[((size_t)j * C + c) * O + o] = t->data[((size_t)o * C + c) * 3 + j];
    return r;
}
/* [R][K] -> [K][R] */
static float *mat_T(const float *a, int R, int K) {
    float *r = (float *)malloc(sizeof(float) * R * K);
    for (int i = 0; i < R; i++)
        for (int k = 0; k < K; k++) r[(size_t)k * R + i] = a[(size_t)i * K + k];
    return r;
}
static float *dup(const tensor_t *t) {
    float *r = (float *)malloc(sizeof(float) * t->numel);
    memcpy(r, t->data, sizeof(float) * t->numel);
    return r;
}

static int build_branch(branch_t *b, int sr, const tensor_t *w, int nw, const tensor_t *bs, int nbs) {
    const char *p = sr == 16000 ? "_model." : "_model_8k.";
    b->sr = sr;
    b->n = sr == 16000 ? 512 : 256;
    b->ctx = b->n / 8;      /* context_size_samples: 64 / 32 */
    b->nfft = b->n / 2;     /* filter_length: 256 / 128 */
    b->hop = b->nfft / 2;   /* hop_length: 128 / 64 */
    b->pad = b->nfft / 4;   /* ReflectionPad1d((0, 64 / 32)) */
    b->F = b->nfft / 2 + 1; /* 129 / 65 */
    const tensor_t *basis = find(bs, nbs, p, "stft.forward_basis_buffer");
    const tensor_t *w0 = find(w, nw, p, "encoder.0.reparam_conv.weight"), *b0 = find(w, nw, p, "encoder.0.reparam_conv.bias");
    const tensor_t *w1 = find(w, nw, p, "encoder.1.reparam_conv.weight"), *b1 = find(w, nw, p, "encoder.1.reparam_conv.bias");
    const tensor_t *w2 = find(w, nw, p, "encoder.2.reparam_conv.weight"), *b2 = find(w, nw, p, "encoder.2.reparam_conv.bias");
    const tensor_t *w3 = find(w, nw, p, "encoder.3.reparam_conv.weight"), *b3 = find(w, nw, p, "encoder.3.reparam_conv.bias");
    const tensor_t *wih = find(w, nw, p, "decoder.rnn.weight_ih"), *whh = find(w, nw, p, "decoder.rnn.weight_hh");
    const tensor_t *bih = find(w, nw, p, "decoder.rnn.bias_ih"), *bhh = find(w, nw, p, "decoder.rnn.bias_hh");
    const tensor_t *wo = find(w, nw, p, "decoder.decoder.2.weight"), *bo = find(w, nw, p, "decoder.decoder.2.bias");
    if (!basis || !w0 || !b0 || !w1 || !b1 || !w2 || !b2 || !w3 || !b3 || !wih || !whh || !bih || !bhh || !wo || !bo) return -1;
    if ((int)basis->dims[0] != 2 * b->F || (int)basis->dims[2] != b->nfft || (int)w0->dims[1] != b->F) return -2;
    b->basisT = mat_T(basis->data, 2 * b->F, b->nfft);
    b->w0T = conv_T(w0); b->b0 = dup(b0);
    b->w1T = conv_T(w1); b->b1 = dup(b1);
    b->w2T = conv_T(w2); b->b2 = dup(b2);
    b->w3T = conv_T(w3); b->b3 = dup(b3);
    b->wihT = mat_T(wih->data, GATES, HID);
    b->whhT = mat_T(whh->data, GATES, HID);
    b->bih = dup(bih); b->bhh = dup(bhh);
    b->wout = dup(wo); b->bout = bo->data[0];
    return 0;
}

static void free_tensors(tensor_t *t, int n) {
    for (int i = 0; i < n; i++) free(t[i].data);
    free(t);
}

svad_oracle_t *svad_oracle_load(const char *weights_path, const char *basis_path) {
    tensor_t *w = NULL, *bs = NULL;
    int nw = 0, nbs = 0;
    if (read_container(weights_path, &w, &nw)) return NULL;
    if (read_container(basis_path, &bs, &nbs)) { free_tensors(w, nw); return NULL; }
    svad_oracle_t *o = (svad_oracle_t *)calloc(1, sizeof *o);
    int rc = build_branch(&o->br[0], 16000, w, nw, bs, nbs) || build_branch(&o->br[1], 8000, w, nw, bs, nbs);
    free_tensors(w, nw);
    free_tensors(bs, nbs);
    if (rc) { free(o); return NULL; }
    return o;
}

void svad_oracle_free(svad_oracle_t *o) {
    if (!o) return;
    for (int i = 0; i < 2; i++) {
        branch_t *b = &o->br[i];
        free(b->basisT); free(b->w0T); free(b->b0); free(b->w1T); free(b->b1); free(b->w2T); free(b->b2);
        free(b->w3T); free(b->b3); free(b->wihT); free(b->whhT); free(b->bih); free(b->bhh); free(b->wout);
    }
    free(o);
}

/* ---------------------------------------------------------------- arithmetic */

/* Y[m][0..N) += sum_k X_m[k] * Wt[k][0..N)  for m < M; a NULL row is an all-zero row (conv zero padding,
 * silero_vad.jit::...reparam_conv: padding=[1]).  Register-blocked MB x NB tile kept in vector registers over
 * the whole k loop (gcc vectorises the j loops), so the kernel runs near the FMA rate instead of the L1 rate. */
#define MB 4
#define NB 32
static const float zero_row[640] = {0};
static void gemm_rows(int M, int N, int K, const float *const *xrows, const float *Wt, float *const *yrows) {
    for (int m0 = 0; m0 < M; m0 += MB) {
        const float *x[MB];
        float *y[MB];
        float sink[NB];
        const int mb = M - m0 < MB ? M - m0 : MB;
        int any = 0;
        for (int m = 0; m < MB; m++) {
            x[m] = (m < mb && xrows[m0 + m]) ? xrows[m0 + m] : zero_row;
            y[m] = (m < mb) ? yrows[m0 + m] : NULL;
            any |= (m < mb && xrows[m0 + m] != NULL);
        }
        if (!any) continue;
        int n0 = 0;
        for (; n0 + NB <= N; n0 += NB) {
            float acc[MB][NB];
            for (int m = 0; m < MB; m++)
                for (int j = 0; j < NB; j++) acc[m][j] = y[m] ? y[m][n0 + j] : 0.0f;
            for (int k = 0; k < K; k++) {
                const float *w = Wt + (size_t)k * N + n0;
                for (int m = 0; m < MB; m++) {
                    const float xv = x[m][k];
                    for (int j = 0; j < NB; j++) acc[m][j] += xv * w[j];
                }
            }
            for (int m = 0; m < MB; m++) {
                float *dst = y[m] ? y[m] + n0 : sink;
                for (int j = 0; j < NB; j++) dst[j] = acc[m][j];
            }
        }
        if (n0 < N) { /* ragged tail of the STFT basis (258 / 130 columns) */
            const int nb = N - n0;
            float acc[MB][NB];
            for (int m = 0; m < MB; m++)
                for (int j = 0; j < nb; j++) acc[m][j] = y[m] ? y[m][n0 + j] : 0.0f;
            for (int k = 0; k < K; k++) {
                const float *w = Wt + (size_t)k * N + n0;
                for (int m = 0; m < MB; m++) {
                    const float xv = x[m][k];
                    for (int j = 0; j < nb; j++) acc[m][j] += xv * w[j];
                }
            }
            for (int m = 0; m < MB; m++)
                if (y[m]) for (int j = 0; j < nb; j++) y[m][n0 + j] = acc[m][j];
        }
    }
}

static inline float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

typedef struct {
    float xp[SB][640];
    float spec[SB * 4][258];
    float mag[SB][4][132];
    float e0[SB][4][128];
    float e1[SB][2][64];
    float e2[SB][64];
    float e3[SB][128];
    float gates[SB][GATES];
} scratch_t;

/* One chunk step for nb (<= SB) streams.
 *   x1   : nb pointers to ctx+n contiguous samples (context already prepended)
 *   h, c : nb pointers to 128 floats each, updated in place
 *   prob : nb outputs */
static void step_block(const branch_t *b, int nb, const float *const *x1, float *const *h, float *const *c, float *prob,
                       scratch_t *s) {
    const int L = b->ctx + b->n, F = b->F, N2 = 2 * b->F;
    const float *xr[4 * SB];
    float *yr[4 * SB];

    /* reflect-pad right: xp[L+i] = x1[L-2-i]  (ReflectionPad1d, silero_vad.jit::_model.stft.padding) */
    for (int i = 0; i < nb; i++) {
        memcpy(s->xp[i], x1[i], sizeof(float) * L);
        for (int k = 0; k < b->pad; k++) s->xp[i][L + k] = x1[i][L - 2 - k];
    }
    /* STFT as strided conv with the basis: spec[f][r] = sum_m basis[r][m] * xp[hop*f + m] */
    for (int i = 0; i < nb; i++)
        for (int f = 0; f < 4; f++) {
            xr[i * 4 + f] = s->xp[i] + b->hop * f;
            yr[i * 4 + f] = s->spec[i * 4 + f];
            memset(s->spec[i * 4 + f], 0, sizeof(float) * N2);
        }
    gemm_rows(nb * 4, N2, b->nfft, xr, b->basisT, yr);
    for (int i = 0; i < nb; i++)
        for (int f = 0; f < 4; f++) {
            const float *sp = s->spec[i * 4 + f];
            for (int k = 0; k < F; k++) s->mag[i][f][k] = sqrtf(sp[k] * sp[k] + sp[F + k] * sp[F + k]);
        }
    /* enc0: 129/65 -> 128, k3 s1 p1, T 4 -> 4 */
    for (int i = 0; i < nb; i++)
        for (int t = 0; t < 4; t++) memcpy(s->e0[i][t], b->b0, sizeof(float) * 128);
    for (int j = 0; j < 3; j++) {
        for (int i = 0; i < nb; i++)
            for (int t = 0; t < 4; t++) {
                int ti = t + j - 1;
                xr[i * 4 + t] = (ti < 0 || ti > 3) ? NULL : s->mag[i][ti];
                yr[i * 4 + t] = s->e0[i][t];
            }
        gemm_rows(nb * 4, 128, F, xr, b->w0T + (size_t)j * F * 128, yr);
    }
    for (int i = 0; i < nb; i++)
        for (int t = 0; t < 4; t++)
            for (int o = 0; o < 128; o++) s->e0[i][t][o] = fmaxf(s->e0[i][t][o], 0.0f);
    /* enc1: 128 -> 64, k3 s2 p1, T 4 -> 2 */
    for (int i = 0; i < nb; i++)
        for (int t = 0; t < 2; t++) memcpy(s->e1[i][t], b->b1, sizeof(float) * 64);
    for (int j = 0; j < 3; j++) {
        for (int i = 0; i < nb; i++)
            for (int t = 0; t < 2; t++) {
                int ti = 2 * t + j - 1;
                xr[i * 2 + t] = (ti < 0 || ti > 3) ? NULL : s->e0[i][ti];
                yr[i * 2 + t] = s->e1[i][t];
            }
        gemm_rows(nb * 2, 64, 128, xr, b->w1T + (size_t)j * 128 * 64, yr);
    }
    for (int i = 0; i < nb; i++)
        for (int t = 0; t < 2; t++)
            for (int o = 0; o < 64; o++) s->e1[i][t][o] = fmaxf(s->e1[i][t][o], 0.0f);
    /* enc2: 64 -> 64, k3 s2 p1, T 2 -> 1 */
    for (int i = 0; i < nb; i++) memcpy(s->e2[i], b->b2, sizeof(float) * 64);
    for (int j = 0; j < 3; j++) {
        for (int i = 0; i < nb; i++) {
            int ti = j - 1;
            xr[i] = (ti < 0 || ti > 1) ? NULL : s->e1[i][ti];
            yr[i] = s->e2[i];
        }
        gemm_rows(nb, 64, 64, xr, b->w2T + (size_t)j * 64 * 64, yr);
    }
    for (int i = 0; i < nb; i++)
        for (int o = 0; o < 64; o++) s->e2[i][o] = fmaxf(s->e2[i][o], 0.0f);
    /* enc3: 64 -> 128, k3 s1 p1, T 1 -> 1 */
    for (int i = 0; i < nb; i++) memcpy(s->e3[i], b->b3, sizeof(float) * 128);
    for (int j = 0; j < 3; j++) {
        for (int i = 0; i < nb; i++) {
            int ti = j - 1;
            xr[i] = (ti != 0) ? NULL : s->e2[i];
            yr[i] = s->e3[i];
        }
        gemm_rows(nb, 128, 64, xr, b->w3T + (size_t)j * 64 * 128, yr);
    }
    for (int i = 0; i < nb; i++)
        for (int o = 0; o < 128; o++) s->e3[i][o] = fmaxf(s->e3[i][o], 0.0f);
    /* LSTM cell: gates = W_ih x + b_ih + W_hh h + b_hh ; order i, f, g, o */
    for (int i = 0; i < nb; i++) {
        for (int g = 0; g < GATES; g++) s->gates[i][g] = b->bih[g] + b->bhh[g];
        xr[i] = s->e3[i];
        yr[i] = s->gates[i];
    }
    gemm_rows(nb, GATES, HID, xr, b->wihT, yr);
    for (int i = 0; i < nb; i++) xr[i] = h[i];
    gemm_rows(nb, GATES, HID, xr, b->whhT, yr);
    for (int i = 0; i < nb; i++) {
        const float *g = s->gates[i];
        float acc = b->bout;
        for (int j = 0; j < HID; j++) {
            float ig = sigmoidf_(g[j]), fg = sigmoidf_(g[HID + j]), gg = tanhf(g[2 * HID + j]), og = sigmoidf_(g[3 * HID + j]);
            float cn = fg * c[i][j] + ig * gg;
            float hn = og * tanhf(cn);
            c[i][j] = cn;
            h[i][j] = hn;
            acc += b->wout[j] * fmaxf(hn, 0.0f); /* Dropout(eval)=id -> ReLU -> Conv1d(128,1,1) */
        }
        prob[i] = sigmoidf_(acc); /* mean over the length-1 time axis is the identity */
    }
}

static const branch_t *pick(const svad_oracle_t *o, int sr) {
    if (sr == 16000) return &o->br[0];
    if (sr == 8000) return &o->br[1];
    return NULL;
}

/* Stateless step, the ONNX contract (utils_vad.py:80-82; examples/cpp/silero-vad-onnx.cpp:176-195):
 *   input f32[B, ctx+n], state f32[2,B,128] -> out f32[B], state_out f32[2,B,128]. */
int svad_oracle_step(const svad_oracle_t *o, int sr, int B, const float *input, const float *state_in, float *out,
                     float *state_out, int nthreads) {
    const branch_t *b = pick(o, sr);
    if (!b || B < 0) return -1;
    const int L = b->ctx + b->n;
    if (state_out != state_in) memcpy(state_out, state_in, sizeof(float) * 2 * (size_t)B * HID);
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#else
    (void)nthreads;
#endif
#pragma omp parallel
    {
        scratch_t *s = (scratch_t *)malloc(sizeof *s);
#pragma omp for schedule(static)
        for (int b0 = 0; b0 < B; b0 += SB) {
            int nb = B - b0 < SB ? B - b0 : SB;
            const float *x1[SB];
            float *h[SB], *c[SB];
            for (int i = 0; i < nb; i++) {
                x1[i] = input + (size_t)(b0 + i) * L;
                h[i] = state_out + (size_t)(b0 + i) * HID;
                c[i] = state_out + ((size_t)B + b0 + i) * HID;
            }
            step_block(b, nb, x1, h, c, out + b0, s);
        }
        free(s);
    }
    return 0;
}

/* Bulk path = audio_forward (utils_vad.py:94-110) with explicit carried state:
 *   audio f32[B, L] row stride `ld`; the tail is zero-padded to a multiple of n (utils_vad.py:100-102);
 *   state f32[2,B,128] and context f32[B,ctx] are read and updated in place (NULL = start from zeros,
 *   i.e. what reset_states() gives); probs f32[B, T], T = ceil(L/n). */
int svad_oracle_audio_forward(const svad_oracle_t *o, int sr, int B, long L, long ld, const float *audio, float *state,
                              float *context, float *probs, int nthreads) {
    const branch_t *b = pick(o, sr);
    if (!b || B < 0 || L < 0) return -1;
    const int n = b->n, ctx = b->ctx;
    const long T = (L + n - 1) / n;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#else
    (void)nthreads;
#endif
#pragma omp parallel
    {
        scratch_t *s = (scratch_t *)malloc(sizeof *s);
        float(*win)[576] = (float(*)[576])malloc(sizeof(float) * SB * 576);
        float(*hh)[HID] = (float(*)[HID])malloc(sizeof(float) * SB * HID);
        float(*cc)[HID] = (float(*)[HID])malloc(sizeof(float) * SB * HID);
        float pr[SB];
#pragma omp for schedule(dynamic, 1)
        for (int b0 = 0; b0 < B; b0 += SB) {
            int nb = B - b0 < SB ? B - b0 : SB;
            const float *x1[SB];
            float *h[SB], *c[SB];
            for (int i = 0; i < nb; i++) {
                x1[i] = win[i]; h[i] = hh[i]; c[i] = cc[i];
                if (state) {
                    memcpy(hh[i], state + (size_t)(b0 + i) * HID, sizeof(float) * HID);
                    memcpy(cc[i], state + ((size_t)B + b0 + i) * HID, sizeof(float) * HID);
                } else {
                    memset(hh[i], 0, sizeof(float) * HID);
                    memset(cc[i], 0, sizeof(float) * HID);
                }
                if (context) memcpy(win[i], context + (size_t)(b0 + i) * ctx, sizeof(float) * ctx);
                else memset(win[i], 0, sizeof(float) * ctx);
            }
            for (long t = 0; t < T; t++) {
                for (int i = 0; i < nb; i++) {
                    const float *src = audio + (size_t)(b0 + i) * ld + t * n;
                    long avail = L - t * n;
                    if (avail >= n) memcpy(win[i] + ctx, src, sizeof(float) * n);
                    else {
                        memcpy(win[i] + ctx, src, sizeof(float) * avail);
                        memset(win[i] + ctx + avail, 0, sizeof(float) * (n - avail));
                    }
                }
                step_block(b, nb, x1, h, c, pr, s);
                for (int i = 0; i < nb; i++) {
                    probs[(size_t)(b0 + i) * T + t] = pr[i];
                    memmove(win[i], win[i] + n, sizeof(float) * ctx); /* new context = last ctx samples */
                }
            }
            for (int i = 0; i < nb; i++) {
                if (state) {
                    memcpy(state + (size_t)(b0 + i) * HID, hh[i], sizeof(float) * HID);
                    memcpy(state + ((size_t)B + b0 + i) * HID, cc[i], sizeof(float) * HID);
                }
                if (context) memcpy(context + (size_t)(b0 + i) * ctx, win[i], sizeof(float) * ctx);
            }
        }
        free(s); free(win); free(hh); free(cc);
    }
    return 0;
}

int svad_oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

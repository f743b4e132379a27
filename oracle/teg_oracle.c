/*
 * teg_oracle.c -- CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement of the reference Teg C backend for the hot path
 * (evaluation of a reduced, base-language Teg program by midpoint quadrature):
 *
 *   tree -> IR lowering         /root/reference/teg/passes/compile.py:51-216
 *   per-op C statements         /root/reference/teg/ir/passes/to_c.py:260-386, c_math.py:34-74
 *   real if/else guards         /root/reference/teg/ir/passes/to_c.py:459-487
 *   midpoint rule               /root/reference/data/templates/rectangular_rule.template.c:1-11
 *   vector (Tup) arithmetic     /root/reference/teg/include/teg_generic_array.h:8-69
 *   one evaluation per call     /root/reference/teg/eval/c_eval.py:129-157
 *
 * Two further quadrature modes are restated here for the backend's `integration_mode` option: the numpy backend's
 * trapezoid rule (/root/reference/teg/eval/numpy_eval.py:57-68; pinned to committed outputs of that backend) and
 * `monte_carlo_uniform`, which the reference names (to_c.py:389-393) but never defines -- PARITY UNPINNED for that
 * mode: only its generator (Philox4x32-10) is checked against published known-answer vectors.
 *
 * Instead of generating C text and compiling it per program, this file *interprets* the
 * program (TEGP text, see teg_b200/tegp.py for the grammar), performing exactly the IEEE
 * operations the generated C performs, in working precision (float or double), with the
 * reference's sample placement and left-fold accumulation.  g++ -O3 without -ffast-math and
 * without -mfma neither re-associates nor contracts, so the interpreter is expected to be
 * BIT-IDENTICAL to the reference's compiled output for + - * / sqrt (libm transcendentals are
 * the same glibc routines).  That expectation is pinned by tests/test_oracle_pinning.py
 * against (a) the reference's own generated C compiled from /root/reference into
 * oracle/_ref/ and (b) golden vectors recorded from the reference C backend
 * (tests/golden/).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library.  The product (teg_b200/) never does.
 *
 * Values are memoised per node with logical time stamps: a memo is valid until any variable
 * the node mentions is re-bound.  This is pure caching of a pure function (it makes the
 * interpreter skip loop-invariant recomputation, as g++ -O3 does for the generated C) and
 * cannot change any result.
 *
 * Build: gcc -O2 -std=c99 -fPIC -shared [-fopenmp] teg_oracle.c -o libteg_oracle.so -lm
 *        (-ffp-contract=off is passed by the recipe; x86-64 without -mfma has no FMA anyway)
 */
#include <ctype.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define TEGO_MAXV 16
#define TEGO_MAXLET 16

enum { FN_SQR = 1, FN_SQRT, FN_SIN, FN_COS, FN_ASIN, FN_ATAN2, FN_EXP };

typedef struct {
    char op;
    int nk;
    int *kids;
    float cval_f;       /* 'C': strtof of the decimal literal  (what `<lit>f` means to the C compiler) */
    double cval_d;      /* 'C': strtod of the decimal literal */
    int label;          /* 'V': label id ; 'T': integration-variable label id */
    int fn;             /* 'F' */
    int nlet;           /* 'L': number of bindings */
    int *let_labels;
    int nfree;          /* free variable label ids (sorted, unique) */
    int *free;
} Node;

typedef struct Prog {
    int nnodes;
    Node *nodes;
    int root;
    int nlabels;
    char **labels;
    int nargs;
    int *arg_labels;
} Prog;

static __thread char tego_err[256];

static void tego_set_error(const char *msg) {
    snprintf(tego_err, sizeof tego_err, "%s", msg ? msg : "unknown error");
}

const char *tego_last_error(void) { return tego_err; }

static int label_id(Prog *p, const char *name) {
    for (int i = 0; i < p->nlabels; i++)
        if (strcmp(p->labels[i], name) == 0) return i;
    p->labels = (char **)realloc(p->labels, sizeof(char *) * (p->nlabels + 1));
    p->labels[p->nlabels] = strdup(name);
    return p->nlabels++;
}

static int fn_id(const char *s) {
    if (!strcmp(s, "Sqr")) return FN_SQR;
    if (!strcmp(s, "Sqrt")) return FN_SQRT;
    if (!strcmp(s, "Sin")) return FN_SIN;
    if (!strcmp(s, "Cos")) return FN_COS;
    if (!strcmp(s, "ASin")) return FN_ASIN;
    if (!strcmp(s, "ATan2")) return FN_ATAN2;
    if (!strcmp(s, "Exp")) return FN_EXP;
    return 0;
}

/* sorted-unique union of a and b, minus the labels in `drop` */
static int *set_union(const int *a, int na, const int *b, int nb, int *nout) {
    int *r = (int *)malloc(sizeof(int) * (na + nb + 1));
    int i = 0, j = 0, n = 0;
    while (i < na || j < nb) {
        int v;
        if (j >= nb || (i < na && a[i] < b[j])) v = a[i++];
        else if (i >= na || b[j] < a[i]) v = b[j++];
        else { v = a[i]; i++; j++; }
        r[n++] = v;
    }
    *nout = n;
    return r;
}

static int *set_minus(const int *a, int na, const int *drop, int nd, int *nout) {
    int *r = (int *)malloc(sizeof(int) * (na + 1));
    int n = 0;
    for (int i = 0; i < na; i++) {
        int keep = 1;
        for (int j = 0; j < nd; j++) if (drop[j] == a[i]) keep = 0;
        if (keep) r[n++] = a[i];
    }
    *nout = n;
    return r;
}

void tego_free(Prog *p) {
    if (!p) return;
    for (int i = 0; i < p->nnodes; i++) { free(p->nodes[i].kids); free(p->nodes[i].let_labels); free(p->nodes[i].free); }
    for (int i = 0; i < p->nlabels; i++) free(p->labels[i]);
    free(p->nodes); free(p->labels); free(p->arg_labels); free(p);
}

static char *next_tok(char **cur) {
    char *s = *cur;
    while (*s && isspace((unsigned char)*s)) s++;
    if (!*s) return NULL;
    char *start = s;
    while (*s && !isspace((unsigned char)*s)) s++;
    if (*s) { *s = 0; s++; }
    *cur = s;
    return start;
}

#define NEED(tok) do { if (!(tok)) { tego_set_error("truncated TEGP text"); goto fail; } } while (0)

Prog *tego_load(const char *text) {
    char *buf = strdup(text), *cur = buf, *t;
    Prog *p = (Prog *)calloc(1, sizeof(Prog));
    t = next_tok(&cur); NEED(t);
    if (strcmp(t, "TEGP")) { tego_set_error("not a TEGP program"); goto fail; }
    t = next_tok(&cur); NEED(t);
    if (strcmp(t, "1")) { tego_set_error("unsupported TEGP version"); goto fail; }
    t = next_tok(&cur); NEED(t); /* name */
    t = next_tok(&cur); NEED(t);
    t = next_tok(&cur); NEED(t); /* nodes */
    t = next_tok(&cur); NEED(t);
    p->nnodes = atoi(t);
    p->nodes = (Node *)calloc(p->nnodes > 0 ? p->nnodes : 1, sizeof(Node));
    for (int i = 0; i < p->nnodes; i++) {
        Node *nd = &p->nodes[i];
        t = next_tok(&cur); NEED(t);
        nd->op = t[0];
        int nk = 0;
        switch (nd->op) {
        case 'C':
            t = next_tok(&cur); NEED(t);
            nd->cval_f = strtof(t, NULL);
            nd->cval_d = strtod(t, NULL);
            break;
        case 'V':
            t = next_tok(&cur); NEED(t);
            nd->label = label_id(p, t);
            t = next_tok(&cur); NEED(t); /* default: unused by the oracle */
            break;
        case 'F':
            t = next_tok(&cur); NEED(t);
            nd->fn = fn_id(t);
            if (!nd->fn) { tego_set_error("unknown smooth function"); goto fail; }
            nk = 1;
            break;
        case 'A': case 'M': case 'B': case '&': case '|': nk = 2; break;
        case 'I': nk = 1; break;
        case 'S': nk = 3; break;
        case 'T': nk = 3; break;
        case 'U':
            t = next_tok(&cur); NEED(t);
            nk = atoi(t);
            break;
        case 'L': {
            t = next_tok(&cur); NEED(t);
            nd->nlet = atoi(t);
            nd->let_labels = (int *)malloc(sizeof(int) * (nd->nlet + 1));
            nd->nk = nd->nlet + 1;
            nd->kids = (int *)malloc(sizeof(int) * nd->nk);
            for (int j = 0; j < nd->nlet; j++) {
                t = next_tok(&cur); NEED(t);
                nd->let_labels[j] = label_id(p, t);
                t = next_tok(&cur); NEED(t);
                nd->kids[j] = atoi(t);
            }
            t = next_tok(&cur); NEED(t);
            nd->kids[nd->nlet] = atoi(t);
            break;
        }
        default:
            tego_set_error("unknown TEGP op");
            goto fail;
        }
        if (nd->op != 'L') {
            nd->nk = nk;
            nd->kids = (int *)malloc(sizeof(int) * (nk + 1));
            for (int j = 0; j < nk; j++) { t = next_tok(&cur); NEED(t); nd->kids[j] = atoi(t); }
            if (nd->op == 'T') { t = next_tok(&cur); NEED(t); nd->label = label_id(p, t); }
        }
        for (int j = 0; j < nd->nk; j++)
            if (nd->kids[j] < 0 || nd->kids[j] >= i) { tego_set_error("children must precede parents"); goto fail; }
    }
    t = next_tok(&cur); NEED(t); /* root */
    t = next_tok(&cur); NEED(t);
    p->root = atoi(t);
    if (p->root < 0 || p->root >= p->nnodes) { tego_set_error("bad root"); goto fail; }
    t = next_tok(&cur); NEED(t); /* args */
    t = next_tok(&cur); NEED(t);
    p->nargs = atoi(t);
    p->arg_labels = (int *)malloc(sizeof(int) * (p->nargs + 1));
    for (int j = 0; j < p->nargs; j++) { t = next_tok(&cur); NEED(t); p->arg_labels[j] = label_id(p, t); }

    /* free-variable sets, bottom-up (scoping as compile.py: Teg binds its dvar in the body only,
       LetIn binds its labels in the body only) */
    for (int i = 0; i < p->nnodes; i++) {
        Node *nd = &p->nodes[i];
        int *acc = NULL, na = 0;
        if (nd->op == 'V') {
            acc = (int *)malloc(sizeof(int)); acc[0] = nd->label; na = 1;
        } else {
            acc = (int *)malloc(sizeof(int)); na = 0;
            for (int j = 0; j < nd->nk; j++) {
                Node *kid = &p->nodes[nd->kids[j]];
                int *kf = kid->free, nkf = kid->nfree, *tmp = NULL;
                if (nd->op == 'T' && j == 2) { tmp = set_minus(kf, nkf, &nd->label, 1, &nkf); kf = tmp; }
                if (nd->op == 'L' && j == nd->nlet) { tmp = set_minus(kf, nkf, nd->let_labels, nd->nlet, &nkf); kf = tmp; }
                int nn; int *u = set_union(acc, na, kf, nkf, &nn);
                free(acc); free(tmp);
                acc = u; na = nn;
            }
        }
        nd->free = acc; nd->nfree = na;
    }
    free(buf);
    return p;
fail:
    free(buf);
    tego_free(p);
    return NULL;
}

int tego_num_args(const Prog *p) { return p->nargs; }

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)

/* ---- fp32 instantiation: float overloads of <math.h> as the generated C++ resolves them ---- */
/* Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11), the published
 * counter-based generator: key (k0, k1), counter (c0, c1, c2, c3) -> 4 x 32 random bits.  Used by the Monte Carlo
 * integration mode only; teg_b200/codegen.py emits the same text into the generated kernels. */
void tego_philox4x32_10(unsigned int k0, unsigned int k1, unsigned int c0, unsigned int c1, unsigned int c2,
                               unsigned int c3, unsigned int out[4]) {
    for (int r = 0; r < 10; r++) {
        const unsigned long long p0 = 0xD2511F53ull * c0, p1 = 0xCD9E8D57ull * c2;
        const unsigned int n0 = (unsigned int)(p1 >> 32) ^ c1 ^ k0, n1 = (unsigned int)p1;
        const unsigned int n2 = (unsigned int)(p0 >> 32) ^ c3 ^ k1, n3 = (unsigned int)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

#define T float
#define SUF(x) CAT(x, f)
#define M_SQRT sqrtf
#define M_SIN sinf
#define M_COS cosf
#define M_ASIN asinf
#define M_ATAN2 atan2f
#include "teg_oracle_impl.h"
#undef T
#undef SUF
#undef M_SQRT
#undef M_SIN
#undef M_COS
#undef M_ASIN
#undef M_ATAN2

/* ---- fp64 instantiation ---- */
#define T double
#define SUF(x) CAT(x, d)
#define M_SQRT sqrt
#define M_SIN sin
#define M_COS cos
#define M_ASIN asin
#define M_ATAN2 atan2
#include "teg_oracle_impl.h"
#undef T
#undef SUF

int tego_eval_f32(const Prog *p, const float *in, long n, int num_samples, float *out, int k, int nthreads) {
    return tego_eval_f(p, in, n, num_samples, out, k, nthreads, 0, 0, 0);
}

int tego_eval_f64(const Prog *p, const double *in, long n, int num_samples, double *out, int k, int nthreads) {
    return tego_eval_d(p, in, n, num_samples, out, k, nthreads, 0, 0, 0);
}

/* quadrature: 0 = midpoint rule of the C backend, 1 = trapezoid rule of the numpy backend (numpy_eval.py:57-68),
 * 2 = uniform Monte Carlo (seed = generator key, inst_base = index of the first instance of this call) */
int tego_eval_f32_q(const Prog *p, const float *in, long n, int num_samples, float *out, int k, int nthreads, int q,
                    unsigned long long seed, long inst_base) {
    return tego_eval_f(p, in, n, num_samples, out, k, nthreads, q, seed, inst_base);
}

int tego_eval_f64_q(const Prog *p, const double *in, long n, int num_samples, double *out, int k, int nthreads, int q,
                    unsigned long long seed, long inst_base) {
    return tego_eval_d(p, in, n, num_samples, out, k, nthreads, q, seed, inst_base);
}

/* size of the program's result (1 for scalars, k for Tup programs), probed on one dummy instance */
int tego_out_size(const Prog *p) {
    Ctxd *c = ctx_newd(p, 1, 0);
    for (int a = 0; a < p->nlabels; a++) {   /* bind everything so unbound inputs cannot fail the probe */
        c->bind_val[a].n = 1; c->bind_val[a].v[0] = 0.5; c->bind_stamp[a] = ++c->clock;
    }
    Vald r = evald(c, p->root);
    int n = c->err ? -1 : r.n;
    if (c->err) tego_set_error(c->err);
    ctx_freed(c);
    return n;
}

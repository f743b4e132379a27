/*
 * TEST INFRASTRUCTURE ONLY -- precision-generic body of the CPU oracle (see teg_oracle.c).
 * Included twice with T = float / double.  Every routine cites the reference code it restates.
 */
#ifndef T
#error "define T, SUF(name) and the libm names before including"
#endif

typedef struct { int n; T v[TEGO_MAXV]; } SUF(Val);

/* u in [0, 1) from the Philox words: 24 bits for float, 53 bits for double (exact conversions, exact scalings) */
static T SUF(tego_uniform)(unsigned long long seed, unsigned long long inst, unsigned int depth, unsigned int i) {
    unsigned int r[4];
    tego_philox4x32_10((unsigned int)seed, (unsigned int)(seed >> 32), (unsigned int)inst, (unsigned int)(inst >> 32),
                       depth, i, r);
    if (sizeof(T) == 4) return (T)(r[0] >> 8) * (T)(1.0 / 16777216.0);
    return (T)(((unsigned long long)(r[0] >> 5) << 26) | (unsigned long long)(r[1] >> 6)) * (T)(1.0 / 9007199254740992.0);
}

typedef struct {
    const Prog *p;
    int num_samples;
    int quadrature;             /* 0 = midpoint (the C backend), 1 = trapezoid on linspace (the numpy backend),
                                   2 = uniform Monte Carlo with a counter-based generator (see case 'T') */
    unsigned long long seed;    /* Monte Carlo: key of the generator */
    unsigned long long inst;    /* Monte Carlo: index of the instance being evaluated */
    int depth;                  /* Monte Carlo: nesting depth of the integral being evaluated */
    int *memo_depth;            /* depth at which a memoised value was computed (structurally shared integrals) */
    /* current binding of every label (program inputs, let variables, integration variables) */
    SUF(Val) *bind_val;
    long *bind_stamp;           /* 0 = unbound */
    /* memo per node: value + logical time at which it was computed */
    SUF(Val) *memo_val;
    unsigned char *memo_bool;
    long *memo_stamp;
    long clock;
    const char *err;
} SUF(Ctx);

static int SUF(memo_valid)(SUF(Ctx) *c, int id) {
    const Node *nd = &c->p->nodes[id];
    long st = c->memo_stamp[id];
    if (st == 0) return 0;
    if (c->quadrature == 2 && c->memo_depth[id] != c->depth) return 0;
    for (int i = 0; i < nd->nfree; i++)
        if (c->bind_stamp[nd->free[i]] >= st) return 0;
    return 1;
}

static SUF(Val) SUF(eval)(SUF(Ctx) *c, int id);

/* Bool / And / Or: reference compile.py:200-213 -> IR_CompareLT / IR_LAnd / IR_LOr, emitted as
 * `a < b`, `a && b`, `a || b` by to_c.py:298-331.  allow_eq is ignored by the reference. */
static int SUF(eval_bool)(SUF(Ctx) *c, int id) {
    const Node *nd = &c->p->nodes[id];
    if (SUF(memo_valid)(c, id)) return c->memo_bool[id];
    int r = 0;
    switch (nd->op) {
    case 'B': {
        SUF(Val) a = SUF(eval)(c, nd->kids[0]);
        SUF(Val) b = SUF(eval)(c, nd->kids[1]);
        if (a.n != 1 || b.n != 1) { c->err = "vector-valued comparison used as a condition"; return 0; }
        r = a.v[0] < b.v[0];
        break;
    }
    case '&': {
        /* both operands are materialised before `&&` in the emitted C; they are pure, so the
         * short-circuit makes no observable difference */
        int a = SUF(eval_bool)(c, nd->kids[0]);
        int b = SUF(eval_bool)(c, nd->kids[1]);
        r = a && b;
        break;
    }
    case '|': {
        int a = SUF(eval_bool)(c, nd->kids[0]);
        int b = SUF(eval_bool)(c, nd->kids[1]);
        r = a || b;
        break;
    }
    default:
        c->err = "expected a boolean expression";
        return 0;
    }
    c->memo_bool[id] = (unsigned char)r;
    c->memo_stamp[id] = ++c->clock;
    c->memo_depth[id] = c->depth;
    return r;
}

/* element-wise binary op with scalar broadcast: to_c.py:263-283 (+ teg_generic_array.h:23-69) */
static SUF(Val) SUF(binary)(SUF(Ctx) *c, char op, SUF(Val) a, SUF(Val) b) {
    SUF(Val) r;
    if (a.n != b.n && a.n != 1 && b.n != 1) { c->err = "Binary operands have incompatible sizes"; r.n = 1; r.v[0] = 0; return r; }
    r.n = a.n > b.n ? a.n : b.n;
    for (int i = 0; i < r.n; i++) {
        T x = a.v[a.n == 1 ? 0 : i], y = b.v[b.n == 1 ? 0 : i];
        r.v[i] = (op == 'A') ? x + y : (op == 'M') ? x * y : x / y;
    }
    return r;
}

static SUF(Val) SUF(eval)(SUF(Ctx) *c, int id) {
    const Node *nd = &c->p->nodes[id];
    SUF(Val) r;
    r.n = 1; r.v[0] = 0;
    if (c->err) return r;
    if (nd->op == 'V') {            /* compile.py:62-67: a label lookup, innermost binding wins */
        if (c->bind_stamp[nd->label] == 0) { c->err = "unbound variable"; return r; }
        return c->bind_val[nd->label];
    }
    if (SUF(memo_valid)(c, id)) return c->memo_val[id];
    switch (nd->op) {
    case 'C':                       /* to_c.py:231-247: `{value}f` literal in fp32, `{value}` in fp64 */
        r.v[0] = SUF(nd->cval_);
        break;
    case 'A': case 'M':             /* compile.py:185-198 -> to_c.py:310-319 */
        r = SUF(binary)(c, nd->op, SUF(eval)(c, nd->kids[0]), SUF(eval)(c, nd->kids[1]));
        break;
    case 'I': {                     /* compile.py:69-74: IR_Divide(out, 1.0, x) -> `1.0f / x` */
        SUF(Val) one; one.n = 1; one.v[0] = (T)1.0;
        r = SUF(binary)(c, 'D', one, SUF(eval)(c, nd->kids[0]));
        break;
    }
    case 'F': {                     /* compile.py:76-81 -> c_math.py:34-74 */
        SUF(Val) a = SUF(eval)(c, nd->kids[0]);
        if (nd->fn == FN_ATAN2) {
            if (a.n != 2) { c->err = "atan2() takes exactly 2 arguments"; return r; }
            r.n = 1; r.v[0] = M_ATAN2(a.v[0], a.v[1]);
            break;
        }
        r.n = a.n;
        for (int i = 0; i < a.n; i++) {
            T x = a.v[i];
            switch (nd->fn) {
            case FN_SQR:  r.v[i] = x * x; break;
            case FN_SQRT: r.v[i] = M_SQRT(x); break;
            case FN_SIN:  r.v[i] = M_SIN(x); break;
            case FN_COS:  r.v[i] = M_COS(x); break;
            case FN_ASIN: r.v[i] = M_ASIN(x); break;
            default: c->err = "Math operation has no defined C transformation"; return r;  /* Exp: to_c.py:338-341 */
            }
        }
        break;
    }
    case 'S': {                     /* compile.py:83-103 -> to_c.py:459-487: a real if/else, only the
                                       taken branch runs (this is what skips guarded delta integrals) */
        int cond = SUF(eval_bool)(c, nd->kids[0]);
        if (c->err) return r;
        r = SUF(eval)(c, cond ? nd->kids[1] : nd->kids[2]);
        break;
    }
    case 'U': {                     /* compile.py:137-152 -> to_c.py:371-386; scalars only (typing.py:207-219) */
        if (nd->nk > TEGO_MAXV) { c->err = "tuple wider than TEGO_MAXV"; return r; }
        r.n = nd->nk;
        for (int i = 0; i < nd->nk; i++) {
            SUF(Val) e = SUF(eval)(c, nd->kids[i]);
            if (e.n != 1) { c->err = "Teg IR currently does not support multidimensional packing"; return r; }
            r.v[i] = e.v[0];
        }
        if (r.n == 1) { /* Tup of one element stays scalar-sized */ }
        break;
    }
    case 'L': {                     /* compile.py:154-180: new_exprs are evaluated in the OUTER scope
                                       (parallel let), then the body with the labels bound */
        int k = nd->nlet;
        SUF(Val) vals[TEGO_MAXLET];
        SUF(Val) saved[TEGO_MAXLET];
        long saved_stamp[TEGO_MAXLET];
        if (k > TEGO_MAXLET) { c->err = "too many let bindings"; return r; }
        for (int i = 0; i < k; i++) vals[i] = SUF(eval)(c, nd->kids[i]);
        for (int i = 0; i < k; i++) {
            int lab = nd->let_labels[i];
            saved[i] = c->bind_val[lab]; saved_stamp[i] = c->bind_stamp[lab];
            c->bind_val[lab] = vals[i]; c->bind_stamp[lab] = ++c->clock;
        }
        r = SUF(eval)(c, nd->kids[k]);
        for (int i = k - 1; i >= 0; i--) {
            int lab = nd->let_labels[i];
            c->bind_val[lab] = saved[i];
            c->bind_stamp[lab] = saved_stamp[i] ? ++c->clock : 0;
        }
        break;
    }
    case 'T': {                     /* compile.py:105-135 -> to_c.py:396-456 instantiating
                                       data/templates/rectangular_rule.template.c:1-11 verbatim:
                                         step = ((T)(ub - lb)) / N;
                                         for i in [0,N): x = lb + step * (i + (T)0.5); out = out + f(x) * step; */
        SUF(Val) lo = SUF(eval)(c, nd->kids[0]);
        SUF(Val) hi = SUF(eval)(c, nd->kids[1]);
        if (lo.n != 1 || hi.n != 1) { c->err = "Bounds are of different types"; return r; }
        int lab = nd->label;
        SUF(Val) saved = c->bind_val[lab]; long saved_stamp = c->bind_stamp[lab];
        SUF(Val) out; out.n = 1; out.v[0] = 0;
        if (c->quadrature == 2) {
            /* monte_carlo_uniform: the reference names the mode (to_c.py:389-393) but ships neither a template nor
             * a definition, so the semantics are this backend's (teg_b200/codegen.py, Emitter._mc_loop, same text):
             *   u_i = philox4x32-10(key = seed, counter = (instance lo, instance hi, depth, i)) mapped to [0, 1)
             *   x_i = lb + (ub - lb) * u_i;  step = ((T)(ub - lb)) / N;  out = out + f(x_i) * step   (left fold)
             * The points depend on the instance, the nesting depth and the sample index only: integrals of the same
             * depth share their points (common random numbers), which keeps loop fusion and hoisting exact. */
            const int N = c->num_samples;
            T width = hi.v[0] - lo.v[0];
            T step = ((T)width) / N;
            int first = 1;
            const int depth = c->depth;
            for (unsigned int i = 0; i < (unsigned int)N; i++) {
                SUF(Val) x; x.n = 1;
                x.v[0] = lo.v[0] + width * SUF(tego_uniform)(c->seed, c->inst, (unsigned int)depth, i);
                c->bind_val[lab] = x; c->bind_stamp[lab] = ++c->clock;
                c->depth = depth + 1;
                SUF(Val) f = SUF(eval)(c, nd->kids[2]);
                c->depth = depth;
                if (c->err) break;
                if (first) { out.n = f.n; for (int j = 0; j < f.n; j++) out.v[j] = 0; first = 0; }
                for (int j = 0; j < f.n; j++) out.v[j] = out.v[j] + f.v[j] * step;
            }
        } else if (c->quadrature == 1) {
            /* numpy backend, /root/reference/teg/eval/numpy_eval.py:57-68:
             *   samples, step = np.linspace(lower, upper, N, retstep=True)   (N points, end points included,
             *                    x_i = i*step + lower, x_{N-1} = upper exactly)
             *   value = (1 if lower < upper else -1) * step * sum(y[:-1] + y[1:]) / 2
             * restated with a left fold of the pair sums (numpy's pairwise np.sum differs by rounding only) */
            const int N = c->num_samples;
            T step = (hi.v[0] - lo.v[0]) / (T)(N - 1);
            SUF(Val) prev; prev.n = 1; prev.v[0] = 0;
            int first = 1;
            for (int i = 0; i < N; i++) {
                SUF(Val) x; x.n = 1;
                x.v[0] = (i == N - 1) ? hi.v[0] : (T)i * step + lo.v[0];
                c->bind_val[lab] = x; c->bind_stamp[lab] = ++c->clock;
                SUF(Val) f = SUF(eval)(c, nd->kids[2]);
                if (c->err) break;
                if (first) { out.n = f.n; for (int j = 0; j < f.n; j++) out.v[j] = 0; first = 0; }
                else for (int j = 0; j < f.n; j++) out.v[j] = out.v[j] + (prev.v[j] + f.v[j]);
                prev = f;
            }
            T sgn = (lo.v[0] < hi.v[0]) ? (T)1 : (T)-1;
            for (int j = 0; j < out.n; j++) out.v[j] = ((sgn * step) * out.v[j]) / (T)2;
        } else {
        T step = ((T)(hi.v[0] - lo.v[0])) / c->num_samples;
        int first = 1;
        for (unsigned int i = 0; i < (unsigned int)c->num_samples; i++) {
            SUF(Val) x; x.n = 1;
            x.v[0] = lo.v[0] + step * (i + (T)(0.5));
            c->bind_val[lab] = x; c->bind_stamp[lab] = ++c->clock;
            SUF(Val) f = SUF(eval)(c, nd->kids[2]);
            if (c->err) break;
            if (first) { out.n = f.n; for (int j = 0; j < f.n; j++) out.v[j] = 0; first = 0; }
            for (int j = 0; j < f.n; j++) out.v[j] = out.v[j] + f.v[j] * step;
        }
        }
        c->bind_val[lab] = saved;
        c->bind_stamp[lab] = saved_stamp ? ++c->clock : 0;
        r = out;
        break;
    }
    default:
        c->err = "boolean expression used as a value";
        return r;
    }
    c->memo_val[id] = r;
    c->memo_stamp[id] = ++c->clock;
    return r;
}

static SUF(Ctx) *SUF(ctx_new)(const Prog *p, int num_samples, int quadrature) {
    SUF(Ctx) *c = (SUF(Ctx) *)calloc(1, sizeof(SUF(Ctx)));
    c->p = p; c->num_samples = num_samples; c->quadrature = quadrature;
    c->memo_depth = (int *)calloc(p->nnodes ? p->nnodes : 1, sizeof(int));
    c->bind_val = (SUF(Val) *)calloc(p->nlabels ? p->nlabels : 1, sizeof(SUF(Val)));
    c->bind_stamp = (long *)calloc(p->nlabels ? p->nlabels : 1, sizeof(long));
    c->memo_val = (SUF(Val) *)calloc(p->nnodes, sizeof(SUF(Val)));
    c->memo_bool = (unsigned char *)calloc(p->nnodes, 1);
    c->memo_stamp = (long *)calloc(p->nnodes, sizeof(long));
    return c;
}

static void SUF(ctx_free)(SUF(Ctx) *c) {
    free(c->bind_val); free(c->bind_stamp); free(c->memo_val); free(c->memo_bool); free(c->memo_stamp); free(c->memo_depth); free(c);
}

/* Batch driver.  in: SoA [n_args][n] (argument a of instance i at in[a*n+i]); out: SoA [k][n].
 * Plays the role of C_EvalMode.eval (c_eval.py:129-157) called once per instance, minus the
 * process spawn and the 6-digit stdout round trip. */
int SUF(tego_eval_)(const Prog *p, const T *in, long n, int num_samples, T *out, int k, int nthreads,
                    int quadrature, unsigned long long seed, long inst_base) {
    int failed = 0;
    const char *msg = NULL;
    if (nthreads < 1) nthreads = 1;
#ifdef _OPENMP
#pragma omp parallel num_threads(nthreads)
#endif
    {
        SUF(Ctx) *c = SUF(ctx_new)(p, num_samples, quadrature);
        c->seed = seed;
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
        for (long i = 0; i < n; i++) {
            if (failed) continue;
            c->inst = (unsigned long long)(inst_base + i);
            c->depth = 0;
            for (int a = 0; a < p->nargs; a++) {
                int lab = p->arg_labels[a];
                c->bind_val[lab].n = 1;
                c->bind_val[lab].v[0] = in[(long)a * n + i];
                c->bind_stamp[lab] = ++c->clock;
            }
            SUF(Val) r = SUF(eval)(c, p->root);
            if (!c->err && r.n != k) c->err = "output size mismatch";
            if (c->err) {
#ifdef _OPENMP
#pragma omp critical
#endif
                { failed = 1; msg = c->err; }
                continue;
            }
            for (int j = 0; j < k; j++) out[(long)j * n + i] = r.v[j];
        }
        SUF(ctx_free)(c);
    }
    if (failed) { tego_set_error(msg); return 1; }
    return 0;
}

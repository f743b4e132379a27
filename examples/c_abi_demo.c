/*
 * c_abi_demo.c -- libtegb200 from plain C: no Python, no torch, no CUDA headers.
 *
 *   python -m teg_b200 -c -t CUDA -f single -s 50 -m readme -o readme.cu prog.py     (writes readme.cu + readme_desc.h)
 *   gcc -std=c99 -I include -I . examples/c_abi_demo.c -o demo -L teg_b200 -ltegb200 -Wl,-rpath,teg_b200
 *   ./demo readme.cu
 *
 * Evaluates the README program  int_0^1 [x < t] dx  for 1000 values of t through tegb_compile /
 * tegb_program_create / tegb_program_eval_host and checks the closed form clamp(t, 0, 1) to midpoint error.
 * This is what the reference's generated main() + one process per evaluation (teg/eval/c_eval.py:58-157) becomes.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "readme_desc.h"
#include "tegb200.h"

int main(int argc, char **argv) {
    if (argc < 2) { fprintf(stderr, "usage: %s readme.cu\n", argv[0]); return 2; }
    FILE *f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 2; }
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    char *src = (char *)malloc(sz + 1);
    if (fread(src, 1, sz, f) != (size_t)sz) return 2;
    src[sz] = 0;
    fclose(f);

    tegb_module *mod = NULL;
    tegb_program *prog = NULL;
    if (tegb_compile(src, "sm_100a", NULL, &mod)) { fprintf(stderr, "compile: %s\n", tegb_last_error()); return 1; }
    if (tegb_program_create(mod, &readme_desc, 0, &prog)) { fprintf(stderr, "create: %s\n", tegb_last_error()); return 1; }

    enum { N = 1000 };
    float t[N], out[N];
    for (int i = 0; i < N; i++) t[i] = -0.25f + 1.5f * (float)i / (N - 1);
    const void *in_ptrs[1] = {t};
    void *out_ptrs[1] = {out};
    if (tegb_program_eval_host(prog, NULL, in_ptrs, out_ptrs, N)) { fprintf(stderr, "eval: %s\n", tegb_last_error()); return 1; }

    double worst = 0, sum = 0;
    for (int i = 0; i < N; i++) {
        double want = t[i] < 0 ? 0 : (t[i] > 1 ? 1 : t[i]);
        double e = fabs(out[i] - want);
        if (e > worst) worst = e;
        sum += out[i];
    }
    printf("c_abi_demo: %d instances, checksum %.6f, max |error| vs clamp(t,0,1) = %.4g, launches = %lld\n",
           N, sum, worst, (long long)tegb_launch_count());
    tegb_program_destroy(prog);
    tegb_module_destroy(mod);
    free(src);
    return worst <= 0.5 / 50 + 1e-6 ? 0 : 1;
}

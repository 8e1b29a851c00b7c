/* oracle/ref_main.c -- TEST INFRASTRUCTURE: `longcallD call` with the UNMODIFIED reference linked as a shared library
 * (oracle/_ref/liblcdref.so), so that the functions of the per-region worker stay interposable: the end-to-end parity
 * test preloads longcalld_b200/dropin/liblcd_dropin.so and compares the VCF with the reference's own. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
int call_var_main(int argc, char *argv[]);
extern char *CMD;
int main(int argc, char *argv[]) {
    size_t n = 0;
    for (int i = 0; i < argc; ++i) n += strlen(argv[i]) + 1;
    CMD = (char*)calloc(n + 1, 1);
    for (int i = 0; i < argc; ++i) { if (i) strcat(CMD, " "); strcat(CMD, argv[i]); }
    if (argc < 2 || strcmp(argv[1], "call") != 0) { fprintf(stderr, "usage: %s call [options] ref.fa in.bam\n", argv[0]); return 1; }
    return call_var_main(argc - 1, argv + 1);
}

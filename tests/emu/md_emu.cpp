// tests/emu/md_emu.cpp -- TEST INFRASTRUCTURE: runs the product's MD-tag walk (longcalld_b200/csrc/md_device.cuh: count pass, then fill pass, as
// the kernels do) on the host for one read.  Same signature as the oracle's lcd_oracle_md_to_eqx.
#include "cuda_emu.h"
#include "../../longcalld_b200/csrc/md_device.cuh"

using namespace lcd::md;

extern "C" long long emu_md_to_eqx(int n_cigar, const uint32_t *cigar, const char *md, uint32_t *out, long long cap) {
    uint8_t active = 1; int32_t nc0 = n_cigar, nc = 0, status = 0; long long off0 = 0, md_off = 0, cnt[2] = {0, 0}, first[2] = {0, 0}, off = 0;
    KernelArgs a; memset(&a, 0, sizeof(a));
    a.n_reads_total = 1; a.read_active = &active; a.n_cigar0 = &nc0; a.cigar_off0 = &off0; a.cigar0 = cigar; a.md_off = &md_off; a.md = md;
    a.cnt = cnt; a.first = first; a.n_cigar = &nc; a.cigar_off = &off; a.cigar = out; a.status = &status;
    count_read(a, 0);
    if (status == MD_MISMATCH) return -2;
    if (status == MD_EQX_OP) return -3;
    if (cnt[0] > cap) return -1;
    first[1] = cnt[0];
    fill_read(a, 0);
    return nc;
}

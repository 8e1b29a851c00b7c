/* oracle/md.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * The MD-tag variant of the difference-list pass, collect_digar_from_MD_tag (reference src/bam_utils.c:1003-1174), differs from the =/X
 * variant (:701-841, oracle/digar.c) only in where the '=' runs and mismatches of an 'M' op come from: the walk over the MD string
 * (:1037-1083), including its handling of a run that continues over an insertion (last_eq_len), of "0" between two mismatches and after a
 * deletion, and of zero-length runs.  This file restates that walk as a conversion of (CIGAR with M, MD) into the equivalent =/X CIGAR --
 * one '=' op per piece the reference emits, one 1-base 'X' op per mismatch -- so that MD-tagged reads go through the =/X path.
 * Pinned in tests/test_oracle_md.py: the unmodified collect_digar_from_MD_tag on (M CIGAR, MD) == the =/X oracle on the converted CIGAR.
 */
#include <ctype.h>
#include <stdint.h>
#include <stdlib.h>
#include "lcd_oracle.h"

enum { CMATCH = 0, CINS = 1, CDEL = 2, CREF_SKIP = 3, CSOFT = 4, CHARD = 5, CEQUAL = 7, CDIFF = 8 };

/* returns the number of ops written (<= cap), -1 when cap is too small, -2 when MD and CIGAR do not match (the reference exits), -3 for an
 * =/X op in the input (the reference exits) */
int64_t lcd_oracle_md_to_eqx(int n_cigar, const uint32_t *cigar, const char *md, uint32_t *out, int64_t cap) {
    int64_t n = 0; int md_i = 0; long last_eq_len = 0;
#define EMIT(op, len) do { if (n >= cap) return -1; out[n++] = ((uint32_t)(len) << 4) | (op); } while (0)
    for (int i = 0; i < n_cigar; ++i) {
        const int op = cigar[i] & 15; const long len = cigar[i] >> 4;
        if (op == CMATCH) {
            long m_len = len, eq_len;
            for (;;) {
                if (last_eq_len > 0) {
                    if (last_eq_len >= m_len) { EMIT(CEQUAL, m_len); last_eq_len -= m_len; m_len = 0; }
                    else { EMIT(CEQUAL, last_eq_len); m_len -= last_eq_len; md_i = 0; last_eq_len = 0; }
                } else if (isdigit((unsigned char)md[md_i])) {
                    char *end; eq_len = strtol(md + md_i, &end, 10); md = end;            /* the reference moves its base pointer and restarts md_i */
                    if (eq_len > m_len) { last_eq_len = eq_len - m_len; eq_len = m_len; }
                    else if (eq_len == 0) { md_i = 0; continue; }
                    EMIT(CEQUAL, eq_len); m_len -= eq_len; md_i = 0;
                } else if (isalpha((unsigned char)md[md_i])) {
                    EMIT(CDIFF, 1); m_len -= 1;
                    if (md[md_i + 1] == '\0' || md[md_i + 1] != '0') md_i++; else md_i += 2;   /* skip the 0 after a mismatch */
                } else return -2;
                if (m_len <= 0) break;
            }
        } else if (op == CDEL) {
            EMIT(CDEL, len);
            md_i++;
            while (md[md_i] && isalpha((unsigned char)md[md_i])) md_i++;
            if (md[md_i] == '0') md_i++;                                                  /* skip the 0 after a deletion */
        } else if (op == CEQUAL || op == CDIFF) return -3;
        else EMIT(op, len);
    }
#undef EMIT
    return n;
}

/* Plain C99 consumer of include/skm_b200.h: proves the boundary is a C ABI (no C++ / torch types) and exercises
 * the entry points that need no device: library version, the alphabet LUT (alphabet.py:88-96 + vectorize.py:193-195)
 * and the FASTA ingest (kmerize.smk:90-129).  Built and run by tests/test_abi.py with `gcc -std=c99`. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "skm_b200.h"

int main(void) {
    if (skm_version() < 10000) return 1;
    /* miqs: 20 residues -> 10 symbols */
    uint8_t lut[256];
    const char *from = "ACDEFGHIKLMNPQRSTVWY", *to = "ACDDFGHIKIIDPIKSSIFF", *syms = "ACDFGHIKPS";
    if (skm_lut_build(from, to, 20, syms, 10, lut) != SKM_OK) { fprintf(stderr, "%s\n", skm_last_error()); return 2; }
    if (lut['E'] != lut['D'] || lut['X'] != SKM_INVALID_SYMBOL || lut['A'] != 0 || lut['a'] != SKM_INVALID_SYMBOL) return 3;
    /* bad arguments are reported, not crashed on */
    if (skm_lut_build(from, to, 20, syms, 300, lut) == SKM_OK) return 4;
    if (strlen(skm_last_error()) == 0) return 5;
    /* FASTA: two records, wrapped lines, CRLF, blanks */
    const char *txt = "junk\n>sp|P1|A first\nACDE\r\nFG H\n>sp|P2|B\n\nKLMN*\n";
    int64_t nseq = 0, nres = 0, idb = 0;
    if (skm_fasta_scan((const uint8_t *)txt, (int64_t)strlen(txt), 2, &nseq, &nres, &idb) != SKM_OK) return 6;
    if (nseq != 2 || nres != 12 || idb != 14) { fprintf(stderr, "%lld %lld %lld\n", (long long)nseq, (long long)nres, (long long)idb); return 7; }
    uint8_t res[16], ids[32];
    int64_t off[3], idoff[3];
    if (skm_fasta_pack((const uint8_t *)txt, (int64_t)strlen(txt), 2, res, off, ids, idoff) != SKM_OK) return 8;
    if (memcmp(res, "ACDEFGHKLMN*", 12) != 0 || off[0] != 0 || off[1] != 7 || off[2] != 12) return 9;
    if (memcmp(ids, "sp|P1|Asp|P2|B", 14) != 0 || idoff[1] != 7 || idoff[2] != 14) return 10;
    /* a compute entry point without a device / with NULL buffers fails with a code and a message */
    if (skm_count_dense(NULL, 10, NULL, 1, NULL, 10, 3, NULL, 1000, 1000, 32, NULL, 10, NULL) == SKM_OK) return 11;
    printf("cabi ok: version %d, %lld records, %lld residues\n", skm_version(), (long long)nseq, (long long)nres);
    return 0;
}

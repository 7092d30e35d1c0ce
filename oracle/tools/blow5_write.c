/* blow5_write.c -- TEST INFRASTRUCTURE. Writes a BLOW5 file from the flat dump format of
 * blow5_dump.c read on stdin, with a chosen experiment_type ("genomic_dna" | "rna") so that the
 * reference's drna_detect (src/misc.c:34-60) selects DNA or RNA detector parameters.
 * Uses the slow5lib write API (slow5.h:525-612). */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <slow5/slow5.h>

int main(int argc, char **argv) {
    if (argc < 3) { fprintf(stderr, "usage: %s out.blow5 genomic_dna|rna < dump.bin\n", argv[0]); return 1; }
    slow5_file_t *sp = slow5_open(argv[1], "w");
    if (!sp) { fprintf(stderr, "cannot open %s for writing\n", argv[1]); return 1; }
    slow5_hdr_t *h = sp->header;
    if (slow5_hdr_add("experiment_type", h) < 0 || slow5_hdr_set("experiment_type", argv[2], 0, h) < 0) return 2;
    if (slow5_hdr_add("sequencing_kit", h) < 0 ||
        slow5_hdr_set("sequencing_kit", strcmp(argv[2], "rna") ? "sqk-lsk109" : "sqk-rna002", 0, h) < 0) return 2;
    if (slow5_hdr_write(sp) < 0) return 3;
    for (;;) {
        uint32_t idl;
        if (fread(&idl, 4, 1, stdin) != 1) break;
        slow5_rec_t *rec = slow5_rec_init();
        rec->read_id = (char *)malloc(idl + 1);
        if (fread(rec->read_id, 1, idl, stdin) != idl) return 4;
        rec->read_id[idl] = 0;
        rec->read_id_len = (uint16_t)idl;
        uint64_t n;
        if (fread(&n, 8, 1, stdin) != 1) return 4;
        if (fread(&rec->digitisation, 8, 1, stdin) != 1) return 4;
        if (fread(&rec->offset, 8, 1, stdin) != 1) return 4;
        if (fread(&rec->range, 8, 1, stdin) != 1) return 4;
        rec->read_group = 0;
        rec->sampling_rate = 4000.0;
        rec->len_raw_signal = n;
        rec->raw_signal = (int16_t *)malloc(n * sizeof(int16_t));
        if (fread(rec->raw_signal, 2, n, stdin) != n) return 4;
        if (slow5_write(rec, sp) < 0) return 5;
        slow5_rec_free(rec);
    }
    slow5_close(sp);
    return 0;
}

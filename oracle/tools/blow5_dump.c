/* blow5_dump.c -- TEST INFRASTRUCTURE. Dumps the fields of every record that the hot path uses
 * (slow5_rec_t: read_id, digitisation, offset, range, len_raw_signal, raw_signal; slow5.h:274-286)
 * to a flat little-endian binary stream on stdout:
 *   per record: u32 id_len, id bytes, u64 n, f64 digitisation, f64 offset, f64 range, i16[n]
 * Links the reference's vendored slow5lib (oracle/_ref/libslow5.a). */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <slow5/slow5.h>

int main(int argc, char **argv) {
    if (argc < 2) { fprintf(stderr, "usage: %s reads.blow5 > dump.bin\n", argv[0]); return 1; }
    slow5_file_t *sp = slow5_open(argv[1], "r");
    if (!sp) { fprintf(stderr, "cannot open %s\n", argv[1]); return 1; }
    slow5_rec_t *rec = NULL;
    int ret;
    while ((ret = slow5_get_next(&rec, sp)) >= 0) {
        uint32_t idl = (uint32_t)strlen(rec->read_id);
        uint64_t n = rec->len_raw_signal;
        fwrite(&idl, 4, 1, stdout);
        fwrite(rec->read_id, 1, idl, stdout);
        fwrite(&n, 8, 1, stdout);
        fwrite(&rec->digitisation, 8, 1, stdout);
        fwrite(&rec->offset, 8, 1, stdout);
        fwrite(&rec->range, 8, 1, stdout);
        fwrite(rec->raw_signal, 2, n, stdout);
    }
    if (ret != SLOW5_ERR_EOF) { fprintf(stderr, "slow5_get_next error %d\n", ret); return 1; }
    slow5_rec_free(rec);
    slow5_close(sp);
    return 0;
}

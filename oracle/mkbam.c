/* mkbam.c — TEST INFRASTRUCTURE ONLY: SAM text -> BAM + .bai, for end-to-end runs of the reference
 * binary (oracle/_ref/longphase-s) on synthetic data.  There is no samtools in the image.
 * usage: mkbam in.sam out.bam */
#include <stdio.h>
#include <htslib/sam.h>

int main(int argc, char **argv) {
    if (argc != 3) { fprintf(stderr, "usage: %s in.sam out.bam\n", argv[0]); return 2; }
    samFile *in = sam_open(argv[1], "r");
    if (!in) { perror(argv[1]); return 1; }
    sam_hdr_t *h = sam_hdr_read(in);
    samFile *out = sam_open(argv[2], "wb");
    if (!h || !out || sam_hdr_write(out, h) < 0) { fprintf(stderr, "cannot write %s\n", argv[2]); return 1; }
    bam1_t *b = bam_init1();
    long n = 0;
    int r;
    while ((r = sam_read1(in, h, b)) >= 0) {
        if (sam_write1(out, h, b) < 0) { fprintf(stderr, "write error\n"); return 1; }
        n++;
    }
    if (r < -1) { fprintf(stderr, "parse error after %ld records\n", n); return 1; }
    bam_destroy1(b);
    sam_hdr_destroy(h);
    sam_close(in);
    if (sam_close(out) < 0) return 1;
    if (sam_index_build(argv[2], 0) < 0) { fprintf(stderr, "index failed\n"); return 1; }
    fprintf(stderr, "%ld records\n", n);
    return 0;
}

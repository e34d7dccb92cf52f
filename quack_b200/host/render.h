/* render.h -- host-side consumer of the accumulator: percent transform + SVG output, byte-identical
 * with the reference's transform() and draw() (quack.c:230-293, 295-856) and the tag printing of
 * svg.c:12-103.  O(max_length x 97) work, stays on the host. */
#ifndef QB_RENDER_H
#define QB_RENDER_H

#include <stdint.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QR_ROW 97 /* base_information as u64[97]: scores[91], content[4], length_count, kmer_count */

typedef struct {
  uint64_t *rows;               /* [max_length][97]; transformed in place */
  uint64_t max_length;          /* becomes the number of bins when binning applies */
  uint64_t original_max_length; /* set by qr_transform */
  uint64_t n_reads;             /* number_of_sequences */
} qr_data;

/* quack.c:230-293: optional 100-bp binning above 3000 bp, cumulative adapter counts, integer score
 * percentages, single-precision ceil() percentages for length and adapter counts.  `log` receives
 * "Binning...\n" (the reference prints it to stderr). */
void qr_transform(qr_data *d, FILE *log);

/* quack.c:295-856: one file's panels.  position 0 = left/only, 1 = right; adapters_used toggles the
 * adapter panel.  Output goes to `out` through the shared indentation state. */
void qr_draw(const qr_data *d, int position, int adapters_used, FILE *out);

/* quack.c:879-909 and 923-925: document frame around the panels. */
void qr_begin_document(int paired, int adapters_used, const char *name, FILE *out);
void qr_end_document(const char *name, FILE *out);

/* Convenience used by the tests: whole SVG for 1 or 2 already-accumulated files. */
void qr_render_all(qr_data *first, qr_data *second, int adapters_used, const char *name, FILE *out, FILE *log);

#ifdef __cplusplus
}
#endif
#endif

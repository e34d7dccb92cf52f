/* render.c -- see render.h.  A restatement, not a copy: tags are emitted by three small printf
 * helpers that take the whole attribute list as one format string, and the panel code is organised
 * around those; what must match the reference is the byte stream, which tests/test_host_cpu.py
 * compares with the reference binary's output for every CLI combination. */
#include "render.h"

#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ tag helpers (svg.c:59-103) */

static int g_level = 0; /* nesting depth: two spaces per level (svg.c:8-9, 65-66) */

static void put_indent(FILE *out) {
  for (int i = 0; i < g_level; i++) fputs("  ", out);
}

/* "<name" + attributes + ">" ; attributes arrive as one printf format, each ` key="value"` */
static void tag_open(FILE *out, const char *name, const char *fmt, ...) {
  va_list ap;
  put_indent(out);
  g_level++;
  fprintf(out, "<%s", name);
  va_start(ap, fmt);
  vfprintf(out, fmt, ap);
  va_end(ap);
  fputs(">\n", out);
}

static void tag_single(FILE *out, const char *name, const char *fmt, ...) {
  va_list ap;
  put_indent(out);
  fprintf(out, "<%s", name);
  va_start(ap, fmt);
  vfprintf(out, fmt, ap);
  va_end(ap);
  fputs("/>\n", out);
}

static void tag_close(FILE *out, const char *name) {
  if (g_level > 0) g_level--;
  put_indent(out);
  fprintf(out, "</%s>\n", name);
}

/* the three text macros of quack.c:15-50 */
static void axis_label(FILE *out, int x, int y, int rot, const char *label) {
  tag_open(out, "text",
           " x=\"%d\" fill=\"#AAA\" y=\"%d\" font-family=\"sans-serif\" font-size=\"15px\" text-anchor=\"middle\""
           " transform=\"rotate(%d)\"",
           x, y, rot);
  fprintf(out, "%s\n", label);
  tag_close(out, "text");
}

static void axis_number(FILE *out, int x, int y, const char *anchor, int number) {
  tag_open(out, "text", " x=\"%d\" fill=\"#AAA\" y=\"%d\" font-family=\"sans-serif\" font-size=\"10px\" text-anchor=\"%s\"",
           x, y, anchor);
  fprintf(out, "%d\n", number);
  tag_close(out, "text");
}

static void center_label_open(FILE *out, int x, int y, const char *fill) {
  tag_open(out, "text",
           " x=\"%d\" y=\"%d\" fill=\"%s\" font-family=\"sans-serif\" font-size=\"15px\" font-weight=\"bold\""
           " text-anchor=\"middle\"",
           x, y, fill);
}

static void panel_caption(FILE *out, int y, const char *fill, const char *caption) {
  tag_open(out, "text", " y=\"%d\" fill=\"%s\" x=\"%d\" font-family=\"sans-serif\" font-size=\"15px\"", y, fill, 5);
  fprintf(out, "%s\n", caption);
  tag_close(out, "text");
}

static void percent_axis(FILE *out, int position, int label_at, int zero_y, int hundred_y) {
  /* quack.c:515-523, 680-688, 738-746: the 0 / 100 "Percent" axis beside a panel */
  if (position == 0) {
    axis_label(out, -label_at, -5, -90, "Percent");
    axis_number(out, -5, zero_y, "end", 0);
    axis_number(out, -5, hundred_y, "end", 100);
  } else {
    axis_label(out, label_at, -455, 90, "Percent");
    axis_number(out, 455, zero_y, "start", 0);
    axis_number(out, 455, hundred_y, "start", 100);
  }
}

/* ------------------------------------------------------------------ transform (quack.c:230-293) */

void qr_transform(qr_data *d, FILE *log) {
  uint64_t *rows = d->rows;
  d->original_max_length = d->max_length;
  if (d->max_length > 3000) { /* quack.c:234-262 */
    if (log) fprintf(log, "Binning...\n");
    uint64_t bin = 0;
    for (uint64_t pos = 1; pos < d->max_length; pos++) {
      if (pos % 100 == 0) { /* entering a new bin: its row is reused, all but kmer_count cleared */
        bin++;
        memset(rows + bin * QR_ROW, 0, 96 * sizeof(uint64_t));
      }
      uint64_t *dst = rows + bin * QR_ROW;
      const uint64_t *src = rows + pos * QR_ROW;
      for (int j = 0; j < QR_ROW; j++) dst[j] = dst[j] + src[j];
    }
    d->max_length = bin;
  }
  for (uint64_t pos = 1; pos < d->max_length; pos++) /* quack.c:264-266 */
    rows[pos * QR_ROW + 96] += rows[(pos - 1) * QR_ROW + 96];
  for (uint64_t pos = 0; pos < d->max_length; pos++) { /* quack.c:269-291 */
    uint64_t *row = rows + pos * QR_ROW;
    int score_sum = 0; /* the reference sums into an int */
    for (int j = 0; j < 91; j++) score_sum = (int)((uint64_t)score_sum + row[j]);
    if (score_sum != 0)
      for (int j = 0; j < 91; j++) row[j] = 100 * row[j] / (uint64_t)(int64_t)score_sum;
    /* product and quotient in single precision, then ceil in double (quack.c:288-289) */
    row[95] = (uint64_t)ceil(100 * (float)row[95] / d->n_reads);
    row[96] = (uint64_t)ceil(100 * (float)row[96] / (float)d->n_reads);
  }
}

/* ------------------------------------------------------------------ draw (quack.c:295-856) */

void qr_draw(const qr_data *d, int position, int adapters_used, FILE *out) {
  const uint64_t *rows = d->rows;
  const int L = (int)d->max_length; /* printed with %d in the reference */
  const int n_reads = (int)d->n_reads;
  int x, y, i, j;

  /* encoding: phred33 iff some position has a percentage below raw score index 31 (quack.c:303-322) */
  const char *encoding = "phred64";
  int offset = 31;
  for (i = 0; i < L && offset; i++)
    for (j = 0; j < 31; j++)
      if (rows[(size_t)i * QR_ROW + j] > 0) {
        encoding = "phred33";
        offset = 0;
        break;
      }

  /* max score, score distribution, per-position mean (quack.c:325-341) */
  int max_score = 40;
  uint64_t total_counts[91] = {0};
  uint64_t number_of_bases = 0;
  float *averages = (float *)malloc(sizeof(float) * (size_t)(L > 0 ? L : 1));
  for (i = 0; i < L; i++) {
    int sum = 0;
    for (j = offset; j < 91; j++) {
      const uint64_t pct = rows[(size_t)i * QR_ROW + j];
      if (pct > 0 && (j - offset) > max_score) max_score = j - offset;
      total_counts[j - offset] += pct;
      sum += (j - offset) * (int)pct;
    }
    number_of_bases++;
    averages[i] = sum / 100.0;
  }

  /* file stats line (quack.c:344-362) */
  tag_open(out, "text", " x=\"%d\" y=\"%d\" text-anchor=\"middle\" font-family=\"sans-serif\" font-size=\"15px\" fill=\"#555\"",
           position == 0 ? 355 : 835, 20);
  tag_open(out, "tspan", "%s", "");
  fprintf(out, "%d", n_reads);
  tag_close(out, "tspan");
  tag_open(out, "tspan", " fill=\"#888\"");
  fprintf(out, "&#160;reads with endcoding&#160;");
  tag_close(out, "tspan");
  tag_open(out, "tspan", "%s", "");
  fprintf(out, "%s", encoding);
  tag_close(out, "tspan");
  tag_close(out, "text");

  tag_open(out, "g", " transform=\"translate(%d %d)\"", 5, 25); /* rug plot */

  /* horizontal tick marks (quack.c:369-384) */
  x = position == 1 ? 1000 : 100;
  for (i = 10; i < 100; i += 10) {
    y = 105 + i * 250 / 100;
    tag_single(out, "line", " x1=\"%d\" x2=\"%d\" y1=\"%d\" y2=\"%d\" stroke=\"black\" stroke-width=\"%f\"", x, x + 100, y, y,
               (i % 20 == 10) ? 1.0 : 0.5);
  }

  tag_open(out, "g", " transform=\"translate(%d 0)\"", position == 1 ? 610 : 130); /* vertical section */

  /* vertical tick marks (quack.c:393-405) */
  y = adapters_used == 0 ? 400 : 500;
  for (i = 10; i < 100; i += 10) {
    x = i * 450 / 100;
    tag_single(out, "line", " x1=\"%d\" x2=\"%d\" y1=\"%d\" y2=\"%d\" stroke=\"black\" stroke-width=\"%f\"", x, x, 10, y,
               (i % 20 == 10) ? 1.0 : 0.5);
  }

  /* ---- base content (quack.c:408-523): four stacked polylines of cumulative raw counts ---- */
  tag_open(out, "g", " transform=\"translate(%d,%d) scale(%d, %d)\"", 0, 100, 1, -1);
  tag_open(out, "svg", " width=\"%d\" height=\"%d\" preserveAspectRatio=\"none\" viewBox=\"0 0 %d %d\"", 450, 100, L, n_reads);
  tag_single(out, "rect", " width=\"100%%\" height=\"100%%\" fill=\"#CCC\"");
  {
    static const char *colors[4] = {"#648964", "#89bc89", "#84accf", "#5d7992"};
    /* the reference builds four strings of at most 25 bytes per position (snprintf + strncat) */
    const size_t cap = 25 * (size_t)(L > 0 ? L : 1) + 64;
    char *pts[4];
    size_t used[4];
    for (i = 0; i < 4; i++) {
      pts[i] = (char *)malloc(cap);
      used[i] = 0;
    }
    y = 0;
    for (i = 0; i < 4; i++) { /* first point sits on the left edge */
      y += (int)rows[91 + i];
      used[i] += (size_t)snprintf(pts[i] + used[i], cap - used[i], "0,%d ", y);
    }
    for (x = 0; x < L; x++) {
      y = 0;
      for (i = 0; i < 4; i++) {
        y += (int)rows[(size_t)x * QR_ROW + 91 + i];
        used[i] += (size_t)snprintf(pts[i] + used[i], cap - used[i], "%d.5,%d ", x, y);
      }
    }
    y = 0;
    for (i = 0; i < 4; i++) { /* last point on the right edge */
      y += (int)rows[(size_t)(L - 1) * QR_ROW + 91 + i];
      used[i] += (size_t)snprintf(pts[i] + used[i], cap - used[i], "%d,%d ", L, y);
    }
    for (i = 3; i >= 0; i--) /* G, C, T, A so that they stack */
      tag_single(out, "polyline", " points=\"0,0 %s %d,0\" fill=\"%s\" stroke=\"none\"", pts[i], L, colors[i]);
    for (i = 0; i < 4; i++) free(pts[i]);

    tag_close(out, "svg");
    tag_close(out, "g");
    panel_caption(out, 95, "#CCC", "Base Content Percentage");
    if (position == 0) {
      static const char *labels[4] = {"%A", "%T", "%C", "%G"};
      for (i = 0; i < 4; i++) {
        center_label_open(out, 465, 20 * (4 - i), colors[i]);
        fprintf(out, "%s", labels[i]);
        tag_close(out, "text");
      }
    }
  }
  percent_axis(out, position, 50, 100, 5);

  /* ---- heatmap (quack.c:525-628) ---- */
  tag_open(out, "g", " transform=\"translate(%d,%d) scale(%d, %d)\"", 0, 355, 1, -1);
  tag_open(out, "svg", " width=\"%d\" height=\"%d\" preserveAspectRatio=\"none\" viewBox=\"0 0 %d %d\"", 450, 250, L, max_score);
  {
    static const char *bands[3] = {"#ccebc5", "#ffffcc", "#fbb4ae"}; /* green, yellow, red */
    const int heights[3] = {max_score, 28, 20};
    for (i = 0; i < 3; i++)
      tag_single(out, "rect", " x=\"%d\" y=\"%d\" width=\"100%%\" height=\"%d\" stroke=\"none\" fill=\"%s\"", 0, 0, heights[i],
                 bands[i]);
  }
  {
    const size_t cap = 15 * (size_t)(L > 0 ? L : 1) + 64;
    char *line = (char *)malloc(cap);
    size_t used = (size_t)snprintf(line, cap, "0,%0.2f ", averages[0]);
    for (x = 0; x < L; x++) {
      for (y = 0; y < max_score; y++) {
        const uint64_t pct = rows[(size_t)x * QR_ROW + y + offset];
        if (pct > 0)
          tag_single(out, "rect",
                     " x=\"%d\" y=\"%d\" fill-opacity=\"%f\" width=\"%d\" height=\"%d\" stroke=\"none\" stroke-width=\"%d\""
                     " fill=\"black\"",
                     x, y, (float)pct / 100.0, 1, 1, 0);
      }
      used += (size_t)snprintf(line + used, cap - used, "%d.5,%0.2f ", x, averages[x]);
    }
    snprintf(line + used, cap - used, "%d,%0.2f", L, averages[L - 1]);
    tag_single(out, "polyline",
               " points=\"%s\" stroke=\"black\" stroke-width=\"%f\" stroke-opacity=\"%f\" fill=\"none\" stroke-linejoin=\"round\"",
               line, 0.5, 0.5);
    free(line);
  }
  tag_close(out, "svg");
  tag_close(out, "g");
  panel_caption(out, 350, "#888", "Per Base Sequence Quality");
  if (position == 0) {
    const int ys[3] = {112, 112 + (int)((max_score - 28) * 250 / max_score), 112 + (int)((max_score - 20) * 250 / max_score)};
    const int vals[3] = {max_score, 28, 20};
    for (i = 0; i < 3; i++) {
      center_label_open(out, 465, ys[i], "#888");
      fprintf(out, "%d", vals[i]);
      tag_close(out, "text");
    }
  }

  /* ---- length distribution (quack.c:631-688) and adapter distribution (quack.c:691-747) ---- */
  for (int panel = 0; panel < (adapters_used == 1 ? 2 : 1); panel++) {
    const int col = panel == 0 ? 95 : 96;
    const int top = panel == 0 ? 360 : 465;
    tag_open(out, "svg", " x=\"%d\" y=\"%d\" width=\"%d\" height=\"%d\" preserveAspectRatio=\"none\" viewBox=\"0 0 %d 100\"", 0, top,
             450, 100, L);
    tag_single(out, "rect", " width=\"100%%\" height=\"100%%\" fill=\"#EEE\"");
    for (x = 0; x < L; x++) {
      const uint64_t v = rows[(size_t)x * QR_ROW + col];
      if (v > 0)
        tag_single(out, "rect", " x=\"%d\" y=\"%d\" width=\"%d\" height=\"%d\" stroke=\"none\" fill=\"steelblue\"", x, 0, 1, (int)v);
    }
    tag_close(out, "svg");
    panel_caption(out, top + 95, "#888", panel == 0 ? "Length Distribution" : "Adapter Distribution");
    percent_axis(out, position, top + 50, top + 10, top + 100);
  }

  /* bottom axis (quack.c:749-755) */
  y = 470 + (adapters_used == 1 ? 105 : 0);
  axis_label(out, 225, y + 5, 0, "Base Pairs");
  axis_number(out, 0, y, "middle", 0);
  axis_number(out, 450, y, "middle", L);
  tag_close(out, "g"); /* vertical section */

  /* ---- score distribution beside the heatmap (quack.c:760-852) ---- */
  tag_open(out, "g", " transform=\"translate(%d,%d) scale(%d, %d)\"", position == 0 ? 125 : 1065, 355, position == 0 ? -1 : 1, -1);
  tag_open(out, "svg", " width=\"%d\" height=\"%d\" preserveAspectRatio=\"none\" viewBox=\"0 0 100 %d\"", 100, 250, max_score);
  tag_single(out, "rect", " width=\"100%%\" height=\"100%%\" fill=\"#EEE\"");
  for (y = 0; y < max_score; y++)
    if (total_counts[y] > 0)
      tag_single(out, "rect", " x=\"%d\" y=\"%d\" width=\"%d\" height=\"%d\" stroke=\"none\" fill=\"steelblue\"", 0, y,
                 (int)(total_counts[y] / number_of_bases), 1);
  tag_close(out, "svg");
  tag_close(out, "g");
  {
    const int left = position == 0;
    const int tx = left ? 30 : 1070;
    tag_open(out, "text", " y=\"%d\" fill=\"#888\" x=\"%d\" font-family=\"sans-serif\" font-size=\"15px\"", 335, tx);
    tag_open(out, "tspan", "%s", "");
    fprintf(out, "Score\n");
    tag_close(out, "tspan");
    tag_open(out, "tspan", " dy=\"%d\" x=\"%d\"", 15, tx);
    fprintf(out, "Distribution\n");
    tag_close(out, "tspan");
    tag_close(out, "text");
    if (left) {
      axis_label(out, 72, 100, 0, "Percent");
      axis_number(out, 25, 100, "middle", 100);
      axis_label(out, -230, 20, -90, "Score");
      axis_number(out, 20, 110, "end", max_score);
      axis_number(out, 20, 355, "end", 1);
    } else {
      axis_label(out, 1115, 100, 0, "Percent");
      axis_number(out, 1165, 100, "middle", 100);
      axis_label(out, 230, -1170, 90, "Score");
      axis_number(out, 1170, 110, "start", max_score);
      axis_number(out, 1170, 355, "start", 1);
    }
  }
  tag_close(out, "g"); /* rug plot */
  free(averages);
}

/* ------------------------------------------------------------------ document frame */

void qr_begin_document(int paired, int adapters_used, const char *name, FILE *out) {
  const int width = paired ? 1195 : 615;
  int height = adapters_used ? 610 : 510;
  if (name) height += 30;
  g_level = 0;
  tag_open(out, "svg",
           " width=\"%d\" height=\"%d\" viewBox=\"%d %d %d %d\" xmlns=\"http://www.w3.org/2000/svg\""
           " xmlns:xlink=\"http://www.w3.org/1999/xlink\"",
           width, height, 0, 0, width, height);
  if (name) { /* quack.c:894-909 */
    tag_open(out, "text", " x=\"%d\" y=\"%d\" font-family=\"sans-serif\" text-anchor=\"middle\" font-size=\"30px\" fill=\"black\"",
             width / 2, 30);
    fprintf(out, "%s", name);
    tag_close(out, "text");
    tag_open(out, "g", " transform=\"translate(%d %d)\"", 0, 30);
  }
}

void qr_end_document(const char *name, FILE *out) {
  if (name) tag_close(out, "g");
  tag_close(out, "svg");
}

void qr_render_all(qr_data *first, qr_data *second, int adapters_used, const char *name, FILE *out, FILE *log) {
  qr_begin_document(second != NULL, adapters_used, name, out);
  qr_transform(first, log);
  qr_draw(first, 0, adapters_used, out);
  if (second) {
    qr_transform(second, log);
    qr_draw(second, 1, adapters_used, out);
  }
  qr_end_document(name, out);
}

/* Test/binding entry point: render 1 or 2 raw accumulators (as returned by qb_finish) to a file.
 * rows are copied, so the caller's arrays stay raw.  Returns 0, -1 if the file cannot be opened. */
int qr_render_to_path(const uint64_t *rows1, uint64_t max_length1, uint64_t n_reads1, const uint64_t *rows2,
                      uint64_t max_length2, uint64_t n_reads2, int adapters_used, const char *name,
                      const char *path) {
  FILE *out = fopen(path, "wb");
  if (!out) return -1;
  qr_data a = {0}, b = {0};
  a.rows = (uint64_t *)malloc(sizeof(uint64_t) * QR_ROW * (size_t)(max_length1 ? max_length1 : 1));
  memcpy(a.rows, rows1, sizeof(uint64_t) * QR_ROW * (size_t)max_length1);
  a.max_length = max_length1;
  a.n_reads = n_reads1;
  if (rows2) {
    b.rows = (uint64_t *)malloc(sizeof(uint64_t) * QR_ROW * (size_t)(max_length2 ? max_length2 : 1));
    memcpy(b.rows, rows2, sizeof(uint64_t) * QR_ROW * (size_t)max_length2);
    b.max_length = max_length2;
    b.n_reads = n_reads2;
  }
  qr_render_all(&a, rows2 ? &b : NULL, adapters_used, name, out, NULL);
  free(a.rows);
  free(b.rows);
  fclose(out);
  return 0;
}

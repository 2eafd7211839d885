/* TEST INFRASTRUCTURE -- CPU oracle, residue classification and cascade lattice VQ (see
 * orc.h).  Restated from the algorithm of /root/reference/lib/res0.c:316-640,715-799. */
#include <stdlib.h>
#include <string.h>
#include "orc.h"

static int first_class(const orc_residue *r, int a, int b, int wildcard) {
  int k;
  for (k = 0; k < r->partitions - 1; k++)
    if (a <= r->classmetric1[k] && ((wildcard && r->classmetric2[k] < 0) || (wildcard ? b < r->classmetric2[k] : b <= r->classmetric2[k])))
      break;
  return k;
}

/* res1_class/_01class (res0.c:406-468,735) and res2_class/_2class (res0.c:473-526,759).
   partword is [rows][partvals]; returns the number of rows (0 == the reference returns NULL). */
int orc_residue_class(const orc_residue *r, int **in, const int *nonzero, int ch, int *partword) {
  int partvals = (r->end - r->begin) / r->grouping, used = 0, i, j, k;
  for (i = 0; i < ch; i++) if (nonzero[i]) used++;
  if (!used) return 0;
  if (r->type == 2) {
    int l = r->begin / ch;
    for (i = 0; i < partvals; i++) {
      int mag = 0, ang = 0;
      for (j = 0; j < r->grouping; j += ch, l++) {
        if (abs(in[0][l]) > mag) mag = abs(in[0][l]);
        for (k = 1; k < ch; k++) if (abs(in[k][l]) > ang) ang = abs(in[k][l]);
      }
      partword[i] = first_class(r, mag, ang, 0);
    }
    return 1;
  } else {
    float scale = 100. / r->grouping;
    int row = 0;
    for (j = 0; j < ch; j++) {
      if (!nonzero[j]) continue;
      for (i = 0; i < partvals; i++) {
        const int *v = in[j] + r->begin + i * r->grouping;
        int mx = 0, ent = 0;
        for (k = 0; k < r->grouping; k++) { int a = abs(v[k]); if (a > mx) mx = a; ent += a; }
        ent *= scale;
        partword[row * partvals + i] = first_class(r, mx, ent, 1);
      }
      row++;
    }
    return used;
  }
}

/* local_book_besterror, res0.c:316-378: nearest lattice point by per-component rounding; when that
   entry has no codeword, exhaustive search over the used entries (first minimum wins); the chosen
   point is subtracted in place. */
static int quantize(const orc_book *b, int *a) {
  int dim = b->dim, qv = b->quantvals, ze = qv >> 1, index = 0, i, o;
  int pt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (i = 0, o = dim; i < dim; i++) {
    int v, m;
    --o;
    v = (b->delta != 1) ? (a[o] - b->minval + (b->delta >> 1)) / b->delta : a[o] - b->minval;
    m = (v < ze) ? ((ze - v) << 1) - 1 : ((v - ze) << 1);
    index = index * qv + (m < 0 ? 0 : (m >= qv ? qv - 1 : m));
    pt[o] = v * b->delta + b->minval;
  }
  if (b->len[index] == 0) {
    int e[8] = {0, 0, 0, 0, 0, 0, 0, 0}, best = -1, j;
    int top = b->minval + b->delta * (b->quantvals - 1);
    for (i = 0; i < b->entries; i++) {
      if (b->len[i] > 0) {
        int d = 0;
        for (j = 0; j < dim; j++) d += (e[j] - a[j]) * (e[j] - a[j]);
        if (best == -1 || d < best) { memcpy(pt, e, sizeof(pt)); best = d; index = i; }
      }
      for (j = 0; j < 8 && e[j] >= top; j++) e[j] = 0;
      if (j < 8) { if (e[j] >= 0) e[j] += b->delta; e[j] = -e[j]; }
    }
  }
  if (index > -1) for (i = 0; i < dim; i++) a[i] -= pt[i];
  return index;
}

/* res1_forward / res2_forward -> _01forward -> _encodepart (res0.c:384-404,528-640,715-799).
   Emits (book,entry) pairs in the reference's emission order; returns their number. */
int orc_residue_forward(const orc_residue *r, int **in, const int *nonzero, int ch, const int *partword,
                        int n2, int *syms, int max_syms) {
  int partvals = (r->end - r->begin) / r->grouping, used = 0, rows, ns = 0, s, i, j, k;
  int *vec[8], *inter = NULL;
  for (i = 0; i < ch; i++) if (nonzero[i]) used++;
  if (!used) return 0;
  if (r->type == 2) {
    inter = malloc(sizeof(int) * ch * n2);
    for (i = 0; i < ch; i++) for (j = 0; j < n2; j++) inter[j * ch + i] = in[i][j];
    vec[0] = inter; rows = 1;
  } else {
    rows = 0;
    for (i = 0; i < ch; i++) if (nonzero[i]) vec[rows++] = in[i];
  }
  for (s = 0; s < r->stages; s++) {
    for (i = 0; i < partvals;) {
      if (s == 0)
        for (j = 0; j < rows; j++) {
          int val = partword[j * partvals + i];
          for (k = 1; k < r->phrase_dim; k++) { val *= r->partitions; if (i + k < partvals) val += partword[j * partvals + i + k]; }
          if (val < r->books[r->groupbook].entries && ns < max_syms) { syms[ns * 2] = r->groupbook; syms[ns * 2 + 1] = val; ns++; }
        }
      for (k = 0; k < r->phrase_dim && i < partvals; k++, i++) {
        int off = i * r->grouping + r->begin;
        for (j = 0; j < rows; j++) {
          int cls = partword[j * partvals + i];
          if (r->secondstages[cls] & (1 << s)) {
            int bk = r->partbooks[cls * 8 + s], t;
            if (bk >= 0) {
              const orc_book *b = r->books + bk;
              for (t = 0; t < r->grouping / b->dim; t++) {
                int e = quantize(b, vec[j] + off + t * b->dim);
                if (ns < max_syms) { syms[ns * 2] = bk; syms[ns * 2 + 1] = e; ns++; }
              }
            }
          }
        }
      }
    }
  }
  free(inter);
  return ns;
}

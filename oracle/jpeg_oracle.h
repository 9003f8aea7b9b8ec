/*
 * oracle/jpeg_oracle.h -- TEST INFRASTRUCTURE.  CPU restatement of the
 * reference's JPEG-encode hot path (see jpeg_oracle.c).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg may use it; the
 * product (mozjpeg_b200/, libb200jpeg.so) never links or calls it.
 *
 * It consumes the same plain parameter block as the product's C-ABI
 * (include/b200jpeg.h, data declaration only).
 */
#ifndef JPEG_ORACLE_H
#define JPEG_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#include "../include/b200jpeg.h"

#ifdef __cplusplus
extern "C" {
int orc_encode_coefs(const b200jpeg_params *p, const int16_t *const *planes, const size_t *pitch_blocks, uint8_t **out, size_t *outsize);
int orc_encode_raw(const b200jpeg_params *p, const uint8_t *const *planes, const size_t *pitch, uint8_t **out, size_t *outsize);
#endif

typedef struct {
  int ncomp;
  int wib[4], hib[4];      /* real blocks per component                       */
  int wpad[4], hpad[4];    /* padded to whole interleaved MCUs (dummy blocks)  */
  int16_t *plain[4];       /* plain-quantized coefs  [hpad][wpad][64] natural  */
  int16_t *raw[4];         /* raw DCT output (x8)    [hpad][wpad][64] natural  */
  int16_t *final_[4];      /* coefs entering the scans (after trellis)          */
  /* Huffman tables used by the trellis rate model, per component */
  b200jpeg_huff_tbl trellis_dc[4], trellis_ac[4];
  /* tables written for each scan (index = scan number; [0]=dc,[1]=ac per tbl slot) */
  int nscans;
  b200jpeg_huff_tbl scan_dc[B200JPEG_MAX_SCANS][4], scan_ac[B200JPEG_MAX_SCANS][4];
  size_t scan_bytes[B200JPEG_MAX_SCANS];   /* entropy-coded bytes per scan (stuffed) */
} orc_debug;

void orc_debug_free(orc_debug *d);

/* Whole encode.  pixels: 8-bit interleaved, input_components per pixel.
 * *out is malloc'ed.  dbg may be NULL.  Returns 0 or a B200JPEG_ERR_* code. */
int orc_encode(const b200jpeg_params *p, const uint8_t *pixels, size_t row_pitch,
               uint8_t **out, size_t *outsize, orc_debug *dbg);
void orc_free(void *p);

/* Stage-level entry points (each restates one reference function). */
void orc_rgb_to_ycc(int r, int g, int b, int *y, int *cb, int *cr);       /* jccolor.c:213-246, jccolext.c:30-75 */
void orc_fdct_islow(int *data);                                           /* jfdctint.c:142-286 */
void orc_deringing(int *data, int q0);                                    /* jcdctmgr.c:416-498 */
int  orc_quantize_coef(int x, int q);                                     /* jcdctmgr.c:611-682 */
void orc_gen_optimal_table(long *freq257, b200jpeg_huff_tbl *out);        /* jchuff.c:947-1106 */
int  orc_make_derived(const b200jpeg_huff_tbl *t, int is_dc, unsigned *ehufco, unsigned char *ehufsi); /* jchuff.c:231-318 */
/* quantize_trellis on one block row (jcdctmgr.c:936-1330), natural-order blocks */
void orc_trellis_row(const b200jpeg_params *p, const unsigned char *dcsi, const unsigned char *acsi,
                     int16_t *coef_blocks, const int16_t *src, int num_blocks,
                     const uint16_t *qtbl, int16_t *last_dc_val, const int16_t *coef_above, const int16_t *src_above, int Ss, int Se, double *norm_src, double *norm_coef);

#ifdef __cplusplus
}
#endif
#endif

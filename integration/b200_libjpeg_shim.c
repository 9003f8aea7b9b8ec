/*
 * b200_libjpeg_shim.c -- the reference-side binding of libb200jpeg, as an
 * interposition library for the libjpeg API (JPEG_LIB_VERSION 62).
 *
 * Loaded in front of the reference's libjpeg (LD_PRELOAD, or linked before it),
 * it takes over the three calls that bracket the encode hot path,
 *
 *     jpeg_start_compress   (jcapistd.c:44-70)
 *     jpeg_write_scanlines  (jcapistd.c:90-135)
 *     jpeg_finish_compress  (jcapimin.c:176-229)
 *
 * and runs the image through the C-ABI of include/b200jpeg.h (sm_100a kernels).
 * Everything else -- jpeg_create_compress, jpeg_set_defaults, jpeg_set_quality,
 * jpeg_c_set_*_param, destination managers, error handling -- stays the
 * reference's own code, so an unmodified application (the reference's `cjpeg`
 * binary, in tests/test_libjpeg_shim.py) produces its files on the GPU.
 *
 * Parameter sets the device path does not cover (B200JPEG_ERR_UNSUPPORTED from
 * b200jpeg_start_compress: JDCT_IFAST, smoothing, arithmetic coding, raw data,
 * 12-bit through this 8-bit entry point, ...) and hosts without a CUDA device fall through to the reference's
 * implementation of the same three functions (dlsym RTLD_NEXT): that is the
 * REFERENCE running, not a CPU path of this library.  B200_SHIM_VERBOSE=1
 * reports on stderr which path an image took; B200_SHIM_REQUIRE=1 turns a
 * fall-through into error_exit (used by the tests, which must not pass on the
 * reference's code).
 *
 * Built against the reference's own headers (jpeglib.h / jpegint.h and the
 * generated jconfig.h under oracle/_ref/cfg) by integration/Makefile; nothing
 * of the reference is copied into this repository.
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define JPEG_INTERNALS
#include "jinclude.h"
#include "jpeglib.h"        /* with JPEG_INTERNALS this pulls in jpegint.h and jerror.h */

#include "b200jpeg.h"

#define MAX_ACTIVE 16
static struct { j_compress_ptr cinfo; b200jpeg_encoder *enc; } g_active[MAX_ACTIVE];
static b200jpeg_encoder *g_idle_enc;          /* encoders are reused: creating one costs a CUDA context */
static int g_no_device;

static int verbose(void) { const char *v = getenv("B200_SHIM_VERBOSE"); return v && v[0] == '1'; }
static int required(void) { const char *v = getenv("B200_SHIM_REQUIRE"); return v && v[0] == '1'; }

typedef void (*start_fn)(j_compress_ptr, boolean);
typedef JDIMENSION (*write_fn)(j_compress_ptr, JSAMPARRAY, JDIMENSION);
typedef void (*finish_fn)(j_compress_ptr);
typedef void (*abort_fn)(j_compress_ptr);
typedef void (*marker_fn)(j_compress_ptr, int, const JOCTET *, unsigned int);

static void *next_sym(const char *name)
{
  void *p = dlsym(RTLD_NEXT, name);
  if (!p) { fprintf(stderr, "b200 shim: the reference's %s is not behind this library\n", name); abort(); }
  return p;
}

static int find_active(j_compress_ptr cinfo)
{
  for (int i = 0; i < MAX_ACTIVE; i++) if (g_active[i].cinfo == cinfo) return i;
  return -1;
}
static void release_slot(int i)
{
  if (g_idle_enc) b200jpeg_encoder_destroy(g_active[i].enc); else g_idle_enc = g_active[i].enc;
  g_active[i].cinfo = NULL; g_active[i].enc = NULL;
}

/* the encoder-relevant state of the reference's object -> b200jpeg_params.  Returns 0 if the
 * object uses something the parameter block cannot express (then the reference encodes it). */
static int fill_params(j_compress_ptr cinfo, boolean write_all_tables, b200jpeg_params *p)
{
  int i, ci;
  memset(p, 0, sizeof(*p));
  if (!write_all_tables) return 0;                       /* abbreviated datastreams: reference only */
  if (cinfo->data_precision != 8 || cinfo->arith_code || cinfo->raw_data_in || cinfo->master->lossless) return 0;
  if (cinfo->num_components > B200JPEG_MAX_COMPONENTS || cinfo->num_scans > B200JPEG_MAX_SCANS) return 0;
  switch (cinfo->in_color_space) {
  case JCS_GRAYSCALE: p->in_color_space = B200JPEG_CS_GRAYSCALE; break;
  case JCS_RGB: case JCS_EXT_RGB: p->in_color_space = B200JPEG_CS_RGB; break;
  case JCS_YCbCr: p->in_color_space = B200JPEG_CS_YCbCr; break;
  default: return 0;
  }
  switch (cinfo->jpeg_color_space) {
  case JCS_GRAYSCALE: p->jpeg_color_space = B200JPEG_CS_GRAYSCALE; break;
  case JCS_YCbCr: p->jpeg_color_space = B200JPEG_CS_YCbCr; break;
  case JCS_RGB: p->jpeg_color_space = B200JPEG_CS_RGB; break;
  default: return 0;
  }
  p->image_width = (int)cinfo->image_width;   p->image_height = (int)cinfo->image_height;
  p->input_components = cinfo->input_components;
  p->data_precision = cinfo->data_precision;
  p->num_components = cinfo->num_components;
  for (ci = 0; ci < cinfo->num_components; ci++) {
    jpeg_component_info *c = &cinfo->comp_info[ci];
    p->comp_info[ci].component_id = c->component_id;
    p->comp_info[ci].h_samp_factor = c->h_samp_factor;  p->comp_info[ci].v_samp_factor = c->v_samp_factor;
    p->comp_info[ci].quant_tbl_no = c->quant_tbl_no;
    p->comp_info[ci].dc_tbl_no = c->dc_tbl_no;          p->comp_info[ci].ac_tbl_no = c->ac_tbl_no;
  }
  for (i = 0; i < NUM_QUANT_TBLS; i++) if (cinfo->quant_tbl_ptrs[i]) {
    for (int k = 0; k < DCTSIZE2; k++) p->quant_tbl[i][k] = cinfo->quant_tbl_ptrs[i]->quantval[k];
    p->quant_tbl_present[i] = 1;
  }
  for (i = 0; i < NUM_HUFF_TBLS; i++) {
    if (cinfo->dc_huff_tbl_ptrs[i]) { memcpy(p->dc_huff_tbl[i].bits, cinfo->dc_huff_tbl_ptrs[i]->bits, 17);
      memcpy(p->dc_huff_tbl[i].huffval, cinfo->dc_huff_tbl_ptrs[i]->huffval, 256); p->dc_huff_tbl[i].present = 1; }
    if (cinfo->ac_huff_tbl_ptrs[i]) { memcpy(p->ac_huff_tbl[i].bits, cinfo->ac_huff_tbl_ptrs[i]->bits, 17);
      memcpy(p->ac_huff_tbl[i].huffval, cinfo->ac_huff_tbl_ptrs[i]->huffval, 256); p->ac_huff_tbl[i].present = 1; }
  }
  p->num_scans = cinfo->scan_info ? cinfo->num_scans : 0;
  for (i = 0; i < p->num_scans; i++) {
    p->scan_info[i].comps_in_scan = cinfo->scan_info[i].comps_in_scan;
    for (int k = 0; k < MAX_COMPS_IN_SCAN; k++) p->scan_info[i].component_index[k] = cinfo->scan_info[i].component_index[k];
    p->scan_info[i].Ss = cinfo->scan_info[i].Ss; p->scan_info[i].Se = cinfo->scan_info[i].Se;
    p->scan_info[i].Ah = cinfo->scan_info[i].Ah; p->scan_info[i].Al = cinfo->scan_info[i].Al;
  }
  p->optimize_coding = cinfo->optimize_coding;   p->dct_method = cinfo->dct_method;
  p->restart_interval = (int)cinfo->restart_interval; p->restart_in_rows = cinfo->restart_in_rows;
  p->smoothing_factor = cinfo->smoothing_factor;
  p->write_JFIF_header = cinfo->write_JFIF_header; p->write_Adobe_marker = cinfo->write_Adobe_marker;
  p->JFIF_major_version = cinfo->JFIF_major_version; p->JFIF_minor_version = cinfo->JFIF_minor_version;
  p->density_unit = cinfo->density_unit; p->X_density = cinfo->X_density; p->Y_density = cinfo->Y_density;
  /* mozjpeg extension block, jpegint.h:93-135 */
  p->compress_profile = cinfo->master->compress_profile;  p->optimize_scans = cinfo->master->optimize_scans;
  p->trellis_quant = cinfo->master->trellis_quant;        p->trellis_quant_dc = cinfo->master->trellis_quant_dc;
  p->trellis_eob_opt = cinfo->master->trellis_eob_opt;    p->use_scans_in_trellis = cinfo->master->use_scans_in_trellis;
  p->trellis_q_opt = cinfo->master->trellis_q_opt;        p->overshoot_deringing = cinfo->master->overshoot_deringing;
  p->trellis_freq_split = cinfo->master->trellis_freq_split; p->trellis_num_loops = cinfo->master->trellis_num_loops;
  p->lambda_log_scale1 = cinfo->master->lambda_log_scale1; p->lambda_log_scale2 = cinfo->master->lambda_log_scale2;
  p->trellis_delta_dc_weight = cinfo->master->trellis_delta_dc_weight;
  p->use_lambda_weight_tbl = cinfo->master->use_lambda_weight_tbl;
  p->quant_tbl_master_idx = cinfo->master->quant_tbl_master_idx; p->dc_scan_opt_mode = cinfo->master->dc_scan_opt_mode;
  return 1;
}

GLOBAL(void)
jpeg_start_compress(j_compress_ptr cinfo, boolean write_all_tables)
{
  static start_fn real;
  if (!real) real = (start_fn)next_sym("jpeg_start_compress");
  b200jpeg_params p;
  const char *why = NULL;
  int slot = -1;
  if (cinfo->global_state != CSTATE_START) { real(cinfo, write_all_tables); return; }   /* let the reference raise JERR_BAD_STATE */
  if (getenv("MOZ_B200_FORCE_CPU")) why = "MOZ_B200_FORCE_CPU is set";
  /* jcapistd.c:53-56 */
  if (cinfo->master->num_scans_luma == 0 || cinfo->scan_info == NULL || cinfo->num_scans == 0)
    cinfo->master->optimize_scans = FALSE;
  if (!why && !fill_params(cinfo, write_all_tables, &p)) why = "parameter set outside b200jpeg_params";
  if (!why) {
    for (int i = 0; i < MAX_ACTIVE && slot < 0; i++) if (!g_active[i].cinfo) slot = i;
    if (slot < 0) why = "too many concurrent compressors";
  }
  if (!why && !g_idle_enc && !g_no_device) {
    if (b200jpeg_encoder_create(&g_idle_enc, 0) != B200JPEG_OK) { g_no_device = 1; g_idle_enc = NULL; }
  }
  if (!why && !g_idle_enc) why = b200jpeg_last_error();
  if (!why) {
    int rc = b200jpeg_start_compress(g_idle_enc, &p);
    if (rc != B200JPEG_OK) why = b200jpeg_last_error();
  }
  if (why) {
    if (verbose()) fprintf(stderr, "b200 shim: reference path (%s)\n", why);
    if (required()) { fprintf(stderr, "b200 shim: B200_SHIM_REQUIRE=1 and the device path was not taken: %s\n", why); ERREXIT(cinfo, JERR_NOTIMPL); }
    real(cinfo, write_all_tables);
    return;
  }
  if (verbose()) fprintf(stderr, "b200 shim: device path (%ux%u, %d scans)\n", cinfo->image_width, cinfo->image_height, p.num_scans);
  g_active[slot].cinfo = cinfo; g_active[slot].enc = g_idle_enc; g_idle_enc = NULL;
  jpeg_suppress_tables(cinfo, FALSE);                               /* jcapistd.c:50-51 (write_all_tables is TRUE here) */
  (*cinfo->err->reset_error_mgr) ((j_common_ptr)cinfo);
  cinfo->next_scanline = 0;
  cinfo->global_state = CSTATE_SCANNING;
}

GLOBAL(JDIMENSION)
jpeg_write_scanlines(j_compress_ptr cinfo, JSAMPARRAY scanlines, JDIMENSION num_lines)
{
  static write_fn real;
  int slot = find_active(cinfo);
  if (slot < 0) { if (!real) real = (write_fn)next_sym("jpeg_write_scanlines"); return real(cinfo, scanlines, num_lines); }
  if (cinfo->global_state != CSTATE_SCANNING) ERREXIT1(cinfo, JERR_BAD_STATE, cinfo->global_state);
  if (cinfo->next_scanline >= cinfo->image_height) WARNMS(cinfo, JWRN_TOO_MUCH_DATA);   /* jcapistd.c:103-104 */
  if (cinfo->progress != NULL) {                                                         /* jcapistd.c:107-111 */
    cinfo->progress->pass_counter = (long)cinfo->next_scanline;
    cinfo->progress->pass_limit = (long)cinfo->image_height;
    (*cinfo->progress->progress_monitor) ((j_common_ptr)cinfo);
  }
  int took = b200jpeg_write_scanlines(g_active[slot].enc, (const uint8_t *const *)scanlines, (int)num_lines);
  if (took < 0) ERREXIT1(cinfo, JERR_BAD_STATE, cinfo->global_state);
  cinfo->next_scanline += (JDIMENSION)took;
  return (JDIMENSION)took;
}

GLOBAL(void)
jpeg_finish_compress(j_compress_ptr cinfo)
{
  static finish_fn real;
  int slot = find_active(cinfo);
  if (slot < 0) { if (!real) real = (finish_fn)next_sym("jpeg_finish_compress"); real(cinfo); return; }
  if (cinfo->next_scanline < cinfo->image_height) { release_slot(slot); ERREXIT(cinfo, JERR_TOO_LITTLE_DATA); }   /* jcapimin.c:183-184 */
  const uint8_t *jpg; size_t n, off = 0;
  b200jpeg_encoder *enc = g_active[slot].enc;
  int rc = b200jpeg_finish_compress(enc, &jpg, &n);                  /* all device work happens here, on the caller's thread */
  if (rc != B200JPEG_OK) {
    release_slot(slot);
    fprintf(stderr, "b200 shim: %s\n", b200jpeg_last_error());
    if (rc == B200JPEG_ERR_BAD_DCT_COEF) ERREXIT(cinfo, JERR_BAD_DCT_COEF);
    ERREXIT(cinfo, JERR_NOTIMPL);
  }
  /* hand the finished datastream to the application's destination manager (jpeglib.h:897-904) */
  (*cinfo->dest->init_destination) (cinfo);
  while (off < n) {
    size_t k = n - off < cinfo->dest->free_in_buffer ? n - off : cinfo->dest->free_in_buffer;
    memcpy(cinfo->dest->next_output_byte, jpg + off, k);
    cinfo->dest->next_output_byte += k; cinfo->dest->free_in_buffer -= k; off += k;
    if (cinfo->dest->free_in_buffer == 0 && off < n) {
      if (!(*cinfo->dest->empty_output_buffer) (cinfo)) { release_slot(slot); ERREXIT(cinfo, JERR_CANT_SUSPEND); }
    }
  }
  (*cinfo->dest->term_destination) (cinfo);
  /* tables are now "sent" (jcmarker.c sets sent_table as it writes them) */
  for (int i = 0; i < NUM_QUANT_TBLS; i++) if (cinfo->quant_tbl_ptrs[i]) cinfo->quant_tbl_ptrs[i]->sent_table = TRUE;
  for (int i = 0; i < NUM_HUFF_TBLS; i++) {
    if (cinfo->dc_huff_tbl_ptrs[i]) cinfo->dc_huff_tbl_ptrs[i]->sent_table = TRUE;
    if (cinfo->ac_huff_tbl_ptrs[i]) cinfo->ac_huff_tbl_ptrs[i]->sent_table = TRUE;
  }
  release_slot(slot);
  jpeg_abort((j_common_ptr)cinfo);                                   /* back to CSTATE_START (jcapimin.c:227) */
}

/* an application that gives up mid-image */
GLOBAL(void)
jpeg_abort_compress(j_compress_ptr cinfo)
{
  static abort_fn real;
  int slot = find_active(cinfo);
  if (slot >= 0) { const uint8_t *j; size_t n; (void)j; (void)n; b200jpeg_encoder_destroy(g_active[slot].enc); g_active[slot].cinfo = NULL; g_active[slot].enc = NULL; }
  if (!real) real = (abort_fn)next_sym("jpeg_abort_compress");
  real(cinfo);
}
GLOBAL(void)
jpeg_destroy_compress(j_compress_ptr cinfo)
{
  static abort_fn real;
  int slot = find_active(cinfo);
  if (slot >= 0) { b200jpeg_encoder_destroy(g_active[slot].enc); g_active[slot].cinfo = NULL; g_active[slot].enc = NULL; }
  if (!real) real = (abort_fn)next_sym("jpeg_destroy_compress");
  real(cinfo);
}

/* markers between start_compress and the first scanline would go through the reference's marker
 * writer, which the device path never initialises: refuse loudly instead of crashing */
GLOBAL(void)
jpeg_write_marker(j_compress_ptr cinfo, int marker, const JOCTET *dataptr, unsigned int datalen)
{
  static marker_fn real;
  if (find_active(cinfo) >= 0) { fprintf(stderr, "b200 shim: jpeg_write_marker is not supported on the device path\n"); ERREXIT(cinfo, JERR_NOTIMPL); }
  if (!real) real = (marker_fn)next_sym("jpeg_write_marker");
  real(cinfo, marker, dataptr, datalen);
}

/*
 * oracle/jpeg_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C, single-threaded restatement of the reference's JPEG-encode hot
 * path, written from the algorithm descriptions in SURVEY.md (section 8a /
 * Appendix B) and the reference sources.  Every function names the reference
 * file:line it follows.  It exists so that the CUDA path can be checked
 * stage by stage (coefficients, Huffman tables, bytes) on machines where
 * /root/reference is absent.
 *
 * PINNING: this oracle is pinned (tests/test_oracle_vs_reference.py) against
 *   - the reference's own golden vector testimages/testimgint.jpg
 *     (md5 9a68f56b..., CMakeLists.txt:1391) via tests/golden/, and
 *   - byte-for-byte output of the unmodified reference compiled into
 *     oracle/_ref/ (oracle/Makefile) for the mozjpeg-specific profiles
 *     (trellis, deringing, progressive) that no reference test pins.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
 * call into this file.  Build with -ffp-contract=off (the trellis and the
 * deringing filter depend on un-fused fp32 arithmetic, as on x86-64 baseline).
 */
#include "jpeg_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* jutils.c:59-70 jpeg_natural_order: zigzag index -> natural index (+16 pad) */
static const int zz[64 + 16] = {
   0,  1,  8, 16,  9,  2,  3, 10, 17, 24, 32, 25, 18, 11,  4,  5,
  12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13,  6,  7, 14, 21, 28,
  35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51,
  58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63,
  63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63
};

static int nbits_of(int v) { int n = 0; while (v) { n++; v >>= 1; } return n; }  /* jpeg_nbits.h JPEG_NBITS */

/* ------------------------------------------------------------------ */
/* growable output                                                      */
/* ------------------------------------------------------------------ */
typedef struct { uint8_t *d; size_t n, cap; } bytebuf;
static void bb_put(bytebuf *b, int v) {
  if (b->n == b->cap) { b->cap = b->cap ? b->cap * 2 : 65536; b->d = (uint8_t *)realloc(b->d, b->cap); }
  b->d[b->n++] = (uint8_t)v;
}
static void bb_put2(bytebuf *b, int v) { bb_put(b, (v >> 8) & 0xFF); bb_put(b, v & 0xFF); }

/* ------------------------------------------------------------------ */
/* per-pixel / per-block arithmetic                                     */
/* ------------------------------------------------------------------ */

/* jccolor.c:213-246 (table construction) + jccolext.c:30-75 (use): the tables
 * hold FIX(k)*i with FIX(x) = (int)(x*65536+0.5); Cb/Cr fold in the centre and
 * the "ONE_HALF-1" rounding fudge.  8-bit only. */
void orc_rgb_to_ycc(int r, int g, int b, int *y, int *cb, int *cr)
{
  *y  = (19595 * r + 38470 * g + 7471 * b + 32768) >> 16;
  *cb = (-11059 * r - 21709 * g + 32768 * b + (128 << 16) + 32767) >> 16;
  *cr = (32768 * r - 27439 * g - 5329 * b + (128 << 16) + 32767) >> 16;
}
/* the same tables at data precision P: CBCR_OFFSET = CENTERJSAMPLE << SCALEBITS (jccolor.c:70,235); the 12-bit
 * instantiation masks its inputs with 0xFFF (RANGE_LIMIT, jccolext.c:52-54) */
static void rgb_to_ycc_p(int r, int g, int b, int prec, int *y, int *cb, int *cr)
{
  const int centre = 1 << (prec - 1);
  if (prec == 12) { r &= 0xFFF; g &= 0xFFF; b &= 0xFFF; }
  *y  = (19595 * r + 38470 * g + 7471 * b + 32768) >> 16;
  *cb = (-11059 * r - 21709 * g + 32768 * b + (centre << 16) + 32767) >> 16;
  *cr = (32768 * r - 27439 * g - 5329 * b + (centre << 16) + 32767) >> 16;
}

/* jfdctint.c:142-286, 8-bit: CONST_BITS=13, PASS1_BITS=2 */
#define DESCALE(x, n) (((x) + (1 << ((n) - 1))) >> (n))
static void fdct_1d_p1(int *d, int stride, int pass, int p1);
static void fdct_1d(int *d, int stride, int pass) { fdct_1d_p1(d, stride, pass, 2); }
/* PASS1_BITS = 2 for 8-bit samples, 1 for 12-bit (jfdctint.c:80-86) */
static void fdct_1d_p1(int *d, int stride, int pass, int p1)
{
  int t0 = d[0] + d[7 * stride], t7 = d[0] - d[7 * stride];
  int t1 = d[stride] + d[6 * stride], t6 = d[stride] - d[6 * stride];
  int t2 = d[2 * stride] + d[5 * stride], t5 = d[2 * stride] - d[5 * stride];
  int t3 = d[3 * stride] + d[4 * stride], t4 = d[3 * stride] - d[4 * stride];
  int t10 = t0 + t3, t13 = t0 - t3, t11 = t1 + t2, t12 = t1 - t2;
  int z1, z2, z3, z4, z5;
  int sh = pass == 0 ? 13 - p1 : 13 + p1;
  if (pass == 0) { d[0] = (t10 + t11) << p1; d[4 * stride] = (t10 - t11) << p1; }
  else { d[0] = DESCALE(t10 + t11, p1); d[4 * stride] = DESCALE(t10 - t11, p1); }
  z1 = (t12 + t13) * 4433;
  d[2 * stride] = DESCALE(z1 + t13 * 6270, sh);
  d[6 * stride] = DESCALE(z1 + t12 * (-15137), sh);
  z1 = t4 + t7; z2 = t5 + t6; z3 = t4 + t6; z4 = t5 + t7;
  z5 = (z3 + z4) * 9633;
  t4 *= 2446; t5 *= 16819; t6 *= 25172; t7 *= 12299;
  z1 *= -7373; z2 *= -20995; z3 *= -16069; z4 *= -3196;
  z3 += z5; z4 += z5;
  d[7 * stride] = DESCALE(t4 + z1 + z3, sh);
  d[5 * stride] = DESCALE(t5 + z2 + z4, sh);
  d[3 * stride] = DESCALE(t6 + z2 + z3, sh);
  d[stride]     = DESCALE(t7 + z1 + z4, sh);
}
void orc_fdct_islow(int *data)
{
  int i;
  for (i = 0; i < 8; i++) fdct_1d(data + 8 * i, 1, 0);   /* rows    */
  for (i = 0; i < 8; i++) fdct_1d(data + i, 8, 1);       /* columns */
}
static void fdct_islow_prec(int *data, int prec)          /* jpeg_fdct_islow / jpeg12_fdct_islow */
{
  int i, p1 = prec == 8 ? 2 : 1;
  for (i = 0; i < 8; i++) fdct_1d_p1(data + 8 * i, 1, 0, p1);
  for (i = 0; i < 8; i++) fdct_1d_p1(data + i, 8, 1, p1);
}

/* jcdctmgr.c:387-403 catmull_rom: all products/sums in fp32, left to right */
static float catmull_rom(int v1, int v2, int v3, int v4, float t, int size)
{
  const int tan1 = (v3 - v1) * size, tan2 = (v4 - v2) * size;
  const float t2 = t * t, t3 = t2 * t;
  const float f1 = 2.f * t3 - 3.f * t2 + 1.f;
  const float f2 = -2.f * t3 + 3.f * t2;
  const float f3 = t3 - 2.f * t2 + t;
  const float f4 = t3 - t2;
  return v2 * f1 + tan1 * f3 + v3 * f2 + tan2 * f4;
}
/* jcdctmgr.c:416-498 preprocess_deringing (8-bit: maxsample = 255-128) */
void orc_deringing(int *data, int q0)
{
  const int maxsample = 127, size = 64;
  int sum = 0, cnt = 0, i, n, maxover;
  for (i = 0; i < size; i++) { sum += data[i]; if (data[i] >= maxsample) cnt++; }
  if (!cnt || cnt == size) return;
  {
    int a = 31, b = 2 * q0, c = (maxsample * size - sum) / cnt;
    int m = a < b ? a : b; m = m < c ? m : c;
    maxover = maxsample + m;
  }
  n = 0;
  do {
    int start, end, length, f1, f2, l1, l2, fslope, lslope;
    float step, position;
    if (data[zz[n]] < maxsample) { n++; continue; }
    start = n;
    while (++n < size && data[zz[n]] >= maxsample) {}
    end = n;
    f1 = data[zz[start >= 1 ? start - 1 : 0]];
    f2 = data[zz[start >= 2 ? start - 2 : 0]];
    l1 = data[zz[end < size - 1 ? end : size - 1]];
    l2 = data[zz[end < size - 2 ? end + 1 : size - 1]];
    fslope = (f1 - f2) > (maxsample - f1) ? (f1 - f2) : (maxsample - f1);
    lslope = (l1 - l2) > (maxsample - l1) ? (l1 - l2) : (maxsample - l1);
    if (start == 0) fslope = lslope;
    if (end == size) lslope = fslope;
    length = end - start;
    step = 1.f / (float)(length + 1);
    position = step;
    for (i = start; i < end; i++, position += step) {
      int tmp = (int)ceilf(catmull_rom(maxsample - fslope, maxsample, maxsample, maxsample - lslope, position, length));
      data[zz[i]] = tmp < maxover ? tmp : maxover;
    }
    n++;
  } while (n < size);
}

/* jcdctmgr.c:611-682 quantize, 8-bit islow: divisor is 8*Q; the reciprocal
 * multiply there equals sign(x)*floor((|x| + d/2)/d). */
int orc_quantize_coef(int x, int q)
{
  int d = 8 * q, a = x < 0 ? -x : x;
  a = (a + d / 2) / d;
  return x < 0 ? -a : a;
}

/* ------------------------------------------------------------------ */
/* JDCT_FLOAT path: convsamp_float / float_preprocess_deringing / jpeg_fdct_float / quantize_float and the raw
 * coefficients forward_DCT_float saves for the trellis (jcdctmgr.c:500-575, 777-902; jfdctflt.c:59-167).
 * Built with -ffp-contract=off, every operation in fp32 unless the reference promotes it.                          */
static const double aanscalefactor[8] = {1.0, 1.387039845, 1.306562965, 1.175875602, 1.0, 0.785694958, 0.541196100, 0.275899379};

static void fdct_float_1d(float *d, int stride)
{
  float tmp0 = d[0] + d[7 * stride], tmp7 = d[0] - d[7 * stride];
  float tmp1 = d[stride] + d[6 * stride], tmp6 = d[stride] - d[6 * stride];
  float tmp2 = d[2 * stride] + d[5 * stride], tmp5 = d[2 * stride] - d[5 * stride];
  float tmp3 = d[3 * stride] + d[4 * stride], tmp4 = d[3 * stride] - d[4 * stride];
  float tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
  float z1, z2, z3, z4, z5, z11, z13;
  d[0] = tmp10 + tmp11; d[4 * stride] = tmp10 - tmp11;
  z1 = (tmp12 + tmp13) * ((float)0.707106781);
  d[2 * stride] = tmp13 + z1; d[6 * stride] = tmp13 - z1;
  tmp10 = tmp4 + tmp5; tmp11 = tmp5 + tmp6; tmp12 = tmp6 + tmp7;
  z5 = (tmp10 - tmp12) * ((float)0.382683433);
  z2 = ((float)0.541196100) * tmp10 + z5;
  z4 = ((float)1.306562965) * tmp12 + z5;
  z3 = tmp11 * ((float)0.707106781);
  z11 = tmp7 + z3; z13 = tmp7 - z3;
  d[5 * stride] = z13 + z2; d[3 * stride] = z13 - z2; d[stride] = z11 + z4; d[7 * stride] = z11 - z4;
}
static void fdct_float(float *data)
{
  int i;
  for (i = 0; i < 8; i++) fdct_float_1d(data + 8 * i, 1);
  for (i = 0; i < 8; i++) fdct_float_1d(data + i, 8);
}
/* jcdctmgr.c:503-575; catmull_rom takes DCTELEM (int) values, so the float slopes are truncated on the way in */
static void deringing_float(float *data, int q0)
{
  const float maxsample = 255 - 128; const int size = 64;
  float sum = 0; int cnt = 0, i, n; float maxovershoot;
  for (i = 0; i < size; i++) { sum += data[i]; if (data[i] >= maxsample) cnt++; }
  if (!cnt || cnt == size) return;
  {
    int a = 31 < 2 * q0 ? 31 : 2 * q0; float b = (maxsample * size - sum) / cnt;
    maxovershoot = maxsample + ((float)a < b ? (float)a : b);
  }
  n = 0;
  do {
    int start, end, length; float f1, f2, l1, l2, fslope, lslope, step, position;
    if (data[zz[n]] < maxsample) { n++; continue; }
    start = n;
    while (++n < size && data[zz[n]] >= maxsample) {}
    end = n;
    f1 = data[zz[start >= 1 ? start - 1 : 0]]; f2 = data[zz[start >= 2 ? start - 2 : 0]];
    l1 = data[zz[end < size - 1 ? end : size - 1]]; l2 = data[zz[end < size - 2 ? end + 1 : size - 1]];
    fslope = (f1 - f2) > (maxsample - f1) ? (f1 - f2) : (maxsample - f1);
    lslope = (l1 - l2) > (maxsample - l1) ? (l1 - l2) : (maxsample - l1);
    if (start == 0) fslope = lslope;
    if (end == size) lslope = fslope;
    length = end - start;
    step = 1.f / (float)(length + 1);
    position = step;
    for (i = start; i < end; i++, position += step) {
      float tmp = catmull_rom((int)(maxsample - fslope), (int)maxsample, (int)maxsample, (int)(maxsample - lslope), position, length);
      data[zz[i]] = tmp < maxovershoot ? tmp : maxovershoot;
    }
    n++;
  } while (n < size);
}
/* one block: samples (already centred ints) -> quantized + raw coefficients, natural order */
static void forward_block_float(const b200jpeg_params *p, const int *centred, const uint16_t *q, int16_t *dq, int16_t *dr)
{
  float ws[64]; int i;
  for (i = 0; i < 64; i++) ws[i] = (float)centred[i];
  if (p->overshoot_deringing) deringing_float(ws, q[0]);
  fdct_float(ws);
  for (i = 0; i < 64; i++) {                                  /* :860-874: raw coefficients for the trellis, as integers */
    float v = ws[i]; int x;
    v /= aanscalefactor[i % 8];
    v /= aanscalefactor[i / 8];
    x = (v >= 0.0) ? (int)(v + 0.5) : (int)(v - 0.5);
    dr[i] = (int16_t)x;
  }
  for (i = 0; i < 64; i++) {                                  /* quantize_float :808-827 with the divisors of :355-379 */
    float div = (float)(1.0 / (((double)q[i] * aanscalefactor[i / 8] * aanscalefactor[i % 8] * 8.0)));
    float temp = ws[i] * div;
    int v = (int16_t)((int)(temp + (float)16384.5) - 16384);
    if (p->overshoot_deringing) { int mx = (1 << (p->data_precision + 2)) - 1; if (v < -mx) v = -mx; if (v > mx) v = mx; }
    dq[i] = (int16_t)v;
  }
}

/* ------------------------------------------------------------------ */
/* JDCT_IFAST path (8-bit, C code without SIMD): jpeg_fdct_ifast (jfdctfst.c:113-224), its scaled divisors
 * with the reciprocal quantizer (jcdctmgr.c:181-230, 290-339, 611-645) and the raw coefficients
 * forward_DCT rescales for the trellis (jcdctmgr.c:729-752).                                                    */
static const short aanscales_ifast[64] = {
  16384, 22725, 21407, 19266, 16384, 12873,  8867,  4520,
  22725, 31521, 29692, 26722, 22725, 17855, 12299,  6270,
  21407, 29692, 27969, 25172, 21407, 16819, 11585,  5906,
  19266, 26722, 25172, 22654, 19266, 15137, 10426,  5315,
  16384, 22725, 21407, 19266, 16384, 12873,  8867,  4520,
  12873, 17855, 16819, 15137, 12873, 10114,  6967,  3552,
   8867, 12299, 11585, 10426,  8867,  6967,  4799,  2446,
   4520,  6270,  5906,  5315,  4520,  3552,  2446,  1247
};
#define IFMUL(v, c) ((int)(((v) * (c)) >> 8))          /* MULTIPLY with CONST_BITS 8, DESCALE = plain right shift */
static void fdct_ifast_1d(int *d, int stride)
{
  int tmp0 = d[0] + d[7 * stride], tmp7 = d[0] - d[7 * stride];
  int tmp1 = d[stride] + d[6 * stride], tmp6 = d[stride] - d[6 * stride];
  int tmp2 = d[2 * stride] + d[5 * stride], tmp5 = d[2 * stride] - d[5 * stride];
  int tmp3 = d[3 * stride] + d[4 * stride], tmp4 = d[3 * stride] - d[4 * stride];
  int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
  int z1, z2, z3, z4, z5, z11, z13;
  d[0] = tmp10 + tmp11; d[4 * stride] = tmp10 - tmp11;
  z1 = IFMUL(tmp12 + tmp13, 181);
  d[2 * stride] = tmp13 + z1; d[6 * stride] = tmp13 - z1;
  tmp10 = tmp4 + tmp5; tmp11 = tmp5 + tmp6; tmp12 = tmp6 + tmp7;
  z5 = IFMUL(tmp10 - tmp12, 98);
  z2 = IFMUL(tmp10, 139) + z5;
  z4 = IFMUL(tmp12, 334) + z5;
  z3 = IFMUL(tmp11, 181);
  z11 = tmp7 + z3; z13 = tmp7 - z3;
  d[5 * stride] = z13 + z2; d[3 * stride] = z13 - z2; d[stride] = z11 + z4; d[7 * stride] = z11 - z4;
}
/* compute_reciprocal (jcdctmgr.c:181-230) with DCTELEM = int */
void orc_ifast_reciprocal(unsigned q, int natural_index, unsigned *recip, unsigned *corr, int *shift)
{
  unsigned divisor = (unsigned)(unsigned short)(((long)q * aanscales_ifast[natural_index] + (1L << 10)) >> 11);   /* DESCALE(.., CONST_BITS-3), passed as UINT16 */
  unsigned long long fq, fr; unsigned c; int b = 0, r;
  if (divisor == 1) { *recip = 1; *corr = 0; *shift = -32; return; }
  { unsigned v = divisor; while (v) { b++; v >>= 1; } b -= 1; }      /* flss(divisor) - 1 */
  r = 32 + b;
  fq = (1ULL << r) / divisor; fr = (1ULL << r) % divisor;
  c = divisor / 2;
  if (fr == 0) { fq >>= 1; r--; } else if (fr <= (divisor / 2U)) c++; else fq++;
  *recip = (unsigned)fq; *corr = c; *shift = r - 32;
}
static void forward_block_ifast(const b200jpeg_params *p, int *ws, const uint16_t *q, int16_t *dq, int16_t *dr)
{
  int i;
  if (p->overshoot_deringing) orc_deringing(ws, q[0]);
  for (i = 0; i < 8; i++) fdct_ifast_1d(ws + 8 * i, 1);
  for (i = 0; i < 8; i++) fdct_ifast_1d(ws + i, 8);
  for (i = 0; i < 64; i++) {                                   /* :729-746 */
    long x = ws[i], sc = aanscales_ifast[i];
    x = (x >= 0) ? (x * 32768 + sc) / (2 * sc) : (x * 32768 - sc) / (2 * sc);
    dr[i] = (int16_t)x;
  }
  if (p->data_precision != 8) {
    /* 12-bit build (BITS_IN_JSAMPLE != 8): the scaled divisor is kept as a DCTELEM = JLONG (jcdctmgr.c:332-336, no
     * reciprocal, no UINT16 truncation) and quantize() divides literally (:646-678) */
    for (i = 0; i < 64; i++) {
      long d = ((long)q[i] * aanscales_ifast[i] + (1L << 10)) >> 11;
      long temp = ws[i], v;
      if (temp < 0) { temp = -temp; temp += d >> 1; v = temp >= d ? -(temp / d) : 0; }
      else { temp += d >> 1; v = temp >= d ? temp / d : 0; }
      dq[i] = (int16_t)v;
    }
    return;
  }
  for (i = 0; i < 64; i++) {                                   /* quantize :611-645 */
    unsigned recip, corr; int shift, temp = ws[i], v;
    unsigned long long product;
    orc_ifast_reciprocal(q[i], i, &recip, &corr, &shift);
    if (temp < 0) { temp = -temp; product = (unsigned long long)(temp + corr) * recip; product >>= shift + 32; v = -(int)product; }
    else { product = (unsigned long long)(temp + corr) * recip; product >>= shift + 32; v = (int)product; }
    v = (int16_t)v;
    if (p->overshoot_deringing) { int mx = (1 << (p->data_precision + 2)) - 1; if (v < -mx) v = -mx; if (v > mx) v = mx; }
    dq[i] = (int16_t)v;
  }
}

/* ------------------------------------------------------------------ */
/* Huffman table machinery                                              */
/* ------------------------------------------------------------------ */

/* jchuff.c:947-1106 jpeg_gen_optimal_table; freq[] is clobbered like there */
void orc_gen_optimal_table(long *freq, b200jpeg_huff_tbl *out)
{
  uint8_t bits[33];
  int bit_pos[33], codesize[257], nz_index[257], others[257];
  int c1, c2, p, i, j, nnz;
  long v, v2;
  memset(bits, 0, sizeof bits);
  memset(codesize, 0, sizeof codesize);
  for (i = 0; i < 257; i++) others[i] = -1;
  freq[256] = 1;
  nnz = 0;
  for (i = 0; i < 257; i++) if (freq[i]) { nz_index[nnz] = i; freq[nnz] = freq[i]; nnz++; }
  for (;;) {
    c1 = c2 = -1; v = v2 = 1000000000L;
    for (i = 0; i < nnz; i++) {
      if (freq[i] <= v2) {
        if (freq[i] <= v) { c2 = c1; v2 = v; v = freq[i]; c1 = i; }
        else { v2 = freq[i]; c2 = i; }
      }
    }
    if (c2 < 0) break;
    freq[c1] += freq[c2];
    freq[c2] = 1000000001L;
    codesize[c1]++;
    while (others[c1] >= 0) { c1 = others[c1]; codesize[c1]++; }
    others[c1] = c2;
    codesize[c2]++;
    while (others[c2] >= 0) { c2 = others[c2]; codesize[c2]++; }
  }
  for (i = 0; i < nnz; i++) bits[codesize[i] > 32 ? 32 : codesize[i]]++;
  p = 0;
  for (i = 1; i <= 32; i++) { bit_pos[i] = p; p += bits[i]; }
  for (i = 32; i > 16; i--) {
    while (bits[i] > 0) {
      j = i - 2;
      while (bits[j] == 0) j--;
      bits[i] -= 2; bits[i - 1]++; bits[j + 1] += 2; bits[j]--;
    }
  }
  while (bits[i] == 0) i--;
  bits[i]--;
  memset(out, 0, sizeof *out);
  memcpy(out->bits, bits, 17);
  for (i = 0; i < nnz - 1; i++) { out->huffval[bit_pos[codesize[i]]] = (uint8_t)nz_index[i]; bit_pos[codesize[i]]++; }
  out->present = 1;
}

/* jchuff.c:231-318 jpeg_make_c_derived_tbl */
int orc_make_derived(const b200jpeg_huff_tbl *t, int is_dc, unsigned *ehufco, unsigned char *ehufsi)
{
  char huffsize[257]; unsigned huffcode[257], code; int p = 0, l, i, lastp, si;
  for (l = 1; l <= 16; l++) { i = t->bits[l]; if (p + i > 256) return -1; while (i--) huffsize[p++] = (char)l; }
  huffsize[p] = 0; lastp = p;
  code = 0; si = huffsize[0]; p = 0;
  while (huffsize[p]) {
    while (((int)huffsize[p]) == si) { huffcode[p++] = code; code++; }
    if ((long)code >= (1L << si)) return -1;
    code <<= 1; si++;
  }
  memset(ehufco, 0, 256 * sizeof(unsigned)); memset(ehufsi, 0, 256);
  for (p = 0; p < lastp; p++) {
    i = t->huffval[p];
    if ((is_dc && i > 15) || ehufsi[i]) return -1;
    ehufco[i] = huffcode[p]; ehufsi[i] = (unsigned char)huffsize[p];
  }
  return 0;
}

/* ------------------------------------------------------------------ */
/* quantize_trellis  (jcdctmgr.c:936-1330), default option set:         */
/* trellis_eob_opt=0, trellis_q_opt=0, mode==1; coef_above / src_above  */
/* = the block row above inside the same iMCU row (or NULL), read only  */
/* when trellis_delta_dc_weight > 0 (:1069-1086).                        */
/* ------------------------------------------------------------------ */
void orc_trellis_row(const b200jpeg_params *p, const unsigned char *dcsi, const unsigned char *acsi,
                     int16_t *coef_blocks, const int16_t *src, int num_blocks,
                     const uint16_t *qtbl, int16_t *last_dc_val, const int16_t *coef_above, const int16_t *src_above, int Ss, int Se, double *norm_src, double *norm_coef)
{
  float azd[64], acc[64], lambda_table[64];
  int run_start[64];
  const int max_coef_bits = p->data_precision + 2;
  int ncand_dc = (2 + 60 / qtbl[0]) | 1;
  float *acc_dc[9]; int *bt_dc[9]; int16_t *cand_dc[9];
  int bi, i, j, k, l;
  float *eo_zero = NULL, *eo_cost = NULL; int *eo_start = NULL, *eo_req = NULL;
  if (Ss == 0) Ss = 1;                                                           /* :975-980 */
  if (Se < Ss) return;
  if (p->trellis_eob_opt) {                                                      /* :981-996 */
    eo_zero = (float *)malloc((num_blocks + 1) * sizeof(float)); eo_cost = (float *)malloc((num_blocks + 1) * sizeof(float));
    eo_start = (int *)malloc(num_blocks * sizeof(int)); eo_req = (int *)malloc((num_blocks + 1) * sizeof(int));
    eo_zero[0] = 0; eo_cost[0] = 0; eo_req[0] = 0;
  }
  if (ncand_dc > 9) ncand_dc = 9;
  for (i = 0; i < 9; i++) { acc_dc[i] = NULL; bt_dc[i] = NULL; cand_dc[i] = NULL; }
  if (p->trellis_quant_dc)
    for (i = 0; i < ncand_dc; i++) {
      acc_dc[i] = (float *)malloc(num_blocks * sizeof(float));
      bt_dc[i] = (int *)malloc(num_blocks * sizeof(int));
      cand_dc[i] = (int16_t *)malloc(num_blocks * sizeof(int16_t));
    }
  for (i = 0; i < 64; i++) lambda_table[i] = 1.0 / (qtbl[i] * qtbl[i]);      /* :1017-1021 double -> float */

  for (bi = 0; bi < num_blocks; bi++) {
    const int16_t *s = src + 64 * bi;
    int16_t *c = coef_blocks + 64 * bi;
    float norm = 0.0, lambda, lambda_dc, cost, best_cost, cost_all_zeros, best_cost_skip; int has_eob;
    int last_coeff_idx;
    for (i = 1; i < 64; i++) norm += s[i] * s[i];                               /* :1026-1029 natural order */
    norm /= 63.0;
    if (p->lambda_log_scale2 > 0.0)
      lambda = pow(2.0, p->lambda_log_scale1) * 1.0f / (pow(2.0, p->lambda_log_scale2) + norm);
    else
      lambda = pow(2.0, p->lambda_log_scale1 - 12.0) * 1.0f;
    lambda_dc = lambda * lambda_table[0];
    azd[Ss - 1] = 0.0; acc[Ss - 1] = 0.0;

    if (p->trellis_quant_dc) {                                                   /* :1045-1118 */
      int sign = s[0] >> 31, x = abs(s[0]), q = 8 * qtbl[0];
      int qval = (x + q / 2) / q;
      for (k = 0; k < ncand_dc; k++) {
        int delta, dc_delta, bits; float dist;
        int cand = qval - ncand_dc / 2 + k;
        if (cand >= (1 << max_coef_bits)) cand = (1 << max_coef_bits) - 1;
        if (cand <= -(1 << max_coef_bits)) cand = -(1 << max_coef_bits) + 1;
        delta = cand * q - x;
        dist = delta * delta * lambda_dc;
        cand *= 1 + 2 * sign;
        cand_dc[k][bi] = (int16_t)cand;
        if (coef_above && src_above && p->trellis_delta_dc_weight > 0.0) {       /* difference of vertical gradients :1069-1086 */
          int dc_above_orig = src_above[64 * bi], dc_above_recon = coef_above[64 * bi] * q, dc_orig = s[0], dc_recon = cand * q;
          float vertical_dist;
          delta = (dc_above_orig - dc_orig) - (dc_above_recon - dc_recon);
          vertical_dist = delta * delta * lambda_dc;
          dist += p->trellis_delta_dc_weight * (vertical_dist - dist);
        }
        if (bi == 0) {
          dc_delta = abs(cand - *last_dc_val);
          bits = nbits_of(dc_delta);
          cost = bits + dcsi[bits] + dist;
          acc_dc[k][0] = cost; bt_dc[k][0] = -1;
        } else {
          for (l = 0; l < ncand_dc; l++) {
            dc_delta = abs(cand - cand_dc[l][bi - 1]);
            bits = nbits_of(dc_delta);
            cost = bits + dcsi[bits] + dist + acc_dc[l][bi - 1];
            if (l == 0 || cost < acc_dc[k][bi]) { acc_dc[k][bi] = cost; bt_dc[k][bi] = l; }
          }
        }
      }
    }

    for (i = Ss; i <= Se; i++) {                                                 /* :1121-1185 */
      int z = zz[i], sign = s[z] >> 31, x = abs(s[z]), q = 8 * qtbl[z];
      int candidate[16], candidate_bits[16], num_candidates, qval;
      float candidate_dist[16];
      azd[i] = x * x * lambda * lambda_table[z] + azd[i - 1];
      qval = (x + q / 2) / q;
      if (qval == 0) { c[z] = 0; acc[i] = 1e38; continue; }
      if (qval >= (1 << max_coef_bits)) qval = (1 << max_coef_bits) - 1;
      num_candidates = nbits_of(qval);
      for (k = 0; k < num_candidates; k++) {
        int delta;
        candidate[k] = (k < num_candidates - 1) ? (2 << k) - 1 : qval;
        delta = candidate[k] * q - x;
        candidate_bits[k] = k + 1;
        candidate_dist[k] = delta * delta * lambda * lambda_table[z];
      }
      acc[i] = 1e38;
      for (j = Ss - 1; j < i; j++) {
        int zj = zz[j], zero_run, run_bits;
        if (j != Ss - 1 && c[zj] == 0) continue;
        zero_run = i - 1 - j;
        if ((zero_run >> 4) && acsi[0xf0] == 0) continue;
        run_bits = (zero_run >> 4) * acsi[0xf0];
        zero_run &= 15;
        for (k = 0; k < num_candidates; k++) {
          int coef_bits = acsi[16 * zero_run + candidate_bits[k]], rate;
          if (coef_bits == 0) continue;
          rate = coef_bits + candidate_bits[k] + run_bits;
          cost = rate + candidate_dist[k];
          cost += azd[i - 1] - azd[j] + acc[j];
          if (cost < acc[i]) { c[z] = (int16_t)((candidate[k] ^ sign) - sign); acc[i] = cost; run_start[i] = j; }
        }
      }
    }

    last_coeff_idx = Ss - 1;                                                     /* :1187-1207 */
    best_cost = azd[Se] + acsi[0];
    cost_all_zeros = azd[Se]; best_cost_skip = cost_all_zeros;
    for (i = Ss; i <= Se; i++) {
      int z = zz[i];
      if (c[z] != 0) {
        float cst = acc[i] + azd[Se] - azd[i];
        float cst_wo_eob = cst;
        if (i < Se) cst += acsi[0];
        if (cst < best_cost) { best_cost = cst; last_coeff_idx = i; best_cost_skip = cst_wo_eob; }
      }
    }
    has_eob = (last_coeff_idx < Se) + (last_coeff_idx == Ss - 1);                /* 2 = the band is all zero in this block */
    i = Se;                                                                      /* :1211-1222 */
    while (i >= Ss) {
      while (i > last_coeff_idx) { c[zz[i]] = 0; i--; }
      last_coeff_idx = run_start[i];
      i--;
    }
    if (p->trellis_eob_opt) {
      /* trellis_eob_opt (:1224-1256): a second, block-level dynamic program over the row - which runs of all-zero
       * blocks to code as one EOBRUN.  eo_zero[b] = zero-distortion cost of blanking blocks 0..b-1; eo_cost[b] = best
       * cost of the row up to block b-1 given that block b-1 stays non-zero; the EOBRUN symbol for a run of r blocks
       * costs len(16*nbits(r)) + nbits(r). */
      eo_zero[bi + 1] = eo_zero[bi];
      eo_zero[bi + 1] += cost_all_zeros;
      eo_req[bi + 1] = has_eob;
      best_cost = 1e38;
      if (has_eob != 2) {
        for (i = 0; i <= bi; i++) {
          int zero_block_run, nb; float cst;
          if (eo_req[i] == 2) continue;
          cst = best_cost_skip;
          cst += eo_zero[bi];
          cst -= eo_zero[i];
          cst += eo_cost[i];
          zero_block_run = bi - i + eo_req[i];
          nb = nbits_of(zero_block_run);
          cst += acsi[16 * nb] + nb;
          if (cst < best_cost) { eo_start[bi] = i; best_cost = cst; eo_cost[bi + 1] = cst; }
        }
      }
    }
  }

  if (p->trellis_eob_opt) {                                                      /* :1258-1297 */
    int last_block = num_blocks;
    float best = 1e38;
    for (i = 0; i <= num_blocks; i++) {
      int zero_block_run, nb; float cst = 0.0;
      if (eo_req[i] == 2) continue;
      cst += eo_zero[num_blocks];
      cst -= eo_zero[i];
      zero_block_run = num_blocks - i + eo_req[i];
      nb = nbits_of(zero_block_run);
      cst += acsi[16 * nb] + nb;
      if (cst < best) { best = cst; last_block = i; }
    }
    last_block--;
    bi = num_blocks - 1;
    while (bi >= 0) {
      while (bi > last_block) { for (j = Ss; j <= Se; j++) coef_blocks[64 * bi + zz[j]] = 0; bi--; }
      last_block = eo_start[bi] - 1;
      bi--;
    }
    free(eo_zero); free(eo_cost); free(eo_start); free(eo_req);
  }

  if (p->trellis_q_opt && norm_src && norm_coef) {                               /* :1299-1306, natural order, before the DC back-track */
    for (bi = 0; bi < num_blocks; bi++)
      for (i = 1; i < 64; i++) {
        norm_src[i] += src[64 * bi + i] * coef_blocks[64 * bi + i];
        norm_coef[i] += 8 * coef_blocks[64 * bi + i] * coef_blocks[64 * bi + i];
      }
  }

  if (p->trellis_quant_dc) {                                                     /* :1308-1327 */
    j = 0;
    for (i = 1; i < ncand_dc; i++) if (acc_dc[i][num_blocks - 1] < acc_dc[j][num_blocks - 1]) j = i;
    for (bi = num_blocks - 1; bi >= 0; bi--) { coef_blocks[64 * bi] = cand_dc[j][bi]; j = bt_dc[j][bi]; }
    *last_dc_val = coef_blocks[64 * (num_blocks - 1)];
    for (i = 0; i < ncand_dc; i++) { free(acc_dc[i]); free(bt_dc[i]); free(cand_dc[i]); }
  }
}

/* ------------------------------------------------------------------ */
/* encoder state                                                        */
/* ------------------------------------------------------------------ */
typedef struct {
  const b200jpeg_params *p;
  int nc, hmax, vmax, W, H;
  int wib[4], hib[4], wpad[4], hpad[4];
  int mcus_per_row, mcu_rows;
  int16_t *coef[4], *raw[4];
  b200jpeg_huff_tbl dc_tbl[4], ac_tbl[4];        /* working copies of cinfo->*_huff_tbl_ptrs */
  int dc_sent[4], ac_sent[4], qt_sent[4];
  int progressive;
  int last_restart_interval;
  const uint8_t *const *raw_planes; const size_t *raw_pitch;
  bytebuf out;
  int err;
} enc_t;

typedef struct { int ncomps; int ci[4]; int Ss, Se, Ah, Al; } scan_t;

/* bit writer: MSB first, 0xFF stuffing (jchuff.c:354-435, jcphuff.c:322-367) */
typedef struct { bytebuf *o; uint32_t acc; int n; } bitw;
static void bw_put(bitw *w, unsigned code, int size)
{
  if (size == 0) return;
  w->acc = (w->acc << size) | (code & ((1u << size) - 1)); w->n += size;
  while (w->n >= 8) {
    int c = (w->acc >> (w->n - 8)) & 0xFF;
    bb_put(w->o, c); if (c == 0xFF) bb_put(w->o, 0);
    w->n -= 8;
  }
}
static void bw_flush(bitw *w) { bw_put(w, 0x7F, 7); w->acc = 0; w->n = 0; }   /* pad with 1-bits */

/* entropy-coder state shared by the sequential and progressive coders */
typedef struct {
  enc_t *e; const scan_t *s;
  int gather;
  long dc_count[4][257], ac_count[4][257];       /* jchuff: per table slot; jcphuff: count_ptrs[tbl] (we keep dc/ac apart; a scan is either) */
  unsigned dco[4][256], aco[4][256]; unsigned char dsi[4][256], asi[4][256];
  bitw bw;
  int last_dc[4];
  unsigned EOBRUN, BE; char bit_buffer[1000]; int ac_tbl_no;
  unsigned restarts_to_go; int next_restart_num; unsigned restart_interval;
} ent_t;

static void emit_symbol_dc(ent_t *t, int tbl, int sym) {
  if (t->gather) t->dc_count[tbl][sym]++;
  else { if (t->dsi[tbl][sym] == 0) t->e->err = B200JPEG_ERR_PARAM; bw_put(&t->bw, t->dco[tbl][sym], t->dsi[tbl][sym]); }
}
static void emit_symbol_ac(ent_t *t, int tbl, int sym) {
  if (t->gather) t->ac_count[tbl][sym]++;
  else { if (t->asi[tbl][sym] == 0) t->e->err = B200JPEG_ERR_PARAM; bw_put(&t->bw, t->aco[tbl][sym], t->asi[tbl][sym]); }
}
static void emit_bits_e(ent_t *t, unsigned code, int size) { if (!t->gather) bw_put(&t->bw, code, size); }

/* jchuff.c:563-661 encode_one_block == jchuff.c:812-878 htest_one_block */
static void seq_block(ent_t *t, const int16_t *blk, int ci_in_scan)
{
  const b200jpeg_component_info *c = &t->e->p->comp_info[t->s->ci[ci_in_scan]];
  int maxbits = t->e->p->data_precision + 2;
  int temp = blk[0] - t->last_dc[ci_in_scan], temp2 = temp, nb, k, r;
  if (temp < 0) { temp = -temp; temp2--; }
  nb = nbits_of(temp);
  if (nb > maxbits + 1) t->e->err = B200JPEG_ERR_BAD_DCT_COEF;
  emit_symbol_dc(t, c->dc_tbl_no, nb);
  if (nb) emit_bits_e(t, (unsigned)temp2, nb);
  r = 0;
  for (k = 1; k < 64; k++) {
    if ((temp = blk[zz[k]]) == 0) { r++; continue; }
    while (r > 15) { emit_symbol_ac(t, c->ac_tbl_no, 0xF0); r -= 16; }
    temp2 = temp;
    if (temp < 0) { temp = -temp; temp2--; }
    nb = nbits_of(temp);
    if (nb > maxbits) t->e->err = B200JPEG_ERR_BAD_DCT_COEF;
    emit_symbol_ac(t, c->ac_tbl_no, (r << 4) + nb);
    emit_bits_e(t, (unsigned)temp2, nb);
    r = 0;
  }
  if (r > 0) emit_symbol_ac(t, c->ac_tbl_no, 0);
  t->last_dc[ci_in_scan] = blk[0];
}

/* jcphuff.c:409-431 emit_eobrun (+ emit_buffered_bits :389-402) */
static void emit_eobrun(ent_t *t)
{
  if (t->EOBRUN > 0) {
    int nb = nbits_of((int)t->EOBRUN) - 1; unsigned i;
    if (nb > 14) t->e->err = B200JPEG_ERR_PARAM;
    emit_symbol_ac(t, t->ac_tbl_no, nb << 4);
    if (nb) emit_bits_e(t, t->EOBRUN, nb);
    t->EOBRUN = 0;
    for (i = 0; i < t->BE; i++) emit_bits_e(t, (unsigned)t->bit_buffer[i], 1);
    t->BE = 0;
  }
}
/* jcphuff.c:468-548 encode_mcu_DC_first (one block) */
static void prog_dc_first(ent_t *t, const int16_t *blk, int ci_in_scan)
{
  const b200jpeg_component_info *c = &t->e->p->comp_info[t->s->ci[ci_in_scan]];
  int Al = t->s->Al, maxbits = t->e->p->data_precision + 2;
  int temp2 = ((int)blk[0]) >> Al, temp = temp2 - t->last_dc[ci_in_scan], temp3, nb;
  t->last_dc[ci_in_scan] = temp2;
  temp3 = temp >> 31; temp ^= temp3; temp -= temp3; temp2 = temp ^ temp3;
  nb = nbits_of(temp);
  if (nb > maxbits + 1) t->e->err = B200JPEG_ERR_BAD_DCT_COEF;
  emit_symbol_dc(t, c->dc_tbl_no, nb);
  if (nb) emit_bits_e(t, (unsigned)temp2, nb);
}
/* jcphuff.c:648-737 encode_mcu_AC_first */
static void prog_ac_first(ent_t *t, const int16_t *blk)
{
  int Ss = t->s->Ss, Se = t->s->Se, Al = t->s->Al, maxbits = t->e->p->data_precision + 2;
  int k, r = 0, any = 0, temp, temp2, nb;
  for (k = Ss; k <= Se; k++) { int a = abs(blk[zz[k]]) >> Al; if (a) { any = 1; break; } }
  if (any && t->EOBRUN > 0) emit_eobrun(t);
  for (k = Ss; k <= Se; k++) {
    temp = blk[zz[k]];
    temp2 = temp >> 31; temp ^= temp2; temp -= temp2; temp >>= Al;
    if (temp == 0) { r++; continue; }
    temp2 ^= temp;
    while (r > 15) { emit_symbol_ac(t, t->ac_tbl_no, 0xF0); r -= 16; }
    nb = nbits_of(temp);
    if (nb > maxbits) t->e->err = B200JPEG_ERR_BAD_DCT_COEF;
    emit_symbol_ac(t, t->ac_tbl_no, (r << 4) + nb);
    emit_bits_e(t, (unsigned)temp2, nb);
    r = 0;
  }
  if (r > 0) { t->EOBRUN++; if (t->EOBRUN == 0x7FFF) emit_eobrun(t); }
}
/* jcphuff.c:746-786 encode_mcu_DC_refine (one block) */
static void prog_dc_refine(ent_t *t, const int16_t *blk) { emit_bits_e(t, (unsigned)(((int)blk[0]) >> t->s->Al), 1); }
/* jcphuff.c:817-1017 encode_mcu_AC_refine */
static void prog_ac_refine(ent_t *t, const int16_t *blk)
{
  int Ss = t->s->Ss, Se = t->s->Se, Al = t->s->Al;
  int absv[64], k, r = 0, EOB = 0, temp; unsigned BR = 0, i;
  char *BR_buffer = t->bit_buffer + t->BE;
  for (k = Ss; k <= Se; k++) { temp = abs(blk[zz[k]]) >> Al; absv[k] = temp; if (temp == 1) EOB = k; }
  for (k = Ss; k <= Se; k++) {
    if ((temp = absv[k]) == 0) { r++; continue; }
    while (r > 15 && k <= EOB) {
      emit_eobrun(t);
      emit_symbol_ac(t, t->ac_tbl_no, 0xF0);
      r -= 16;
      for (i = 0; i < BR; i++) emit_bits_e(t, (unsigned)BR_buffer[i], 1);
      BR_buffer = t->bit_buffer; BR = 0;
    }
    if (temp > 1) { BR_buffer[BR++] = (char)(temp & 1); continue; }
    emit_eobrun(t);
    emit_symbol_ac(t, t->ac_tbl_no, (r << 4) + 1);
    emit_bits_e(t, blk[zz[k]] < 0 ? 0 : 1, 1);
    for (i = 0; i < BR; i++) emit_bits_e(t, (unsigned)BR_buffer[i], 1);
    BR_buffer = t->bit_buffer; BR = 0; r = 0;
  }
  if (r > 0 || BR > 0) {
    t->EOBRUN++; t->BE += BR;
    if (t->EOBRUN == 0x7FFF || t->BE > (1000 - 64 + 1)) emit_eobrun(t);
  }
}

/* ------------------------------------------------------------------ */
/* scan walking: compress_output MCU assembly (jccoefct.c:498-553) with */
/* per_scan_setup geometry (jcmaster.c:518-601)                          */
/* ------------------------------------------------------------------ */
static long scan_num_mcus(const enc_t *e, const scan_t *s)
{
  if (s->ncomps == 1) return (long)e->wib[s->ci[0]] * e->hib[s->ci[0]];
  return (long)e->mcus_per_row * e->mcu_rows;
}
static long scan_mcus_per_row(const enc_t *e, const scan_t *s) { return s->ncomps == 1 ? e->wib[s->ci[0]] : e->mcus_per_row; }

static void run_scan(ent_t *t)
{
  enc_t *e = t->e; const scan_t *s = t->s;
  long n = scan_num_mcus(e, s), per_row = scan_mcus_per_row(e, s), m;
  int k;
  for (k = 0; k < 4; k++) t->last_dc[k] = 0;
  t->EOBRUN = 0; t->BE = 0;
  t->restart_interval = e->p->restart_interval;
  if (e->p->restart_in_rows > 0) { long nominal = (long)e->p->restart_in_rows * per_row; t->restart_interval = (unsigned)(nominal < 65535L ? nominal : 65535L); }
  t->restarts_to_go = t->restart_interval; t->next_restart_num = 0;
  if (s->Ss > 0 || e->progressive) t->ac_tbl_no = e->p->comp_info[s->ci[0]].ac_tbl_no;
  for (m = 0; m < n; m++) {
    long row = m / per_row, col = m % per_row;
    int bi = 0, x, y;
    if (t->restart_interval && t->restarts_to_go == 0) {
      /* jchuff.c:668-686 / jcphuff.c:438-460 emit_restart */
      if (e->progressive) emit_eobrun(t);
      if (!t->gather) { bw_flush(&t->bw); bb_put(t->bw.o, 0xFF); bb_put(t->bw.o, 0xD0 + t->next_restart_num); }
      for (k = 0; k < 4; k++) t->last_dc[k] = 0;
      t->EOBRUN = 0; t->BE = 0;
      t->restarts_to_go = t->restart_interval; t->next_restart_num = (t->next_restart_num + 1) & 7;
    }
    for (k = 0; k < s->ncomps; k++) {
      int ci = s->ci[k];
      int mw = s->ncomps == 1 ? 1 : e->p->comp_info[ci].h_samp_factor;
      int mh = s->ncomps == 1 ? 1 : e->p->comp_info[ci].v_samp_factor;
      for (y = 0; y < mh; y++) for (x = 0; x < mw; x++, bi++) {
        const int16_t *blk = e->coef[ci] + ((size_t)(row * mh + y) * e->wpad[ci] + col * mw + x) * 64;
        if (!e->progressive) seq_block(t, blk, k);
        else if (s->Ah == 0) { if (s->Ss == 0) prog_dc_first(t, blk, k); else prog_ac_first(t, blk); }
        else { if (s->Ss == 0) prog_dc_refine(t, blk); else prog_ac_refine(t, blk); }
      }
    }
    if (t->restart_interval) t->restarts_to_go--;
  }
  if (e->progressive) emit_eobrun(t);
  if (!t->gather) bw_flush(&t->bw);
}

/* finish_pass_gather (jchuff.c:1113-1148) / finish_pass_gather_phuff (jcphuff.c:1050-1085) */
static void tables_from_counts(ent_t *t)
{
  enc_t *e = t->e; const scan_t *s = t->s; int k, did_dc[4] = {0, 0, 0, 0}, did_ac[4] = {0, 0, 0, 0};
  for (k = 0; k < s->ncomps; k++) {
    const b200jpeg_component_info *c = &e->p->comp_info[s->ci[k]];
    int want_dc = !e->progressive || (s->Ss == 0 && s->Ah == 0);
    int want_ac = !e->progressive || (s->Ss != 0);
    if (want_dc && !did_dc[c->dc_tbl_no]) { orc_gen_optimal_table(t->dc_count[c->dc_tbl_no], &e->dc_tbl[c->dc_tbl_no]); e->dc_sent[c->dc_tbl_no] = 0; did_dc[c->dc_tbl_no] = 1; }
    if (want_ac && !did_ac[c->ac_tbl_no]) { orc_gen_optimal_table(t->ac_count[c->ac_tbl_no], &e->ac_tbl[c->ac_tbl_no]); e->ac_sent[c->ac_tbl_no] = 0; did_ac[c->ac_tbl_no] = 1; }
  }
}

static void gather_scan(enc_t *e, const scan_t *s, int trellis_passes, ent_t *t)
{
  int k, i, j;
  memset(t, 0, sizeof *t);
  t->e = e; t->s = s; t->gather = 1;
  if (e->progressive && trellis_passes && s->Ss != 0) {
    /* jcphuff.c:257-264: make sure every codeword the trellis may price has a length */
    for (k = 0; k < s->ncomps; k++) { int tb = e->p->comp_info[s->ci[k]].ac_tbl_no; for (i = 0; i < 16; i++) for (j = 0; j < 12; j++) t->ac_count[tb][16 * i + j] = 1; }
  }
  run_scan(t);
  tables_from_counts(t);
}

/* ------------------------------------------------------------------ */
/* markers (jcmarker.c)                                                  */
/* ------------------------------------------------------------------ */
static void emit_dht_one(enc_t *e, int idx, int is_ac)               /* jcmarker.c:256-291 */
{
  b200jpeg_huff_tbl *h = is_ac ? &e->ac_tbl[idx] : &e->dc_tbl[idx]; int *sent = is_ac ? &e->ac_sent[idx] : &e->dc_sent[idx];
  int len = 0, i;
  if (*sent) return;
  bb_put2(&e->out, 0xFFC4);
  for (i = 1; i <= 16; i++) len += h->bits[i];
  bb_put2(&e->out, len + 2 + 1 + 16);
  bb_put(&e->out, idx + (is_ac ? 0x10 : 0));
  for (i = 1; i <= 16; i++) bb_put(&e->out, h->bits[i]);
  for (i = 0; i < len; i++) bb_put(&e->out, h->huffval[i]);
  *sent = 1;
}
static int emit_multi_dht(enc_t *e, const scan_t *s)                  /* jcmarker.c:293-401 */
{
  int i, j, length = 2, dclens[4] = {0, 0, 0, 0}, aclens[4] = {0, 0, 0, 0};
  int dcseen[4] = {-1, -1, -1, -1}, acseen[4] = {-1, -1, -1, -1};
  if (e->p->compress_profile == B200JPEG_PROFILE_FASTEST) return 0;
  for (i = 0; i < s->ncomps; i++) {
    const b200jpeg_component_info *c = &e->p->comp_info[s->ci[i]];
    int dcidx = c->dc_tbl_no, acidx = c->ac_tbl_no, seen = 0;
    if (s->Ss == 0 && s->Ah == 0) {
      if (e->dc_sent[dcidx]) continue;
      for (j = 0; j < 4; j++) seen += (dcseen[j] == dcidx);
      if (seen) continue;
      dcseen[i] = dcidx;
      for (j = 1; j <= 16; j++) dclens[i] += e->dc_tbl[dcidx].bits[j];
      length += dclens[i] + 16 + 1;
    }
    if (s->Se) {
      if (e->ac_sent[acidx]) continue;
      seen = 0;
      for (j = 0; j < 4; j++) seen += (acseen[j] == acidx);
      if (seen) continue;
      acseen[i] = acidx;
      for (j = 1; j <= 16; j++) aclens[i] += e->ac_tbl[acidx].bits[j];
      length += aclens[i] + 16 + 1;
    }
  }
  if (length > (1 << 16) - 1) return 0;
  bb_put2(&e->out, 0xFFC4); bb_put2(&e->out, length);
  for (i = 0; i < s->ncomps; i++) {
    const b200jpeg_component_info *c = &e->p->comp_info[s->ci[i]];
    int dcidx = c->dc_tbl_no, acidx = c->ac_tbl_no;
    if (s->Ss == 0 && s->Ah == 0 && !e->dc_sent[dcidx]) {
      bb_put(&e->out, dcidx);
      for (j = 1; j <= 16; j++) bb_put(&e->out, e->dc_tbl[dcidx].bits[j]);
      for (j = 0; j < dclens[i]; j++) bb_put(&e->out, e->dc_tbl[dcidx].huffval[j]);
      e->dc_sent[dcidx] = 1;
    }
    if (s->Se && !e->ac_sent[acidx]) {
      bb_put(&e->out, acidx + 0x10);
      for (j = 1; j <= 16; j++) bb_put(&e->out, e->ac_tbl[acidx].bits[j]);
      for (j = 0; j < aclens[i]; j++) bb_put(&e->out, e->ac_tbl[acidx].huffval[j]);
      e->ac_sent[acidx] = 1;
    }
  }
  return 1;
}
static void write_file_header(enc_t *e)                              /* jcmarker.c:649-663, 529-561 */
{
  const b200jpeg_params *p = e->p;
  bb_put2(&e->out, 0xFFD8);
  e->last_restart_interval = 0;
  if (p->write_JFIF_header) {
    bb_put2(&e->out, 0xFFE0); bb_put2(&e->out, 16);
    bb_put(&e->out, 'J'); bb_put(&e->out, 'F'); bb_put(&e->out, 'I'); bb_put(&e->out, 'F'); bb_put(&e->out, 0);
    bb_put(&e->out, p->JFIF_major_version); bb_put(&e->out, p->JFIF_minor_version);
    bb_put(&e->out, p->density_unit); bb_put2(&e->out, p->X_density); bb_put2(&e->out, p->Y_density);
    bb_put(&e->out, 0); bb_put(&e->out, 0);
  }
  if (p->write_Adobe_marker) {                                        /* jcmarker.c:564-596 */
    bb_put2(&e->out, 0xFFEE); bb_put2(&e->out, 14);
    bb_put(&e->out, 'A'); bb_put(&e->out, 'd'); bb_put(&e->out, 'o'); bb_put(&e->out, 'b'); bb_put(&e->out, 'e');
    bb_put2(&e->out, 100); bb_put2(&e->out, 0); bb_put2(&e->out, 0);
    bb_put(&e->out, p->jpeg_color_space == B200JPEG_CS_YCbCr ? 1 : 0);
  }
}
static void write_frame_header(enc_t *e)                             /* jcmarker.c:674-735, 140-254, 464-491 */
{
  const b200jpeg_params *p = e->p; int ci, i, prec = 0, multi = 1, is_baseline;
  int precs[4] = {0, 0, 0, 0};
  if (p->compress_profile == B200JPEG_PROFILE_FASTEST) multi = 0;
  if (multi) for (ci = 0; ci < e->nc; ci++) { int t = p->comp_info[ci].quant_tbl_no; if (e->qt_sent[t]) multi = 0; }
  if (multi) {                                                         /* emit_multi_dqt */
    int seen[4] = {0, 0, 0, 0}, size = 0;
    for (ci = 0; ci < e->nc; ci++) { int t = p->comp_info[ci].quant_tbl_no; for (i = 0; i < 64; i++) precs[ci] = !!(precs[ci] + (p->quant_tbl[t][i] > 255)); prec += precs[ci]; }
    bb_put2(&e->out, 0xFFDB);
    for (ci = 0; ci < e->nc; ci++) { int t = p->comp_info[ci].quant_tbl_no; if (!seen[t]) { size += 64 * (precs[ci] + 1) + 1; seen[t] = 1; } }
    bb_put2(&e->out, size + 2);
    for (ci = 0; ci < e->nc; ci++) {
      int t = p->comp_info[ci].quant_tbl_no;
      if (e->qt_sent[t]) continue;
      bb_put(&e->out, t + (precs[ci] << 4));
      for (i = 0; i < 64; i++) { unsigned q = p->quant_tbl[t][zz[i]]; if (precs[ci]) bb_put(&e->out, q >> 8); bb_put(&e->out, q & 0xFF); }
      e->qt_sent[t] = 1;
    }
  } else {
    for (ci = 0; ci < e->nc; ci++) {                                   /* emit_dqt per component */
      int t = p->comp_info[ci].quant_tbl_no, pr = 0;
      for (i = 0; i < 64; i++) if (p->quant_tbl[t][i] > 255) pr = 1;
      if (!e->qt_sent[t]) {
        bb_put2(&e->out, 0xFFDB); bb_put2(&e->out, pr ? 64 * 2 + 1 + 2 : 64 + 1 + 2);
        bb_put(&e->out, t + (pr << 4));
        for (i = 0; i < 64; i++) { unsigned q = p->quant_tbl[t][zz[i]]; if (pr) bb_put(&e->out, q >> 8); bb_put(&e->out, q & 0xFF); }
        e->qt_sent[t] = 1;
      }
      prec += pr;
    }
  }
  if (e->progressive || p->data_precision != 8) is_baseline = 0;
  else {
    is_baseline = 1;
    for (ci = 0; ci < e->nc; ci++) if (p->comp_info[ci].dc_tbl_no > 1 || p->comp_info[ci].ac_tbl_no > 1) is_baseline = 0;
    if (prec && is_baseline) is_baseline = 0;
  }
  bb_put2(&e->out, e->progressive ? 0xFFC2 : (is_baseline ? 0xFFC0 : 0xFFC1));
  bb_put2(&e->out, 3 * e->nc + 2 + 5 + 1);
  bb_put(&e->out, p->data_precision); bb_put2(&e->out, e->H); bb_put2(&e->out, e->W);
  bb_put(&e->out, e->nc);
  for (ci = 0; ci < e->nc; ci++) {
    bb_put(&e->out, p->comp_info[ci].component_id);
    bb_put(&e->out, (p->comp_info[ci].h_samp_factor << 4) + p->comp_info[ci].v_samp_factor);
    bb_put(&e->out, p->comp_info[ci].quant_tbl_no);
  }
}
static void write_scan_header(enc_t *e, const scan_t *s, unsigned restart_interval)   /* jcmarker.c:744-784, 494-526 */
{
  int i;
  if (!emit_multi_dht(e, s)) {
    for (i = 0; i < s->ncomps; i++) {
      const b200jpeg_component_info *c = &e->p->comp_info[s->ci[i]];
      if (s->Ss == 0 && s->Ah == 0) emit_dht_one(e, c->dc_tbl_no, 0);
      if (s->Se) emit_dht_one(e, c->ac_tbl_no, 1);
    }
  }
  if ((int)restart_interval != e->last_restart_interval) {
    bb_put2(&e->out, 0xFFDD); bb_put2(&e->out, 4); bb_put2(&e->out, (int)restart_interval);
    e->last_restart_interval = (int)restart_interval;
  }
  bb_put2(&e->out, 0xFFDA); bb_put2(&e->out, 2 * s->ncomps + 2 + 1 + 3);
  bb_put(&e->out, s->ncomps);
  for (i = 0; i < s->ncomps; i++) {
    const b200jpeg_component_info *c = &e->p->comp_info[s->ci[i]];
    int td = (s->Ss == 0 && s->Ah == 0) ? c->dc_tbl_no : 0, ta = s->Se ? c->ac_tbl_no : 0;
    bb_put(&e->out, c->component_id); bb_put(&e->out, (td << 4) + ta);
  }
  bb_put(&e->out, s->Ss); bb_put(&e->out, s->Se); bb_put(&e->out, (s->Ah << 4) + s->Al);
}

/* ------------------------------------------------------------------ */
/* sample planes -> coefficient planes                                  */
/* ------------------------------------------------------------------ */

/* Dummy-block rules (jccoefct.c:312-345 == :443-476): right-edge dummies copy
 * the DC of the block to their left; bottom dummy rows copy, per MCU, the DC
 * of the last block of that MCU in the row above; all AC = 0. */
static void fill_dummy_blocks(enc_t *e, int ci)
{
  int h = e->p->comp_info[ci].h_samp_factor, wib = e->wib[ci], hib = e->hib[ci], wpad = e->wpad[ci], hpad = e->hpad[ci];
  int r, b, m;
  int16_t *c = e->coef[ci];
  for (r = 0; r < hib; r++)
    for (b = wib; b < wpad; b++) { int16_t *blk = c + ((size_t)r * wpad + b) * 64; memset(blk, 0, 128); blk[0] = blk[-64]; }
  for (r = hib; r < hpad; r++)
    for (m = 0; m < wpad / h; m++) {
      int16_t last = c[((size_t)(r - 1) * wpad + m * h + h - 1) * 64];
      for (b = 0; b < h; b++) { int16_t *blk = c + ((size_t)r * wpad + m * h + b) * 64; memset(blk, 0, 128); blk[0] = last; }
    }
}

/* samples are uint8 (8-bit) or uint16 holding 12-bit values (J12SAMPLE, jmorecfg.h) */
static int forward_all(enc_t *e, const uint8_t *pix, size_t pitch)
{
  const uint8_t *const *rawp = e->raw_planes; const size_t *rawpitch = e->raw_pitch;     /* jpeg_write_raw_data input, or NULL */
  const b200jpeg_params *p = e->p;
  const int prec = p->data_precision, centre = 1 << (prec - 1);
  int W = e->W, H = e->H, ci, x, y;
  /* full-resolution converted planes (jccolor.c) */
  uint16_t *full[4] = {0, 0, 0, 0};
  for (ci = 0; ci < e->nc; ci++) full[ci] = (uint16_t *)malloc((size_t)W * H * 2);
  if (rawp) goto planes_ready;                       /* raw data: no colour conversion, no downsampling, no edge expansion */
#define IN(xx, cc) (prec == 8 ? (int)row[p->input_components * (xx) + (cc)] : (int)((const uint16_t *)row)[p->input_components * (xx) + (cc)])
  for (y = 0; y < H; y++) {
    const uint8_t *row = pix + (size_t)y * pitch;
    for (x = 0; x < W; x++) {
      if (p->in_color_space == B200JPEG_CS_RGB && p->jpeg_color_space == B200JPEG_CS_YCbCr) {
        int Y, Cb, Cr; rgb_to_ycc_p(IN(x, 0), IN(x, 1), IN(x, 2), prec, &Y, &Cb, &Cr);
        full[0][(size_t)y * W + x] = Y; full[1][(size_t)y * W + x] = Cb; full[2][(size_t)y * W + x] = Cr;
      } else if (p->in_color_space == B200JPEG_CS_RGB && p->jpeg_color_space == B200JPEG_CS_GRAYSCALE) {
        int Y, Cb, Cr; rgb_to_ycc_p(IN(x, 0), IN(x, 1), IN(x, 2), prec, &Y, &Cb, &Cr);   /* jccolor.c rgb_gray_convert: same Y */
        full[0][(size_t)y * W + x] = Y;
      } else {                                                                                     /* null_convert / grayscale_convert */
        for (ci = 0; ci < e->nc; ci++) full[ci][(size_t)y * W + x] = (uint16_t)IN(x, ci);
      }
    }
  }
#undef IN
planes_ready:
  for (ci = 0; ci < e->nc; ci++) {
    const b200jpeg_component_info *c = &p->comp_info[ci];
    int hx = e->hmax / c->h_samp_factor, vx = e->vmax / c->v_samp_factor;
    int ow = e->wib[ci] * 8, oh = e->hib[ci] * 8;
    int groups = (H + e->vmax - 1) / e->vmax;                /* row groups holding real data (jcprepct.c:135-192) */
    int rows_avail = groups * c->v_samp_factor;
    uint16_t *plane = (uint16_t *)malloc((size_t)ow * oh * 2);
    int numpix = hx * vx, bx, by, i;
    if (rawp) {                                       /* compress_first_pass reads hib*8 rows of wib*8 samples straight from the caller's planes */
      for (y = 0; y < oh; y++) for (x = 0; x < ow; x++) plane[(size_t)y * ow + x] = rawp[ci][(size_t)y * rawpitch[ci] + x];
    } else if (p->smoothing_factor) {
      /* Input smoothing (jcsample.c:298-455).  Some component always takes a smoothing method (the full-size one), so
       * the downsampler asks for context rows (jcsample.c:482,517) and the pre-processing controller runs
       * pre_process_context (jcprepct.c:201-262): the first row is replicated upwards (:226-234), rows past the bottom
       * are replicas of the last INPUT row (:246-252) and every output row - padding included - is downsampled from
       * them.  expand_right_edge + the first/last-column special cases equal clamping the column to [0, W-1]. */
      const long SF = p->smoothing_factor;
#define AT(yy_, xx_) ((long)full[ci][(size_t)((yy_) < 0 ? 0 : (yy_) > H - 1 ? H - 1 : (yy_)) * W + ((xx_) < 0 ? 0 : (xx_) > W - 1 ? W - 1 : (xx_))])
      for (y = 0; y < oh; y++) for (x = 0; x < ow; x++) {
        long val;
        if (hx == 1 && vx == 1) {                                            /* fullsize_smooth_downsample :405-455 */
          long member = AT(y, x), neigh = -member; int dy, dx;
          for (dy = -1; dy <= 1; dy++) for (dx = -1; dx <= 1; dx++) neigh += AT(y + dy, x + dx);
          val = (member * (65536L - SF * 512L) + neigh * (SF * 64) + 32768) >> 16;
        } else if (hx == 2 && vx == 2) {                                     /* h2v2_smooth_downsample :306-397 */
          int Y = 2 * y, X = 2 * x;
          long member = AT(Y, X) + AT(Y, X + 1) + AT(Y + 1, X) + AT(Y + 1, X + 1);
          long edge = AT(Y - 1, X) + AT(Y - 1, X + 1) + AT(Y + 2, X) + AT(Y + 2, X + 1) + AT(Y, X - 1) + AT(Y, X + 2) + AT(Y + 1, X - 1) + AT(Y + 1, X + 2);
          long corner = AT(Y - 1, X - 1) + AT(Y - 1, X + 2) + AT(Y + 2, X - 1) + AT(Y + 2, X + 2);
          val = (member * (16384 - SF * 80) + (2 * edge + corner) * (SF * 16) + 32768) >> 16;
        } else {                                                             /* no smoothing variant: h2v1 / int_downsample */
          long sum = 0; int u, v;
          for (v = 0; v < vx; v++) for (u = 0; u < hx; u++) sum += AT(y * vx + v, x * hx + u);
          if (hx == 2 && vx == 1) val = (sum + (x & 1)) >> 1; else val = (sum + numpix / 2) / numpix;
        }
        plane[(size_t)y * ow + x] = (uint16_t)val;
      }
#undef AT
    } else
    for (y = 0; y < oh; y++) {
      int yy = y < rows_avail ? y : rows_avail - 1;          /* expand_bottom_edge on the downsampled rows */
      int g = yy / c->v_samp_factor, s = yy % c->v_samp_factor;
      for (x = 0; x < ow; x++) {
        int sum = 0, u, v, val;
        for (v = 0; v < vx; v++) {
          int iy = g * e->vmax + s * vx + v; if (iy > H - 1) iy = H - 1;
          for (u = 0; u < hx; u++) { int ix = x * hx + u; if (ix > W - 1) ix = W - 1; sum += full[ci][(size_t)iy * W + ix]; }
        }
        if (hx == 1 && vx == 1) val = sum;                                   /* jcsample.c:199-211 */
        else if (hx == 2 && vx == 1) val = (sum + (x & 1)) >> 1;             /* jcsample.c:226-254 bias 0,1,.. */
        else if (hx == 2 && vx == 2) val = (sum + 1 + (x & 1)) >> 2;         /* jcsample.c:263-295 bias 1,2,.. */
        else val = (sum + numpix / 2) / numpix;                              /* jcsample.c:151-190 */
        plane[(size_t)y * ow + x] = (uint16_t)val;
      }
    }
    /* forward_DCT on every real block (jcdctmgr.c:693-772) */
    for (by = 0; by < e->hib[ci]; by++) for (bx = 0; bx < e->wib[ci]; bx++) {
      int ws[64]; const uint16_t *q = p->quant_tbl[c->quant_tbl_no];
      int16_t *dq = e->coef[ci] + ((size_t)by * e->wpad[ci] + bx) * 64, *dr = e->raw[ci] + ((size_t)by * e->wpad[ci] + bx) * 64;
      for (y = 0; y < 8; y++) for (x = 0; x < 8; x++) ws[8 * y + x] = plane[(size_t)(by * 8 + y) * ow + bx * 8 + x] - centre;   /* convsamp :576-604 */
      if (p->dct_method == B200JPEG_DCT_FLOAT) { forward_block_float(p, ws, q, dq, dr); continue; }
      if (p->dct_method == B200JPEG_DCT_IFAST) { forward_block_ifast(p, ws, q, dq, dr); continue; }
      if (p->overshoot_deringing) orc_deringing(ws, q[0]);
      fdct_islow_prec(ws, prec);
      for (i = 0; i < 64; i++) {
        int v = orc_quantize_coef(ws[i], q[i]);                /* 12-bit: the literal division of :646-678 is the same function */
        dr[i] = (int16_t)ws[i];                                /* (12-bit raw values can exceed int16; only the trellis reads them and it is off) */
        if (p->overshoot_deringing) { int mx = (1 << (p->data_precision + 2)) - 1; if (v < -mx) v = -mx; if (v > mx) v = mx; }   /* :761-770 */
        dq[i] = (int16_t)v;
      }
    }
    free(plane);
    fill_dummy_blocks(e, ci);
  }
  for (ci = 0; ci < e->nc; ci++) free(full[ci]);
  return 0;
}

/* compress_trellis_pass over the whole image for one component (jccoefct.c:356-486) */
static void trellis_component(enc_t *e, int ci, int Ss, int Se, double (*norm_src)[64], double (*norm_coef)[64])
{
  const b200jpeg_params *p = e->p; const b200jpeg_component_info *c = &p->comp_info[ci];
  unsigned co[256]; unsigned char dcsi[256], acsi[256];
  int v = c->v_samp_factor, imcu, br;
  orc_make_derived(&e->dc_tbl[c->dc_tbl_no], 1, co, dcsi);
  orc_make_derived(&e->ac_tbl[c->ac_tbl_no], 0, co, acsi);
  for (imcu = 0; imcu * v < e->hib[ci]; imcu++) {
    int16_t lastDC = 0;
    for (br = 0; br < v && imcu * v + br < e->hib[ci]; br++) {
      size_t off = (size_t)(imcu * v + br) * e->wpad[ci] * 64;
      size_t up = off - (size_t)e->wpad[ci] * 64;                 /* lastblockrow = buffer[block_row-1] only inside the iMCU row (jccoefct.c:420) */
      orc_trellis_row(p, dcsi, acsi, e->coef[ci] + off, e->raw[ci] + off, e->wib[ci], p->quant_tbl[c->quant_tbl_no], &lastDC,
                      br > 0 ? e->coef[ci] + up : NULL, br > 0 ? e->raw[ci] + up : NULL, Ss, Se,
                      norm_src ? norm_src[c->quant_tbl_no] : NULL, norm_coef ? norm_coef[c->quant_tbl_no] : NULL);
    }
  }
  fill_dummy_blocks(e, ci);
}

/* ------------------------------------------------------------------ */
void orc_free(void *p) { free(p); }
void orc_debug_free(orc_debug *d) { int i; for (i = 0; i < 4; i++) { free(d->plain[i]); free(d->raw[i]); free(d->final_[i]); } memset(d, 0, sizeof *d); }

static const uint8_t *const *g_raw_planes; static const size_t *g_raw_pitch;   /* set by orc_encode_raw around its call of orc_encode (test infrastructure: single-threaded) */
static const int16_t *const *g_coef_planes; static const size_t *g_coef_pitch;  /* likewise, set by orc_encode_coefs */
int orc_encode(const b200jpeg_params *p, const uint8_t *pixels, size_t row_pitch,
               uint8_t **out, size_t *outsize, orc_debug *dbg);
/* jpeg_write_coefficients (jctrans.c:39-66): quantized coefficients in, entropy coding only.  planes[ci] = hib rows of
 * wib JBLOCKs (natural order), pitch_blocks[ci] blocks per row.  Dummy blocks as compress_output makes them
 * (jctrans.c:352-362: AC 0, DC of the previous block of the MCU) - the same values fill_dummy_blocks produces. */
int orc_encode_coefs(const b200jpeg_params *p, const int16_t *const *planes, const size_t *pitch_blocks, uint8_t **out, size_t *outsize)
{
  int rc;
  if (p->trellis_quant) return B200JPEG_ERR_PARAM;        /* jpeg_copy_critical_parameters turns it off (jctrans.c:103) */
  g_coef_planes = planes; g_coef_pitch = pitch_blocks;
  rc = orc_encode(p, (const uint8_t *)planes[0], 0, out, outsize, NULL);
  g_coef_planes = NULL; g_coef_pitch = NULL;
  return rc;
}
/* jpeg_write_raw_data (jcapistd.c:145-195): component planes instead of pixels */
int orc_encode_raw(const b200jpeg_params *p, const uint8_t *const *planes, const size_t *pitch, uint8_t **out, size_t *outsize)
{
  int rc;
  if (p->data_precision != 8) return B200JPEG_ERR_UNSUPPORTED;
  g_raw_planes = planes; g_raw_pitch = pitch;
  rc = orc_encode(p, planes[0], pitch[0], out, outsize, NULL);
  g_raw_planes = NULL; g_raw_pitch = NULL;
  return rc;
}
int orc_encode(const b200jpeg_params *p, const uint8_t *pixels, size_t row_pitch,
               uint8_t **out, size_t *outsize, orc_debug *dbg)
{
  enc_t E; enc_t *e = &E; int ci, si, nscans; scan_t scans[B200JPEG_MAX_SCANS];
  ent_t *t = (ent_t *)malloc(sizeof(ent_t));
  memset(e, 0, sizeof *e);
  *out = NULL; *outsize = 0;
  if (dbg) memset(dbg, 0, sizeof *dbg);
  static b200jpeg_params P;                                /* mutable copy: trellis_q_opt rewrites the quantization tables (test infrastructure: single-threaded) */
  P = *p; p = &P;
  e->p = p; e->nc = p->num_components; e->W = p->image_width; e->H = p->image_height;
  e->raw_planes = g_raw_planes; e->raw_pitch = g_raw_pitch;
  /* 12-bit: no JBUF_REQUANT => no trellis (jccoefct.c:132-138); deringing is not usable at 12 bits (jcdctmgr.c:419) */
  if (p->data_precision == 12 && !g_coef_planes && (p->trellis_quant || p->overshoot_deringing)) { free(t); return B200JPEG_ERR_UNSUPPORTED; }   /* (no forward stage when transcoding) */
  if ((p->data_precision != 8 && p->data_precision != 12) || p->trellis_num_loops < 1) { free(t); return B200JPEG_ERR_UNSUPPORTED; }
  e->hmax = e->vmax = 1;
  for (ci = 0; ci < e->nc; ci++) { if (p->comp_info[ci].h_samp_factor > e->hmax) e->hmax = p->comp_info[ci].h_samp_factor; if (p->comp_info[ci].v_samp_factor > e->vmax) e->vmax = p->comp_info[ci].v_samp_factor; }
  e->mcus_per_row = (e->W + e->hmax * 8 - 1) / (e->hmax * 8); e->mcu_rows = (e->H + e->vmax * 8 - 1) / (e->vmax * 8);
  for (ci = 0; ci < e->nc; ci++) {                                       /* jcmaster.c:215-236 */
    int h = p->comp_info[ci].h_samp_factor, v = p->comp_info[ci].v_samp_factor;
    if (e->hmax % h || e->vmax % v) { free(t); return B200JPEG_ERR_UNSUPPORTED; }
    e->wib[ci] = (e->W * h + e->hmax * 8 - 1) / (e->hmax * 8); e->hib[ci] = (e->H * v + e->vmax * 8 - 1) / (e->vmax * 8);
    e->wpad[ci] = e->mcus_per_row * h; e->hpad[ci] = e->mcu_rows * v;
    e->coef[ci] = (int16_t *)calloc((size_t)e->wpad[ci] * e->hpad[ci] * 64, 2);
    e->raw[ci] = (int16_t *)calloc((size_t)e->wpad[ci] * e->hpad[ci] * 64, 2);
  }
  for (ci = 0; ci < 4; ci++) { e->dc_tbl[ci] = p->dc_huff_tbl[ci]; e->ac_tbl[ci] = p->ac_huff_tbl[ci]; }

  /* scan list + progressive_mode (validate_script, jcmaster.c:252-436) */
  if (p->num_scans > 0) {
    nscans = p->num_scans;
    for (si = 0; si < nscans; si++) { int k; scans[si].ncomps = p->scan_info[si].comps_in_scan; for (k = 0; k < 4; k++) scans[si].ci[k] = p->scan_info[si].component_index[k];
      scans[si].Ss = p->scan_info[si].Ss; scans[si].Se = p->scan_info[si].Se; scans[si].Ah = p->scan_info[si].Ah; scans[si].Al = p->scan_info[si].Al; }
    e->progressive = (scans[0].Ss != 0 || scans[0].Se != 63);
  } else {
    nscans = 1; scans[0].ncomps = e->nc; for (ci = 0; ci < 4; ci++) scans[0].ci[ci] = ci; scans[0].Ss = 0; scans[0].Se = 63; scans[0].Ah = scans[0].Al = 0;
    e->progressive = 0;
  }
  {
    int optimize = p->optimize_coding || e->progressive || p->data_precision == 12;   /* jcmaster.c:1091-1094, :1102-1105 */
    if (p->trellis_quant && !optimize) { free(t); return B200JPEG_ERR_UNSUPPORTED; }

    write_file_header(e);                                                /* jcinit.c:149 */
    if (g_coef_planes) {                                                 /* transcoding: no pass 0, the arrays are given */
      for (ci = 0; ci < e->nc; ci++) {
        int by;
        for (by = 0; by < e->hib[ci]; by++)
          memcpy(e->coef[ci] + (size_t)by * e->wpad[ci] * 64, g_coef_planes[ci] + (size_t)by * g_coef_pitch[ci] * 64, (size_t)e->wib[ci] * 128);
        fill_dummy_blocks(e, ci);
      }
    } else
    forward_all(e, pixels, row_pitch);                                   /* pass 0 data path */
    if (dbg) { dbg->ncomp = e->nc; for (ci = 0; ci < e->nc; ci++) { size_t n = (size_t)e->wpad[ci] * e->hpad[ci] * 64 * 2;
        dbg->wib[ci] = e->wib[ci]; dbg->hib[ci] = e->hib[ci]; dbg->wpad[ci] = e->wpad[ci]; dbg->hpad[ci] = e->hpad[ci];
        dbg->plain[ci] = (int16_t *)malloc(n); memcpy(dbg->plain[ci], e->coef[ci], n); dbg->raw[ci] = (int16_t *)malloc(n); memcpy(dbg->raw[ci], e->raw[ci], n); } }

    if (p->trellis_quant) {
      /* trellis phase (jcmaster.c:443-467,612-715,968-1035): per component,
       * gather on the plain-quantized coefs (Ss=1..63 for jcphuff), build
       * tables, requantize.  The re-gather the reference does inside the
       * trellis pass only produces tables that are overwritten before use. */
      double norm_src[4][64], norm_coef[4][64]; int pass_number = 0;
      memset(norm_src, 0, sizeof norm_src); memset(norm_coef, 0, sizeof norm_coef);
      for (ci = 0; ci < e->nc; ci++) {
        /* use_scans_in_trellis: the component's AC band is split at trellis_freq_split and each half gets its own
         * statistics -> tables -> quantize_trellis pair of passes (select_scan_parameters, jcmaster.c:451-467) */
        const int nband = p->use_scans_in_trellis ? 2 : 1; int band, loop;
        /* trellis_num_loops: the component's rounds are simply repeated (pass -> component = pass / (2 or 4 * loops),
         * jcmaster.c:453-465): later rounds gather on the already requantized coefficients */
        for (loop = 0; loop < p->trellis_num_loops; loop++)
        for (band = 0; band < nband; band++) {
          const int group = e->nc * (nband == 2 ? 4 : 2);     /* passes between two table updates (jcmaster.c:687-698,1014-1030) */
          scan_t ts; ts.ncomps = 1; ts.ci[0] = ci; ts.Ah = ts.Al = 0;
          ts.Ss = (nband == 2 && band == 1) ? p->trellis_freq_split + 1 : 1;
          ts.Se = (nband == 2 && band == 0) ? p->trellis_freq_split : 63;
          {
            scan_t gs = ts;
            if (!e->progressive) gs.Ss = 0;   /* jchuff ignores Ss/Se: DC+AC statistics over the whole block */
            gather_scan(e, &gs, 1, t);
          }
          if (dbg) { dbg->trellis_dc[ci] = e->dc_tbl[p->comp_info[ci].dc_tbl_no]; dbg->trellis_ac[ci] = e->ac_tbl[p->comp_info[ci].ac_tbl_no]; }
          pass_number++;                                       /* the statistics pass */
          if (p->trellis_q_opt && pass_number % group == 1) memset(norm_src, 0, sizeof norm_src), memset(norm_coef, 0, sizeof norm_coef);
          trellis_component(e, ci, ts.Ss, ts.Se, norm_src, norm_coef);
          if (p->trellis_q_opt && (pass_number + 1) % group == 0) {
            /* trellis_q_opt: every table entry becomes the least-squares dequantizer of what the trellis kept */
            int ti, jj;
            for (ti = 0; ti < 4; ti++) for (jj = 1; jj < 64; jj++) if (norm_coef[ti][jj] != 0.0) {
              int q = (int)(norm_src[ti][jj] / norm_coef[ti][jj] + 0.5);
              if (q > 254) q = 254;
              if (q < 1) q = 1;
              P.quant_tbl[ti][jj] = (uint16_t)q;
            }
          }
          pass_number++;                                       /* the trellis pass */
        }
      }
    }
    if (dbg) for (ci = 0; ci < e->nc; ci++) { size_t n = (size_t)e->wpad[ci] * e->hpad[ci] * 64 * 2; dbg->final_[ci] = (int16_t *)malloc(n); memcpy(dbg->final_[ci], e->coef[ci], n); }

    if (p->optimize_scans && p->num_scans > 0) {
      /* scan search: finish_pass_master / select_scans / select_scan_parameters (jcmaster.c:443-515, 773-962, 968-1035)
       * followed literally: every candidate scan the reference codes is coded into its own buffer (frame header
       * included for scan 0, scan header for all: the memory destination is installed before write_scan_header,
       * :668-681), select_scans() runs after each one and may jump ahead, and at the end the chosen buffers are
       * copied out in the order of :911-957. */
      const int num_scans_luma_dc = 1, Al_max_luma = 3, nsplits = 5, num_scans_chroma_dc = 3, Al_max_chroma = 2;
      const int num_scans_luma = num_scans_luma_dc + (3 * Al_max_luma + 2) + (2 * nsplits + 1);
      const int luma_freq_split_scan_start = num_scans_luma_dc + 3 * Al_max_luma + 2;
      const int chroma_freq_split_scan_start = num_scans_luma + num_scans_chroma_dc + (6 * Al_max_chroma + 4);
      bytebuf *sb = (bytebuf *)calloc((size_t)nscans, sizeof(bytebuf));
      unsigned long best_cost = 0;
      int best_Al_luma = 0, best_Al_chroma = 0, best_freq_split_idx_luma = 0, best_freq_split_idx_chroma = 0, interleave_chroma_dc = 0;
      int scan_number = 0, done = 0, i;
      bytebuf real_out = e->out;
#define SZ(k) ((unsigned long)sb[k].n)
#define COPY(k) do { size_t q_; for (q_ = 0; q_ < sb[k].n; q_++) bb_put(&e->out, sb[k].d[q_]); } while (0)
      if (nscans != ((e->nc == 3) ? 64 : 23)) { free(sb); free(t); return B200JPEG_ERR_PARAM; }
      while (!done && !e->err) {
        scan_t sc = scans[scan_number]; int k, next_scan_number, base_scan_idx = 0; unsigned ri;
        /* select_scan_parameters :477-488: the frequency-split scans are coded at the best Al found so far */
        if (scan_number >= luma_freq_split_scan_start && scan_number < num_scans_luma) sc.Al = best_Al_luma;
        if (scan_number >= chroma_freq_split_scan_start && scan_number < nscans) sc.Al = best_Al_chroma;
        if (!(sc.Ss == 0 && sc.Ah != 0)) gather_scan(e, &sc, 0, t);
        memset(&e->out, 0, sizeof e->out);
        if (scan_number == 0) write_frame_header(e);
        ri = p->restart_interval;
        if (p->restart_in_rows > 0) { long nominal = (long)p->restart_in_rows * scan_mcus_per_row(e, &sc); ri = (unsigned)(nominal < 65535L ? nominal : 65535L); }
        write_scan_header(e, &sc, ri);
        memset(t, 0, sizeof *t); t->e = e; t->s = &sc; t->gather = 0; t->bw.o = &e->out;
        for (k = 0; k < 4; k++) { orc_make_derived(&e->dc_tbl[k], 1, t->dco[k], t->dsi[k]); orc_make_derived(&e->ac_tbl[k], 0, t->aco[k], t->asi[k]); }
        run_scan(t);
        sb[scan_number] = e->out;
        /* select_scans(cinfo, scan_number + 1) */
        next_scan_number = scan_number + 1;
        if (next_scan_number > 1 && next_scan_number <= luma_freq_split_scan_start) {
          if ((next_scan_number - 1) % 3 == 2) {
            int Al = (next_scan_number - 1) / 3; unsigned long cost = SZ(next_scan_number - 2) + SZ(next_scan_number - 1);
            for (i = 0; i < Al; i++) cost += SZ(3 + 3 * i);
            if (Al == 0 || cost < best_cost) { best_cost = cost; best_Al_luma = Al; }
            else scan_number = luma_freq_split_scan_start - 1;
          }
        } else if (next_scan_number > luma_freq_split_scan_start && next_scan_number <= num_scans_luma) {
          if (next_scan_number == luma_freq_split_scan_start + 1) { best_freq_split_idx_luma = 0; best_cost = SZ(next_scan_number - 1); }
          else if ((next_scan_number - luma_freq_split_scan_start) % 2 == 1) {
            int idx = (next_scan_number - luma_freq_split_scan_start) >> 1; unsigned long cost = SZ(next_scan_number - 2) + SZ(next_scan_number - 1);
            if (cost < best_cost) { best_cost = cost; best_freq_split_idx_luma = idx; }
            if ((idx == 2 && best_freq_split_idx_luma == 0) || (idx == 3 && best_freq_split_idx_luma != 2) || (idx == 4 && best_freq_split_idx_luma != 4))
              scan_number = num_scans_luma - 1;
          }
        } else if (nscans > num_scans_luma) {
          if (next_scan_number == num_scans_luma + num_scans_chroma_dc) {
            base_scan_idx = num_scans_luma;
            interleave_chroma_dc = SZ(base_scan_idx) <= SZ(base_scan_idx + 1) + SZ(base_scan_idx + 2);
          } else if (next_scan_number > num_scans_luma + num_scans_chroma_dc && next_scan_number <= chroma_freq_split_scan_start) {
            base_scan_idx = num_scans_luma + num_scans_chroma_dc;
            if ((next_scan_number - base_scan_idx) % 6 == 4) {
              int Al = (next_scan_number - base_scan_idx) / 6;
              unsigned long cost = SZ(next_scan_number - 4) + SZ(next_scan_number - 3) + SZ(next_scan_number - 2) + SZ(next_scan_number - 1);
              for (i = 0; i < Al; i++) { cost += SZ(base_scan_idx + 4 + 6 * i); cost += SZ(base_scan_idx + 5 + 6 * i); }
              if (Al == 0 || cost < best_cost) { best_cost = cost; best_Al_chroma = Al; }
              else scan_number = chroma_freq_split_scan_start - 1;
            }
          } else if (next_scan_number > chroma_freq_split_scan_start && next_scan_number <= nscans) {
            if (next_scan_number == chroma_freq_split_scan_start + 2) { best_freq_split_idx_chroma = 0; best_cost = SZ(next_scan_number - 2) + SZ(next_scan_number - 1); }
            else if ((next_scan_number - chroma_freq_split_scan_start) % 4 == 2) {
              int idx = (next_scan_number - chroma_freq_split_scan_start) >> 2;
              unsigned long cost = SZ(next_scan_number - 4) + SZ(next_scan_number - 3) + SZ(next_scan_number - 2) + SZ(next_scan_number - 1);
              if (cost < best_cost) { best_cost = cost; best_freq_split_idx_chroma = idx; }
              if ((idx == 2 && best_freq_split_idx_chroma == 0) || (idx == 3 && best_freq_split_idx_chroma != 2) || (idx == 4 && best_freq_split_idx_chroma != 4))
                scan_number = nscans - 1;
            }
          }
        }
        if (scan_number == nscans - 1) {
          int Al, min_Al = best_Al_luma < best_Al_chroma ? best_Al_luma : best_Al_chroma;
          e->out = real_out;
          COPY(0);
          if (nscans > num_scans_luma && p->dc_scan_opt_mode != 0) {
            base_scan_idx = num_scans_luma;
            if (interleave_chroma_dc && p->dc_scan_opt_mode != 1) COPY(base_scan_idx);
            else { COPY(base_scan_idx + 1); COPY(base_scan_idx + 2); }
          }
          if (best_freq_split_idx_luma == 0) COPY(luma_freq_split_scan_start);
          else { COPY(luma_freq_split_scan_start + 2 * (best_freq_split_idx_luma - 1) + 1); COPY(luma_freq_split_scan_start + 2 * (best_freq_split_idx_luma - 1) + 2); }
          for (Al = best_Al_luma - 1; Al >= min_Al; Al--) COPY(3 + 3 * Al);
          if (nscans > num_scans_luma) {
            if (best_freq_split_idx_chroma == 0) { COPY(chroma_freq_split_scan_start); COPY(chroma_freq_split_scan_start + 1); }
            else { int q2; for (q2 = 2; q2 <= 5; q2++) COPY(chroma_freq_split_scan_start + 4 * (best_freq_split_idx_chroma - 1) + q2); }
            base_scan_idx = num_scans_luma + num_scans_chroma_dc;
            for (Al = best_Al_chroma - 1; Al >= min_Al; Al--) { COPY(base_scan_idx + 6 * Al + 4); COPY(base_scan_idx + 6 * Al + 5); }
          }
          for (Al = min_Al - 1; Al >= 0; Al--) {
            COPY(3 + 3 * Al);
            if (nscans > num_scans_luma) { COPY(base_scan_idx + 6 * Al + 4); COPY(base_scan_idx + 6 * Al + 5); }
          }
          real_out = e->out;
          done = 1;
        }
        scan_number++;
      }
      e->out = real_out;
      for (i = 0; i < nscans; i++) free(sb[i].d);
      free(sb);
#undef SZ
#undef COPY
    } else
    for (si = 0; si < nscans && !e->err; si++) {
      const scan_t *s = &scans[si]; size_t before; int k;
      unsigned ri;
      if (optimize && !(e->progressive && s->Ss == 0 && s->Ah != 0)) gather_scan(e, s, 0, t);   /* huff_opt_pass; skipped for DC refinement (jcmaster.c:650-662) */
      if (si == 0) write_frame_header(e);
      ri = p->restart_interval;
      if (p->restart_in_rows > 0) { long nominal = (long)p->restart_in_rows * scan_mcus_per_row(e, s); ri = (unsigned)(nominal < 65535L ? nominal : 65535L); }
      write_scan_header(e, s, ri);
      memset(t, 0, sizeof *t); t->e = e; t->s = s; t->gather = 0; t->bw.o = &e->out;
      for (k = 0; k < 4; k++) { orc_make_derived(&e->dc_tbl[k], 1, t->dco[k], t->dsi[k]); orc_make_derived(&e->ac_tbl[k], 0, t->aco[k], t->asi[k]); }
      before = e->out.n;
      run_scan(t);
      if (dbg && si < B200JPEG_MAX_SCANS) { dbg->scan_bytes[si] = e->out.n - before; for (k = 0; k < 4; k++) { dbg->scan_dc[si][k] = e->dc_tbl[k]; dbg->scan_ac[si][k] = e->ac_tbl[k]; } }
    }
    if (dbg) dbg->nscans = nscans;
    bb_put2(&e->out, 0xFFD9);                                            /* write_file_trailer */
  }
  for (ci = 0; ci < e->nc; ci++) { free(e->coef[ci]); free(e->raw[ci]); }
  free(t);
  if (e->err) { free(e->out.d); return e->err; }
  *out = e->out.d; *outsize = e->out.n;
  return 0;
}

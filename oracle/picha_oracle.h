/* picha_oracle.h -- CPU restatement of picha's pixel hot path (TEST INFRASTRUCTURE ONLY;
 * see picha_oracle.c for the pinning statement and the reference citations). */
#ifndef PICHA_ORACLE_H
#define PICHA_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

/* src/picha.h:79-92 */
enum { PO_RGB = 0, PO_RGBA, PO_GREY, PO_GREYA, PO_R16, PO_R16G16, PO_R16G16B16, PO_R16G16B16A16, PO_NUM_PIXELS };
/* src/resize.cc:151-160 */
enum { PO_CUBIC = 0, PO_LANCZOS, PO_CATMULROM, PO_MITCHEL, PO_BOX, PO_TRIANGLE, PO_NUM_FILTERS };

int po_pixel_bytes(int pixel);
int po_pixel_channels(int pixel);
int po_pixel_deep(int pixel);
int po_row_stride(int width, int pixel);

/* One axis of makeContribs (src/resize.cc:19-50). Returns the number of weights. */
int po_make_contribs(int filter, float fwidth, int srcsize, int dstsize,
                     int *left, int *right, int *woff, float *weights, int cap);

/* resizeImage (src/resize.cc:66-134, :270-280). 0 on success. */
int po_resize(int filter, float fwidth,
              const unsigned char *src, int sstride, int sw, int sh,
              unsigned char *dst, int dstride, int dw, int dh, int pixel);

/* getSettings (src/colorconvert.cc:6-22): NaN means "option not given". */
void po_resolve_color_settings(double r, double g, double b, float out[3]);

/* doColorConvert (src/colorconvert.cc:171-188) with already-normalised weights. */
int po_color_convert(float r, float g, float b,
                     const unsigned char *src, int sstride, int w, int h, int spixel,
                     unsigned char *dst, int dstride, int dpixel);

/* cmyk_to_rgb (src/jpegcodec.cc:36-42) over every row of a w x h image: 4 bytes per pixel in,
 * 3 out.  The JPEG codec itself cannot be compiled here (no libjpeg headers), so this one is
 * pinned by restatement only; the formula is three integer operations. */
int po_cmyk_to_rgb(const unsigned char *cmyk, int sstride, int w, int h, unsigned char *rgb, int dstride);

#ifdef __cplusplus
}
#endif
#endif

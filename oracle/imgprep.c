/* oracle/imgprep.c -- TEST INFRASTRUCTURE (see oracle.h).  Image preparation per view.
 * Follows libs/tex/texture_view.cpp:42-132 and MVE image_tools [UPSTREAM-RECALL]. */
#include "oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* texture_view.cpp:42-94: flood fill from the four corners through 4-connected pixels whose
 * channel sum is 0; those become invalid.  Visiting order does not change the result. */
void orc_validity_mask(const uint8_t *rgb, int w, int h, uint8_t *mask)
{
    size_t n = (size_t)w * h;
    memset(mask, 1, n);
    uint8_t *checked = (uint8_t *)calloc(n, 1);
    int32_t *stack = (int32_t *)malloc(sizeof(int32_t) * (n + 4));
    size_t sp = 0;
    int cx[4] = {0, 0, w - 1, w - 1};
    int cy[4] = {0, h - 1, 0, h - 1};
    for (int i = 0; i < 4; ++i) {
        size_t id = (size_t)cx[i] + (size_t)cy[i] * w;
        if (!checked[id]) { checked[id] = 255; stack[sp++] = (int32_t)id; }
    }
    while (sp) {
        int32_t id = stack[--sp];
        int x = id % w, y = id / w;
        int sum = rgb[3 * (size_t)id] + rgb[3 * (size_t)id + 1] + rgb[3 * (size_t)id + 2];
        if (sum != 0) continue;
        mask[id] = 0;
        int nx[4] = {x + 1, x, x - 1, x};
        int ny[4] = {y, y + 1, y, y - 1};
        for (int i = 0; i < 4; ++i) {
            if (0 <= nx[i] && nx[i] < w && 0 <= ny[i] && ny[i] < h) {
                size_t nid = (size_t)nx[i] + (size_t)ny[i] * w;
                if (!checked[nid]) { checked[nid] = 255; stack[sp++] = (int32_t)nid; }
            }
        }
    }
    free(stack);
    free(checked);
}

/* texture_view.cpp:109-132.  NB the quirk: the border write at :116 hits the array that is
 * swapped away at :131, so image-border pixels are NOT invalidated; interior invalid pixels
 * invalidate their 3x3 neighbourhood. */
void orc_erode_validity_mask(uint8_t *mask, int w, int h)
{
    size_t n = (size_t)w * h;
    uint8_t *eroded = (uint8_t *)malloc(n);
    memcpy(eroded, mask, n);
    for (int y = 1; y < h - 1; ++y)
        for (int x = 1; x < w - 1; ++x) {
            if (mask[x + (size_t)y * w]) continue;
            for (int j = -1; j <= 1; ++j)
                for (int i = -1; i <= 1; ++i)
                    eroded[(x + i) + (size_t)(y + j) * w] = 0;
        }
    memcpy(mask, eroded, n);
    free(eroded);
}

/* texture_view.cpp:102-107: desaturate<uint8>(DESATURATE_LUMINANCE) then sobel_edge<uint8>.
 * [UPSTREAM-RECALL] MVE: luminance = interpolate(r,g,b; .21f,.72f,.07f) -> (u8)(sum + 0.5f);
 * Sobel 3x3 in double, image border 0, value = (u8)min(255.0, sqrt(gx^2+gy^2)). */
void orc_gradient_magnitude(const uint8_t *rgb, int w, int h, uint8_t *grad)
{
    size_t n = (size_t)w * h;
    uint8_t *bw = (uint8_t *)malloc(n);
    for (size_t i = 0; i < n; ++i) {
        float v = (float)rgb[3 * i] * 0.21f + (float)rgb[3 * i + 1] * 0.72f
            + (float)rgb[3 * i + 2] * 0.07f + 0.5f;
        bw[i] = (uint8_t)v;
    }
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            size_t p = x + (size_t)y * w;
            if (y == 0 || y == h - 1 || x == 0 || x == w - 1) { grad[p] = 0; continue; }
            double gx = 1.0 * bw[p + 1 - w] - 1.0 * bw[p - 1 - w] + 2.0 * bw[p + 1]
                - 2.0 * bw[p - 1] + 1.0 * bw[p + 1 + w] - 1.0 * bw[p - 1 + w];
            double gy = 1.0 * bw[p + w - 1] - 1.0 * bw[p - w - 1] + 2.0 * bw[p + w]
                - 2.0 * bw[p - w] + 1.0 * bw[p + w + 1] - 1.0 * bw[p - w + 1];
            double g = sqrt(gx * gx + gy * gy);
            grad[p] = (uint8_t)(g < 255.0 ? g : 255.0);
        }
    free(bw);
}

/*
 * conv_ref.c - plain C restatement of the TensorFlow ops on the DCSCN hot path.  TEST INFRASTRUCTURE ONLY
 * (see oracle/dcscn_oracle.py header): it cross-checks the torch-based oracle's reading of the TF
 * semantics with nothing but explicit loops, and is what `bench.py --impl reference` can fall back to
 * when torch's CPU convolution is unavailable.  Never linked into the product library.
 *
 * Each function cites the reference call site it restates (paths under /root/reference):
 *   conv2d_same_nhwc   tf.nn.conv2d(x, w, [1,1,1,1], "SAME")      helper/tf_graph.py:105  (+ tf.add bias :109)
 *   prelu_nhwc         relu(x) + alpha * (x - |x|) * 0.5           helper/tf_graph.py:94
 *   depth_to_space_dcr tf.depth_to_space(x, r)                     helper/tf_graph.py:248
 *   depthwise_same_nhwc first half of tf.nn.separable_conv2d      helper/tf_graph.py:161
 *
 * Build:  gcc -O2 -fopenmp -shared -fPIC -o oracle/libconv_ref.so oracle/conv_ref.c   (oracle/Makefile)
 */
#include <stddef.h>

/* x [n,h,w,cin] NHWC, w [k,k,cin,cout] HWIO, bias [cout] or NULL, y [n,h,w,cout]; accumulation in double. */
void conv2d_same_nhwc(const float* x, const float* w, const float* bias, float* y, int n, int h, int wd, int cin,
                      int cout, int k) {
  const int pad = k / 2; /* SAME, stride 1, odd k: k/2 zeros on every side */
#pragma omp parallel for collapse(2)
  for (int b = 0; b < n; ++b)
    for (int oy = 0; oy < h; ++oy)
      for (int ox = 0; ox < wd; ++ox)
        for (int co = 0; co < cout; ++co) {
          double acc = 0.0;
          for (int ky = 0; ky < k; ++ky) {
            const int iy = oy + ky - pad;
            if (iy < 0 || iy >= h) continue;
            for (int kx = 0; kx < k; ++kx) {
              const int ix = ox + kx - pad;
              if (ix < 0 || ix >= wd) continue;
              const float* xp = x + (((size_t)b * h + iy) * wd + ix) * cin;
              const float* wp = w + ((size_t)(ky * k + kx) * cin) * cout + co;
              for (int ci = 0; ci < cin; ++ci) acc += (double)xp[ci] * (double)wp[(size_t)ci * cout];
            }
          }
          if (bias) acc += (double)bias[co];
          y[(((size_t)b * h + oy) * wd + ox) * cout + co] = (float)acc;
        }
}

/* depthwise k x k, channel multiplier 1: w [k,k,c,1] */
void depthwise_same_nhwc(const float* x, const float* w, float* y, int n, int h, int wd, int c, int k) {
  const int pad = k / 2;
#pragma omp parallel for collapse(2)
  for (int b = 0; b < n; ++b)
    for (int oy = 0; oy < h; ++oy)
      for (int ox = 0; ox < wd; ++ox)
        for (int ch = 0; ch < c; ++ch) {
          double acc = 0.0;
          for (int ky = 0; ky < k; ++ky) {
            const int iy = oy + ky - pad;
            if (iy < 0 || iy >= h) continue;
            for (int kx = 0; kx < k; ++kx) {
              const int ix = ox + kx - pad;
              if (ix < 0 || ix >= wd) continue;
              acc += (double)x[(((size_t)b * h + iy) * wd + ix) * c + ch] * (double)w[(size_t)(ky * k + kx) * c + ch];
            }
          }
          y[(((size_t)b * h + oy) * wd + ox) * c + ch] = (float)acc;
        }
}

void prelu_nhwc(float* x, const float* alpha, size_t pixels, int c) {
  for (size_t p = 0; p < pixels; ++p)
    for (int ch = 0; ch < c; ++ch) {
      float v = x[p * c + ch];
      float relu = v > 0.f ? v : 0.f;
      float absv = v < 0.f ? -v : v;
      x[p * c + ch] = relu + alpha[ch] * (v - absv) * 0.5f;
    }
}

/* x [n,h,w,r*r*c] -> y [n,h*r,w*r,c], DCR: input channel (i*r + j)*c + ch goes to (oy*r+i, ox*r+j, ch) */
void depth_to_space_dcr(const float* x, float* y, int n, int h, int wd, int c, int r) {
  for (int b = 0; b < n; ++b)
    for (int oy = 0; oy < h; ++oy)
      for (int ox = 0; ox < wd; ++ox)
        for (int i = 0; i < r; ++i)
          for (int j = 0; j < r; ++j)
            for (int ch = 0; ch < c; ++ch)
              y[(((size_t)b * h * r + (oy * r + i)) * (wd * r) + (ox * r + j)) * c + ch] =
                  x[(((size_t)b * h + oy) * wd + ox) * (r * r * c) + (i * r + j) * c + ch];
}

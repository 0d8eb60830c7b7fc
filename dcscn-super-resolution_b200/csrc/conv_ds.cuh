// Depthwise-separable DCSCN layers (reference: helper/tf_graph.py:155-216 `depthwise_separable_conv2d` /
// `build_depthwise_separable_conv`, i.e. tf.nn.separable_conv2d: depthwise k x k with channel multiplier 1 and no
// bias, then pointwise 1x1, then +bias, then PReLU; used for EVERY layer of a `--depthwise_separable` graph incl. the
// 1x1 A1/B1 (a per-channel scale) and R-CNN1).
//
// These graphs are tiny (c-DCSCN: <= 131 channels, 14,240 MAC per LR pixel at x4) and bound by HBM traffic, not
// math: one fused kernel per layer on CUDA cores, fp32 NHWC activations, no tensor cores (a 131 x 24 contraction per
// pixel does not fill a UMMA tile).  The depthwise result never leaves registers / shared memory.
#pragma once
#include <cstdint>

namespace dcscn {

struct DsLayerParams {
  int n_img, H, W;          // resolution of this layer (input == output)
  int ksz;                  // depthwise kernel size (1 or 3)
  int cin, cout;
  const float* src;         // [N,H,W,src_pitch] fp32, already offset to the first input channel
  int src_pitch;
  const float* dw;          // [k*k][cin]   depthwise_W  [k,k,cin,1]
  const float* pw;          // [cin][cout]  pointwise_W  [1,1,cin,cout]
  const float* bias;        // [cout] or null
  const float* alpha;       // [cout] or null (no activation)
  // output: plain channel slot, or depth_to_space scatter (DCR), optionally + x2 (final layer)
  float* dst;
  int dst_pitch;            // channels per pixel of dst
  int dst_off;              // first channel written (plain mode)
  int d2s_r;                // 0 = plain; otherwise block size, dst is [N, r*H, r*W, dst_pitch] and cout = r*r*dst_cout
  int d2s_cout;
  const float* add;         // null, or [N,H,W] tensor added to channel 0 (cout must be 1): tf.add(H[-1], x2)
};

constexpr int kDsPix = 64;      // pixels per CTA (one row segment)
constexpr int kDsThreads = 256;

// One CTA: a run of kDsPix consecutive pixels of one image row.  Phase 1: depthwise outputs [pix][cin] into shared
// memory (threads stride over (pixel, channel): channel fastest -> coalesced NHWC reads).  Phase 2: pointwise
// [pix][cout] with the pointwise filter read through the read-only cache (warp-uniform per output channel).
__global__ void __launch_bounds__(kDsThreads) ds_layer_kernel(const DsLayerParams p) {
  extern __shared__ float s_dw[];  // [kDsPix][cin]
  const int segs_per_row = (p.W + kDsPix - 1) / kDsPix;
  const int seg = blockIdx.x % segs_per_row;
  const int rowid = blockIdx.x / segs_per_row;      // img * H + y
  const int y = rowid % p.H;
  const int img = rowid / p.H;
  const int x0 = seg * kDsPix;
  const int npix = (p.W - x0) < kDsPix ? (p.W - x0) : kDsPix;
  const int half = p.ksz >> 1;
  const float* img_base = p.src + (size_t)img * p.H * p.W * p.src_pitch;

  for (int i = threadIdx.x; i < npix * p.cin; i += blockDim.x) {
    const int c = i % p.cin, px = i / p.cin;
    const int x = x0 + px;
    float acc = 0.f;
    for (int t = 0; t < p.ksz * p.ksz; ++t) {
      const int yy = y + t / p.ksz - half, xx = x + t % p.ksz - half;
      if (yy >= 0 && yy < p.H && xx >= 0 && xx < p.W)
        acc = fmaf(__ldg(img_base + ((size_t)yy * p.W + xx) * p.src_pitch + c), __ldg(p.dw + t * p.cin + c), acc);
    }
    s_dw[px * p.cin + c] = acc;
  }
  __syncthreads();

  for (int i = threadIdx.x; i < npix * p.cout; i += blockDim.x) {
    const int co = i % p.cout, px = i / p.cout;
    const float* d = s_dw + px * p.cin;
    float acc = 0.f;
    for (int c = 0; c < p.cin; ++c) acc = fmaf(d[c], __ldg(p.pw + (size_t)c * p.cout + co), acc);
    if (p.bias) acc += __ldg(p.bias + co);
    if (p.alpha) acc = acc > 0.f ? acc : __ldg(p.alpha + co) * acc;
    const int x = x0 + px;
    if (p.d2s_r == 0) {
      const size_t o = (((size_t)img * p.H + y) * p.W + x);
      if (p.add) acc += __ldg(p.add + o);
      p.dst[o * p.dst_pitch + p.dst_off + co] = acc;
    } else {
      // DCR: input channel (i*r + j)*C + c -> (y*r + i, x*r + j, c)   (tf.depth_to_space, tf_graph.py:248)
      const int r = p.d2s_r, ij = co / p.d2s_cout, c = co - ij * p.d2s_cout;
      const int ii = ij / r, jj = ij - ii * r;
      const size_t o = (((size_t)img * p.H * r + (size_t)(y * r + ii)) * (p.W * r) + (size_t)(x * r + jj));
      p.dst[o * p.dst_pitch + c] = acc;
    }
  }
}

}  // namespace dcscn

// ff_pre.cuh — K0: crop preprocessing.  cv2.resize(face,(224,224),INTER_AREA) + cv2.cvtColor(RGB2BGR)
// (/root/reference/CViT-main/cvit_prediction.py:114-115, :96-97, :141-142) for a batch of variable-size uint8
// HWC crops, optionally also emitting the normalised fp32 NCHW tensor of cvit_prediction.py:209-215.
//
// OpenCV's INTER_AREA has three regimes (modules/imgproc/src/resize.cpp); each is restated so that the uint8
// result matches OpenCV's (see oracle/resize_oracle.py for the CPU restatement it is tested against):
//   FAST   both scales integer >= 1 : block sum * (1/area) in fp32, round-half-even; 2x2 -> (s+2)>>2
//   FRAC   both scales >= 1         : computeResizeAreaTab weights (fp64 -> fp32), fp32 accumulate in table order
//   LINEAR any up-scaling           : bilinear in "area mode", 11-bit fixed-point taps, integer arithmetic
#pragma once
#include <cmath>
#include "ff_ptx.cuh"

namespace ff {

enum { PRE_FAST = 0, PRE_FRAC = 1, PRE_LINEAR = 2 };

struct CropDesc {
  const uint8_t* ptr;
  int h, w, pitch;
  int mode, isx, isy;
};

struct AreaSpan {
  int n;        // number of table entries for this destination index
  int sx1, sx2;
  double fs1, fs2, cell;
  bool head, tail;
};

__device__ __forceinline__ AreaSpan area_span(int d, double scale, int ssize) {
  AreaSpan s;
  s.fs1 = __dmul_rn(static_cast<double>(d), scale);      // explicit roundings: an fma contraction of d*scale + scale
  s.fs2 = __dadd_rn(s.fs1, scale);                         // would move a table boundary by one ulp
  s.cell = fmin(scale, ssize - s.fs1);
  int sx1 = static_cast<int>(ceil(s.fs1)), sx2 = static_cast<int>(floor(s.fs2));
  sx2 = min(sx2, ssize - 1);
  sx1 = min(sx1, sx2);
  s.sx1 = sx1;
  s.sx2 = sx2;
  s.head = (sx1 - s.fs1) > 1e-3;
  s.tail = (s.fs2 - sx2) > 1e-3;
  s.n = (s.head ? 1 : 0) + (sx2 - sx1) + (s.tail ? 1 : 0);
  return s;
}
// entry e of the span: source index and fp32 weight, in OpenCV's table order
__device__ __forceinline__ void area_entry(const AreaSpan& s, int e, int* si, float* alpha) {
  if (s.head && e == 0) {
    *si = s.sx1 - 1;
    *alpha = static_cast<float>((s.sx1 - s.fs1) / s.cell);
    return;
  }
  const int m = e - (s.head ? 1 : 0);
  if (m < s.sx2 - s.sx1) {
    *si = s.sx1 + m;
    *alpha = static_cast<float>(1.0 / s.cell);
    return;
  }
  *si = s.sx2;
  *alpha = static_cast<float>(fmin(fmin(s.fs2 - s.sx2, 1.0), s.cell) / s.cell);
}

__device__ __forceinline__ void linear_tap(int d, int ssize, int dsize, int* sx, int* a0, int* a1, bool* edge) {
  const double scale = static_cast<double>(ssize) / dsize, inv = static_cast<double>(dsize) / ssize;
  int s = static_cast<int>(floor(__dmul_rn(static_cast<double>(d), scale)));
  float fx = static_cast<float>(__dsub_rn(static_cast<double>(d + 1), __dmul_rn(static_cast<double>(s + 1), inv)));
  fx = fx <= 0.0f ? 0.0f : fx - floorf(fx);
  if (s < 0) { fx = 0.0f; s = 0; }
  bool e = false;
  if (s + 1 >= ssize) {
    e = true;
    if (s >= ssize - 1) { fx = 0.0f; s = ssize - 1; }
  }
  const float c0 = 1.0f - fx;
  *a0 = max(-32768, min(32767, __float2int_rn(c0 * 2048.0f)));
  *a1 = max(-32768, min(32767, __float2int_rn(fx * 2048.0f)));
  *sx = s;
  *edge = e;
}

__device__ __forceinline__ uint8_t sat_u8_rn(float v) { return static_cast<uint8_t>(max(0, min(255, __float2int_rn(v)))); }

// Per-destination-index tables (what OpenCV keeps in xtab/ytab / xofs+ialpha): computed once per block in fp64
// exactly as before, then shared by the block's 896 output pixels.
struct AreaTabEntry {          // FRAC: entries of one destination index, in OpenCV's table order
  int first;                   // source index of entry 0
  int n;                       // number of entries
  float a_first, a_mid, a_last;
};
struct LinTabEntry { int s, a0, a1, edge; };

__device__ __forceinline__ AreaTabEntry make_area_tab(int d, double scale, int ssize) {
  const AreaSpan sp = area_span(d, scale, ssize);
  AreaTabEntry t;
  t.n = sp.n;
  t.first = sp.head ? sp.sx1 - 1 : sp.sx1;
  int si;
  float al = 0.f;
  t.a_first = t.a_mid = t.a_last = 0.f;
  if (sp.n > 0) { area_entry(sp, 0, &si, &al); t.a_first = al; }
  if (sp.n > 1) { area_entry(sp, sp.n - 1, &si, &al); t.a_last = al; }
  t.a_mid = static_cast<float>(1.0 / sp.cell);
  // entry e (0 < e < n-1) is always a full source pixel with weight 1/cell; entry 0 is the head (or a full pixel,
  // for which a_first == a_mid) and entry n-1 the tail (or a full pixel)
  return t;
}
__device__ __forceinline__ float area_alpha(const AreaTabEntry& t, int e) {
  return e == 0 ? t.a_first : (e == t.n - 1 ? t.a_last : t.a_mid);
}

// Destination size D x D: 224 for the CViT crops, 128 for the BlazeFace tiles (helpers_face_extract_1.py:195).
// Block = 4 output rows of one crop; one thread produces 4 consecutive output pixels of one row
// (12 output bytes = three aligned 32-bit stores), so a warp covers 128 consecutive output pixels.
template <int D>
__global__ void __launch_bounds__(256)
preprocess_kernel(const CropDesc* __restrict__ crops, int n, int swap_rb, uint8_t* __restrict__ out_u8,
                  float* __restrict__ out_norm) {
  constexpr int ROWS = 4;
  static_assert(D % 4 == 0 && ROWS * D / 4 <= 256 && D + ROWS <= 256, "block shape");
  __shared__ AreaTabEntry s_ax[D];
  __shared__ AreaTabEntry s_ay[ROWS];
  __shared__ LinTabEntry s_lx[D];
  __shared__ LinTabEntry s_ly[ROWS];
  const int i = blockIdx.y;
  if (i >= n) return;
  const CropDesc c = crops[i];
  const int row0 = blockIdx.x * ROWS;
  const int tid = threadIdx.x;
  if (c.mode == PRE_FRAC) {
    if (tid < D) s_ax[tid] = make_area_tab(tid, static_cast<double>(c.w) / D, c.w);
    else if (tid < D + ROWS) s_ay[tid - D] = make_area_tab(row0 + tid - D, static_cast<double>(c.h) / D, c.h);
  } else if (c.mode == PRE_LINEAR) {
    if (tid < D + ROWS) {
      const bool isx = tid < D;
      LinTabEntry e;
      bool edge;
      linear_tap(isx ? tid : row0 + tid - D, isx ? c.w : c.h, D, &e.s, &e.a0, &e.a1, &edge);
      e.edge = edge ? 1 : 0;
      if (isx) s_lx[tid] = e; else s_ly[tid - D] = e;
    }
  }
  __syncthreads();
  if (tid >= ROWS * D / 4) return;
  const int rl = tid / (D / 4);
  const int dy = row0 + rl, dx0 = (tid - rl * (D / 4)) * 4;
  const int pix0 = dy * D + dx0;
  uint32_t packed[3] = {0u, 0u, 0u};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int dx = dx0 + j;
    int res[3];
    if (c.mode == PRE_FAST) {
      int sum[3] = {0, 0, 0};
      for (int sy = 0; sy < c.isy; ++sy) {
        const uint8_t* row = c.ptr + static_cast<size_t>(dy * c.isy + sy) * c.pitch + static_cast<size_t>(dx) * c.isx * 3;
        for (int sx = 0; sx < c.isx; ++sx) {
          sum[0] += row[sx * 3]; sum[1] += row[sx * 3 + 1]; sum[2] += row[sx * 3 + 2];
        }
      }
      if (c.isx == 2 && c.isy == 2) {
        for (int k = 0; k < 3; ++k) res[k] = (sum[k] + 2) >> 2;
      } else {
        const float scale = 1.0f / static_cast<float>(c.isx * c.isy);
        for (int k = 0; k < 3; ++k) res[k] = sat_u8_rn(__fmul_rn(static_cast<float>(sum[k]), scale));
      }
    } else if (c.mode == PRE_FRAC) {
      const AreaTabEntry tx = s_ax[dx], ty = s_ay[rl];
      float sum[3] = {0.f, 0.f, 0.f};
      for (int ey = 0; ey < ty.n; ++ey) {
        const float beta = area_alpha(ty, ey);
        const uint8_t* p = c.ptr + static_cast<size_t>(ty.first + ey) * c.pitch + static_cast<size_t>(tx.first) * 3;
        float buf[3] = {0.f, 0.f, 0.f};
        for (int ex = 0; ex < tx.n; ++ex, p += 3) {
          const float alpha = area_alpha(tx, ex);
          for (int k = 0; k < 3; ++k) buf[k] = __fadd_rn(buf[k], __fmul_rn(static_cast<float>(p[k]), alpha));
        }
        for (int k = 0; k < 3; ++k)
          sum[k] = (ey == 0) ? __fmul_rn(beta, buf[k]) : __fadd_rn(sum[k], __fmul_rn(beta, buf[k]));
      }
      for (int k = 0; k < 3; ++k) res[k] = sat_u8_rn(sum[k]);
    } else {
      const LinTabEntry ex = s_lx[dx], ey = s_ly[rl];
      const int sx1 = min(ex.s + 1, c.w - 1), sy1 = min(ey.s + 1, c.h - 1);
      const uint8_t* r0 = c.ptr + static_cast<size_t>(ey.s) * c.pitch;
      const uint8_t* r1 = c.ptr + static_cast<size_t>(sy1) * c.pitch;
      for (int k = 0; k < 3; ++k) {
        int h0, h1;
        if (ex.edge) { h0 = r0[ex.s * 3 + k] * 2048; h1 = r1[ex.s * 3 + k] * 2048; }
        else {
          h0 = r0[ex.s * 3 + k] * ex.a0 + r0[sx1 * 3 + k] * ex.a1;
          h1 = r1[ex.s * 3 + k] * ex.a0 + r1[sx1 * 3 + k] * ex.a1;
        }
        const int v = (((ey.a0 * (h0 >> 4)) >> 16) + ((ey.a1 * (h1 >> 4)) >> 16) + 2) >> 2;
        res[k] = max(0, min(255, v));
      }
    }
    if (swap_rb) { const int t = res[0]; res[0] = res[2]; res[2] = t; }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int byte = j * 3 + k;
      packed[byte >> 2] |= static_cast<uint32_t>(res[k]) << (8 * (byte & 3));
    }
    if (out_norm) {
      const float mean[3] = {0.485f, 0.456f, 0.406f}, sd[3] = {0.229f, 0.224f, 0.225f};
      for (int k = 0; k < 3; ++k)
        out_norm[(static_cast<size_t>(i) * 3 + k) * D * D + pix0 + j] =
            __fdiv_rn(__fdiv_rn(static_cast<float>(res[k]), 255.0f) - mean[k], sd[k]);
    }
  }
  if (out_u8) {
    uint32_t* o = reinterpret_cast<uint32_t*>(out_u8 + (static_cast<size_t>(i) * D * D + pix0) * 3);   // 12-byte aligned
    o[0] = packed[0]; o[1] = packed[1]; o[2] = packed[2];
  }
}

// Which of OpenCV's three INTER_AREA regimes a (h x w) -> (D x D) resize falls into (resize.cpp: integer scale ->
// resizeAreaFast_, fractional down-scale -> resizeArea_, any up-scaling -> the bilinear kernel in area mode).
inline CropDesc make_crop_desc(const uint8_t* ptr, int h, int w, int pitch, int D) {
  CropDesc d;
  d.ptr = ptr; d.h = h; d.w = w; d.pitch = pitch;
  const double sx = static_cast<double>(w) / D, sy = static_cast<double>(h) / D;
  if (sx >= 1.0 && sy >= 1.0) {
    const int isx = static_cast<int>(lrint(sx)), isy = static_cast<int>(lrint(sy));
    const bool fast = fabs(sx - isx) < 2.220446049250313e-16 && fabs(sy - isy) < 2.220446049250313e-16;
    d.mode = fast ? PRE_FAST : PRE_FRAC; d.isx = isx; d.isy = isy;
  } else {
    d.mode = PRE_LINEAR; d.isx = d.isy = 1;
  }
  return d;
}

}  // namespace ff

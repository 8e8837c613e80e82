// Scene-graph constraint metrics: validate_constrains / validate_constrains_changes of the reference's
// helpers/metrics_3dfront.py:57-306 -- the last CPU stage of scripts/eval_3dfront.py (:209-210, :305), there a Python loop over the
// triples with two device->host copies per triple.  Here: one thread per triple, the boxes never leave the GPU.
//
// The reference's arithmetic is reproduced type for type: box entries are float32 (numpy scalars of a float32 array), differences,
// volumes and heights are float32 operations, thresholds compare in float32; corners_from_box / box3d_iou / close_dis run in
// float64 on float32-rounded inputs (np.dot with a float64 identity promotes).  Boxes are axis-aligned there (corners_from_box
// ignores the angle), the bird's-eye overlap is a literal Sutherland-Hodgman clip (polygon_clip :390-434) followed by the area
// of the clipped polygon (ConvexHull(...).volume of a convex polygon = its shoelace area).
#include "ops.cuh"

#include <math.h>

namespace echo {
namespace {

enum Rel : int { R_LEFT = 0, R_RIGHT, R_FRONT, R_BEHIND, R_BIGGER, R_SMALLER, R_TALLER, R_SHORTER, R_STANDING, R_CLOSE, R_SYMM, R_COUNT };

struct Box {
  float l, h, w, px, py, pz;   // [l, h, w, px, py, pz(, angle)]: l along z, h along y, w along x; (px, py, pz) = bottom centre
};

__device__ inline Box load_box(const float* __restrict__ boxes, int64_t i, int D) {
  const float* b = boxes + i * D;
  return {b[0], b[1], b[2], b[3], b[4], b[5]};
}

// corners_from_box(with_translation=True): x = +-w/2 + px, y = {h, 0} + py, z = +-l/2 + pz; corner order as the reference's lists
__device__ inline void corners(const Box& b, double (&c)[8][3]) {
  const float hw = b.w / 2.f, hl = b.l / 2.f;   // float32 halves
  const double xs[8] = {hw, hw, -hw, -hw, hw, hw, -hw, -hw};
  const double ys[8] = {b.h, b.h, b.h, b.h, 0., 0., 0., 0.};
  const double zs[8] = {hl, -hl, -hl, hl, hl, -hl, -hl, hl};
  for (int i = 0; i < 8; ++i) {
    c[i][0] = xs[i] + (double)b.px;
    c[i][1] = ys[i] + (double)b.py;
    c[i][2] = zs[i] + (double)b.pz;
  }
}

// polygon_clip(subject, clip) (Sutherland-Hodgman, strict `inside`) + area of the result; 0 when the result is empty
__device__ double clip_area(const double (&subj)[4][2], const double (&clip)[4][2]) {
  double out[16][2], in[16][2];
  int n_out = 4;
  for (int i = 0; i < 4; ++i) { out[i][0] = subj[i][0]; out[i][1] = subj[i][1]; }
  double cp1x = clip[3][0], cp1y = clip[3][1];
  for (int ci = 0; ci < 4; ++ci) {
    const double cp2x = clip[ci][0], cp2y = clip[ci][1];
    const int n_in = n_out;
    for (int i = 0; i < n_in; ++i) { in[i][0] = out[i][0]; in[i][1] = out[i][1]; }
    n_out = 0;
    double sx = in[n_in - 1][0], sy = in[n_in - 1][1];
    for (int i = 0; i < n_in; ++i) {
      const double ex = in[i][0], ey = in[i][1];
      const bool e_in = (cp2x - cp1x) * (ey - cp1y) > (cp2y - cp1y) * (ex - cp1x);
      const bool s_in = (cp2x - cp1x) * (sy - cp1y) > (cp2y - cp1y) * (sx - cp1x);
      if (e_in != s_in) {   // computeIntersection()
        const double dcx = cp1x - cp2x, dcy = cp1y - cp2y, dpx = sx - ex, dpy = sy - ey;
        const double n1 = cp1x * cp2y - cp1y * cp2x, n2 = sx * ey - sy * ex;
        const double n3 = 1.0 / (dcx * dpy - dcy * dpx);
        if (n_out < 16) { out[n_out][0] = (n1 * dpx - n2 * dcx) * n3; out[n_out][1] = (n1 * dpy - n2 * dcy) * n3; ++n_out; }
      }
      if (e_in && n_out < 16) { out[n_out][0] = ex; out[n_out][1] = ey; ++n_out; }
      sx = ex; sy = ey;
    }
    cp1x = cp2x; cp1y = cp2y;
    if (n_out == 0) return 0.0;
  }
  double a = 0.0;   // shoelace
  for (int i = 0; i < n_out; ++i) {
    const int j = (i + 1) % n_out;
    a += out[i][0] * out[j][1] - out[j][0] * out[i][1];
  }
  return 0.5 * fabs(a);
}

// box3d_iou(box1, box2, with_translation=True)[0]: intersection volume over the SMALLER box's volume
__device__ double iou3d(const Box& b1, const Box& b2) {
  double c1[8][3], c2[8][3];
  corners(b1, c1);
  corners(b2, c2);
  double r1[4][2], r2[4][2];
  for (int i = 0; i < 4; ++i) { r1[i][0] = c1[i][2]; r1[i][1] = c1[i][0]; r2[i][0] = c2[i][2]; r2[i][1] = c2[i][0]; }
  const double inter_area = clip_area(r1, r2);
  const double ymax = fmin(c1[0][1], c2[0][1]), ymin = fmax(c1[4][1], c2[4][1]);
  const double inter_vol = inter_area * fmax(0.0, ymax - ymin);
  auto vol = [](const double (&c)[8][3]) {
    auto d = [&](int i, int j) {
      const double x = c[i][0] - c[j][0], y = c[i][1] - c[j][1], z = c[i][2] - c[j][2];
      return sqrt(x * x + y * y + z * z);
    };
    return d(0, 1) * d(1, 2) * d(0, 4);
  };
  return inter_vol / fmin(vol(c1), vol(c2));
}

// close_dis: min over the 8 x 8 corner pairs of sqrt(|a|^2 + |b|^2 - 2 a.b); a NaN entry makes np.min NaN (and the test pass)
__device__ double close_dis(const Box& b1, const Box& b2) {
  double c1[8][3], c2[8][3];
  corners(b1, c1);
  corners(b2, c2);
  double best = INFINITY;
  bool nan = false;
  for (int i = 0; i < 8; ++i)
    for (int j = 0; j < 8; ++j) {
      double d = -2.0 * (c1[i][0] * c2[j][0] + c1[i][1] * c2[j][1] + c1[i][2] * c2[j][2]);
      d += c1[i][0] * c1[i][0] + c1[i][1] * c1[i][1] + c1[i][2] * c1[i][2];
      d += c2[j][0] * c2[j][0] + c2[j][1] * c2[j][1] + c2[j][2] * c2[j][2];
      d = sqrt(d);
      if (d != d) nan = true;
      best = fmin(best, d);
    }
  return nan ? NAN : best;
}

__device__ inline float l2(float ax, float ay, float bx, float by) {   // cal_l2_distance on float32 scalars
  const float dx = bx - ax, dy = by - ay;
  return sqrtf(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
}

__global__ void validate_kernel(const int64_t* __restrict__ triples, int64_t T, const float* __restrict__ boxes, int64_t N, int D,
                                const int32_t* __restrict__ keep, int changes_mode, const int32_t* __restrict__ rel_of_pred, int n_preds,
                                int strict, float overlap_threshold, int8_t* __restrict__ out_rel, int8_t* __restrict__ out_ok) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= T) return;
  int8_t rel = -1, ok = 0;
  const int64_t s = triples[3 * i], p = triples[3 * i + 1], o = triples[3 * i + 2];
  bool take = s >= 0 && s < N && o >= 0 && o < N && p >= 0 && p < n_preds;
  if (take && keep) take = changes_mode ? (keep[s] == 0 || keep[o] == 0) : (keep[s] == 1 && keep[o] == 1);
  const int r = take ? rel_of_pred[p] : -1;
  if (r >= 0) {
    const Box a = load_box(boxes, s, D), b = load_box(boxes, o, D);
    const float thr = overlap_threshold;
    bool good = false;
    switch (r) {
      case R_LEFT:    good = !(__fsub_rn(a.pz, b.pz) > -0.05f || (strict && iou3d(a, b) > (double)thr)); break;
      case R_RIGHT:   good = !(__fsub_rn(a.pz, b.pz) < 0.05f || (strict && iou3d(a, b) > (double)thr)); break;
      case R_FRONT:   good = !(__fsub_rn(a.px, b.px) < -0.05f || (strict && iou3d(a, b) > (double)thr)); break;
      case R_BEHIND:  good = !(__fsub_rn(a.px, b.px) > 0.05f || (strict && iou3d(a, b) > (double)thr)); break;
      case R_BIGGER:
      case R_SMALLER: {
        const float sv = __fmul_rn(__fmul_rn(a.l, a.h), a.w), ov = __fmul_rn(__fmul_rn(b.l, b.h), b.w);
        const float q = __fdiv_rn(__fsub_rn(sv, ov), sv);
        good = r == R_BIGGER ? !(q < 0.15f) : !(q > -0.15f);
        break;
      }
      case R_TALLER:
      case R_SHORTER: {
        const float hs = __fadd_rn(a.py, a.h), ho = __fadd_rn(b.py, b.h);
        const float q = __fdiv_rn(__fsub_rn(hs, ho), hs);
        good = r == R_TALLER ? !(q < 0.1f) : !(q > -0.1f);
        break;
      }
      case R_STANDING: good = fabsf(__fsub_rn(a.py, b.py)) < 0.04f; break;
      case R_CLOSE: {
        const double d = close_dis(a, b);
        good = !(d > 0.45);
        break;
      }
      case R_SYMM:
        good = l2(-a.px, -a.pz, b.px, b.pz) < 0.45f || l2(-a.px, a.pz, b.px, b.pz) < 0.45f || l2(a.px, -a.pz, b.px, b.pz) < 0.45f;
        break;
      default: break;
    }
    rel = (int8_t)r;
    ok = good ? 1 : 0;
  }
  out_rel[i] = rel;
  out_ok[i] = ok;
}

}  // namespace

void validate_constraints(const int64_t* triples, int64_t T, const float* boxes, int64_t N, int D, const int32_t* keep, bool changes_mode,
                          const int32_t* rel_of_pred_dev, int n_preds, bool strict, float overlap_threshold, int8_t* out_rel, int8_t* out_ok,
                          cudaStream_t s) {
  if (T == 0) return;
  validate_kernel<<<cdiv(T, 128), 128, 0, s>>>(triples, T, boxes, N, D, keep, changes_mode ? 1 : 0, rel_of_pred_dev, n_preds, strict ? 1 : 0,
                                                overlap_threshold, out_rel, out_ok);
  ECHO_LAUNCH_CHECK();
}

}  // namespace echo
